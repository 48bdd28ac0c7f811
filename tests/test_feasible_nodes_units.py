"""actions/common/feasible_nodes_test.go TestFeasibleNodes on the oracle (CPU): the whole-GPU cases over the CPU node, the
node with an idle GPU and the node with a releasing GPU (:20-33); the fraction / GPU-memory / MIG cases (:190-250) need
the GPU-sharing tables that are out of scope."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dsl  # noqa: E402
from oracle_lib import Oracle, lib  # noqa: E402

from kai_scheduler_b200 import abi  # noqa: E402

NODES = {"cpu-node": {"GPUs": 0}, "idle-gpu-node": {"GPUs": 1}, "releasing-gpu-node": {"GPUs": 1}, "full-gpu-node": {"GPUs": 1}}
GPU_NODES = ["idle-gpu-node", "releasing-gpu-node"]
HOLDERS = [{"Name": "leaving", "QueueName": "q", "RequiredGPUsPerTask": 1, "Tasks": [{"State": "Releasing", "NodeName": "releasing-gpu-node"}]},
           {"Name": "staying", "QueueName": "q", "RequiredGPUsPerTask": 1, "Tasks": [{"State": "Running", "NodeName": "full-gpu-node"}]}]


def _feasible(job, nodes=NODES):
    topo = {"Nodes": nodes, "Queues": [{"Name": "q", "DeservedGPUs": 4}],
            "Jobs": [job] + [h for h in HOLDERS if h["Tasks"][0]["NodeName"] in nodes]}
    snap, meta = dsl.build_snapshot(topo)
    o = Oracle(abi.make_config())
    o.load(snap)
    l = lib()
    l.kai_oracle_feasible_nodes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
    out = np.zeros(max(1, snap.n_nodes), dtype=np.int32)
    assert l.kai_oracle_feasible_nodes(o._h, meta["job_names"].index(job["Name"]), out.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    return sorted(n for i, n in enumerate(meta["node_names"]) if out[i])


def _job(*tasks):
    """tasks: GPUs per pod (0 = a CPU / memory pod)."""
    job = {"Name": "job", "QueueName": "q", "Priority": 100, "Tasks": [{"State": "Pending"} for _ in tasks]}
    if len(set(tasks)) == 1:
        job["RequiredGPUsPerTask"] = tasks[0]
        if tasks[0] == 0:
            job["RequiredCPUsPerTask"] = 1.0
    return job, tasks


def test_no_nodes():  # :89-104
    job, _ = _job(1)
    assert _feasible(job, nodes={}) == []


def test_cpu_only_job_keeps_every_node():  # :106-121
    job, _ = _job(0)
    assert _feasible(job) == sorted(NODES)


@pytest.mark.parametrize("pods", [(2,), (2, 2)])  # :123-165 whole GPU job, distributed whole GPU job
def test_whole_gpu_job_keeps_nodes_with_idle_or_releasing_gpus(pods):
    job, _ = _job(*pods)
    assert _feasible(job) == GPU_NODES  # neither the CPU node nor the node whose only GPU is taken by a running pod


def test_mixed_requests_keep_every_node():  # :167-188 one whole-GPU pod and one CPU pod
    topo_job = {"Name": "job", "QueueName": "q", "Priority": 100, "RequiredGPUsPerTask": 2,
                "Tasks": [{"State": "Pending"}, {"State": "Pending", "RequiredGPUs": 0}]}
    snap, meta = dsl.build_snapshot({"Nodes": NODES, "Queues": [{"Name": "q", "DeservedGPUs": 4}], "Jobs": [topo_job] + HOLDERS})
    # the table format has one GPU count per job: clear the second pod's GPU request in the packed snapshot
    t = meta["task_names"].index("job-1")
    snap.task_req = snap.task_req.copy()
    snap.task_req[t, 2] = 0.0
    snap.task_req[t, 0] = 1000.0
    o = Oracle(abi.make_config())
    o.load(snap)
    l = lib()
    l.kai_oracle_feasible_nodes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
    out = np.zeros(snap.n_nodes, dtype=np.int32)
    assert l.kai_oracle_feasible_nodes(o._h, meta["job_names"].index("job"), out.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    assert sorted(n for i, n in enumerate(meta["node_names"]) if out[i]) == sorted(NODES)
