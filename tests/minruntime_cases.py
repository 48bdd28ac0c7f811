"""Action-level min-runtime scenarios shared by the oracle tests (CPU) and the engine parity tests (GPU).

The reference has no action tables for plugins/minruntime (its tests stop at the plugin's functions, see
tests/test_minruntime.py); these clusters are built with the reference's test DSL and the protection inputs are set on
the packed snapshot.  Each case: (name, topology, action, config kwargs, queue min-runtimes, job last starts (seconds
before now), expected {task: (node, status)} reasoned from minruntime.go).
"""
import numpy as np

import dsl
from kai_scheduler_b200 import abi

NOW = 10_000.0


def _running(name, queue, node, n=1, gpus=1, min_available=None, priority=50):
    job = {"Name": name, "QueueName": queue, "Priority": priority, "RequiredGPUsPerTask": gpus,
           "Tasks": [{"State": "Running", "NodeName": node} for _ in range(n)]}
    if min_available is not None:
        job["RootSubGroupSet"] = {"podsets": [{"name": dsl.DEFAULT_SUBGROUP, "min_available": min_available}]}
    return job


def _pending(name, queue, n=1, gpus=1, priority=50):
    return {"Name": name, "QueueName": queue, "Priority": priority, "RequiredGPUsPerTask": gpus,
            "Tasks": [{"State": "Pending"} for _ in range(n)]}


def _reclaim_topology(victims, reclaimer_gpus=1, node_gpus=2):
    return {"Nodes": {"node0": {"GPUs": node_gpus}},
            "Queues": [{"Name": "q0", "DeservedGPUs": 0, "GPUOverQuotaWeight": 0},
                       {"Name": "q1", "DeservedGPUs": node_gpus, "GPUOverQuotaWeight": 1}],
            "Jobs": victims + [_pending("reclaimer", "q1", gpus=reclaimer_gpus)]}


def _preempt_topology(victims, preemptor_gpus=1, node_gpus=2):
    return {"Nodes": {"node0": {"GPUs": node_gpus}},
            "Queues": [{"Name": "q0", "DeservedGPUs": node_gpus, "GPUOverQuotaWeight": 1}],
            "Jobs": victims + [_pending("preemptor", "q0", gpus=preemptor_gpus, priority=150)]}


R, P = "Releasing", "Pipelined"
CASES = [
    # ---- reclaim: two 1-GPU victims of an over-quota queue, the reclaimer needs one GPU ----
    ("reclaim-unprotected", _reclaim_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "reclaim", {}, {}, {}, "one-victim"),
    ("reclaim-first-choice-protected", _reclaim_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "reclaim", {"default_reclaim_min_runtime_s": 60}, {}, {"@first": 10, "@other": 600}, "other-victim"),
    ("reclaim-both-protected", _reclaim_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "reclaim", {"default_reclaim_min_runtime_s": 60}, {}, {"v-old": 10, "v-new": 20}, "none"),
    ("reclaim-queue-value-overrides-default", _reclaim_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "reclaim", {"default_reclaim_min_runtime_s": 60}, {"q0": (None, 5)}, {"v-old": 10, "v-new": 20}, "one-victim"),
    ("reclaim-window-over", _reclaim_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "reclaim", {"default_reclaim_min_runtime_s": 60}, {}, {"v-old": 61, "v-new": 60}, "one-victim"),
    # ---- reclaim from a protected ELASTIC job (3 pods, minAvailable 1) on a 3-GPU node ----
    ("reclaim-elastic-surplus", _reclaim_topology([_running("el", "q0", "node0", n=3, min_available=1)], reclaimer_gpus=2, node_gpus=3),
     "reclaim", {"default_reclaim_min_runtime_s": 60}, {}, {"el": 10}, "elastic-two"),
    ("reclaim-elastic-below-min", _reclaim_topology([_running("el", "q0", "node0", n=3, min_available=1)], reclaimer_gpus=3, node_gpus=3),
     "reclaim", {"default_reclaim_min_runtime_s": 60}, {}, {"el": 10}, "none"),
    ("reclaim-elastic-unprotected", _reclaim_topology([_running("el", "q0", "node0", n=3, min_available=1)], reclaimer_gpus=3, node_gpus=3),
     "reclaim", {"default_reclaim_min_runtime_s": 60}, {}, {"el": 100}, "elastic-all"),
    # ---- preempt: same queue, lower priority victims ----
    ("preempt-unprotected", _preempt_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "preempt", {}, {}, {}, "one-victim"),
    ("preempt-both-protected", _preempt_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "preempt", {"default_preempt_min_runtime_s": 30}, {}, {"v-old": 10, "v-new": 29}, "none"),
    ("preempt-queue-value", _preempt_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "preempt", {"default_preempt_min_runtime_s": 5}, {"default": (30, None)}, {"v-old": 10, "v-new": 29}, "none"),
    ("preempt-reclaim-value-is-irrelevant", _preempt_topology([_running("v-old", "q0", "node0"), _running("v-new", "q0", "node0")]),
     "preempt", {"default_reclaim_min_runtime_s": 600}, {"q0": (None, 600)}, {"v-old": 10, "v-new": 29}, "one-victim"),
    ("preempt-elastic-below-min", _preempt_topology([_running("el", "q0", "node0", n=2, min_available=1)], preemptor_gpus=2),
     "preempt", {"default_preempt_min_runtime_s": 30}, {}, {"el": 10}, "none"),
    ("preempt-elastic-surplus", _preempt_topology([_running("el", "q0", "node0", n=2, min_available=1)], preemptor_gpus=1),
     "preempt", {"default_preempt_min_runtime_s": 30}, {}, {"el": 10}, "elastic-one"),
]


def build(case, first_choice=None):
    """-> (snapshot, meta, config).  `@first` / `@other` in the last-start table name the victim the unprotected run
    picks first (passed in by the caller) and the other one."""
    name, topo, action, cfg, queue_mrt, last_start, _ = case
    snap, meta = dsl.build_snapshot(topo)
    Q, J = snap.n_queues, snap.n_jobs
    pre, rec = np.full(Q, -1.0), np.full(Q, -1.0)
    for q, (p, r) in queue_mrt.items():
        i = meta["queue_names"].index(q)
        pre[i] = -1.0 if p is None else p
        rec[i] = -1.0 if r is None else r
    ls = np.full(J, -1.0)
    for job, ago in last_start.items():
        if job == "@first":
            job = first_choice
        elif job == "@other":
            job = "v-new" if first_choice == "v-old" else "v-old"
        ls[meta["job_names"].index(job)] = NOW - ago
    snap.now_s = NOW
    snap.queue_preempt_min_runtime_s, snap.queue_reclaim_min_runtime_s, snap.job_last_start_s = pre, rec, ls
    return snap, meta, abi.make_config(**cfg)


def outcome(res, meta):
    status_names = {v: k for k, v in abi.POD_STATUS_NAMES.items()}
    return {n: status_names[int(res.task_status[t])] for t, n in enumerate(meta["task_names"])}
