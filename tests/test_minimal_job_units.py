"""MinimalJobRepresentatives (actions/common/minimal_job_comparison.go) restated in the oracle, pinned on
minimal_job_comparison_test.go:36-380 (CPU): IsEasierToSchedule for single pods, several pods and mixed pod statuses,
and UpdateRepresentative.  Pod requests are (cpu milli, memory bytes, gpus); the two fractional-GPU cases of the Go
table are out of scope (shared GPUs are refused at load).
"""
import ctypes as C

import numpy as np
import pytest

from kai_scheduler_b200 import abi
from oracle_lib import Oracle, lib

Mi = 2 ** 20
P, F, RUN = abi.POD_PENDING, abi.POD_FAILED, abi.POD_RUNNING


def two_jobs(rep_pods, job_pods):
    """pods: [(cpu, mem, gpu)] or [((cpu, mem, gpu), status)]; job 0 = representative, job 1 = candidate."""
    def norm(pods):
        return [p if isinstance(p[0], tuple) else (p, P) for p in pods]

    rep_pods, job_pods = norm(rep_pods), norm(job_pods)
    pods = rep_pods + job_pods
    T = len(pods)
    status = np.array([s for _, s in pods], dtype=np.int32)
    req = np.array([[r[0], r[1], r[2], 1.0] for r, _ in pods], dtype=np.float64).reshape(T, 4)
    alloc = np.array([[1e9], [1e15], [1e6], [1e6]])
    idle = alloc.copy()
    for t in range(T):
        if status[t] == RUN:
            idle[:, 0] -= req[t]
    return abi.Snapshot(
        n_res=4, node_allocatable=alloc, node_idle=idle, node_releasing=np.zeros((4, 1)), node_name_rank=np.zeros(1, dtype=np.int32),
        node_flags=np.full(1, abi.NODE_READY, dtype=np.uint32), queue_parent=np.array([-1], dtype=np.int32),
        queue_priority=np.array([100], dtype=np.int32), queue_creation=np.zeros(1, dtype=np.int64),
        queue_uid_rank=np.zeros(1, dtype=np.int32), queue_deserved=np.full((3, 1), -1.0), queue_limit=np.full((3, 1), -1.0),
        queue_oqw=np.ones((3, 1)), job_queue=np.zeros(2, dtype=np.int32), job_priority=np.full(2, 50, dtype=np.int32),
        job_order_rank=np.arange(2, dtype=np.int32), job_flags=np.full(2, abi.JOB_PREEMPTIBLE, dtype=np.uint32),
        job_podset_begin=np.arange(3, dtype=np.int32), podset_min_available=np.ones(2, dtype=np.int32),
        podset_task_begin=np.array([0, len(rep_pods), T], dtype=np.int32), task_status=status,
        task_node=np.where(status == RUN, 0, -1).astype(np.int32), task_req=req,
        task_order_rank=np.concatenate([np.arange(len(rep_pods)), np.arange(len(job_pods))]).astype(np.int32),
        job_signature=np.zeros(2, dtype=np.int32))


def _call(fn_name, rep_pods, job_pods):
    o = Oracle()
    o.load(two_jobs(rep_pods, job_pods))
    fn = getattr(lib(), fn_name)
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return bool(fn(o._h, 1, 0))


SINGLE = [  # :41-137 "single pod": (representative, job, easier)
    ("empty job", (100, 0, 0), (0, 0, 0), True), ("empty representative", (0, 0, 0), (100, 0, 0), False),
    ("equal jobs", (100, 0, 0), (100, 0, 0), False), ("cpu only over", (100, 0, 0), (200, 0, 0), False),
    ("cpu only under", (100, 0, 0), (50, 0, 0), True), ("memory only over", (0, 100 * Mi, 0), (0, 200 * Mi, 0), False),
    ("memory only under", (0, 100 * Mi, 0), (0, 50 * Mi, 0), True),
    ("cpu over memory under", (100, 100 * Mi, 0), (200, 50 * Mi, 0), False),
    ("cpu under memory over", (100, 100 * Mi, 0), (50, 200 * Mi, 0), False),
    ("gpu only over", (0, 0, 1), (0, 0, 2), False), ("gpu only under", (0, 0, 2), (0, 0, 1), True),
]


@pytest.mark.parametrize("name,rep,job,easier", SINGLE, ids=[c[0] for c in SINGLE])
def test_is_easier_to_schedule_single_pod(name, rep, job, easier):
    assert _call("kai_oracle_job_easier_to_schedule", [rep], [job]) == easier


c = lambda m: (m, 0, 0)  # noqa: E731
MULTI = [  # :139-196 "multiple pods"
    ("empty representatives", [c(0), c(0)], [c(100), c(100)], False),
    ("empty job", [c(100), c(100)], [c(0), c(0)], True),
    ("only one pod is small, other is equal", [c(100), c(100)], [c(100), c(50)], True),
    ("total smaller, but one pod is bigger", [c(1000), c(100), c(100)], [c(500), c(500), c(500)], True),
]


@pytest.mark.parametrize("name,rep,job,easier", MULTI, ids=[x[0] for x in MULTI])
def test_is_easier_to_schedule_multiple_pods(name, rep, job, easier):
    assert _call("kai_oracle_job_easier_to_schedule", rep, job) == easier


STATUSES = [  # :198-283 "multiple pods, different pod statuses": only Pending pods take part
    ("representative has a failing pod", [(c(100), P), (c(100), F)], [(c(100), P), (c(100), P)], False),
    ("job has one pod not pending", [(c(100), P), (c(100), P)], [(c(100), P), (c(100), F)], True),
    ("the running pod is bigger", [(c(100), P), (c(100), P)], [(c(100), P), (c(150), RUN)], True),
]


@pytest.mark.parametrize("name,rep,job,easier", STATUSES, ids=[x[0] for x in STATUSES])
def test_is_easier_to_schedule_pod_statuses(name, rep, job, easier):
    assert _call("kai_oracle_job_easier_to_schedule", rep, job) == easier


UPDATE = [  # :286-330 UpdateRepresentative: does the job become the representative
    ("only one pod is small, other is equal", [c(100), c(100)], [c(100), c(50)], True),
    ("total smaller, but one pod is bigger", [c(1000), c(100), c(100)], [c(500), c(500), c(500)], False),
]


@pytest.mark.parametrize("name,rep,job,replaced", UPDATE, ids=[x[0] for x in UPDATE])
def test_update_representative(name, rep, job, replaced):
    assert _call("kai_oracle_job_replaces_representative", rep, job) == replaced
