"""plugins/minruntime in the engine's solver against the oracle (GPU): the hand-built scenarios of
tests/minruntime_cases.py and victim workloads with seeded start times and per-queue min-runtimes."""
import numpy as np
import pytest

import minruntime_cases as mc
from kai_scheduler_b200 import abi, synthetic
from kai_scheduler_b200.engine import Engine
from oracle_lib import Oracle
from test_engine_gpu import assert_same

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", mc.CASES, ids=[c[0] for c in mc.CASES])
def test_min_runtime_cases_gpu(case):
    first = None
    if "@first" in case[5]:
        base = next(c for c in mc.CASES if c[0] == "reclaim-unprotected")
        snap, meta, cfg = mc.build(base)
        e = Engine(cfg)
        e.load(snap)
        out = mc.outcome(e.run("reclaim"), meta)
        e.close()
        first = [n for n, st in out.items() if st == "Releasing"][0].rsplit("-", 1)[0]
    snap, meta, cfg = mc.build(case, first)
    e, o = Engine(cfg), Oracle(cfg)
    e.load(snap)
    o.load(snap)
    assert_same(e.run(case[2]), o.run(case[2]))
    e.close()


@pytest.mark.parametrize("method", [abi.RESOLVE_LCA, abi.RESOLVE_QUEUE])
@pytest.mark.parametrize("action", ["reclaim", "preempt"])
def test_min_runtime_victim_workload(method, action):
    """64 nodes of running 1-GPU pods in three over-quota queues + reclaimers; a third of the victims started inside
    their window (windows differ per queue and per department), so the victim sets change; engine == oracle."""
    # preempt only frees the pods of the preemptor's own queue (~ a third of a node): smaller pods there
    gpus = 4.0 if action == "reclaim" else 2.0
    snap = synthetic.reclaim_snapshot(64, victim_queues=3, reclaimer_jobs=6, reclaimer_tasks=2, reclaimer_gpus=gpus)
    rng = np.random.default_rng(7)
    J, Q = snap.n_jobs, snap.n_queues
    now = 50_000.0
    snap.now_s = now
    snap.job_last_start_s = now - rng.choice(np.array([5.0, 50.0, 500.0, -1.0]), size=J)
    snap.job_last_start_s[rng.random(J) < 0.1] = -1.0
    snap.queue_reclaim_min_runtime_s = rng.choice(np.array([-1.0, 0.0, 20.0, 100.0]), size=Q)
    snap.queue_preempt_min_runtime_s = rng.choice(np.array([-1.0, 10.0, 100.0]), size=Q)
    if action == "preempt":  # preemptors: give the pending jobs a higher priority inside the victims' queues
        pending = np.array([snap.task_status[snap.podset_task_begin[snap.job_podset_begin[j]]] == abi.POD_PENDING for j in range(J)])
        snap.job_priority = np.where(pending, 75, snap.job_priority).astype(np.int32)
        snap.job_queue = np.where(pending, snap.job_queue[0], snap.job_queue).astype(np.int32)
    cfg = abi.make_config(reclaim_resolve_method=method, default_reclaim_min_runtime_s=30.0, default_preempt_min_runtime_s=30.0)
    e, o = Engine(cfg), Oracle(cfg)
    e.load(snap)
    o.load(snap)
    re_, ro = e.run(action), o.run(action)
    assert_same(re_, ro)
    # and the protection is not vacuous: the unprotected run evicts a different set
    o2 = Oracle(abi.make_config(reclaim_resolve_method=method))
    plain = synthetic.reclaim_snapshot(64, victim_queues=3, reclaimer_jobs=6, reclaimer_tasks=2, reclaimer_gpus=gpus)
    if action == "preempt":
        plain.job_priority, plain.job_queue = snap.job_priority, snap.job_queue
    o2.load(plain)
    r2 = o2.run(action)
    assert not np.array_equal(r2.task_status, ro.task_status)
    e.close()
