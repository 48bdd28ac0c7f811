"""Engine vs oracle over whole five-action cycles on the seeded clusters of tests/test_node_accounting_fuzz.py (GPU).

The CPU suite checks the oracle's node / queue accounting invariants on these 400 clusters; here the same clusters run
through the C ABI on the GPU and every action's outcome (bindings, statuses, visit order, node Idle / Releasing, queue
tables) must equal the oracle's, bit for bit.  These clusters reach the Statement corners the reference tests in
framework/statement_checkpoint_test.go (evict -> pipeline onto the own node -> rollback; node_info.go:495-513).
"""
import numpy as np
import pytest

import dsl
from kai_scheduler_b200 import abi, synthetic
from kai_scheduler_b200.engine import Engine
from oracle_lib import Oracle
from test_engine_gpu import assert_same
from test_snapshot_io import _random_topology

pytestmark = pytest.mark.gpu

ACTIONS = ["allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"]


@pytest.mark.parametrize("chunk", range(8))
def test_cycle_fuzz_engine_equals_oracle(chunk):
    for seed in range(chunk * 50, (chunk + 1) * 50):
        rng = np.random.default_rng(5000 + seed)
        snap, _meta = dsl.build_snapshot(_random_topology(rng))
        cfg = abi.make_config(allow_consolidating_reclaim=True, max_consolidation_preemptees=-1)
        e, o = Engine(cfg), Oracle(cfg)
        e.load(snap)
        o.load(snap)
        for act in ACTIONS:
            re_, ro = e.run(act), o.run(act)
            try:
                assert_same(re_, ro)
            except AssertionError as ex:
                raise AssertionError(f"seed {seed} action {act}: {ex}") from None
            assert re_.pods_evicted == ro.pods_evicted, f"seed {seed} action {act}"
        e.close()
        o.close()


@pytest.mark.parametrize("kw", [
    dict(n_nodes=48, running_per_node=7, victim_queues=2, reclaimer_jobs=12, reclaimer_tasks=2, reclaimer_gpus=3.0),
    dict(n_nodes=64, running_per_node=7, victim_queues=3, reclaimer_jobs=20, reclaimer_tasks=3, reclaimer_gpus=2.0),
    dict(n_nodes=96, running_per_node=6, victim_queues=4, reclaimer_jobs=24, reclaimer_tasks=2, reclaimer_gpus=3.0),
    dict(n_nodes=8, running_per_node=6, victim_queues=3, reclaimer_jobs=4, reclaimer_tasks=2, reclaimer_gpus=3.0),
    dict(n_nodes=24, running_per_node=6, victim_queues=3, reclaimer_jobs=20, reclaimer_tasks=3, reclaimer_gpus=3.0),
])
def test_remove_of_a_task_that_left_its_node(kw):
    """Workloads where consolidation evicts a victim, pipelines it back onto its own node and rolls back: the reference's
    RemoveTask is then a no-op (node_info.go:495-513) and unevict re-adds the pod (statement.go:171-183); and workloads
    where a victim moved by consolidation is evicted and re-placed by reclaim (entries on three nodes)."""
    snap = synthetic.reclaim_snapshot(**kw)
    e, o = Engine(), Oracle()
    e.load(snap)
    o.load(snap)
    for act in ("allocate", "consolidation", "reclaim", "preempt"):
        re_, ro = e.run(act), o.run(act)
        assert_same(re_, ro)
        assert re_.pods_evicted == ro.pods_evicted
    e.close()
    o.close()
