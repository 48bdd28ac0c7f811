"""Node accounting invariant of the oracle under a whole cycle, on seeded random clusters (CPU only).

After every action the per-node Idle / Releasing vectors of the result must equal what the reference's NodeInfo arithmetic
gives for the final task table (pkg/scheduler/api/node_info/node_info.go:337-420 addTaskResources / removeTaskResources):

  * Allocated / Binding / Bound / Running / Releasing entries take their request out of Idle,
  * Releasing entries add it to Releasing, Pipelined entries take it out of Releasing,
  * a victim that was evicted on node A and pipelined on node B in the same session keeps TWO entries, Releasing on A and
    Pipelined on B (Statement.Pipeline with a different node adds the task to B and leaves A's entry alone,
    framework/statement.go:193-240); the result carries one (node, status) per task, so the entry on A is tracked here.

The proportion plugin's per-queue Allocated / AllocatedNotPreemptible (open-session sum over allocated statuses plus the
Allocate / Deallocate event handlers, plugins/proportion/proportion.go:347-372, :440-500) must likewise equal the sum over
the final table's active-allocated tasks, accumulated up the queue's parent chain.

Evicting a Pipelined task is legal (activeAllocatedStatuses includes Pipelined, pod_status.go:66) and can drive Idle below
zero exactly as in the reference — the invariant is on the arithmetic, not on the sign.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dsl  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from test_snapshot_io import _random_topology  # noqa: E402

from kai_scheduler_b200 import abi  # noqa: E402

S = abi.POD_STATUS_NAMES
TAKES_IDLE = S["Allocated"] | S["Binding"] | S["Bound"] | S["Running"] | S["Releasing"]
ACTIVE_ALLOCATED = S["Allocated"] | S["Pipelined"] | S["Binding"] | S["Bound"] | S["Running"]
ACTIONS = ["allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"]


def _entries(status, node, ghosts):
    """(task, node, status) entries held by the nodes: one per placed task plus the Releasing leftovers of moved victims."""
    out = [(t, int(node[t]), int(status[t])) for t in range(len(status)) if node[t] >= 0 and (int(status[t]) & (TAKES_IDLE | S["Pipelined"]))]
    return out + [(t, n, S["Releasing"]) for t, n in ghosts]


def _account(base_free, base_rel, req, entries):
    idle, rel = base_free.copy(), base_rel.copy()
    for t, n, st in entries:
        if st & TAKES_IDLE:
            idle[:, n] -= req[t]
        if st == S["Releasing"]:
            rel[:, n] += req[t]
        elif st == S["Pipelined"]:
            rel[:, n] -= req[t]
    return idle, rel


@pytest.mark.parametrize("chunk", range(8))
def test_node_vectors_follow_the_task_table(chunk):
    for seed in range(chunk * 50, (chunk + 1) * 50):
        rng = np.random.default_rng(5000 + seed)
        snap, meta = dsl.build_snapshot(_random_topology(rng))
        req = np.asarray(snap.task_req, dtype=np.float64)
        status, node, ghosts = snap.task_status.copy(), snap.task_node.copy(), []
        # what the nodes offer before any session task is counted (allocatable minus foreign pods)
        zero = np.zeros_like(snap.node_idle)
        used, held = _account(zero, zero, req, _entries(status, node, ghosts))
        base_free, base_rel = snap.node_idle - used, snap.node_releasing - held
        task_job = np.zeros(snap.n_tasks, dtype=int)
        for j in range(snap.n_jobs):
            for ps in range(snap.job_podset_begin[j], snap.job_podset_begin[j + 1]):
                task_job[snap.podset_task_begin[ps]:snap.podset_task_begin[ps + 1]] = j
        o = Oracle(abi.make_config(allow_consolidating_reclaim=True, max_consolidation_preemptees=-1))
        o.load(snap)
        for act in ACTIONS:
            res = o.run(act)
            for t in range(len(status)):
                moved = node[t] >= 0 and res.task_node[t] != node[t] and (int(status[t]) & ACTIVE_ALLOCATED)
                if moved and int(res.task_status[t]) == S["Pipelined"]:
                    ghosts.append((t, int(node[t])))
            status, node = res.task_status.copy(), res.task_node.copy()
            idle, rel = _account(base_free, base_rel, req, _entries(status, node, ghosts))
            np.testing.assert_allclose(res.node_idle, idle, rtol=0, atol=1e-9, err_msg=f"seed {seed} {act} idle")
            np.testing.assert_allclose(res.node_releasing, rel, rtol=0, atol=1e-9, err_msg=f"seed {seed} {act} releasing")
            alloc, fixed = np.zeros((3, snap.n_queues)), np.zeros((3, snap.n_queues))
            for t in np.flatnonzero(status & ACTIVE_ALLOCATED):
                j, q = task_job[t], int(snap.job_queue[task_job[t]])
                while q >= 0:
                    alloc[:, q] += req[t, :3]
                    if not int(snap.job_flags[j]) & abi.JOB_PREEMPTIBLE:
                        fixed[:, q] += req[t, :3]
                    q = int(snap.queue_parent[q])
            np.testing.assert_allclose(res.queue_allocated, alloc, rtol=0, atol=1e-9, err_msg=f"seed {seed} {act} queue allocated")
            np.testing.assert_allclose(res.queue_allocated_non_preemptible, fixed, rtol=0, atol=1e-9,
                                       err_msg=f"seed {seed} {act} queue non-preemptible")
        o.close()
