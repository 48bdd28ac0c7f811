import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from kai_scheduler_b200 import abi, synthetic
from oracle_lib import Oracle
from test_node_accounting_fuzz import _account, _entries, S, ACTIVE_ALLOCATED
snap = synthetic.reclaim_snapshot(n_nodes=8, running_per_node=6, victim_queues=3, reclaimer_jobs=4, reclaimer_tasks=2, reclaimer_gpus=3.0)
req = np.asarray(snap.task_req, dtype=np.float64)
status, node, ghosts = snap.task_status.copy(), snap.task_node.copy(), []
zero = np.zeros_like(snap.node_idle)
used, held = _account(zero, zero, req, _entries(status, node, ghosts))
base_free, base_rel = snap.node_idle - used, snap.node_releasing - held
o = Oracle(); o.load(snap)
for act in ("allocate", "consolidation", "reclaim", "preempt"):
    res = o.run(act)
    for t in range(len(status)):
        moved = node[t] >= 0 and res.task_node[t] != node[t] and (int(status[t]) & ACTIVE_ALLOCATED)
        if moved and int(res.task_status[t]) in (S["Pipelined"], S["Releasing"]):
            ghosts.append((t, int(node[t])))
    status, node = res.task_status.copy(), res.task_node.copy()
    idle, rel = _account(base_free, base_rel, req, _entries(status, node, ghosts))
    print(act, "idle ok", np.array_equal(res.node_idle, idle), "rel ok", np.array_equal(res.node_releasing, rel), "ghosts", ghosts)
    if not np.array_equal(res.node_idle, idle):
        print(" oracle idle gpu", res.node_idle[2], "\n expect     ", idle[2])
        print(" oracle rel gpu", res.node_releasing[2], "\n expect    ", rel[2])
    print(" status/node of moved:", [(t, int(status[t]), int(node[t])) for t, _ in ghosts])
