import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from kai_scheduler_b200 import abi, synthetic
from oracle_lib import Oracle
from test_node_accounting_fuzz import _account, _entries, S, ACTIVE_ALLOCATED
snap = synthetic.reclaim_snapshot(n_nodes=48, running_per_node=7, victim_queues=2, reclaimer_jobs=12, reclaimer_tasks=2, reclaimer_gpus=3.0)
req = np.asarray(snap.task_req, dtype=np.float64)
status, node, ghosts = snap.task_status.copy(), snap.task_node.copy(), []
zero = np.zeros_like(snap.node_idle)
used, held = _account(zero, zero, req, _entries(status, node, ghosts))
base_free, base_rel = snap.node_idle - used, snap.node_releasing - held
o = Oracle(); o.load(snap)
for act in ("allocate", "consolidation"):
    res = o.run(act)
    for t in range(len(status)):
        moved = node[t] >= 0 and res.task_node[t] != node[t] and (int(status[t]) & ACTIVE_ALLOCATED)
        if moved and int(res.task_status[t]) in (S["Pipelined"], S["Releasing"]):
            ghosts.append((t, int(node[t])))
    status, node = res.task_status.copy(), res.task_node.copy()
    idle, rel = _account(base_free, base_rel, req, _entries(status, node, ghosts))
    bad = sorted(set(np.argwhere(res.node_idle != idle)[:,1].tolist()) | set(np.argwhere(res.node_releasing != rel)[:,1].tolist()))
    print(act, "bad nodes", bad, "ghosts", ghosts)
    for n in bad:
        print(" node", n, "oracle I", res.node_idle[:, n], "L", res.node_releasing[:, n], "expect I", idle[:, n], "L", rel[:, n])
        print("   tasks now on node", [(int(t), int(status[t])) for t in np.flatnonzero(node == n)], "initially", [(int(t), int(status[t]), int(node[t])) for t in np.flatnonzero(snap.task_node == n)])
