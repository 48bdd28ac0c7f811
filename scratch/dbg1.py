import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from kai_scheduler_b200 import abi, synthetic
from kai_scheduler_b200.engine import Engine
from oracle_lib import Oracle

def run(kw, acts=("allocate", "consolidation", "reclaim", "preempt")):
    snap = synthetic.reclaim_snapshot(**kw)
    e, o = Engine(), Oracle()
    e.load(snap); o.load(snap)
    for act in acts:
        re_, ro = e.run(act), o.run(act)
        bad = []
        for f in ("task_node", "task_status", "visits", "node_idle", "node_releasing", "queue_allocated"):
            a, b = getattr(re_, f), getattr(ro, f)
            if a.shape != b.shape or not np.array_equal(a, b):
                bad.append(f)
        print(kw, act, "evicted", re_.pods_evicted, ro.pods_evicted, "placed", re_.pods_placed, ro.pods_placed, "BAD" if bad else "ok", bad)
        if bad:
            for f in ("node_idle", "node_releasing"):
                a, b = getattr(re_, f), getattr(ro, f)
                d = np.argwhere(a != b)
                nodes = sorted(set(int(x[1]) for x in d))
                print(f, "nodes", nodes)
                for n in nodes[:4]:
                    print("  node", n, "engine", a[:, n], "oracle", b[:, n])
                    ts = np.flatnonzero((ro.task_node == n))
                    print("  tasks on node (oracle):", [(int(t), int(ro.task_status[t]), int(re_.task_status[t]), int(snap.task_node[t]), int(snap.task_status[t])) for t in ts])
                    ts0 = np.flatnonzero((snap.task_node == n))
                    print("  tasks initially on node:", [(int(t), int(ro.task_node[t]), int(ro.task_status[t])) for t in ts0])
            return False
    e.close(); o.close()
    return True

run(dict(n_nodes=64, running_per_node=7, victim_queues=3, reclaimer_jobs=20, reclaimer_tasks=3, reclaimer_gpus=2.0))
# search smaller
import itertools
for n_nodes, rpn, vq, rj, rt, rg in itertools.product([8, 16, 24], [6, 7], [2, 3], [4, 8, 20], [2, 3], [2.0, 3.0]):
    ok = run(dict(n_nodes=n_nodes, running_per_node=rpn, victim_queues=vq, reclaimer_jobs=rj, reclaimer_tasks=rt, reclaimer_gpus=rg))
    if not ok:
        break
