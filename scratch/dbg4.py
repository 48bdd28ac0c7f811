import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from kai_scheduler_b200 import abi, synthetic
from oracle_lib import Oracle
snap = synthetic.config_snapshot("config3-cycle-small")
o = Oracle(); o.load(snap)
for a in ("allocate", "reclaim"):
    t0=time.time(); r = o.run(a); print(a, "placed", r.pods_placed, "evicted", r.pods_evicted, "visits", len(r.visits), f"{time.time()-t0:.3f}s")
print("pending left", int((r.task_status == abi.POD_PENDING).sum()))
