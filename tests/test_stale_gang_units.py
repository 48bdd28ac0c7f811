"""stalegangeviction grace period on the oracle (CPU): stalegangeviction.go:42-62 over one cluster of gangs that fell
below minAvailable at different instants."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dsl  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

from kai_scheduler_b200 import abi  # noqa: E402

DURATIONS = [None, 1, 29, 30, 59, 60, 61, 4000]  # seconds in stale state; None = no timestamp yet


def stale_cluster():
    """Eight gangs of three (minAvailable 3) with one failed pod each, one healthy gang that still carries a timestamp,
    one gang with a succeeded pod (never stale, job_info.go:418-420) and one elastic job above its minAvailable."""
    jobs = []
    for k, d in enumerate(DURATIONS):
        job = {"Name": f"stale-{k}", "QueueName": "q-1", "RequiredGPUsPerTask": 1, "Priority": 50, "Tasks": [
            {"State": "Running", "NodeName": f"node-{k % 4}"}, {"State": "Running", "NodeName": f"node-{(k + 1) % 4}"},
            {"State": "Failed", "NodeName": f"node-{k % 4}"}]}
        if d is not None:
            job["StaleDuration"] = float(d)
        jobs.append(job)
    jobs.append({"Name": "healthy", "QueueName": "q-1", "RequiredGPUsPerTask": 1, "Priority": 50, "StaleDuration": 4000.0,
                 "Tasks": [{"State": "Running", "NodeName": "node-0"}, {"State": "Running", "NodeName": "node-1"}]})
    jobs.append({"Name": "done", "QueueName": "q-1", "RequiredGPUsPerTask": 1, "Priority": 50, "StaleDuration": 4000.0,
                 "Tasks": [{"State": "Running", "NodeName": "node-2"}, {"State": "Succeeded", "NodeName": "node-2"},
                           {"State": "Failed", "NodeName": "node-2"}]})
    jobs.append({"Name": "elastic", "QueueName": "q-1", "RequiredGPUsPerTask": 1, "Priority": 50, "StaleDuration": 4000.0,
                 "RootSubGroupSet": {"podsets": [{"name": dsl.DEFAULT_SUBGROUP, "min_available": 1}]},
                 "Tasks": [{"State": "Running", "NodeName": "node-3"}, {"State": "Failed", "NodeName": "node-3"}]})
    return {"Nodes": {f"node-{i}": {"GPUs": 8} for i in range(4)},
            "Queues": [{"Name": "q-1", "ParentQueue": "d-1", "DeservedGPUs": 32}],
            "Departments": [{"Name": "d-1", "DeservedGPUs": 32}], "Jobs": jobs}


@pytest.mark.parametrize("grace,expected", [(-1, []), (0, list(range(8))), (30, [3, 4, 5, 6, 7]), (60, [5, 6, 7]),
                                            (3600, [7]), (4001, [])])
def test_grace_period_selects_the_gangs(grace, expected):
    snap, meta = dsl.build_snapshot(stale_cluster())
    o = Oracle(abi.make_config(staleness_grace_period_s=grace))
    o.load(snap)
    res = o.run("stalegangeviction")
    releasing = abi.POD_STATUS_NAMES["Releasing"]
    evicted_jobs = sorted({meta["job_names"][meta["task_job"][t]] for t in range(snap.n_tasks) if res.task_status[t] == releasing})
    assert evicted_jobs == [f"stale-{k}" for k in expected]
    assert res.pods_evicted == 2 * len(expected)  # the two running pods of each gang; the failed pod stays Failed
    assert sorted(meta["job_names"][j] for j, outcome in res.visits if outcome == 1) == evicted_jobs
