"""plugins/proportion/reclaimable restated in the oracle, pinned on the reference's own unit tests (CPU).

  * CanReclaimResources: the two literal tables of reclaimable_test.go:34-531, transcribed mechanically
    (tests/golden/can_reclaim_resources.json, 15 cases);
  * Reclaimable: the Ginkgo scenarios of reclaimable_test.go:533-1160 (BeforeEach fixture + per-case mutations),
    transcribed by hand below with the line of every `It`.
Resources are ordered (cpu, memory, gpu); a queue row is {Deserved, FairShare, Allocated, AllocatedNotPreemptible} (+
MaxAllowed for Reclaimable: unlimited on the GPU rows, the zero value on the others, as the Go fixtures have it).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle_lib import lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RES = ("CPU", "Memory", "GPU")
FIELDS = ("Deserved", "FairShare", "Allocated", "AllocatedNotPreemptible")


def _lib():
    l = lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    l.kai_oracle_can_reclaim_resources.argtypes = [dp, dp, C.c_int]
    l.kai_oracle_can_reclaim_resources.restype = C.c_int
    l.kai_oracle_reclaimable.argtypes = [C.c_int, ip, dp, C.c_double, C.c_int, C.c_int, dp, C.c_int, ip, dp]
    l.kai_oracle_reclaimable.restype = C.c_int
    return l


def _dp(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


CAN_RECLAIM = json.load(open(os.path.join(GOLDEN, "can_reclaim_resources.json")))


@pytest.mark.parametrize("case", CAN_RECLAIM, ids=[c["name"] for c in CAN_RECLAIM])
def test_can_reclaim_resources(case):
    share = [[case["share"][r][f] for f in FIELDS] for r in RES]
    _s, ps = _dp(share)
    _r, pr = _dp(case["req"])
    assert bool(_lib().kai_oracle_can_reclaim_resources(ps, pr, int(case["preemptible"]))) == case["can_reclaim"]


def test_can_reclaim_table_is_complete():
    assert len(CAN_RECLAIM) == 15 and sum(c["preemptible"] for c in CAN_RECLAIM) == 7


# ------------------------------------------------------------------------------------------------- Reclaimable
def gpu_queue(parent, deserved, fair, allocated, alloc_np=0.0):
    """buildQueues (reclaimable_test.go:1166-1186): only the GPU share is filled, CPU / memory are zero."""
    return {"parent": parent, "CPU": [0, 0, 0, 0, 0], "Memory": [0, 0, 0, 0, 0], "GPU": [deserved, fair, allocated, alloc_np, -1.0]}


def reclaimable(queues, reclaimer_queue, victims, preemptible=True, req=(0, 0, 1), multiplier=1.0):
    names = list(queues)
    idx = {n: i for i, n in enumerate(names)}
    _p, pp = _ip([idx.get(queues[n]["parent"], -1) for n in names])
    _s, ps = _dp([[queues[n][r] for r in RES] for n in names])
    _r, pr = _dp(req)
    _vq, pvq = _ip([idx[q] for q, _ in victims])
    _vr, pvr = _dp([[0, 0, g] for _, g in victims])
    return bool(_lib().kai_oracle_reclaimable(len(names), pp, ps, multiplier, idx[reclaimer_queue], int(preemptible), pr,
                                              len(victims), pvq, pvr))


def single_department():  # reclaimable_test.go:540-614
    return {"p1": gpu_queue("default", 3, 3, 2), "p2": gpu_queue("default", 2, 2, 3), "default": gpu_queue("", 5, 5, 5)}


def set_gpu(q, **kw):
    for k, v in kw.items():
        q["GPU"][FIELDS.index(k)] = v


def test_single_department():
    V = [("p2", 1.0)]  # one running 1-GPU pod of the reclaimee in p2
    q = single_department()
    assert reclaimable(q, "p1", V) is True  # :616
    q = single_department()  # :620
    set_gpu(q["p2"], Allocated=2)
    set_gpu(q["default"], Allocated=4)
    assert reclaimable(q, "p1", V) is False
    q = single_department()  # :626
    set_gpu(q["p2"], Allocated=1)
    set_gpu(q["default"], Allocated=3)
    assert reclaimable(q, "p1", V) is False
    q = single_department()  # :632 fair share of p2 raised to 3
    set_gpu(q["p2"], FairShare=3)
    assert reclaimable(q, "p1", V) is True
    q = single_department()  # :638 (AddResourceShare(7 - p2.FairShare) on the department: 5 + 5 = 10)
    set_gpu(q["p1"], Allocated=3)
    set_gpu(q["default"], Allocated=6, Deserved=7, FairShare=10)
    assert reclaimable(q, "p1", V) is False
    q = single_department()  # :646
    set_gpu(q["p1"], Allocated=3)
    set_gpu(q["default"], Allocated=6)
    assert reclaimable(q, "p1", V) is False
    q = single_department()  # :652 non-preemptible reclaimer, p1 deserved 2
    set_gpu(q["p1"], Deserved=2)
    assert reclaimable(q, "p1", V, preemptible=False) is True
    q = single_department()  # :658
    set_gpu(q["p1"], Deserved=2)
    set_gpu(q["default"], Deserved=3)
    assert reclaimable(q, "p1", V, preemptible=False) is True
    q = single_department()  # :665
    set_gpu(q["p1"], Deserved=2)
    set_gpu(q["default"], Deserved=3, AllocatedNotPreemptible=3)
    assert reclaimable(q, "p1", V, preemptible=False) is False
    q = single_department()  # :673
    set_gpu(q["p1"], Deserved=2)
    set_gpu(q["default"], Deserved=1, FairShare=1, Allocated=3)
    assert reclaimable(q, "p1", V) is True


def multiple_departments():  # reclaimable_test.go:716-776
    return {"p1": gpu_queue("d1", 3, 3, 2), "p2": gpu_queue("d2", 2, 2, 3), "d1": gpu_queue("", 3, 3, 2), "d2": gpu_queue("", 2, 2, 3)}


def test_multiple_departments():
    V = [("p2", 1.0)]
    assert reclaimable(multiple_departments(), "p1", V) is True  # :779
    q = multiple_departments()  # :783
    set_gpu(q["p2"], Allocated=2)
    set_gpu(q["d2"], Allocated=2)
    assert reclaimable(q, "p1", V) is False
    q = multiple_departments()  # :789
    set_gpu(q["p1"], Allocated=1)
    set_gpu(q["d1"], Allocated=1)
    set_gpu(q["p2"], FairShare=4)
    set_gpu(q["d2"], FairShare=4)
    assert reclaimable(q, "p1", V) is True


def test_multiple_hierarchy_levels():
    # :830 three levels on both sides, the reclaimee holds a 2-GPU pod
    q = {"left-top": gpu_queue("", 1, 1, 0), "left-mid": gpu_queue("left-top", 1, 1, 0), "left-leaf": gpu_queue("left-mid", 1, 1, 0),
         "right-top": gpu_queue("", 1, 1, 2), "right-mid": gpu_queue("right-top", 1, 1, 2), "right-leaf": gpu_queue("right-mid", 1, 1, 2)}
    assert reclaimable(q, "left-leaf", [("right-leaf", 2.0)]) is True
    # :876 the reclaimer's top queue would go over its quota
    q = {"left-top": gpu_queue("", 1, 1, 1), "left-top-oq-leaf": gpu_queue("left-top", 0, 0, 1), "left-mid": gpu_queue("left-top", 1, 1, 0),
         "left-leaf": gpu_queue("left-mid", 1, 1, 0), "right-top": gpu_queue("", 1, 1, 2), "right-mid": gpu_queue("right-top", 1, 1, 2),
         "right-leaf": gpu_queue("right-mid", 1, 1, 2)}
    assert reclaimable(q, "left-leaf", [("right-leaf", 2.0)]) is False

    def same_branch():  # :929-967, :980-1018
        return {"top": gpu_queue("", 2, 2, 2), "mid1": gpu_queue("top", 1, 1, 0.5), "mid2": gpu_queue("top", 1, 1, 1.5),
                "left-leaf1": gpu_queue("mid1", 1, 1, 0), "left-leaf2": gpu_queue("mid1", 0, 0, 0.5), "right-leaf": gpu_queue("mid2", 1, 1, 1.5)}

    assert reclaimable(same_branch(), "left-leaf1", [("right-leaf", 1.5)]) is False  # :928
    assert reclaimable(same_branch(), "left-leaf1", [("right-leaf", 1.5), ("left-leaf2", 0.5)]) is True  # :979
    # :1045 the reclaimer's utilisation ratio is lower than the reclaimee's but over 1
    q = {"d1": gpu_queue("", 4, 4, 4), "d1-project-1": gpu_queue("d1", 3, 1, 0), "d1-project-2": gpu_queue("d1", 1, 3, 4),
         "d2": gpu_queue("", 3, 3, 7), "d2-project-1": gpu_queue("d2", 3, 3, 7)}
    assert reclaimable(q, "d1-project-1", [("d2-project-1", 1.0)]) is True
    # :1104 a resource nobody asks for (CPU ratio 3.0 vs 1.0 on the departments) must not block the GPU reclaim
    q = {"d1": gpu_queue("", 4, 4, 4), "d1-project-1": gpu_queue("d1", 1, 1, 1), "d2": gpu_queue("", 3, 3, 7),
         "d2-project-1": gpu_queue("d2", 3, 3, 7)}
    q["d1"]["CPU"] = [0, 1000, 3000, 0, 0]
    q["d2"]["CPU"] = [0, 1000, 1000, 0, 0]
    assert reclaimable(q, "d1-project-1", [("d2-project-1", 1.0)]) is True


# ------------------------------------------------------------------------------------------------- idle-GPU filter
# accumulated_scenario_filters/idle_gpus/idle_gpus_test.go:106-199 Test_greedyMatchRequirements: (requirements,
# holder capacities in the holders' order, want)
GREEDY = [
    ("empty requirements always match", [], [1.0], True),
    ("zero requirements are skipped", [0, 0], [], True),
    ("single requirement matched to holder", [0.5], [1.0], True),
    ("single requirement exceeds holder capacity", [0.5], [0.0], False),
    ("virtual allocation prevents double-use of same holder", [1.0, 0.5], [1.0], False),
    ("bin-packing: two requirements fit in one holder", [1.0, 0.5], [1.5], True),
    ("second holder used after first is saturated", [1.0, 1.0], [2.0, 1.0], True),
    ("early termination: best holder below requirement", [2.0], [1.0], False),
]


@pytest.mark.parametrize("name,req,cap,want", GREEDY, ids=[c[0] for c in GREEDY])
def test_greedy_match_requirements(name, req, cap, want):
    l = lib()
    dp = C.POINTER(C.c_double)
    l.kai_oracle_greedy_match.argtypes = [C.c_int, dp, C.c_int, dp]
    _r, pr = _dp(req if req else [0.0])
    _c, pc = _dp(cap if cap else [0.0])
    assert bool(l.kai_oracle_greedy_match(len(req), pr, len(cap), pc)) == want


# ------------------------------------------------------------------------------------------------- reclaim strategies
# strategies_test.go:22-806: three literal tables, transcribed mechanically (tests/golden/reclaim_strategies.json).
# Where the Go test passes `reclaimeeQueue.GetAllocatedShare()` as the remaining share, that is the Allocated column.
STRATEGIES = json.load(open(os.path.join(GOLDEN, "reclaim_strategies.json")))
FIELDS5 = ("Deserved", "FairShare", "Allocated", "AllocatedNotPreemptible", "MaxAllowed")


@pytest.mark.parametrize("case", STRATEGIES, ids=[f"{c['context']}: {c['name']}" for c in STRATEGIES])
def test_reclaim_strategies(case):
    l = lib()
    dp = C.POINTER(C.c_double)
    l.kai_oracle_reclaim_strategy.argtypes = [C.c_int, dp, dp, dp, dp]
    rows = lambda q: [[q[r][f] for f in FIELDS5] for r in RES]  # noqa: E731
    _a, pa = _dp(rows(case["reclaimer"]))
    _b, pb = _dp(rows(case["reclaimee"]))
    _q, pq = _dp(case.get("reclaimer_req", [0.0, 0.0, 0.0]))
    if "remaining" in case:
        remaining = [case["remaining"].get(k, 0.0) for k in ("CpuResource", "MemoryResource", "GpuResource")]
    else:
        remaining = [case["reclaimee"][r]["Allocated"] for r in RES]
    _m, pm = _dp(remaining)
    strategy = 1 if case["context"].startswith("Guarantee") else 0
    assert bool(l.kai_oracle_reclaim_strategy(strategy, pa, pb, pq, pm)) == case["expected"]


def test_reclaim_strategy_tables_are_complete():
    assert len(STRATEGIES) == 24


# ------------------------------------------------------------------------------------------------- capacity policy
# capacity_policy_test.go:24-1080: four literal tables (14 cases), transcribed mechanically
# (tests/golden/capacity_policy.json); the job's requirement is getRequiredQuota over its pending pods
CAPACITY = json.load(open(os.path.join(GOLDEN, "capacity_policy.json")))


@pytest.mark.parametrize("case", CAPACITY, ids=[f"{c['function']}: {c['name']}" for c in CAPACITY])
def test_capacity_policy(case):
    l = lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    l.kai_oracle_capacity_schedulable.argtypes = [C.c_int, ip, dp, C.c_int, C.c_int, dp, C.c_int]
    names = list(case["queues"])
    idx = {n: i for i, n in enumerate(names)}
    _p, pp = _ip([idx.get(case["queues"][n]["parent"], -1) for n in names])
    _s, ps = _dp([[[case["queues"][n][r][f] for f in FIELDS5] for r in RES] for n in names])
    _r, pr = _dp(case["req"])
    mode = 1 if case["function"] == "IsNonPreemptibleJobOverQuota" else 0
    got = l.kai_oracle_capacity_schedulable(len(names), pp, ps, idx[case["queue"]], int(case["preemptible"]), pr, mode)
    assert bool(got) == case["schedulable"]


def test_capacity_policy_tables_are_complete():
    assert len(CAPACITY) == 14


# max_allowed_check_test.go (isOverLimit 8, resultsOverLimit 6) and quota_check_test.go (isAllocatedNonPreemptibleOverQuota 4,
# resultsWithNonPreemptibleOverQuota 5): the two halves of the policy on their own, with the table's requested share
CHECKS = json.load(open(os.path.join(GOLDEN, "capacity_checks.json")))


@pytest.mark.parametrize("case", CHECKS, ids=[f"{c['function']}: {c['name']}" for c in CHECKS])
def test_capacity_checks(case):
    l = lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    l.kai_oracle_capacity_schedulable.argtypes = [C.c_int, ip, dp, C.c_int, C.c_int, dp, C.c_int]
    names = list(case["queues"])
    idx = {n: i for i, n in enumerate(names)}
    _p, pp = _ip([idx.get(case["queues"][n]["parent"], -1) for n in names])
    _s, ps = _dp([[[case["queues"][n][r][f] for f in FIELDS5] for r in RES] for n in names])
    _r, pr = _dp(case["req"])
    quota_only = case["function"] in ("isAllocatedNonPreemptibleOverQuota", "resultsWithNonPreemptibleOverQuota")
    got = l.kai_oracle_capacity_schedulable(len(names), pp, ps, idx[case["queue"]], int(case["preemptible"]), pr, int(quota_only))
    assert bool(got) == case["schedulable"]


def test_capacity_check_tables_are_complete():
    from collections import Counter
    assert Counter(c["function"] for c in CHECKS) == {"isOverLimit": 8, "resultsOverLimit": 6,
                                                      "isAllocatedNonPreemptibleOverQuota": 4, "resultsWithNonPreemptibleOverQuota": 5}
