"""Known-answer tests that pin single functions of the oracle to the reference's own unit tests (CPU only).

  nodeplacement.json      plugins/nodeplacement/nodepack_test.go (exact f64 expected binpack scores per node),
                          nodespread_test.go (spread scores)
  resource_division.json  plugins/proportion/resource_division/resource_division_test.go, "two queues" table
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle_lib import lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NODEPLACEMENT = json.load(open(os.path.join(GOLDEN, "nodeplacement.json")))
RESOURCE_DIVISION = json.load(open(os.path.join(GOLDEN, "resource_division.json")))


@pytest.mark.parametrize("case", NODEPLACEMENT["binpack"], ids=[c["name"] for c in NODEPLACEMENT["binpack"]])
def test_binpack_scores_bit_exact(case):
    """getMinMaxPerNode + getScoreOfCurrentNode (pack.go:45-86): reflect.DeepEqual on float64 in the reference."""
    nodes = case["nodes"]
    mn, mx = np.finfo(np.float64).max, 0.0
    for nd in nodes.values():
        if nd["allocatable_gpus"] == 0:
            continue
        mn = min(mn, nd["idle_gpus"])
        mx = max(mx, nd["idle_gpus"])
    for name, nd in nodes.items():
        got = lib().kai_oracle_binpack_score(mn, mx, nd["idle_gpus"], nd["allocatable_gpus"])
        assert got == nd["expected_score"], f"{case['name']} node {name}: {got!r} != {nd['expected_score']!r}"


@pytest.mark.parametrize("case", NODEPLACEMENT["spread"])
def test_spread_scores(case):
    got = lib().kai_oracle_spread_score(case["non_allocated"], case["count"])
    assert got == case["expected_score"]


@pytest.mark.parametrize("case", RESOURCE_DIVISION, ids=[c["name"] for c in RESOURCE_DIVISION])
def test_set_resource_share_two_queues(case):
    ids = sorted(case["queues"])
    n = len(ids)
    q = [case["queues"][i] for i in ids]

    def arr(key, dt=np.float64):
        return np.array([x[key] for x in q], dtype=dt)

    deserved, limit, oqw, request = arr("deserved"), arr("max_allowed"), arr("oqw"), arr("request")
    usage = np.zeros(n)
    prio = arr("priority", np.int32)
    creation = np.zeros(n, dtype=np.int64)  # v1.Now() for every queue: ties fall through to the UID
    uid_rank = np.arange(n, dtype=np.int32)
    fair = arr("fair_share")
    dp = C.POINTER(C.c_double)
    rem = lib().kai_oracle_set_resource_share(
        n, case["total"], case["k_value"], deserved.ctypes.data_as(dp), limit.ctypes.data_as(dp), oqw.ctypes.data_as(dp),
        request.ctypes.data_as(dp), usage.ctypes.data_as(dp), prio.ctypes.data_as(C.POINTER(C.c_int32)),
        creation.ctypes.data_as(C.POINTER(C.c_int64)), uid_rank.ctypes.data_as(C.POINTER(C.c_int32)),
        fair.ctypes.data_as(dp))
    assert rem == case["expected_remaining"]
    for i, qid in enumerate(ids):
        if qid in case["expected_share"]:
            assert fair[i] == case["expected_share"][qid], f"queue {qid}"


# ---------------------------------------------------------------------------------------------------------------
# resource_division_test.go:26-223 — the single-queue / two-queue "sanity" contexts (Ginkgo BeforeEach fixture + a
# mutation per `It`), transcribed by hand: (line, fixture overrides, amount, expected remaining, expected FairShare)
# ---------------------------------------------------------------------------------------------------------------
def _division(fn_name, queues, amount, k_value=0.0):
    n = len(queues)
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)

    def arr(key, dt=np.float64):
        return np.array([q[key] for q in queues], dtype=dt)

    deserved, limit, oqw, request, fair = arr("Deserved"), arr("MaxAllowed"), arr("OverQuotaWeight"), arr("Request"), arr("FairShare")
    usage, prio = np.zeros(n), np.array([q.get("Priority", 100) for q in queues], dtype=np.int32)
    creation, uid_rank = np.zeros(n, dtype=np.int64), np.arange(n, dtype=np.int32)
    fn = getattr(lib(), fn_name)
    fn.argtypes = [C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, dp, ip, lp, ip, dp]
    fn.restype = C.c_double
    rem = fn(n, amount, k_value, deserved.ctypes.data_as(dp), limit.ctypes.data_as(dp), oqw.ctypes.data_as(dp),
             request.ctypes.data_as(dp), usage.ctypes.data_as(dp), prio.ctypes.data_as(ip), creation.ctypes.data_as(lp),
             uid_rank.ctypes.data_as(ip), fair.ctypes.data_as(dp))
    return rem, fair.tolist()


WITHIN_QUOTA = dict(Deserved=3, FairShare=0, OverQuotaWeight=0, MaxAllowed=-1, Request=2)  # :31-47
SET_RESOURCE_SHARE_SINGLE = [
    (50, {}, 2, 0, 2), (55, {}, 3, 1, 2), (60, {"MaxAllowed": 2}, 3, 1, 2), (66, {}, 1, 0, 2), (71, {"Request": 5}, 7, 4, 3),
    (77, {"Deserved": 1.5}, 2, 0.5, 1.5), (83, {"Request": 1.5}, 2, 0.5, 1.5), (89, {"Deserved": 0}, 2, 2, 0),
]


@pytest.mark.parametrize("line,override,total,remaining,fair", SET_RESOURCE_SHARE_SINGLE, ids=[f"L{c[0]}" for c in SET_RESOURCE_SHARE_SINGLE])
def test_set_resource_share_single_queue(line, override, total, remaining, fair):
    rem, got = _division("kai_oracle_set_resource_share", [dict(WITHIN_QUOTA, **override)], total)
    assert rem == remaining and got == [fair]


OVER_QUOTA = dict(Deserved=3, FairShare=3, OverQuotaWeight=1, MaxAllowed=-1, Request=5)  # :102-118
DIVIDE_OVER_QUOTA_SINGLE = [
    (121, {}, 2, 0, 5), (126, {}, 3, 1, 5), (131, {"MaxAllowed": 4}, 2, 1, 4), (137, {"OverQuotaWeight": 0}, 2, 2, 3),
    (143, {"Request": 4.5}, 2, 0.5, 4.5), (149, {}, 0.5, 0, 3.5), (154, {"Deserved": 0, "FairShare": 0}, 6, 1, 5),
    (161, {}, 0, 0, 3), (167, {"OverQuotaWeight": 0}, 10, 10, 3),
]


@pytest.mark.parametrize("line,override,amount,remaining,fair", DIVIDE_OVER_QUOTA_SINGLE, ids=[f"L{c[0]}" for c in DIVIDE_OVER_QUOTA_SINGLE])
def test_divide_over_quota_single_queue(line, override, amount, remaining, fair):
    rem, got = _division("kai_oracle_divide_over_quota", [dict(OVER_QUOTA, **override)], amount)
    assert rem == remaining and got == [fair]


def test_divide_over_quota_two_queues_zero_weight():  # :176-223
    q1 = dict(OVER_QUOTA, OverQuotaWeight=0)
    q2 = dict(OVER_QUOTA, Request=3)
    rem, got = _division("kai_oracle_divide_over_quota", [q1, q2], 10)
    assert rem == 10 and got[0] == 3


# resource_division_test.go:559-697 "three queues / no pre-allocated resources" (CPU; the memory contexts :877-1015
# repeat the same numbers): (line, per-queue overrides, total, expected FairShare of queues 1..3); remaining is 0
THREE = dict(Deserved=5000, FairShare=0, OverQuotaWeight=5000, MaxAllowed=-1, Request=5000)  # :564-612
THREE_QUEUES = [
    (619, [{}, {}, {}], 10000, [5000, 5000, 5000]),
    (633, [{"Request": 6000}] * 3, 18000, [6000, 6000, 6000]),
    (647, [{"Request": 7000}, {"Request": 2000}, {"Request": 2000}], 10000, [6000, 2000, 2000]),
    (655, [{"Request": 7000}, {"Request": 5000}, {"Request": 2000}], 10000, [5000, 5000, 2000]),
    (671, [{"Request": 7000, "MaxAllowed": 7000}, {"Request": 2000}, {"Request": 2000}], 10000, [6000, 2000, 2000]),
    (689, [{}, {"Deserved": 0}, {}], 10000, [5000, 0, 5000]),
]


@pytest.mark.parametrize("line,overrides,total,fair", THREE_QUEUES, ids=[f"L{c[0]}" for c in THREE_QUEUES])
def test_set_resource_share_three_queues(line, overrides, total, fair):
    rem, got = _division("kai_oracle_set_resource_share", [dict(THREE, **o) for o in overrides], total)
    assert rem == 0 and got == fair


# :699-787 "all resources allocated between 2 queues": the same fixture with Allocated = 5000 on queues 1 and 2 —
# Allocated does not enter the division (GetRequestableShare, resource_share.go:40-45), so only the numbers are kept.
# None = the reference does not assert that queue.
THREE_QUEUES_ALLOCATED = [
    (706, [{}, {}, {}], [5000, 5000, 5000]),
    (720, [{"Request": 7000}, {"Request": 2000}, {"Request": 2000}], [6000, 2000, 2000]),
    (727, [{"Request": 7000}, {"Request": 5000}, {"Request": 2000}], [5000, 5000, 2000]),
    (743, [{"Request": 7000, "MaxAllowed": 7000}, {"Request": 2000}, {"Request": 2000}], [6000, 2000, 2000]),
    (750, [{"Request": 7000, "MaxAllowed": 7000}, {"Request": 5000}, {"Request": 2000}], [5000, 5000, 2000]),
    (773, [{"Request": 10000}, {"Deserved": 0}, {"Request": 10000}], [5000, None, 5000]),
    (780, [{}, {"Deserved": 0}, {}], [5000, None, 5000]),
]


@pytest.mark.parametrize("line,overrides,fair", THREE_QUEUES_ALLOCATED, ids=[f"L{c[0]}" for c in THREE_QUEUES_ALLOCATED])
def test_set_resource_share_three_queues_allocated(line, overrides, fair):
    rem, got = _division("kai_oracle_set_resource_share", [dict(THREE, **o) for o in overrides], 10000)
    assert rem == 0
    for g, w in zip(got, fair):
        assert w is None or g == w


# resource_division_test.go:1111-2020 — the data-driven SetResourcesShare table (10 cases over 4 contexts, three
# resources per queue), transcribed mechanically into tests/golden/set_resources_share.json
SET_RESOURCES_SHARE = json.load(open(os.path.join(GOLDEN, "set_resources_share.json")))


@pytest.mark.parametrize("case", SET_RESOURCES_SHARE, ids=[f"{c['context']}: {c['name']}" for c in SET_RESOURCES_SHARE])
def test_set_resources_share_table(case):
    ids = sorted(case["queues"])
    total_key = {"GPU": "GpuResource", "CPU": "CpuResource", "Memory": "MemoryResource"}
    for res in ("GPU", "CPU", "Memory"):
        queues = [dict(case["queues"][i][res], Priority=case["queues"][i]["priority"]) for i in ids]
        rem, got = _division("kai_oracle_set_resource_share", queues, case["total"].get(total_key[res], 0.0))
        for i, qid in enumerate(ids):
            if qid in case["expected"]:
                assert got[i] == case["expected"][qid][res], f"{res} share of queue {qid}"


def test_divides_the_remainder_even_when_using_priorities():  # resource_division_test.go:401-470
    base = dict(FairShare=0, OverQuotaWeight=2, MaxAllowed=-1)
    queues = [dict(base, Deserved=2, Request=5, Priority=2), dict(base, Deserved=2, Request=5, Priority=2),
              dict(base, Deserved=1, Request=0, Priority=1)]
    rem, got = _division("kai_oracle_set_resource_share", queues, 5)
    assert rem == 0.0 and got == [3, 2, 0]


# ---------------------------------------------------------------------------------------------------------------
# queue_order_test.go:44-300 TestGetQueueOrderResult (6 cases), transcribed by hand.  A row is the 8-field
# ResourceShare {Deserved, FairShare, MaxAllowed, OverQuotaWeight, Allocated, AllocatedNotPreemptible, Request, Usage};
# CPU and memory rows are zero-valued.  The two GPU-memory cases ask for `devices x gpuMemory / minNodeGPUMemory`
# GPUs (GetTasksToAllocateInitResource with minNodeGPUMemory = 10000): 2 x 0.5 vs 2 x 1.0, and 4 x 0.8 vs 4 x 0.2.
# ---------------------------------------------------------------------------------------------------------------
def _gpu_row(deserved, fair, allocated, request):
    return [deserved, fair, -1, 1, allocated, 0, request, 0]


L_FIRST, R_FIRST = -1, 1
QUEUE_ORDER = [
    ("fair share starvation", _gpu_row(2, 99, 20, 99), _gpu_row(2, 2, 0, 2), 0, 0, 0, 0, 0, R_FIRST),
    ("quota starvation beats priority", _gpu_row(2, 99, 20, 99), _gpu_row(2, 2, 0, 2), 1, 0, 0, 0, 0, R_FIRST),
    ("priority when both are satisfied", _gpu_row(2, 99, 20, 99), _gpu_row(2, 2, 3, 99), 1, 0, 0, 0, 0, L_FIRST),
    ("priority, quota < rQueue < fair share", _gpu_row(2, 99, 20, 99), _gpu_row(2, 80, 3, 99), 1, 0, 0, 0, 0, L_FIRST),
    ("lower GPU-memory request first", _gpu_row(50, 50, 40, 50), _gpu_row(50, 50, 40, 50), 0, 0, 1.0, 2.0, 100, L_FIRST),
    ("higher GPU-memory request later", _gpu_row(50, 50, 40, 50), _gpu_row(50, 50, 40, 50), 0, 0, 3.2, 0.8, 100, R_FIRST),
]


@pytest.mark.parametrize("name,l_gpu,r_gpu,l_prio,r_prio,l_req,r_req,total_gpu,expected", QUEUE_ORDER, ids=[c[0] for c in QUEUE_ORDER])
def test_queue_order_result(name, l_gpu, r_gpu, l_prio, r_prio, l_req, r_req, total_gpu, expected):
    dp = C.POINTER(C.c_double)

    def rows(gpu):
        a = np.zeros((3, 8))
        a[2] = gpu
        return a

    l, r = rows(l_gpu), rows(r_gpu)
    lreq, rreq, total = np.array([0.0, 0.0, l_req]), np.array([0.0, 0.0, r_req]), np.array([0.0, 0.0, float(total_gpu)])
    got = lib().kai_oracle_queue_order(l.ctypes.data_as(dp), r.ctypes.data_as(dp), l_prio, r_prio, 0, 0,
                                       lreq.ctypes.data_as(dp), rreq.ctypes.data_as(dp), total.ctypes.data_as(dp))
    assert got == expected


# plugins/topology/node_scoring_test.go:106-257 TestCalculateNodeScores: the score of the nodes of the i-th of n
# preferred-level domains in sorted-tree order, in units of scores.Topology (= 10000, plugins/scores/scores.go)
@pytest.mark.parametrize("n,expected", [(1, [10.0]), (3, [3.0, 6.0, 10.0]), (4, [2.0, 5.0, 7.0, 10.0]), (2, [5.0, 10.0])])
def test_topology_position_scores(n, expected):
    fn = lib().kai_oracle_topology_position_score
    fn.argtypes, fn.restype = [C.c_int, C.c_int], C.c_double
    assert [fn(i, n) for i in range(n)] == [e * 10000.0 for e in expected]


# ---------------------------------------------------------------------------------------------------------------
# proportion_test.go:265-523 "Set fair share for 2 hierarchy queues - simplified": the level recursion of setFairShare
# (proportion.go:403-423) over d1{q1}, d2{q2}; every queue starts with GPU {Deserved 2, OverQuotaWeight 1, Request 100,
# MaxAllowed unlimited}.  (line, total GPUs, deserved / weight / priority overrides, expected fair share)
# ---------------------------------------------------------------------------------------------------------------
TWO_LEVELS = [
    (380, 4, {}, {}, {}, {"d1": 2, "q1": 2, "d2": 2, "q2": 2}),
    (393, 4, {"d1": 3, "d2": 1}, {}, {}, {"d1": 3, "q1": 3, "d2": 1, "q2": 2}),
    (410, 12, {"d1": 3, "d2": 1}, {}, {}, {"d1": 7, "q1": 7, "d2": 5, "q2": 5}),
    (427, 12, {"d1": 3, "d2": 1}, {"d1": 1, "d2": 7}, {}, {"d1": 4, "q1": 4, "d2": 8, "q2": 8}),
    (448, 12, {"d1": 3, "d2": 1}, {"d1": 7, "d2": 1}, {}, {"d1": 10, "q1": 10, "d2": 2, "q2": 2}),
    (469, 12, {"d1": 1, "q1": 1, "d2": 1, "q2": 1}, {"d1": 7, "d2": 1}, {"d1": 1, "d2": 2}, {"d1": 1, "q1": 1, "d2": 11, "q2": 11}),
    (496, 12, {"d1": 1, "q1": 1, "d2": 1, "q2": 1}, {"d1": 1, "d2": 7}, {"d1": 2, "d2": 1}, {"d1": 11, "q1": 11, "d2": 1, "q2": 1}),
]


def fair_share_tree(queues, total_gpu, k_value=0.0):
    """queues: name -> dict(parent, priority, GPU=(Deserved, MaxAllowed, OverQuotaWeight, Request)); GPU shares back."""
    names = list(queues)
    idx = {n: i for i, n in enumerate(names)}
    n = len(names)
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    parent = np.array([idx.get(queues[q]["parent"], -1) for q in names], dtype=np.int32)
    prio = np.array([queues[q].get("priority", 0) for q in names], dtype=np.int32)
    creation = np.zeros(n, dtype=np.int64)
    uid = np.argsort(np.argsort(np.array(names, dtype=object))).astype(np.int32)
    rows = np.zeros((n, 3, 4))
    for i, q in enumerate(names):
        rows[i, 2] = queues[q]["GPU"]
    total = np.array([0.0, 0.0, float(total_gpu)])
    out = np.zeros((n, 3))
    fn = lib().kai_oracle_set_fair_share_tree
    fn.argtypes = [C.c_int, ip, ip, lp, ip, dp, dp, C.c_double, dp]
    assert fn(n, parent.ctypes.data_as(ip), prio.ctypes.data_as(ip), creation.ctypes.data_as(lp), uid.ctypes.data_as(ip),
              rows.ctypes.data_as(dp), total.ctypes.data_as(dp), k_value, out.ctypes.data_as(dp)) == 0
    return {q: out[i, 2] for i, q in enumerate(names)}


@pytest.mark.parametrize("line,total,deserved,weight,priority,expected", TWO_LEVELS, ids=[f"L{c[0]}" for c in TWO_LEVELS])
def test_set_fair_share_two_levels(line, total, deserved, weight, priority, expected):
    queues = {q: {"parent": p, "priority": priority.get(q, 0), "GPU": (deserved.get(q, 2), -1, weight.get(q, 1), 100)}
              for q, p in (("d1", ""), ("q1", "d1"), ("d2", ""), ("q2", "d2"))}
    assert fair_share_tree(queues, total) == expected


def test_set_fair_share_multi_hierarchy():  # proportion_test.go:43-262 (two literal cases; OverQuotaWeight 0 unless given)
    q = lambda parent, deserved, request, oqw=0: {"parent": parent, "GPU": (deserved, -1, oqw, request)}  # noqa: E731
    got = fair_share_tree({"top-queue": q("", 3, 3), "mid-queue": q("top-queue", 3, 3), "leaf-queue-1": q("mid-queue", 2, 2),
                           "leaf-queue-2": q("mid-queue", 1, 1)}, 3)
    assert got == {"top-queue": 3, "mid-queue": 3, "leaf-queue-1": 2, "leaf-queue-2": 1}
    got = fair_share_tree({"top-queue-1": q("", 2, 2), "child-queue-1": q("top-queue-1", 1, 1), "child-queue-2": q("top-queue-1", 1, 1),
                           "top-queue-2": q("", 2, 2), "child-queue-3": q("top-queue-2", 0, 0),
                           "child-queue-4": q("top-queue-2", 1, 2, oqw=1)}, 4)
    assert got == {"top-queue-1": 2, "child-queue-1": 1, "child-queue-2": 1, "top-queue-2": 2, "child-queue-3": 0, "child-queue-4": 2}
