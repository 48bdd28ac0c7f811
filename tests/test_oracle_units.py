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
