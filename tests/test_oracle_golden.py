"""Pin the CPU oracle against the reference's own action tables (CPU only).

Each case is a `test_utils.TestTopologyBasic` literal transcribed from
pkg/scheduler/actions/*/*_test.go by tests/golden/gen_fixtures.py; the assertions are
`MatchExpectedAndRealTasks` (pkg/scheduler/test_utils/test_utils.go:121-314) restated in tests/dsl.py.
"""
import pytest

import dsl
from fixtures import action_cases, case_config
from oracle_lib import Oracle

ALLOCATE = action_cases(["allocate__"], single_action="allocate")


@pytest.mark.parametrize("cid,case", ALLOCATE, ids=[c[0] for c in ALLOCATE])
def test_allocate_tables(cid, case):
    snap, meta = dsl.build_snapshot(case["topology"])
    o = Oracle(case_config(case))
    o.load(snap)
    res = o.run("allocate")
    errs = dsl.check_expectations(case["topology"], meta, res, snap)
    assert not errs, f"{case['source']} #{case['index']} {case['name']}: {errs}"


RECLAIM = action_cases(["reclaim__"], single_action="reclaim")
CONSOLIDATION = action_cases(["consolidation__"], single_action="consolidation")
PREEMPT = action_cases(["preempt__"], single_action="preempt")


@pytest.mark.parametrize("cid,case", RECLAIM + CONSOLIDATION + PREEMPT, ids=[c[0] for c in RECLAIM + CONSOLIDATION + PREEMPT])
def test_solver_tables(cid, case):
    """reclaim (40+5+9+11+6 tables) and consolidation (20+6): victims Releasing, preemptor Pipelined, moved victims
    Pipelined on their new node — the victim SETS the reference's tests pin."""
    snap, meta = dsl.build_snapshot(case["topology"])
    o = Oracle()
    o.load(snap)
    res = o.run(case["actions"][0])
    errs = dsl.check_expectations(case["topology"], meta, res, snap)
    assert not errs, f"{case['source']} #{case['index']} {case['name']}: {errs}"


STALE = action_cases(["stalegangeviction__"], single_action="stalegangeviction")


@pytest.mark.parametrize("cid,case", STALE, ids=[c[0] for c in STALE])
def test_stale_gang_eviction_table(cid, case):
    """actions/stalegangeviction/stalegangeviction_test.go: a gang below minAvailable (whole job or one sub group) is
    evicted once it has been stale for the 60 s grace period the test driver sets; 1 s of staleness or a satisfied gang
    with a leftover timestamp are left alone."""
    snap, meta = dsl.build_snapshot(case["topology"])
    assert case["config"] == {"staleness_grace_period_s": 60}
    o = Oracle(case_config(case))
    o.load(snap)
    res = o.run("stalegangeviction")
    errs = dsl.check_expectations(case["topology"], meta, res, snap)
    assert not errs, f"{case['source']} #{case['index']} {case['name']}: {errs}"
    # the same table with the grace period switched off (< 0) evicts nothing; with 0 every stale gang goes at once
    evicted = {}
    for grace in (-1, 0):
        o2 = Oracle(case_config(case, staleness_grace_period_s=grace))
        o2.load(snap)
        evicted[grace] = o2.run("stalegangeviction").pods_evicted
    assert evicted[-1] == 0 and evicted[0] >= res.pods_evicted
    if "recently stale" in case["name"]:
        assert res.pods_evicted == 0 and evicted[0] == 1


INTEGRATION = action_cases(["integration_tests__"])


@pytest.mark.parametrize("cid,case", INTEGRATION, ids=[c[0] for c in INTEGRATION])
def test_integration_tables(cid, case):
    """actions/integration_tests/**: allocate, consolidation, reclaim, preempt, stalegangeviction on a session that is
    rebuilt every round from the previous round's outcome (integration_tests_utils.go:41-140)."""
    errs = dsl.run_integration_case(case, lambda: Oracle())
    assert not errs, f"{case['source']} #{case['index']} {case['name']}: {errs[:3]}"


def test_case_counts():
    assert len(INTEGRATION) == 60 and len(STALE) == 6
    assert len(RECLAIM) == 66 and len(CONSOLIDATION) == 25 and len(PREEMPT) == 31
    # the transcription must not silently lose cases (allocate 21 + gang 6 + elastic 7 + subgroups 7)
    assert len(ALLOCATE) == 64  # allocate 22 + gang 6 + elastic 7 + subgroups 8 + topology 21
