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


INTEGRATION = action_cases(["integration_tests__"])


@pytest.mark.parametrize("cid,case", INTEGRATION, ids=[c[0] for c in INTEGRATION])
def test_integration_tables(cid, case):
    """actions/integration_tests/**: allocate, consolidation, reclaim, preempt, stalegangeviction on a session that is
    rebuilt every round from the previous round's outcome (integration_tests_utils.go:41-140)."""
    errs = dsl.run_integration_case(case, lambda: Oracle())
    assert not errs, f"{case['source']} #{case['index']} {case['name']}: {errs[:3]}"


def test_case_counts():
    assert len(INTEGRATION) == 60
    assert len(RECLAIM) == 66 and len(CONSOLIDATION) == 25 and len(PREEMPT) == 31
    # the transcription must not silently lose cases (allocate 21 + gang 6 + elastic 7 + subgroups 7)
    assert len(ALLOCATE) == 64  # allocate 22 + gang 6 + elastic 7 + subgroups 8 + topology 21
