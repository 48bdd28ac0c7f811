"""plugins/minruntime restated in the oracle, pinned on the reference's own unit tests (CPU).

Transcribed from pkg/scheduler/plugins/minruntime/resolver_test.go (queue fixture :342-431, expectations :38-340) and
minruntime_test.go (:98-330): durations in seconds, nil = -1.
"""
import numpy as np
import pytest

from kai_scheduler_b200 import abi, synthetic
from oracle_lib import Oracle

# createTestQueues (resolver_test.go:342-431): name -> (parent, preempt min runtime, reclaim min runtime)
QUEUES = {
    "dev": ("", 5, 10), "prod": ("", 20, 30), "research": ("", 4, 6),
    "dev-team1": ("dev", None, 8), "dev-team2": ("dev", 3, None),
    "prod-team1": ("prod", None, 25), "prod-team2": ("prod", 15, 35),
    "research-project": ("research", 7, 9),
}


def _queue_snapshot(queues=QUEUES, jobs=(), now=1000.0):
    """A snapshot that only carries the queue tree (+ optional running jobs: (queue, last_start, min_available, pods))."""
    names = list(queues)
    qi = {n: i for i, n in enumerate(names)}
    Q = len(names)
    snap = synthetic.benchmark_snapshot(2, 0, n_queues=1)
    snap.queue_parent = np.array([qi.get(queues[n][0], -1) for n in names], dtype=np.int32)
    snap.queue_priority = np.full(Q, 100, dtype=np.int32)
    snap.queue_creation = np.arange(Q, dtype=np.int64)
    snap.queue_uid_rank = np.argsort(np.argsort(np.array(names, dtype=object))).astype(np.int32)
    snap.queue_deserved = np.full((3, Q), -1.0)
    snap.queue_limit = np.full((3, Q), -1.0)
    snap.queue_oqw = np.ones((3, Q))
    snap.queue_preempt_min_runtime_s = np.array([-1.0 if queues[n][1] is None else queues[n][1] for n in names])
    snap.queue_reclaim_min_runtime_s = np.array([-1.0 if queues[n][2] is None else queues[n][2] for n in names])
    snap.now_s = now
    J = len(jobs)
    snap.job_queue = np.array([qi[j[0]] for j in jobs], dtype=np.int32).reshape(J)
    snap.job_priority = np.full(J, 50, dtype=np.int32)
    snap.job_order_rank = np.arange(J, dtype=np.int32)
    snap.job_flags = np.full(J, abi.JOB_PREEMPTIBLE, dtype=np.uint32)
    snap.job_podset_begin = np.arange(J + 1, dtype=np.int32)
    snap.podset_min_available = np.array([j[2] for j in jobs], dtype=np.int32).reshape(J)
    counts = [j[3] for j in jobs]
    snap.podset_task_begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    T = int(sum(counts))
    snap.task_status = np.full(T, abi.POD_RUNNING, dtype=np.int32)
    snap.task_node = np.zeros(T, dtype=np.int32)
    snap.task_req = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (T, 1))
    snap.task_order_rank = np.concatenate([np.arange(c) for c in counts]).astype(np.int32) if T else np.zeros(0, np.int32)
    snap.node_idle = snap.node_idle.copy()
    snap.node_idle[3, 0] -= T
    snap.job_last_start_s = np.array([-1.0 if j[1] is None else j[1] for j in jobs]).reshape(J)
    return snap, qi


def _oracle(snap, **cfg):
    o = Oracle(abi.make_config(**cfg))
    o.load(snap)
    return o


@pytest.mark.parametrize("queue,expected", [
    ("prod-team2", 15), ("prod-team1", 20), ("dev-team1", 5), ("prod", 20), ("research", 4),
])
def test_preempt_min_runtime(queue, expected):  # resolver_test.go:38-80
    snap, qi = _queue_snapshot()
    o = _oracle(snap, default_preempt_min_runtime_s=2, default_reclaim_min_runtime_s=1)
    assert o.min_runtime(False, -1, qi[queue]) == expected


def test_nil_queues_fall_back_to_the_defaults():  # resolver_test.go:82-100,134-146,199-205
    snap, qi = _queue_snapshot()
    o = _oracle(snap, default_preempt_min_runtime_s=2, default_reclaim_min_runtime_s=1)
    assert o.min_runtime(False, -1, -1) == 2
    assert o.min_runtime(True, -1, qi["dev-team2"]) == 1 and o.min_runtime(True, qi["dev-team1"], -1) == 1
    q = _oracle(snap, default_preempt_min_runtime_s=2, default_reclaim_min_runtime_s=1, reclaim_resolve_method=abi.RESOLVE_QUEUE)
    assert q.min_runtime(True, -1, qi["dev-team2"]) == 1


@pytest.mark.parametrize("pending,victim,expected", [
    ("dev-team1", "prod-team2", 35), ("dev-team1", "dev-team2", 10), ("dev-team1", "research-project", 9),
])
def test_reclaim_min_runtime_queue_method(pending, victim, expected):  # resolver_test.go:103-119,257-260
    snap, qi = _queue_snapshot()
    o = _oracle(snap, default_reclaim_min_runtime_s=1, reclaim_resolve_method=abi.RESOLVE_QUEUE)
    assert o.min_runtime(True, qi[pending], qi[victim]) == expected


@pytest.mark.parametrize("pending,victim,expected", [
    ("dev-team1", "prod-team2", 30),      # different top-level queues: the victim's top-level value
    ("dev-team2", "dev-team1", 8),        # LCA dev, victim-side child has a value
    ("prod-team1", "prod-team2", 35),
    ("prod-team2", "prod-team2", 35),     # same queue
    ("dev-team1", "research-project", 6),
    ("dev-team1", "dev-team2", 10),       # victim-side child unset -> walks up to dev
])
def test_reclaim_min_runtime_lca(pending, victim, expected):  # resolver_test.go:149-197,243-265
    snap, qi = _queue_snapshot()
    o = _oracle(snap, default_reclaim_min_runtime_s=1)
    assert o.min_runtime(True, qi[pending], qi[victim]) == expected


def test_lca_edge_cases():  # resolver_test.go:283-338
    queues = dict(QUEUES)
    queues["orphan"] = ("", None, 7)          # a queue whose parent is missing is a root of its own
    queues["no-reclaim"] = ("", None, None)
    queues["leaf"] = ("no-reclaim", None, None)
    snap, qi = _queue_snapshot(queues)
    o = _oracle(snap, default_reclaim_min_runtime_s=1)
    assert o.min_runtime(True, qi["dev-team1"], qi["orphan"]) == 7
    assert o.min_runtime(True, qi["dev-team1"], qi["leaf"]) == 1


def test_filters():  # minruntime_test.go:98-200 (defaults 5 s / 3 s, LCA)
    now = 1000.0
    jobs = [("dev-team1", None, 1, 1),          # 0 pending job's queue stand-in
            ("prod-team2", now - 10, 1, 1),     # 1 started 10 s ago: preempt 15 s -> protected; reclaim LCA 30 s -> protected
            ("prod-team2", now - 30, 1, 1),     # 2 30 s ago: preempt window over; reclaim LCA 30 s: now == until -> not protected
            ("prod-team2", None, 1, 1),         # 3 never started
            ("prod-team2", now - 20, 1, 1),     # 4 reclaim: 20 s < 30 s (LCA) and < 35 s (queue)
            ("prod-team2", now - 40, 1, 1),     # 5 reclaim window over under both methods
            ("prod-team2", now - 32, 1, 1),     # 6 between the LCA (30) and queue (35) values
            ("prod-team2", now - 10, 1, 3)]     # 7 elastic: never filtered
    snap, _ = _queue_snapshot(jobs=jobs, now=now)
    o = _oracle(snap, default_preempt_min_runtime_s=5, default_reclaim_min_runtime_s=3)
    assert [o.min_runtime_protected(False, 0, v) for v in range(1, 8)] == [True, False, False, False, False, False, False]
    assert [o.min_runtime_protected(True, 0, v) for v in range(1, 8)] == [True, False, False, True, False, False, False]
    q = _oracle(snap, default_reclaim_min_runtime_s=3, reclaim_resolve_method=abi.RESOLVE_QUEUE)
    assert [q.min_runtime_protected(True, 0, v) for v in range(1, 8)] == [True, True, False, True, False, True, False]


# ---------------------------------------------------------------------------------------------- whole actions
import minruntime_cases as mc  # noqa: E402


def _run(case, first_choice=None):
    snap, meta, cfg = mc.build(case, first_choice)
    o = Oracle(cfg)
    o.load(snap)
    return mc.outcome(o.run(case[2]), meta)


def _first_choice(action):
    base = next(c for c in mc.CASES if c[0] == f"{action}-unprotected")
    out = _run(base)
    gone = [n for n, st in out.items() if st == "Releasing"]
    assert len(gone) == 1
    return gone[0].rsplit("-", 1)[0]


@pytest.mark.parametrize("case", mc.CASES, ids=[c[0] for c in mc.CASES])
def test_actions_respect_min_runtime(case):
    first = _first_choice("reclaim") if "@first" in case[5] else None
    out = _run(case, first)
    expect = case[6]
    releasing = sorted(n for n, st in out.items() if st == "Releasing")
    pending_job = "reclaimer" if case[2] == "reclaim" else "preemptor"
    placed = out[pending_job + "-0"] == "Pipelined"
    if expect == "none":
        assert not releasing and not placed, out
    elif expect == "one-victim":
        assert len(releasing) == 1 and placed, out
    elif expect == "other-victim":
        other = "v-new" if first == "v-old" else "v-old"
        assert releasing == [other + "-0"] and placed, out
    elif expect == "elastic-one":
        assert len(releasing) == 1 and placed, out
    elif expect == "elastic-two":
        assert len(releasing) == 2 and placed, out
    elif expect == "elastic-all":
        assert len(releasing) == 3 and placed, out
    else:
        raise AssertionError(expect)


# ---------------------------------------------------------------------------------------------- wire format
from kai_scheduler_b200 import snapshot_io as sio  # noqa: E402


def test_duration_strings():  # minruntime_test.go:341-385 parseMinRuntime
    for text in ["5s", "10s", "10m5s", "1m", "1.5s", "2m30s", "1h", "1h30m", "2h45m15s", "2d4h30m", "5w4d12h"]:
        assert sio.parse_duration(text) > 0
    assert sio.parse_duration("2d4h30m") == 2 * 86400 + 4 * 3600 + 1800 and sio.parse_duration("1.5s") == 1.5
    for text in ["5", "1h2", "2h45m15", "abc", "1h-30m", "dfdsfdfdf"]:
        with pytest.raises(ValueError):
            sio.parse_duration(text)
    for text in ["-5s", "-10m", "-1h", "-2d", "-3w"]:
        assert sio.parse_duration(text) < 0
    # plugin arguments: unparsable and negative values fall back to 0 (minruntime.go:43-55)
    doc = {"config": {"tiers": [{"plugins": [{"name": "minruntime", "arguments": {
        "defaultReclaimMinRuntime": "-5s", "defaultPreemptMinRuntime": "1h2", "reclaimResolveMethod": "bogus"}}]}]}}
    kw, _ = sio._parse_config(doc)
    assert kw["default_reclaim_min_runtime_s"] == 0 and kw["default_preempt_min_runtime_s"] == 0
    assert "reclaim_resolve_method" not in kw
    doc["config"]["tiers"][0]["plugins"][0]["arguments"] = {"defaultReclaimMinRuntime": "2m", "reclaimResolveMethod": "queue"}
    kw, _ = sio._parse_config(doc)
    assert kw["default_reclaim_min_runtime_s"] == 120 and kw["reclaim_resolve_method"] == abi.RESOLVE_QUEUE


@pytest.mark.parametrize("case", mc.CASES, ids=[c[0] for c in mc.CASES])
def test_min_runtime_through_the_wire_format(case):
    """Queue min-runtimes (QueueSpec), the last-start annotation and the plugin arguments survive dump -> pack, and the
    repacked cluster behaves the same."""
    first = _first_choice("reclaim") if "@first" in case[5] else None
    snap, meta, cfg = mc.build(case, first)
    o = Oracle(cfg)
    o.load(snap)
    want = mc.outcome(o.run(case[2]), meta)
    doc = sio.dump_cluster(snap, actions=[case[2]], names=meta,
                           config=dict(case[3], allow_consolidating_reclaim=True, max_consolidation_preemptees=-1))
    snap2, meta2, kw, actions = sio.pack_cluster(doc)
    assert actions == [case[2]]
    if snap2.job_last_start_s is None:  # nothing set anywhere: the arrays stay NULL
        assert (snap.job_last_start_s <= 0).all() and (snap.queue_preempt_min_runtime_s < 0).all() \
            and (snap.queue_reclaim_min_runtime_s < 0).all()
    else:
        assert snap2.now_s == snap.now_s
        perm = [meta["queue_names"].index(q) for q in meta2["queue_names"]]
        assert np.array_equal(snap2.queue_preempt_min_runtime_s, snap.queue_preempt_min_runtime_s[perm])
        assert np.array_equal(snap2.queue_reclaim_min_runtime_s, snap.queue_reclaim_min_runtime_s[perm])
        jperm = [meta["job_names"].index(j) for j in meta2["job_names"]]
        assert np.array_equal(snap2.job_last_start_s, snap.job_last_start_s[jperm])
    o2 = Oracle(abi.make_config(**kw))
    o2.load(snap2)
    assert mc.outcome(o2.run(case[2]), meta2) == want
