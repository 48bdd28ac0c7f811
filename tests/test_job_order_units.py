"""JobsOrderByQueues (actions/utils/job_order_by_queue.go) + the job / queue comparators, pinned on the reference's own
unit tests (CPU): the pop order of actions/utils/job_order_by_queue_test.go, observed as the visiting order of an
`allocate` run on a cluster where nothing fits (a failed attempt changes no queue state, so visits = pops).

The reference's sessions there register priority + elastic JobOrderFns and no queue-order plugin: queues fall back to
(CreationTimestamp, UID) with all timestamps equal (framework/session_plugins.go:284-299).  On the production path the
proportion plugin is registered and its last resort is creation time alone, "not before" meaning the right-hand queue
(queue_order.go:235-240) — for equal timestamps that is not an order at all and the pop order is heap mechanics.  The
cases are therefore run with distinct creation times that follow the UID order: both comparators then give the order
the reference's tests expect, and every share comparison before that ties because nothing has any resources.
"""
import numpy as np
import pytest

from kai_scheduler_b200 import abi
from oracle_lib import Oracle


def order_snapshot(queues: dict, jobs: list):
    """queues: name -> parent name ("" = top level); jobs: (name, priority, queue[, creation rank])."""
    qnames = list(queues)
    qi = {n: i for i, n in enumerate(qnames)}
    Q, J = len(qnames), len(jobs)
    uid_rank = np.argsort(np.argsort(np.array(qnames, dtype=object))).astype(np.int32)
    # JobOrderFn ends with (CreationTimestamp, UID): the tests leave both unset, jobs of one queue differ by priority
    job_rank = np.array([j[3] if len(j) > 3 else i for i, j in enumerate(jobs)], dtype=np.int32)
    zeros = np.zeros((4, 1))
    snap = abi.Snapshot(
        n_res=4, node_allocatable=zeros.copy(), node_idle=zeros.copy(), node_releasing=zeros.copy(),
        node_name_rank=np.zeros(1, dtype=np.int32), node_flags=np.full(1, abi.NODE_READY, dtype=np.uint32),
        queue_parent=np.array([qi.get(queues[n], -1) for n in qnames], dtype=np.int32),
        queue_priority=np.full(Q, 100, dtype=np.int32), queue_creation=uid_rank.astype(np.int64), queue_uid_rank=uid_rank,
        queue_deserved=np.full((3, Q), -1.0), queue_limit=np.full((3, Q), -1.0), queue_oqw=np.ones((3, Q)),
        job_queue=np.array([qi[j[2]] for j in jobs], dtype=np.int32), job_priority=np.array([j[1] for j in jobs], dtype=np.int32),
        job_order_rank=job_rank, job_flags=np.full(J, abi.JOB_PREEMPTIBLE, dtype=np.uint32),
        job_podset_begin=np.arange(J + 1, dtype=np.int32), podset_min_available=np.ones(J, dtype=np.int32),
        podset_task_begin=np.arange(J + 1, dtype=np.int32), task_status=np.full(J, abi.POD_PENDING, dtype=np.int32),
        task_node=np.full(J, -1, dtype=np.int32), task_req=np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (J, 1)),
        task_order_rank=np.zeros(J, dtype=np.int32))
    return snap


def pop_order(queues, jobs):
    o = Oracle()
    o.load(order_snapshot(queues, jobs))
    res = o.run("allocate")
    assert res.pods_placed == 0
    return [jobs[int(j)][0] for j, _ in res.visits]


def test_numerical_priority_within_same_queue():  # job_order_by_queue_test.go:42-144
    queues = {"test-queue": "test-parent", "test-parent": ""}
    jobs = [("p150", 150, "test-queue"), ("p255", 255, "test-queue"), ("p160", 160, "test-queue"), ("p200", 200, "test-queue")]
    assert pop_order(queues, jobs) == ["p255", "p200", "p160", "p150"]


HIERARCHY = [  # TestNLevelQueueHierarchy, job_order_by_queue_test.go:798-975 (push-job cases initialise the same tree)
    ("three level hierarchy", {"root": "", "dept1": "root", "dept2": "root", "team1": "dept1", "team2": "dept1", "team3": "dept2"},
     [("job1-team1-p100", 100, "team1"), ("job2-team2-p200", 200, "team2"), ("job3-team3-p150", 150, "team3"),
      ("job4-team1-p250", 250, "team1")],
     ["job4-team1-p250", "job1-team1-p100", "job2-team2-p200", "job3-team3-p150"]),
    ("four level hierarchy", {"org": "", "div1": "org", "dept1": "div1", "team1": "dept1"}, [("deep-job", 100, "team1")], ["deep-job"]),
    ("single level hierarchy", {"default": ""}, [("job1-default-p100", 100, "default"), ("job2-default-p200", 200, "default")],
     ["job2-default-p200", "job1-default-p100"]),
    ("two level hierarchy", {"root": "", "leaf1": "root", "leaf2": "root"},
     [("job1-leaf1-p100", 100, "leaf1"), ("job2-leaf2-p200", 200, "leaf2")], ["job1-leaf1-p100", "job2-leaf2-p200"]),
    ("mixed depth hierarchy", {"root": "", "leaf1": "root", "dept": "root", "team": "dept"},
     [("job1-shallow-p150", 150, "leaf1"), ("job2-deep-p200", 200, "team")], ["job2-deep-p200", "job1-shallow-p150"]),
    ("multiple root queues", {"root1": "", "leaf1": "root1", "root2": "", "leaf2": "root2"},
     [("job1-root1-p100", 100, "leaf1"), ("job2-root2-p200", 200, "leaf2")], ["job1-root1-p100", "job2-root2-p200"]),
    ("multiple single level root queues", {"queue-a": "", "queue-b": "", "queue-c": ""},
     [("job-a-p100", 100, "queue-a"), ("job-b-p300", 300, "queue-b"), ("job-c-p200", 200, "queue-c")],
     ["job-a-p100", "job-b-p300", "job-c-p200"]),
    ("push job builds n-level tree", {"root": "", "dept": "root", "team": "dept"},
     [("job1-p100", 100, "team"), ("job2-p200", 200, "team")], ["job2-p200", "job1-p100"]),
    ("push job to single level queue", {"default": ""}, [("pushed-job", 100, "default")], ["pushed-job"]),
    ("tree cleanup after all jobs popped", {"root": "", "dept1": "root", "dept2": "root", "team1": "dept1", "team2": "dept2"},
     [("job1-team1", 200, "team1"), ("job2-team2", 100, "team2")], ["job1-team1", "job2-team2"]),
]


@pytest.mark.parametrize("name,queues,jobs,expected", HIERARCHY, ids=[c[0] for c in HIERARCHY])
def test_n_level_queue_hierarchy(name, queues, jobs, expected):
    assert pop_order(queues, jobs) == expected


# ---------------------------------------------------------------------------------------------------------------
# plugins/elastic/elastic_test.go:18-512 TestJobOrderFn: the elastic comparator (elastic.go:25-65) between two pod groups
# (minAvailable, pod statuses) -> -1 (left first) / 0 / 1.  Observed as the pop order of two jobs of one queue with equal
# priority, run with both creation orders: a 0 lets the older job go first, -1 / 1 do not care about age.  Every job gets
# one extra Pending pod so that allocate visits it; Pending pods are not active-allocated, the compared state is unchanged.
# ---------------------------------------------------------------------------------------------------------------
RUNNING, ALLOCATED, BOUND, RELEASING = abi.POD_RUNNING, abi.POD_ALLOCATED, abi.POD_BOUND, abi.POD_RELEASING
ELASTIC = [
    ("no pods", 0, [], 0, [], 0),
    ("running single pod", 1, [RUNNING], 1, [RUNNING], 0),
    ("allocated pod counts as allocated", 1, [ALLOCATED], 1, [RUNNING], 0),
    ("bound pod counts as allocated", 1, [BOUND], 1, [RUNNING], 0),
    ("releasing pod doesn't count as allocated", 1, [RELEASING], 1, [RUNNING], -1),
    ("pod group with min pods against pod group with no pods", 1, [RUNNING], 1, [], 1),
    ("less than min against min", 2, [RUNNING], 2, [RUNNING, RUNNING], -1),
    ("less than min against more than min", 2, [RUNNING], 2, [RUNNING, RUNNING, RUNNING], -1),
    ("min against more than min", 1, [RUNNING], 1, [RUNNING, RUNNING], -1),
    ("min against less than min", 1, [RUNNING], 3, [RUNNING, RUNNING], 1),
    ("more than min against min", 1, [RUNNING, RUNNING], 1, [RUNNING], 1),
    ("more than min against min (named 'less than min' in the Go table)", 1, [RUNNING, RUNNING], 1, [RUNNING], 1),
]


def _elastic_order(l_min, l_st, r_min, r_st, l_is_older):
    statuses = [l_st + [abi.POD_PENDING], r_st + [abi.POD_PENDING]]
    flat = np.array(statuses[0] + statuses[1], dtype=np.int32)
    T = len(flat)
    on_node = flat != abi.POD_PENDING
    alloc = np.array([[0.0], [0.0], [0.0], [float(on_node.sum())]])  # every pod slot is taken: nothing fits
    idle = alloc.copy()
    idle[3, 0] = 0.0
    rel = np.zeros((4, 1))
    rel[3, 0] = float((flat == RELEASING).sum())
    snap = abi.Snapshot(
        n_res=4, node_allocatable=alloc, node_idle=idle, node_releasing=rel, node_name_rank=np.zeros(1, dtype=np.int32),
        node_flags=np.full(1, abi.NODE_READY, dtype=np.uint32), queue_parent=np.array([-1], dtype=np.int32),
        queue_priority=np.array([100], dtype=np.int32), queue_creation=np.zeros(1, dtype=np.int64),
        queue_uid_rank=np.zeros(1, dtype=np.int32), queue_deserved=np.full((3, 1), -1.0), queue_limit=np.full((3, 1), -1.0),
        queue_oqw=np.ones((3, 1)), job_queue=np.zeros(2, dtype=np.int32), job_priority=np.full(2, 50, dtype=np.int32),
        job_order_rank=np.array([0, 1] if l_is_older else [1, 0], dtype=np.int32),
        job_flags=np.full(2, abi.JOB_PREEMPTIBLE, dtype=np.uint32), job_podset_begin=np.arange(3, dtype=np.int32),
        podset_min_available=np.array([l_min, r_min], dtype=np.int32),
        podset_task_begin=np.array([0, len(statuses[0]), T], dtype=np.int32), task_status=flat,
        task_node=np.where(on_node, 0, -1).astype(np.int32), task_req=np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (T, 1)),
        task_order_rank=np.concatenate([np.arange(len(statuses[0])), np.arange(len(statuses[1]))]).astype(np.int32))
    o = Oracle()
    o.load(snap)
    res = o.run("allocate")  # (a Releasing pod's slot may take one Pending pod: only the first pop is read)
    return [int(j) for j, _ in res.visits]


@pytest.mark.parametrize("name,l_min,l_st,r_min,r_st,want", ELASTIC, ids=[c[0] for c in ELASTIC])
def test_elastic_job_order(name, l_min, l_st, r_min, r_st, want):
    for l_is_older in (True, False):
        first = _elastic_order(l_min, l_st, r_min, r_st, l_is_older)[0]
        expected_first = 0 if want < 0 else (1 if want > 0 else (0 if l_is_older else 1))
        assert first == expected_first, f"l_is_older={l_is_older}"


# ---------------------------------------------------------------------------------------------------------------
# scheduler_util/priority_queue_test.go:13-305: the PriorityQueue every job / queue / task order is built on, over the
# oracle's restatement of Go's container/heap (push, pop, peek, fix, and the max-size eviction)
# ---------------------------------------------------------------------------------------------------------------
import ctypes as C  # noqa: E402

from oracle_lib import lib  # noqa: E402

PUSH, POP, PEEK, FIX0, LEN = 0, 1, 2, 3, 4
EMPTY = -(2 ** 31)


def pq(ops, max_size=-1):
    n = len(ops)
    o = np.array([x[0] for x in ops], dtype=np.int32)
    v = np.array([x[1] if len(x) > 1 else 0 for x in ops], dtype=np.int32)
    out = np.zeros(n, dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    fn = lib().kai_oracle_priority_queue_exercise
    fn.argtypes, fn.restype = [C.c_int, C.c_int, ip, ip, ip], None
    fn(max_size, n, o.ctypes.data_as(ip), v.ctypes.data_as(ip), out.ctypes.data_as(ip))
    return out.tolist()


def test_priority_queue_push_and_pop():  # :13-108
    assert pq([(PUSH, 2), (PUSH, 3), (PUSH, 1), (LEN,), (POP,)])[-2:] == [3, 1]            # add item
    assert pq([(PUSH, 1), (PUSH, 3), (PUSH, 2), (LEN,), (POP,)])[-2:] == [3, 1]            # add less prioritized item
    assert pq([(PUSH, 2), (PUSH, 3), (PUSH, 4), (PUSH, 1), (LEN,), (POP,)], max_size=2)[-2:] == [2, 1]  # limited queue size


def test_priority_queue_peek_fix_len():  # :110-305
    assert pq([(PUSH, 2), (PUSH, 3), (LEN,), (PEEK,), (LEN,)])[-3:] == [2, 2, 2]           # basic peek
    assert pq([(LEN,), (PEEK,), (POP,)]) == [0, EMPTY, EMPTY]                               # no items
    assert pq([(PUSH, 2), (PUSH, 3), (FIX0, 4), (PEEK,)])[-1] == 3                          # basic fix
    order = pq([(PUSH, v) for v in (5, 1, 4, 2, 3)] + [(POP,)] * 5)[-5:]
    assert order == [1, 2, 3, 4, 5]
