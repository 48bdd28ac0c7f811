"""GetTasksToAllocate and its helpers (api/podgroup_info/allocation_info.go:27-177) restated in the oracle, pinned on
allocation_info_test.go (CPU): Test_GetTasksToAllocate (:62-217), Test_getTasksPriorityQueue (:309-394),
Test_getNumTasksToAllocate (:396-461) and Test_getNumAllocatableTasks (:463-535), transcribed by hand.

The Go tests build a PodGroupInfo with NewPodGroupInfo, which always carries an empty `default-sub-group` PodSet with
minAvailable 1 (job_info.go:109-111): it is part of the transcription because it counts as an unsatisfied sub group.
Sub groups are ordered by name and tasks by UID there, which is what PodSet order and task_order_rank encode here.
"""
import ctypes as C

import numpy as np
import pytest

from kai_scheduler_b200 import abi
from oracle_lib import Oracle, lib

P, A, RUN, REL, SUC = abi.POD_PENDING, abi.POD_ALLOCATED, abi.POD_RUNNING, abi.POD_RELEASING, abi.POD_SUCCEEDED


def job_snapshot(podsets: dict, with_default=True):
    """podsets: name -> (minAvailable, [(task name, status)]); one job, one queue, one roomy node."""
    sets = dict(podsets)
    if with_default and "default-sub-group" not in sets:
        sets["default-sub-group"] = (1, [])
    names = sorted(sets)
    tasks, ps_min, begin = [], [], [0]
    for n in names:
        ps_min.append(sets[n][0])
        tasks += sets[n][1]
        begin.append(len(tasks))
    T = len(tasks)
    uid_rank = np.argsort(np.argsort(np.array([t[0] for t in tasks], dtype=object))).astype(np.int32) if T else np.zeros(0, np.int32)
    alloc = np.array([[1e6], [1e12], [64.0], [110.0]])
    status = np.array([t[1] for t in tasks], dtype=np.int32).reshape(T)
    on_node = np.isin(status, [A, RUN, REL])
    idle = alloc.copy()
    idle[:, 0] -= on_node.sum() * np.array([1000.0, 1e9, 0.0, 1.0])
    rel = np.zeros((4, 1))
    rel[:, 0] = (status == REL).sum() * np.array([1000.0, 1e9, 0.0, 1.0])
    snap = abi.Snapshot(
        n_res=4, node_allocatable=alloc, node_idle=idle, node_releasing=rel, node_name_rank=np.zeros(1, dtype=np.int32),
        node_flags=np.full(1, abi.NODE_READY, dtype=np.uint32), queue_parent=np.array([-1], dtype=np.int32),
        queue_priority=np.array([100], dtype=np.int32), queue_creation=np.zeros(1, dtype=np.int64),
        queue_uid_rank=np.zeros(1, dtype=np.int32), queue_deserved=np.full((3, 1), -1.0), queue_limit=np.full((3, 1), -1.0),
        queue_oqw=np.ones((3, 1)), job_queue=np.zeros(1, dtype=np.int32), job_priority=np.array([50], dtype=np.int32),
        job_order_rank=np.zeros(1, dtype=np.int32), job_flags=np.array([abi.JOB_PREEMPTIBLE], dtype=np.uint32),
        job_podset_begin=np.array([0, len(names)], dtype=np.int32), podset_min_available=np.array(ps_min, dtype=np.int32),
        podset_task_begin=np.array(begin, dtype=np.int32), task_status=status,
        task_node=np.where(on_node, 0, -1).astype(np.int32), task_req=np.tile(np.array([1000.0, 1e9, 0.0, 1.0]), (T, 1)),
        task_order_rank=uid_rank)
    return snap, [t[0] for t in tasks]


def tasks_to_allocate(podsets, real=True, virtual=(), with_default=True):
    snap, names = job_snapshot(podsets, with_default)
    o = Oracle()
    o.load(snap)
    l = lib()
    l.kai_oracle_tasks_to_allocate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]
    l.kai_oracle_set_task_virtual.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for name in virtual:
        assert l.kai_oracle_set_task_virtual(o._h, names.index(name), 1) == 0
    out = (C.c_int32 * 64)()
    n = l.kai_oracle_tasks_to_allocate(o._h, 0, int(real), out, 64)
    return [names[out[i]] for i in range(n)]


GET_TASKS_TO_ALLOCATE = [  # allocation_info_test.go:71-192
    ("single pending task", {"subGroup1": (1, [("task1", P)])}, ["task1"]),
    ("multiple pending tasks", {"subGroup2": (2, [("task1", P), ("task2", P)])}, ["task1", "task2"]),
    ("one allocated and one pending", {"subGroup3": (1, [("task1", A), ("task2", P)])}, ["task2"]),
    ("pending in multiple subgroups, subGroups below minAvailable",
     {"subGroup1": (1, [("task1", P)]), "subGroup2": (1, [("task2", P)])}, ["task1", "task2"]),
    ("no allocatable tasks", {"subGroup4": (1, [("task1", A)])}, []),
    ("two subgroups, allocation left only in second",
     {"subGroup1": (1, [("task1", RUN)]), "subGroup2": (1, [("task2", RUN), ("task3", P)])}, ["task3"]),
    ("three subgroups, last two are not gang satisfied",
     {"subGroup1": (1, [("task1", RUN)]), "subGroup2": (1, [("task2", P)]), "subGroup3": (1, [("task3", P)])}, ["task2", "task3"]),
    ("three subgroups, all gang satisfied, allocation left in the last two",
     {"subGroup1": (1, [("task1", RUN)]), "subGroup2": (1, [("task2", RUN), ("task3", P)]),
      "subGroup3": (1, [("task4", RUN), ("task5", P)])}, ["task3"]),
]


@pytest.mark.parametrize("name,podsets,want", GET_TASKS_TO_ALLOCATE, ids=[c[0] for c in GET_TASKS_TO_ALLOCATE])
def test_get_tasks_to_allocate(name, podsets, want):
    assert tasks_to_allocate(podsets) == want


# Test_getTasksPriorityQueue (:309-394): one PodSet "subGroup1" with minAvailable 1; (tasks, real, queue length, first).
# The queue is what ShouldAllocate lets through; with minAvailable 1 and nothing allocated GetTasksToAllocate takes one
# task from its head, so `first` is observable (and the length, when it is 0).
TASK_QUEUE = [
    ("one pending task", [("task1", P)], True, (), 1, "task1"),
    ("one allocated and one pending task, only allocatable", [("task1", A), ("task2", P)], True, (), 1, "task2"),
    ("only allocated tasks", [("task1", A), ("task2", A)], True, (), 0, None),
    ("releasing and pending tasks", [("task1", REL), ("task2", P)], True, (), 1, "task2"),
    ("releasing and pending tasks (virtual allocation)", [("task1", REL), ("task2", P)], False, ("task1",), 2, "task1"),
    ("empty queue", [], True, (), 0, None),
]


@pytest.mark.parametrize("name,tasks,real,virtual,want_len,want_first", TASK_QUEUE, ids=[c[0] for c in TASK_QUEUE])
def test_tasks_priority_queue(name, tasks, real, virtual, want_len, want_first):
    got = tasks_to_allocate({"subGroup1": (1, tasks)}, real=real, virtual=virtual, with_default=False)
    if want_len == 0:
        assert got == []
    else:
        assert got[0] == want_first


# Test_getNumTasksToAllocate (:396-461): (minAvailable, statuses, want) — the number of tasks taken from one PodSet
NUM_TASKS = [
    ("pending equal to minAvailable", 3, [P, P, P], 3),
    ("allocated equal to minAvailable, plus pending", 2, [A, A, P], 1),
    ("allocated above minAvailable, extra allocatable pending", 2, [A, A, A, P], 1),
    ("allocated less than minAvailable, rest pending", 4, [A, A, P, P], 2),
    ("all allocated, at minAvailable", 3, [A, A, A], 0),
]


@pytest.mark.parametrize("name,min_available,statuses,want", NUM_TASKS, ids=[c[0] for c in NUM_TASKS])
def test_num_tasks_to_allocate(name, min_available, statuses, want):
    tasks = [(f"task-{i}", s) for i, s in enumerate(statuses)]
    assert len(tasks_to_allocate({"sg": (min_available, tasks)}, with_default=False)) == want


# Test_getNumAllocatableTasks (:463-535): how many tasks ShouldAllocate lets through; observed with a minAvailable large
# enough that GetTasksToAllocate takes them all
ALLOCATABLE = [
    ("no tasks", [], True, (), 0), ("all pending", [P, P, P], True, (), 3), ("pending and running", [P, RUN], True, (), 1),
    ("pending and releasing - real allocation", [P, REL], True, (), 1),
    ("pending and releasing - non-real allocation", [P, REL], False, ("task-1",), 2),
    ("allocated and succeeded", [A, SUC], True, (), 0), ("all succeeded", [SUC, SUC], True, (), 0),
]


@pytest.mark.parametrize("name,statuses,real,virtual,want", ALLOCATABLE, ids=[c[0] for c in ALLOCATABLE])
def test_num_allocatable_tasks(name, statuses, real, virtual, want):
    tasks = [(f"task-{i}", s) for i, s in enumerate(statuses)]
    assert len(tasks_to_allocate({"sg": (16, tasks)}, real=real, virtual=virtual, with_default=False)) == want


# plugins/subgrouporder/subgroup_order_test.go:30-102 TestSubGroupOrderFn: PodSetOrderFn (subgroup_order.go:31-62) between
# two PodSets (minAvailable, allocated pods) -> left / right / equal.  Observed through GetTasksToAllocate: each PodSet
# also holds one Pending pod and the first task returned comes from the PodSet that orders first; "equal" falls back to
# the name order, so every case runs with both name orders.
L, R_, EQ = -1, 1, 0
SUBGROUP_ORDER = [
    ("both below minAvailable, should be equal", 3, 1, 4, 2, EQ),
    ("left below, right above minAvailable", 3, 1, 3, 5, L),
    ("right below, left above minAvailable", 3, 5, 3, 1, R_),
    ("both above minAvailable, left lower allocation ratio", 2, 4, 4, 9, L),
    ("both above minAvailable, right lower allocation ratio", 2, 10, 4, 9, R_),
    ("both above minAvailable, equal allocation ratio", 2, 4, 4, 8, EQ),
]


@pytest.mark.parametrize("name,l_min,l_alloc,r_min,r_alloc,want", SUBGROUP_ORDER, ids=[c[0] for c in SUBGROUP_ORDER])
def test_subgroup_order(name, l_min, l_alloc, r_min, r_alloc, want):
    for l_name, r_name in (("a-left", "b-right"), ("b-left", "a-right")):
        podsets = {l_name: (l_min, [(f"{l_name}-{i}", A) for i in range(l_alloc)] + [(f"{l_name}-pending", P)]),
                   r_name: (r_min, [(f"{r_name}-{i}", A) for i in range(r_alloc)] + [(f"{r_name}-pending", P)])}
        first = tasks_to_allocate(podsets, with_default=False)[0]
        expected = l_name if want == L else (r_name if want == R_ else min(l_name, r_name))
        assert first == f"{expected}-pending", (l_name, r_name)
