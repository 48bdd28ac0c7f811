"""ctypes mirror of include/kai_engine.h (the C ABI of libkaigpu.so).

Only layout definitions and numpy<->C marshalling live here; no scheduling
logic.  Field order must match the header exactly.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

KAI_ABI_VERSION = 7
KAI_MAX_RES = 8
KAI_QRES = 3
RES_CPU, RES_MEM, RES_GPU, RES_PODS = 0, 1, 2, 3
UNLIMITED = -1.0

OK = 0
ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE = -1, -2, -3, -4, -5

POD_PENDING, POD_GATED, POD_ALLOCATED, POD_PIPELINED = 1, 2, 4, 8
POD_BINDING, POD_BOUND, POD_RUNNING, POD_RELEASING = 16, 32, 64, 128
POD_SUCCEEDED, POD_FAILED, POD_UNKNOWN, POD_DELETED = 256, 512, 1024, 2048
POD_ACTIVE_USED = POD_ALLOCATED | POD_PIPELINED | POD_BINDING | POD_BOUND | POD_RUNNING | POD_RELEASING
POD_STATUS_NAMES = {
    "Pending": POD_PENDING, "Gated": POD_GATED, "Allocated": POD_ALLOCATED, "Pipelined": POD_PIPELINED,
    "Binding": POD_BINDING, "Bound": POD_BOUND, "Running": POD_RUNNING, "Releasing": POD_RELEASING,
    "Succeeded": POD_SUCCEEDED, "Failed": POD_FAILED, "Unknown": POD_UNKNOWN, "Deleted": POD_DELETED,
}

NODE_READY, NODE_NOT_CPU_ONLY = 1, 2
JOB_PREEMPTIBLE = 1
ACTION_ALLOCATE, ACTION_CONSOLIDATION, ACTION_RECLAIM = 1, 2, 3
ACTION_PREEMPT = 4
ACTION_STALEGANGEVICTION = 5
ACTIONS = {"allocate": ACTION_ALLOCATE, "consolidation": ACTION_CONSOLIDATION, "reclaim": ACTION_RECLAIM,
           "preempt": ACTION_PREEMPT, "stalegangeviction": ACTION_STALEGANGEVICTION}
PLACEMENT_BINPACK, PLACEMENT_SPREAD = 0, 1
RESOLVE_LCA, RESOLVE_QUEUE = 0, 1
PEER_HANDLE_BYTES = 64

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_lp = C.POINTER(C.c_int64)


class KaiConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("gpu_placement", C.c_int32),
        ("cpu_placement", C.c_int32),
        ("k_value", C.c_double),
        ("saturation_multiplier", C.c_double),
        ("max_consolidation_preemptees", C.c_int32),
        ("allow_consolidating_reclaim", C.c_int32),
        ("shard_rank", C.c_int32),
        ("shard_count", C.c_int32),
        ("use_scheduling_signatures", C.c_int32),
        ("staleness_grace_period_s", C.c_int32),
        ("reclaim_resolve_method", C.c_int32),
        ("default_reclaim_min_runtime_s", C.c_double),
        ("default_preempt_min_runtime_s", C.c_double),
    ]


class KaiSnapshot(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_res", C.c_int32), ("n_nodes", C.c_int32), ("n_queues", C.c_int32),
        ("n_jobs", C.c_int32), ("n_podsets", C.c_int32), ("n_tasks", C.c_int32), ("n_pred_classes", C.c_int32),
        ("node_allocatable", _dp), ("node_idle", _dp), ("node_releasing", _dp), ("node_name_rank", _ip),
        ("node_flags", _up), ("node_gpu_count", _dp), ("node_foreign", _dp),
        ("queue_parent", _ip), ("queue_priority", _ip), ("queue_creation", _lp), ("queue_uid_rank", _ip),
        ("queue_deserved", _dp), ("queue_limit", _dp), ("queue_oqw", _dp), ("queue_usage", _dp),
        ("job_queue", _ip), ("job_priority", _ip), ("job_order_rank", _ip), ("job_flags", _up),
        ("job_podset_begin", _ip),
        ("podset_min_available", _ip), ("podset_task_begin", _ip),
        ("task_status", _ip), ("task_node", _ip), ("task_req", _dp), ("task_order_rank", _ip),
        ("task_nominated", _ip), ("task_pred_class", _ip),
        ("pred_mask", _up),
        ("job_signature", _ip),
        ("n_topologies", C.c_int32), ("reserved1", C.c_int32),
        ("topology_level_begin", _ip), ("node_domain", _ip),
        ("job_topology", _ip), ("job_required_level", _ip), ("job_preferred_level", _ip),
        ("n_subgroup_sets", C.c_int32), ("reserved2", C.c_int32),
        ("job_sgs_begin", _ip), ("sgs_parent", _ip), ("sgs_name_rank", _ip), ("sgs_topology", _ip),
        ("sgs_required_level", _ip), ("sgs_preferred_level", _ip), ("podset_sgs", _ip), ("podset_topology", _ip),
        ("podset_required_level", _ip), ("podset_preferred_level", _ip),
        ("now_s", C.c_double), ("queue_preempt_min_runtime_s", _dp), ("queue_reclaim_min_runtime_s", _dp),
        ("job_last_start_s", _dp), ("job_stale_since_s", _dp),
        ("structure_epoch", C.c_uint64),
    ]


class KaiJobVisit(C.Structure):
    _fields_ = [("job", C.c_int32), ("outcome", C.c_int32)]


class KaiResult(C.Structure):
    _fields_ = [
        ("n_tasks", C.c_int32), ("task_node", _ip), ("task_status", _ip),
        ("n_visits", C.c_int32), ("visits", C.POINTER(KaiJobVisit)),
        ("n_queues", C.c_int32), ("queue_fair_share", _dp), ("queue_allocated", _dp),
        ("queue_allocated_non_preemptible", _dp), ("queue_request", _dp), ("total_resource", _dp),
        ("n_nodes", C.c_int32), ("node_idle", _dp), ("node_releasing", _dp),
        ("pods_placed", C.c_int64), ("pods_evicted", C.c_int64),
    ]


class KaiStats(C.Structure):
    _fields_ = [
        ("upload_ms", C.c_double), ("open_session_ms", C.c_double), ("action_ms", C.c_double),
        ("download_ms", C.c_double), ("decisions", C.c_int64), ("nodes_scanned", C.c_int64),
        ("kernel_launches", C.c_int64), ("algorithmic_bytes", C.c_int64),
    ]


def make_config(device: int = 0, gpu_placement: int = PLACEMENT_BINPACK, cpu_placement: int = PLACEMENT_BINPACK,
                k_value: float = 1.0, saturation_multiplier: float = 1.0, max_consolidation_preemptees: int = -1,
                allow_consolidating_reclaim: bool = True, shard_rank: int = 0, shard_count: int = 1,
                use_scheduling_signatures: bool = False, staleness_grace_period_s: int = 0,
                reclaim_resolve_method: int = 0, default_reclaim_min_runtime_s: float = 0.0,
                default_preempt_min_runtime_s: float = 0.0) -> KaiConfig:
    return KaiConfig(KAI_ABI_VERSION, device, gpu_placement, cpu_placement, k_value, saturation_multiplier,
                     max_consolidation_preemptees, int(allow_consolidating_reclaim), shard_rank, shard_count,
                     int(use_scheduling_signatures), staleness_grace_period_s, reclaim_resolve_method,
                     max(0.0, default_reclaim_min_runtime_s), max(0.0, default_preempt_min_runtime_s))


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


@dataclass
class Snapshot:
    """Host-side SoA snapshot (numpy).  Array shapes follow kai_engine.h."""

    n_res: int
    node_allocatable: np.ndarray  # [R, N] f64
    node_idle: np.ndarray
    node_releasing: np.ndarray
    node_name_rank: np.ndarray  # [N] i32
    node_flags: np.ndarray  # [N] u32
    queue_parent: np.ndarray
    queue_priority: np.ndarray
    queue_creation: np.ndarray
    queue_uid_rank: np.ndarray
    queue_deserved: np.ndarray  # [3, Q]
    queue_limit: np.ndarray
    queue_oqw: np.ndarray
    job_queue: np.ndarray
    job_priority: np.ndarray
    job_order_rank: np.ndarray
    job_flags: np.ndarray
    job_podset_begin: np.ndarray  # [J+1]
    podset_min_available: np.ndarray
    podset_task_begin: np.ndarray  # [S+1]
    task_status: np.ndarray
    task_node: np.ndarray
    task_req: np.ndarray  # [T, R]
    task_order_rank: np.ndarray
    node_gpu_count: np.ndarray | None = None
    node_foreign: np.ndarray | None = None
    queue_usage: np.ndarray | None = None
    task_nominated: np.ndarray | None = None
    task_pred_class: np.ndarray | None = None
    pred_mask: np.ndarray | None = None  # [C, ceil(N/32)] u32
    job_signature: np.ndarray | None = None  # [J] i32 scheduling-constraints signature class, -1 = unique
    topology_level_begin: np.ndarray | None = None  # [n_topologies + 1]
    node_domain: np.ndarray | None = None           # [n_levels_total, N] i32
    job_topology: np.ndarray | None = None          # [J] i32, -1 none
    job_required_level: np.ndarray | None = None    # [J] i32
    job_preferred_level: np.ndarray | None = None   # [J] i32
    job_sgs_begin: np.ndarray | None = None         # [J+1] SubGroupSet tree (see include/kai_engine.h)
    sgs_parent: np.ndarray | None = None
    sgs_name_rank: np.ndarray | None = None
    sgs_topology: np.ndarray | None = None
    sgs_required_level: np.ndarray | None = None
    sgs_preferred_level: np.ndarray | None = None
    podset_sgs: np.ndarray | None = None
    podset_topology: np.ndarray | None = None
    podset_required_level: np.ndarray | None = None
    podset_preferred_level: np.ndarray | None = None
    now_s: float = 0.0                                   # min-runtime protection (plugins/minruntime)
    queue_preempt_min_runtime_s: np.ndarray | None = None  # [Q] f64 seconds, < 0 = not set
    queue_reclaim_min_runtime_s: np.ndarray | None = None
    job_last_start_s: np.ndarray | None = None             # [J] f64 seconds, <= 0 = never started
    job_stale_since_s: np.ndarray | None = None            # [J] f64 seconds, <= 0 = no staleness timestamp (stalegangeviction)
    structure_epoch: int = 0                               # != 0 and unchanged: only the per-cycle columns are reloaded (ABI v7)
    names: dict = field(default_factory=dict)  # optional: node/job/task/queue names for reporting
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_nodes(self):
        return int(self.node_name_rank.shape[0])

    @property
    def n_queues(self):
        return int(self.queue_parent.shape[0])

    @property
    def n_jobs(self):
        return int(self.job_queue.shape[0])

    @property
    def n_podsets(self):
        return int(self.podset_min_available.shape[0])

    @property
    def n_tasks(self):
        return int(self.task_status.shape[0])

    def host_bytes(self) -> int:
        tot = 0
        for v in self.__dict__.values():
            if isinstance(v, np.ndarray):
                tot += v.nbytes
        return tot

    def to_c(self) -> KaiSnapshot:
        """Build the C struct; the numpy buffers are kept alive on self._keep."""
        keep = []

        def p(a, dtype, ptr):
            if a is None:
                return C.cast(None, ptr)
            b = _arr(a, dtype)
            keep.append(b)
            return b.ctypes.data_as(ptr)

        R, N, Q = self.n_res, self.n_nodes, self.n_queues
        assert self.node_allocatable.shape == (R, N), self.node_allocatable.shape
        assert self.node_idle.shape == (R, N) and self.node_releasing.shape == (R, N)
        assert self.queue_deserved.shape == (3, Q)
        assert self.task_req.shape == (self.n_tasks, R)
        assert self.job_podset_begin.shape[0] == self.n_jobs + 1
        assert self.podset_task_begin.shape[0] == self.n_podsets + 1
        s = KaiSnapshot()
        s.abi_version = KAI_ABI_VERSION
        s.n_res, s.n_nodes, s.n_queues = R, N, Q
        s.n_jobs, s.n_podsets, s.n_tasks = self.n_jobs, self.n_podsets, self.n_tasks
        s.n_pred_classes = 0 if self.pred_mask is None else int(self.pred_mask.shape[0])
        s.node_allocatable = p(self.node_allocatable, np.float64, _dp)
        s.node_idle = p(self.node_idle, np.float64, _dp)
        s.node_releasing = p(self.node_releasing, np.float64, _dp)
        s.node_name_rank = p(self.node_name_rank, np.int32, _ip)
        s.node_flags = p(self.node_flags, np.uint32, _up)
        s.node_gpu_count = p(self.node_gpu_count, np.float64, _dp)
        s.node_foreign = p(self.node_foreign, np.float64, _dp)
        s.queue_parent = p(self.queue_parent, np.int32, _ip)
        s.queue_priority = p(self.queue_priority, np.int32, _ip)
        s.queue_creation = p(self.queue_creation, np.int64, _lp)
        s.queue_uid_rank = p(self.queue_uid_rank, np.int32, _ip)
        s.queue_deserved = p(self.queue_deserved, np.float64, _dp)
        s.queue_limit = p(self.queue_limit, np.float64, _dp)
        s.queue_oqw = p(self.queue_oqw, np.float64, _dp)
        s.queue_usage = p(self.queue_usage, np.float64, _dp)
        s.job_queue = p(self.job_queue, np.int32, _ip)
        s.job_priority = p(self.job_priority, np.int32, _ip)
        s.job_order_rank = p(self.job_order_rank, np.int32, _ip)
        s.job_flags = p(self.job_flags, np.uint32, _up)
        s.job_podset_begin = p(self.job_podset_begin, np.int32, _ip)
        s.podset_min_available = p(self.podset_min_available, np.int32, _ip)
        s.podset_task_begin = p(self.podset_task_begin, np.int32, _ip)
        s.task_status = p(self.task_status, np.int32, _ip)
        s.task_node = p(self.task_node, np.int32, _ip)
        s.task_req = p(self.task_req, np.float64, _dp)
        s.task_order_rank = p(self.task_order_rank, np.int32, _ip)
        s.task_nominated = p(self.task_nominated, np.int32, _ip)
        s.task_pred_class = p(self.task_pred_class, np.int32, _ip)
        s.pred_mask = p(self.pred_mask, np.uint32, _up)
        s.job_signature = p(self.job_signature, np.int32, _ip)
        s.n_topologies = 0 if self.topology_level_begin is None else int(len(self.topology_level_begin) - 1)
        s.topology_level_begin = p(self.topology_level_begin, np.int32, _ip)
        s.node_domain = p(self.node_domain, np.int32, _ip)
        s.job_topology = p(self.job_topology, np.int32, _ip)
        s.job_required_level = p(self.job_required_level, np.int32, _ip)
        s.job_preferred_level = p(self.job_preferred_level, np.int32, _ip)
        s.n_subgroup_sets = 0 if self.sgs_parent is None else int(len(self.sgs_parent))
        for name in ("job_sgs_begin", "sgs_parent", "sgs_name_rank", "sgs_topology", "sgs_required_level", "sgs_preferred_level",
                     "podset_sgs", "podset_topology", "podset_required_level", "podset_preferred_level"):
            setattr(s, name, p(getattr(self, name), np.int32, _ip))
        s.now_s = float(self.now_s)
        for name in ("queue_preempt_min_runtime_s", "queue_reclaim_min_runtime_s", "job_last_start_s", "job_stale_since_s"):
            setattr(s, name, p(getattr(self, name), np.float64, _dp))
        s.structure_epoch = int(self.structure_epoch)
        self._keep = keep
        return s


@dataclass
class Result:
    """Copy of a kai_result into numpy arrays (the C arrays are engine-owned)."""

    task_node: np.ndarray
    task_status: np.ndarray
    visits: np.ndarray  # [V, 2] (job, outcome)
    queue_fair_share: np.ndarray  # [3, Q]
    queue_allocated: np.ndarray
    queue_allocated_non_preemptible: np.ndarray
    queue_request: np.ndarray
    total_resource: np.ndarray
    node_idle: np.ndarray  # [R, N]
    node_releasing: np.ndarray
    pods_placed: int
    pods_evicted: int

    @staticmethod
    def from_c(r: KaiResult, n_res: int) -> "Result":
        def a(ptr, n, dtype):
            if n == 0 or not ptr:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

        T, Q, N, V = r.n_tasks, r.n_queues, r.n_nodes, r.n_visits
        if V:
            raw = np.ctypeslib.as_array(C.cast(r.visits, _ip), shape=(2 * V,)).astype(np.int32, copy=True)
            visits = raw.reshape(V, 2)
        else:
            visits = np.zeros((0, 2), np.int32)
        return Result(
            task_node=a(r.task_node, T, np.int32), task_status=a(r.task_status, T, np.int32), visits=visits,
            queue_fair_share=a(r.queue_fair_share, 3 * Q, np.float64).reshape(3, Q),
            queue_allocated=a(r.queue_allocated, 3 * Q, np.float64).reshape(3, Q),
            queue_allocated_non_preemptible=a(r.queue_allocated_non_preemptible, 3 * Q, np.float64).reshape(3, Q),
            queue_request=a(r.queue_request, 3 * Q, np.float64).reshape(3, Q),
            total_resource=a(r.total_resource, 3, np.float64),
            node_idle=a(r.node_idle, n_res * N, np.float64).reshape(n_res, N),
            node_releasing=a(r.node_releasing, n_res * N, np.float64).reshape(n_res, N),
            pods_placed=int(r.pods_placed), pods_evicted=int(r.pods_evicted))


def bind_engine_api(lib: C.CDLL, prefix: str) -> None:
    """Declare argtypes/restypes of the create/load/run/... entry points for `prefix` in {kai_engine, kai_oracle}."""
    vp = C.c_void_p
    getattr(lib, f"{prefix}_create").argtypes = [C.POINTER(KaiConfig), C.POINTER(vp)]
    getattr(lib, f"{prefix}_create").restype = C.c_int
    getattr(lib, f"{prefix}_load_snapshot").argtypes = [vp, C.POINTER(KaiSnapshot)]
    getattr(lib, f"{prefix}_load_snapshot").restype = C.c_int
    getattr(lib, f"{prefix}_run").argtypes = [vp, C.c_int, C.POINTER(KaiResult)]
    getattr(lib, f"{prefix}_run").restype = C.c_int
    getattr(lib, f"{prefix}_fair_share").argtypes = [vp, C.POINTER(KaiResult)]
    getattr(lib, f"{prefix}_fair_share").restype = C.c_int
    getattr(lib, f"{prefix}_stats").argtypes = [vp, C.POINTER(KaiStats)]
    getattr(lib, f"{prefix}_stats").restype = C.c_int
    getattr(lib, f"{prefix}_destroy").argtypes = [vp]
    getattr(lib, f"{prefix}_destroy").restype = None
