"""Cluster snapshot wire format <-> SoA snapshot.

The reference's `snapshot` plugin serves a zip holding `snapshot.json`: the scheduler configuration, the
SchedulerParams and the raw Kubernetes objects the cache listed (pkg/scheduler/plugins/snapshot/snapshot.go:33-66);
`cmd/snapshot-tool` replays one scheduling cycle from such a file (cmd/snapshot-tool/main.go:60-121).  This module
is the packer a shim needs on that boundary:

  pack_cluster(doc)          raw objects -> (abi.Snapshot, meta, config kwargs, action names), following
                             `ClusterInfo.Snapshot()` (pkg/scheduler/cache/cluster_info/cluster_info.go:118-228)
  dump_cluster(snap, ...)    SoA snapshot -> a snapshot.json document the reference's snapshot-tool can replay
                             (how the synthetic BASELINE configs are handed to the stock Go path on a box with Go)
  read_snapshot_zip / write_snapshot_zip

Scope: whole-GPU pods, CPU / memory / pods / extended scalar resources, queues (both fairness levels), PodGroups with
SubGroups and topology constraints, priority classes, BindRequests, Topology CRs, and the node-local Kubernetes
predicates that need no other pod's state (node conditions, spec.unschedulable, nodeSelector, required node affinity,
taints/tolerations), evaluated here into `pred_mask` classes exactly as SURVEY.md §8(c) assigns them to the host.
Fractional GPU / GPU memory / MIG / DRA requests, inter-pod affinity, topology spread, host ports and volume binding
raise UnsupportedSnapshot (strict=True) or are recorded in meta["ignored"] (strict=False).

Pure host code: no CUDA, no oracle.
"""
from __future__ import annotations

import io
import json
import math
import zipfile
from datetime import datetime, timezone
from fractions import Fraction

import numpy as np

from . import abi

SNAPSHOT_FILE_NAME = "snapshot.json"  # plugins/snapshot/snapshot.go:33-35
POD_GROUP_ANNOTATION = "pod-group-name"  # pkg/common/constants/constants.go:35
SUBGROUP_LABEL = "kai.scheduler/subgroup-name"  # constants.go:56
GPU_COUNT_LABEL = "nvidia.com/gpu.count"  # constants.go:55
TASK_ORDER_LABEL = "kai.scheduler/task-priority"
DEFAULT_SUBGROUP = "default-sub-group"
LAST_START_ANNOTATION = "kai.scheduler/last-start-timestamp"  # constants.go:43
STALE_ANNOTATION = "kai.scheduler/stale-podgroup-timestamp"  # constants.go:42
DEFAULT_QUEUE_PRIORITY = 100  # constants.go:13
DEFAULT_PODGROUP_PRIORITY = 50  # constants.go:14
NON_PREEMPTIBLE_THRESHOLD = 100  # pkg/common/podgroup/preemptible.go:10
DEFAULT_SCHEDULER_NAME = "kai-scheduler"
GPU_NAMES = ("nvidia.com/gpu", "amd.com/gpu")
MEGA = 1e6  # Queue CR memory quota/limit unit (plugins/proportion/proportion.go:48-50,327-328)


class UnsupportedSnapshot(ValueError):
    """The document uses a feature outside the packer's scope."""


# ---------------------------------------------------------------- resource.Quantity (k8s.io/apimachinery v0.34.3)
_BIN = {"Ki": 2 ** 10, "Mi": 2 ** 20, "Gi": 2 ** 30, "Ti": 2 ** 40, "Pi": 2 ** 50, "Ei": 2 ** 60}
_DEC = {"n": Fraction(1, 10 ** 9), "u": Fraction(1, 10 ** 6), "m": Fraction(1, 1000), "": Fraction(1), "k": Fraction(10 ** 3),
        "M": Fraction(10 ** 6), "G": Fraction(10 ** 9), "T": Fraction(10 ** 12), "P": Fraction(10 ** 15), "E": Fraction(10 ** 18)}


def parse_quantity(q) -> Fraction:
    """Exact value of a Kubernetes quantity ("500m", "20Gi", "2e3", 4, "1.5")."""
    if isinstance(q, (int, float)):
        return Fraction(q)
    s = str(q).strip()
    if not s:
        raise ValueError("empty quantity")
    i = 0
    if s[i] in "+-":
        i += 1
    while i < len(s) and (s[i].isdigit() or s[i] == "."):
        i += 1
    num, suffix = s[:i], s[i:]
    if num in ("", "+", "-", "."):
        raise ValueError(f"bad quantity {q!r}")
    val = Fraction(num)
    if suffix in _BIN:
        return val * _BIN[suffix]
    if suffix in _DEC:
        return val * _DEC[suffix]
    if suffix[:1] in "eE":
        return val * Fraction(10) ** int(suffix[1:])
    raise ValueError(f"bad quantity suffix in {q!r}")


def quantity_value(q) -> int:
    """Quantity.Value(): rounded up to an integer."""
    return math.ceil(parse_quantity(q))


def quantity_milli_value(q) -> int:
    """Quantity.MilliValue(): thousandths, rounded up."""
    return math.ceil(parse_quantity(q) * 1000)


def _format_int(v: float) -> str:
    if v != int(v):
        raise UnsupportedSnapshot(f"non-integral resource value {v}")
    return str(int(v))


def _epoch(ts) -> int:
    """metav1.Time (RFC 3339) -> seconds; missing = 0."""
    if not ts:
        return 0
    if isinstance(ts, (int, float)):
        return int(ts)
    return int(datetime.fromisoformat(str(ts).replace("Z", "+00:00")).timestamp())


_DUR_UNITS = {"ns": 1e-9, "us": 1e-6, "µs": 1e-6, "μs": 1e-6, "ms": 1e-3, "s": 1.0, "m": 60.0, "h": 3600.0, "d": 86400.0,
              "w": 604800.0}


def parse_duration(text) -> float:
    """Go duration string ("1h30m", "1.5s", "2d4h30m", "5w4d12h" — the plugin arguments also take days and weeks,
    plugins/minruntime/minruntime_test.go:343-351) -> seconds.  A number is taken as nanoseconds (time.Duration's JSON
    form).  Raises ValueError on anything else ("5", "1h2", "1h-30m")."""
    if isinstance(text, (int, float)):
        return float(text) * 1e-9
    s = str(text).strip()
    if s in ("0", "+0", "-0"):
        return 0.0
    sign = 1.0
    if s[:1] in "+-":
        sign = -1.0 if s[0] == "-" else 1.0
        s = s[1:]
    if not s:
        raise ValueError(f"bad duration {text!r}")
    total, i = 0.0, 0
    while i < len(s):
        j = i
        while j < len(s) and (s[j].isdigit() or s[j] == "."):
            j += 1
        if j == i or s[i:j] == ".":
            raise ValueError(f"bad duration {text!r}")
        k = j
        while k < len(s) and not (s[k].isdigit() or s[k] == "."):
            k += 1
        unit = s[j:k]
        if unit not in _DUR_UNITS:
            raise ValueError(f"bad duration unit in {text!r}")
        total += float(s[i:j]) * _DUR_UNITS[unit]
        i = k
    return sign * total


def format_duration(sec: float) -> str:
    return f"{sec:g}s"


def _rfc3339(sec: int) -> str:
    return datetime.fromtimestamp(int(sec), tz=timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ")


# ---------------------------------------------------------------- zip
def read_snapshot_zip(path_or_bytes) -> dict:
    """cmd/snapshot-tool/main.go:123-150 loadSnapshot."""
    src = io.BytesIO(path_or_bytes) if isinstance(path_or_bytes, (bytes, bytearray)) else path_or_bytes
    with zipfile.ZipFile(src) as z:
        if SNAPSHOT_FILE_NAME not in z.namelist():
            raise FileNotFoundError(SNAPSHOT_FILE_NAME)
        with z.open(SNAPSHOT_FILE_NAME) as f:
            return json.load(f)


def write_snapshot_zip(path, doc: dict) -> None:
    """plugins/snapshot/snapshot.go:206-228: one deflated member named snapshot.json."""
    with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_DEFLATED) as z:
        z.writestr(SNAPSHOT_FILE_NAME, json.dumps(doc))


# ---------------------------------------------------------------- raw objects -> resource vectors
class _ResourceNames:
    """ResourceVectorMap (api/resource_info/resource_vector.go:23-36,151-201): cpu, memory, gpu, pods, extras."""

    def __init__(self):
        self.names = ["cpu", "memory", "gpu", "pods"]
        self.index = {n: i for i, n in enumerate(self.names)}

    def slot(self, name: str) -> int:
        if name not in self.index:
            self.index[name] = len(self.names)
            self.names.append(name)
        return self.index[name]


def _is_scalar_resource_name(name: str) -> bool:
    """k8s_internal.IsScalarResourceName (k8s_internal/kubernetes_helpers.go:12-15) over the v1helper predicates of
    k8s.io/kubernetes v1.34.2 (not vendored; published rules): extended resources (a domain-qualified name outside
    kubernetes.io/ and not starting with `requests.`), hugepages-*, kubernetes.io/-prefixed native resources and
    attachable-volumes-*."""
    if name.startswith("hugepages-") or name.startswith("attachable-volumes-") or "kubernetes.io/" in name:
        return True
    native = "/" not in name
    return not native and not name.startswith("requests.")


def _is_mig(name: str) -> bool:
    return name.startswith("nvidia.com/mig-")


def _resource_list(rl: dict, names: _ResourceNames, request: bool, what: str) -> dict:
    """ResourceFromResourceList / RequirementsFromResourceList (resource_info.go:53-79, resource_requirment.go:45-72)
    -> {slot: value}.  A node's extra resources that no pod requests get no column: the fit test only walks the
    scalars a pod asks for (base_resources.go:90-105)."""
    out = {}
    for name, q in (rl or {}).items():
        if not request and name not in names.index and name not in GPU_NAMES and not _is_mig(name):
            continue
        if name == "cpu":
            v, k = quantity_milli_value(q), 0
        elif name == "memory":
            v, k = quantity_value(q), 1
        elif name in GPU_NAMES:
            frac = parse_quantity(q)
            if frac != int(frac):
                raise UnsupportedSnapshot(f"{what}: fractional {name} quantity")
            v, k = int(frac), 2
        elif name == "pods":
            if request:
                continue  # a pod's own `pods` request is always 1 (pod_info.go:390)
            v, k = quantity_value(q), 3
        elif _is_mig(name):
            if parse_quantity(q) != 0:
                raise UnsupportedSnapshot(f"{what}: MIG resource {name}")
            continue
        elif name in ("ephemeral-storage", "storage"):
            v, k = quantity_value(q), names.slot(name)
        elif _is_scalar_resource_name(name):
            v, k = quantity_milli_value(q), names.slot(name)  # scalar resources are kept in milli-units
        else:
            continue  # not a resource the scheduler accounts (resource_requirment.go:61-69)
        if v != 0:
            out[k] = out.get(k, 0) + v
    return out


def _pod_request(pod: dict, names: _ResourceNames) -> dict:
    """getPodResourceRequest (api/pod_info/pod_info.go:373-393): max(sum of containers, each init container) +
    overhead; pods = 1."""
    what = f"pod {pod['metadata'].get('namespace', '')}/{pod['metadata']['name']}"
    spec = pod.get("spec") or {}
    total = {}
    for c in spec.get("containers") or []:
        for k, v in _resource_list((c.get("resources") or {}).get("requests"), names, True, what).items():
            total[k] = total.get(k, 0) + v
    for c in spec.get("initContainers") or []:
        for k, v in _resource_list((c.get("resources") or {}).get("requests"), names, True, what).items():
            total[k] = max(total.get(k, 0), v)
    for k, v in _resource_list(spec.get("overhead"), names, True, what).items():
        if k != 2:  # result.Add(&overheadReq.BaseResource): the GPU part of an overhead is not added
            total[k] = total.get(k, 0) + v
    total[3] = 1
    return total


def _bind_request_failed(br: dict) -> bool:
    """BindRequestInfo.IsFailed (api/bindrequest_info/binrequest_info.go:84-92): failed for good once the attempts reach
    the backoff limit (or no limit is set)."""
    status, spec = br.get("status") or {}, br.get("spec") or {}
    if status.get("phase") != "Failed":
        return False
    if spec.get("backoffLimit") is None:
        return True
    return int(status.get("failedAttempts", 0) or 0) >= int(spec["backoffLimit"])


KAI_UTILITY_APPS = ("kai-resource-reservation", "scaling-pod")  # conf/global_config.go:25-26


def _is_kai_utility_pod(pod: dict) -> bool:
    """pod_info.IsKaiUtilityPod (api/pod_info/utility_pods.go): GPU reservation and scale-adjust pods, by `app` label."""
    return (pod["metadata"].get("labels") or {}).get("app") in KAI_UTILITY_APPS


def _task_status(pod: dict, bind_request) -> int:
    """getTaskStatus (pod_info.go:414-446)."""
    phase = (pod.get("status") or {}).get("phase", "")
    deleting = pod["metadata"].get("deletionTimestamp") is not None
    spec = pod.get("spec") or {}
    if phase == "Running":
        return abi.POD_RELEASING if deleting else abi.POD_RUNNING
    if phase == "Pending":
        if deleting:
            return abi.POD_RELEASING
        if spec.get("nodeName"):
            return abi.POD_BOUND
        if bind_request is not None:
            return abi.POD_BINDING
        if spec.get("schedulingGates"):
            return abi.POD_GATED
        return abi.POD_PENDING
    if phase == "Succeeded":
        return abi.POD_SUCCEEDED
    if phase == "Failed":
        return abi.POD_FAILED
    return abi.POD_UNKNOWN


ACTIVE_USED = (abi.POD_ALLOCATED | abi.POD_PIPELINED | abi.POD_BINDING | abi.POD_BOUND | abi.POD_RUNNING | abi.POD_RELEASING)
ACTIVE_ALLOCATED = (abi.POD_ALLOCATED | abi.POD_BINDING | abi.POD_BOUND | abi.POD_RUNNING)


# ---------------------------------------------------------------- node-local predicates
def _node_ready(node: dict) -> bool:
    """CheckNodeConditionPredicate (scheduler_util/scheduler_utils.go:12-40)."""
    if (node.get("spec") or {}).get("unschedulable"):
        return False
    for c in (node.get("status") or {}).get("conditions") or []:
        if c.get("type") == "Ready":
            if c.get("status") != "True":
                return False
        elif c.get("type") in ("MemoryPressure", "DiskPressure", "PIDPressure", "NetworkUnavailable"):
            if c.get("status") != "False":
                return False
    return True


def _match_expression(labels: dict, e: dict, fields: bool = False, node_name: str = "") -> bool:
    key, op, values = e.get("key"), e.get("operator"), e.get("values") or []
    if fields:
        present, val = key == "metadata.name", node_name
    else:
        present, val = key in labels, labels.get(key)
    if op == "In":
        return present and val in values
    if op == "NotIn":
        return not (present and val in values)
    if op == "Exists":
        return present
    if op == "DoesNotExist":
        return not present
    if op in ("Gt", "Lt"):
        try:
            a, b = int(val), int(values[0])
        except (TypeError, ValueError, IndexError):
            return False
        return present and (a > b if op == "Gt" else a < b)
    raise UnsupportedSnapshot(f"node selector operator {op!r}")


def _tolerates(tol: dict, taint: dict) -> bool:
    """v1.Toleration.ToleratesTaint."""
    if tol.get("effect") and tol["effect"] != taint.get("effect"):
        return False
    if tol.get("key") and tol["key"] != taint.get("key"):
        return False
    op = tol.get("operator") or "Equal"
    if op == "Exists":
        return True
    if op == "Equal":
        return (tol.get("value") or "") == (taint.get("value") or "")
    return False


def _pod_fits_node(cons: dict, node: dict) -> bool:
    labels = node["metadata"].get("labels") or {}
    for k, v in (cons.get("nodeSelector") or {}).items():
        if labels.get(k) != v:
            return False
    terms = cons.get("terms")
    if terms is not None:
        ok = False
        for t in terms:
            exprs, fields = t.get("matchExpressions") or [], t.get("matchFields") or []
            if not exprs and not fields:
                continue  # an empty term matches nothing
            if all(_match_expression(labels, e) for e in exprs) and \
               all(_match_expression(labels, e, True, node["metadata"]["name"]) for e in fields):
                ok = True
                break
        if not ok:
            return False
    for taint in (node.get("spec") or {}).get("taints") or []:
        if taint.get("effect") not in ("NoSchedule", "NoExecute"):
            continue
        if not any(_tolerates(t, taint) for t in cons.get("tolerations") or []):
            return False
    return True


def _pod_constraints(pod: dict, strict: bool, ignored: list) -> dict:
    spec = pod.get("spec") or {}
    what = f"pod {pod['metadata'].get('namespace', '')}/{pod['metadata']['name']}"

    def unsupported(msg):
        if strict:
            raise UnsupportedSnapshot(f"{what}: {msg}")
        ignored.append(f"{what}: {msg}")

    aff = spec.get("affinity") or {}
    if aff.get("podAffinity") or aff.get("podAntiAffinity"):
        unsupported("inter-pod affinity")
    if spec.get("topologySpreadConstraints"):
        unsupported("topology spread constraints")
    for c in (spec.get("containers") or []) + (spec.get("initContainers") or []):
        if any(p.get("hostPort") for p in c.get("ports") or []):
            unsupported("host ports")
    if spec.get("resourceClaims"):
        unsupported("DRA resource claims")
    if any(v.get("persistentVolumeClaim") for v in spec.get("volumes") or []):
        unsupported("persistent volume claims")
    ann = pod["metadata"].get("annotations") or {}
    for k in ("gpu-fraction", "gpu-memory", "gpu-fraction-num-devices"):
        if ann.get(k):
            raise UnsupportedSnapshot(f"{what}: annotation {k} (GPU sharing)")
    req = ((aff.get("nodeAffinity") or {}).get("requiredDuringSchedulingIgnoredDuringExecution") or {})
    return {"nodeSelector": spec.get("nodeSelector") or None,
            "terms": req.get("nodeSelectorTerms") if req else None,
            "tolerations": spec.get("tolerations") or None}


def _pod_signature(pod: dict) -> str:
    """The fields schedulingConstraintsSignature hashes (api/pod_info/scheduling_constraints_signature.go:19-96),
    as a canonical string: equal strings <=> equal reference signatures for the supported fields."""
    spec = pod.get("spec") or {}
    ports = [p.get("hostPort", 0) for c in (spec.get("containers") or []) + (spec.get("initContainers") or [])
             for p in c.get("ports") or []]
    tol = [[t.get("key", ""), t.get("operator", ""), t.get("value", ""), t.get("effect", "")]
           for t in spec.get("tolerations") or []] if spec.get("tolerations") is not None else None
    return json.dumps([spec.get("nodeSelector"), spec.get("affinity"), tol, spec.get("priorityClassName", ""),
                       spec.get("priority"), spec.get("topologySpreadConstraints"), ports], sort_keys=True)


# ---------------------------------------------------------------- config
def _parse_config(doc: dict):
    """conf.SchedulerConfiguration + SchedulerParams -> (make_config kwargs, action names)."""
    conf = doc.get("config") or {}
    params = doc.get("schedulerParams") or {}
    actions = [a.strip() for a in (conf.get("actions") or "allocate").split(",") if a.strip()]
    kw = {}
    place = {"binpack": abi.PLACEMENT_BINPACK, "spread": abi.PLACEMENT_SPREAD}
    for tier in conf.get("tiers") or []:
        for p in tier.get("plugins") or []:
            args = p.get("arguments") or {}
            if p.get("name") == "nodeplacement":  # plugins/nodeplacement/nodeplacement.go:53-73
                if "gpu" in args:
                    kw["gpu_placement"] = place[args["gpu"]]
                if "cpu" in args:
                    kw["cpu_placement"] = place[args["cpu"]]
            elif p.get("name") == "minruntime":  # plugins/minruntime/minruntime.go:43-78
                for key, field in (("defaultReclaimMinRuntime", "default_reclaim_min_runtime_s"),
                                   ("defaultPreemptMinRuntime", "default_preempt_min_runtime_s")):
                    if key in args:
                        try:
                            kw[field] = max(0.0, parse_duration(args[key]))  # unparsable or negative -> 0
                        except ValueError:
                            kw[field] = 0.0
                if args.get("reclaimResolveMethod") == "queue":  # anything else falls back to lca
                    kw["reclaim_resolve_method"] = abi.RESOLVE_QUEUE
            elif p.get("name") == "proportion":  # plugins/proportion/proportion.go:68-85
                if "kValue" in args:
                    kw["k_value"] = float(args["kValue"])
                if "relcaimerSaturationMultiplier" in args:
                    kw["saturation_multiplier"] = float(args["relcaimerSaturationMultiplier"])
    # omitted = Go's zero value = consolidation disabled (actions/consolidation/consolidation.go:36-39); -1 = no limit
    kw["max_consolidation_preemptees"] = int(params.get("maxNumberConsolidationPreemptees", 0) or 0)
    kw["use_scheduling_signatures"] = bool(params.get("useSchedulingSignatures", False))
    kw["allow_consolidating_reclaim"] = bool(params.get("allowConsolidatingReclaim", False))
    grace_ns = params.get("globalDefaultStalenessGracePeriod", 0) or 0  # time.Duration marshals as nanoseconds
    kw["staleness_grace_period_s"] = -1 if grace_ns < 0 else int(grace_ns // 10 ** 9)
    return kw, actions


# ---------------------------------------------------------------- pack
def pack_cluster(doc: dict, strict: bool = True, now: float | None = None):
    """ClusterInfo.Snapshot() over the raw objects of a snapshot.json document.

    Returns (snapshot, meta, config_kwargs, actions).  meta: node_names, queue_names, job_names, task_names (pod
    names), task_uids, task_job, resource_names, ignored.  Index order: nodes and queues by name, jobs by PodGroup
    name, tasks by PodSet then (creation, UID) — any order is valid for the engine, this one is reproducible.
    `now` (seconds since the epoch) is the instant min-runtime windows are measured against: the document's own
    `capturedAt` extension if present, else the argument, else the current time (the reference reads time.Now()).
    """
    raw = doc.get("rawObjects") or {}
    params = doc.get("schedulerParams") or {}
    if params.get("restrictSchedulingNodes"):
        raise UnsupportedSnapshot("restrictSchedulingNodes")
    scheduler_name = params.get("schedulerName") or DEFAULT_SCHEDULER_NAME
    ignored: list = []
    names = _ResourceNames()

    # SchedulingNodePoolParams.GetLabelSelector (conf/scheduler_conf.go:95-112): the partition selector the listers of
    # nodes, queues and pod groups apply (cache/cluster_info/data_lister/kubernetes_lister.go:101-117); pods are not
    # filtered, the ones of other partitions simply find no node
    part = params.get("partitionParams") or {}
    pool_key, pool_value = part.get("NodePoolLabelKey") or "", part.get("NodePoolLabelValue") or ""

    def in_partition(obj) -> bool:
        if not pool_key:
            return True
        labels = obj["metadata"].get("labels") or {}
        return labels.get(pool_key) == pool_value if pool_value else pool_key not in labels

    # ---- nodes (cluster_info.go:230-260, node_info.go:107-152) ----
    nodes = sorted((n for n in raw.get("nodes") or [] if in_partition(n)), key=lambda n: n["metadata"]["name"].encode())
    node_names = [n["metadata"]["name"] for n in nodes]
    nindex = {n: i for i, n in enumerate(node_names)}
    if len(nindex) != len(nodes):
        raise ValueError("duplicate node names")
    N = len(nodes)

    # ---- bind requests (cluster_info.go:328-349): only those naming a known node ----
    bind_requests = {}
    for br in raw.get("bindRequests") or []:
        spec = br.get("spec") or {}
        if spec.get("selectedNode") in nindex:
            if spec.get("selectedGPUGroups") or (spec.get("receivedResourceType") or "Regular") not in ("Regular", ""):
                raise UnsupportedSnapshot("bind request for a shared GPU")
            bind_requests[(br["metadata"].get("namespace", ""), spec.get("podName"))] = br

    # ---- pods (cluster_info.go:294-326,460-490) ----
    pods = raw.get("pods") or []
    pod_rows = []
    for pod in pods:
        md = pod["metadata"]
        br = bind_requests.get((md.get("namespace", ""), md["name"]))
        if br is not None and _bind_request_failed(br):  # BindRequestMap.GetBindRequestForPod (binrequest_info.go:16-27)
            br = None
        status = _task_status(pod, br)
        node_name = (pod.get("spec") or {}).get("nodeName") or (br["spec"]["selectedNode"] if br else "")
        pod_rows.append(dict(pod=pod, status=status, node=nindex.get(node_name, -1), req=_pod_request(pod, names),
                             group=(md.get("annotations") or {}).get(POD_GROUP_ANNOTATION, ""),
                             cons=_pod_constraints(pod, strict, ignored)))
    node_alloc = [_resource_list((n.get("status") or {}).get("allocatable"), names, False, f"node {n['metadata']['name']}")
                  for n in nodes]
    R = len(names.names)
    if R > abi.KAI_MAX_RES:
        raise UnsupportedSnapshot(f"{R} resource dimensions (max {abi.KAI_MAX_RES})")

    alloc = np.zeros((R, N))
    for n, rl in enumerate(node_alloc):
        for k, v in rl.items():
            alloc[k, n] = float(v)
    # populateDRAGPUs (cluster_info.go:262-292): GPUs published through DRA ResourceSlices of a GPU driver are added to
    # the node (NodeInfo.AddDRAGPUs) and mark it as a DRA node
    dra_node = np.zeros(N, dtype=bool)
    for sl in raw.get("resourceSlices") or []:
        spec = sl.get("spec") or {}
        if spec.get("allNodes") or "gpu" not in str(spec.get("driver", "")).lower():  # resources.IsGPUDeviceClass
            continue
        n = nindex.get(spec.get("nodeName") or "", -1)
        if n >= 0 and spec.get("devices"):
            alloc[2, n] += float(len(spec["devices"]))
            dra_node[n] = True
    idle = alloc.copy()
    rel = np.zeros((R, N))
    foreign = np.zeros((3, N))
    for row in pod_rows:  # NodeInfo.AddTasksToNode / addTaskResources (node_info.go:417-436,457-493)
        n = row["node"]
        if n < 0 or not (row["status"] & ACTIVE_USED):
            continue
        vec = np.zeros(R)
        for k, v in row["req"].items():
            vec[k] = float(v)
        idle[:, n] -= vec
        if row["status"] == abi.POD_RELEASING:
            rel[:, n] += vec
        sched = (row["pod"].get("spec") or {}).get("schedulerName", "")
        if sched != scheduler_name and not _is_kai_utility_pod(row["pod"]):  # proportion.go:276-286
            foreign[0, n] += vec[0]
            foreign[1, n] += vec[1]
            foreign[2, n] += vec[2]
    ready = np.array([_node_ready(n) for n in nodes], dtype=bool)
    flags = (np.where(ready, abi.NODE_READY, 0) | np.where(dra_node, abi.NODE_NOT_CPU_ONLY, 0)).astype(np.uint32)
    gpu_count = alloc[2].copy() if N else np.zeros(0)
    for n, node in enumerate(nodes):  # GetNumberOfGPUsInNode (node_info.go:630-651)
        lv = (node["metadata"].get("labels") or {}).get(GPU_COUNT_LABEL)
        if lv is not None:
            try:
                gpu_count[n] = float(int(lv))
            except ValueError:
                pass

    # ---- queues (cache/cluster_info/queue.go:56-140) ----
    queues_raw = {q["metadata"]["name"]: q for q in raw.get("queues") or [] if in_partition(q)}
    qrows = {}
    if params.get("fullHierarchyFairness"):
        for name, q in queues_raw.items():
            qrows[name] = dict(q.get("spec") or {}, _created=_epoch(q["metadata"].get("creationTimestamp")))
    else:  # ProjectLevelFairness: one synthetic parent, top-level queues dropped
        unl = {"quota": -1, "overQuotaWeight": 1, "limit": -1}
        qrows["default"] = {"resources": {"gpu": unl, "cpu": unl, "memory": unl}, "_created": 2 ** 62}
        for name, q in queues_raw.items():
            spec = dict(q.get("spec") or {})
            if spec.get("parentQueue"):
                spec["parentQueue"] = "default"
                qrows[name] = dict(spec, _created=_epoch(q["metadata"].get("creationTimestamp")))
    changed = True
    while changed:  # cleanQueueOrphans: a queue whose parent is missing goes, with its subtree
        changed = False
        for name in list(qrows):
            p = qrows[name].get("parentQueue") or ""
            if p and p not in qrows:
                del qrows[name]
                changed = True
    queue_names = sorted(qrows, key=lambda s: s.encode())
    qindex = {n: i for i, n in enumerate(queue_names)}
    Q = len(queue_names)
    q_parent = np.full(Q, -1, dtype=np.int32)
    q_prio = np.zeros(Q, dtype=np.int32)
    q_created = np.zeros(Q, dtype=np.int64)
    q_des, q_lim, q_oqw = np.zeros((3, Q)), np.zeros((3, Q)), np.zeros((3, Q))
    q_pre_mrt, q_rec_mrt = np.full(Q, -1.0), np.full(Q, -1.0)
    for i, name in enumerate(queue_names):
        spec = qrows[name]
        if spec.get("parentQueue"):
            q_parent[i] = qindex[spec["parentQueue"]]
        q_prio[i] = spec["priority"] if spec.get("priority") is not None else DEFAULT_QUEUE_PRIORITY
        for arr, key in ((q_pre_mrt, "preemptMinRuntime"), (q_rec_mrt, "reclaimMinRuntime")):
            if spec.get(key) is not None:  # *metav1.Duration (queue_types.go:40-46)
                arr[i] = parse_duration(spec[key])
        q_created[i] = spec["_created"]
        res = spec.get("resources") or {}
        for r, key, scale in ((0, "cpu", 1.0), (1, "memory", MEGA), (2, "gpu", 1.0)):
            e = res.get(key) or {}
            quota, limit = float(e.get("quota", 0) or 0), float(e.get("limit", 0) or 0)
            q_des[r, i] = max(-1.0, quota * scale) if scale != 1.0 else quota
            q_lim[r, i] = max(-1.0, limit * scale) if scale != 1.0 else limit
            q_oqw[r, i] = float(e.get("overQuotaWeight", 0) or 0)
    q_uid_rank = np.arange(Q, dtype=np.int32)  # queue_names is sorted byte-wise already

    # ---- priority classes (cluster_info.go:500-534) ----
    prio_classes = {pc["metadata"]["name"]: int(pc.get("value", 0)) for pc in raw.get("priorityClasses") or []}
    default_prio = DEFAULT_PODGROUP_PRIORITY
    for pc in raw.get("priorityClasses") or []:
        if pc.get("globalDefault"):
            default_prio = int(pc.get("value", 0))
            break

    # ---- topologies ----
    topologies = raw.get("topologies") or []
    tnames = [t["metadata"]["name"] for t in topologies]
    level_labels = [[lv["nodeLabel"] for lv in (t.get("spec") or {}).get("levels") or []] for t in topologies]
    level_begin = np.zeros(len(topologies) + 1, dtype=np.int32)
    for k, labels in enumerate(level_labels):
        level_begin[k + 1] = level_begin[k] + len(labels)
    node_domain = np.full((int(level_begin[-1]), N), -1, dtype=np.int32)
    for k, labels in enumerate(level_labels):  # DomainID = label values joined by "." (topology_structs.go:76-82)
        for li in range(len(labels)):
            ids = {}
            for n, node in enumerate(nodes):
                nl = node["metadata"].get("labels") or {}
                if all(lb in nl for lb in labels[:li + 1]):
                    ids[n] = ".".join(nl[lb] for lb in labels[:li + 1])
            order = {d: i for i, d in enumerate(sorted(set(ids.values()), key=lambda s: s.encode()))}
            for n, d in ids.items():
                node_domain[level_begin[k] + li, n] = order[d]

    def constraint(tc):
        if not tc or not tc.get("topology"):
            return (-1, -1, -1)
        if tc["topology"] not in tnames:
            return (-2, -1, -1)  # "Requested topology does not exist" (plugins/topology/job_filtering.go:41-47)
        k = tnames.index(tc["topology"])
        lv = []
        for key in ("requiredTopologyLevel", "preferredTopologyLevel"):
            lv.append((level_labels[k].index(tc[key]) if tc[key] in level_labels[k] else -2) if tc.get(key) else -1)
        return (k, lv[0], lv[1])

    # ---- pod groups (cluster_info.go:351-412; job_info.go:160-216; subgroup_info/factory.go:16-120) ----
    by_group: dict = {}
    for row in pod_rows:
        if row["group"]:
            by_group.setdefault(row["group"], []).append(row)
    def up_for_scheduler(pg) -> bool:
        """isPodGroupUpForScheduler (cluster_info.go:585-602): a pod group with a scheduling backoff that this node
        pool already marked unschedulable (its last SchedulingCondition names the pool) is left out."""
        backoff = (pg.get("spec") or {}).get("schedulingBackoff")
        if backoff is None or int(backoff) == -1:  # utils.NoSchedulingBackoff, also the default
            return True
        last, last_id = None, None
        for cond in (pg.get("status") or {}).get("schedulingConditions") or []:  # utils.GetLastSchedulingCondition
            try:
                cid = int(cond.get("transitionID", ""))
            except ValueError:
                cid = -1
            if last is None or cid > last_id:
                last, last_id = cond, cid
        if last is None:
            return True
        pool = (pg["metadata"].get("labels") or {}).get(pool_key) or "default"  # utils.GetNodePoolNameFromLabels
        return (last.get("nodePool") or "") != pool

    pgs = sorted((g for g in raw.get("podGroups") or [] if in_partition(g) and up_for_scheduler(g)),
                 key=lambda g: g["metadata"]["name"].encode())
    job_names, job_queue, job_prio, job_flags, job_created, job_last_start, job_stale_since = [], [], [], [], [], [], []
    job_podset_begin, podset_min, podset_task_begin = [0], [], [0]
    t_status, t_node, t_req, t_rank, t_names, t_uids, t_job, t_cons, t_nominated = [], [], [], [], [], [], [], [], []
    job_sgs_begin, sgs_parent, sgs_names, sgs_con, ps_sgs, ps_con = [0], [], [], [], [], []
    sig_keys = []
    for pg in pgs:
        spec = pg.get("spec") or {}
        name = pg["metadata"]["name"]
        ji = len(job_names)
        job_names.append(name)
        qi = qindex.get(spec.get("queue", ""), -1)
        job_queue.append(qi)
        if qi >= 0:
            prio = prio_classes.get(spec.get("priorityClassName", ""), default_prio)
            pre = spec.get("preemptibility") or ""
            preemptible = pre == "preemptible" or (pre != "non-preemptible" and prio < NON_PREEMPTIBLE_THRESHOLD)
        else:  # queue validation failed: priority and preemptibility stay at their zero values
            prio, preemptible = 0, False
        job_prio.append(prio)
        job_flags.append(abi.JOB_PREEMPTIBLE if preemptible else 0)
        job_created.append(_epoch(pg["metadata"].get("creationTimestamp")))
        started = (pg["metadata"].get("annotations") or {}).get(LAST_START_ANNOTATION)
        try:  # job_info.go:185-193: an unparsable timestamp is ignored
            job_last_start.append(float(_epoch(started)) if started else -1.0)
        except ValueError:
            job_last_start.append(-1.0)
        stale = (pg["metadata"].get("annotations") or {}).get(STALE_ANNOTATION)
        try:  # job_info.go:174-182 StalenessInfo.TimeStamp, an unparsable value is ignored
            job_stale_since.append(float(_epoch(stale)) if stale else -1.0)
        except ValueError:
            job_stale_since.append(-1.0)

        # SubGroup tree: entries with children are sets, the others PodSets
        subgroups = spec.get("subGroups") or []
        sg_by_name = {}
        for sg in subgroups:
            if sg["name"] in sg_by_name:
                raise ValueError(f"podgroup {name}: subgroup {sg['name']} already exists")
            sg_by_name[sg["name"]] = sg
        children = {}
        for sg in subgroups:
            children.setdefault(sg.get("parent") or "", []).append(sg["name"])
        root_gi = len(sgs_parent)
        sgs_parent.append(-1)
        sgs_names.append("")
        sgs_con.append(constraint(spec.get("topologyConstraint")))
        set_index = {"": root_gi}

        def add_set(sname):
            if sname in set_index:
                return set_index[sname]
            sg = sg_by_name[sname]
            parent = sg.get("parent") or ""
            if parent and (parent not in sg_by_name or parent not in children):
                raise ValueError(f"podgroup {name}: parent {parent} of {sname} not found")
            pgi = add_set(parent)
            set_index[sname] = len(sgs_parent)
            sgs_parent.append(pgi)
            sgs_names.append(sname)
            sgs_con.append(constraint(sg.get("topologyConstraint")))
            return set_index[sname]

        podsets = []  # (name, minAvailable, set index, constraint)
        for sg in subgroups:
            if sg["name"] in children:
                add_set(sg["name"])
        for sg in subgroups:
            if sg["name"] not in children:
                parent = sg.get("parent") or ""
                if parent and parent not in set_index:
                    raise ValueError(f"podgroup {name}: parent {parent} of {sg['name']} not found")
                podsets.append((sg["name"], max(int(sg.get("minMember", 0) or 0), 1), set_index[parent],
                                constraint(sg.get("topologyConstraint"))))
        if not podsets:
            podsets.append((DEFAULT_SUBGROUP, max(int(spec.get("minMember", 0) or 0), 1), root_gi, (-1, -1, -1)))
        podsets.sort(key=lambda p: p[0].encode())
        job_sgs_begin.append(len(sgs_parent))

        rows = by_group.get(name, [])

        def task_key(row):  # TaskOrderFn (plugins/taskorder/task_order.go:28-63) then creation, UID
            md = row["pod"]["metadata"]
            lab = (md.get("labels") or {}).get(TASK_ORDER_LABEL)
            try:
                pr = (0, -int(lab)) if lab is not None else (1, 0)
            except ValueError:
                pr = (0, float("inf"))
            return pr + (_epoch(md.get("creationTimestamp")), str(md.get("uid") or md["name"]).encode())

        ranked = sorted(rows, key=task_key)
        rank_of = {id(r): i for i, r in enumerate(ranked)}
        ps_sigs = []
        for ps_name, ps_min, ps_set, ps_c in podsets:
            podset_min.append(ps_min)
            ps_sgs.append(ps_set)
            ps_con.append(ps_c)
            member_sigs, n_members, n_active = [], 0, 0
            for row in ranked:
                md = row["pod"]["metadata"]
                if ((md.get("labels") or {}).get(SUBGROUP_LABEL) or DEFAULT_SUBGROUP) != ps_name:
                    continue
                n_members += 1
                st = row["status"]
                t_status.append(st)
                t_node.append(row["node"] if st & ACTIVE_USED else -1)
                vec = [0.0] * R
                for k, v in row["req"].items():
                    vec[k] = float(v)
                t_req.append(vec)
                t_rank.append(rank_of[id(row)])
                t_names.append(md["name"])
                t_uids.append(str(md.get("uid") or md["name"]))
                t_job.append(ji)
                t_cons.append(row["cons"])
                t_nominated.append(nindex.get((row["pod"].get("status") or {}).get("nominatedNodeName", ""), -1))
                if st & ACTIVE_ALLOCATED:
                    n_active += 1
                else:
                    member_sigs.append(_pod_signature(row["pod"]))
            podset_task_begin.append(len(t_status))
            if n_members == n_active:
                ps_sigs.append("")  # subgroup_info/podset.go:155-159
            else:
                chain, g = [repr(ps_c)], ps_set
                while g >= 0:
                    chain.append(repr(sgs_con[g]))
                    g = sgs_parent[g]
                ps_sigs.append(json.dumps(["|".join(chain), sorted(member_sigs)]))
        job_podset_begin.append(len(podset_min))
        sig_keys.append(json.dumps(sorted(ps_sigs)))

    J, T = len(job_names), len(t_status)
    order = sorted(range(J), key=lambda i: (job_created[i], job_names[i].encode()))
    job_order_rank = np.zeros(J, dtype=np.int32)
    for r_, i in enumerate(order):
        job_order_rank[i] = r_
    sig_ids = {}
    job_signature = np.array([sig_ids.setdefault(k, len(sig_ids)) for k in sig_keys], dtype=np.int32).reshape(J)
    sgs_name_rank = np.zeros(len(sgs_parent), dtype=np.int32)
    for j in range(J):
        b, e = job_sgs_begin[j], job_sgs_begin[j + 1]
        for r_, g in enumerate(sorted(range(b, e), key=lambda g: sgs_names[g].encode())):
            sgs_name_rank[g] = r_

    # ---- predicate classes: node conditions + node-local k8s filters ----
    words = (N + 31) // 32
    class_ids, masks = {}, []
    t_class = np.full(T, -1, dtype=np.int32)
    need_mask = not bool(ready.all())
    for t, cons in enumerate(t_cons):
        # PredicateByNodeResourcesType (node_info.go:326-333): a device-plugin GPU request is rejected on a DRA node
        wants_gpu = bool(dra_node.any()) and t_req[t][2] > 0
        key = json.dumps([cons, wants_gpu], sort_keys=True)
        if key not in class_ids:
            fits = np.array([bool(ready[n]) and not (wants_gpu and dra_node[n]) and _pod_fits_node(cons, nodes[n])
                             for n in range(N)], dtype=bool)
            if fits.all() and not need_mask:
                class_ids[key] = -1
            else:
                bits = np.zeros(words * 32, dtype=np.uint32)
                bits[:N] = fits
                masks.append((bits.reshape(words, 32) << np.arange(32, dtype=np.uint32)).sum(axis=1).astype(np.uint32))
                class_ids[key] = len(masks) - 1
        t_class[t] = class_ids[key]

    kw = {}
    if masks:
        kw.update(task_pred_class=t_class, pred_mask=np.stack(masks).astype(np.uint32))
    if any(v >= 0 for v in t_nominated):
        kw["task_nominated"] = np.array(t_nominated, dtype=np.int32)
    if foreign.any():
        kw["node_foreign"] = foreign
    if not np.array_equal(gpu_count, alloc[2]):
        kw["node_gpu_count"] = gpu_count
    if len(topologies):
        kw.update(topology_level_begin=level_begin, node_domain=node_domain)
    has_tree = len(topologies) or any(job_sgs_begin[j + 1] - job_sgs_begin[j] > 1 for j in range(J))
    if has_tree:
        kw.update(job_sgs_begin=np.array(job_sgs_begin, dtype=np.int32), sgs_parent=np.array(sgs_parent, dtype=np.int32),
                  sgs_name_rank=sgs_name_rank,
                  sgs_topology=np.array([c[0] for c in sgs_con], dtype=np.int32).reshape(-1),
                  sgs_required_level=np.array([c[1] for c in sgs_con], dtype=np.int32).reshape(-1),
                  sgs_preferred_level=np.array([c[2] for c in sgs_con], dtype=np.int32).reshape(-1),
                  podset_sgs=np.array(ps_sgs, dtype=np.int32).reshape(-1),
                  podset_topology=np.array([c[0] for c in ps_con], dtype=np.int32).reshape(-1),
                  podset_required_level=np.array([c[1] for c in ps_con], dtype=np.int32).reshape(-1),
                  podset_preferred_level=np.array([c[2] for c in ps_con], dtype=np.int32).reshape(-1))
    if (q_pre_mrt >= 0).any() or (q_rec_mrt >= 0).any() or any(v > 0 for v in job_last_start + job_stale_since):
        import time
        captured = doc.get("capturedAt")
        kw.update(queue_preempt_min_runtime_s=q_pre_mrt, queue_reclaim_min_runtime_s=q_rec_mrt,
                  job_last_start_s=np.array(job_last_start, dtype=np.float64).reshape(J),
                  now_s=float(_epoch(captured)) if captured else (float(now) if now is not None else time.time()))
        if any(v > 0 for v in job_stale_since):
            kw["job_stale_since_s"] = np.array(job_stale_since, dtype=np.float64).reshape(J)
    usage = _queue_usage(doc, queue_names)
    if usage is not None:
        kw["queue_usage"] = usage

    snap = abi.Snapshot(
        n_res=R, node_allocatable=alloc, node_idle=idle, node_releasing=rel,
        node_name_rank=np.arange(N, dtype=np.int32), node_flags=flags,
        queue_parent=q_parent, queue_priority=q_prio, queue_creation=q_created, queue_uid_rank=q_uid_rank,
        queue_deserved=q_des, queue_limit=q_lim, queue_oqw=q_oqw,
        job_queue=np.array(job_queue, dtype=np.int32).reshape(J), job_priority=np.array(job_prio, dtype=np.int32).reshape(J),
        job_order_rank=job_order_rank, job_flags=np.array(job_flags, dtype=np.uint32).reshape(J),
        job_podset_begin=np.array(job_podset_begin, dtype=np.int32),
        podset_min_available=np.array(podset_min, dtype=np.int32).reshape(-1),
        podset_task_begin=np.array(podset_task_begin, dtype=np.int32),
        task_status=np.array(t_status, dtype=np.int32).reshape(T), task_node=np.array(t_node, dtype=np.int32).reshape(T),
        task_req=np.array(t_req, dtype=np.float64).reshape(T, R), task_order_rank=np.array(t_rank, dtype=np.int32).reshape(T),
        job_signature=job_signature, **kw)
    meta = {"node_names": node_names, "queue_names": queue_names, "job_names": job_names, "task_names": t_names,
            "task_uids": t_uids, "task_job": np.array(t_job, dtype=np.int32).reshape(T),
            "resource_names": list(names.names), "ignored": ignored}
    cfg_kw, actions = _parse_config(doc)
    return snap, meta, cfg_kw, actions


def _queue_usage(doc: dict, queue_names):
    """Optional `queueUsage: {queue: {cpu, memory, gpu}}` (normalised historical usage, the ClusterUsage the usage DB
    would return — cache/usagedb): not part of the reference's snapshot.json, accepted as an extension."""
    u = doc.get("queueUsage")
    if not u:
        return None
    out = np.zeros((3, len(queue_names)))
    for i, name in enumerate(queue_names):
        e = u.get(name) or {}
        out[0, i], out[1, i], out[2, i] = float(e.get("cpu", 0)), float(e.get("memory", 0)), float(e.get("gpu", 0))
    return out


# ---------------------------------------------------------------- dump
_BASE_EPOCH = 1_700_000_000


def dump_cluster(snap: "abi.Snapshot", actions=("allocate",), config: dict | None = None, names: dict | None = None) -> dict:
    """SoA snapshot -> snapshot.json document (the inverse of pack_cluster up to names).

    Names come from `names` ({"node_names", "queue_names", "job_names", "task_names"}) or are generated so that every
    ordering the scheduler derives from strings (node name, queue UID, PodSet and SubGroupSet names) equals the rank
    arrays of the snapshot; creation timestamps are generated from the order ranks the same way.  Session-only pod
    statuses (Allocated, Pipelined), predicate classes, foreign-pod rows and extra resource dimensions cannot be
    expressed as raw objects without more information and raise UnsupportedSnapshot.
    """
    names = names or {}
    config = dict(config or {})
    N, Q, J, T, R = snap.n_nodes, snap.n_queues, snap.n_jobs, snap.n_tasks, snap.n_res
    if R != 4:
        raise UnsupportedSnapshot("extra resource dimensions need their names")
    if snap.pred_mask is not None or snap.node_foreign is not None:
        raise UnsupportedSnapshot("predicate classes / foreign pods have no raw-object form here")
    wn = max(1, len(str(max(N - 1, 0))))
    node_names = names.get("node_names") or [f"node-{int(r):0{wn}d}" for r in snap.node_name_rank]
    wq = max(1, len(str(max(Q - 1, 0))))
    queue_names = names.get("queue_names") or [f"queue-{int(r):0{wq}d}" for r in snap.queue_uid_rank]
    wj = max(1, len(str(max(J - 1, 0))))
    job_names = names.get("job_names") or [f"job-{j:0{wj}d}" for j in range(J)]
    task_names = names.get("task_names")

    # topologies: label value of level l = zero-padded dense domain id, which keeps the DomainID order
    topologies, level_keys = [], []
    if snap.topology_level_begin is not None:
        lb = snap.topology_level_begin
        for k in range(len(lb) - 1):
            keys = [f"kai.topology/t{k}-l{l}" for l in range(int(lb[k + 1] - lb[k]))]
            level_keys.append(keys)
            topologies.append({"metadata": {"name": f"topology-{k}"}, "spec": {"levels": [{"nodeLabel": x} for x in keys]}})
    nodes = []
    for n in range(N):
        labels = {}
        if snap.node_gpu_count is not None and snap.node_gpu_count[n] != snap.node_allocatable[2, n]:
            labels[GPU_COUNT_LABEL] = _format_int(snap.node_gpu_count[n])
        for k, keys in enumerate(level_keys):
            for l, key in enumerate(keys):
                d = int(snap.node_domain[snap.topology_level_begin[k] + l, n])
                if d >= 0:
                    labels[key] = f"{d:08d}"
        a = snap.node_allocatable[:, n]
        node = {"metadata": {"name": node_names[n], "labels": labels},
                "spec": {},
                "status": {"allocatable": {"cpu": _format_int(a[0]) + "m", "memory": _format_int(a[1]),
                                           "nvidia.com/gpu": _format_int(a[2]), "pods": _format_int(a[3])},
                           "conditions": [{"type": "Ready", "status": "True" if snap.node_flags[n] & abi.NODE_READY else "False"}]}}
        nodes.append(node)

    queues = []
    for q in range(Q):
        res = {}
        for r, key, scale in ((0, "cpu", 1.0), (1, "memory", MEGA), (2, "gpu", 1.0)):
            d, l = float(snap.queue_deserved[r, q]), float(snap.queue_limit[r, q])
            res[key] = {"quota": d / scale if (scale != 1.0 and d >= 0) else d,
                        "limit": l / scale if (scale != 1.0 and l >= 0) else l,
                        "overQuotaWeight": float(snap.queue_oqw[r, q])}
        spec = {"resources": res, "priority": int(snap.queue_priority[q])}
        for arr, key in ((snap.queue_preempt_min_runtime_s, "preemptMinRuntime"), (snap.queue_reclaim_min_runtime_s, "reclaimMinRuntime")):
            if arr is not None and arr[q] >= 0:
                spec[key] = format_duration(float(arr[q]))
        if snap.queue_parent[q] >= 0:
            spec["parentQueue"] = queue_names[int(snap.queue_parent[q])]
        queues.append({"metadata": {"name": queue_names[q], "creationTimestamp": _rfc3339(_BASE_EPOCH + int(snap.queue_creation[q]))},
                       "spec": spec})

    prios = sorted({int(p) for p in snap.job_priority})
    priority_classes = [{"metadata": {"name": f"priority-{p}"}, "value": p} for p in prios]

    def con_obj(topo, req, pref):
        if topo < 0:
            return None
        keys = level_keys[int(topo)]
        out = {"topology": f"topology-{int(topo)}"}
        if req >= 0:
            out["requiredTopologyLevel"] = keys[int(req)]
        if pref >= 0:
            out["preferredTopologyLevel"] = keys[int(pref)]
        return out

    pod_groups, pods, bind_requests = [], [], []
    status_names = {v: k for k, v in abi.POD_STATUS_NAMES.items()}
    for j in range(J):
        b, e = int(snap.job_podset_begin[j]), int(snap.job_podset_begin[j + 1])
        spec = {"queue": queue_names[int(snap.job_queue[j])] if snap.job_queue[j] >= 0 else "missing-queue",
                "priorityClassName": f"priority-{int(snap.job_priority[j])}",
                "preemptibility": "preemptible" if snap.job_flags[j] & abi.JOB_PREEMPTIBLE else "non-preemptible"}
        ps_names = {}
        tree = snap.job_sgs_begin is not None
        if tree:
            gb, ge = int(snap.job_sgs_begin[j]), int(snap.job_sgs_begin[j + 1])
            c = con_obj(snap.sgs_topology[gb], snap.sgs_required_level[gb], snap.sgs_preferred_level[gb])
        else:
            gb = ge = 0
            c = con_obj(snap.job_topology[j], snap.job_required_level[j], snap.job_preferred_level[j]) \
                if snap.job_topology is not None else None
        if c:
            spec["topologyConstraint"] = c
        single_default = (e - b == 1) and (not tree or (ge - gb == 1 and (snap.podset_topology is None or snap.podset_topology[b] < 0)))
        if single_default:
            spec["minMember"] = int(snap.podset_min_available[b])
            ps_names[b] = None
        else:
            subgroups = []
            set_names = {gb: None}
            for g in range(gb + 1, ge):  # nested sets, parents before children
                set_names[g] = f"set-{int(snap.sgs_name_rank[g]):04d}"
            for g in range(gb + 1, ge):
                sg = {"name": set_names[g]}
                if set_names[int(snap.sgs_parent[g])] is not None:
                    sg["parent"] = set_names[int(snap.sgs_parent[g])]
                c = con_obj(snap.sgs_topology[g], snap.sgs_required_level[g], snap.sgs_preferred_level[g])
                if c:
                    sg["topologyConstraint"] = c
                subgroups.append(sg)
            for ps in range(b, e):
                ps_names[ps] = f"podset-{ps - b:04d}"
                sg = {"name": ps_names[ps], "minMember": int(snap.podset_min_available[ps])}
                if tree and set_names.get(int(snap.podset_sgs[ps])) is not None:
                    sg["parent"] = set_names[int(snap.podset_sgs[ps])]
                if snap.podset_topology is not None:
                    c = con_obj(snap.podset_topology[ps], snap.podset_required_level[ps], snap.podset_preferred_level[ps])
                    if c:
                        sg["topologyConstraint"] = c
                subgroups.append(sg)
            have_children = {sg.get("parent") for sg in subgroups}
            for sg in subgroups:  # a set without children would be read back as a PodSet
                if sg["name"].startswith("set-") and sg["name"] not in have_children:
                    raise UnsupportedSnapshot("empty SubGroupSet")
            spec["subGroups"] = subgroups
        pg_md = {"name": job_names[j], "namespace": "default",
                 "creationTimestamp": _rfc3339(_BASE_EPOCH + int(snap.job_order_rank[j]))}
        if snap.job_last_start_s is not None and snap.job_last_start_s[j] > 0:
            if snap.job_last_start_s[j] != int(snap.job_last_start_s[j]):
                raise UnsupportedSnapshot("sub-second last-start timestamps (RFC 3339 annotation)")
            pg_md["annotations"] = {LAST_START_ANNOTATION: _rfc3339(int(snap.job_last_start_s[j]))}
        if snap.job_stale_since_s is not None and snap.job_stale_since_s[j] > 0:
            if snap.job_stale_since_s[j] != int(snap.job_stale_since_s[j]):
                raise UnsupportedSnapshot("sub-second staleness timestamps (RFC 3339 annotation)")
            pg_md.setdefault("annotations", {})[STALE_ANNOTATION] = _rfc3339(int(snap.job_stale_since_s[j]))
        pod_groups.append({"metadata": pg_md, "spec": spec})
        for ps in range(b, e):
            for t in range(int(snap.podset_task_begin[ps]), int(snap.podset_task_begin[ps + 1])):
                st = int(snap.task_status[t])
                tname = task_names[t] if task_names else f"{job_names[j]}-{t - int(snap.podset_task_begin[b]):04d}"
                req = snap.task_req[t]
                requests = {}
                if req[0]:
                    requests["cpu"] = _format_int(req[0]) + "m"
                if req[1]:
                    requests["memory"] = _format_int(req[1])
                if req[2]:
                    requests["nvidia.com/gpu"] = _format_int(req[2])
                md = {"name": tname, "namespace": "default", "uid": tname,
                      "annotations": {POD_GROUP_ANNOTATION: job_names[j]}, "labels": {},
                      "creationTimestamp": _rfc3339(_BASE_EPOCH + int(snap.task_order_rank[t]))}
                if ps_names[ps] is not None:
                    md["labels"][SUBGROUP_LABEL] = ps_names[ps]
                pspec = {"schedulerName": config.get("scheduler_name", DEFAULT_SCHEDULER_NAME),
                         "containers": [{"name": "main", "resources": {"requests": requests}}]}
                pstat = {}
                node = node_names[int(snap.task_node[t])] if snap.task_node[t] >= 0 else ""
                if st == abi.POD_PENDING:
                    pstat["phase"] = "Pending"
                elif st == abi.POD_GATED:
                    pstat["phase"] = "Pending"
                    pspec["schedulingGates"] = [{"name": "kai.scheduler/gate"}]
                elif st == abi.POD_RUNNING:
                    pstat["phase"] = "Running"
                    pspec["nodeName"] = node
                elif st == abi.POD_RELEASING:
                    pstat["phase"] = "Running"
                    pspec["nodeName"] = node
                    md["deletionTimestamp"] = _rfc3339(_BASE_EPOCH)
                elif st == abi.POD_BOUND:
                    pstat["phase"] = "Pending"
                    pspec["nodeName"] = node
                elif st == abi.POD_BINDING:
                    pstat["phase"] = "Pending"
                    bind_requests.append({"metadata": {"name": tname, "namespace": "default"},
                                          "spec": {"podName": tname, "selectedNode": node, "receivedResourceType": "Regular"}})
                elif st in (abi.POD_SUCCEEDED, abi.POD_FAILED, abi.POD_UNKNOWN):
                    pstat["phase"] = status_names[st]
                else:
                    raise UnsupportedSnapshot(f"task status {status_names.get(st, st)} exists only inside a session")
                if snap.task_nominated is not None and snap.task_nominated[t] >= 0:
                    pstat["nominatedNodeName"] = node_names[int(snap.task_nominated[t])]
                pods.append({"metadata": md, "spec": pspec, "status": pstat})

    place = {abi.PLACEMENT_BINPACK: "binpack", abi.PLACEMENT_SPREAD: "spread"}
    # the default tier (conf_util/scheduler_conf_util.go:38-60) without the workload-specific and HTTP plugins
    plugins = [{"name": n} for n in ("predicates", "proportion", "priority", "elastic", "nodeavailability",
                                     "gpusharingorder", "gpupack", "resourcetype", "subgrouporder", "taskorder",
                                     "nominatednode", "nodeplacement", "minruntime", "topology")]
    for p in plugins:
        if p["name"] == "nodeplacement":
            p["arguments"] = {"gpu": place[config.get("gpu_placement", abi.PLACEMENT_BINPACK)],
                              "cpu": place[config.get("cpu_placement", abi.PLACEMENT_BINPACK)]}
        if p["name"] == "minruntime":
            p["arguments"] = {"defaultReclaimMinRuntime": format_duration(config.get("default_reclaim_min_runtime_s", 0.0)),
                              "defaultPreemptMinRuntime": format_duration(config.get("default_preempt_min_runtime_s", 0.0)),
                              "reclaimResolveMethod": "queue" if config.get("reclaim_resolve_method") == abi.RESOLVE_QUEUE else "lca"}
        if p["name"] == "proportion" and ("k_value" in config or "saturation_multiplier" in config):
            p["arguments"] = {"kValue": str(config.get("k_value", 1.0)),
                              "relcaimerSaturationMultiplier": str(config.get("saturation_multiplier", 1.0))}
    params = {"schedulerName": config.get("scheduler_name", DEFAULT_SCHEDULER_NAME), "fullHierarchyFairness": True,
              "useSchedulingSignatures": bool(config.get("use_scheduling_signatures", False)),
              "allowConsolidatingReclaim": bool(config.get("allow_consolidating_reclaim", True))}
    params["maxNumberConsolidationPreemptees"] = int(config.get("max_consolidation_preemptees", -1))
    grace = config.get("staleness_grace_period_s", 0)
    if grace:
        params["globalDefaultStalenessGracePeriod"] = int(grace) * 10 ** 9 if grace > 0 else -1
    doc = {"config": {"actions": ", ".join(actions), "tiers": [{"plugins": plugins}]},
           "schedulerParams": params,
           "rawObjects": {"pods": pods, "nodes": nodes, "queues": queues, "podGroups": pod_groups,
                          "bindRequests": bind_requests, "priorityClasses": priority_classes, "topologies": topologies}}
    if snap.job_last_start_s is not None or snap.queue_preempt_min_runtime_s is not None or snap.job_stale_since_s is not None:
        if snap.now_s != int(snap.now_s):
            raise UnsupportedSnapshot("sub-second capture time")
        doc["capturedAt"] = _rfc3339(int(snap.now_s))  # extension: the reference's tool uses time.Now() at replay
    if snap.queue_usage is not None:
        doc["queueUsage"] = {queue_names[q]: {"cpu": float(snap.queue_usage[0, q]), "memory": float(snap.queue_usage[1, q]),
                                              "gpu": float(snap.queue_usage[2, q])} for q in range(Q)}
    return doc
