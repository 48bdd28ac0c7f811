"""The reference's declarative test DSL -> SoA snapshot (test infrastructure).

Mirrors pkg/scheduler/test_utils (reference @72de64fc): a `TestTopologyBasic`
table (transcribed to JSON by tests/golden/gen_fixtures.py) is turned into the
kai_snapshot the engine and the oracle consume, following the same rules as
`test_utils.BuildSession` (test_utils_builder.go:260-289):

  jobs    jobs_fake/jobs.go:53-98     stable sort by Priority desc; job i created now-(n-i) min;
                                      UID = Name; preemptible iff priority < 100 unless explicit
  tasks   jobs_fake/jobs.go:188-291   pod "<job>-<k>", UID = name; cpu default "1" (1000m),
                                      memory default "1G"; pods = 1; best effort = empty request
  nodes   nodes_fake/nodes.go:31-36,171-224   cpu "20000" cores-quantity = 2e7 m, memory "20G",
                                      pods 110 (or MaxTaskNum), running tasks pre-added
  queues  test_utils_builder.go:96-225        queue k created now+k min, CPU/memory quota/limit -1,
                                      default department when none given
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kai_scheduler_b200 import abi  # noqa: E402

ACTIVE_USED = {"Allocated", "Pipelined", "Binding", "Bound", "Running", "Releasing"}
DEFAULT_SUBGROUP = "default-sub-group"


def _parse_quantity_cpu(cores: float) -> float:
    return float(cores) * 1000.0


DSL_NOW_S = 1.7e9  # the session instant of tables that carry durations


def build_snapshot(topo: dict):
    """Return (Snapshot, meta) for a transcribed TestTopologyBasic.

    meta: node_names, job_names (snapshot order), task_names, task_job, queue_names.
    """
    topo = dict(topo)
    queues = [dict(q) for q in (topo.get("Queues") or [])]
    departments = [dict(d) for d in (topo.get("Departments") or [])]
    # test_utils_builder.go:160-177 addDefaultDepartmentIfNeeded
    if not departments and not topo.get("DisableDefaultDepartment"):
        for q in queues:
            q["ParentQueue"] = "default"
        departments = [{"Name": "default", "DeservedGPUs": -1.0, "MaxAllowedGPUs": -1.0}]

    # ---- queues (leaf queues first, then departments; creation = index minutes) ----
    qnames, qparent, qprio, qcreate = [], [], [], []
    qd, ql, qw = [], [], []
    for k, q in enumerate(queues):
        qnames.append(q["Name"])
        qparent.append(q.get("ParentQueue") or "")
        qprio.append(q["Priority"] if q.get("Priority") is not None else 100)
        qcreate.append(k * 60)
        max_gpu = q.get("MaxAllowedGPUs", 0) or 0
        gpu_limit = max_gpu if max_gpu != 0 else -1.0
        cpu_q = q["DeservedCPUs"] if q.get("DeservedCPUs") is not None else -1.0
        mem_q = q["DeservedMemory"] if q.get("DeservedMemory") is not None else -1.0
        cpu_l = q["MaxAllowedCPUs"] if q.get("MaxAllowedCPUs") is not None else -1.0
        mem_l = q["MaxAllowedMemory"] if q.get("MaxAllowedMemory") is not None else -1.0
        # proportion.go:48-50,327-328: memory quota/limit are in MB in the Queue CR
        qd.append([cpu_q, max(-1.0, mem_q * 1e6), q.get("DeservedGPUs", 0) or 0])
        ql.append([cpu_l, max(-1.0, mem_l * 1e6), gpu_limit])
        qw.append([1.0, 1.0, q.get("GPUOverQuotaWeight", 0) or 0])
    for d_i, d in enumerate(departments):
        qnames.append(d["Name"])
        qparent.append(d.get("ParentQueue") or "")
        qprio.append(100)
        qcreate.append(d_i * 60)
        max_gpu = d.get("MaxAllowedGPUs", 0) or 0
        gpu_limit = max_gpu if max_gpu != 0 else -1.0
        cpu_l = d["MaxAllowedCPUs"] if d.get("MaxAllowedCPUs") is not None else -1.0
        mem_l = d["MaxAllowedMemory"] if d.get("MaxAllowedMemory") is not None else -1.0
        dg = d.get("DeservedGPUs", 0) or 0
        qd.append([-1.0, -1.0, dg])
        ql.append([cpu_l, max(-1.0, mem_l * 1e6), gpu_limit])
        qw.append([1.0, 1.0, dg])  # department OverQuotaWeight = DeservedGPUs (:194-198)
    qindex = {n: i for i, n in enumerate(qnames)}
    Q = len(qnames)
    # cache/cluster_info/queue.go:90-129 UpdateQueueHierarchy: orphans (missing parent) are deleted
    parent_idx = []
    for i in range(Q):
        p = qparent[i]
        parent_idx.append(qindex[p] if p in qindex else (-1 if p == "" else -2))
    if any(p == -2 for p in parent_idx):
        raise ValueError("queue with missing parent: unsupported in DSL")
    uid_rank = np.argsort(np.argsort(np.array(qnames, dtype=object), kind="stable"), kind="stable")

    # ---- nodes ----
    node_names = sorted((topo.get("Nodes") or {}).keys())
    nindex = {n: i for i, n in enumerate(node_names)}
    N = len(node_names)
    R = 4
    alloc = np.zeros((R, N))
    for n, name in enumerate(node_names):
        nd = topo["Nodes"][name]
        alloc[0, n] = float(nd["CPUMillis"]) * 1000.0 if (nd.get("CPUMillis") or 0) > 0 else 2e7
        alloc[1, n] = float(nd["CPUMemory"]) if (nd.get("CPUMemory") or 0) > 0 else 2e10
        alloc[2, n] = float(nd.get("GPUs", 0) or 0)
        alloc[3, n] = float(nd["MaxTaskNum"]) if nd.get("MaxTaskNum") is not None else 110.0
    idle = alloc.copy()
    rel = np.zeros((R, N))
    name_rank = np.arange(N, dtype=np.int32)  # node_names is already sorted byte-wise
    flags = np.full(N, abi.NODE_READY, dtype=np.uint32)

    # ---- jobs ----
    jobs = [dict(j) for j in (topo.get("Jobs") or [])]
    jobs.sort(key=lambda j: -(j.get("Priority", 0) or 0))  # stable, like sort.SliceStable
    nj = len(jobs)
    job_names, job_queue, job_prio, job_flags, job_creation = [], [], [], [], []
    job_podset_begin = [0]
    podset_min, podset_task_begin = [], [0]
    t_status, t_node, t_req, t_rank, t_names, t_job = [], [], [], [], [], []
    t_fixture_node = []
    t_affinity = []
    job_podset_names = []
    for ji, job in enumerate(jobs):
        job_names.append(job["Name"])
        job_queue.append(qindex.get(job.get("QueueName", ""), -1))
        prio = int(job.get("Priority", 0) or 0)
        job_prio.append(prio)
        age = job.get("JobAgeInMinutes", 0) or 0
        job_creation.append(-(age if age != 0 else (nj - ji)) * 60)
        pre = job.get("Preemptibility") or ""
        preemptible = pre == "preemptible" or (pre != "non-preemptible" and prio < 100)
        job_flags.append(abi.JOB_PREEMPTIBLE if preemptible else 0)
        tasks = job.get("Tasks") or []
        # requests
        if job.get("IsBestEffortJob"):
            base = [0.0, 0.0, 0.0, 1.0]
        else:
            cpu = _parse_quantity_cpu(job["RequiredCPUsPerTask"]) if job.get("RequiredCPUsPerTask") else 1000.0
            mem = float(job["RequiredMemoryPerTask"]) if job.get("RequiredMemoryPerTask") else 1e9
            base = [cpu, mem, float(job.get("RequiredGPUsPerTask", 0) or 0), 1.0]
        # podsets: jobs_fake.go:117-134
        root = job.get("RootSubGroupSet")
        if root:
            podsets = [(p["name"], int(p["min_available"])) for p in root["podsets"]]
        else:
            podsets = []
        names_in_sets = {p[0] for p in podsets}
        if any(not t.get("SubGroupName") for t in tasks) and DEFAULT_SUBGROUP not in names_in_sets:
            podsets.append((DEFAULT_SUBGROUP, len(tasks)))
        podsets.sort(key=lambda p: p[0])  # PodSet name order (session_plugins.go:261-270 fallback)
        if job.get("JobNotReadyForSsn"):
            # integration tables: job exists but is not ready -> model as gated-like by min_available > tasks
            podsets = [(n_, m_ + len(tasks) + 1) for n_, m_ in podsets]
        # task order: TaskOrderFn = taskorder label priority desc, then creation (equal), then UID (pod name)
        order_keys = []
        for k, t in enumerate(tasks):
            uid = f"{job['Name']}-{k}"
            pr = t.get("Priority")
            order_keys.append((0 if pr is not None else 1, -(pr or 0), uid, k))
        order_rank = {k: r for r, (_, _, _, k) in enumerate(sorted(order_keys))}
        job_podset_names.append([p_[0] for p_ in podsets])
        for ps_name, ps_min in podsets:
            podset_min.append(ps_min)
            for k, t in enumerate(tasks):
                sg = t.get("SubGroupName") or DEFAULT_SUBGROUP
                if sg != ps_name:
                    continue
                st = t.get("State", "Pending")
                t_status.append(abi.POD_STATUS_NAMES[st])
                node = t.get("NodeName") or ""
                t_node.append(nindex[node] if (st in ACTIVE_USED and node in nindex) else -1)
                req = list(base)
                if t.get("RequiredGPUs") is not None and not job.get("IsBestEffortJob"):
                    req[2] = float(t["RequiredGPUs"])
                t_req.append(req)
                t_rank.append(order_rank[k])
                t_names.append(f"{job['Name']}-{k}")
                t_job.append(ji)
                t_affinity.append(tuple(t.get("NodeAffinityNames") or ()))
                # a fixture may name a node on a Pending task; PodInfo.NodeName keeps it and the reference's
                # matcher compares it (test_utils.go:236-244) although it never reaches the algorithm
                t_fixture_node.append(node if t_node[-1] < 0 else "")
            podset_task_begin.append(len(t_status))
        job_podset_begin.append(len(podset_min))
    # job order rank under (CreationTimestamp, UID)
    order = sorted(range(nj), key=lambda i: (job_creation[i], job_names[i]))
    job_order_rank = np.zeros(nj, dtype=np.int32)
    for r, i in enumerate(order):
        job_order_rank[i] = r

    # ---- pre-place active-used tasks on nodes (node_info.go:457-493), sorted pod-key order is irrelevant for sums
    T = len(t_status)
    t_req_a = np.array(t_req, dtype=np.float64).reshape(T, R)
    for t in range(T):
        n = t_node[t]
        if n < 0:
            continue
        st = t_status[t]
        if st == abi.POD_RELEASING:
            rel[:, n] += t_req_a[t]
            idle[:, n] -= t_req_a[t]
        elif st == abi.POD_PIPELINED:
            rel[:, n] -= t_req_a[t]
        else:
            idle[:, n] -= t_req_a[t]
    # nodes_fake.go:103-112: with MaxTaskNum the idle pods are clamped at 0
    for n, name in enumerate(node_names):
        if topo["Nodes"][name].get("MaxTaskNum") is not None and idle[3, n] < 0:
            idle[3, n] = 0

    # ---- Topology CRs -> per-level domain ids (topology_plugin.go:57-110; DomainID = joined label values) ----
    topo_kw = {}
    topologies = topo.get("Topologies") or []
    if topologies:
        tnames = [tp["ObjectMeta"]["Name"] for tp in topologies]
        level_begin = [0]
        level_labels = []
        for tp in topologies:
            labels = [lv["NodeLabel"] for lv in tp["Spec"]["Levels"]]
            level_labels.append(labels)
            level_begin.append(level_begin[-1] + len(labels))
        node_domain = np.full((level_begin[-1], N), -1, dtype=np.int32)
        for k, labels in enumerate(level_labels):
            for li in range(len(labels)):
                ids = {}
                for n, name in enumerate(node_names):
                    nl = topo["Nodes"][name].get("Labels") or {}
                    if any(lb not in nl for lb in labels[:li + 1]):
                        continue
                    ids[n] = ".".join(nl[lb] for lb in labels[:li + 1])
                order = {d: i for i, d in enumerate(sorted(set(ids.values())))}
                for n, d in ids.items():
                    node_domain[level_begin[k] + li, n] = order[d]
        def con(tc):
            """TopologyConstraintInfo -> (topology, required level, preferred level) indices; -2 = unknown name"""
            if not tc or not tc.get("Topology"):
                return (-1, -1, -1)
            if tc["Topology"] not in tnames:
                return (-2, -1, -1)  # "Requested topology does not exist" (job_filtering.go:41-47)
            k = tnames.index(tc["Topology"])
            lv = []
            for key in ("RequiredLevel", "PreferredLevel"):
                lv.append((level_labels[k].index(tc[key]) if tc[key] in level_labels[k] else -2) if tc.get(key) else -1)
            return (k, lv[0], lv[1])

        # SubGroupSet tree per job: sets in pre-order (root first), PodSets attached to their set
        job_sgs_begin, sgs_parent, sgs_names, sgs_con = [0], [], [], []
        ps_sgs = [0] * len(podset_min)
        ps_con = [(-1, -1, -1)] * len(podset_min)
        for ji, job in enumerate(jobs):
            root = job.get("RootSubGroupSet") or {}
            tree = root.get("tree") or {"name": "root", "constraint": root.get("topology_constraint"), "podsets": [], "groups": []}
            ps_names = job_podset_names[ji]
            first_ps = job_podset_begin[ji]
            placed = set()

            def walk(g, parent):
                gi = len(sgs_parent)
                sgs_parent.append(parent)
                sgs_names.append(g["name"])
                sgs_con.append(con(g.get("constraint")))
                for p in g["podsets"]:
                    k = ps_names.index(p["name"])
                    ps_sgs[first_ps + k] = gi
                    ps_con[first_ps + k] = con(p.get("constraint"))
                    placed.add(p["name"])
                for c in g["groups"]:
                    walk(c, gi)
                return gi

            root_gi = walk(tree, -1)
            for k, name in enumerate(ps_names):  # the default PodSet (and flat lists) hang off the root
                if name not in placed:
                    ps_sgs[first_ps + k] = root_gi
            job_sgs_begin.append(len(sgs_parent))
        sgs_name_rank = np.zeros(len(sgs_parent), dtype=np.int32)
        for ji in range(nj):
            b, e = job_sgs_begin[ji], job_sgs_begin[ji + 1]
            order = sorted(range(b, e), key=lambda g: sgs_names[g])
            for r_, g in enumerate(order):
                sgs_name_rank[g] = r_
        topo_kw = dict(topology_level_begin=np.array(level_begin, dtype=np.int32), node_domain=node_domain,
                       job_sgs_begin=np.array(job_sgs_begin, dtype=np.int32), sgs_parent=np.array(sgs_parent, dtype=np.int32),
                       sgs_name_rank=sgs_name_rank, sgs_topology=np.array([c[0] for c in sgs_con], dtype=np.int32),
                       sgs_required_level=np.array([c[1] for c in sgs_con], dtype=np.int32),
                       sgs_preferred_level=np.array([c[2] for c in sgs_con], dtype=np.int32),
                       podset_sgs=np.array(ps_sgs, dtype=np.int32), podset_topology=np.array([c[0] for c in ps_con], dtype=np.int32),
                       podset_required_level=np.array([c[1] for c in ps_con], dtype=np.int32),
                       podset_preferred_level=np.array([c[2] for c in ps_con], dtype=np.int32))

    # tasks_fake/tasks.go:98-116: NodeAffinityNames = required node affinity `kai.scheduler/type In names`; every fake
    # node carries that label with its own name unless the fixture overrides it (nodes_fake/nodes.go:182-191).  A k8s
    # Filter result: handed to the engine as a predicate class.
    pred_kw = {}
    if any(t_affinity):
        def node_type(name):
            labels = topo["Nodes"][name].get("Labels") or {}
            return labels.get("tasks_fake.NodeAffinityKey", labels.get("kai.scheduler/type", name))
        classes, masks = {}, []
        t_class = np.full(T, -1, dtype=np.int32)
        words = (N + 31) // 32
        for t, names in enumerate(t_affinity):
            if not names:
                continue
            if names not in classes:
                bits = np.zeros(words * 32, dtype=np.uint64)
                bits[:N] = [node_type(n) in names for n in node_names]
                masks.append((bits.reshape(words, 32) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32))
                classes[names] = len(masks) - 1
            t_class[t] = classes[names]
        pred_kw = dict(task_pred_class=t_class, pred_mask=np.stack(masks).astype(np.uint32))

    snap = abi.Snapshot(
        n_res=R,
        node_allocatable=alloc, node_idle=idle, node_releasing=rel,
        node_name_rank=name_rank, node_flags=flags,
        queue_parent=np.array(parent_idx, dtype=np.int32), queue_priority=np.array(qprio, dtype=np.int32),
        queue_creation=np.array(qcreate, dtype=np.int64), queue_uid_rank=uid_rank.astype(np.int32),
        queue_deserved=np.array(qd, dtype=np.float64).reshape(Q, 3).T.copy(),
        queue_limit=np.array(ql, dtype=np.float64).reshape(Q, 3).T.copy(),
        queue_oqw=np.array(qw, dtype=np.float64).reshape(Q, 3).T.copy(),
        job_queue=np.array(job_queue, dtype=np.int32), job_priority=np.array(job_prio, dtype=np.int32),
        job_order_rank=job_order_rank, job_flags=np.array(job_flags, dtype=np.uint32),
        job_podset_begin=np.array(job_podset_begin, dtype=np.int32),
        podset_min_available=np.array(podset_min, dtype=np.int32),
        podset_task_begin=np.array(podset_task_begin, dtype=np.int32),
        task_status=np.array(t_status, dtype=np.int32), task_node=np.array(t_node, dtype=np.int32),
        task_req=t_req_a, task_order_rank=np.array(t_rank, dtype=np.int32),
        **topo_kw, **pred_kw,
    )
    if any(j.get("StaleDuration") is not None for j in jobs):
        # jobs_fake/jobs.go:50,92 StaleDuration -> StalenessInfo.TimeStamp = now - duration (seconds in the fixtures)
        snap.now_s = DSL_NOW_S
        snap.job_stale_since_s = np.array([DSL_NOW_S - float(j["StaleDuration"]) if j.get("StaleDuration") is not None else -1.0
                                           for j in jobs], dtype=np.float64)
    meta = {
        "node_names": node_names, "job_names": job_names, "task_names": t_names,
        "task_job": np.array(t_job, dtype=np.int32), "queue_names": qnames, "task_fixture_node": t_fixture_node,
    }
    return snap, meta


def check_expectations(topo: dict, meta: dict, res: "abi.Result", snap: "abi.Snapshot"):
    """MatchExpectedAndRealTasks (test_utils.go:121-314) restated: returns a list of mismatch strings."""
    errs = []
    status_name = {v: k for k, v in abi.POD_STATUS_NAMES.items()}
    tj = meta["task_job"]
    for jname, exp in (topo.get("JobExpectedResults") or {}).items():
        if jname not in meta["job_names"]:
            errs.append(f"job {jname} missing")
            continue
        ji = meta["job_names"].index(jname)
        gpus = 0.0
        for t in np.nonzero(tj == ji)[0]:
            st = status_name[int(res.task_status[t])]
            if st != exp.get("Status"):
                errs.append(f"job {jname} task {meta['task_names'][t]}: status {st}, expected {exp.get('Status')}")
            want = exp.get("NodeName") or ""
            got = meta["node_names"][res.task_node[t]] if res.task_node[t] >= 0 else meta["task_fixture_node"][t]
            if want and got != want:
                errs.append(f"job {jname} task {meta['task_names'][t]}: node {got!r}, expected {want!r}")
            gpus += snap.task_req[t, 2]
        if gpus != float(exp.get("GPUsRequired", 0) or 0):
            errs.append(f"job {jname}: GPUsRequired {gpus} expected {exp.get('GPUsRequired')}")
    for tname, exp in (topo.get("TaskExpectedResults") or {}).items():
        if tname not in meta["task_names"]:
            continue  # test_utils.go:213-219 iterates the session's tasks: an expectation naming no task is never read
        t = meta["task_names"].index(tname)
        st = status_name[int(res.task_status[t])]
        if st != exp.get("Status"):
            errs.append(f"task {tname}: status {st}, expected {exp.get('Status')}")
        want = exp.get("NodeName") or ""
        got = meta["node_names"][res.task_node[t]] if res.task_node[t] >= 0 else meta["task_fixture_node"][t]
        if want and got != want:
            errs.append(f"task {tname}: node {got!r}, expected {want!r}")
    for nname, exp in (topo.get("ExpectedNodesResources") or {}).items():
        n = meta["node_names"].index(nname)
        if res.node_idle[2, n] != float(exp.get("IdleGPUs", 0) or 0):
            errs.append(f"node {nname}: idle GPUs {res.node_idle[2, n]} expected {exp.get('IdleGPUs')}")
        if res.node_releasing[2, n] != float(exp.get("ReleasingGPUs", 0) or 0):
            errs.append(f"node {nname}: releasing GPUs {res.node_releasing[2, n]} expected {exp.get('ReleasingGPUs')}")
    return errs


# ---------------- multi-round integration harness (integration_tests_utils.go:41-140) ----------------
class _SnapshotAsResult:
    """MatchExpectedAndRealTasks on a freshly built session (prepareSessionForMatch): the snapshot itself."""

    def __init__(self, snap):
        self.task_status = snap.task_status
        self.task_node = snap.task_node
        self.node_idle = snap.node_idle
        self.node_releasing = snap.node_releasing


def _carry_over(topo: dict, meta: dict, res):
    """runSchedulerOneRound's write-back of the session into the test topology (:96-125)."""
    status_name = {v: k for k, v in abi.POD_STATUS_NAMES.items()}
    jobs = {j["Name"]: j for j in topo["Jobs"]}
    for t, tname in enumerate(meta["task_names"]):
        jname, k = tname.rsplit("-", 1)
        job = jobs[jname]
        task = job["Tasks"][int(k)]
        st = status_name[int(res.task_status[t])]
        node = meta["node_names"][res.task_node[t]] if res.task_node[t] >= 0 else (task.get("NodeName") or "")
        if st == "Releasing":
            if job.get("DeleteJobInTest"):
                task["NodeName"], task["State"] = node, "Releasing"
            else:
                task["NodeName"], task["State"] = "", "Pending"
        elif st == "Pipelined":
            task["NodeName"], task["State"] = "", "Pending"
        elif st == "Binding":
            task["NodeName"], task["State"] = node, "Running"
        else:
            task["NodeName"], task["State"] = node, st


def run_integration_case(case: dict, make_runner):
    """make_runner() -> object with load(snap) / run(action).  Returns the list of mismatches."""
    import copy
    topo = copy.deepcopy(case["topology"])
    for j in topo["Jobs"]:
        j["Tasks"] = [dict(t or {}) for t in j["Tasks"]]
    until = case.get("rounds_until_match") or 2
    after = case.get("rounds_after_match") or 5

    def one_round():
        snap, meta = build_snapshot(topo)
        r = make_runner()
        r.load(snap)
        res = None
        for a in case["actions"]:
            res = r.run(a)
        if hasattr(r, "close"):
            r.close()
        _carry_over(topo, meta, res)
        return snap, meta, res

    for _ in range(until):
        one_round()
    snap, meta = build_snapshot(topo)
    errs = [f"after {until} rounds: {e}" for e in check_expectations(topo, meta, _SnapshotAsResult(snap), snap)]
    for i in range(after):
        snap, meta, res = one_round()
        errs += [f"stability round {i}: {e}" for e in check_expectations(topo, meta, res, snap)]
    return errs
