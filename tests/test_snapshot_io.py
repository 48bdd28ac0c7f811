"""snapshot.json wire format <-> SoA (kai_scheduler_b200/snapshot_io.py), CPU only.

Three kinds of evidence:
  * Quantity arithmetic against known answers of k8s.io/apimachinery's resource.Quantity;
  * a hand-written document exercising the cluster_info.Snapshot() rules (pod status, init containers, overhead, bind
    requests, orphan queues, priority classes, foreign pods, node conditions, nodeSelector / affinity / taints);
  * every single-action reference table (tests/golden/actions) pushed through dump -> zip -> pack and run on the
    oracle: the table's expected bindings must still hold after the trip through raw Kubernetes objects.
"""
import io
import json
import os
import tempfile

import numpy as np
import pytest

import dsl
from fixtures import action_cases, case_needs_predicates
from oracle_lib import Oracle

from kai_scheduler_b200 import abi, snapshot_io as sio, synthetic


# ---------------------------------------------------------------------------------------------- quantities
@pytest.mark.parametrize("text,value,milli", [
    ("100m", 1, 100), ("1", 1, 1000), ("2.5", 3, 2500), ("1Gi", 2 ** 30, 2 ** 30 * 1000), ("1G", 10 ** 9, 10 ** 12),
    ("128Mi", 128 * 2 ** 20, 128 * 2 ** 20 * 1000), ("1e3", 1000, 10 ** 6), ("5E-1", 1, 500), ("1n", 1, 1),
    ("1500u", 1, 2), ("0", 0, 0), ("20000", 20000, 2 * 10 ** 7), (4, 4, 4000), ("12k", 12000, 12 * 10 ** 6),
])
def test_quantity_known_answers(text, value, milli):
    assert sio.quantity_value(text) == value
    assert sio.quantity_milli_value(text) == milli


def test_quantity_rejects_garbage():
    for bad in ("", "abc", "1Zi", "--1"):
        with pytest.raises(ValueError):
            sio.parse_quantity(bad)


# ---------------------------------------------------------------------------------------------- hand-written document
def _pod(name, group=None, phase="Pending", node=None, requests=None, **extra):
    md = {"name": name, "namespace": "ns", "uid": "uid-" + name, "creationTimestamp": "2025-01-01T00:00:00Z",
          "annotations": {}, "labels": {}}
    if group:
        md["annotations"][sio.POD_GROUP_ANNOTATION] = group
    spec = {"schedulerName": "kai-scheduler",
            "containers": [{"name": "c", "resources": {"requests": requests or {"cpu": "1", "memory": "1G", "nvidia.com/gpu": "1"}}}]}
    if node:
        spec["nodeName"] = node
    pod = {"metadata": md, "spec": spec, "status": {"phase": phase}}
    for k, v in extra.items():
        if k in ("labels", "deletionTimestamp", "uid", "creationTimestamp"):
            if k == "labels":
                md["labels"].update(v)
            else:
                md[k] = v
        elif k == "nominated":
            pod["status"]["nominatedNodeName"] = v
        else:
            spec[k] = v
    return pod


def _node(name, gpus=8, labels=None, taints=None, ready=True, unschedulable=False, extra=None):
    alloc = {"cpu": "16", "memory": "64Gi", "nvidia.com/gpu": str(gpus), "pods": "110", "ephemeral-storage": "100Gi",
             "hugepages-2Mi": "0"}
    alloc.update(extra or {})
    return {"metadata": {"name": name, "labels": labels or {}},
            "spec": {"taints": taints or [], "unschedulable": unschedulable},
            "status": {"allocatable": alloc,
                       "conditions": [{"type": "Ready", "status": "True" if ready else "False"},
                                      {"type": "MemoryPressure", "status": "False"}]}}


def _queue(name, parent=None, gpu=(0, -1, 1), created="2025-01-01T00:00:00Z", priority=None, mem=(-1, -1, 1)):
    spec = {"resources": {"gpu": dict(zip(("quota", "limit", "overQuotaWeight"), gpu)),
                          "cpu": {"quota": -1, "limit": -1, "overQuotaWeight": 1},
                          "memory": dict(zip(("quota", "limit", "overQuotaWeight"), mem))}}
    if parent:
        spec["parentQueue"] = parent
    if priority is not None:
        spec["priority"] = priority
    return {"metadata": {"name": name, "creationTimestamp": created}, "spec": spec}


def _handwritten():
    nodes = [
        _node("node-b", labels={"zone": "z1", "rack": "r2", "pool": "a"}),
        _node("node-a", labels={"zone": "z1", "rack": "r1", "pool": "b", sio.GPU_COUNT_LABEL: "16"}),
        _node("node-c", labels={"zone": "z2", "rack": "r1"}, taints=[{"key": "dedicated", "value": "x", "effect": "NoSchedule"}]),
        _node("node-d", ready=False, labels={"zone": "z2", "rack": "r3"}),
        _node("node-e", gpus=0, extra={"example.com/widget": "4"}, labels={"zone": "z2", "rack": "r3"}),
    ]
    queues = [
        _queue("dept", gpu=(16, -1, 2)),
        _queue("team-a", parent="dept", gpu=(8, 12, 1), created="2025-01-01T00:01:00Z", priority=200, mem=(1000, 2000, 1)),
        _queue("team-b", parent="dept", gpu=(8, -1, 1), created="2025-01-01T00:02:00Z"),
        _queue("orphan", parent="nowhere"),
        _queue("orphan-child", parent="orphan"),
    ]
    pods = [
        # running gang member on node-a, one releasing on node-b
        _pod("train-0", "train", "Running", "node-a"),
        _pod("train-1", "train", "Running", "node-b", deletionTimestamp="2025-01-02T00:00:00Z"),
        # pending pods: init container larger than the sum, overhead, task priority label, selector / toleration
        _pod("infer-0", "infer", requests={"cpu": "500m", "memory": "1Gi", "nvidia.com/gpu": "2"},
             initContainers=[{"name": "i", "resources": {"requests": {"cpu": "2", "memory": "512Mi"}}}],
             overhead={"cpu": "250m", "memory": "64Mi"}, nodeSelector={"pool": "a"}),
        _pod("infer-1", "infer", labels={sio.TASK_ORDER_LABEL: "7"}, nominated="node-b",
             tolerations=[{"key": "dedicated", "operator": "Exists"}],
             affinity={"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
                 {"matchExpressions": [{"key": "zone", "operator": "In", "values": ["z2"]}]},
                 {"matchExpressions": [{"key": "rack", "operator": "NotIn", "values": ["r1", "r3"]}]}]}}}),
        _pod("widget-0", "widget", requests={"cpu": "1", "example.com/widget": "2"}),
        # bind request in flight, gated pod, bound pod
        _pod("binding-0", "misc"),
        _pod("gated-0", "misc", schedulingGates=[{"name": "g"}], uid="uid-aaa"),
        _pod("bound-0", "misc", "Pending", "node-a"),
        # a pod of another scheduler holding GPUs, a pod without a group, a finished pod
        _pod("foreign-0", None, "Running", "node-b", schedulerName="default-scheduler",
             requests={"cpu": "2", "memory": "2G", "nvidia.com/gpu": "3"}),
        _pod("loose-0", None, "Running", "node-e", requests={"cpu": "1"}),
        _pod("done-0", "misc", "Succeeded", "node-a"),
        _pod("lost-0", "no-such-group"),
    ]
    pod_groups = [
        {"metadata": {"name": "train", "namespace": "ns", "creationTimestamp": "2025-01-01T00:00:10Z"},
         "spec": {"queue": "team-a", "minMember": 2, "priorityClassName": "train"}},
        {"metadata": {"name": "infer", "namespace": "ns", "creationTimestamp": "2025-01-01T00:00:05Z"},
         "spec": {"queue": "team-b", "minMember": 1, "priorityClassName": "inference",
                  "topologyConstraint": {"topology": "dc", "requiredTopologyLevel": "zone", "preferredTopologyLevel": "rack"}}},
        {"metadata": {"name": "widget", "namespace": "ns", "creationTimestamp": "2025-01-01T00:00:05Z"},
         "spec": {"queue": "team-b", "preemptibility": "non-preemptible"}},
        {"metadata": {"name": "misc", "namespace": "ns"}, "spec": {"queue": "gone", "minMember": 3}},
    ]
    return {
        "config": {"actions": "allocate, reclaim",
                   "tiers": [{"plugins": [{"name": "nodeplacement", "arguments": {"gpu": "spread", "cpu": "binpack"}},
                                          {"name": "proportion", "arguments": {"kValue": "0.5"}}]}]},
        "schedulerParams": {"schedulerName": "kai-scheduler", "fullHierarchyFairness": True, "useSchedulingSignatures": True,
                            "maxNumberConsolidationPreemptees": 16, "globalDefaultStalenessGracePeriod": 60 * 10 ** 9},
        "rawObjects": {
            "pods": pods, "nodes": nodes, "queues": queues, "podGroups": pod_groups,
            "bindRequests": [{"metadata": {"name": "br", "namespace": "ns"},
                              "spec": {"podName": "binding-0", "selectedNode": "node-b"}},
                             {"metadata": {"name": "br2", "namespace": "ns"},
                              "spec": {"podName": "gated-0", "selectedNode": "deleted-node"}}],
            "priorityClasses": [{"metadata": {"name": "train"}, "value": 50},
                                {"metadata": {"name": "inference"}, "value": 125},
                                {"metadata": {"name": "fallback"}, "value": 75, "globalDefault": True}],
            "topologies": [{"metadata": {"name": "dc"}, "spec": {"levels": [{"nodeLabel": "zone"}, {"nodeLabel": "rack"}]}}],
        },
    }


def test_handwritten_document():
    snap, meta, kw, actions = sio.pack_cluster(_handwritten())
    assert actions == ["allocate", "reclaim"]
    assert kw["gpu_placement"] == abi.PLACEMENT_SPREAD and kw["cpu_placement"] == abi.PLACEMENT_BINPACK
    assert kw["k_value"] == 0.5 and kw["use_scheduling_signatures"] and kw["max_consolidation_preemptees"] == 16
    assert kw["staleness_grace_period_s"] == 60 and kw["allow_consolidating_reclaim"] is False

    # nodes by name; the widget resource gets a column because a pod requests it, storage / hugepages do not
    assert meta["node_names"] == ["node-a", "node-b", "node-c", "node-d", "node-e"]
    assert meta["resource_names"] == ["cpu", "memory", "gpu", "pods", "example.com/widget"]
    A, I, L = snap.node_allocatable, snap.node_idle, snap.node_releasing
    assert A[0, 0] == 16000 and A[1, 0] == 64 * 2 ** 30 and A[2, 0] == 8 and A[3, 0] == 110
    assert A[4, 4] == 4000 and A[2, 4] == 0
    # node-a: train-0 running + bound-0 bound (done-0 finished, not counted)
    assert I[2, 0] == 8 - 2 and I[0, 0] == 16000 - 2000 and I[3, 0] == 108
    # node-b: train-1 releasing (1 GPU) + binding-0 via its bind request (1) + foreign-0 (3)
    assert I[2, 1] == 8 - 5 and L[2, 1] == 1 and L[0, 1] == 1000 and I[3, 1] == 107
    assert snap.node_foreign[2, 1] == 3 and snap.node_foreign[0, 1] == 2000 and snap.node_foreign.sum() == 3 + 2000 + 2e9
    assert I[0, 4] == 16000 - 1000  # the loose pod still holds CPU on node-e
    assert list(snap.node_flags & abi.NODE_READY) == [1, 1, 1, 0, 1]
    assert list(snap.node_gpu_count) == [16, 8, 8, 8, 0]

    # queues: orphans dropped with their subtree; memory quota in MB -> bytes
    assert meta["queue_names"] == ["dept", "team-a", "team-b"]
    assert list(snap.queue_parent) == [-1, 0, 0] and list(snap.queue_priority) == [100, 200, 100]
    assert snap.queue_deserved[2].tolist() == [16, 8, 8] and snap.queue_limit[2].tolist() == [-1, 12, -1]
    assert snap.queue_deserved[1].tolist() == [-1, 1e9, -1] and snap.queue_limit[1].tolist() == [-1, 2e9, -1]
    assert snap.queue_oqw[2].tolist() == [2, 1, 1]
    assert snap.queue_creation[2] - snap.queue_creation[1] == 60

    # jobs by name: infer, misc, train, widget
    assert meta["job_names"] == ["infer", "misc", "train", "widget"]
    assert list(snap.job_queue) == [2, -1, 1, 2]
    assert list(snap.job_priority) == [125, 0, 50, 75]  # class value, queue missing -> zero value, class, global default
    assert list(snap.job_flags & abi.JOB_PREEMPTIBLE) == [0, 0, 1, 0]
    # creation: misc has none (epoch 0) < infer == widget (UID breaks the tie) < train
    assert list(snap.job_order_rank) == [1, 0, 3, 2]
    assert list(snap.podset_min_available) == [1, 3, 2, 1]

    t = {n: i for i, n in enumerate(meta["task_names"])}
    assert "lost-0" not in t and "foreign-0" not in t and len(t) == 9
    st = snap.task_status
    assert st[t["train-0"]] == abi.POD_RUNNING and st[t["train-1"]] == abi.POD_RELEASING
    assert st[t["binding-0"]] == abi.POD_BINDING and snap.task_node[t["binding-0"]] == 1
    assert st[t["gated-0"]] == abi.POD_GATED and st[t["bound-0"]] == abi.POD_BOUND and st[t["done-0"]] == abi.POD_SUCCEEDED
    assert snap.task_node[t["done-0"]] == -1 and snap.task_node[t["infer-0"]] == -1
    # infer-0: max(sum, init) per resource + overhead: cpu max(500, 2000) + 250, memory max(1Gi, 512Mi) + 64Mi
    assert snap.task_req[t["infer-0"]].tolist() == [2250, 2 ** 30 + 64 * 2 ** 20, 2, 1, 0]
    assert snap.task_req[t["widget-0"]].tolist() == [1000, 0, 0, 1, 2000]
    # task order inside `infer`: the labelled pod first; inside `misc`: equal creation -> UID order
    assert snap.task_order_rank[t["infer-1"]] == 0 and snap.task_order_rank[t["infer-0"]] == 1
    misc = sorted(["binding-0", "gated-0", "bound-0", "done-0"], key=lambda n: "uid-aaa" if n == "gated-0" else "uid-" + n)
    assert [snap.task_order_rank[t[n]] for n in misc] == [0, 1, 2, 3]
    assert snap.task_nominated[t["infer-1"]] == 1 and snap.task_nominated[t["infer-0"]] == -1

    # predicate classes: bit n = node n passes
    def mask(name):
        c = snap.task_pred_class[t[name]]
        assert c >= 0
        return [int(snap.pred_mask[c, n // 32] >> (n % 32)) & 1 for n in range(5)]

    assert mask("train-0") == [1, 1, 0, 0, 1]  # untolerated taint on node-c, node-d not ready
    assert mask("infer-0") == [0, 1, 0, 0, 0]  # pool=a
    assert mask("infer-1") == [0, 1, 1, 0, 1]  # zone z2 (c, e; d not ready) or rack not in {r1, r3} (b); taint tolerated

    # topology: zone ids z1=0, z2=1; rack DomainIDs z1.r1, z1.r2, z2.r1, z2.r3
    assert snap.node_domain[0].tolist() == [0, 0, 1, 1, 1]
    assert snap.node_domain[1].tolist() == [0, 1, 2, 3, 3]
    j = meta["job_names"].index("infer")
    g = snap.job_sgs_begin[j]
    assert (snap.sgs_topology[g], snap.sgs_required_level[g], snap.sgs_preferred_level[g]) == (0, 0, 1)
    assert snap.sgs_topology[snap.job_sgs_begin[0 if j else 1]] in (-1, 0)

    # signatures hash constraints, not requests: `train` (its Releasing pod is not active-allocated) and `widget`
    # both have unconstrained pods and no topology constraint, `infer` and `misc` differ
    sig = dict(zip(meta["job_names"], snap.job_signature.tolist()))
    assert sig["train"] == sig["widget"] and len(set(sig.values())) == 3

    # the packed snapshot is loadable and schedulable
    o = Oracle(abi.make_config(**kw))
    o.load(snap)
    res = o.run("allocate")
    assert res.task_node[t["infer-0"]] == 1  # the only node with pool=a


def test_project_level_fairness_and_unsupported():
    doc = _handwritten()
    doc["schedulerParams"]["fullHierarchyFairness"] = False
    snap, meta, _, _ = sio.pack_cluster(doc)
    # top-level queues are dropped, the others hang off the synthetic `default` parent (queue.go:24-49,68-80)
    assert meta["queue_names"] == ["default", "orphan", "orphan-child", "team-a", "team-b"]
    assert list(snap.queue_parent) == [-1, 0, 0, 0, 0]
    assert snap.queue_deserved[:, 0].tolist() == [-1, -1, -1]

    bad = _handwritten()
    bad["rawObjects"]["pods"][2]["metadata"]["annotations"]["gpu-fraction"] = "0.5"
    with pytest.raises(sio.UnsupportedSnapshot):
        sio.pack_cluster(bad)
    soft = _handwritten()
    soft["rawObjects"]["pods"][2]["spec"]["topologySpreadConstraints"] = [{"maxSkew": 1}]
    with pytest.raises(sio.UnsupportedSnapshot):
        sio.pack_cluster(soft)
    _, meta, _, _ = sio.pack_cluster(soft, strict=False)
    assert any("topology spread" in m for m in meta["ignored"])


def test_identical_pod_groups_share_a_signature():
    doc = _handwritten()
    raw = doc["rawObjects"]
    raw["podGroups"].append({"metadata": {"name": "widget2", "namespace": "ns"},
                             "spec": {"queue": "team-a", "preemptibility": "non-preemptible"}})
    raw["pods"].append(_pod("widget2-0", "widget2", requests={"cpu": "3", "example.com/widget": "1"}))
    snap, meta, _, _ = sio.pack_cluster(doc)
    sig = dict(zip(meta["job_names"], snap.job_signature.tolist()))
    assert sig["widget"] == sig["widget2"]  # requests are not part of the signature (scheduling_constraints_signature.go)
    assert sig["infer"] != sig["widget"] and sig["misc"] != sig["widget"]


# ---------------------------------------------------------------------------------------------- reference tables
TABLES = (action_cases(["allocate__"], single_action="allocate") + action_cases(["reclaim__"], single_action="reclaim")
          + action_cases(["consolidation__"], single_action="consolidation") + action_cases(["preempt__"], single_action="preempt")
          + action_cases(["stalegangeviction__"], single_action="stalegangeviction"))
_trip_stats = {"run": 0, "skipped": 0}


def _affinity_of(case, meta):
    """{pod name: NodeAffinityNames} and {node name: its `kai.scheduler/type` label} of a table (tasks_fake/tasks.go:98-116,
    nodes_fake/nodes.go:182-191)."""
    pods = {}
    for job in case["topology"].get("Jobs") or []:
        for k, t in enumerate(job.get("Tasks") or []):
            if t.get("NodeAffinityNames"):
                pods[f"{job['Name']}-{k}"] = t["NodeAffinityNames"]
    if not pods:
        return None
    nodes = {}
    for name, nd in (case["topology"].get("Nodes") or {}).items():
        labels = nd.get("Labels") or {}
        nodes[name] = labels.get("tasks_fake.NodeAffinityKey", labels.get("kai.scheduler/type", name))
    return pods, nodes


def _round_trip(snap, actions, names=None, config=None, affinity=None):
    doc = sio.dump_cluster(snap, actions=actions, names=names, config=config)
    if affinity:
        pods, nodes = affinity
        for node in doc["rawObjects"]["nodes"]:
            node["metadata"]["labels"]["kai.scheduler/type"] = nodes[node["metadata"]["name"]]
        for pod in doc["rawObjects"]["pods"]:
            if pod["metadata"]["name"] in pods:
                pod["spec"]["affinity"] = {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
                    {"matchExpressions": [{"key": "kai.scheduler/type", "operator": "In", "values": pods[pod["metadata"]["name"]]}]}]}}}
    buf = io.BytesIO()
    sio.write_snapshot_zip(buf, doc)
    return sio.pack_cluster(sio.read_snapshot_zip(buf.getvalue()))


@pytest.mark.parametrize("cid,case", TABLES, ids=[c[0] for c in TABLES])
def test_reference_tables_through_the_wire_format(cid, case):
    snap, meta = dsl.build_snapshot(case["topology"])
    place = {"binpack": abi.PLACEMENT_BINPACK, "spread": abi.PLACEMENT_SPREAD}
    config = {k: place[v] if isinstance(v, str) else v for k, v in (case.get("config") or {}).items()}
    if case_needs_predicates(case):  # the table's node affinity goes back into the pods, the packer re-derives the masks
        snap.task_pred_class = snap.pred_mask = None
    try:
        snap2, meta2, kw, actions = _round_trip(snap, case["actions"], names=meta, config=config,
                                                affinity=_affinity_of(case, meta))
    except sio.UnsupportedSnapshot as e:  # tables that start from session-only statuses (Allocated / Pipelined)
        _trip_stats["skipped"] += 1
        pytest.skip(str(e))
    _trip_stats["run"] += 1
    assert actions == case["actions"]
    assert sorted(meta2["task_names"]) == sorted(meta["task_names"])
    # same cluster: per-node tables by name
    perm = [meta["node_names"].index(n) for n in meta2["node_names"]]
    assert np.array_equal(snap2.node_allocatable, snap.node_allocatable[:, perm])
    assert np.array_equal(snap2.node_releasing, snap.node_releasing[:, perm])
    fixture_node = dict(zip(meta["task_names"], meta["task_fixture_node"]))
    meta2["task_fixture_node"] = [fixture_node[n] for n in meta2["task_names"]]
    kw["max_consolidation_preemptees"] = -1  # what the reference's action tests run with
    kw["allow_consolidating_reclaim"] = True
    o = Oracle(abi.make_config(**kw))
    o.load(snap2)
    res = o.run(case["actions"][0])
    errs = dsl.check_expectations(case["topology"], meta2, res, snap2)
    assert not errs, f"{case['source']} #{case['index']} {case['name']}: {errs}"


def test_most_tables_survive_the_trip():
    if _trip_stats["run"] + _trip_stats["skipped"] < len(TABLES):
        pytest.skip("run after the table tests")
    assert _trip_stats["run"] >= 150, _trip_stats


# ---------------------------------------------------------------------------------------------- synthetic configs
def _by_rank(res_nodes, rank):
    out = np.empty_like(res_nodes)
    out[:, rank] = res_nodes
    return out


@pytest.mark.parametrize("name", ["config1", "cycle5-small", "config4-small"])
def test_synthetic_configs_round_trip(name):
    """BASELINE configs without names: generated names must reproduce every rank array, so the oracle's outcome on the
    repacked snapshot is the original outcome up to the index permutation (nodes by name rank, queues by UID rank)."""
    snap = synthetic.config_snapshot(name)
    actions = synthetic.CONFIG_ACTIONS.get(name, ["allocate"])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "snapshot.zip")
        sio.write_snapshot_zip(path, sio.dump_cluster(snap, actions=actions, config={"allow_consolidating_reclaim": True}))
        snap2, meta2, kw, actions2 = sio.pack_cluster(sio.read_snapshot_zip(path))
    assert actions2 == list(actions)
    assert snap2.n_tasks == snap.n_tasks and snap2.n_jobs == snap.n_jobs and snap2.n_queues == snap.n_queues
    assert np.array_equal(_by_rank(snap.node_idle, snap.node_name_rank), snap2.node_idle)
    assert np.array_equal(snap2.job_order_rank, snap.job_order_rank)  # jobs keep their index order
    # tasks come back in task-order inside their PodSet: match them by (job, order rank)
    def task_jobs(sn):
        tj = np.zeros(sn.n_tasks, dtype=np.int64)
        for j in range(sn.n_jobs):
            b, e = sn.podset_task_begin[sn.job_podset_begin[j]], sn.podset_task_begin[sn.job_podset_begin[j + 1]]
            tj[b:e] = j
        return tj

    key1 = task_jobs(snap) * 10 ** 6 + snap.task_order_rank
    key2 = task_jobs(snap2) * 10 ** 6 + snap2.task_order_rank
    assert len(set(key1.tolist())) == snap.n_tasks and sorted(key1.tolist()) == sorted(key2.tolist())
    o1 = np.argsort(key1)
    perm = np.empty(snap.n_tasks, dtype=np.int64)  # perm[t2] = the original task
    perm[np.argsort(key2)] = o1
    assert np.array_equal(snap2.task_status, snap.task_status[perm])
    assert np.array_equal(snap2.task_req, snap.task_req[perm])
    if snap.node_domain is not None:
        assert np.array_equal(_by_rank(snap.node_domain, snap.node_name_rank), snap2.node_domain)
    a, b = Oracle(abi.make_config()), Oracle(abi.make_config(**kw))
    a.load(snap)
    b.load(snap2)
    for act in actions:
        ra, rb = a.run(act), b.run(act)
        assert np.array_equal(ra.task_status[perm], rb.task_status), act
        node_a = np.where(ra.task_node >= 0, snap.node_name_rank[np.maximum(ra.task_node, 0)], -1)
        assert np.array_equal(node_a[perm], rb.task_node), act
        assert ra.pods_placed == rb.pods_placed and ra.pods_evicted == rb.pods_evicted


# ---------------------------------------------------------------------------------------------- node_info fit test
# api/node_info/node_info_test.go:677-794 TestIsTaskAllocatable: node allocatable, running pods and a candidate pod as
# Kubernetes quantities -> does the pod fit the node's Idle resources.  Run through the whole wire path: the packer's
# Quantity arithmetic and request computation (incl. pod overhead), then one allocate cycle on the oracle.
# (node allocatable, [running pod requests], candidate requests, candidate overhead, expected)
def _rl(cpu, mem, gpu=None, pods=None):
    rl = {"cpu": cpu, "memory": mem}
    if gpu is not None:
        rl["nvidia.com/gpu"] = gpu
    if pods is not None:
        rl["pods"] = pods
    return rl


IS_TASK_ALLOCATABLE = [
    ("add pod with not enough cpu and memory", _rl("2000m", "2G", pods="110"), [_rl("1000m", "1G")], _rl("2000m", "2G"), None, False),
    ("add pod with not enough cpu - 1 millicpu", _rl("2000m", "2G", pods="110"), [_rl("1000m", "1G")], _rl("1001m", "1G"), None, False),
    ("add pod with not enough memory - 1 Kb", _rl("2000m", "2G", pods="110"), [_rl("1000m", "2G")], _rl("1000m", "1Ki"), None, False),
    ("add pod with enough cpu and memory", _rl("2000m", "2G", pods="110"), [_rl("1000m", "1G")], _rl("1000m", "1G"), None, True),
    ("task in capacity but not in available", _rl("1000m", "1G", pods="110"), [_rl("1000m", "1G")], _rl("1000m", "1G"), None, False),
    ("missing gpu", _rl("2000m", "2G", pods="110"), [], _rl("1000m", "1G", gpu="1"), None, False),
    ("already used gpu so missing gpu", _rl("2000m", "2G", "1", "110"), [_rl("1000m", "1G", gpu="1")], _rl("1000m", "1G", gpu="1"), None, False),
    ("enough cpu memory and gpu", _rl("2000m", "2G", "2", "110"), [_rl("1000m", "1G", gpu="1")], _rl("1000m", "1G", gpu="1"), None, True),
    ("pod with overhead that fits without overhead but not with overhead", _rl("2000m", "2G", pods="110"), [_rl("1000m", "1G")],
     _rl("500m", "500M"), _rl("600m", "600M"), False),
    ("pod with overhead that doesn't fit even without overhead", _rl("2000m", "2G", pods="110"), [_rl("1000m", "1G")],
     _rl("1500m", "1500M"), _rl("100m", "100M"), False),
    ("pod without overhead that doesn't fit", _rl("2000m", "2G", pods="110"), [_rl("1000m", "1G")], _rl("1500m", "1500M"), None, False),
    ("pod with overhead that fits with overhead", _rl("2000m", "2G", pods="110"), [_rl("1000m", "1G")],
     _rl("500m", "500M"), _rl("100m", "100M"), True),
]


@pytest.mark.parametrize("name,node,running,candidate,overhead,expected", IS_TASK_ALLOCATABLE, ids=[c[0] for c in IS_TASK_ALLOCATABLE])
def test_is_task_allocatable_through_the_wire_format(name, node, running, candidate, overhead, expected):
    pods = [_pod(f"p{i}", "running", "Running", "n1", requests=r) for i, r in enumerate(running)]
    extra = {"overhead": overhead} if overhead else {}
    pods.append(_pod("podToAllocate", "candidate", requests=candidate, **extra))
    doc = {"config": {"actions": "allocate"}, "schedulerParams": {"fullHierarchyFairness": True},
           "rawObjects": {"pods": pods,
                          "nodes": [{"metadata": {"name": "n1"}, "spec": {}, "status": {"allocatable": node}}],
                          "queues": [_queue("q", gpu=(-1, -1, 1))],
                          "podGroups": [{"metadata": {"name": g, "namespace": "ns"}, "spec": {"queue": "q", "minMember": 1}}
                                        for g in ("running", "candidate")]}}
    snap, meta, kw, actions = sio.pack_cluster(doc)
    o = Oracle(abi.make_config(**kw))
    o.load(snap)
    res = o.run("allocate")
    t = meta["task_names"].index("podToAllocate")
    assert (res.task_status[t] == abi.POD_BINDING) == expected


# ---------------------------------------------------------------------------------------------- cluster totals
# plugins/proportion/proportion_test.go:526-800 "Get Node Resources" (getNodeResources, proportion.go:258-289): what one
# node contributes to the fair-share totals.  (name, allocatable, pods = (scheduler, phase, labels, requests), want)
def _npod(i, scheduler, phase, requests, labels=None, on_node=True):
    p = _pod(f"pod-{i}", None, phase, "n1" if on_node else None, requests=requests, schedulerName=scheduler)
    p["metadata"]["labels"].update(labels or {})
    return p


NODE_RESOURCES = [
    ("cpu + memory node", _rl("8000m", "10G"), [], [8000, 1e10, 0]),
    ("gpu node", _rl("8000m", "10G", gpu="2"), [], [8000, 1e10, 2]),
    ("ignore extra resources", {"A": "4"}, [], [0, 0, 0]),
    ("Count out resources for non-related pods", _rl("8000m", "10G"),
     [("kai-scheduler", "Running", None, _rl("2", "2G")), ("default-scheduler", "Running", None, _rl("1", "1G"))], [7000, 9e9, 0]),
    ("consider reservation pods", _rl("8000m", "10G"),
     [("kai-scheduler", "Running", None, _rl("2", "2G")), ("default-scheduler", "Running", None, _rl("1", "1G")),
      ("default-scheduler", "Running", {"app": "kai-resource-reservation"}, _rl("1", "1G"))], [7000, 9e9, 0]),
    ("consider scaler pods", _rl("8000m", "10G"),
     [("kai-scheduler", "Running", None, _rl("2", "2G")), ("default-scheduler", "Running", None, _rl("1", "1G")),
      ("default-scheduler", "Running", {"app": "scaling-pod"}, _rl("1", "1G"))], [7000, 9e9, 0]),
    ("Do not count out resources for non-related pods if non active", _rl("8000m", "10G"),
     [("default-scheduler", "Pending", None, _rl("1", "1G")), ("default-scheduler", "Succeeded", None, _rl("1", "1G"))], [8000, 1e10, 0]),
]


@pytest.mark.parametrize("name,allocatable,pods,want", NODE_RESOURCES, ids=[c[0] for c in NODE_RESOURCES])
def test_node_resources_for_fair_share_totals(name, allocatable, pods, want):
    raw_pods = [_npod(i, sched, phase, req, labels, on_node=phase == "Running") for i, (sched, phase, labels, req) in enumerate(pods)]
    doc = {"config": {"actions": "allocate"}, "schedulerParams": {"fullHierarchyFairness": True},
           "rawObjects": {"pods": raw_pods, "nodes": [{"metadata": {"name": "n1"}, "spec": {}, "status": {"allocatable": allocatable}}],
                          "queues": [_queue("q")], "podGroups": []}}
    snap, meta, kw, _ = sio.pack_cluster(doc)
    o = Oracle(abi.make_config(**kw))
    o.load(snap)
    assert o.fair_share().total_resource.tolist() == [float(x) for x in want]


# ---------------------------------------------------------------------------------------------- cluster_info_test.go
def _cluster(nodes=(), pods=(), queues=(), pod_groups=(), priority_classes=(), params=None):
    return {"config": {"actions": "allocate"}, "schedulerParams": dict({"fullHierarchyFairness": True}, **(params or {})),
            "rawObjects": {"nodes": list(nodes), "pods": list(pods), "queues": list(queues), "podGroups": list(pod_groups),
                           "priorityClasses": list(priority_classes)}}


def _cpu_node(name, labels=None):
    return {"metadata": {"name": name, "labels": labels or {}}, "spec": {}, "status": {"allocatable": {"cpu": "10", "pods": "110"}}}


def _cpu_pod(name, node, phase="Running"):
    return {"metadata": {"name": name, "namespace": "ns", "uid": name}, "spec": {"nodeName": node, "containers": [
        {"resources": {"requests": {"cpu": "2"}}}]}, "status": {"phase": phase}}


def test_snapshot_nodes():  # cache/cluster_info/cluster_info_test.go:244-501 TestSnapshotNodes (BasicUsage, Finished job, node pool)
    snap, meta, _, _ = sio.pack_cluster(_cluster(nodes=[_cpu_node("node-1")], pods=[_cpu_pod("p", "node-1")]))
    assert snap.node_idle[:, 0].tolist() == [8000, 0, 0, 109] and snap.node_releasing[:, 0].tolist() == [0, 0, 0, 0]
    snap, _, _, _ = sio.pack_cluster(_cluster(nodes=[_cpu_node("node-1")], pods=[_cpu_pod("p", "node-1", "Succeeded")]))
    assert snap.node_idle[:, 0].tolist() == [10000, 0, 0, 110]
    pool = {"partitionParams": {"NodePoolLabelKey": "pool", "NodePoolLabelValue": "pool-a"}}
    snap, meta, _, _ = sio.pack_cluster(_cluster(
        nodes=[_cpu_node("node-1", {"pool": "pool-a"}), _cpu_node("node-2", {"pool": "pool-b"})],
        pods=[_cpu_pod("p1", "node-1"), _cpu_pod("p2", "node-2")], params=pool))
    assert meta["node_names"] == ["node-1"] and snap.node_idle[:, 0].tolist() == [8000, 0, 0, 109]


def test_snapshot_queues_and_flat_hierarchy():  # cluster_info_test.go:1326-1480
    def q(name, parent=None, labels=None, gpu=None):
        spec = {"resources": {"gpu": gpu or {"quota": 2}}}
        if parent:
            spec["parentQueue"] = parent
        return {"metadata": {"name": name, "labels": labels or {}}, "spec": spec}

    # TestSnapshotQueues: the default partition selector is "label absent" -> the queue of another node pool is dropped
    doc = _cluster(queues=[q("department0", gpu={"quota": 4}), q("department0-a", labels={"nodepool": "nodepool-a"}),
                           q("queue0", parent="department0")], params={"partitionParams": {"NodePoolLabelKey": "nodepool"}})
    snap, meta, _, _ = sio.pack_cluster(doc)
    assert meta["queue_names"] == ["department0", "queue0"] and list(snap.queue_parent) == [-1, 0]
    assert snap.queue_deserved[2].tolist() == [4, 2]
    # TestSnapshotFlatHierarchy: fullHierarchyFairness off -> a synthetic `default` parent (quota -1, weight 1, limit -1),
    # the departments dropped, their queues re-parented
    dept = {"quota": 4, "overQuotaWeight": 2, "limit": 10}
    lab = {"nodepool": "nodepool-a"}
    doc = _cluster(queues=[q("department0", labels=lab, gpu=dept), q("department1", labels=lab, gpu=dept),
                           q("queue0", "department0", lab, {}), q("queue1", "department1", lab, {})],
                   params={"fullHierarchyFairness": False,
                           "partitionParams": {"NodePoolLabelKey": "nodepool", "NodePoolLabelValue": "nodepool-a"}})
    snap, meta, _, _ = sio.pack_cluster(doc)
    assert meta["queue_names"] == ["default", "queue0", "queue1"] and list(snap.queue_parent) == [-1, 0, 0]
    assert snap.queue_deserved[:, 0].tolist() == [-1, -1, -1] and snap.queue_limit[:, 0].tolist() == [-1, -1, -1]
    assert snap.queue_oqw[:, 0].tolist() == [1, 1, 1]


def test_pod_group_priority_resolution():  # cluster_info_test.go:1482-1506,1660-1731
    def pg(name, cls=None):
        spec = {"queue": "q"}
        if cls:
            spec["priorityClassName"] = cls
        return {"metadata": {"name": name, "namespace": "ns"}, "spec": spec}

    def prio(classes):
        snap, meta, _, _ = sio.pack_cluster(_cluster(queues=[_queue("q")], pod_groups=[pg("with-class", "my-priority"), pg("without")],
                                                      priority_classes=classes))
        return dict(zip(meta["job_names"], snap.job_priority.tolist()))

    # TestGetPodGroupPriority / TestGetDefaultPriorityNotExists: the named class wins, no global default -> 50
    assert prio([{"metadata": {"name": "my-priority"}, "value": 2}]) == {"with-class": 2, "without": 50}
    # TestGetDefaultPriority: a globalDefault class is the fallback
    assert prio([{"metadata": {"name": "my-priority"}, "value": 2, "globalDefault": True}]) == {"with-class": 2, "without": 2}
    # TestGetPodGroupPriorityNotExistingPriority / TestGetDefaultPriorityWithError: unknown class -> the default (50 here)
    assert prio([]) == {"with-class": 50, "without": 50}


def test_snapshot_pod_groups():  # cluster_info_test.go:957-1275 BasicUsage / NotExistingQueue + job_info.go:200-216
    pod = {"metadata": {"name": "test-pod", "namespace": "ns", "uid": "test-pod", "annotations": {sio.POD_GROUP_ANNOTATION: "podGroup-0"}},
           "spec": {"containers": []}, "status": {"phase": "Pending"}}
    pg = {"metadata": {"name": "podGroup-0", "uid": "ABC"}, "spec": {"queue": "queue-0"}}
    snap, meta, _, _ = sio.pack_cluster(_cluster(queues=[_queue("queue-0")], pods=[pod], pod_groups=[pg]))
    assert meta["job_names"] == ["podGroup-0"] and list(snap.job_queue) == [0] and list(snap.podset_min_available) == [1]
    assert meta["task_names"] == ["test-pod"] and snap.task_req[0].tolist() == [0, 0, 0, 1]
    pg["spec"]["queue"] = "queue-1"  # NotExistingQueue: the job is kept, without a queue (and gets a fit error upstream)
    snap, meta, _, _ = sio.pack_cluster(_cluster(queues=[_queue("queue-0")], pods=[pod], pod_groups=[pg]))
    assert list(snap.job_queue) == [-1] and meta["task_names"] == ["test-pod"]
    # minMember feeds the default PodSet, SubGroups replace it (setSubGroups)
    pg["spec"] = {"queue": "queue-0", "minMember": 3}
    snap, _, _, _ = sio.pack_cluster(_cluster(queues=[_queue("queue-0")], pods=[pod], pod_groups=[pg]))
    assert list(snap.podset_min_available) == [3]
    pg["spec"] = {"queue": "queue-0", "minMember": 3, "subGroups": [{"name": "a", "minMember": 2}, {"name": "b"}]}
    snap, meta, _, _ = sio.pack_cluster(_cluster(queues=[_queue("queue-0")], pods=[pod], pod_groups=[pg]))
    assert list(snap.podset_min_available) == [2, 1] and meta["task_names"] == []  # the unlabelled pod matches no PodSet


UP_FOR_SCHEDULER = [  # cluster_info_test.go:1825-1926 TestIsPodGroupUpForScheduler: (backoff, pool label, last condition's pool, kept)
    ("Infinite schedulingBackoff", -1, "nodepoola", "nodepoola", True),
    ("Nil schedulingBackoff", None, "nodepoolb", "nodepoolb", True),
    ("No last scheduling condition", 1, "nodepoolb", None, True),
    ("No last scheduling condition - default node pool", 1, "default", None, True),
    ("unassigned by condition from different node pool", 1, "nodepoola", "different-nodepool", True),
    ("unassigned by condition", 1, "nodepoolc", "nodepoolc", False),
    ("unassigned by condition - default node pool", 1, "default", "different-nodepool", True),
    ("unassigned by condition - default node pool 2", 1, "nodepoolc", "default", True),
    ("unassigned by condition - default node pool (same)", 1, "default", "default", False),
]


@pytest.mark.parametrize("name,backoff,pool,last_pool,kept", UP_FOR_SCHEDULER, ids=[c[0] for c in UP_FOR_SCHEDULER])
def test_is_pod_group_up_for_scheduler(name, backoff, pool, last_pool, kept):
    key = "kai.scheduler/node-pool"
    pg = {"metadata": {"name": "test-pg", "labels": {} if pool in ("default", "") else {key: pool}}, "spec": {"queue": "q"},
          "status": {"schedulingConditions": [] if last_pool is None else [{"nodePool": last_pool}]}}
    if backoff is not None:
        pg["spec"]["schedulingBackoff"] = backoff
    # the scheduler of the pool the pod group lives in: value = the pool ("" selects unlabelled objects, the default pool)
    params = {"partitionParams": {"NodePoolLabelKey": key, "NodePoolLabelValue": "" if pool == "default" else pool}}
    queue = _queue("q")
    queue["metadata"]["labels"] = dict(pg["metadata"]["labels"])
    _, meta, _, _ = sio.pack_cluster(_cluster(queues=[queue], pod_groups=[pg], params=params))
    assert (meta["job_names"] == ["test-pg"]) == kept


def _bind_doc(requests, extra_pods=()):
    pod = {"metadata": {"name": "pod-1", "namespace": "namespace-1", "uid": "pod-1", "annotations": {sio.POD_GROUP_ANNOTATION: "podgroup-1"}},
           "spec": {"containers": [{"resources": {"requests": {"cpu": "2"}}}]}, "status": {"phase": "Pending"}}
    pods = [pod]
    for name in extra_pods:
        other = json.loads(json.dumps(pod))
        other["metadata"]["name"] = other["metadata"]["uid"] = name
        pods.append(other)
    doc = _cluster(nodes=[{"metadata": {"name": "node-1"}, "spec": {}, "status": {"allocatable": {"cpu": "10"}}}], pods=pods,
                   queues=[{"metadata": {"name": "queue-0"}, "spec": {}}],
                   pod_groups=[{"metadata": {"name": "podgroup-1", "namespace": "namespace-1"}, "spec": {"queue": "queue-0"}}])
    doc["rawObjects"]["bindRequests"] = requests
    return doc


def test_bind_requests():  # cache/cluster_info/cluster_info_test.go:503-955 TestBindRequests
    def br(pod="pod-1", node="node-1", **kw):
        spec = {"podName": pod, "selectedNode": node}
        if "backoff" in kw:
            spec["backoffLimit"] = kw["backoff"]
        status = {"phase": kw["phase"], "failedAttempts": kw.get("failed", 0)} if "phase" in kw else {}
        return {"metadata": {"name": "my-pod-1234", "namespace": "namespace-1"}, "spec": spec, "status": status}

    def outcome(requests, extra=()):
        snap, meta, _, _ = sio.pack_cluster(_bind_doc(requests, extra))
        t = meta["task_names"].index("pod-1")
        return int(snap.task_status[t]), snap.node_idle[0, 0]

    assert outcome([br()]) == (abi.POD_BINDING, 8000)  # :547 waiting for binding: the node already holds the pod's CPU
    assert outcome([br(backoff=5, phase="Failed", failed=2)]) == (abi.POD_BINDING, 8000)  # :593 failing, below the limit
    assert outcome([br(backoff=5, phase="Failed", failed=5)]) == (abi.POD_PENDING, 10000)  # :644 reached the limit: stale
    assert outcome([br(pod="not-pod-1")], extra=["not-pod-1"]) == (abi.POD_PENDING, 8000)  # :696 the request is another pod's
    assert outcome([br(node="node-2", phase="Failed")]) == (abi.POD_PENDING, 10000)  # :749 unknown node: not for this snapshot


def test_task_order_fn():  # plugins/taskorder/task_order_test.go:17-53 + session_plugins.go:244-259 (creation, then UID)
    def order(l_label, r_label):
        def pod(name, label):
            p = _pod(name, "g")
            if label is not None:
                p["metadata"]["labels"][sio.TASK_ORDER_LABEL] = label
            return p
        doc = _cluster(queues=[_queue("q")], pods=[pod("l", l_label), pod("r", r_label)],
                       pod_groups=[{"metadata": {"name": "g", "namespace": "ns"}, "spec": {"queue": "q"}}])
        snap, meta, _, _ = sio.pack_cluster(doc)
        rank = dict(zip(meta["task_names"], snap.task_order_rank.tolist()))
        return -1 if rank["l"] < rank["r"] else 1

    # TaskOrderFn(l, r): 1 / 0 / -1; a 0 falls through to creation (equal here) and then the UID ("uid-l" < "uid-r")
    assert order("1", "2") == 1      # the higher task-priority goes first
    assert order("2", "2") == -1     # equal -> UID
    assert order("2", "1") == -1
    assert order(None, "1") == 1     # a labelled pod precedes an unlabelled one
    assert order("1", None) == -1
    assert order(None, None) == -1   # equal -> UID


def test_requirements_from_resource_list():  # api/resource_info/resource_requirment_test.go:14-72
    names = sio._ResourceNames()
    rl = lambda d: sio._resource_list(d, names, True, "test")  # noqa: E731
    assert rl({"cpu": "1", "memory": "5G"}) == {0: 1000, 1: 5_000_000_000}
    assert rl({"cpu": "-1"}) == {0: -1000}
    assert rl({"nvidia.com/gpu": "1"}) == {2: 1} and rl({"amd.com/gpu": "2"}) == {2: 2}
    assert rl({"kai.scheduler/test-resource": "1"}) == {4: 1000} and names.names[4] == "kai.scheduler/test-resource"
    assert rl({"unsupported": "2"}) == {}  # not a scalar resource name: not counted
    assert rl({"hugepages-2Mi": "1", "attachable-volumes-x": "1", "requests.foo/bar": "1"}) == {5: 1000, 6: 1000}


def _slice(name, node, driver, count):  # cluster_info_test.go:2523-2539 createTestResourceSlice
    return {"metadata": {"name": name}, "spec": {"nodeName": node, "driver": driver, "devices": [{"name": f"device-{i}"} for i in range(count)]}}


def test_snapshot_nodes_with_dra_gpus():  # cluster_info_test.go:2411-2521 TestSnapshotNodesWithDRAGPUs
    def dra(nodes, slices):
        doc = _cluster(nodes=[{"metadata": {"name": n}, "spec": {}, "status": {"allocatable": {}}} for n in nodes])
        doc["rawObjects"]["resourceSlices"] = slices
        snap, meta, _, _ = sio.pack_cluster(doc)
        return {n: (snap.node_allocatable[2, i], bool(snap.node_flags[i] & abi.NODE_NOT_CPU_ONLY)) for i, n in enumerate(meta["node_names"])}

    assert dra(["node-1"], [_slice("slice-1", "node-1", "nvidia.com/gpu", 4)]) == {"node-1": (4, True)}
    assert dra(["node-1", "node-2"], [_slice("slice-1", "node-1", "nvidia.com/gpu", 4), _slice("slice-2", "node-2", "nvidia.com/gpu", 8)]) == \
        {"node-1": (4, True), "node-2": (8, True)}
    assert dra(["node-1"], []) == {"node-1": (0, False)}
    assert dra(["node-1"], [_slice("slice-nvidia", "node-1", "nvidia.com/gpu", 4), _slice("slice-amd", "node-1", "amd.com/gpu", 2)]) == \
        {"node-1": (6, True)}
    assert dra(["node-1"], [_slice("slice-net", "node-1", "example.com/nic", 3)]) == {"node-1": (0, False)}


def test_device_plugin_gpu_pods_avoid_dra_nodes():  # node_info.go:326-333 PredicateByNodeResourcesType
    doc = _cluster(nodes=[_node("dra-node", gpus=0), _node("plugin-node", gpus=2)], queues=[_queue("q", gpu=(-1, -1, 1))],
                   pods=[_pod("gpu-pod", "g"), _pod("cpu-pod", "c", requests={"cpu": "1"})],
                   pod_groups=[{"metadata": {"name": g, "namespace": "ns"}, "spec": {"queue": "q"}} for g in ("g", "c")])
    doc["rawObjects"]["resourceSlices"] = [_slice("s", "dra-node", "gpu.nvidia.com", 8)]
    snap, meta, kw, _ = sio.pack_cluster(doc)
    t = {n: i for i, n in enumerate(meta["task_names"])}
    gpu_class, cpu_class = snap.task_pred_class[t["gpu-pod"]], snap.task_pred_class[t["cpu-pod"]]
    assert gpu_class >= 0 and int(snap.pred_mask[gpu_class, 0]) == 0b10  # only plugin-node (index 1)
    assert cpu_class < 0 or int(snap.pred_mask[cpu_class, 0]) == 0b11
    o = Oracle(abi.make_config(**kw))
    o.load(snap)
    res = o.run("allocate")
    assert meta["node_names"][res.task_node[t["gpu-pod"]]] == "plugin-node"  # binpack would otherwise prefer... the DRA node is excluded


# ---------------------------------------------------------------------------------------------- seeded fuzz
def _random_topology(rng):
    n_nodes, n_queues = int(rng.integers(1, 6)), int(rng.integers(1, 4))
    nodes = {f"node{i}": {"GPUs": int(rng.integers(0, 9))} for i in range(n_nodes)}
    departments = [{"Name": f"dept{d}", "DeservedGPUs": int(rng.integers(1, 12))} for d in range(int(rng.integers(1, 3)))]
    queues = [{"Name": f"queue{q}", "ParentQueue": departments[int(rng.integers(0, len(departments)))]["Name"],
               "DeservedGPUs": int(rng.integers(0, 6)), "GPUOverQuotaWeight": int(rng.integers(0, 3)),
               "MaxAllowedGPUs": int(rng.choice([0, 0, 4, 8]))} for q in range(n_queues)]
    free = {n: v["GPUs"] for n, v in nodes.items()}
    jobs = []
    for j in range(int(rng.integers(1, 9))):
        gpus = int(rng.choice([0, 1, 1, 2, 4]))
        tasks = []
        for _ in range(int(rng.integers(1, 4))):
            fits = [n for n, f in free.items() if f >= gpus]
            if fits and rng.random() < 0.4:
                node = fits[int(rng.integers(0, len(fits)))]
                free[node] -= gpus
                tasks.append({"State": str(rng.choice(["Running", "Running", "Releasing"])), "NodeName": node})
            else:
                tasks.append({"State": "Pending"})
        job = {"Name": f"job{j}", "QueueName": queues[int(rng.integers(0, n_queues))]["Name"],
               "Priority": int(rng.choice([50, 50, 75, 100, 125])), "RequiredGPUsPerTask": gpus, "Tasks": tasks}
        if gpus == 0:
            job["RequiredCPUsPerTask"] = float(rng.choice([0.5, 1, 2]))
        if rng.random() < 0.3:
            job["RootSubGroupSet"] = {"podsets": [{"name": dsl.DEFAULT_SUBGROUP, "min_available": int(rng.integers(1, len(tasks) + 1))}]}
        jobs.append(job)
    # Keep the job index order equal on both sides of the trip (the DSL orders jobs by priority, the packer by name): where
    # the reference iterates a Go map of jobs the restatements go by ascending index, and that order can decide — e.g.
    # proportion.reclaimableFn appends the victims' resources of a queue in map order (proportion.go:143-160) and
    # Reclaimable checks the running remainder before each subtraction, so victims of different sizes make the verdict
    # order-dependent in the reference itself (seed 1144 of this generator hits it).
    jobs.sort(key=lambda j: -j["Priority"])
    for i, job in enumerate(jobs):
        job["Name"] = f"job{i:02d}"
    return {"Nodes": nodes, "Queues": queues, "Departments": departments, "Jobs": jobs}


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_round_trip_keeps_the_outcome(seed):
    """Random small clusters (running / releasing / pending pods, gangs, priorities, departments, limits): the cluster
    rebuilt from its own raw-object dump must schedule exactly like the original, action after action."""
    rng = np.random.default_rng(1000 + seed)
    topo = _random_topology(rng)
    snap, meta = dsl.build_snapshot(topo)
    actions = ["allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"]
    snap2, meta2, kw, actions2 = _round_trip(snap, actions, names=meta,
                                             config={"allow_consolidating_reclaim": True, "max_consolidation_preemptees": -1})
    assert actions2 == actions
    a, b = Oracle(abi.make_config()), Oracle(abi.make_config(**kw))
    a.load(snap)
    b.load(snap2)
    status_names = {v: k for k, v in abi.POD_STATUS_NAMES.items()}

    def outcome(res, m):
        return {n: (m["node_names"][res.task_node[t]] if res.task_node[t] >= 0 else "", status_names[int(res.task_status[t])])
                for t, n in enumerate(m["task_names"])}

    for act in actions:
        assert outcome(a.run(act), meta) == outcome(b.run(act), meta2), act
