"""The C++ host side above the C ABI (host/kai_host.hpp): the reference's Action / Session surface mirrored in C++
(the reference's host language, Go, has no toolchain in this image).  The driver host/kai_host_test builds a
framework::Session from a cluster description, resolves the Actions by name from the registry and executes them, like
pkg/scheduler/test_utils RunTests; the expectations are the reference's own tables.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import dsl
from fixtures import action_cases, case_needs_predicates
from kai_scheduler_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "kai_host_test")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "kai_scheduler_b200", "csrc")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")], stdout=subprocess.DEVNULL)


def test_registry_resolves_default_actions():
    """framework.RegisterAction / GetAction semantics (framework/plugins.go:47-62): no GPU needed."""
    _build()
    out = subprocess.run([BIN, "--registry"], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[:6] == ["allocate registered", "consolidation registered", "reclaim registered", "preempt registered",
                       "stalegangeviction registered", "unknown missing"]


def write_case(path, snap: abi.Snapshot, meta: dict, actions, topo=None, cfg: dict | None = None):
    R, N, Q = snap.n_res, snap.n_nodes, int(snap.queue_parent.shape[0])
    qn = meta["queue_names"]
    with open(path, "w") as f:
        f.write(f"R {R}\n")
        for q in range(Q):
            parent = qn[snap.queue_parent[q]] if snap.queue_parent[q] >= 0 else "-"
            vals = " ".join(repr(float(x)) for tab in (snap.queue_deserved, snap.queue_limit, snap.queue_oqw) for x in tab[:, q])
            f.write(f"queue {qn[q]} {parent} {int(snap.queue_priority[q])} {int(snap.queue_creation[q])} {vals}\n")
        for n in range(N):
            f.write(f"node {meta['node_names'][n]} " + " ".join(repr(float(x)) for x in snap.node_allocatable[:, n]) + "\n")
        for tp in (topo or {}).get("Topologies") or []:
            f.write(f"topology {tp['ObjectMeta']['Name']} " + " ".join(lv["NodeLabel"] for lv in tp["Spec"]["Levels"]) + "\n")
        for nname, nd in ((topo or {}).get("Nodes") or {}).items():
            for k_, v_ in (nd.get("Labels") or {}).items():
                k_ = "kai.scheduler/type" if k_ == "tasks_fake.NodeAffinityKey" else k_
                f.write(f"label {nname} {k_} {v_}\n")
        for j, name in enumerate(meta["job_names"]):
            pre = 1 if snap.job_flags[j] & abi.JOB_PREEMPTIBLE else 0
            f.write(f"job {name} {qn[snap.job_queue[j]]} {int(snap.job_priority[j])} {pre} {int(snap.job_order_rank[j])}\n")
            jdef = next((jd for jd in (topo or {}).get("Jobs", []) if jd["Name"] == name), {})
            root = jdef.get("RootSubGroupSet") or {}
            tree = root.get("tree")
            ps_parent, ps_tc = {}, {}

            def tc_fields(tc):
                if not tc or not tc.get("Topology"):
                    return "- - -"
                return f"{tc['Topology']} {tc.get('RequiredLevel') or '-'} {tc.get('PreferredLevel') or '-'}"

            if root.get("topology_constraint"):
                f.write(f"rootconstraint {name} {tc_fields(root['topology_constraint'])}\n")
            if tree:
                if tree.get("constraint"):
                    f.write(f"rootconstraint {name} {tc_fields(tree['constraint'])}\n")

                def walk(g, parent):
                    for p_ in g["podsets"]:
                        ps_parent[p_["name"]] = "-" if parent is None else g["name"]
                        ps_tc[p_["name"]] = p_.get("constraint")
                    for c in g["groups"]:
                        f.write(f"set {name} {c['name']} {'-' if parent is None else g['name']} {tc_fields(c.get('constraint'))}\n")
                        walk(c, g)

                walk(tree, None)
            # PodSets in name order (= snapshot order); the engine-side name keeps that order
            real_names = sorted(ps_parent) if ps_parent else []
            for ps in range(snap.job_podset_begin[j], snap.job_podset_begin[j + 1]):
                k = ps - snap.job_podset_begin[j]
                extra = ""
                if real_names and len(real_names) == snap.job_podset_begin[j + 1] - snap.job_podset_begin[j]:
                    rn = real_names[k]
                    extra = f" {ps_parent[rn]} {tc_fields(ps_tc[rn])}"
                f.write(f"podset {name} ps{k:03d} {int(snap.podset_min_available[ps])}{extra}\n")
        for j, name in enumerate(meta["job_names"]):
            for ps in range(snap.job_podset_begin[j], snap.job_podset_begin[j + 1]):
                for t in range(snap.podset_task_begin[ps], snap.podset_task_begin[ps + 1]):
                    node = meta["node_names"][snap.task_node[t]] if snap.task_node[t] >= 0 else "-"
                    req = " ".join(repr(float(x)) for x in snap.task_req[t])
                    f.write(f"task {name} ps{ps - snap.job_podset_begin[j]:03d} {meta['task_names'][t]} {int(snap.task_status[t])} "
                            f"{node} {int(snap.task_order_rank[t])} {req}\n")
        for jd in (topo or {}).get("Jobs") or []:
            for k, t in enumerate(jd.get("Tasks") or []):
                if t.get("NodeAffinityNames"):
                    f.write(f"affinity {jd['Name']}-{k} " + " ".join(t["NodeAffinityNames"]) + "\n")
        cfg = cfg or {}
        f.write(f"conf {cfg.get('gpu_placement', 0)} {cfg.get('cpu_placement', 0)} {cfg.get('default_reclaim_min_runtime_s', 0.0)!r} "
                f"{cfg.get('default_preempt_min_runtime_s', 0.0)!r} {cfg.get('reclaim_resolve_method', 0)} {float(snap.now_s)!r}\n")
        if snap.queue_preempt_min_runtime_s is not None:
            for q in range(Q):
                f.write(f"queuemrt {qn[q]} {float(snap.queue_preempt_min_runtime_s[q])!r} {float(snap.queue_reclaim_min_runtime_s[q])!r}\n")
        if snap.job_last_start_s is not None:
            for j, name in enumerate(meta["job_names"]):
                f.write(f"jobstart {name} {float(snap.job_last_start_s[j])!r}\n")
        if snap.job_stale_since_s is not None:
            for j, name in enumerate(meta["job_names"]):
                f.write(f"jobstale {name} {float(snap.job_stale_since_s[j])!r}\n")
        if "staleness_grace_period_s" in cfg:
            f.write(f"grace {int(cfg['staleness_grace_period_s'])}\n")
        f.write("actions " + " ".join(actions) + "\n")


TOPO_CASES = action_cases(["allocate__allocateTopology"], single_action="allocate")


@pytest.mark.parametrize("cid,case", TOPO_CASES, ids=[c[0] for c in TOPO_CASES])
def test_cpp_packing_of_topologies_and_subgroup_tree(cid, case):
    """packSnapshot (C++) against tests/dsl.py on the topology tables: per-level domain ids of every node and the
    SubGroupSet tree with its constraints (no GPU: --dump-packed stops before the engine)."""
    _build()
    snap, meta = dsl.build_snapshot(case["topology"])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "case.txt")
        write_case(path, snap, meta, case["actions"], case["topology"])
        out = subprocess.run([BIN, path, "--dump-packed"], capture_output=True, text=True, check=True).stdout.split("\n")
    nidx = {n: i for i, n in enumerate(meta["node_names"])}
    jidx = {n: i for i, n in enumerate(meta["job_names"])}
    seen_sets = 0
    for line in out:
        f = line.split()
        if not f:
            continue
        if f[0] == "node_domain":
            assert int(f[3]) == snap.node_domain[int(f[1]), nidx[f[2]]], line
        elif f[0] == "set":
            j = jidx[f[1]]
            g = snap.job_sgs_begin[j] + int(f[2])
            par = snap.sgs_parent[g]
            assert int(f[4]) == (-1 if par < 0 else par - snap.job_sgs_begin[j]), line
            assert [int(x) for x in f[8:11]] == [snap.sgs_topology[g], snap.sgs_required_level[g], snap.sgs_preferred_level[g]], line
            # name rank: same relative order among the nested sets of the job (the root's own rank is irrelevant)
            seen_sets += 1
        elif f[0] == "podset":
            j = jidx[f[1]]
            ps = snap.job_podset_begin[j] + int(f[2])
            assert int(f[4]) == snap.podset_min_available[ps], line
            assert int(f[6]) == snap.podset_sgs[ps] - snap.job_sgs_begin[j], line
            assert [int(x) for x in f[8:11]] == [snap.podset_topology[ps], snap.podset_required_level[ps], snap.podset_preferred_level[ps]], line
    assert seen_sets == len(snap.sgs_parent)
    if case_needs_predicates(case):
        masks = {}
        for line in out:
            f = line.split()
            if f and f[0] == "pred":  # pred <pod> <class> <mask words...>
                masks[f[1]] = [int(x) for x in f[3:]]
        for t, name in enumerate(meta["task_names"]):
            c = snap.task_pred_class[t]
            want = [] if c < 0 else [int(x) for x in snap.pred_mask[c]]
            assert masks.get(name, []) == want, name


STALE_CASES = action_cases(["stalegangeviction__"], single_action="stalegangeviction")


@pytest.mark.parametrize("cid,case", STALE_CASES, ids=[c[0] for c in STALE_CASES])
def test_cpp_packing_of_staleness(cid, case):
    """PodGroupInfo.StalenessInfo.TimeStamp and the grace period reach the C ABI unchanged (no GPU)."""
    _build()
    snap, meta = dsl.build_snapshot(case["topology"])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "case.txt")
        write_case(path, snap, meta, case["actions"], case["topology"], cfg=case["config"])
        out = subprocess.run([BIN, path, "--dump-packed"], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {f[1]: (float(f[2]), float(f[4]), int(f[6])) for f in (l.split() for l in out) if f and f[0] == "stale"}
    want = {}
    if snap.job_stale_since_s is not None:
        want = {n: (float(snap.job_stale_since_s[j]), float(snap.now_s), 60) for j, n in enumerate(meta["job_names"])
                if snap.job_stale_since_s[j] > 0}
    assert got == want


class _Res:
    pass


CASES = (TOPO_CASES + action_cases(["allocate__allocate_subgroups"], single_action="allocate")[-2:]
         + action_cases(["allocate__allocate"], single_action="allocate")[:12]
         + [c for c in action_cases(["allocate__", "reclaim__"]) if c[1].get("config") or case_needs_predicates(c[1])][:4]
         + action_cases(["reclaim__"], single_action="reclaim")[:12]
         + action_cases(["consolidation__"], single_action="consolidation")[:8] + action_cases(["preempt__"], single_action="preempt")[:8])


@pytest.mark.gpu
@pytest.mark.parametrize("cid,case", CASES, ids=[c[0] for c in CASES])
def test_reference_tables_through_cpp_shim(cid, case):
    _build()
    snap, meta = dsl.build_snapshot(case["topology"])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "case.txt")
        place = {"binpack": 0, "spread": 1}
        write_case(path, snap, meta, case["actions"], case["topology"],
                   cfg={k: place[v] for k, v in (case.get("config") or {}).items()})
        out = subprocess.run([BIN, path], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().split("\n")]
    by_name = {l[0]: l for l in lines if l[0] != "cache"}
    res = _Res()
    res.task_status = np.array([int(by_name[n][1]) for n in meta["task_names"]], dtype=np.int32)
    nidx = {n: i for i, n in enumerate(meta["node_names"])}
    res.task_node = np.array([nidx.get(by_name[n][2], -1) for n in meta["task_names"]], dtype=np.int32)
    topo = dict(case["topology"])
    topo.pop("ExpectedNodesResources", None)  # the C++ session replays statuses / bindings; node tables stay in the engine
    res.node_idle = res.node_releasing = None
    errs = dsl.check_expectations(topo, meta, res, snap)
    assert not errs, f"{case['source']} #{case['index']}: {errs}"
    # CacheMocking caps (test_utils.go CacheRequirements): binds / evictions / pipelines must not exceed them
    cache = [l for l in lines if l[0] == "cache"][0]
    caps = ((case["topology"].get("Mocks") or {}).get("CacheRequirements") or {})
    for key, got in (("NumberOfCacheBinds", int(cache[1])), ("NumberOfCacheEvictions", int(cache[2])),
                     ("NumberOfPipelineActions", int(cache[3]))):
        if caps.get(key) is not None:
            assert got <= caps[key], (key, got, caps[key])


import minruntime_cases as mc  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("case", mc.CASES, ids=[c[0] for c in mc.CASES])
def test_min_runtime_through_cpp_shim(case):
    """QueueInfo.{Preempt,Reclaim}MinRuntime, PodGroupInfo.LastStartTimestamp and the plugin arguments packed by the C++
    mirror: the replayed session must end where the oracle ends."""
    from oracle_lib import Oracle
    _build()
    first = None
    if "@first" in case[5]:
        snap, meta, cfg = mc.build(next(c for c in mc.CASES if c[0] == "reclaim-unprotected"))
        o = Oracle(cfg)
        o.load(snap)
        first = [n for n, st in mc.outcome(o.run("reclaim"), meta).items() if st == "Releasing"][0].rsplit("-", 1)[0]
    snap, meta, cfg = mc.build(case, first)
    o = Oracle(cfg)
    o.load(snap)
    want = o.run(case[2])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "case.txt")
        write_case(path, snap, meta, [case[2]], case[1], cfg=case[3])
        out = subprocess.run([BIN, path], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    by_name = {l.split()[0]: l.split() for l in out.stdout.strip().split("\n") if not l.startswith("cache")}
    got_status = [int(by_name[n][1]) for n in meta["task_names"]]
    assert got_status == [int(x) for x in want.task_status]
