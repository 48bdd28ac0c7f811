#!/usr/bin/env python
"""Transcribe the reference's Go table tests into JSON fixtures (run in the build container only).

    python tests/golden/gen_fixtures.py            # rewrites tests/golden/*.json

Reads /root/reference (read-only) and writes small JSON files next to this script.
The fixtures, not the reference, travel to the GPU box.  Sources transcribed:

  actions/*.json          pkg/scheduler/actions/{allocate,reclaim,consolidation}/*_test.go and
                          pkg/scheduler/actions/integration_tests/{allocate,reclaim,consolidation,
                          consolidation_and_reclaim}/*_test.go  (test_utils.TestTopologyBasic tables:
                          cluster in -> node name / status per job out)
  nodepack.json           plugins/nodeplacement/nodepack_test.go, nodespread_test.go
  resource_division.json  plugins/proportion/resource_division/resource_division_test.go

Every case keeps the Go literal's field names; identifiers are resolved to their
values.  Cases using features outside the engine's scope (fractional GPUs, MIG,
GPU memory, DRA, pod/node affinity, custom scheduler conf) are kept but marked
`"supported": false` with a reason, so that the skipped set is explicit.
"""
from __future__ import annotations

import glob
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from go_literal import find_literals  # noqa: E402

REF = "/root/reference/pkg/scheduler"

IDENTS = {
    "constants.PriorityTrainNumber": 50,
    "constants.PriorityInteractivePreemptibleNumber": 75,
    "constants.PriorityBuildNumber": 100,
    "constants.PriorityInferenceNumber": 125,
    "common_info.NoMaxAllowedResource": -1.0,
    "commonconstants.UnlimitedResourceQuantity": -1.0,
    "enginev2alpha2.Preemptible": "preemptible",
    "enginev2alpha2.NonPreemptible": "non-preemptible",
    "v2alpha2.Preemptible": "preemptible",
    "v2alpha2.NonPreemptible": "non-preemptible",
    "node_info.MigStrategySingle": "single",
    "node_info.MigStrategyMixed": "mixed",
    "node_info.MigStrategyNone": "",
    "podgroup_info.DefaultSubGroup": "default-sub-group",
    "constants.DefaultQueuePriority": 100,
    "commonconstants.DefaultQueuePriority": 100,
    "time.Minute": 60.0,
    "time.Second": 1.0,
    "time.Hour": 3600.0,
}


def resolve(x):
    if isinstance(x, dict):
        if "__ident" in x and len(x) == 1:
            name = x["__ident"]
            if name.startswith("pod_status."):
                return name.split(".", 1)[1]
            if name in IDENTS:
                return IDENTS[name]
            return {"__ident": name}
        if "__binop" in x:
            l, r = resolve(x["l"]), resolve(x["r"])
            if isinstance(l, (int, float)) and isinstance(r, (int, float)):
                return {"*": l * r, "+": l + r, "-": l - r, "/": l / r if r else 0}[x["__binop"]]
            if isinstance(l, str) and isinstance(r, str) and x["__binop"] == "+":
                return l + r
            return {"__binop": x["__binop"], "l": l, "r": r}
        if "__call" in x:
            if x["__call"] == "jobs_fake.DefaultSubGroup":
                return {"podsets": [{"name": "default-sub-group", "min_available": resolve(x["args"][0])}]}
            if x["__call"].startswith("test_utils.Create") and len(x["args"]) == 1:
                return resolve(x["args"][0])
            if x["__call"] == "subgroup_info.NewSubGroupSet":
                return {"__unsupported": "inline SubGroupSet constructor (topology constraint)"}
            return {"__call": x["__call"], "args": [resolve(a) for a in x["args"]]}
        if "__func" in x:
            return parse_subgroup_func(x["__func"])
        return {k: resolve(v) for k, v in x.items()}
    if isinstance(x, list):
        return [resolve(v) for v in x]
    return x


_PODSET = re.compile(r'NewPodSet \( "([^"]+)" , (\d+) , (nil|[^)]*)\)')


def parse_subgroup_func(body: str):
    """RootSubGroupSet: func() *SubGroupSet { root := NewSubGroupSet(...); root.AddPodSet(NewPodSet(name, min, nil)) ... }"""
    n_sets = body.count("NewSubGroupSet")
    podsets = []
    unsupported = None
    for m in _PODSET.finditer(body):
        podsets.append({"name": m.group(1), "min_available": int(m.group(2))})
        if m.group(3).strip() != "nil":
            unsupported = "podset topology constraint"
    if n_sets > 1:
        unsupported = "nested SubGroupSets"
    if "TopologyConstraint" in body:
        unsupported = "subgroup topology constraint"
    out = {"podsets": podsets}
    if unsupported or not podsets:
        out["__unsupported"] = unsupported or "unparsed RootSubGroupSet"
    return out


def find_unresolved(x, path=""):
    out = []
    if isinstance(x, dict):
        for k in ("__ident", "__call", "__unsupported", "__binop", "__neg"):
            if k in x:
                out.append(f"{path}:{k}={x[k] if not isinstance(x[k], (dict, list)) else '...'}")
        for k, v in x.items():
            out += find_unresolved(v, f"{path}.{k}")
    elif isinstance(x, list):
        for i, v in enumerate(x):
            out += find_unresolved(v, f"{path}[{i}]")
    return out


def classify(topo: dict) -> str | None:
    """Return a skip reason if the case uses features outside the engine's scope, else None."""
    if topo.get("Topologies"):
        return "topology CRs"
    mocks = topo.get("Mocks") or {}
    if isinstance(mocks, dict) and mocks.get("SchedulerConf"):
        return "custom SchedulerConf"
    if isinstance(mocks, dict) and mocks.get("GPUMetric"):
        return "GPU metric mocks"
    for key in ("TestDRAObjects", "ResourceClaims", "ResourceSlices", "DeviceClasses"):
        if topo.get(key):
            return "DRA objects"
    for name, node in (topo.get("Nodes") or {}).items():
        if node.get("MigStrategy") or node.get("MigInstances"):
            return "MIG node"
        if node.get("GpuMemorySynced") is not None or node.get("GPUMemory"):
            return "GPU memory"
        if node.get("Labels"):
            return "node labels"
    for job in topo.get("Jobs") or []:
        g = job.get("RequiredGPUsPerTask", 0) or 0
        if float(g) != int(g):
            return "fractional GPU"
        if job.get("RequiredGpuMemory") or job.get("RequiredMultiFractionDevicesPerTask"):
            return "GPU memory / multi-fraction"
        if job.get("DeleteJobInTest") or job.get("StaleDuration") is not None:
            return "deletion / staleness"
        for t in job.get("Tasks") or []:
            for k in ("NodeAffinityNames", "PodAffinityLabels", "PodAffinityTopologyKey",
                      "PodAntiAffinityTopologyKey", "RequiredMigInstances", "IsLegacyMigTask",
                      "ResourceClaimTemplates", "ResourceClaimNames", "GPUGroups"):
                if t.get(k):
                    return f"task {k}"
    for q in topo.get("Queues") or []:
        if q.get("V1") or q.get("UseOnlyFreeCPUResources") or q.get("InteractiveTimeoutInMinutes"):
            return "legacy queue fields"
    unresolved = find_unresolved(topo)
    if unresolved:
        return "unresolved: " + "; ".join(unresolved[:3])
    return None


ACTION_SUITES = [
    # (glob relative to actions/, action list run by the reference's test driver)
    ("allocate/allocate_test.go", ["allocate"]),
    ("allocate/allocateGang_test.go", ["allocate"]),
    ("allocate/allocateElastic_test.go", ["allocate"]),
    ("allocate/allocate_subgroups_test.go", ["allocate"]),
    ("allocate/allocateTopology_test.go", ["allocate"]),
    ("reclaim/reclaim_test.go", ["reclaim"]),
    ("reclaim/reclaimGang_test.go", ["reclaim"]),
    ("reclaim/reclaimDepartments_test.go", ["reclaim"]),
    ("reclaim/reclaim_elastic_test.go", ["reclaim"]),
    ("reclaim/reclaim_sub_group_test.go", ["reclaim"]),
    ("consolidation/consolidation_test.go", ["consolidation"]),
    ("consolidation/consolidation_subgroups_test.go", ["consolidation"]),
    # integration tables run the whole default action list for several rounds
    ("integration_tests/allocate/allocate_test.go", ["allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"]),
    ("integration_tests/reclaim/reclaim_test.go", ["allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"]),
    ("integration_tests/consolidation/consolidation_test.go", ["allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"]),
    ("integration_tests/consolidation_and_reclaim/consolidation_and_reclaim_test.go",
     ["allocate", "consolidation", "reclaim", "preempt", "stalegangeviction"]),
]


def gen_actions():
    os.makedirs(os.path.join(HERE, "actions"), exist_ok=True)
    summary = {}
    for rel, actions in ACTION_SUITES:
        path = os.path.join(REF, "actions", rel)
        src = open(path).read()
        # integration tables wrap the topology in TestTopologyMetadata{TestTopologyBasic: ..., RoundsUntilMatch: n}
        rounds = {}
        metas = find_literals(src, "integration_tests_utils.TestTopologyMetadata") if "TestTopologyMetadata" in src else []
        # find_literals returns the outermost `[]TestTopologyMetadata{...}` slice literal first; flatten
        flat = []
        for m in metas:
            if isinstance(m, list):
                flat += m
            else:
                flat.append(m)
        cases = []
        if flat and all(isinstance(m, dict) and "TestTopologyBasic" in m for m in flat):
            for m in flat:
                m = resolve(m)
                topo = m["TestTopologyBasic"]
                cases.append((topo, m.get("RoundsUntilMatch"), m.get("RoundsAfterMatch")))
        else:
            for topo in find_literals(src, "test_utils.TestTopologyBasic"):
                cases.append((resolve(topo), None, None))
        out = []
        for i, (topo, r_until, r_after) in enumerate(cases):
            topo.pop("__type", None)
            reason = classify(topo)
            out.append({
                "source": f"pkg/scheduler/actions/{rel}",
                "index": i,
                "name": topo.get("Name", ""),
                "actions": actions,
                "rounds_until_match": r_until,
                "rounds_after_match": r_after,
                "supported": reason is None,
                "skip_reason": reason,
                "topology": topo,
            })
        name = rel.replace("/", "__").replace("_test.go", "") + ".json"
        with open(os.path.join(HERE, "actions", name), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        summary[rel] = (len(out), sum(1 for c in out if c["supported"]))
    return summary


def main():
    s = gen_actions()
    for k, (n, ok) in s.items():
        print(f"{k}: {n} cases, {ok} in scope")


if __name__ == "__main__":
    main()
