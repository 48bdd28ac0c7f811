"""Loading of the transcribed reference tables (tests/golden/actions/*.json)."""
import glob
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def action_cases(prefixes, single_action=None):
    """Yield (id, case) for in-scope cases of the suites whose file name starts with one of `prefixes`."""
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN, "actions", "*.json"))):
        base = os.path.basename(path)[:-5]
        if not any(base.startswith(p) for p in prefixes):
            continue
        for case in json.load(open(path)):
            if not case["supported"]:
                continue
            if single_action and case["actions"] != [single_action]:
                continue
            out.append((f"{base}[{case['index']}]", case))
    return out


def case_config(case, **overrides):
    """kai_config for a table: the reference's test defaults plus what the table's own SchedulerConf sets
    (gen_fixtures keeps `nodeplacement` arguments as case["config"])."""
    from kai_scheduler_b200 import abi
    place = {"binpack": abi.PLACEMENT_BINPACK, "spread": abi.PLACEMENT_SPREAD}
    kw = {}
    for key, value in (case.get("config") or {}).items():
        kw[key] = place[value] if key.endswith("_placement") else value
    kw.update(overrides)
    return abi.make_config(**kw)


def case_needs_predicates(case) -> bool:
    """Tables whose pods carry node affinity: they need pred_mask classes (not in the C++ mirror's text format)."""
    return any(t.get("NodeAffinityNames") for j in case["topology"].get("Jobs") or [] for t in j.get("Tasks") or [])
