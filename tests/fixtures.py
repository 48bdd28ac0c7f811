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
