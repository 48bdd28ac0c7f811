"""snapshot zip -> engine (kai_scheduler_b200/snapshot_tool.py) against the oracle on the same packed snapshot (GPU)."""
import os
import tempfile

import numpy as np
import pytest

import dsl
from fixtures import action_cases, case_needs_predicates
from kai_scheduler_b200 import abi, snapshot_io as sio, snapshot_tool, synthetic
from kai_scheduler_b200.engine import Engine
from oracle_lib import Oracle
from test_engine_gpu import assert_same
from test_snapshot_io import _handwritten

pytestmark = pytest.mark.gpu


def _both(doc):
    snap, meta, kw, actions = sio.pack_cluster(doc)
    e, o = Engine(abi.make_config(**kw)), Oracle(abi.make_config(**kw))
    e.load(snap)
    o.load(snap)
    for a in actions:
        assert_same(e.run(a), o.run(a))
    e.close()
    return actions


def test_handwritten_document_engine_vs_oracle():
    """5 resource columns, predicate classes from selectors / affinity / taints / node conditions, a nominated node,
    foreign pods, a topology constraint, spread placement, kValue 0.5: allocate then reclaim."""
    assert _both(_handwritten()) == ["allocate", "reclaim"]


@pytest.mark.parametrize("name", ["config1", "cycle5-small", "config4-small"])
def test_replay_tool_on_synthetic_zip(name, capsys):
    snap = synthetic.config_snapshot(name)
    actions = synthetic.CONFIG_ACTIONS.get(name, ["allocate"])
    doc = sio.dump_cluster(snap, actions=actions, config={"allow_consolidating_reclaim": True})
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "snapshot.zip")
        sio.write_snapshot_zip(path, doc)
        assert snapshot_tool.main(["--filename", path]) == 0
        out = capsys.readouterr().out
        assert out.count('"action"') == len(actions)
        records, res, meta = snapshot_tool.replay(sio.read_snapshot_zip(path))
    snap2, _, kw, _ = sio.pack_cluster(doc)
    o = Oracle(abi.make_config(**kw))
    o.load(snap2)
    for a in actions:
        ro = o.run(a)
    np.testing.assert_array_equal(res.task_node, ro.task_node)
    np.testing.assert_array_equal(res.task_status, ro.task_status)
    assert [r.get("action") for r in records[1:]] == list(actions)


TOPOLOGY_TABLES = [c for c in action_cases(["allocate__allocateTopology"], single_action="allocate")
                   if not case_needs_predicates(c[1])]  # predicate classes have no raw-object form in dump_cluster


@pytest.mark.parametrize("cid,case", TOPOLOGY_TABLES, ids=[c[0] for c in TOPOLOGY_TABLES])
def test_topology_tables_through_the_wire_format_gpu(cid, case):
    snap, meta = dsl.build_snapshot(case["topology"])
    doc = sio.dump_cluster(snap, actions=case["actions"], names=meta, config={"allow_consolidating_reclaim": True})
    _both(doc)
