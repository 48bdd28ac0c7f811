"""Replay one scheduling cycle from a snapshot zip on the GPU engine.

The counterpart of the reference's `cmd/snapshot-tool` (main.go:33-121: `-filename snapshot.zip`, load the snapshot,
open a session, run the configured actions once): same input file, same action list from the file's own
`config.actions`, the engine behind the C ABI instead of the stock actions.  Prints one JSON line per action
(latency, pods placed / evicted) and, with --bindings, the resulting pod -> node / status table.

    python -m kai_scheduler_b200.snapshot_tool --filename snapshot.zip [--bindings] [--device 0]

No CPU fallback: without libkaigpu.so or a CUDA device this fails.
"""
from __future__ import annotations

import argparse
import json
import sys
import time

from . import abi, engine, snapshot_io


def replay(doc: dict, device: int = 0, strict: bool = True):
    """-> (per-action records, final Result, meta).  Session state carries from one action to the next."""
    snap, meta, kw, actions = snapshot_io.pack_cluster(doc, strict=strict)
    unknown = [a for a in actions if a not in abi.ACTIONS]
    if unknown:
        raise snapshot_io.UnsupportedSnapshot(f"actions {unknown}")
    eng = engine.Engine(abi.make_config(device=device, **kw))
    try:
        t0 = time.perf_counter()
        eng.load(snap)
        records = [{"phase": "load_snapshot", "ms": (time.perf_counter() - t0) * 1e3, "nodes": snap.n_nodes,
                    "queues": snap.n_queues, "pod_groups": snap.n_jobs, "pods": snap.n_tasks}]
        res = None
        for a in actions:
            t0 = time.perf_counter()
            res = eng.run(a)
            records.append({"action": a, "ms": (time.perf_counter() - t0) * 1e3, "pods_placed": int(res.pods_placed),
                            "pods_evicted": int(res.pods_evicted), "jobs_visited": len(res.visits)})
        return records, res, meta
    finally:
        eng.close()


def bindings(res, meta):
    status_names = {v: k for k, v in abi.POD_STATUS_NAMES.items()}
    out = []
    for t, name in enumerate(meta["task_names"]):
        n = int(res.task_node[t])
        out.append({"pod": name, "pod_group": meta["job_names"][int(meta["task_job"][t])],
                    "node": meta["node_names"][n] if n >= 0 else "", "status": status_names[int(res.task_status[t])]})
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--filename", "-filename", required=True, help="snapshot zip (plugins/snapshot format)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--bindings", action="store_true", help="print the pod -> node table after the last action")
    ap.add_argument("--lenient", action="store_true", help="ignore (and list) pod constraints outside the packer's scope")
    args = ap.parse_args(argv)
    doc = snapshot_io.read_snapshot_zip(args.filename)
    records, res, meta = replay(doc, device=args.device, strict=not args.lenient)
    for r in records:
        print(json.dumps(r))
    if meta["ignored"]:
        print(json.dumps({"ignored": meta["ignored"]}))
    if args.bindings and res is not None:
        for b in bindings(res, meta):
            print(json.dumps(b))
    return 0


if __name__ == "__main__":
    sys.exit(main())
