"""The C-ABI library loads and exports every symbol include/kai_engine.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from kai_scheduler_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "kai_engine.h")).read()
    return sorted(set(re.findall(r"\b(kai_[a-z_]+)\s*\(", src)) - {"kai_engine_h"})


def test_header_symbols_exported():
    lib = engine.lib()
    declared = _declared_symbols()
    assert set(declared) == set(engine.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"libkaigpu.so does not export {name}"
    assert lib.kai_abi_version() == abi.KAI_ABI_VERSION


def test_struct_sizes_match_header():
    # field order/size sanity: the C side rejects a mismatching abi_version, sizes are checked here
    assert C.sizeof(abi.KaiConfig) == 4 * 4 + 8 * 2 + 4 * 7 + 4 + 8 * 2  # 4 bytes of padding before the doubles
    assert C.sizeof(abi.KaiSnapshot) == 8 * 4 + 30 * 8 + 2 * 4 + 5 * 8 + 2 * 4 + 10 * 8 + 5 * 8 + 8
    assert C.sizeof(abi.KaiJobVisit) == 8


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = abi.make_config()
    h = C.c_void_p()
    rc = engine.lib().kai_engine_create(C.byref(cfg), C.byref(h))
    assert rc == abi.ERR_NO_DEVICE
    with pytest.raises(engine.EngineError):
        engine.Engine()


def test_invalid_arguments():
    lib = engine.lib()
    cfg = abi.make_config()
    cfg.abi_version = 99
    h = C.c_void_p()
    assert lib.kai_engine_create(C.byref(cfg), C.byref(h)) == abi.ERR_INVALID
    assert lib.kai_engine_create(None, C.byref(h)) == abi.ERR_INVALID


def test_fractional_gpu_requests_are_refused():
    """Shared-GPU pods need tables the ABI does not carry: both libraries must refuse them, not approximate."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_lib import Oracle
    from kai_scheduler_b200 import synthetic
    snap = synthetic.benchmark_snapshot(4, 3, n_queues=1)
    snap.task_req = snap.task_req.copy()
    snap.task_req[1, 2] = 0.5
    o = Oracle()
    with pytest.raises(RuntimeError, match="fractional GPU"):
        o.load(snap)
