"""ctypes loader for the CPU oracle (oracle/libkaioracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kai_scheduler_b200 import abi  # noqa: E402

_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libkaioracle.so")
        src = os.path.join(ROOT, "oracle", "kai_oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libkaioracle.so"],
                                  stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(path)
        abi.bind_engine_api(_LIB, "kai_oracle")
        _LIB.kai_oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        _LIB.kai_oracle_last_error.argtypes = [C.c_void_p]
        _LIB.kai_oracle_last_error.restype = C.c_char_p
        _LIB.kai_oracle_binpack_score.argtypes = [C.c_double] * 4
        _LIB.kai_oracle_binpack_score.restype = C.c_double
        _LIB.kai_oracle_spread_score.argtypes = [C.c_double] * 2
        _LIB.kai_oracle_spread_score.restype = C.c_double
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        _LIB.kai_oracle_set_resource_share.argtypes = [C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, dp, ip, lp,
                                                       ip, dp]
        _LIB.kai_oracle_set_resource_share.restype = C.c_double
        _LIB.kai_oracle_queue_order.argtypes = [dp, dp, C.c_int, C.c_int, C.c_int64, C.c_int64, dp, dp, dp]
        _LIB.kai_oracle_queue_order.restype = C.c_int
        _LIB.kai_oracle_min_runtime.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _LIB.kai_oracle_min_runtime.restype = C.c_double
        _LIB.kai_oracle_min_runtime_protected.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _LIB.kai_oracle_min_runtime_protected.restype = C.c_int
    return _LIB


class Oracle:
    """Same call surface as kai_scheduler_b200.engine.Engine, backed by the CPU oracle."""

    def __init__(self, cfg: abi.KaiConfig | None = None, threads: int = 1):
        self._lib = lib()
        self._cfg = cfg or abi.make_config()
        self._h = C.c_void_p()
        rc = self._lib.kai_oracle_create(C.byref(self._cfg), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"kai_oracle_create failed: {rc}")
        self._lib.kai_oracle_set_threads(self._h, threads)
        self._n_res = 4

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle error {rc}: {self._lib.kai_oracle_last_error(self._h).decode()}")

    def load(self, snap: abi.Snapshot):
        c = snap.to_c()
        self._n_res = snap.n_res
        self._check(self._lib.kai_oracle_load_snapshot(self._h, C.byref(c)))

    def run(self, action) -> abi.Result:
        a = abi.ACTIONS[action] if isinstance(action, str) else action
        r = abi.KaiResult()
        self._check(self._lib.kai_oracle_run(self._h, a, C.byref(r)))
        return abi.Result.from_c(r, self._n_res)

    def fair_share(self) -> abi.Result:
        r = abi.KaiResult()
        self._check(self._lib.kai_oracle_fair_share(self._h, C.byref(r)))
        return abi.Result.from_c(r, self._n_res)

    def min_runtime(self, reclaim: bool, pending_queue: int, victim_queue: int) -> float:
        return self._lib.kai_oracle_min_runtime(self._h, int(reclaim), pending_queue, victim_queue)

    def min_runtime_protected(self, reclaim: bool, pending_job: int, victim_job: int) -> bool:
        return bool(self._lib.kai_oracle_min_runtime_protected(self._h, int(reclaim), pending_job, victim_job))

    def node_entries(self):
        """(task, node, status) of every clone the nodes hold (NodeInfo.PodInfos)."""
        import numpy as np
        fn = self._lib.kai_oracle_node_entries
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3 + [C.c_int]
        cap = 1 << 16
        while True:
            t, n, s = (np.zeros(cap, dtype=np.int32) for _ in range(3))
            k = fn(self._h, *(a.ctypes.data_as(C.POINTER(C.c_int32)) for a in (t, n, s)), cap)
            if k <= cap:
                return list(zip(t[:k].tolist(), n[:k].tolist(), s[:k].tolist()))
            cap = k

    def stats(self) -> abi.KaiStats:
        s = abi.KaiStats()
        self._check(self._lib.kai_oracle_stats(self._h, C.byref(s)))
        return s

    def close(self):
        if self._h:
            self._lib.kai_oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
