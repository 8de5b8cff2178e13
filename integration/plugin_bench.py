"""ctypes access to integration/_build/libplugin_bench.so: n panel-major ``ocp_qp_in`` objects (the reference's structs) handed to
the plugin's batched entry ``ocp_qp_cuipm_batch_solve`` -- the call a user of the patched libacados makes.  Used by bench.py
(``e2e_plugin``) and tests/test_plugin.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libplugin_bench.so")


def available() -> bool:
    return os.path.exists(LIB)


class PluginBatch:
    def __init__(self, batch, opts):
        self.lib = C.CDLL(LIB)
        self.batch = batch
        self._shape = batch.shape.as_ctypes()
        self.lib.plugin_bench_create.restype = C.c_void_p
        self.h = self.lib.plugin_bench_create(C.byref(self._shape), C.c_int(batch.nbatch), C.c_void_p(batch.qp.ctypes.data), C.byref(opts))
        if not self.h:
            raise RuntimeError("plugin_bench_create failed")

    def run(self, reps: int = 1):
        """``reps`` calls of the batched entry; returns (worst acados status, seconds per call)."""
        sec = np.zeros(reps)
        self.lib.plugin_bench_run.restype = C.c_int
        st = self.lib.plugin_bench_run(C.c_void_p(self.h), C.c_int(reps), C.c_void_p(sec.ctypes.data))
        return st, sec

    def solutions(self):
        nb = self.batch.nbatch
        sol = self.batch.layout.new_sol(nb)
        it = np.zeros(nb, dtype=np.int32)
        st = np.zeros(nb, dtype=np.int32)
        self.lib.plugin_bench_get(C.c_void_p(self.h), C.byref(self._shape), C.c_void_p(sol.ctypes.data), C.c_void_p(it.ctypes.data),
                                  C.c_void_p(st.ctypes.data))
        return sol, it, st

    def close(self):
        if self.h:
            self.lib.plugin_bench_destroy(C.c_void_p(self.h))
            self.h = None
