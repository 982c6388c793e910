"""ctypes binding of oracle/_ref/libmsb_ref.so: the CPU loop through the reference's own
plugin API (models/base.hpp, recarray row_accessor), built by `make -C oracle ref` in the
build container.  Test / bench infrastructure only.  Returns None when it was not built."""
import ctypes as C
import os

import numpy as np

import oracle_lib as ol

LIB = os.path.join(ol.ORACLE_DIR, "_ref", "libmsb_ref.so")
_P, _SZ = C.c_void_p, C.c_size_t


class RefHarness(object):
    def __init__(self, lib):
        self.lib = lib
        lib.ref_score_rows.restype = C.c_int
        lib.ref_score_rows.argtypes = [C.POINTER(ol.OrcModel), _SZ, _P, _P, _SZ, _P, _P, _P, C.POINTER(ol.OrcType), _SZ, _SZ, C.c_int, _P]
        lib.ref_perf_group.restype = C.c_double
        lib.ref_perf_group.argtypes = [C.c_int, C.c_uint, _SZ, _SZ, C.POINTER(C.c_double)]

    def score_rows(self, descs, hp_flat, ss, logprior, view, nthreads=1, row_lo=0, row_hi=None):
        orc = ol.load()
        models, types, data, mask = orc._pack(descs, view)
        row_hi = view.size() if row_hi is None else row_hi
        K = ss.shape[0]
        hp_flat = np.ascontiguousarray(hp_flat, np.float64); ss = np.ascontiguousarray(ss, np.float64)
        lp = np.ascontiguousarray(logprior, np.float64)
        out = np.zeros((row_hi - row_lo, K), np.float32)
        rc = self.lib.ref_score_rows(models, len(descs), hp_flat.ctypes.data, ss.ctypes.data, K, lp.ctypes.data,
                                     data.ctypes.data, mask.ctypes.data if mask is not None else None, types,
                                     row_lo, row_hi, nthreads, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("ref_score_rows failed")
        return out

    def noop_overhead_ns(self, D=1000, niters=20000):
        """ns per call of the reference's own noop model through its plugin API (models/noop.hpp)"""
        self.lib.ref_noop_overhead_ns.restype = C.c_double
        self.lib.ref_noop_overhead_ns.argtypes = [_SZ, _SZ]
        return self.lib.ref_noop_overhead_ns(D, niters)

    def perf_group(self, family=ol.BB, dim=0, D=1000, niters=2000):
        s = C.c_double()
        return self.lib.ref_perf_group(family, dim, D, niters, C.byref(s)), s.value


_ref = False


def load():
    global _ref
    if _ref is False:
        _ref = RefHarness(C.CDLL(LIB)) if os.path.exists(LIB) else None
    return _ref
