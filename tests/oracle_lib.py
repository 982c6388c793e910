"""ctypes binding of oracle/libmsb_oracle.so -- the CPU checker (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libmsb_oracle.so")

BB, BNB, GP, NICH, DD, NIW, BBNC, DM = range(8)
FAMILY = {"dm": DM, "bbnc": BBNC, "bb": BB, "bnb": BNB, "gp": GP, "nich": NICH, "dd": DD, "niw": NIW}


class OrcModel(C.Structure):
    _fields_ = [("family", C.c_int32), ("dim", C.c_uint32)]


class OrcType(C.Structure):
    _fields_ = [("prim", C.c_int32), ("n", C.c_uint32), ("vec", C.c_int32)]


_P, _SZ, _D = C.c_void_p, C.c_size_t, C.c_double


class Oracle(object):
    def __init__(self, lib):
        self.lib = lib
        lib.orc_hp_size.restype = _SZ; lib.orc_hp_size.argtypes = [C.POINTER(OrcModel)]
        lib.orc_ss_size.restype = _SZ; lib.orc_ss_size.argtypes = [C.POINTER(OrcModel)]
        lib.orc_score_value.restype = _D
        lib.orc_score_value.argtypes = [C.POINTER(OrcModel), _P, _P, _P, C.c_int]
        lib.orc_add_value.restype = None
        lib.orc_add_value.argtypes = [C.POINTER(OrcModel), _P, _P, _P, C.c_int]
        lib.orc_remove_value.restype = None
        lib.orc_remove_value.argtypes = [C.POINTER(OrcModel), _P, _P, _P, C.c_int]
        lib.orc_score_rows.restype = None
        lib.orc_score_rows.argtypes = [C.POINTER(OrcModel), _SZ, _P, _P, _SZ, _P, _P, _P, C.POINTER(OrcType), _SZ, _SZ, C.c_int, C.c_int, _P]
        lib.orc_score_rows_f32.restype = None
        lib.orc_score_rows_f32.argtypes = [C.POINTER(OrcModel), _SZ, _P, _P, _SZ, _P, _P, _P, C.POINTER(OrcType), _SZ, _SZ, C.c_int, _P]
        lib.orc_update_rows.restype = None
        lib.orc_update_rows.argtypes = [C.POINTER(OrcModel), _SZ, _P, _P, _SZ, _P, _P, _P, C.POINTER(OrcType), _SZ, _SZ, _P, _P, C.c_int]
        lib.orc_expf.restype = C.c_float; lib.orc_expf.argtypes = [C.c_float]
        lib.orc_sample_discrete_log.restype = C.c_int64
        lib.orc_sample_discrete_log.argtypes = [_P, _SZ, C.c_float]
        lib.orc_sample_rows.restype = None
        lib.orc_sample_rows.argtypes = [_P, _SZ, _SZ, _SZ, _P, _P]
        lib.orc_philox_u01.restype = C.c_float; lib.orc_philox_u01.argtypes = [C.c_uint64] * 3
        lib.orc_sample_rows_libm.restype = None
        lib.orc_sample_rows_libm.argtypes = [_P, _SZ, _SZ, _SZ, _P, _P]
        lib.orc_philox_u01_rows.restype = None
        lib.orc_philox_u01_rows.argtypes = [C.c_uint64, C.c_uint64, _SZ, C.c_uint64, _P]
        lib.orc_philox_raw.restype = None; lib.orc_philox_raw.argtypes = [C.c_uint64] * 3 + [_P]

    # -- models ------------------------------------------------------------
    @staticmethod
    def model(desc):
        d = desc()
        return OrcModel(FAMILY[d.name()], int(d._param() or 0))

    def hp_size(self, m):
        return self.lib.orc_hp_size(C.byref(m))

    def ss_size(self, m):
        return self.lib.orc_ss_size(C.byref(m))

    @staticmethod
    def flat_hp(desc, hp=None):
        """flat hyperparameter vector in the order of include/mscope_b200.h enum msb_family"""
        d = desc()
        hp = dict(d.default_hyperparams(), **(hp or {}))
        n = d.name()
        if n in ("bb", "bbnc"): return np.array([hp["alpha"], hp["beta"]], np.float64)
        if n == "bnb": return np.array([hp["alpha"], hp["beta"], hp["r"]], np.float64)
        if n == "gp": return np.array([hp["alpha"], hp["inv_beta"]], np.float64)
        if n == "nich": return np.array([hp["mu"], hp["kappa"], hp["sigmasq"], hp["nu"]], np.float64)
        if n in ("dd", "dm"): return np.asarray(hp["alphas"], np.float64)
        if n == "niw":
            return np.concatenate([np.asarray(hp["mu"], np.float64).ravel(), [hp["kappa"]],
                                   np.asarray(hp["psi"], np.float64).ravel(), [hp["nu"]]])
        raise ValueError(n)

    def score_value(self, m, hp, ss, x, prec=64):
        hp = np.ascontiguousarray(hp, np.float64); ss = np.ascontiguousarray(ss, np.float64)
        x = np.ascontiguousarray(np.atleast_1d(x), np.float64)
        return self.lib.orc_score_value(C.byref(m), hp.ctypes.data, ss.ctypes.data, x.ctypes.data, prec)

    def add_value(self, m, hp, ss, x, prec=64):
        x = np.ascontiguousarray(np.atleast_1d(x), np.float64)
        self.lib.orc_add_value(C.byref(m), hp.ctypes.data, ss.ctypes.data, x.ctypes.data, prec)

    def remove_value(self, m, hp, ss, x, prec=64):
        x = np.ascontiguousarray(np.atleast_1d(x), np.float64)
        self.lib.orc_remove_value(C.byref(m), hp.ctypes.data, ss.ctypes.data, x.ctypes.data, prec)

    # -- marginal likelihoods ----------------------------------------------------
    def score_data(self, m, hp, ss):
        hp = np.ascontiguousarray(hp, np.float64); ss = np.ascontiguousarray(ss, np.float64)
        self.lib.orc_score_data.restype = _D
        self.lib.orc_score_data.argtypes = [C.POINTER(OrcModel), _P, _P]
        return self.lib.orc_score_data(C.byref(m), hp.ctypes.data, ss.ctypes.data)

    # -- posterior-predictive draws (group::sample_value) ----------------------------
    def sample_value(self, m, hp, ss, seed, counter, n):
        hp = np.ascontiguousarray(hp, np.float64); ss = np.ascontiguousarray(ss, np.float64)
        width = m.dim if m.family == NIW else 1
        out = np.zeros((n, width), np.float64)
        self.lib.orc_sample_value.restype = C.c_int
        self.lib.orc_sample_value.argtypes = [C.POINTER(OrcModel), _P, _P, C.c_uint64, C.c_uint64, _SZ, _P]
        rc = self.lib.orc_sample_value(C.byref(m), hp.ctypes.data, ss.ctypes.data, seed, counter, n, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("multinomial sampling unimplemented" if rc == -1 else "scale matrix is not positive definite")
        return out if width > 1 else out[:, 0]

    def score_assignment(self, assign, alpha, prec=32):
        a = np.ascontiguousarray(assign, np.int64)
        if prec == 32:
            self.lib.orc_score_assignment.restype = C.c_float
            self.lib.orc_score_assignment.argtypes = [_P, _SZ, C.c_float]
            return self.lib.orc_score_assignment(a.ctypes.data, a.size, float(alpha))
        self.lib.orc_score_assignment64.restype = _D
        self.lib.orc_score_assignment64.argtypes = [_P, _SZ, _D]
        return self.lib.orc_score_assignment64(a.ctypes.data, a.size, float(alpha))

    # -- batched ---------------------------------------------------------------
    def _pack(self, descs, view):
        models = (OrcModel * len(descs))(*[self.model(d) for d in descs])
        types = (OrcType * len(descs))(*[OrcType(t.prim, t.n, t.vec) for t in view.types()])
        data, mask = view.raw()
        return models, types, data, mask

    def ss_total(self, descs):
        return sum(self.ss_size(self.model(d)) for d in descs)

    def score_rows(self, descs, hp_flat, ss, logprior, view, row_lo=0, row_hi=None, prec=64, nthreads=4, f32=False):
        models, types, data, mask = self._pack(descs, view)
        row_hi = view.size() if row_hi is None else row_hi
        K = ss.shape[0]
        hp_flat = np.ascontiguousarray(hp_flat, np.float64); ss = np.ascontiguousarray(ss, np.float64)
        lp = np.ascontiguousarray(logprior, np.float64)
        mptr = mask.ctypes.data if mask is not None else None
        if f32:
            out = np.zeros((row_hi - row_lo, K), np.float32)
            self.lib.orc_score_rows_f32(models, len(descs), hp_flat.ctypes.data, ss.ctypes.data, K, lp.ctypes.data,
                                        data.ctypes.data, mptr, types, row_lo, row_hi, nthreads, out.ctypes.data)
        else:
            out = np.zeros((row_hi - row_lo, K), np.float64)
            self.lib.orc_score_rows(models, len(descs), hp_flat.ctypes.data, ss.ctypes.data, K, lp.ctypes.data,
                                    data.ctypes.data, mptr, types, row_lo, row_hi, prec, nthreads, out.ctypes.data)
        return out

    def update_rows(self, descs, hp_flat, ss, counts, view, old, new, row_lo=0, row_hi=None, prec=64):
        models, types, data, mask = self._pack(descs, view)
        row_hi = view.size() if row_hi is None else row_hi
        K = ss.shape[0]
        hp_flat = np.ascontiguousarray(hp_flat, np.float64)
        assert ss.flags.c_contiguous and ss.dtype == np.float64 and counts.dtype == np.float64
        old_p = np.ascontiguousarray(old, np.int32) if old is not None else None
        new_p = np.ascontiguousarray(new, np.int32) if new is not None else None
        self.lib.orc_update_rows(models, len(descs), hp_flat.ctypes.data, ss.ctypes.data, K, counts.ctypes.data,
                                 data.ctypes.data, mask.ctypes.data if mask is not None else None, types, row_lo, row_hi,
                                 old_p.ctypes.data if old_p is not None else None,
                                 new_p.ctypes.data if new_p is not None else None, prec)

    # -- sampler -----------------------------------------------------------------
    def expf(self, x):
        return self.lib.orc_expf(float(x))

    def sample_rows(self, scores, u):
        s = np.ascontiguousarray(scores, np.float32); u = np.ascontiguousarray(u, np.float32)
        out = np.zeros(s.shape[0], np.int32)
        self.lib.orc_sample_rows(s.ctypes.data, s.shape[0], s.shape[1], s.shape[1], u.ctypes.data, out.ctypes.data)
        return out

    def sample_rows_libm(self, scores, u):
        """the same walk with glibc's expf (util.hpp:131): how often does the exponential change a draw?"""
        s = np.ascontiguousarray(scores, np.float32); u = np.ascontiguousarray(u, np.float32)
        out = np.zeros(s.shape[0], np.int32)
        self.lib.orc_sample_rows_libm(s.ctypes.data, s.shape[0], s.shape[1], s.shape[1], u.ctypes.data, out.ctypes.data)
        return out

    def philox_u01_rows(self, seed, row_lo, n, sweep):
        out = np.zeros(n, np.float32)
        self.lib.orc_philox_u01_rows(seed, row_lo, n, sweep, out.ctypes.data)
        return out

    def philox_u01(self, seed, row, sweep):
        return self.lib.orc_philox_u01(seed, row, sweep)

    def philox_raw(self, seed, row, sweep):
        out = (C.c_uint32 * 4)()
        self.lib.orc_philox_raw(seed, row, sweep, out)
        return [int(v) for v in out]


_oracle = None


def load():
    global _oracle
    if _oracle is None:
        src = os.path.join(ORACLE_DIR, "msb_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libmsb_oracle.so"])
        _oracle = Oracle(C.CDLL(LIB))
    return _oracle


def logprior(counts, alpha):
    """log(pseudocount), group_manager.hpp:274-283 (float arithmetic like the reference)"""
    counts = np.asarray(counts, np.float64)
    nempty = int((counts == 0).sum())
    pseudo = np.where(counts != 0, counts.astype(np.float32),
                      np.float32(alpha) / np.float32(max(nempty, 1))).astype(np.float32)
    return np.log(pseudo.astype(np.float64)).astype(np.float32).astype(np.float64)


def build_suffstats(orc, descs, hp_flat, view, z, K, prec=64):
    """suffstats of K groups after add_value of every row to group z[row]"""
    ss = np.zeros((K, orc.ss_total(descs)), np.float64)
    counts = np.zeros(K, np.float64)
    orc.update_rows(descs, hp_flat, ss, counts, view, None, z, prec=prec)
    return ss, counts
