"""Synthetic planted-mixture datasets of the BASELINE.json shapes.

Used by bench.py and the tests.  Counter-based generator (numpy Philox), seed 73
as in the reference's own micro-benchmark (bin/perf_group.cpp:19).
"""
import numpy as np
import numpy.ma as ma

from . import models as M


def _rng(seed, stream):
    return np.random.Generator(np.random.Philox(key=[int(seed), int(stream)]))


def _categorical(rng, theta_kc, z):
    """x[n] ~ Cat(theta[z[n]]) without materialising an N x C matrix"""
    K, Cn = theta_kc.shape
    cdf = np.cumsum(theta_kc, axis=1)
    cdf /= cdf[:, -1:]
    flat = (cdf + np.arange(K)[:, None]).ravel()
    u = rng.random(z.shape[0])
    idx = np.searchsorted(flat, u + z, side="right")
    x = idx - z * Cn
    return np.clip(x, 0, Cn - 1)


def make_dataset(models, n, k_true, seed=73, stream=0, mask_frac=0.0, storage=None):
    """returns (structured array [masked if mask_frac>0], planted assignment z)

    storage: optional list of numpy dtypes overriding each model's Value dtype
    (any of the 11 primitive types may back a field, runtime_type.hpp:145-166).
    """
    # the planted mixture (group parameters) depends on the seed only; `stream` selects which rows
    # of that mixture are drawn, so that every rank of a row-sharded run sees the same model
    prng = _rng(seed, 1_000_003)
    rng = _rng(seed, stream)
    z = (np.arange(n) % k_true).astype(np.int64)
    cols, fields = [], []
    for d, m in enumerate(models):
        m = m()
        name = m.name()
        if name in ("bb", "bbnc"):
            p = prng.beta(0.5, 0.5, size=k_true)
            x = rng.random(n) < p[z]
            dt = np.bool_
        elif name == "dd":
            C = m._param()
            theta = prng.dirichlet(np.full(C, 0.5), size=k_true)
            x = _categorical(rng, theta, z)
            dt = np.int32
        elif name == "bnb":
            pk = prng.beta(2.0, 2.0, size=k_true) * 0.8 + 0.1
            x = rng.negative_binomial(1, pk[z])
            dt = np.uint32
        elif name == "gp":
            lam = prng.gamma(2.0, 4.0, size=k_true)
            x = rng.poisson(lam[z])
            dt = np.uint32
        elif name == "nich":
            mu = prng.normal(0.0, 3.0, size=k_true)
            sg = prng.uniform(0.5, 2.0, size=k_true)
            x = mu[z] + sg[z] * rng.standard_normal(n)
            dt = np.float32
        elif name == "dm":
            C = m._param()
            theta = prng.dirichlet(np.full(C, 0.5), size=k_true)
            tot = rng.poisson(20.0, size=n)
            x = np.stack([rng.multinomial(int(tot[i]), theta[z[i]]) for i in range(n)]).astype(np.uint32)
            dt = np.dtype((np.int32, (C,)))
        elif name == "niw":
            dim = m._param()
            mu = prng.normal(0.0, 2.0, size=(k_true, dim))
            A = prng.standard_normal((k_true, dim, dim))
            L = np.linalg.cholesky(A @ A.transpose(0, 2, 1) / dim + 0.1 * np.eye(dim))
            e = rng.standard_normal((n, dim))
            x = np.empty((n, dim))
            for g in range(k_true):  # z = row mod K: the rows of group g are g, g + K, ...  (no N x d x d gather of L[z])
                x[g::k_true] = mu[g] + e[g::k_true] @ L[g].T
            dt = np.dtype((np.float32, (dim,)))
        else:
            raise ValueError(name)
        if storage is not None and storage[d] is not None:
            dt = storage[d]
        fields.append(("f%d" % d, dt))
        cols.append(x)
    arr = np.zeros(n, dtype=np.dtype(fields))
    for d, x in enumerate(cols):
        arr["f%d" % d] = x
    if mask_frac > 0.0:
        mask = np.zeros(n, dtype=[(nm, np.bool_, arr.dtype.fields[nm][0].shape) for nm in arr.dtype.names])
        for nm in arr.dtype.names:
            cell = rng.random(n) < mask_frac
            if mask[nm].ndim == 2:
                mask[nm] = cell[:, None]
            else:
                mask[nm] = cell
        arr = ma.array(arr, mask=mask)
    return arr, z


# the five BASELINE.json configurations (SURVEY.md section 8d "Configs -> shapes")
def config(name):
    name = name.upper()
    if name == "C1":
        return dict(models=[M.bb] * 64, n=10_000, k=50, hp={"alpha": 2.0, "beta": 2.0})
    if name == "C2":
        return dict(models=[M.dd(256)] * 32, n=1_000_000, k=200, storage=[np.uint8] * 32)
    if name == "C3":
        return dict(models=[M.nich] * 128, n=4_000_000, k=500)
    if name == "C4":
        return dict(models=[M.niw(64)], n=1_000_000, k=256)
    if name == "C5":
        return dict(models=[M.bb] * 64 + [M.gp] * 64 + [M.nich] * 64 + [M.dd(16)] * 64, n=1_000_000, k=1000)
    raise ValueError(name)
