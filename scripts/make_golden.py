#!/usr/bin/env python
"""Generates tests/golden/*.json: known-answer vectors the oracle is pinned against.

Run in the BUILD container only (it reads /root/reference, which does not exist
on the GPU box); the JSON it writes is committed.

Sources of truth, in order of authority:
  1. the reference's own in-tree closed forms, imported from
     /root/reference/microscopes/common/vendor/stats.py:
       multivariate_t_loglik (:235-243)  -> niw predictive
       beta_predictive       (:245-255)  -> bb predictive
     (the file is Python-2 era; `numpy.core.umath_tests.inner1d` is shimmed
     with an einsum, nothing else is touched)
  2. scipy.stats / scipy.special in fp64 for every family
  3. hand-computable cases
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import scipy.special as sp
import scipy.stats as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_STATS = "/root/reference/microscopes/common/vendor/stats.py"


def load_reference_stats():
    shim = types.ModuleType("numpy.core.umath_tests")
    shim.inner1d = lambda a, b: np.einsum("...i,...i->...", a, b)
    sys.modules["numpy.core.umath_tests"] = shim
    spec = importlib.util.spec_from_file_location("ref_vendor_stats", REF_STATS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def niw_posterior(mu0, kappa0, psi0, nu0, n, sx, sxx):
    d = len(mu0)
    kn = kappa0 + n
    nun = nu0 + n
    mun = (kappa0 * mu0 + sx) / kn
    psin = psi0 + sxx + kappa0 * np.outer(mu0, mu0) - kn * np.outer(mun, mun)
    dof = nun - d + 1.0
    scale = psin * (kn + 1.0) / (kn * dof)
    return dof, mun, scale


def main():
    ref = load_reference_stats()
    rng = np.random.default_rng(20141118)
    cases = []

    # ---- bb ---------------------------------------------------------------
    for alpha, beta, heads, tails in [(1, 1, 3, 1), (2, 2, 0, 0), (0.5, 1.5, 10, 7), (1, 1, 0, 5), (3.25, 0.75, 1000, 1)]:
        for x in (0, 1):
            refv = float(ref.beta_predictive((alpha + heads, beta + tails), (1, 0) if x else (0, 1)))
            scv = float(sp.betaln(alpha + heads + x, beta + tails + 1 - x) - sp.betaln(alpha + heads, beta + tails))
            cases.append(dict(family="bb", dim=0, hp=[alpha, beta], ss=[heads, tails], x=[x],
                              expect=scv, ref_vendor=refv, source="vendor/stats.py:beta_predictive + scipy betaln"))
    # hand-computable: alpha=beta=1, heads=3, tails=1 => log(4/6), log(2/6)
    cases.append(dict(family="bb", dim=0, hp=[1, 1], ss=[3, 1], x=[1], expect=float(np.log(4 / 6)), source="hand"))
    cases.append(dict(family="bb", dim=0, hp=[1, 1], ss=[3, 1], x=[0], expect=float(np.log(2 / 6)), source="hand"))

    # ---- dd ---------------------------------------------------------------
    for C in (2, 5, 128, 256):
        alphas = rng.uniform(0.2, 2.0, size=C)
        counts = rng.integers(0, 50, size=C).astype(float)
        for x in sorted(set([0, C // 2, C - 1])):
            # Dirichlet-multinomial predictive via betaln-free gamma ratio
            a = alphas + counts
            # (the gamma-ratio form loses ~1e-12 to cancellation at sum(a) ~ 6000, so it is only the cross-check)
            alt = float(sp.gammaln(a[x] + 1) - sp.gammaln(a[x]) + sp.gammaln(a.sum()) - sp.gammaln(a.sum() + 1))
            exp = float(np.log(a[x]) - np.log(a.sum()))
            assert abs(alt - exp) < 1e-10
            cases.append(dict(family="dd", dim=C, hp=alphas.tolist(), ss=[counts.sum()] + counts.tolist(), x=[x],
                              expect=exp, source="log ratio, cross-checked with scipy gammaln to 1e-10"))
    cases.append(dict(family="dd", dim=4, hp=[1, 1, 1, 1], ss=[0, 0, 0, 0, 0], x=[2], expect=float(np.log(0.25)),
                      source="hand: empty group, uniform alpha"))

    # ---- gp: predictive = NegBin(r = alpha+sum, p = b/(1+b)), b = inv_beta+count
    for alpha, inv_beta, count, total in [(1, 1, 0, 0), (2.0, 0.5, 10, 83), (0.7, 3.0, 1000, 7512), (1, 1, 5, 0)]:
        for x in (0, 1, 7, 40, 300):
            a, b = alpha + total, inv_beta + count
            exp = float(st.nbinom.logpmf(x, a, b / (1.0 + b)))
            log_prod = 0.0  # not used by the predictive
            cases.append(dict(family="gp", dim=0, hp=[alpha, inv_beta], ss=[count, total, log_prod], x=[x],
                              expect=exp, source="scipy nbinom.logpmf"))

    # ---- nich: Student-t(nu', mu', sigmasq' (kappa'+1)/kappa') ---------------
    for mu, kappa, sigmasq, nu, n in [(0, 1, 1, 1, 0), (0.5, 2.0, 1.5, 3.0, 7), (-3.0, 0.1, 0.3, 2.0, 250), (0, 1, 1, 1, 8000)]:
        data = rng.normal(1.3, 0.8, size=n)
        mean = float(data.mean()) if n else 0.0
        ctv = float(((data - mean) ** 2).sum()) if n else 0.0
        kn, nun = kappa + n, nu + n
        mun = (kappa * mu + n * mean) / kn
        sn = (nu * sigmasq + ctv + n * kappa * (mu - mean) ** 2 / kn) / nun
        for x in (-2.0, 0.0, 1.3, 1.31, 9.5):
            exp = float(st.t.logpdf(x, df=nun, loc=mun, scale=np.sqrt(sn * (kn + 1) / kn)))
            cases.append(dict(family="nich", dim=0, hp=[mu, kappa, sigmasq, nu], ss=[n, mean, ctv], x=[x],
                              expect=exp, source="scipy t.logpdf"))

    # ---- niw: reference multivariate_t_loglik + scipy multivariate_t ----------
    for d, n in [(2, 0), (3, 5), (8, 40), (64, 0), (64, 300)]:
        mu0 = rng.normal(0, 1, size=d)
        kappa0 = float(rng.uniform(0.5, 2.0))
        A = rng.normal(size=(d, d))
        psi0 = A @ A.T / d + np.eye(d)
        nu0 = d + float(rng.uniform(0, 3))
        X = rng.normal(0.5, 1.2, size=(n, d))
        sx = X.sum(0) if n else np.zeros(d)
        sxx = X.T @ X if n else np.zeros((d, d))
        dof, mun, scale = niw_posterior(mu0, kappa0, psi0, nu0, n, sx, sxx)
        for _ in range(3):
            x = rng.normal(0.5, 1.5, size=d)
            refv = float(np.ravel(ref.multivariate_t_loglik(x, dof, mun, scale))[0])
            scv = float(st.multivariate_t.logpdf(x, loc=mun, shape=scale, df=dof))
            cases.append(dict(family="niw", dim=d,
                              hp=np.concatenate([mu0, [kappa0], psi0.ravel(), [nu0]]).tolist(),
                              ss=np.concatenate([[n], sx, sxx.ravel()]).tolist(), x=x.tolist(),
                              expect=scv, ref_vendor=refv,
                              source="vendor/stats.py:multivariate_t_loglik + scipy multivariate_t"))

    # ---- bnb: beta-negative-binomial predictive (appended last so that the cases above keep their random draws)
    for alpha, beta, r, count, total in [(1, 1, 1, 0, 0), (2.0, 3.0, 2, 10, 37), (0.7, 1.3, 5, 400, 2512), (1, 1, 1, 7, 0)]:
        a, b = alpha + r * count, beta + total
        for x in (0, 1, 6, 40, 250):
            exp = float(st.betanbinom.logpmf(x, r, a, b))
            alt = float(sp.gammaln(r + x) - sp.gammaln(r) - sp.gammaln(x + 1) + sp.betaln(a + r, b + x) - sp.betaln(a, b))
            assert abs(exp - alt) < 1e-9 * max(1.0, abs(exp))
            cases.append(dict(family="bnb", dim=0, hp=[alpha, beta, r], ss=[count, total], x=[x],
                              expect=alt, source="scipy betanbinom.logpmf == gammaln/betaln closed form"))

    # ---- dm: Dirichlet-multinomial predictive of a count vector (src/models/dm.cpp:38-76), own generator so that
    # the cases above keep their random draws
    rng_dm = np.random.default_rng(38_76)
    for C, scale, total in [(3, 1.0, 0), (3, 1.0, 5), (8, 0.5, 40), (16, 2.0, 300), (64, 1.0, 5000), (5, 0.1, 12)]:
        alphas = rng_dm.uniform(0.2, 2.0, size=C) * scale
        counts = rng_dm.multinomial(total, rng_dm.dirichlet(np.ones(C))) if total else np.zeros(C, int)
        for xt in (0, 1, 7, 60):
            x = rng_dm.multinomial(xt, rng_dm.dirichlet(np.ones(C)))
            e = alphas + counts
            exp = float(st.dirichlet_multinomial.logpmf(x, e, xt))
            alt = float(sp.gammaln(xt + 1) - sp.gammaln(x + 1).sum() + sp.gammaln(e.sum()) - sp.gammaln(e.sum() + xt)
                        + (sp.gammaln(e + x) - sp.gammaln(e)).sum())
            assert abs(exp - alt) < 1e-9 * max(1.0, abs(exp))
            cases.append(dict(family="dm", dim=C, hp=alphas.tolist(), ss=counts.astype(float).tolist() + [0.0], x=x.tolist(),
                              expect=alt, source="scipy dirichlet_multinomial.logpmf == dm.cpp:38-76 closed form"))

    out = os.path.join(ROOT, "tests", "golden", "score_value.json")
    with open(out, "w") as f:
        json.dump(dict(generator="scripts/make_golden.py", cases=cases), f)
    print("wrote %d cases to %s" % (len(cases), out))
    worst = max(abs(c["expect"] - c["ref_vendor"]) / max(1.0, abs(c["expect"])) for c in cases if "ref_vendor" in c)
    print("max |scipy - reference vendor/stats.py| (relative): %.3e" % worst)

    # ---- sampler known answers: exhaustive small cases by exact arithmetic ------
    samp = []
    for probs in ([0.5, 0.5], [0.1, 0.2, 0.7], [1.0], [0.25] * 4, [1e-6, 1 - 1e-6]):
        scores = np.log(np.asarray(probs, np.float64)).astype(np.float32)
        cdf = np.cumsum(probs)
        for u in (0.0, 0.05, 0.26, 0.49, 0.51, 0.74, 0.999):
            # skip uniforms within 1e-4 of a cdf boundary (float rounding could go either way)
            if np.min(np.abs(cdf - u)) < 1e-4 and u > 0:
                continue
            k = int(np.searchsorted(cdf, u, side="left")) if u > 0 else 0
            samp.append(dict(scores=scores.tolist(), u=u, expect=min(k, len(probs) - 1)))
    with open(os.path.join(ROOT, "tests", "golden", "sample_discrete_log.json"), "w") as f:
        json.dump(dict(generator="scripts/make_golden.py", cases=samp), f)
    print("wrote %d sampler cases" % len(samp))


if __name__ == "__main__":
    main()
