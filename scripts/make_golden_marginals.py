#!/usr/bin/env python
"""Generates tests/golden/score_data.json: known answers for group::score_data (models/base.hpp:28), the log
marginal likelihood of a group's data, and for group_manager::score_assignment (group_manager.hpp:250-272).

Run in the BUILD container only (reads /root/reference); the JSON is committed.  Sources of truth:
  1. niw: the reference's own in-tree inverse-Wishart partition function
     (microscopes/common/vendor/stats.py:227-231): p(X) = (2 pi)^(-n d / 2) Z(psi', nu') / Z(psi, nu) (kappa/kappa')^(d/2)
  2. every family: scipy.special (betaln, gammaln, multigammaln) + numpy slogdet in fp64, written independently of
     the oracle's C
  3. the chain rule: log p(x_1..x_n) = sum_i log p(x_i | x_<i) with scipy's predictive densities
  4. score_assignment: the Ewens / CRP formula with scipy gammaln, and a direct float64 loop
"""
import json
import os
import sys

import numpy as np
import scipy.special as sp
import scipy.stats as st

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import load_reference_stats, ROOT  # noqa: E402


def main():
    ref = load_reference_stats()
    rng = np.random.default_rng(20141119)
    cases = []

    # ---- bb: betaln ratio; chain rule with Bernoulli predictive --------------------------------------
    for alpha, beta, n in [(1, 1, 0), (1, 1, 4), (0.5, 1.5, 17), (3.25, 0.75, 1001)]:
        x = rng.integers(0, 2, size=n)
        heads, tails = int(x.sum()), int(n - x.sum())
        exp = float(sp.betaln(alpha + heads, beta + tails) - sp.betaln(alpha, beta))
        chain, h, t = 0.0, 0, 0
        for xi in x:
            chain += np.log(((alpha + h) if xi else (beta + t)) / (alpha + beta + h + t))
            h += int(xi); t += int(1 - xi)
        assert abs(chain - exp) < 1e-9 * max(1.0, abs(exp))
        cases.append(dict(family="bb", dim=0, hp=[alpha, beta], ss=[heads, tails], expect=exp, source="scipy betaln; chain rule"))

    # ---- dd: Dirichlet-multinomial ------------------------------------------------------------------
    for C, n in [(2, 0), (5, 9), (16, 300), (256, 5000)]:
        alphas = rng.uniform(0.2, 2.0, size=C)
        counts = np.bincount(rng.integers(0, C, size=n), minlength=C).astype(float)
        exp = float(sp.gammaln(alphas + counts).sum() - sp.gammaln(alphas).sum() + sp.gammaln(alphas.sum()) - sp.gammaln(alphas.sum() + n))
        cases.append(dict(family="dd", dim=C, hp=alphas.tolist(), ss=[float(n)] + counts.tolist(), expect=exp, source="scipy gammaln"))

    # ---- gp: chain rule with the negative-binomial predictive ------------------------------------------
    for alpha, inv_beta, n in [(1, 1, 0), (2.0, 0.5, 10), (0.7, 3.0, 400)]:
        x = rng.poisson(6.0, size=n)
        chain, cnt, tot = 0.0, 0, 0
        for xi in x:
            a, b = alpha + tot, inv_beta + cnt
            chain += float(st.nbinom.logpmf(xi, a, b / (1.0 + b)))
            cnt += 1; tot += int(xi)
        log_prod = float(sp.gammaln(x + 1.0).sum())
        exp = float(sp.gammaln(alpha + tot) - sp.gammaln(alpha) + alpha * np.log(inv_beta) - (alpha + tot) * np.log(inv_beta + cnt) - log_prod)
        assert abs(chain - exp) < 1e-9 * max(1.0, abs(exp)), (chain, exp)
        cases.append(dict(family="gp", dim=0, hp=[alpha, inv_beta], ss=[float(cnt), float(tot), log_prod], expect=exp,
                          source="closed form == chain rule over scipy nbinom.logpmf"))

    # ---- nich: chain rule with the Student-t predictive -------------------------------------------------
    for mu, kappa, sigmasq, nu, n in [(0, 1, 1, 1, 0), (0.5, 2.0, 1.5, 3.0, 7), (-3.0, 0.1, 0.3, 2.0, 250)]:
        x = rng.normal(1.3, 0.8, size=n)
        chain = 0.0
        for i in range(n):
            xs = x[:i]
            m = float(xs.mean()) if i else 0.0
            ctv = float(((xs - m) ** 2).sum()) if i else 0.0
            kn, nun = kappa + i, nu + i
            mun = (kappa * mu + i * m) / kn
            sn = (nu * sigmasq + ctv + i * kappa * (mu - m) ** 2 / kn) / nun
            chain += float(st.t.logpdf(x[i], df=nun, loc=mun, scale=np.sqrt(sn * (kn + 1) / kn)))
        mean = float(x.mean()) if n else 0.0
        ctv = float(((x - mean) ** 2).sum()) if n else 0.0
        kn, nun = kappa + n, nu + n
        sn = (nu * sigmasq + ctv + n * kappa * (mu - mean) ** 2 / kn) / nun
        exp = float(sp.gammaln(nun / 2) - sp.gammaln(nu / 2) + 0.5 * np.log(kappa / kn) + 0.5 * nu * np.log(nu * sigmasq)
                    - 0.5 * nun * np.log(nun * sn) - 0.5 * n * np.log(np.pi))
        assert abs(chain - exp) < 1e-9 * max(1.0, abs(exp)), (chain, exp)
        cases.append(dict(family="nich", dim=0, hp=[mu, kappa, sigmasq, nu], ss=[float(n), mean, ctv], expect=exp,
                          source="closed form == chain rule over scipy t.logpdf"))

    # ---- niw: the reference's inverse-Wishart partition function + scipy multigammaln / slogdet ----------
    for d, n in [(2, 0), (3, 5), (8, 40), (64, 300)]:
        mu0 = rng.normal(0, 1, size=d)
        kappa0 = float(rng.uniform(0.5, 2.0))
        A = rng.normal(size=(d, d))
        psi0 = A @ A.T / d + np.eye(d)
        nu0 = d + float(rng.uniform(0, 3))
        X = rng.normal(0.5, 1.2, size=(n, d))
        sx = X.sum(0) if n else np.zeros(d)
        sxx = X.T @ X if n else np.zeros((d, d))
        kn, nun = kappa0 + n, nu0 + n
        mun = (kappa0 * mu0 + sx) / kn
        psin = psi0 + sxx + kappa0 * np.outer(mu0, mu0) - kn * np.outer(mun, mun)
        refv = float(-0.5 * n * d * np.log(2 * np.pi) + ref.invwishart_log_partitionfunction(psin, nun)
                     - ref.invwishart_log_partitionfunction(psi0, nu0) + 0.5 * d * np.log(kappa0 / kn))
        exp = float(-0.5 * n * d * np.log(np.pi) + sp.multigammaln(nun / 2, d) - sp.multigammaln(nu0 / 2, d)
                    + 0.5 * nu0 * np.linalg.slogdet(psi0)[1] - 0.5 * nun * np.linalg.slogdet(psin)[1] + 0.5 * d * np.log(kappa0 / kn))
        assert abs(refv - exp) < 1e-9 * max(1.0, abs(exp)), (refv, exp)
        cases.append(dict(family="niw", dim=d, hp=np.concatenate([mu0, [kappa0], psi0.ravel(), [nu0]]).tolist(),
                          ss=np.concatenate([[n], sx, sxx.ravel()]).tolist(), expect=exp, ref_vendor=refv,
                          source="vendor/stats.py:invwishart_log_partitionfunction + scipy multigammaln/slogdet"))

    # ---- score_assignment -----------------------------------------------------------------------------
    crp = []
    for n, k, alpha in [(1, 1, 1.0), (6, 2, 1.0), (50, 7, 0.3), (2000, 40, 2.5)]:
        z = rng.integers(0, k, size=n)
        # direct loop of group_manager.hpp:250-272 in float64
        counts, s = {int(z[0]): 1}, 0.0
        for i in range(1, n):
            g = int(z[i])
            s += np.log((counts[g] if g in counts else alpha) / (i + alpha))
            counts[g] = counts.get(g, 0) + 1
        # Ewens: alpha^(K-1) prod Gamma(n_g) / prod_{i=1}^{n-1} (i + alpha)   (entity 0 contributes nothing)
        sizes = np.array(list(counts.values()), float)
        ew = float((len(sizes) - 1) * np.log(alpha) + sp.gammaln(sizes).sum() - (sp.gammaln(n + alpha) - sp.gammaln(1 + alpha)))
        assert abs(ew - s) < 1e-9 * max(1.0, abs(s))
        crp.append(dict(assign=z.tolist(), alpha=alpha, expect=float(s)))

    out = os.path.join(ROOT, "tests", "golden", "score_data.json")
    with open(out, "w") as f:
        json.dump(dict(generator="scripts/make_golden_marginals.py", cases=cases, crp=crp), f)
    print("wrote %d score_data cases and %d score_assignment cases to %s" % (len(cases), len(crp), out))


if __name__ == "__main__":
    main()
