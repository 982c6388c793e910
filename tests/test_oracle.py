"""CPU: the oracle against the golden vectors (reference vendor/stats.py, scipy, hand cases)
and the properties SURVEY.md section 8c lists."""
import json
import os

import numpy as np
import pytest

import common_b200 as cb
import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FAM = {"dm": ol.DM, "bb": ol.BB, "bbnc": ol.BBNC, "bnb": ol.BNB, "dd": ol.DD, "gp": ol.GP, "nich": ol.NICH, "niw": ol.NIW}


def _cases():
    with open(os.path.join(GOLD, "score_value.json")) as f:
        return json.load(f)["cases"]


def test_golden_score_value_fp64(oracle):
    n = 0
    for c in _cases():
        m = ol.OrcModel(FAM[c["family"]], c["dim"])
        got = oracle.score_value(m, c["hp"], c["ss"], c["x"], 64)
        # fp64 conditioning: lgamma(a+x) - lgamma(a) at a ~ 7.5e3 cancels ~1e4 (gp); d=64 Cholesky (niw)
        cond = {"niw": 50.0, "gp": max(1.0, 1e-3 * (c["hp"][0] + c["ss"][1])),
                "bnb": max(1.0, 1e-3 * (sum(c["hp"][:2]) + c["hp"][-1] * c["ss"][0] + c["ss"][1])) if c["family"] == "bnb" else 1.0,
                "dm": max(4.0, 1e-2 * sum(c["ss"])) if c["family"] == "dm" else 1.0,  # lgamma(E + X) - lgamma(E), E = sum(alpha + counts)
                }.get(c["family"], 1.0)
        tol = 1e-12 * max(1.0, abs(c["expect"])) * cond
        assert abs(got - c["expect"]) <= tol, (c["family"], c["source"], got, c["expect"])
        if "ref_vendor" in c:  # the reference's own in-tree closed form
            assert abs(got - c["ref_vendor"]) <= 5e-11 * max(1.0, abs(c["ref_vendor"]))
        n += 1
    assert n >= 94


def test_golden_score_value_fp32_restatement(oracle):
    # the float restatement (reference arithmetic order) is allowed fp32 conditioning error only
    for c in _cases():
        if c["family"] == "niw":
            continue
        m = ol.OrcModel(FAM[c["family"]], c["dim"])
        got = oracle.score_value(m, c["hp"], c["ss"], c["x"], 32)
        # float conditioning of the upstream formulas: nich log(1+z) error grows with nu' = nu + n;
        # gp lgammaf(a+x) - lgammaf(a) cancels terms of size a log a (a = alpha + sum)
        n = c["ss"][0] if c["family"] == "nich" else 1.0
        if c["family"] == "gp":
            a = c["hp"][0] + c["ss"][1] + c["x"][0]
            n = a * np.log(a + 2.0)
        if c["family"] == "bnb":
            a = c["hp"][0] + c["hp"][1] + c["hp"][2] * (c["ss"][0] + 1) + c["ss"][1] + c["x"][0]
            n = 3 * a * np.log(a + 2.0)
        if c["family"] == "dm":  # dim + 1 differences lgammaf(e + x) - lgammaf(e), each cancelling terms of size e log e
            e = np.asarray(c["hp"]) + np.asarray(c["ss"][:-1])
            n = 3 * (np.sum((e + c["x"]) * np.log(e + np.asarray(c["x"]) + 2.0)) + (e.sum() + sum(c["x"])) * np.log(e.sum() + sum(c["x"]) + 2.0))
        tol = 2e-5 * max(1.0, abs(c["expect"])) + 6e-7 * n
        assert abs(got - c["expect"]) <= tol, (c["family"], got, c["expect"])


@pytest.mark.parametrize("desc", [cb.bb, cb.bnb, cb.gp, cb.nich, cb.dd(7), cb.niw(3)])
def test_add_remove_roundtrip_and_direct_suffstats(oracle, desc):
    rng = np.random.default_rng(5)
    m = oracle.model(desc)
    hp = oracle.flat_hp(desc)
    ss = np.zeros(oracle.ss_size(m))
    name = desc().name()

    def draw():
        if name == "bb": return [float(rng.integers(0, 2))]
        if name == "dd": return [float(rng.integers(0, 7))]
        if name == "gp": return [float(rng.poisson(6))]
        if name == "bnb": return [float(rng.negative_binomial(1, 0.3))]
        if name == "nich": return [float(rng.normal(2, 1.5))]
        return rng.normal(0, 1, size=3).tolist()

    xs = [draw() for _ in range(40)]
    for x in xs:
        oracle.add_value(m, hp, ss, x)
    # (iv) score after add == score from directly-set suffstats
    X = np.asarray(xs)
    if name == "bb":
        direct = [X.sum(), len(xs) - X.sum()]
    elif name == "dd":
        direct = [len(xs)] + np.bincount(X[:, 0].astype(int), minlength=7).tolist()
    elif name == "bnb":
        direct = [len(xs), X.sum()]
    elif name == "gp":
        from scipy.special import gammaln
        direct = [len(xs), X.sum(), gammaln(X[:, 0] + 1).sum()]
    elif name == "nich":
        direct = [len(xs), X.mean(), ((X - X.mean()) ** 2).sum()]
    else:
        direct = np.concatenate([[len(xs)], X.sum(0), (X.T @ X).ravel()]).tolist()
    np.testing.assert_allclose(ss, direct, rtol=1e-10, atol=1e-10)
    probe = draw()
    assert abs(oracle.score_value(m, hp, ss, probe) - oracle.score_value(m, hp, np.asarray(direct, float), probe)) < 1e-9
    # (iii) add then remove returns to the initial state (ints exactly, floats within tol)
    for x in reversed(xs):
        oracle.remove_value(m, hp, ss, x)
    if name in ("bb", "dd", "bnb"):
        assert np.all(ss == 0)
    else:
        assert ss[0] == 0
        np.testing.assert_allclose(ss, 0, atol=1e-9)


def test_normalisation(oracle):
    # (v) sum_x exp(score) = 1 for bb / dd, ~1 for gp over a long range
    rng = np.random.default_rng(9)
    m = ol.OrcModel(ol.BB, 0)
    assert abs(sum(np.exp(oracle.score_value(m, [0.3, 2.0], [4, 9], [x])) for x in (0, 1)) - 1) < 1e-12
    m = ol.OrcModel(ol.DD, 256)
    al = rng.uniform(0.1, 2, 256); cn = rng.integers(0, 9, 256).astype(float)
    tot = sum(np.exp(oracle.score_value(m, al, np.concatenate([[cn.sum()], cn]), [x])) for x in range(256))
    assert abs(tot - 1) < 1e-12
    m = ol.OrcModel(ol.GP, 0)
    tot = sum(np.exp(oracle.score_value(m, [2.0, 0.5], [10, 83, 0.0], [x])) for x in range(400))
    assert abs(tot - 1) < 1e-9


def test_empty_group_is_prior_predictive(oracle):
    from scipy import stats
    m = ol.OrcModel(ol.NICH, 0)
    hp = [0.7, 2.0, 1.5, 3.0]
    for x in (-1.0, 0.7, 4.0):
        exp = stats.t.logpdf(x, df=3.0, loc=0.7, scale=np.sqrt(1.5 * 3.0 / 2.0))
        assert abs(oracle.score_value(m, hp, [0, 0, 0], [x]) - exp) < 1e-12


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    def run(ctr, key):
        return oracle.philox_raw(key[0] | key[1] << 32, ctr[0] | ctr[1] << 32, ctr[2] | ctr[3] << 32)
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = [oracle.philox_u01(73, i, 0) for i in range(2000)]
    assert 0.0 <= min(u) and max(u) < 1.0 and abs(np.mean(u) - 0.5) < 0.03


def test_expf_is_a_faithful_exp(oracle):
    xs = np.linspace(-87, 0, 20001).astype(np.float32)
    got = np.array([oracle.expf(x) for x in xs], np.float64)
    ref = np.exp(xs.astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 2.0 * 2 ** -24
    assert oracle.expf(0.0) == 1.0 and oracle.expf(-200.0) == 0.0 and oracle.expf(float("-inf")) == 0.0


def test_sampler_golden_and_semantics(oracle):
    with open(os.path.join(GOLD, "sample_discrete_log.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        s = np.asarray([c["scores"]], np.float32)
        assert oracle.sample_rows(s, [c["u"]])[0] == c["expect"], c
    # util.hpp:149-155: last index is the fallback when the dart never reaches 0
    s = np.log(np.asarray([[0.2, 0.3, 0.5]], np.float32))
    assert oracle.sample_rows(s, [np.float32(1.0)])[0] == 2
    # -inf scores are never drawn; shifting all scores changes nothing (max-subtract, util.hpp:128-130)
    s = np.asarray([[-np.inf, 0.0, -np.inf, 0.0]], np.float32)
    assert set(oracle.sample_rows(np.repeat(s, 50, 0), np.linspace(0.01, 0.99, 50))) == {1, 3}
    # ... except by a dart of exactly 0, which stops at index 0 whatever its probability (util.hpp:151-153)
    assert oracle.sample_rows(s, [0.0])[0] == 0
    rng = np.random.default_rng(3)
    sc = rng.normal(0, 3, size=(200, 17)).astype(np.float32)
    u = rng.random(200).astype(np.float32)
    a = oracle.sample_rows(sc, u)
    # distribution check against exact probabilities (testutil.py-style, chi-square-ish bound)
    big = np.repeat(sc[:1], 20000, 0)
    draws = oracle.sample_rows(big, rng.random(20000).astype(np.float32))
    p = np.exp(sc[0].astype(np.float64)); p /= p.sum()
    freq = np.bincount(draws, minlength=17) / 20000.0
    assert np.max(np.abs(freq - p)) < 0.02
    assert a.min() >= 0 and a.max() < 17


def test_batched_scores_match_single_calls(oracle):
    descs = [cb.bb, cb.gp, cb.nich, cb.dd(5)]
    arr, z = cb.synth.make_dataset(descs, 60, 4, seed=1, mask_frac=0.2)
    view = cb.numpy_dataview(arr)
    hp = np.concatenate([oracle.flat_hp(d) for d in descs])
    ss, counts = ol.build_suffstats(oracle, descs, hp, view, z, 4)
    lp = ol.logprior(counts, 1.0)
    S = oracle.score_rows(descs, hp, ss, lp, view)
    # recompute row 7 / group 2 by hand through the single-value entry point
    i, k = 7, 2
    tot, hoff, soff = lp[k], 0, 0
    for d, desc in enumerate(descs):
        m = oracle.model(desc)
        nh, ns = oracle.hp_size(m), oracle.ss_size(m)
        if not arr.mask["f%d" % d][i]:
            tot += oracle.score_value(m, hp[hoff:hoff + nh], ss[k, soff:soff + ns], [float(arr.data["f%d" % d][i])])
        hoff += nh; soff += ns
    assert abs(S[i, k] - tot) < 1e-12


# ---- marginal likelihoods: group::score_data (base.hpp:28), group_manager::score_assignment (:250-272) ----
def _marginal_cases():
    with open(os.path.join(GOLD, "score_data.json")) as f:
        return json.load(f)


def test_golden_score_data(oracle):
    cases = _marginal_cases()["cases"]
    assert len(cases) >= 18
    for c in cases:
        m = ol.OrcModel(FAM[c["family"]], c["dim"])
        got = oracle.score_data(m, c["hp"], c["ss"])
        tol = 1e-11 * max(1.0, abs(c["expect"])) * (50.0 if c["family"] == "niw" else 1.0)
        assert abs(got - c["expect"]) <= tol, (c["family"], c["source"], got, c["expect"])
        if "ref_vendor" in c:  # the reference's own in-tree inverse-Wishart partition function
            assert abs(got - c["ref_vendor"]) <= 1e-9 * max(1.0, abs(c["ref_vendor"]))


@pytest.mark.parametrize("desc", [cb.bb, cb.gp, cb.nich, cb.dd(7), cb.niw(3)])
def test_score_data_is_the_chain_of_predictives(oracle, desc):
    # log p(x_1..x_n) = sum_i log p(x_i | x_<i): pins score_data on the (golden-pinned) score_value / add_value
    rng = np.random.default_rng(11)
    m = oracle.model(desc)
    hp = oracle.flat_hp(desc)
    ss = np.zeros(oracle.ss_size(m))
    name = desc().name()
    chain = 0.0
    for _ in range(40):
        if name == "bb": x = float(rng.integers(0, 2))
        elif name == "dd": x = float(rng.integers(0, 7))
        elif name == "gp": x = float(rng.poisson(5.0))
        elif name == "nich": x = float(rng.normal(0.7, 1.1))
        else: x = rng.normal(0.2, 1.0, size=3)
        chain += oracle.score_value(m, hp, ss, x, 64)
        oracle.add_value(m, hp, ss, x, 64)
    got = oracle.score_data(m, hp, ss)
    assert abs(got - chain) <= 1e-10 * max(1.0, abs(chain)), (name, got, chain)
    assert oracle.score_data(m, hp, np.zeros_like(ss)) == pytest.approx(0.0, abs=1e-12)  # an empty group


def test_golden_score_assignment(oracle):
    for c in _marginal_cases()["crp"]:
        f32 = oracle.score_assignment(c["assign"], c["alpha"], prec=32)   # the reference's float loop
        f64 = oracle.score_assignment(c["assign"], c["alpha"], prec=64)   # closed form
        assert abs(f64 - c["expect"]) <= 1e-11 * max(1.0, abs(c["expect"]))
        assert abs(f32 - c["expect"]) <= 2e-6 * len(c["assign"]) ** 0.5 * max(1.0, abs(c["expect"]))


def test_bnb_predictive_normalises_and_chains(oracle):
    # sum_x exp(score) ~ 1 over a long range (heavy tail: Beta(3, 9.x) mixing), and the part of the marginal that
    # (count, sum) determine equals the chain of predictives minus the data-only term sum_i log C(x_i + r - 1, x_i)
    from scipy.special import gammaln
    m = ol.OrcModel(ol.BNB, 0)
    hp = np.array([2.0, 3.0, 2.0])
    ss = np.array([4.0, 9.0])
    tot = sum(np.exp(oracle.score_value(m, hp, ss, float(x))) for x in range(0, 20000))
    assert abs(tot - 1.0) < 1e-6
    rng = np.random.default_rng(3)
    ss = np.zeros(2)
    chain = data_term = 0.0
    for _ in range(30):
        x = float(rng.negative_binomial(2, 0.4))
        chain += oracle.score_value(m, hp, ss, x)
        data_term += gammaln(hp[2] + x) - gammaln(hp[2]) - gammaln(x + 1.0)
        oracle.add_value(m, hp, ss, x)
    assert abs(oracle.score_data(m, hp, ss) - (chain - data_term)) < 1e-9 * max(1.0, abs(chain))


def test_bbnc_follows_the_in_tree_source(oracle):
    # src/models/bbnc.cpp: score_value :46-53, add / remove :21-44, score_data :61-73 (hand-computed here)
    m = ol.OrcModel(ol.BBNC, 0)
    hp = np.array([2.0, 3.0])
    ss = np.array([0.3, 0.0, 0.0])
    assert oracle.score_value(m, hp, ss, 1.0) == pytest.approx(np.log(0.3), rel=1e-15)
    assert oracle.score_value(m, hp, ss, 0.0) == pytest.approx(np.log(0.7), rel=1e-15)
    assert oracle.score_value(m, hp, ss, 1.0, 32) == pytest.approx(np.log(0.3), rel=3e-7)
    for x in (1, 1, 0, 1):
        oracle.add_value(m, hp, ss, float(x))
    assert ss.tolist() == [0.3, 3.0, 1.0]
    from scipy.special import betaln
    want = (2 - 1) * np.log(0.3) + (3 - 1) * np.log(0.7) - betaln(2, 3) + 3 * np.log(0.3) + 1 * np.log(0.7)
    assert oracle.score_data(m, hp, ss) == pytest.approx(want, rel=1e-13)
    oracle.remove_value(m, hp, ss, 1.0)
    assert ss.tolist() == [0.3, 2.0, 1.0]
    assert oracle.score_data(m, hp, np.array([1.5, 0.0, 0.0])) == -np.inf      # p outside [0, 1]


def test_dm_follows_the_in_tree_source(oracle):
    # src/models/dm.cpp: add / remove :9-36 (counts += x, ratio += lgamma(sum x + 1) - sum lgamma(x_i + 1)),
    # score_value :38-76, score_data :79-95 = ratio + the Dirichlet-multinomial evidence of the pooled counts
    from scipy.special import gammaln
    rng = np.random.default_rng(12)
    C = 6
    m = ol.OrcModel(ol.DM, C)
    hp = rng.uniform(0.3, 2.0, C)
    ss = np.zeros(C + 1)
    xs = [rng.multinomial(int(rng.integers(0, 30)), rng.dirichlet(np.ones(C))).astype(float) for _ in range(25)]
    chain = 0.0
    for x in xs:   # chain rule: the marginal likelihood is the product of the predictives
        chain += oracle.score_value(m, hp, ss, x)
        oracle.add_value(m, hp, ss, x)
    X = np.asarray(xs)
    assert np.array_equal(ss[:C], X.sum(0))
    ratio = sum(gammaln(x.sum() + 1) - gammaln(x + 1).sum() for x in xs)
    assert ss[C] == pytest.approx(ratio, rel=1e-12)
    assert oracle.score_data(m, hp, ss) == pytest.approx(chain, rel=1e-11)
    # an all-zero row scores log 1 = 0 and the predictive over the C one-count rows sums to 1
    assert oracle.score_value(m, hp, ss, np.zeros(C)) == pytest.approx(0.0, abs=1e-12)
    assert sum(np.exp(oracle.score_value(m, hp, ss, np.eye(C)[i])) for i in range(C)) == pytest.approx(1.0, rel=1e-12)
    for x in reversed(xs):
        oracle.remove_value(m, hp, ss, x)
    assert np.all(ss[:C] == 0) and abs(ss[C]) < 1e-9


def test_sample_value_draws_follow_the_predictive(oracle):
    # group::sample_value (models/base.hpp:29): the checker's draws against the (golden-pinned) predictive --
    # chi-square for the discrete families, Kolmogorov-Smirnov for nich, moments for niw
    from scipy import stats
    n = 40000
    for fam, dim, hp, ss, support in [(ol.BB, 0, [0.7, 1.4], [6, 3], 2), (ol.BBNC, 0, [1.0, 1.0], [0.27, 4, 9], 2),
                                      (ol.DD, 5, [0.5, 1, 2, 0.1, 1], [9, 3, 0, 4, 1, 1], 5),
                                      (ol.GP, 0, [2.0, 0.5], [6, 31, 0.0], 60), (ol.BNB, 0, [3.0, 2.0, 4.0], [5, 9], 400)]:
        m = ol.OrcModel(fam, dim)
        x = oracle.sample_value(m, hp, ss, 99, 0, n)
        assert np.all(x == np.floor(x)) and x.min() >= 0
        pmf = np.exp([oracle.score_value(m, hp, ss, [v]) for v in range(support)])
        obs = np.bincount(np.minimum(x.astype(int), support - 1), minlength=support).astype(float)
        exp = pmf * n
        exp[-1] += n - exp.sum()       # the tail beyond the listed support
        keep = exp > 5
        chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum() + (obs[~keep].sum() - exp[~keep].sum()) ** 2 / max(exp[~keep].sum(), 1.0)
        assert chi2 < stats.chi2.ppf(1 - 1e-6, keep.sum()), (fam, chi2)
    m = ol.OrcModel(ol.NICH, 0)
    hp, ss = [0.5, 2.0, 1.5, 3.0], [7, 1.2, 9.5]
    x = oracle.sample_value(m, hp, ss, 5, 1000, n)
    kappa, nu = 2.0 + 7, 3.0 + 7
    mu = (2.0 * 0.5 + 7 * 1.2) / kappa
    sigmasq = (3.0 * 1.5 + 9.5 + 7 * 2.0 * (0.5 - 1.2) ** 2 / kappa) / nu
    ks = stats.kstest(x, stats.t(df=nu, loc=mu, scale=np.sqrt(sigmasq * (kappa + 1) / kappa)).cdf)
    assert ks.pvalue > 1e-4, ks
    # a shape below 1 takes the boosted branch of the gamma sampler: nu' = 0.6
    x = oracle.sample_value(m, [0.0, 1.0, 1.0, 0.6], [0, 0, 0], 5, 0, n)
    assert stats.kstest(x, stats.t(df=0.6, loc=0.0, scale=np.sqrt(2.0)).cdf).pvalue > 1e-4
    d = 3
    m = ol.OrcModel(ol.NIW, d)
    rng = np.random.default_rng(3)
    data = rng.normal(size=(30, d)) @ np.array([[1.0, 0.4, 0.0], [0.0, 1.0, 0.3], [0.0, 0.0, 0.5]]) + [1.0, -2.0, 0.5]
    hp = np.concatenate([np.zeros(d), [1.0], np.eye(d).ravel(), [d + 2.0]])
    ss = np.concatenate([[30], data.sum(0), (data.T @ data).ravel()])
    x = oracle.sample_value(m, hp, ss, 8, 0, n)
    kn, nun = 31.0, d + 2.0 + 30
    mun = data.sum(0) / kn
    psin = np.eye(d) + data.T @ data - kn * np.outer(mun, mun)
    dof = nun - d + 1
    scale = psin * (kn + 1) / (kn * dof)
    assert np.allclose(x.mean(0), mun, atol=5 * np.sqrt(np.diag(scale).max() / n) * 2)
    assert np.allclose(np.cov(x.T), scale * dof / (dof - 2), rtol=0.05, atol=0.02)
    # the Mahalanobis radius / d of a multivariate t is F(d, dof)
    r = np.einsum("ni,ij,nj->n", x - mun, np.linalg.inv(scale), x - mun) / d
    assert stats.kstest(r, stats.f(d, dof).cdf).pvalue > 1e-4
    # draws are a pure function of (seed, counter + i)
    assert np.array_equal(oracle.sample_value(m, hp, ss, 8, 100, 5), x[100:105])
    with pytest.raises(RuntimeError):
        oracle.sample_value(ol.OrcModel(ol.DM, 3), [1, 1, 1], [0, 0, 0, 0], 1, 0, 1)   # dm.cpp:100-111


def _intree():
    with open(os.path.join(GOLD, "intree_models.json")) as f:
        return json.load(f)


def test_oracle_against_runs_of_the_references_own_python_models(oracle):
    # tests/golden/intree_models.json = outputs of /root/reference/microscopes/dbg/models/{bbnc,dm}.py run in the
    # build container (scripts/make_golden_intree.py): the one place the reference tree itself computes this path
    gold = _intree()
    m = ol.OrcModel(ol.BBNC, 0)
    for r in gold["bbnc"]:
        hp = np.array([r["alpha"], r["beta"]])
        ss = np.array([r["p"], 0.0, 0.0])
        for v in r["values"]:
            oracle.add_value(m, hp, ss, float(v))
        assert ss[1:].tolist() == r["after_add"]
        assert oracle.score_value(m, hp, ss, 1.0) == pytest.approx(r["score_true"], rel=1e-14)
        assert oracle.score_value(m, hp, ss, 0.0) == pytest.approx(r["score_false"], rel=1e-14)
        assert oracle.score_data(m, hp, ss) == pytest.approx(r["score_data"], rel=1e-12)
        for v in r["removed"]:
            oracle.remove_value(m, hp, ss, float(v))
        assert ss[1:].tolist() == r["after_remove"]
        assert oracle.score_data(m, hp, ss) == pytest.approx(r["score_data_after_remove"], rel=1e-12)
    for r in gold["dm"]:
        C = r["dim"]
        m = ol.OrcModel(ol.DM, C)
        hp = np.ones(C)
        ss = np.zeros(C + 1)
        for x in r["rows"]:
            oracle.add_value(m, hp, ss, np.asarray(x, float))
        assert ss[:C].tolist() == r["counts_after_add"]
        assert ss[C] == pytest.approx(r["ratio_after_add"], rel=1e-12, abs=1e-12)
        for x in r["rows"][:r["removed"]]:
            oracle.remove_value(m, hp, ss, np.asarray(x, float))
        assert ss[:C].tolist() == r["counts_after_remove"]
        assert ss[C] == pytest.approx(r["ratio_after_remove"], rel=1e-12, abs=1e-10)
