// test_adapter.cpp -- the reference's own driver loop (bin/perf_group.cpp:76-125) on the GPU adapters,
// plus the batched form.  Compiles against include/microscopes_b200/plugin_api.hpp, or, with
// -DMSB_USE_REFERENCE_HEADERS -I<reference>/include, against the reference's real models/base.hpp.
//
//   test_adapter compile-only   -> exits 0 without touching the GPU
//   test_adapter                -> runs on cuda:0, prints "ok" lines, exits non-zero on mismatch
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "microscopes_b200/gpu_models.hpp"

using namespace microscopes;
using namespace microscopes::common;
using namespace microscopes::models;

static int fails = 0;
#define EXPECT_NEAR(a, b, tol)                                                                   \
  do {                                                                                           \
    const double _a = (a), _b = (b);                                                             \
    if (!(std::fabs(_a - _b) <= (tol) * std::fmax(1.0, std::fabs(_b)))) {                        \
      std::printf("FAIL %s:%d %s = %.9g, expected %.9g\n", __FILE__, __LINE__, #a, _a, _b);      \
      fails++;                                                                                   \
    }                                                                                            \
  } while (0)

// "bags": host only (no device).  For every family: fill the hypers and one group through the named mutators with fixed
// values, print "<family> <dim> <hex of get_hp()> <hex of get_ss()>"; then, for every line "<family> <dim> <hp hex> <ss hex>"
// on stdin, load the bags into fresh objects with set_hp / set_ss and print what get_hp / get_ss return afterwards.
static std::string to_hex(const std::string &b) {
  static const char *d = "0123456789abcdef";
  std::string o;
  for (unsigned char c : b) { o.push_back(d[c >> 4]); o.push_back(d[c & 15]); }
  return o.empty() ? "-" : o;
}
static std::string from_hex(const std::string &h) {
  std::string o;
  if (h == "-") return o;
  for (size_t i = 0; i + 1 < h.size(); i += 2) o.push_back((char)std::stoi(h.substr(i, 2), nullptr, 16));
  return o;
}
static void fill(value_mutator m, double base) {
  for (unsigned i = 0; i < m.shape(); i++) m.set<double>(base + i, i);
}
static int bags_mode() {
  rng_t r(1);
  struct fam { int family; unsigned dim; std::vector<const char *> hp, ss; };
  const std::vector<fam> fams = {
      {MSB_FAMILY_BB, 0, {"alpha", "beta"}, {"heads", "tails"}},
      {MSB_FAMILY_BNB, 0, {"alpha", "beta", "r"}, {"count", "sum"}},
      {MSB_FAMILY_GP, 0, {"alpha", "inv_beta"}, {"count", "sum", "log_prod"}},
      {MSB_FAMILY_NICH, 0, {"mu", "kappa", "sigmasq", "nu"}, {"count", "mean", "count_times_variance"}},
      {MSB_FAMILY_DD, 5, {"alphas"}, {"counts"}},
      {MSB_FAMILY_NIW, 3, {"mu", "kappa", "psi", "nu"}, {"count", "sum_x", "sum_xxT"}},
      {MSB_FAMILY_BBNC, 0, {"alpha", "beta"}, {"p", "heads", "tails"}},
      {MSB_FAMILY_DM, 4, {"alphas"}, {"counts", "ratio"}}};
  for (const auto &f : fams) {
    auto h = gpu_model(f.family, f.dim).create_hypers();
    double base = 1.5;
    for (auto k : f.hp) { fill(h->get_hp_mutator(k), std::strcmp(k, "r") == 0 ? 21.0 : base); base += 10.0; }
    auto g = h->create_group(r);
    base = 3.0;
    for (auto k : f.ss) { fill(g->get_ss_mutator(k), std::strcmp(k, "p") == 0 ? 0.25 : base); base += 7.0; }
    std::printf("%d %u %s %s\n", f.family, f.dim, to_hex(h->get_hp()).c_str(), to_hex(g->get_ss()).c_str());
  }
  std::printf("--\n");
  int family; unsigned dim; char hp[1 << 16], ss[1 << 16];
  while (std::scanf("%d %u %65535s %65535s", &family, &dim, hp, ss) == 4) {
    auto h = gpu_model(family, dim).create_hypers();
    auto g = h->create_group(r);
    try {
      h->set_hp(from_hex(hp));
      g->set_ss(from_hex(ss));
      std::printf("%d %u %s %s\n", family, dim, to_hex(h->get_hp()).c_str(), to_hex(g->get_ss()).c_str());
    } catch (const std::runtime_error &e) { std::printf("%d %u error %s\n", family, dim, e.what()); }
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 1 && std::strcmp(argv[1], "compile-only") == 0) return 0;
  if (argc > 1 && std::strcmp(argv[1], "bags") == 0) return bags_mode();
  rng_t r(73);

  // ---- perf_group.cpp loop: D bb features, 1 row, add -> remove -> score -----------------------
  const size_t D = 16;
  bool data[D];
  for (size_t i = 0; i < D; i++) data[i] = std::bernoulli_distribution(0.5)(r);
  std::vector<runtime_type> types(D, runtime_type(TYPE_B));
  std::vector<std::shared_ptr<hypers>> shares;
  std::vector<std::shared_ptr<group>> groups;
  for (size_t i = 0; i < D; i++) {
    shares.emplace_back(gpu_model(MSB_FAMILY_BB).create_hypers());
    shares.back()->get_hp_mutator("alpha").set<float>(2.0);
    shares.back()->get_hp_mutator("beta").set<float>(2.0);
    groups.emplace_back(shares.back()->create_group(r));
  }
  float score = 0.f;
  for (size_t i = 0; i < D; i++) {
    value_accessor v(reinterpret_cast<const uint8_t *>(&data[i]), nullptr, types[i]);
    groups[i]->add_value(*shares[i], v, r);
    groups[i]->add_value(*shares[i], v, r);
    groups[i]->remove_value(*shares[i], v, r);
    score += groups[i]->score_value(*shares[i], v, r);
  }
  // after one add of x: P(x) = (2 + 1) / (4 + 1)
  EXPECT_NEAR(score, D * std::log(3.0 / 5.0), 1e-5);
  std::printf("ok perf_group loop on gpu_group: score %.6f\n", score);

  // ---- the other families through the same interface ---------------------------------------------
  {
    auto h = gpu_model(MSB_FAMILY_NICH).create_hypers();
    h->get_hp_mutator("mu").set<float>(0.5f);
    h->get_hp_mutator("kappa").set<float>(2.0f);
    h->get_hp_mutator("sigmasq").set<float>(1.5f);
    h->get_hp_mutator("nu").set<float>(3.0f);
    auto g = h->create_group(r);
    const float xs[4] = {1.0f, 2.5f, -0.5f, 1.25f};
    for (float x : xs) g->add_value(*h, value_accessor(&x), r);
    const double n = 4, mean = (1.0 + 2.5 - 0.5 + 1.25) / 4;
    double ctv = 0;
    for (float x : xs) ctv += (x - mean) * (x - mean);
    const double kn = 2 + n, nun = 3 + n, mun = (2 * 0.5 + n * mean) / kn;
    const double sn = (3 * 1.5 + ctv + n * 2 * (0.5 - mean) * (0.5 - mean) / kn) / nun;
    const double lam = kn / ((kn + 1) * sn), x = 0.75;
    const double want = std::lgamma(0.5 * nun + 0.5) - std::lgamma(0.5 * nun) + 0.5 * std::log(lam / (M_PI * nun)) -
                        (0.5 * nun + 0.5) * std::log1p(lam * (x - mun) * (x - mun) / nun);
    const float xf = (float)x;
    EXPECT_NEAR(g->score_value(*h, value_accessor(&xf), r), want, 1e-5);
    double cnt = g->get_ss_mutator("count").accessor().get<double>(0);
    EXPECT_NEAR(cnt, 4.0, 0);
    EXPECT_NEAR(g->get_ss_mutator("mean").accessor().get<double>(0), mean, 1e-12);
    std::printf("ok nich through base.hpp interface\n");
  }
  {
    auto m = gpu_model(MSB_FAMILY_DD, 300);  // beyond the reference's DirichletDiscrete<128> (distributions.hpp:80-81)
    auto h = m.create_hypers();
    auto g = h->create_group(r);
    const int32_t x = 299;
    for (int i = 0; i < 3; i++) g->add_value(*h, value_accessor(&x), r);
    EXPECT_NEAR(g->score_value(*h, value_accessor(&x), r), std::log(4.0 / 303.0), 1e-5);
    const int64_t y = 7;  // any primitive type may back the value (runtime_cast)
    EXPECT_NEAR(g->score_value(*h, value_accessor(&y), r), std::log(1.0 / 303.0), 1e-5);
    bool threw = false;
    try { h->get_hp_mutator("nope"); } catch (const std::runtime_error &) { threw = true; }
    if (!threw) { std::printf("FAIL unknown key did not throw\n"); fails++; }
    std::printf("ok dd(300) through base.hpp interface\n");
  }

  // ---- score_data, sample_value (base.hpp:28-29) and the in-tree models (bbnc, dm) ------------------------------
  {
    auto h = gpu_model(MSB_FAMILY_BB).create_hypers();
    auto g = h->create_group(r);
    const bool t = true, f = false;
    for (int i = 0; i < 3; i++) g->add_value(*h, value_accessor(&t), r);
    g->add_value(*h, value_accessor(&f), r);
    // Beta-Bernoulli evidence with alpha = beta = 1: B(1 + 3, 1 + 1) / B(1, 1) = 1 / 20
    EXPECT_NEAR(g->score_data(*h, r), std::log(1.0 / 20.0), 1e-6);
    int heads = 0;
    const int draws = 400;
    for (int i = 0; i < draws; i++) {
      bool x = false;
      value_mutator vm(&x);
      g->sample_value(*h, vm, r);
      heads += x;
    }
    // predictive P(1) = 4 / 6; 400 draws: mean 266.7, sd 9.4
    if (heads < 210 || heads > 320) { std::printf("FAIL bb sample_value: %d heads of %d\n", heads, draws); fails++; }
    std::printf("ok score_data / sample_value through base.hpp (bb: %d heads of %d, expected ~267)\n", heads, draws);
  }
  {
    auto h = gpu_model(MSB_FAMILY_NICH).create_hypers();
    auto g = h->create_group(r);
    const float xs[3] = {4.0f, 5.0f, 6.0f};
    for (int rep = 0; rep < 20; rep++) for (float x : xs) g->add_value(*h, value_accessor(&x), r);
    double sum = 0;
    for (int i = 0; i < 300; i++) { float x = 0; value_mutator vm(&x); g->sample_value(*h, vm, r); sum += x; }
    EXPECT_NEAR(sum / 300, 60 * 5.0 / 61, 0.05);  // posterior mean, predictive sd ~0.83 -> se 0.05
    std::printf("ok nich sample_value mean %.3f\n", sum / 300);
  }
  {
    auto m = gpu_model(MSB_FAMILY_BBNC);
    auto h = m.create_hypers();
    auto g = h->create_group(r);
    g->get_ss_mutator("p").set<float>(0.25f);
    const bool t = true;
    EXPECT_NEAR(g->score_value(*h, value_accessor(&t), r), std::log(0.25), 1e-6);  // bbnc.cpp:46-53
    g->add_value(*h, value_accessor(&t), r);
    EXPECT_NEAR(g->get_ss_mutator("heads").accessor().get<double>(0), 1.0, 0);
    // bbnc.cpp:61-73 with alpha = beta = 1: log Beta(p; 1, 1) + heads log p + tails log(1 - p)
    EXPECT_NEAR(g->score_data(*h, r), std::log(0.25), 1e-6);
    std::printf("ok bbnc through base.hpp interface\n");
  }
  {
    auto m = gpu_model(MSB_FAMILY_DM, 3);
    auto h = m.create_hypers();
    auto g = h->create_group(r);
    const int32_t a[3] = {2, 0, 1}, b[3] = {1, 1, 0};
    const runtime_type vt(TYPE_I32, 3);
    g->add_value(*h, value_accessor(reinterpret_cast<const uint8_t *>(a), nullptr, vt), r);
    // dm.cpp:38-76: e = (3, 1, 2), E = 6, x = (1, 1, 0): 2!/(1! 1! 0!) * (3 * 1) / (6 * 7)
    EXPECT_NEAR(g->score_value(*h, value_accessor(reinterpret_cast<const uint8_t *>(b), nullptr, vt), r), std::log(2.0 * 3.0 / 42.0), 1e-5);
    EXPECT_NEAR(g->get_ss_mutator("ratio").accessor().get<double>(0), std::log(3.0), 1e-12);  // 3! / (2! 0! 1!)
    // dm.cpp:79-95: ratio + evidence of (2, 0, 1) under alpha = 1: 3 * (1*2 * 1) / (3*4*5)
    EXPECT_NEAR(g->score_data(*h, r), std::log(3.0 * 2.0 / 60.0), 1e-6);
    bool threw = false;
    int32_t out3[3];
    value_mutator vm(reinterpret_cast<uint8_t *>(out3), vt);
    try { g->sample_value(*h, vm, r); } catch (const std::runtime_error &e) { threw = std::string(e.what()).find("unimplemented") != std::string::npos; }
    if (!threw) { std::printf("FAIL dm sample_value did not throw\n"); fails++; }  // dm.cpp:100-111
    std::printf("ok dm through base.hpp interface\n");
  }

  // ---- batched: same numbers as the per-value loop --------------------------------------------------
  {
    const size_t N = 257, K = 3;
    struct __attribute__((packed)) rec { bool b; uint32_t c; float f; };
    std::vector<rec> rows(N);
    for (size_t i = 0; i < N; i++) {
      rows[i].b = std::bernoulli_distribution(0.3 + 0.2 * (i % K))(r);
      rows[i].c = std::poisson_distribution<uint32_t>(2.0 + 3.0 * (i % K))(r);
      rows[i].f = std::normal_distribution<float>(1.0f * (i % K), 1.0f)(r);
    }
    std::vector<runtime_type> t = {runtime_type(TYPE_B), runtime_type(TYPE_U32), runtime_type(TYPE_F32)};
    std::vector<std::shared_ptr<gpu_model>> ms = {std::make_shared<gpu_model>(MSB_FAMILY_BB), std::make_shared<gpu_model>(MSB_FAMILY_GP),
                                                 std::make_shared<gpu_model>(MSB_FAMILY_NICH)};
    batch_scorer bs(ms, reinterpret_cast<const uint8_t *>(rows.data()), nullptr, N, t, K + 1);
    bs.set_alpha(1.0);
    std::vector<int64_t> z(N);
    for (size_t k = 0; k < K; k++) bs.create_group();
    for (size_t i = 0; i < N; i++) z[i] = (int64_t)(i % K);
    bs.add_values(z);
    std::vector<float> S = bs.score_rows();
    // per-value loop over the same data
    std::vector<std::shared_ptr<hypers>> hs;
    std::vector<std::vector<std::shared_ptr<group>>> gs(K);
    for (auto &m : ms) hs.push_back(m->create_hypers());
    for (size_t k = 0; k < K; k++) for (auto &h : hs) gs[k].push_back(h->create_group(r));
    for (size_t i = 0; i < N; i++) {
      const uint8_t *p = reinterpret_cast<const uint8_t *>(&rows[i]);
      gs[i % K][0]->add_value(*hs[0], value_accessor(p, nullptr, t[0]), r);
      gs[i % K][1]->add_value(*hs[1], value_accessor(p + 1, nullptr, t[1]), r);
      gs[i % K][2]->add_value(*hs[2], value_accessor(p + 5, nullptr, t[2]), r);
    }
    for (size_t i : {size_t(0), size_t(100), N - 1}) {
      const uint8_t *p = reinterpret_cast<const uint8_t *>(&rows[i]);
      for (size_t k = 0; k < K; k++) {
        const double cnt = (double)((N - k + K - 1) / K);
        double s = std::log(cnt);
        s += gs[k][0]->score_value(*hs[0], value_accessor(p, nullptr, t[0]), r);
        s += gs[k][1]->score_value(*hs[1], value_accessor(p + 1, nullptr, t[1]), r);
        s += gs[k][2]->score_value(*hs[2], value_accessor(p + 5, nullptr, t[2]), r);
        EXPECT_NEAR(S[i * K + k], s, 2e-5);
      }
    }
    msb_sweep_result res = bs.sweep(73, 0);
    if (res.rows != N || res.units != N * K * 3) { std::printf("FAIL sweep result\n"); fails++; }
    std::printf("ok batch_scorer matches the per-value loop (%zu rows, %llu moved)\n", N, (unsigned long long)res.moved);

    // ---- a pass over host rows from C++: upload + prefetch on the copy stream, refresh swaps the column buffers,
    // the sweep is only enqueued, the assignments come back asynchronously --------------------------------------
    const std::vector<int64_t> before = bs.assignments();
    std::vector<int64_t> out(N, -7);
    for (int pass = 1; pass <= 2; pass++) {
      bs.upload(reinterpret_cast<const uint8_t *>(rows.data()));
      bs.refresh();
      bs.sweep_async(73, (uint64_t)pass);
      bs.assignments_async(out.data(), N);
      bs.assignments_wait();
      const msb_sweep_result r2 = bs.sweep_wait();
      const std::vector<int64_t> now = bs.assignments();
      size_t same = 0;
      for (size_t i = 0; i < N; i++) same += now[i] == out[i];
      if (same != N || r2.rows != N) { std::printf("FAIL streaming pass %d: %zu of %zu assignments agree\n", pass, same, N); fails++; }
    }
    // marginal likelihoods through the adapters (entity_state.hpp:74-86, group_manager.hpp:250-272)
    const float sa = bs.score_assignment(), sl = bs.score_likelihood();
    double per = 0.0;
    for (size_t k = 0; k < K; k++)
      for (size_t d = 0; d < 3; d++) per += bs.score_likelihood(d, k);
    EXPECT_NEAR(sl, per, 1e-4);
    if (!(sa < 0.f) || !(sl < 0.f)) { std::printf("FAIL score_assignment %g score_likelihood %g\n", sa, sl); fails++; }
    std::printf("ok streaming passes and marginal likelihoods from C++ (score_assignment %.3f, score_likelihood %.3f)\n", sa, sl);
  }
  if (fails) { std::printf("%d failure(s)\n", fails); return 1; }
  std::printf("all ok\n");
  return 0;
}
