// ref_harness.cpp -- CPU baseline through the REFERENCE'S OWN plugin API
// (TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE).
//
// Compiled by `make -C oracle ref` against the headers where they lie in
// /root/reference/include (nothing is copied into this repo):
//   microscopes/models/base.hpp                 abstract model / hypers / group
//   microscopes/common/recarray/dataview.hpp    row_accessor (header-only part)
//   microscopes/common/runtime_{type,value}.hpp value_accessor, runtime_cast
//   src/common/runtime_type.cpp                 primitive sizes
// plus oracle/ref_shim/distributions/random_fwd.hpp (3 lines).
//
// The conjugate-family arithmetic of the reference lives in the absent
// `distributions` library, so the classes below restate it in the reference's
// own style: float members, one virtual call per (row, group, feature), libm
// logf/lgammaf (upstream uses table-driven fast_log/fast_lgamma).  The loop is
// bin/perf_group.cpp:95-105 generalised to N rows x K groups and the
// entity_state.hpp:60-72 contract (score = log pseudocount + sum over features).
#include <microscopes/common/recarray/dataview.hpp>
#include <microscopes/models/base.hpp>
#include <microscopes/models/noop.hpp>  // the reference's own stub model: API overhead only

#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "msb_oracle.h"

using namespace microscopes;
using namespace microscopes::common;
using namespace microscopes::common::recarray;

namespace {

// ---- bb ---------------------------------------------------------------------
struct bb_hypers : public models::hypers {
  float alpha = 1.f, beta = 1.f;
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const bb_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &key) override {
    if (key == "alpha") return value_mutator(&alpha);
    if (key == "beta") return value_mutator(&beta);
    throw std::runtime_error("Unknown shared HP param key: " + key);
  }
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "bb"; }
};
struct bb_group : public models::group {
  unsigned heads = 0, tails = 0;
  void add_value(const models::hypers &, const value_accessor &v, rng_t &) override { (v.get<bool>(0) ? heads : tails) += 1; }
  void remove_value(const models::hypers &, const value_accessor &v, rng_t &) override { (v.get<bool>(0) ? heads : tails) -= 1; }
  float score_value(const models::hypers &m, const value_accessor &v, rng_t &) const override {
    const bb_hypers &h = static_cast<const bb_hypers &>(m);
    const float a = h.alpha + heads, b = h.beta + tails;
    return logf((v.get<bool>(0) ? a : b) / (a + b));
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {}
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const bb_group &>(g); }
  value_mutator get_ss_mutator(const std::string &key) override {
    if (key == "heads") return value_mutator(&heads);
    if (key == "tails") return value_mutator(&tails);
    throw std::runtime_error("Unknown group SS param key: " + key);
  }
  std::string debug_str() const override { return "bb"; }
};
std::shared_ptr<models::group> bb_hypers::create_group(rng_t &) const { return std::make_shared<bb_group>(); }

// ---- gp -----------------------------------------------------------------------
struct gp_hypers : public models::hypers {
  float alpha = 1.f, inv_beta = 1.f;
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const gp_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &key) override {
    if (key == "alpha") return value_mutator(&alpha);
    if (key == "inv_beta") return value_mutator(&inv_beta);
    throw std::runtime_error("Unknown shared HP param key: " + key);
  }
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "gp"; }
};
struct gp_group : public models::group {
  unsigned count = 0, sum = 0;
  float log_prod = 0.f;
  void add_value(const models::hypers &, const value_accessor &v, rng_t &) override {
    const unsigned x = v.get<unsigned>(0);
    ++count; sum += x; log_prod += lgammaf(x + 1.f);
  }
  void remove_value(const models::hypers &, const value_accessor &v, rng_t &) override {
    const unsigned x = v.get<unsigned>(0);
    --count; sum -= x; log_prod -= lgammaf(x + 1.f);
  }
  float score_value(const models::hypers &m, const value_accessor &v, rng_t &) const override {
    const gp_hypers &h = static_cast<const gp_hypers &>(m);
    const float a = h.alpha + sum, b = h.inv_beta + count;
    const float x = v.get<unsigned>(0);
    float s = lgammaf(a + x) - lgammaf(a) - lgammaf(x + 1.f);
    s += a * logf(b) - (a + x) * logf(1.f + b);
    return s;
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {}
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const gp_group &>(g); }
  value_mutator get_ss_mutator(const std::string &key) override {
    if (key == "count") return value_mutator(&count);
    if (key == "sum") return value_mutator(&sum);
    if (key == "log_prod") return value_mutator(&log_prod);
    throw std::runtime_error("Unknown group SS param key: " + key);
  }
  std::string debug_str() const override { return "gp"; }
};
std::shared_ptr<models::group> gp_hypers::create_group(rng_t &) const { return std::make_shared<gp_group>(); }

// ---- bbnc (src/models/bbnc.cpp restated: its translation unit needs protobuf and `distributions`) ------------
struct bbnc_hypers : public models::hypers {
  float alpha = 1.f, beta = 1.f;
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const bbnc_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &key) override {
    if (key == "alpha") return value_mutator(&alpha);
    if (key == "beta") return value_mutator(&beta);
    throw std::runtime_error("unknown key: " + key);
  }
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "bbnc"; }
};
struct bbnc_group : public models::group {
  float p = 0.5f;
  unsigned heads = 0, tails = 0;
  void add_value(const models::hypers &, const value_accessor &v, rng_t &) override { if (v.get<bool>(0)) heads++; else tails++; }
  void remove_value(const models::hypers &, const value_accessor &v, rng_t &) override { if (v.get<bool>(0)) heads--; else tails--; }
  float score_value(const models::hypers &, const value_accessor &v, rng_t &) const override {
    return v.get<bool>(0) ? logf(p) : logf(1. - p);
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {}
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const bbnc_group &>(g); }
  value_mutator get_ss_mutator(const std::string &key) override {
    if (key == "p") return value_mutator(&p);
    if (key == "heads") return value_mutator(&heads);
    if (key == "tails") return value_mutator(&tails);
    throw std::runtime_error("unknown key: " + key);
  }
  std::string debug_str() const override { return "bbnc"; }
};
std::shared_ptr<models::group> bbnc_hypers::create_group(rng_t &) const { return std::make_shared<bbnc_group>(); }

// ---- dm (src/models/dm.cpp restated statement by statement: its translation unit needs protobuf and `distributions`;
// lgammaf where upstream has fast_lgamma) ------------------------------------------------------------------------
struct dm_hypers : public models::hypers {
  unsigned dim;
  std::vector<float> alphas;
  explicit dm_hypers(unsigned d) : dim(d), alphas(d, 1.f) {}
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const dm_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &key) override {
    if (key == "alphas") return value_mutator(reinterpret_cast<uint8_t *>(&alphas[0]), runtime_type(TYPE_F32, dim));
    throw std::runtime_error("unknown key: " + key);
  }
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "dm"; }
};
struct dm_group : public models::group {
  std::vector<unsigned> counts;
  float ratio = 0.f;
  explicit dm_group(unsigned d) : counts(d, 0) {}
  void add_value(const models::hypers &, const value_accessor &value, rng_t &) override {  // dm.cpp:9-21
    unsigned count_sum = 0;
    for (size_t i = 0; i < counts.size(); i++) {
      const unsigned ni = value.get<unsigned>(i);
      count_sum += ni;
      counts[i] += ni;
      ratio -= lgammaf(ni + 1);
    }
    ratio += lgammaf(count_sum + 1);
  }
  void remove_value(const models::hypers &, const value_accessor &value, rng_t &) override {  // dm.cpp:23-36
    unsigned count_sum = 0;
    for (size_t i = 0; i < counts.size(); i++) {
      const unsigned ni = value.get<unsigned>(i);
      count_sum += ni;
      counts[i] -= ni;
      ratio += lgammaf(ni + 1);
    }
    ratio -= lgammaf(count_sum + 1);
  }
  float score_value(const models::hypers &m, const value_accessor &value, rng_t &) const override {  // dm.cpp:38-76
    const dm_hypers &h = static_cast<const dm_hypers &>(m);
    float score = 0.;
    unsigned x_sum = 0;
    float a_sum = 0.;
    unsigned n_sum = 0;
    for (size_t i = 0; i < counts.size(); i++) {
      const unsigned xi = value.get<unsigned>(i);
      const float ai = h.alphas[i];
      const unsigned ni = counts[i];
      x_sum += xi;
      a_sum += ai;
      n_sum += ni;
      const float effective_ai = ai + ni;
      score += lgammaf(effective_ai + xi) - lgammaf(effective_ai);
      score -= lgammaf(xi + 1);
    }
    score += lgammaf(x_sum + 1);
    score += lgammaf(a_sum + n_sum) - lgammaf(a_sum + n_sum + x_sum);
    return score;
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {
    throw std::runtime_error("multinomial sampling unimplemented");  // dm.cpp:100-111
  }
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const dm_group &>(g); }
  value_mutator get_ss_mutator(const std::string &key) override {
    if (key == "counts") return value_mutator(reinterpret_cast<uint8_t *>(&counts[0]), runtime_type(TYPE_U32, (unsigned)counts.size()));
    if (key == "ratio") return value_mutator(&ratio);
    throw std::runtime_error("unknown key: " + key);
  }
  std::string debug_str() const override { return "dm"; }
};
std::shared_ptr<models::group> dm_hypers::create_group(rng_t &) const { return std::make_shared<dm_group>(dim); }

// ---- bnb ----------------------------------------------------------------------
struct bnb_hypers : public models::hypers {
  float alpha = 1.f, beta = 1.f;
  unsigned r = 1;
  float rf = 1.f;  // get_hp_mutator("r") hands out a float slot like the other fields; r is read from it
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const bnb_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &key) override {
    if (key == "alpha") return value_mutator(&alpha);
    if (key == "beta") return value_mutator(&beta);
    if (key == "r") return value_mutator(&rf);
    throw std::runtime_error("Unknown shared HP param key: " + key);
  }
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "bnb"; }
};
struct bnb_group : public models::group {
  unsigned count = 0, sum = 0;
  void add_value(const models::hypers &, const value_accessor &v, rng_t &) override { ++count; sum += v.get<unsigned>(0); }
  void remove_value(const models::hypers &, const value_accessor &v, rng_t &) override { --count; sum -= v.get<unsigned>(0); }
  float score_value(const models::hypers &m, const value_accessor &v, rng_t &) const override {
    const bnb_hypers &h = static_cast<const bnb_hypers &>(m);
    const float r = h.rf, a = h.alpha + r * count, b = h.beta + sum, x = v.get<unsigned>(0);
    float s = lgammaf(r + x) - lgammaf(r) - lgammaf(x + 1.f);
    s += lgammaf(a + r) + lgammaf(b + x) - lgammaf(a + r + b + x);
    s -= lgammaf(a) + lgammaf(b) - lgammaf(a + b);
    return s;
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {}
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const bnb_group &>(g); }
  value_mutator get_ss_mutator(const std::string &key) override {
    if (key == "count") return value_mutator(&count);
    if (key == "sum") return value_mutator(&sum);
    throw std::runtime_error("Unknown group SS param key: " + key);
  }
  std::string debug_str() const override { return "bnb"; }
};
std::shared_ptr<models::group> bnb_hypers::create_group(rng_t &) const { return std::make_shared<bnb_group>(); }

// ---- nich ---------------------------------------------------------------------
struct nich_hypers : public models::hypers {
  float mu = 0.f, kappa = 1.f, sigmasq = 1.f, nu = 1.f;
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const nich_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &key) override {
    if (key == "mu") return value_mutator(&mu);
    if (key == "kappa") return value_mutator(&kappa);
    if (key == "sigmasq") return value_mutator(&sigmasq);
    if (key == "nu") return value_mutator(&nu);
    throw std::runtime_error("Unknown shared HP param key: " + key);
  }
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "nich"; }
};
struct nich_group : public models::group {
  unsigned count = 0;
  float mean = 0.f, count_times_variance = 0.f;
  void add_value(const models::hypers &, const value_accessor &v, rng_t &) override {
    const float x = v.get<float>(0);
    ++count;
    const float delta = x - mean;
    mean += delta / count;
    count_times_variance += delta * (x - mean);
  }
  void remove_value(const models::hypers &, const value_accessor &v, rng_t &) override {
    const float x = v.get<float>(0);
    const float total = mean * count, delta = x - mean;
    --count;
    mean = count == 0 ? 0.f : (total - x) / count;
    if (count <= 1) count_times_variance = 0.f;
    else count_times_variance -= delta * (x - mean);
  }
  float score_value(const models::hypers &m, const value_accessor &v, rng_t &) const override {
    const nich_hypers &h = static_cast<const nich_hypers &>(m);
    const float mu1 = h.mu - mean;
    const float kappa = h.kappa + count;
    const float mu = (h.kappa * h.mu + mean * count) / kappa;
    const float nu = h.nu + count;
    const float sigmasq = 1.f / nu * (h.nu * h.sigmasq + count_times_variance + (count * h.kappa * mu1 * mu1) / kappa);
    const float lambda = kappa / ((kappa + 1.f) * sigmasq);
    const float t = v.get<float>(0) - mu;
    float s = lgammaf(0.5f * nu + 0.5f) - lgammaf(0.5f * nu) + 0.5f * logf(lambda / ((float)M_PI * nu));
    s += (-0.5f * nu - 0.5f) * logf(1.f + (lambda * t * t) / nu);
    return s;
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {}
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const nich_group &>(g); }
  value_mutator get_ss_mutator(const std::string &key) override {
    if (key == "count") return value_mutator(&count);
    if (key == "mean") return value_mutator(&mean);
    if (key == "count_times_variance") return value_mutator(&count_times_variance);
    throw std::runtime_error("Unknown group SS param key: " + key);
  }
  std::string debug_str() const override { return "nich"; }
};
std::shared_ptr<models::group> nich_hypers::create_group(rng_t &) const { return std::make_shared<nich_group>(); }

// ---- dd (runtime dim; the reference instantiates DirichletDiscrete<128>, distributions.hpp:80-81) ----
struct dd_hypers : public models::hypers {
  unsigned dim;
  std::vector<float> alphas;
  float alpha_sum = 0.f;
  explicit dd_hypers(unsigned d) : dim(d), alphas(d, 1.f) {}
  void refresh() { alpha_sum = 0.f; for (float a : alphas) alpha_sum += a; }
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const dd_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &key) override {
    if (key == "alphas") return value_mutator(reinterpret_cast<uint8_t *>(&alphas[0]), runtime_type(TYPE_F32, dim));
    throw std::runtime_error("Unknown shared HP param key: " + key);
  }
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "dd"; }
};
struct dd_group : public models::group {
  unsigned count_sum = 0;
  std::vector<unsigned> counts;
  explicit dd_group(unsigned d) : counts(d, 0) {}
  void add_value(const models::hypers &, const value_accessor &v, rng_t &) override { const int x = v.get<int>(0); ++count_sum; ++counts[x]; }
  void remove_value(const models::hypers &, const value_accessor &v, rng_t &) override { const int x = v.get<int>(0); --count_sum; --counts[x]; }
  float score_value(const models::hypers &m, const value_accessor &v, rng_t &) const override {
    const dd_hypers &h = static_cast<const dd_hypers &>(m);
    const int x = v.get<int>(0);
    return logf((h.alphas[x] + counts[x]) / (h.alpha_sum + count_sum));
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {}
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const dd_group &>(g); }
  value_mutator get_ss_mutator(const std::string &key) override {
    if (key == "count_sum") return value_mutator(&count_sum);
    if (key == "counts") return value_mutator(reinterpret_cast<uint8_t *>(&counts[0]), runtime_type(TYPE_U32, (unsigned)counts.size()));
    throw std::runtime_error("Unknown group SS param key: " + key);
  }
  std::string debug_str() const override { return "dd"; }
};
std::shared_ptr<models::group> dd_hypers::create_group(rng_t &) const { return std::make_shared<dd_group>(dim); }

// ---- niw: delegates to the C oracle (no Eigen here); recomputes the posterior per call like upstream ----
struct niw_hypers : public models::hypers {
  orc_model m;
  std::vector<double> hp;
  explicit niw_hypers(unsigned d) : m{ORC_NIW, d}, hp(orc_hp_size(&m), 0.0) {}
  hyperparam_bag_t get_hp() const override { return ""; }
  void set_hp(const hyperparam_bag_t &) override {}
  void set_hp(const models::hypers &s) override { *this = static_cast<const niw_hypers &>(s); }
  value_mutator get_hp_mutator(const std::string &) override { throw std::runtime_error("not supported"); }  // distributions.hpp:112-118
  std::shared_ptr<models::group> create_group(rng_t &rng) const override;
  std::string debug_str() const override { return "niw"; }
};
struct niw_group : public models::group {
  orc_model m;
  std::vector<double> ss;
  explicit niw_group(const orc_model &mm) : m(mm), ss(orc_ss_size(&mm), 0.0) {}
  std::vector<double> val(const value_accessor &v) const {
    std::vector<double> x(v.shape());
    for (unsigned i = 0; i < v.shape(); i++) x[i] = v.get<float>(i);  // element-wise cast loop, distributions.hpp:216-228
    return x;
  }
  void add_value(const models::hypers &h, const value_accessor &v, rng_t &) override {
    auto x = val(v); orc_add_value(&m, static_cast<const niw_hypers &>(h).hp.data(), ss.data(), x.data(), 64);
  }
  void remove_value(const models::hypers &h, const value_accessor &v, rng_t &) override {
    auto x = val(v); orc_remove_value(&m, static_cast<const niw_hypers &>(h).hp.data(), ss.data(), x.data(), 64);
  }
  float score_value(const models::hypers &h, const value_accessor &v, rng_t &) const override {
    auto x = val(v);
    return (float)orc_score_value(&m, static_cast<const niw_hypers &>(h).hp.data(), ss.data(), x.data(), 64);
  }
  float score_data(const models::hypers &, rng_t &) const override { return 0.f; }
  void sample_value(const models::hypers &, value_mutator &, rng_t &) const override {}
  suffstats_bag_t get_ss() const override { return ""; }
  void set_ss(const suffstats_bag_t &) override {}
  void set_ss(const models::group &g) override { *this = static_cast<const niw_group &>(g); }
  value_mutator get_ss_mutator(const std::string &) override { throw std::runtime_error("not supported"); }
  std::string debug_str() const override { return "niw"; }
};
std::shared_ptr<models::group> niw_hypers::create_group(rng_t &) const { return std::make_shared<niw_group>(m); }

struct ref_model : public models::model {
  orc_model m;
  explicit ref_model(const orc_model &mm) : m(mm) {}
  std::shared_ptr<models::hypers> create_hypers() const override {
    switch (m.family) {
      case ORC_BB: return std::make_shared<bb_hypers>();
      case ORC_BNB: return std::make_shared<bnb_hypers>();
      case ORC_BBNC: return std::make_shared<bbnc_hypers>();
      case ORC_GP: return std::make_shared<gp_hypers>();
      case ORC_NICH: return std::make_shared<nich_hypers>();
      case ORC_DD: return std::make_shared<dd_hypers>(m.dim);
      case ORC_DM: return std::make_shared<dm_hypers>(m.dim);
      case ORC_NIW: return std::make_shared<niw_hypers>(m.dim);
      default: throw std::runtime_error("unknown family");
    }
  }
  runtime_type get_runtime_type() const override {
    switch (m.family) {
      case ORC_BB: case ORC_BBNC: return runtime_type(TYPE_B);
      case ORC_BNB: case ORC_GP: return runtime_type(TYPE_U32);
      case ORC_NICH: return runtime_type(TYPE_F32);
      case ORC_DD: return runtime_type(TYPE_I32);
      case ORC_DM: return runtime_type(TYPE_I32, m.dim);  // dm.hpp:186-190
      default: return runtime_type(TYPE_F32, m.dim);
    }
  }
};

struct built_state {
  std::vector<std::shared_ptr<models::hypers>> hypers;               // D
  std::vector<std::vector<std::shared_ptr<models::group>>> groups;   // K x D
  std::vector<runtime_type> types;
  size_t rowsize = 0, maskrowsize = 0;
};

void set_f(value_mutator m, double v) { m.set<float>((float)v, 0); }
void set_u(value_mutator m, double v) { m.set<unsigned>((unsigned)v, 0); }

built_state build(const orc_model *models, size_t D, const double *hp, const double *ss, size_t K, const orc_type *types) {
  built_state st;
  rng_t rng(73);  // bin/perf_group.cpp:19
  std::vector<size_t> hpoff(D), ssoff(D);
  size_t ho = 0, so = 0;
  for (size_t d = 0; d < D; d++) { hpoff[d] = ho; ssoff[d] = so; ho += orc_hp_size(&models[d]); so += orc_ss_size(&models[d]); }
  const size_t SS = so;
  for (size_t d = 0; d < D; d++) {
    st.types.push_back(types[d].vec ? runtime_type((primitive_type)types[d].prim, types[d].n) : runtime_type((primitive_type)types[d].prim));
    auto h = ref_model(models[d]).create_hypers();
    const double *p = hp + hpoff[d];
    switch (models[d].family) {
      case ORC_BB: set_f(h->get_hp_mutator("alpha"), p[0]); set_f(h->get_hp_mutator("beta"), p[1]); break;
      case ORC_BNB: set_f(h->get_hp_mutator("alpha"), p[0]); set_f(h->get_hp_mutator("beta"), p[1]); set_f(h->get_hp_mutator("r"), p[2]); break;
      case ORC_BBNC: set_f(h->get_hp_mutator("alpha"), p[0]); set_f(h->get_hp_mutator("beta"), p[1]); break;
      case ORC_GP: set_f(h->get_hp_mutator("alpha"), p[0]); set_f(h->get_hp_mutator("inv_beta"), p[1]); break;
      case ORC_NICH:
        set_f(h->get_hp_mutator("mu"), p[0]); set_f(h->get_hp_mutator("kappa"), p[1]);
        set_f(h->get_hp_mutator("sigmasq"), p[2]); set_f(h->get_hp_mutator("nu"), p[3]); break;
      case ORC_DD: {
        auto mut = h->get_hp_mutator("alphas");
        for (unsigned i = 0; i < models[d].dim; i++) mut.set<float>((float)p[i], i);
        static_cast<dd_hypers &>(*h).refresh();
        break;
      }
      case ORC_NIW: static_cast<niw_hypers &>(*h).hp.assign(p, p + orc_hp_size(&models[d])); break;
      case ORC_DM: {
        auto mut = h->get_hp_mutator("alphas");
        for (unsigned i = 0; i < models[d].dim; i++) mut.set<float>((float)p[i], i);
        break;
      }
    }
    st.hypers.push_back(h);
  }
  const auto offs = runtime_type::GetOffsetsAndSize(st.types);  // runtime_type.hpp:123-134
  st.rowsize = offs.rowsize_; st.maskrowsize = offs.maskrowsize_;
  st.groups.resize(K);
  for (size_t k = 0; k < K; k++)
    for (size_t d = 0; d < D; d++) {
      auto g = st.hypers[d]->create_group(rng);
      const double *s = ss + k * SS + ssoff[d];
      switch (models[d].family) {
        case ORC_BB: set_u(g->get_ss_mutator("heads"), s[0]); set_u(g->get_ss_mutator("tails"), s[1]); break;
        case ORC_BNB: set_u(g->get_ss_mutator("count"), s[0]); set_u(g->get_ss_mutator("sum"), s[1]); break;
        case ORC_BBNC: set_f(g->get_ss_mutator("p"), s[0]); set_u(g->get_ss_mutator("heads"), s[1]); set_u(g->get_ss_mutator("tails"), s[2]); break;
        case ORC_GP: set_u(g->get_ss_mutator("count"), s[0]); set_u(g->get_ss_mutator("sum"), s[1]); set_f(g->get_ss_mutator("log_prod"), s[2]); break;
        case ORC_NICH: set_u(g->get_ss_mutator("count"), s[0]); set_f(g->get_ss_mutator("mean"), s[1]); set_f(g->get_ss_mutator("count_times_variance"), s[2]); break;
        case ORC_DD: {
          set_u(g->get_ss_mutator("count_sum"), s[0]);
          auto mut = g->get_ss_mutator("counts");
          for (unsigned i = 0; i < models[d].dim; i++) mut.set<unsigned>((unsigned)s[1 + i], i);
          break;
        }
        case ORC_NIW: static_cast<niw_group &>(*g).ss.assign(s, s + orc_ss_size(&models[d])); break;
        case ORC_DM: {
          auto mut = g->get_ss_mutator("counts");
          for (unsigned i = 0; i < models[d].dim; i++) mut.set<unsigned>((unsigned)s[i], i);
          set_f(g->get_ss_mutator("ratio"), s[models[d].dim]);
          break;
        }
      }
      st.groups[k].push_back(g);
    }
  return st;
}

}  // namespace

extern "C" {

// out[(i - row_lo) * K + k] = logprior[k] + sum over unmasked features of
// groups[k][d]->score_value(*hypers[d], acc.get(), rng)   (entity_state.hpp:60-72)
__attribute__((visibility("default"))) int ref_score_rows(
    const orc_model *models, size_t D, const double *hp, const double *ss, size_t K, const double *logprior,
    const uint8_t *data, const uint8_t *mask, const orc_type *types, size_t row_lo, size_t row_hi, int nthreads,
    float *out) {
  try {
    built_state st = build(models, D, hp, ss, K, types);
    if (nthreads < 1) nthreads = 1;
    // everything the loop reads is captured BY VALUE: captured references would live in this frame, next to the locals
    // the calling thread writes while it runs its own share, and every other thread would take those cache misses
    // (measured: 2 threads slower than 1)
    const built_state *stp = &st;
    auto worker = [=](size_t lo, size_t hi) {
      rng_t rng(73);
      const size_t rowsize = stp->rowsize, maskrowsize = stp->maskrowsize;
      const std::vector<runtime_type> *tys = &stp->types;
      for (size_t i = lo; i < hi; i++) {
        const bool *mrow = mask ? reinterpret_cast<const bool *>(mask) + maskrowsize * i : nullptr;
        row_accessor acc(data + rowsize * i, mrow, tys);  // what row_major_dataview::get() builds, dataview.cpp:97-104
        for (size_t k = 0; k < K; k++) {
          float s = (float)logprior[k];
          const std::shared_ptr<models::group> *gk = stp->groups[k].data();
          const std::shared_ptr<models::hypers> *hs = stp->hypers.data();
          acc.reset();
          for (size_t d = 0; d < D; d++, acc.bump())
            if (!acc.anymasked()) s += gk[d]->score_value(*hs[d], acc.get(), rng);
          out[(i - row_lo) * K + k] = s;
        }
      }
    };
    std::vector<std::thread> th;
    const size_t n = row_hi - row_lo;
    if (nthreads == 1) worker(row_lo, row_hi);
    else for (int t = 0; t < nthreads; t++) th.emplace_back(worker, row_lo + n * t / nthreads, row_lo + n * (t + 1) / nthreads);
    for (auto &t : th) t.join();
    return 0;
  } catch (const std::exception &) {
    return 1;
  }
}

// The reference's REAL noop model (include/microscopes/models/noop.hpp, compiled where it lies) in the loop of
// bin/perf_group.cpp:95-105: nothing but the plugin API's own cost per call (virtual dispatch, row_accessor::get /
// bump, value_accessor construction).  Returns ns per score_value call.
__attribute__((visibility("default"))) double ref_noop_overhead_ns(size_t D, size_t niters) {
  rng_t r(73);
  std::vector<uint8_t> data(D, 1);
  std::vector<runtime_type> types(D, runtime_type(TYPE_B));
  row_accessor acc(data.data(), nullptr, &types);
  models::noop_model m;
  std::vector<std::shared_ptr<models::hypers>> shares;
  std::vector<std::shared_ptr<models::group>> groups;
  for (size_t i = 0; i < D; i++) {
    shares.emplace_back(m.create_hypers());
    groups.emplace_back(shares.back()->create_group(r));
  }
  float score = 0.f;
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t n = 0; n < niters; n++) {
    acc.reset();
    for (size_t i = 0; i < D; i++, acc.bump()) score += groups[i]->score_value(*shares[i], acc.get(), r);
  }
  const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
  volatile float sink = score;
  (void)sink;
  return ns / (double)(D * niters);
}

// bin/perf_group.cpp:76-125 itself: D features x 1 row, niters x (add, remove, score); returns ns per call
__attribute__((visibility("default"))) double ref_perf_group(int family, unsigned dim, size_t D, size_t niters, double *score_out) {
  orc_model m{family, dim};
  rng_t r(73);
  std::vector<uint8_t> data(D);
  for (size_t i = 0; i < D; i++) data[i] = std::bernoulli_distribution(0.5)(r);
  std::vector<runtime_type> types(D, runtime_type(TYPE_B));
  row_accessor acc(data.data(), nullptr, &types);
  std::vector<std::shared_ptr<models::hypers>> shares;
  std::vector<std::shared_ptr<models::group>> groups;
  for (size_t i = 0; i < D; i++) {
    shares.emplace_back(ref_model(m).create_hypers());
    if (family == ORC_BB) { shares.back()->get_hp_mutator("alpha").set<float>(2.0, 0); shares.back()->get_hp_mutator("beta").set<float>(2.0, 0); }
    groups.emplace_back(shares.back()->create_group(r));
  }
  float score = 0.f;
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t n = 0; n < niters; n++) {
    acc.reset();
    for (size_t i = 0; i < acc.nfeatures(); i++, acc.bump()) groups[i]->add_value(*shares[i], acc.get(), r);
    acc.reset();
    for (size_t i = 0; i < acc.nfeatures(); i++, acc.bump()) groups[i]->remove_value(*shares[i], acc.get(), r);
    acc.reset();
    for (size_t i = 0; i < acc.nfeatures(); i++, acc.bump()) score += groups[i]->score_value(*shares[i], acc.get(), r);
  }
  const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
  if (score_out) *score_out = score;
  return ns / (double)(niters * D * 3);
}
}
