/*
 * msb_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  Nothing under common_b200/ links or calls it.
 *
 * It restates, in plain C, the algorithm of the reference hot path
 * (datamicroscopes/common).  The per-family arithmetic lives in the
 * third-party library `distributions` (>=2.0.23, conda/microscopes-common/
 * meta.yaml:14,21), which is NOT vendored in the reference tree and not
 * installed here, so the closed forms are restated from the published
 * conjugate-prior maths and anchored on the reference's
 * own call sites:
 *   - forwards:      include/microscopes/models/distributions.hpp:266-285
 *   - field names:   include/microscopes/models/distributions.hpp:21-56,165-199
 *   - value casts:   include/microscopes/common/runtime_type.hpp:145-166
 *   - row layout:    include/microscopes/common/runtime_type.hpp:123-134,
 *                    src/common/recarray/dataview.cpp:97-104
 *   - CRP term:      include/microscopes/common/group_manager.hpp:274-283
 *   - sampler:       include/microscopes/common/util.hpp:125-156
 *   - in-tree closed forms it is pinned against (tests/golden/):
 *                    microscopes/common/vendor/stats.py:235-255
 *
 * PARITY STATUS: the reference tree holds no golden vector for score_value /
 * add_value / remove_value => at the `distributions`
 * boundary parity is UNPINNED.  What IS pinned: bb and niw predictive against
 * the reference's own Python closed forms (vendor/stats.py, run in the build
 * container, fixtures in tests/golden/), bbnc and dm bookkeeping against runs
 * of the reference's in-tree Python models (microscopes/dbg/models/{bbnc,dm}.py ->
 * tests/golden/intree_models.json), every family against scipy in fp64,
 * the dataview layout against the reference's real headers (oracle/_ref).
 */
#ifndef MSB_ORACLE_H
#define MSB_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_BB = 0, ORC_BNB = 1, ORC_GP = 2, ORC_NICH = 3, ORC_DD = 4, ORC_NIW = 5, ORC_BBNC = 6, ORC_DM = 7 };
typedef struct orc_model { int32_t family; uint32_t dim; } orc_model;
typedef struct orc_type { int32_t prim; uint32_t n; int32_t vec; } orc_type;

/* flat field vectors, same order as include/mscope_b200.h enum msb_family */
size_t orc_hp_size(const orc_model *m);
size_t orc_ss_size(const orc_model *m);
size_t orc_prim_size(int prim);
double orc_cast(const uint8_t *px, int prim); /* runtime_cast::cast */

/* prec = 64: closed form in double.  prec = 32: float arithmetic + libm
 * logf/lgammaf in the operation order of the upstream fp32 code ([R]). */
double orc_score_value(const orc_model *m, const double *hp, const double *ss, const double *x, int prec);
void orc_add_value(const orc_model *m, const double *hp, double *ss, const double *x, int prec);
void orc_remove_value(const orc_model *m, const double *hp, double *ss, const double *x, int prec);

/* group::score_data (models/base.hpp:28): log marginal likelihood of the data summarised by ss, fp64 closed form */
double orc_score_data(const orc_model *m, const double *hp, const double *ss);
/* group_manager::score_assignment (group_manager.hpp:250-272): float loop in entity order / fp64 closed form */
float orc_score_assignment(const int64_t *assign, size_t n, float alpha);
double orc_score_assignment64(const int64_t *assign, size_t n, double alpha);

/* batched K x D loop (entity_state.hpp:57-72 semantics, frozen suffstats).
 * hp: concatenation over features; ss: K blocks of the concatenation over features.
 * out[(i-row_lo)*K + k] = logprior[k] + sum over unmasked features. */
void orc_score_rows(const orc_model *models, size_t D, const double *hp, const double *ss, size_t K,
                    const double *logprior, const uint8_t *data, const uint8_t *mask,
                    const orc_type *types, size_t row_lo, size_t row_hi, int prec, int nthreads,
                    double *out);
void orc_score_rows_f32(const orc_model *models, size_t D, const double *hp, const double *ss, size_t K,
                    const double *logprior, const uint8_t *data, const uint8_t *mask,
                    const orc_type *types, size_t row_lo, size_t row_hi, int nthreads,
                    float *out);

/* sampler, util.hpp:125-156, with the exp of DESIGN.md "msb_expf" */
float orc_expf(float x);
int64_t orc_sample_discrete_log(const float *scores, size_t K, float u);
void orc_sample_rows(const float *scores, size_t nrows, size_t K, size_t ld, const float *u, int32_t *out);
/* the same with glibc's expf (what util.hpp:131 really calls): to measure how often a draw differs */
int64_t orc_sample_discrete_log_libm(const float *scores, size_t K, float u);
void orc_sample_rows_libm(const float *scores, size_t nrows, size_t K, size_t ld, const float *u, int32_t *out);
float orc_philox_u01(uint64_t seed, uint64_t row, uint64_t sweep);
void orc_philox_u01_rows(uint64_t seed, uint64_t row_lo, size_t n, uint64_t sweep, float *out);
void orc_philox_raw(uint64_t seed, uint64_t row, uint64_t sweep, uint32_t out[4]);

/* apply remove(old)/add(new) for every row of [row_lo,row_hi) in row order.
 * assign_old/new: column (group) index per row or -1. ss as in orc_score_rows. */
void orc_update_rows(const orc_model *models, size_t D, const double *hp, double *ss, size_t K,
                     double *group_counts, const uint8_t *data, const uint8_t *mask, const orc_type *types,
                     size_t row_lo, size_t row_hi, const int32_t *assign_old, const int32_t *assign_new,
                     int prec);

/* group::sample_value (models/base.hpp:29): n draws from the posterior predictive, draw i from the Philox stream
 * (seed, counter + i); out[n * width], width = dim for niw.  Returns -1 for dm (unimplemented upstream, dm.cpp:100-111). */
int orc_sample_value(const orc_model *m, const double *hp, const double *ss, uint64_t seed, uint64_t counter, size_t n,
                     double *out);

#ifdef __cplusplus
}
#endif
#endif
