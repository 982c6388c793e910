// msb_niw_tc.cuh -- tensor-core (tcgen05) path for the NIW Mahalanobis term, dim == 64.
//
//   scores[n][k] += c0_k + c1_k * log1p(|W_k x_n - b_k|^2 / dof_k)        W_k = L_k^-1 (niw_prepare_kernel)
//
// As a GEMM: Y[n][(k,i)] = sum_j X[n][j] W_k[i][j]  (M = rows, N = 64 outputs x 4 groups = 256, K = 64).
// This is the one genuinely dense contraction of the hot path (BASELINE north_star item 4).
//
//   * B operand: the 4 groups of a block, split into tf32 hi / lo parts, in the UMMA K-major
//     no-swizzle core-matrix layout (8 rows x 16 bytes contiguous), 128 KB resident in shared memory
//     for as long as the CTA works on that group block (bulk-copied once, completes on an mbarrier).
//   * A operand: 128-row tiles of X, read from global memory as fp32 by 4 producer warps, split into
//     tf32 hi / lo and stored straight into the same core-matrix layout (no TMA tensor map needed);
//     two half-tile buffers (k < 32, k >= 32) ping-pong so conversion overlaps the MMAs.
//   * MMA: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32, M=128 N=256 K=8, three
//     products per k-step (hi*hi + hi*lo + lo*hi: fp32-class accuracy, |rel err| ~ 2^-21), fp32
//     accumulators in TMEM, 2 x 256 columns so the epilogue of tile t overlaps the MMAs of tile t+1.
//   * Epilogue: 4 warps tcgen05.ld their 32 TMEM lanes, subtract the bias, square-sum per group,
//     log1p, and add into the score matrix.
// Synchronisation is mbarrier-only (tcgen05.commit arrives on them); every wait is bounded (trap).
#pragma once
#include <cuda_runtime.h>
#include <string>

#include "../../include/mscope_b200.h"
#include "msb_score.cuh"

namespace msb {

namespace niwtc {
constexpr int D = 64;            // dimension this path is built for
constexpr int TM = 128;          // rows per tile (UMMA M)
constexpr int GB = 4;            // groups per block
constexpr bool TRIANGULAR = true;  // skip the structurally zero part of the lower-triangular W_k (see the MMA loop)
constexpr int TN = GB * D;       // UMMA N = 256
constexpr int A_HALF_BYTES = TM * 32 * 4;   // one k-half (32 k) of one part (hi or lo): 16 KB
constexpr int A_BYTES = 4 * A_HALF_BYTES;   // [half][part]: 64 KB
constexpr int B_PART_BYTES = TN * D * 4;    // 64 KB
constexpr int B_BYTES = 2 * B_PART_BYTES;   // hi + lo: 128 KB
constexpr int THREADS = 9 * 32;             // warps 0-3 producers, 4-7 epilogue, 8 MMA issuer
constexpr size_t SMEM_BYTES = (size_t)B_BYTES + A_BYTES + 2 * (TN + 16) * sizeof(float) + 16 * sizeof(uint64_t) + 64;

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor (sm_100): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // layout_type=0 (no swizzle) [61,64)
  const uint64_t lo = (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16);
  const uint64_t hi = (uint64_t)((sbo_bytes >> 4) & 0x3FFF) | (1ull << 14);
  return lo | (hi << 32);
}
// instruction descriptor, kind::tf32: D=F32 (bits 4-5 = 1), A=B=TF32 (bits 7-9, 10-12 = 2), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ constexpr uint32_t idesc_n(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of element (row, k) inside one part of an operand tile with `rowgroups` 8-row groups:
// [k/4][row/8][row%8][k%4]  (core matrix = 8 rows x 16 bytes, contiguous)
__device__ __host__ __forceinline__ uint32_t core_off(uint32_t row, uint32_t k, uint32_t rowgroups) {
  return (k >> 2) * (rowgroups * 128u) + (row >> 3) * 128u + (row & 7u) * 16u + (k & 3u) * 4u;
}
}  // namespace niwtc

// W[k][i][j] (fp32, row-major per group) -> per group block: [hi | lo] parts in UMMA layout; blocks padded with zeros
__global__ void niw_pack_b_kernel(const float *__restrict__ W, int ncols, float *__restrict__ Bop) {
  using namespace niwtc;
  const int gb = blockIdx.x;
  float *dst = Bop + (size_t)gb * (B_BYTES / 4);
  for (int e = threadIdx.x; e < TN * D; e += blockDim.x) {
    const int nn = e / D, j = e % D;
    const int g = nn / D, i = nn % D;
    const int k = gb * GB + g;
    const float w = k < ncols ? W[((size_t)k * D + i) * D + j] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const float lo = w - hi;
    // operand row n' = (i / 8) * 32 + g * 8 + i % 8: the outputs i >= 8 s that contraction step s (j in [8 s, 8 s + 8))
    // can reach -- W is lower triangular -- are then the contiguous rows [32 s, 256)
    const uint32_t off = core_off((uint32_t)((i >> 3) * (GB * 8) + g * 8 + (i & 7)), j, TN / 8) / 4;
    dst[off] = hi;
    dst[B_PART_BYTES / 4 + off] = lo;
  }
}

__global__ void __launch_bounds__(niwtc::THREADS, 1)
niw_tc_kernel(const float *__restrict__ X, const float *__restrict__ Bop, const float *__restrict__ bias,
              const float *__restrict__ coef, int ncols, float *__restrict__ scores, size_t ld, size_t row_lo,
              size_t row_hi, int num_gb_lanes, const float *__restrict__ base, int blocked) {
  // base != nullptr: this kernel is the only writer of the score matrix: scores = base[k] + term (no read);
  // base == nullptr: scores += term (the scalar-feature kernel has already written them)
  using namespace niwtc;
  extern __shared__ __align__(1024) unsigned char niw_smem[];
  unsigned char *sm = niw_smem;
  unsigned char *sB = sm;                      // [part][...]
  unsigned char *sA = sm + B_BYTES;            // [half][part][...]
  float *sBias = reinterpret_cast<float *>(sA + A_BYTES);            // [2][TN + 16]: bias then coef (GB x 4)
  uint64_t *bars = reinterpret_cast<uint64_t *>(sBias + 2 * (TN + 16));
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
  // barriers: 0 b_full, 1 b_free, 2-3 a_full[2], 4-5 a_empty[2], 6-7 acc_full[2], 8-9 acc_empty[2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t nrows = row_hi - row_lo;
  const int nRT = (int)((nrows + TM - 1) / TM);
  const int nGB = (ncols + GB - 1) / GB;
  // Schedule: the grid is G x P CTAs.  CTA (g, p) keeps group block g (then g + G, ...) resident and sweeps
  // the p-th slice of the row tiles, so the G CTAs of a slice read the same X tiles at about the same time:
  // X streams from HBM once per sweep and is shared through L2 instead of being re-read once per group block.
  // (lanes with a smaller index get the leftover CTAs: P_g = number of CTAs c with c % G == g.)
  const int G = (int)num_gb_lanes;
  const int g0 = (int)blockIdx.x % G, part = (int)blockIdx.x / G;
  const int P = ((int)gridDim.x - g0 + G - 1) / G;
  const int rt_lo = (int)((long long)nRT * part / P), rt_hi = (int)((long long)nRT * (part + 1) / P);

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(smem_u32(&bars[2 + i]), 4);   // a_full: one arrive per producer warp
      mbar_init(smem_u32(&bars[4 + i]), 1);   // a_empty: tcgen05.commit
      mbar_init(smem_u32(&bars[6 + i]), 1);   // acc_full: tcgen05.commit
      mbar_init(smem_u32(&bars[8 + i]), 4);   // acc_empty: one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {  // TMEM: all 512 columns (2 accumulators x 256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ===== producers: X tile (fp32) -> tf32 hi / lo halves in core-matrix layout =====
    // The A operand does not depend on the group block: the producers just cycle through this CTA's row
    // tiles.  Global loads for half-tile h + 2 are issued before waiting for the buffer of half-tile h, so
    // their latency hides behind the MMAs instead of serialising with them.
    const int n_mine = rt_hi - rt_lo;
    const int ngb_mine = (nGB - g0 + G - 1) / G;
    const long long total_h = 2ll * n_mine * ngb_mine;
    auto load_half = [&](long long h, float4 (&v)[8]) {
      const int rt = rt_lo + (int)((h >> 1) % n_mine), half = (int)(h & 1);
      const size_t row = row_lo + (size_t)rt * TM + tid;  // tid in [0,128): one row per thread
      const float4 *src = reinterpret_cast<const float4 *>(X + row * D) + half * 8;
      const bool ok = row < row_hi;
#pragma unroll
      for (int c = 0; c < 8; c++) v[c] = ok ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 cur[8], nxt[8], nx2[8];  // two half-tiles in flight: one (0.75 us of MMAs) does not cover the HBM latency under load
    if (total_h > 0) load_half(0, cur);
    if (total_h > 1) load_half(1, nxt);
    for (long long h = 0; h < total_h; h++) {
      if (h + 2 < total_h) load_half(h + 2, nx2);
      const int buf = (int)(h & 1);
      if (h >= 2) mbar_wait(smem_u32(&bars[4 + buf]), (uint32_t)(((h >> 1) - 1) & 1));  // MMAs that read this buffer are done
      unsigned char *hi_base = sA + (size_t)buf * 2 * A_HALF_BYTES;
      unsigned char *lo_base = hi_base + A_HALF_BYTES;
#pragma unroll
      for (int c = 0; c < 8; c++) {  // 8 chunks of 4 k
        const float4 v = cur[c];
        float4 vh, vl;
        vh.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); vl.x = v.x - vh.x;
        vh.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); vl.y = v.y - vh.y;
        vh.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); vl.z = v.z - vh.z;
        vh.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); vl.w = v.w - vh.w;
        const uint32_t off = core_off((uint32_t)tid, (uint32_t)c * 4, TM / 8);
        *reinterpret_cast<float4 *>(hi_base + off) = vh;
        *reinterpret_cast<float4 *>(lo_base + off) = vl;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[2 + buf]));
#pragma unroll
      for (int c = 0; c < 8; c++) { cur[c] = nxt[c]; nxt[c] = nx2[c]; }
    }
  } else if (warp == 8) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      long long h = 0, t = 0;
      int cur_gb = -1;
      uint32_t b_loads = 0;
      for (int gb = g0; gb < nGB; gb += G)
      for (int rt = rt_lo; rt < rt_hi; rt++, t++) {
        if (gb != cur_gb) {  // (re)load the resident B operand
          if (cur_gb >= 0) {
            mma_commit(smem_u32(&bars[1]));
            mbar_wait(smem_u32(&bars[1]), (b_loads - 1) & 1u);  // every MMA that read the old B has completed
          }
          mbar_expect_tx(smem_u32(&bars[0]), B_BYTES);
          const unsigned char *gsrc = reinterpret_cast<const unsigned char *>(Bop) + (size_t)gb * B_BYTES;
          for (int c = 0; c < 4; c++)
            bulk_g2s(smem_u32(sB + (size_t)c * (B_BYTES / 4)), gsrc + (size_t)c * (B_BYTES / 4), B_BYTES / 4, smem_u32(&bars[0]));
          mbar_wait(smem_u32(&bars[0]), b_loads & 1u);
          b_loads++;
          cur_gb = gb;
        }
        const int acc = (int)(t & 1);
        if (t >= 2) mbar_wait(smem_u32(&bars[8 + acc]), (uint32_t)(((t >> 1) - 1) & 1));  // epilogue drained this accumulator
        const uint32_t d_tmem = tmem + (uint32_t)acc * TN;
        for (int half = 0; half < 2; half++, h++) {
          const int buf = (int)(h & 1);
          mbar_wait(smem_u32(&bars[2 + buf]), (uint32_t)((h >> 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem_u32(sA + (size_t)buf * 2 * A_HALF_BYTES), a_lo = a_hi + A_HALF_BYTES;
          const uint32_t b_hi = smem_u32(sB) + (uint32_t)half * (32 / 4) * (TN / 8) * 128u, b_lo = b_hi + B_PART_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {  // K = 8 per instruction: 2 core-matrix columns
            // W is lower triangular: contraction step s = 4 half + ks (j in [8 s, 8 s + 8)) only reaches the outputs
            // i >= 8 s, i.e. the operand rows / accumulator columns [32 s, 256): N shrinks by 32 per step
            // (sum over the 8 steps: 1152 instead of 2048 columns of tensor work)
            const uint32_t st8 = (uint32_t)(half * 4 + ks);
            const uint32_t n0 = (blocked & 4) ? st8 * (GB * 8) : (blocked & 2) ? (uint32_t)half * (TN / 2) : 0u;
            const uint32_t idesc = idesc_n(TN - n0);
            const uint32_t ao = (uint32_t)ks * 2u * (TM / 8) * 128u;
            const uint32_t bo = (uint32_t)ks * 2u * (TN / 8) * 128u + (n0 / 8u) * 128u;
            const uint64_t dah = smem_desc(a_hi + ao, (TM / 8) * 128u, 128u), dal = smem_desc(a_lo + ao, (TM / 8) * 128u, 128u);
            const uint64_t dbh = smem_desc(b_hi + bo, (TN / 8) * 128u, 128u), dbl = smem_desc(b_lo + bo, (TN / 8) * 128u, 128u);
            mma_tf32(d_tmem + n0, dah, dbh, idesc, (half | ks) ? 1u : 0u);
            mma_tf32(d_tmem + n0, dah, dbl, idesc, 1u);
            mma_tf32(d_tmem + n0, dal, dbh, idesc, 1u);
          }
          mma_commit(smem_u32(&bars[4 + buf]));            // A half-buffer free once these MMAs complete
          if (half == 1) mma_commit(smem_u32(&bars[6 + acc]));  // accumulator ready for the epilogue
        }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 4..7: TMEM lanes 32*(warp%4) .. +31 =====
    const int ew = warp & 3;
    const int etid = tid - 128;  // 0..127
    long long t = 0;
    for (int gb = g0; gb < nGB; gb += G)
    for (int rt = rt_lo; rt < rt_hi; rt++, t++) {
      const int acc = (int)(t & 1);
      float *sb = sBias + (size_t)acc * (TN + 16);
      // stage this block's bias and coefficients (the previous use of this buffer ended two tiles ago)
      for (int i = etid; i < TN + 16; i += 128) {
        float v = 0.f;
        if (i < TN) { const int k = gb * GB + i / D; if (k < ncols) v = bias[(size_t)k * D + (i % D)]; }
        else { const int k = gb * GB + (i - TN) / 4; if (k < ncols) v = coef[(size_t)k * 4 + ((i - TN) & 3)]; }
        sb[i] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // old scores of this thread's row (4 consecutive groups = one 16-byte access), fetched BEFORE waiting for
      // the accumulator so that the global-load latency hides behind the MMAs
      const size_t row = row_lo + (size_t)rt * TM + ew * 32 + lane;
      float4 *dst = reinterpret_cast<float4 *>(scores + (row - row_lo) * ld + (size_t)gb * GB);
      // blocked layout (the sweep's, msb_score.cuh): this warp's 32 rows are one 32-row block, lane = row, so the
      // GB values of a thread go to GB different 128-byte lines, each written by the whole warp at once
      float *dstb = scores + (((size_t)rt * (TM / 32) + ew) * ld + (size_t)gb * GB) * 32 + lane;
      float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < row_hi) {
        if (base) old = __ldg(reinterpret_cast<const float4 *>(base + (size_t)gb * GB));
        else if (blocked & 1) old = make_float4(dstb[0], dstb[32], dstb[64], dstb[96]);
        else old = *dst;
      }
      mbar_wait(smem_u32(&bars[6 + acc]), (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // accumulator column n' = (i / 8) * 32 + g * 8 + i % 8 (niw_pack_b_kernel): one 32-column load holds the
      // outputs i in [8 ib, 8 ib + 8) of all four groups
      // The load of the next 32 columns is in flight while the current ones are reduced; two outputs per
      // instruction (FADD2 / FFMA2), even and odd outputs in separate partial sums.
      float q[GB];
      {
        float2 q2[GB];
#pragma unroll
        for (int g = 0; g < GB; g++) q2[g] = make_float2(0.f, 0.f);
        const uint32_t tbase = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * TN);
        uint32_t r[2][32];
        tmem_ld32_issue(tbase, r[0]);
#pragma unroll
        for (int ib = 0; ib < D / 8; ib++) {
          tmem_ld_wait();
          if (ib + 1 < D / 8) tmem_ld32_issue(tbase + (uint32_t)(ib + 1) * 32u, r[(ib + 1) & 1]);
#pragma unroll
          for (int g = 0; g < GB; g++) {
            const float4 b0 = *reinterpret_cast<const float4 *>(sb + g * D + ib * 8);
            const float4 b1 = *reinterpret_cast<const float4 *>(sb + g * D + ib * 8 + 4);
            const uint32_t *v = r[ib & 1] + g * 8;
            float2 y;
            y = __fadd2_rn(make_float2(__uint_as_float(v[0]), __uint_as_float(v[1])), make_float2(-b0.x, -b0.y)); q2[g] = __ffma2_rn(y, y, q2[g]);
            y = __fadd2_rn(make_float2(__uint_as_float(v[2]), __uint_as_float(v[3])), make_float2(-b0.z, -b0.w)); q2[g] = __ffma2_rn(y, y, q2[g]);
            y = __fadd2_rn(make_float2(__uint_as_float(v[4]), __uint_as_float(v[5])), make_float2(-b1.x, -b1.y)); q2[g] = __ffma2_rn(y, y, q2[g]);
            y = __fadd2_rn(make_float2(__uint_as_float(v[6]), __uint_as_float(v[7])), make_float2(-b1.z, -b1.w)); q2[g] = __ffma2_rn(y, y, q2[g]);
          }
        }
#pragma unroll
        for (int g = 0; g < GB; g++) q[g] = q2[g].x + q2[g].y;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[8 + acc]));
      if (row < row_hi) {
        float o[GB] = {old.x, old.y, old.z, old.w};
#pragma unroll
        for (int g = 0; g < GB; g++) {
          const int k = gb * GB + g;
          if (k < ncols && q[g] == q[g]) {  // NaN = masked row: contributes nothing
            const float c0 = sb[TN + g * 4 + 0], c1 = sb[TN + g * 4 + 1], idof = sb[TN + g * 4 + 2];
            o[g] += c0 + c1 * log1pf(q[g] * idof);
          }
        }
        if (blocked & 1) {
#pragma unroll
          for (int g = 0; g < GB; g++) dstb[g * 32] = o[g];
        } else {
          *dst = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // nobody still reads sb when the next-but-one tile restages it
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static inline int niw_tc_init(size_t smem_optin, std::string &err) {
  if (niwtc::SMEM_BYTES > smem_optin) return MSB_OK;  // path disabled, the CUDA-core kernel is used
  cudaError_t e = cudaFuncSetAttribute(niw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)niwtc::SMEM_BYTES);
  if (e != cudaSuccess) { err = std::string("niw_tc_init: ") + cudaGetErrorString(e); return MSB_ERR_CUDA; }
  return MSB_OK;
}
static inline size_t niw_tc_operand_bytes(size_t ncols, uint32_t dim) {
  if (dim != niwtc::D) return 16;
  return ((ncols + niwtc::GB - 1) / niwtc::GB) * (size_t)niwtc::B_BYTES;
}
// scores[row][k] += NIW term for rows [row_lo,row_hi), all ncols groups.  *done = false when this path does not apply.
static inline int niw_tc_score(cudaStream_t stream, uint64_t *launches, const float *X, uint32_t dim, const float *W,
                               const float *bias, const float *coef, float *Bop, size_t ncols, float *scores,
                               size_t ld, size_t row_lo, size_t row_hi, int sm_count, const float *base, bool blocked,
                               bool *done, std::string &err) {
  *done = false;
  if (dim != niwtc::D || getenv("MSB_NO_TENSOR") || row_hi <= row_lo) return MSB_OK;
  const int nGB = (int)((ncols + niwtc::GB - 1) / niwtc::GB);
  niw_pack_b_kernel<<<nGB, 256, 0, stream>>>(W, (int)ncols, Bop);
  (*launches)++;
  const long long nRT = (long long)((row_hi - row_lo + niwtc::TM - 1) / niwtc::TM);
  const int G = std::min(nGB, sm_count);
  // as many CTAs as SMs, but never more than one per (group block, row tile)
  const int grid = (int)std::max<long long>(G, std::min<long long>(sm_count, (long long)G * nRT));
  niw_tc_kernel<<<grid, niwtc::THREADS, niwtc::SMEM_BYTES, stream>>>(X, Bop, bias, coef, (int)ncols, scores, ld, row_lo,
                                                                      row_hi, G, base, (blocked ? 1 : 0) | (getenv("MSB_NIW_TRI2") ? 2 : 0) | (getenv("MSB_NIW_TRI8") ? 4 : 0));
  (*launches)++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { err = std::string("niw_tc_kernel launch: ") + cudaGetErrorString(e); return MSB_ERR_CUDA; }
  *done = true;
  return MSB_OK;
}

}  // namespace msb
