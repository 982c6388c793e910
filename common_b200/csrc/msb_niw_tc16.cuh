// msb_niw_tc16.cuh -- the NIW Mahalanobis GEMM on the fp16 tensor-core path (tcgen05 kind::f16, M128 N256 K16), dim == 64.
//
//   scores[n][k] += c0_k + c1_k * log1p(|W_k x_n - b_k|^2 / dof_k),     Y[n][(k,i)] = sum_j X[n][j] W_k[i][j]
//
// Operands.  A tf32 operand keeps 11 significand bits, so fp32-class accuracy takes a hi / lo split and three products
// per k-step (hi*hi + hi*lo + lo*hi; msb_niw_tc.cuh).  An fp16 operand keeps 11 bits as well: the same three products
// give the same 2^-21 at twice the tensor rate, K = 16 per instruction and half the operand bytes.  fp16's narrow
// exponent is handled by exact power-of-two scales:
//   X'[n][j]      = X[n][j] sx_j,            sx_j   = 2^(9 - e),  max_n |X[n][j]|        in [2^(e-1), 2^e)
//   W'[k][i][j]   = W[k][i][j] r_k / sx_j,   r_k    = 2^(9 - e),  max_ij |W[k][i][j]/sx_j| in [2^(e-1), 2^e)
//   Y'[n][(k,i)]  = sum_j X' W' = r_k Y,     epilogue: t = Y' - r_k b, q' = sum_i t^2 = r_k^2 q  (1 / r_k^2 folded into 1 / dof)
// so every column of X' and every W'_k has its largest entry in [256, 512): products accumulate (fp32, TMEM) below 2^24, the
// hi parts are normal down to 2^-23 of the largest entry and what the subnormal lo parts lose is below 2^-30 of the
// largest term of the sum.  The scales are recomputed from the rows of the sweep itself (niw_colmax_kernel: one more
// pass over X, 256 MB at C4), so data uploaded after bind cannot overflow them.
//
// Kernels of one call:  niw_colmax_kernel -> niw_convert_a16_kernel (A: hi | lo parts of every 128-row tile, in the exact
// core-matrix bytes of a ring buffer, written once per sweep) + niw_pack_b16_kernel (B: the same for every block of 4
// groups, the structurally zero rows of the triangular W_k left out, + the bias operand) -> niw_tc16_kernel:
//   * warp 0 streams A half-tiles (16 KB) into a 6-deep shared-memory ring with bulk copies and keeps B (48 KB per
//     group block) double-buffered one work item ahead;
//   * warps 1 and 18 issue the MMAs, even and odd tiles: per 128-row tile one bias product (ones x -b': the accumulator
//     starts at -b') and four k-steps x three products, N shrinking 256, 192, 128, 64, into the issuer's own 256-column
//     TMEM accumulator.  Their loops are warp-uniform (descriptors in uniform registers, elect.sync around the issue);
//   * sixteen epilogue warps, all on every tile: (TMEM lane quadrant) x (group of the block); four tcgen05.ld
//     32x32b.x16 in flight, the accumulator released before the arithmetic, square sum, log1p, store.  In the sweep's
//     blocked layout a warp's 32 lanes are one 32-row block, so every store is a full 128-byte line;
//   * work items = (slice of <= 32 row tiles) x (group block), dealt round-robin to the CTAs, group block fastest: even
//     load and A shared through L2.
// Synchronisation is mbarrier-only (bulk-copy complete_tx, tcgen05.commit); every wait is bounded (trap).
// What paced it, in the order found (C4, ms per score call): the issuer's instruction stream (4.05 -> 3.29), the bias
// read from shared memory (-> 2.92), barrier waits serialised with the issue (two issuers: -> 2.90 at an SM clock the
// power cap has by then pulled from 1965 to ~1665 MHz: 3 tensor products per algorithmic one).
#pragma once
#include <cuda_fp16.h>

#include "msb_niw_tc.cuh"

namespace msb {

namespace niwtc16 {
using niwtc::D;
using niwtc::GB;
using niwtc::TM;
using niwtc::TN;
constexpr int EPI_WARPS = 16;               // epilogue warps: 4 groups of the block x 4 TMEM lane quadrants
constexpr int THREADS = (3 + EPI_WARPS) * 32;   // warp 0 A / B producer, warps 1 and 18 MMA issuers (even / odd tiles), warps 2-17 epilogue
constexpr int A_HALF_BYTES = TM * 32 * 2;   // one k-half (32 k) of one part (hi or lo): 8 KB
constexpr int NA = 6;                        // A half-tile buffers in the ring (hi + lo each): the copies run three tiles ahead of the MMAs
constexpr int A_BYTES = NA * 2 * A_HALF_BYTES;   // [buffer][part]: 96 KB
// B operand of one group block, compact: W_k is lower triangular, so contraction step st (j in [16 st, 16 st + 16)) only
// reaches the operand rows n' >= 64 st (see niw_pack_b16_kernel) and only those rows are stored: 256, 192, 128, 64 rows of
// 16 k each = 20 KB per part instead of 32
__device__ __host__ constexpr uint32_t b_step_off(uint32_t st) { return st * 8192u - (st * (st - 1u) / 2u) * 2048u; }  // sum_{s<st} (256 - 64 s) * 32
constexpr int B_PART_BYTES = 20480;         // b_step_off(4)
constexpr int BIAS_BYTES = 2 * (TN / 8) * 128;   // the bias operand: TN rows x 16 k (k = 0 hi, k = 1 lo, the rest zero): 8 KB
constexpr int B_BYTES = 2 * B_PART_BYTES + BIAS_BYTES;   // hi + lo + bias: 48 KB
constexpr int NB = 2;                        // B buffers: the next item's group block lands while the current one is multiplied
constexpr int ONES_BYTES = 2 * (TM / 8) * 128;   // the A operand of the bias product: TM rows x 16 k, columns 0 and 1 = 512: 4 KB
constexpr float BIAS_SCALE = 512.f;          // ones column = 512, bias operand = -b' / 512: |b'| < 2^24 stays inside fp16's range
constexpr size_t SMEM_BYTES = (size_t)NB * B_BYTES + A_BYTES + ONES_BYTES + 24 * sizeof(uint64_t) + 64;

// byte offset of element (row, k) inside one part of an fp16 operand tile with `rowgroups` 8-row groups:
// [k/8][row/8][row%8][k%8]  (core matrix = 8 rows x 16 bytes, contiguous)
__device__ __host__ __forceinline__ uint32_t core_off16(uint32_t row, uint32_t k, uint32_t rowgroups) {
  return (k >> 3) * (rowgroups * 128u) + (row >> 3) * 128u + (row & 7u) * 16u + (k & 7u) * 2u;
}
// instruction descriptor, kind::f16: D = F32 (bits 4-5 = 1), A = B = F16 (bits 7-9, 10-12 = 0), K-major both,
// N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ constexpr uint32_t idesc_f16(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the power of two that brings a positive maximum into [256, 512); 1 for an all-zero (or non-finite) row
__device__ __forceinline__ float pow2_scale(float mx) {
  if (!(mx > 0.f) || mx > 3.0e38f) return 1.f;
  int e;
  frexpf(mx, &e);  // mx = m 2^e, m in [0.5, 1)
  return ldexpf(1.f, 9 - e);
}
// x (already scaled) -> fp16 hi and lo parts, two values at a time
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}
// 16 TMEM lanes x 256 bits x 2: thread T gets rows T / 4 (r0 r1, r4 r5) and T / 4 + 8 (r2 r3, r6 r7) of the 16 lanes that
// start at the address's lane, columns 2 (T % 4) + {0, 1} (r0-r3) and 8 + 2 (T % 4) + {0, 1} (r4-r7)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
}  // namespace niwtc16

// colmax[j] = max over the rows [row_lo, row_hi) of |X[row][j]| as float bits (non-negative floats order like
// unsigned integers; NaN rows -- masked -- and infinities are skipped).  colmax must be zeroed before.
__global__ void niw_colmax_kernel(const float *__restrict__ X, size_t row_lo, size_t row_hi, unsigned int *__restrict__ colmax) {
  constexpr int D = niwtc::D;
  __shared__ unsigned int sm[D];
  if (threadIdx.x < D) sm[threadIdx.x] = 0u;
  __syncthreads();
  const int j4 = (threadIdx.x & 15) * 4;  // 16 threads cover one row with float4 loads
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t row = row_lo + (size_t)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); row < row_hi;
       row += (size_t)gridDim.x * (blockDim.x >> 4)) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(X + row * D + j4));
    const float a = fabsf(v.x), b = fabsf(v.y), c = fabsf(v.z), d = fabsf(v.w);
    if (a <= 3.0e38f) m.x = fmaxf(m.x, a);
    if (b <= 3.0e38f) m.y = fmaxf(m.y, b);
    if (c <= 3.0e38f) m.z = fmaxf(m.z, c);
    if (d <= 3.0e38f) m.w = fmaxf(m.w, d);
  }
  atomicMax(&sm[j4 + 0], __float_as_uint(m.x));
  atomicMax(&sm[j4 + 1], __float_as_uint(m.y));
  atomicMax(&sm[j4 + 2], __float_as_uint(m.z));
  atomicMax(&sm[j4 + 3], __float_as_uint(m.w));
  __syncthreads();
  if (threadIdx.x < D) atomicMax(&colmax[threadIdx.x], sm[threadIdx.x]);
}

// X rows [row_lo, row_hi) -> the A operand of niw_tc16_kernel: per 128-row tile and k-half, [hi | lo] fp16 parts of
// X[row][j] sx_j in core-matrix layout (16 KB, the exact bytes of one ring buffer).  Rows past row_hi are zero.
// thread = (row, 8-k chunk), rows fastest: 32-byte reads (one sector each), coalesced 16-byte writes.
__global__ void niw_convert_a16_kernel(const float *__restrict__ X, size_t row_lo, size_t row_hi, const unsigned int *__restrict__ colmax,
                                       unsigned char *__restrict__ A16) {
  using namespace niwtc16;
  __shared__ float s_sx[D];
  if (threadIdx.x < D) s_sx[threadIdx.x] = pow2_scale(__uint_as_float(colmax[threadIdx.x]));
  __syncthreads();
  const size_t rt = blockIdx.x;                 // one block per tile: 128 rows x 8 chunks = 1024 items, 256 threads
  for (int it = threadIdx.x; it < TM * 8; it += blockDim.x) {
    const int r = it % TM, c8 = it / TM;        // c8 in [0, 8): k = 8 c8 .. 8 c8 + 7
    const size_t row = row_lo + rt * TM + r;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row < row_hi) {
      a = __ldg(reinterpret_cast<const float4 *>(X + row * D + c8 * 8));
      b = __ldg(reinterpret_cast<const float4 *>(X + row * D + c8 * 8 + 4));
    }
    const float *sx = s_sx + c8 * 8;
    uint4 vh, vl;
    split2(a.x * sx[0], a.y * sx[1], vh.x, vl.x);
    split2(a.z * sx[2], a.w * sx[3], vh.y, vl.y);
    split2(b.x * sx[4], b.y * sx[5], vh.z, vl.z);
    split2(b.z * sx[6], b.w * sx[7], vh.w, vl.w);
    const int half = c8 >> 2;
    unsigned char *hi_base = A16 + (rt * 2 + half) * (size_t)(2 * A_HALF_BYTES);
    const uint32_t off = core_off16((uint32_t)r, (uint32_t)(c8 & 3) * 8, TM / 8);
    *reinterpret_cast<uint4 *>(hi_base + off) = vh;
    *reinterpret_cast<uint4 *>(hi_base + A_HALF_BYTES + off) = vl;
  }
}

// W[k][i][j] (fp32, row-major per group) -> per group block: [hi | lo] fp16 parts in UMMA layout (compact, see
// b_step_off), scaled, + the bias operand; blocks padded with zeros.  rinv[2 k] = r_k, rinv[2 k + 1] = 1 / r_k^2.
// sx[j] is derived from colmax[j] (as niw_convert_a16_kernel does).
__global__ void niw_pack_b16_kernel(const float *__restrict__ W, const float *__restrict__ bias, int ncols,
                                    const unsigned int *__restrict__ colmax, unsigned char *__restrict__ Bop, float *__restrict__ rinv) {
  using namespace niwtc16;
  __shared__ float s_isx[D];   // 1 / sx_j
  __shared__ float s_r[2 * GB];  // per group: max |W / sx| (as ordered bits), then r_k
  const int gb = blockIdx.x;
  if (threadIdx.x < D) {
    const float sx = pow2_scale(__uint_as_float(colmax[threadIdx.x]));
    s_isx[threadIdx.x] = 1.f / sx;
  }
  __syncthreads();
  // one scale per group: r_k brings max_ij |W[k][i][j] / sx_j| into [256, 512)
  if (threadIdx.x < GB) s_r[threadIdx.x] = 0.f;
  __syncthreads();
  for (int nn = threadIdx.x; nn < TN; nn += blockDim.x) {  // nn = g * D + i
    const int g = nn / D, i = nn % D, k = gb * GB + g;
    float mx = 0.f;
    if (k < ncols)
      for (int j = 0; j < D; j++) mx = fmaxf(mx, fabsf(W[((size_t)k * D + i) * D + j] * s_isx[j]));
    atomicMax(reinterpret_cast<unsigned int *>(&s_r[g]), __float_as_uint(mx));
  }
  __syncthreads();
  if (threadIdx.x < GB) {
    const float r = pow2_scale(s_r[threadIdx.x]);
    rinv[((size_t)gb * GB + threadIdx.x) * 2 + 0] = r;
    rinv[((size_t)gb * GB + threadIdx.x) * 2 + 1] = (1.f / r) * (1.f / r);
    s_r[GB + threadIdx.x] = r;
  }
  __syncthreads();
  __half *dst = reinterpret_cast<__half *>(Bop + (size_t)gb * B_BYTES);
  for (int e = threadIdx.x; e < TN * D; e += blockDim.x) {
    const int nn = e / D, j = e % D;
    const int g = nn / D, i = nn % D;
    const int k = gb * GB + g;
    // operand row n' = (i / 16) * 64 + g * 16 + i % 16: contraction step st = j / 16 of the lower-triangular W_k reaches
    // the rows (= accumulator columns) n' >= 64 st only, and one group's 16 outputs of a step are 16 contiguous columns
    const uint32_t st = (uint32_t)j >> 4, np = (uint32_t)((i >> 4) * (GB * 16) + g * 16 + (i & 15));
    if (np < 64u * st) continue;  // structurally zero (i < 16 st <= j): not stored
    const float w = k < ncols ? W[((size_t)k * D + i) * D + j] * s_isx[j] * s_r[GB + g] : 0.f;   // exact: powers of two
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const uint32_t off = (b_step_off(st) + core_off16(np - 64u * st, (uint32_t)j & 15u, (TN - 64u * st) / 8)) / 2;
    dst[off] = hi;
    dst[B_PART_BYTES / 2 + off] = lo;
  }
  // the bias operand: row n', k = 0: hi part of -b' / 512, k = 1: its lo part (b' = b r_k, the scale of Y'); with the
  // ones operand (columns 0 and 1 = 512) one more MMA starts every accumulator at -b'
  __half *bd = reinterpret_cast<__half *>(Bop + (size_t)gb * B_BYTES + 2 * B_PART_BYTES);
  for (int e = threadIdx.x; e < TN * 16; e += blockDim.x) {
    const int nn = e / 16, kk = e % 16;
    const int g = nn / D, i = nn % D, k = gb * GB + g;
    const uint32_t np = (uint32_t)((i >> 4) * (GB * 16) + g * 16 + (i & 15));
    __half v = __float2half_rn(0.f);
    if (kk < 2 && k < ncols) {
      const float b = -bias[(size_t)k * D + i] * s_r[GB + g] * (1.f / BIAS_SCALE);
      const __half hi = __float2half_rn(b);
      v = kk == 0 ? hi : __float2half_rn(b - __half2float(hi));
    }
    bd[core_off16(np, (uint32_t)kk, TN / 8) / 2] = v;
  }
}

#ifdef MSB_NIW_TRACE
__device__ long long niw_trace[8192];
#define NIW_TRACE(slot) do { if (blockIdx.x == 0 && t >= 8 && t < 40 && lane == 0) niw_trace[(t - 8) * 8 + (slot)] = clock64(); } while (0)
#else
#define NIW_TRACE(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(niwtc16::THREADS, 1)
niw_tc16_kernel(const unsigned char *__restrict__ A16, const unsigned char *__restrict__ Bop, const float *__restrict__ rinv,
                const float *__restrict__ bias, const float *__restrict__ coef, int ncols,
                float *__restrict__ scores, size_t ld, size_t row_lo, size_t row_hi, int slice_tiles,
                const float *__restrict__ base, int blocked) {
  using namespace niwtc16;
  extern __shared__ __align__(1024) unsigned char niw_smem[];
  unsigned char *sm = niw_smem;
  unsigned char *sB = sm;                      // [buffer][part][...]
  unsigned char *sA = sm + NB * B_BYTES;       // [ring slot][part][...]
  unsigned char *sOnes = sA + A_BYTES;         // the A operand of the bias product
  uint64_t *bars = reinterpret_cast<uint64_t *>(sOnes + ONES_BYTES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 23);
  // barriers: b_full[2], b_free[2], a_full[NA], a_empty[NA], acc_full[2], acc_empty[2]
  constexpr int B_FULL = 0, B_FREE = 2, A_FULL = 4, A_EMPTY = 4 + NA, ACC_FULL = 4 + 2 * NA, ACC_EMPTY = 6 + 2 * NA;
  static_assert(ACC_EMPTY + 2 <= 23, "barrier slots");
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform as far as the compiler is concerned: uniform branches and registers
  const size_t nrows = row_hi - row_lo;
  const int nRT = (int)((nrows + TM - 1) / TM);
  const int nGB = (ncols + GB - 1) / GB;
  // Schedule: the work is cut into items (slice of `slice_tiles` row tiles) x (group block), numbered group block
  // fastest, and CTA c takes the items c, c + gridDim.x, ...  At any moment the CTAs are therefore all inside two or three
  // neighbouring slices: their A tiles are shared through L2 (A streams from HBM about once), and every CTA gets the
  // same number of items to within one.
  const int SL = slice_tiles;
  const int nSL = (nRT + SL - 1) / SL;
  const long long nItems = (long long)nSL * nGB;
  if (tid == 0) {
    for (int i = 0; i < NB; i++) {
      mbar_init(smem_u32(&bars[B_FULL + i]), 1);    // b_full: the producer's expect_tx, completed by the bulk copies
      mbar_init(smem_u32(&bars[B_FREE + i]), 2);    // b_free: one tcgen05.commit per issuer after its last MMA of the item
    }
    for (int i = 0; i < NA; i++) {
      mbar_init(smem_u32(&bars[A_FULL + i]), 1);    // a_full: the producer's expect_tx, completed by the bulk copy
      mbar_init(smem_u32(&bars[A_EMPTY + i]), 1);   // a_empty: tcgen05.commit
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(smem_u32(&bars[ACC_FULL + i]), 1);  // acc_full: tcgen05.commit
      mbar_init(smem_u32(&bars[ACC_EMPTY + i]), EPI_WARPS); // acc_empty: one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ones operand: row r, k = 0 and 1 -> 512, everything else 0 (generic-proxy writes, read by the tensor core)
  for (int i = tid; i < ONES_BYTES / 4; i += THREADS) {
    const uint32_t byte = (uint32_t)i * 4u;   // [k/8][row/8][row%8][k%8]: the first 16-byte core row holds k = 0..7
    reinterpret_cast<uint32_t *>(sOnes)[i] = (byte < (TM / 8) * 128u && (byte & 15u) == 0u) ? 0x60006000u : 0u;   // fp16 512 = 0x6000
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {  // TMEM: all 512 columns (2 accumulators x 256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== producer: streams the pre-converted A operand (niw_convert_a16_kernel: scaled fp16 hi / lo parts, already in
    // core-matrix layout, 16 KB per half-tile) into the ring with bulk copies, and keeps the B operand (one block of 4
    // groups, 64 KB) one item ahead: the next item's block is requested a few tiles into the current item -- by then the
    // MMAs of the previous item, which read the buffer it goes into, have long completed, so the wait below does not hold
    // the A ring up -- and lands under ~30 tiles of tensor work.  All 32 lanes walk the (warp-uniform) loop and wait;
    // one elected lane issues the copies.
    auto load_b = [&](int gb, int buf) {
      const uint32_t bar = smem_u32(&bars[B_FULL + buf]);
      mbar_expect_tx(bar, B_BYTES);
      const unsigned char *gsrc = Bop + (size_t)gb * B_BYTES;
#pragma unroll
      for (int c = 0; c < 4; c++)
        bulk_g2s(smem_u32(sB + (size_t)buf * B_BYTES + (size_t)c * (B_BYTES / 4)), gsrc + (size_t)c * (B_BYTES / 4), B_BYTES / 4, bar);
    };
    uint32_t abuf = 0, around = 0;   // ring slot and how many times the ring has wrapped
    int it = 0;
    for (long long item = blockIdx.x; item < nItems; item += gridDim.x, it++) {
      const int sl = (int)(item / nGB);
      const int rt_lo = sl * SL, rt_hi = rt_lo + SL < nRT ? rt_lo + SL : nRT;
      if (it == 0 && elect_one()) load_b((int)(item % nGB), 0);
      const long long next = item + gridDim.x;
      const int trig = rt_lo + 4 < rt_hi - 1 ? rt_lo + 4 : rt_hi - 1;
      for (int rt = rt_lo; rt < rt_hi; rt++) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
          if (around) mbar_wait(smem_u32(&bars[A_EMPTY + abuf]), (around - 1) & 1u);  // the MMAs that read this slot are done
          if (elect_one()) {
            const uint32_t bar = smem_u32(&bars[A_FULL + abuf]);
            mbar_expect_tx(bar, 2 * A_HALF_BYTES);
            bulk_g2s(smem_u32(sA) + abuf * (uint32_t)(2 * A_HALF_BYTES), A16 + ((size_t)rt * 2 + half) * (2 * A_HALF_BYTES), 2 * A_HALF_BYTES, bar);
          }
          __syncwarp();
          if (++abuf == (uint32_t)NA) { abuf = 0; around++; }
        }
        if (rt == trig && next < nItems) {
          const int nb = (it + 1) & 1, u = (it + 1) >> 1;   // use u of buffer nb
          if (u >= 1) mbar_wait(smem_u32(&bars[B_FREE + nb]), (uint32_t)((u - 1) & 1));  // the MMAs of item it - 1 are done
          if (elect_one()) load_b((int)(next % nGB), nb);
          __syncwarp();
        }
      }
    }
  } else if (warp == 1 || warp == 2 + EPI_WARPS) {
    // ===== MMA issuers: warp 1 takes the even tiles (accumulator 0), warp 18 the odd ones (accumulator 1).  All 32 lanes
    // walk the loop (it is warp-uniform: the warp index comes from a shuffle, so the compiler keeps the addresses and
    // descriptors in uniform registers) and wait on the barriers; one elected lane issues the MMAs and commits.
    // Why two: clock64 stamps of a single issuer (scripts/niw_trace.py) show ~1030 of a tile's ~1660 cycles inside the 13
    // tcgen05.mma issues (~80 cycles each: the issue blocks until the tensor pipe takes the instruction, i.e. the queue
    // is about one instruction deep) and ~630 cycles in three satisfied barrier waits, fences and the elect -- during
    // which the pipe runs dry.  With the tiles dealt alternately to two issuers, one waits while the other feeds the pipe.
    // (Before that, one lane under a divergent branch spent ~300 instructions per half-tile on 64-bit ring arithmetic,
    // descriptor assembly and R2UR moves: ~1900 cycles per tile.)
    const int par = warp == 1 ? 0 : 1;
    uint32_t abuf = 0, aphase = 0;
    auto advance = [&]() { if (++abuf == (uint32_t)NA) { abuf = 0; aphase ^= 1u; } };
    int t = 0, it = 0;
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB), sOnes0 = smem_u32(sOnes);
    const uint32_t d_tmem = tmem + (uint32_t)par * TN;
    for (long long item = blockIdx.x; item < nItems; item += gridDim.x, it++) {
      const int sl = (int)(item / nGB);
      const int rt_lo = sl * SL, rt_hi = rt_lo + SL < nRT ? rt_lo + SL : nRT;
      const int bbuf = it & 1;
      mbar_wait(smem_u32(&bars[B_FULL + bbuf]), (uint32_t)((it >> 1) & 1));
      const uint32_t sB_cur = sB0 + (uint32_t)bbuf * B_BYTES;
      for (int rt = rt_lo; rt < rt_hi; rt++, t++) {
        if ((t & 1) != par) { advance(); advance(); continue; }   // the other issuer's tile
        NIW_TRACE(0);
        if (t >= 2) mbar_wait(smem_u32(&bars[ACC_EMPTY + par]), (uint32_t)(((t >> 1) - 1) & 1));  // the epilogue has read this accumulator
        NIW_TRACE(1);
#pragma unroll
        for (int half = 0; half < 2; half++) {
          mbar_wait(smem_u32(&bars[A_FULL + abuf]), aphase);
          NIW_TRACE(2 + half * 2);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = sA0 + abuf * (uint32_t)(2 * A_HALF_BYTES), a_lo = a_hi + A_HALF_BYTES;
          const uint32_t b_hi = sB_cur, b_lo = b_hi + B_PART_BYTES;
          if (elect_one()) {
            if (half == 0) {
              // the accumulator starts at -b' (the group's bias in the scale of Y'): ones operand (columns 0, 1 = 512) x
              // bias operand (k = 0: hi, k = 1: lo part of -b' / 512).  One more K = 16 step at full N (128 of ~1090 tensor
              // cycles per tile) instead of a subtraction per element in the epilogue.
              mma_f16(d_tmem, niwtc::smem_desc(sOnes0, (TM / 8) * 128u, 128u), niwtc::smem_desc(sB_cur + 2 * B_PART_BYTES, (TN / 8) * 128u, 128u),
                      idesc_f16(TN), 0u);
            }
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {  // K = 16 per instruction: 2 core-matrix columns
              // W_k is lower triangular: contraction step st (j in [16 st, 16 st + 16)) only reaches the outputs i >= 16 st,
              // which the operand row order n' = (i / 16) * 64 + g * 16 + i % 16 makes the contiguous rows / accumulator
              // columns [64 st, 256): N = 256, 192, 128, 64 over the four steps, 640 instead of 1024 columns of tensor
              // work per tile, and B keeps those rows only (b_step_off).
              const uint32_t st = (uint32_t)(half * 2 + ks);
              const uint32_t n0 = st * 64u, nn = TN - n0;
              const uint32_t idn = idesc_f16(nn);
              const uint32_t ao = (uint32_t)ks * 2u * (TM / 8) * 128u;
              const uint32_t bo = b_step_off(st);
              const uint64_t dah = niwtc::smem_desc(a_hi + ao, (TM / 8) * 128u, 128u), dal = niwtc::smem_desc(a_lo + ao, (TM / 8) * 128u, 128u);
              const uint64_t dbh = niwtc::smem_desc(b_hi + bo, (nn / 8) * 128u, 128u), dbl = niwtc::smem_desc(b_lo + bo, (nn / 8) * 128u, 128u);
              mma_f16(d_tmem + n0, dah, dbh, idn, 1u);
              mma_f16(d_tmem + n0, dah, dbl, idn, 1u);
              mma_f16(d_tmem + n0, dal, dbh, idn, 1u);
            }
            niwtc::mma_commit(smem_u32(&bars[A_EMPTY + abuf]));            // A half-buffer free once these MMAs complete
            if (half == 1) niwtc::mma_commit(smem_u32(&bars[ACC_FULL + par]));  // accumulator ready for the epilogue
          }
          __syncwarp();
          NIW_TRACE(3 + half * 2);
          advance();
        }
      }
      // this issuer is done with the item's B buffer once its MMAs so far have completed (it commits even when none of
      // the item's tiles was its own: the barrier counts both issuers)
      if (elect_one()) niwtc::mma_commit(smem_u32(&bars[B_FREE + bbuf]));
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..17, all sixteen on EVERY tile.  Warp e reads TMEM lane quadrant warp % 4 (fixed by the
    // hardware: lane = row) and the 64 accumulator columns of group e / 4 of the block: four tcgen05.ld 32x32b.x16
    // issued back to back, ONE wait, and the accumulator is released as soon as the values are in registers -- before
    // any arithmetic (the MMAs of tile t + 2 wait for the readers of tile t).  The accumulator already holds Y' - b'
    // (the bias product of the MMA loop), so what is left per element is one fused multiply-add of the square sum; the
    // group's three coefficients live in registers for the whole item: no shared-memory traffic, no barrier.
    // History of this loop, per 128-row tile against ~1000 cycles of tensor work: the columns split over four / eight
    // warps walking their loads one behind the other with the bias read from shared memory (one LDS.128 per 4 elements,
    // ~3 wavefronts each): ~1900; sixteen warps, loads in flight together: ~1650 once the MMA issue loop no longer paced
    // the kernel; bias in registers through the 16x256b load shape (a thread then needs 16 of the 64 values, but that
    // shape reads TMEM ~5x slower): ~1430.
    const int e = warp - 2;
    const int ew = warp & 3;                 // TMEM lanes 32 ew .. 32 ew + 31
    const int g = e >> 2;                    // group g of the block
    int t = 0;
    for (long long item = blockIdx.x; item < nItems; item += gridDim.x) {
      const int gb = (int)(item % nGB), sl = (int)(item / nGB);
      const int rt_lo = sl * SL, rt_hi = rt_lo + SL < nRT ? rt_lo + SL : nRT;
      const int k = gb * GB + g;
      const bool kok = k < ncols;
      float c0 = 0.f, c1l2 = 0.f, idof = 0.f, base_k = 0.f;
      if (kok) {  // c0, c1 ln 2 (the log below is base 2), 1 / (dof r_k^2): q' = r_k^2 q
        c0 = __ldg(coef + (size_t)k * 4 + 0);
        c1l2 = __ldg(coef + (size_t)k * 4 + 1) * 0.6931471805599453f;
        idof = __ldg(coef + (size_t)k * 4 + 2) * __ldg(rinv + (size_t)k * 2 + 1);
        if (base) base_k = __ldg(base + k);
      }
      // this lane's output element of the item's first tile, then one tile (128 rows) further per step
      const size_t r0 = (size_t)rt_lo * TM + ew * 32 + lane;                      // row index relative to row_lo
      float *dst = (blocked & 1) ? scores + (((size_t)rt_lo * (TM / 32) + ew) * ld + (size_t)k) * 32 + lane
                                 : scores + r0 * ld + (size_t)k;
      const size_t dstep = (blocked & 1) ? (size_t)(TM / 32) * ld * 32 : (size_t)TM * ld;
      long long left = (long long)nrows - (long long)r0;                           // > 0: the row exists
      const uint32_t tbase0 = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(g * 16);
      for (int rt = rt_lo; rt < rt_hi; rt++, t++, dst += dstep, left -= TM) {
        const int acc = t & 1;
        const bool ok = kok && left > 0;
        float old = base_k;
        if (ok && !base) old = *dst;
        mbar_wait(smem_u32(&bars[ACC_FULL + acc]), (uint32_t)((t >> 1) & 1));
        if (warp == 2) NIW_TRACE(6);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // accumulator columns: block i / 16 holds [g][i % 16]; this warp's group is 16 contiguous columns of each block
        const uint32_t tb = tbase0 + (uint32_t)(acc * TN);
        uint32_t r[4][16];
#pragma unroll
        for (int ib = 0; ib < 4; ib++) tmem_ld16_issue(tb + (uint32_t)ib * 64u, r[ib]);
        niwtc::tmem_ld_wait();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[ACC_EMPTY + acc]));
        if (warp == 2) NIW_TRACE(7);
        float2 q2[4];
#pragma unroll
        for (int ib = 0; ib < 4; ib++) {
          q2[ib] = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const float2 y = make_float2(__uint_as_float(r[ib][2 * c]), __uint_as_float(r[ib][2 * c + 1]));
            q2[ib] = __ffma2_rn(y, y, q2[ib]);
          }
        }
        const float q = ((q2[0].x + q2[0].y) + (q2[1].x + q2[1].y)) + ((q2[2].x + q2[2].y) + (q2[3].x + q2[3].y));
        if (ok) {
          float o = old;
          if (q == q) o += fmaf(c1l2, log2_1p_pos(q * idof), c0);  // NaN = masked row: contributes nothing
          *dst = o;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static inline int niw_tc16_init(size_t smem_optin, std::string &err) {
  if (niwtc16::SMEM_BYTES > smem_optin) return MSB_OK;
  cudaError_t e = cudaFuncSetAttribute(niw_tc16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)niwtc16::SMEM_BYTES);
  if (e != cudaSuccess) { err = std::string("niw_tc16_init: ") + cudaGetErrorString(e); return MSB_ERR_CUDA; }
  return MSB_OK;
}
// operand scratch of the fp16 path inside the buffer niw_tc_operand_bytes() sizes (the tf32 operands are twice as big):
// [colmax: D u32][D f32 unused][rinv: nGB x TN f32, 2 used per group][B blocks: nGB x B_BYTES]
static inline size_t niw_tc16_a_bytes(size_t nrows) { return ((nrows + niwtc16::TM - 1) / niwtc16::TM) * (size_t)(4 * niwtc16::A_HALF_BYTES); }
static inline int niw_tc16_score(cudaStream_t stream, uint64_t *launches, KernelProf *prof, const float *X, const float *W, const float *bias,
                                 const float *coef, float *Bop, unsigned char *A16, size_t ncols, float *scores, size_t ld,
                                 size_t row_lo, size_t row_hi, int sm_count, const float *base, bool blocked, bool a_valid, std::string &err) {
  using namespace niwtc16;
  const int nGB = (int)((ncols + GB - 1) / GB);
  unsigned int *colmax = reinterpret_cast<unsigned int *>(Bop);
  float *rinv = Bop + 2 * D;
  unsigned char *Bblk = reinterpret_cast<unsigned char *>(Bop) + (2 * D + (size_t)nGB * TN) * sizeof(float);  // 512 + nGB KB: 16-byte aligned
  auto check = [&](const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return MSB_OK;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return MSB_ERR_CUDA;
  };
  const size_t nrows = row_hi - row_lo;
  const long long nRT = (long long)((nrows + TM - 1) / TM);
  if (!a_valid) {  // a_valid: A16 and colmax already hold exactly these rows of this column data (the caller's version check)
    if (cudaMemsetAsync(colmax, 0, D * sizeof(unsigned int), stream) != cudaSuccess) return check("niw_tc16_score: cudaMemsetAsync");
    const unsigned cm_grid = (unsigned)std::min<size_t>((nrows + 15) / 16, (size_t)sm_count * 8);
    prof->begin("niw_colmax_kernel", stream);
    niw_colmax_kernel<<<cm_grid, 256, 0, stream>>>(X, row_lo, row_hi, colmax);
    prof->end(stream); (*launches)++;
    if (int s = check("niw_colmax_kernel launch")) return s;
    prof->begin("niw_convert_a16_kernel", stream);
    niw_convert_a16_kernel<<<(unsigned)nRT, 256, 0, stream>>>(X, row_lo, row_hi, colmax, A16);
    prof->end(stream); (*launches)++;
    if (int s = check("niw_convert_a16_kernel launch")) return s;
  }
  prof->begin("niw_pack_b16_kernel", stream);
  niw_pack_b16_kernel<<<nGB, 256, 0, stream>>>(W, bias, (int)ncols, colmax, Bblk, rinv);
  prof->end(stream); (*launches)++;
  if (int s = check("niw_pack_b16_kernel launch")) return s;
  // about 24 items per CTA (at most 32 tiles per slice): balance to within a few percent, one B reload per item
  const int SL = (int)std::max<long long>(1, std::min<long long>(32, nRT * nGB / ((long long)sm_count * 24)));
  const long long nItems = ((nRT + SL - 1) / SL) * nGB;
  const int grid = (int)std::min<long long>(sm_count, nItems);
  prof->begin("niw_tc16_kernel", stream);
  niw_tc16_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(A16, Bblk, rinv, bias, coef, (int)ncols, scores, ld, row_lo, row_hi, SL,
                                                         base, blocked ? 1 : 0);
  prof->end(stream); (*launches)++;
  return check("niw_tc16_kernel launch");
}

}  // namespace msb
