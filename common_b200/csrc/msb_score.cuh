// msb_score.cuh -- the score kernel: N rows x K groups, summed over the scalar
// features (bb / dd / gp tables, nich closed form).
//
// Mapping (DESIGN.md "score kernel"):
//   * a block owns NW*RW rows and one k-tile of KT = 32*V groups;
//   * a warp owns RW rows; lane l owns V consecutive groups of the k-tile, so one
//     (row, feature) lookup is ONE conflict-free shared-memory wavefront per 32
//     groups and the row's value x is warp-uniform (shuffled from its owner lane);
//   * acc[RW][V] stays in registers across all features;
//   * the per-(feature, k-tile) parameter chunks are contiguous in global memory
//     in exactly the order the block consumes them, and are streamed through an
//     S-stage shared-memory ring with bulk async copies (cp.async.bulk, SASS
//     UBLKCP) completing on mbarriers -- no register staging, no __syncthreads
//     in the feature loop;
//   * the N x K result is written once with 128*V-byte coalesced stores.
#pragma once
#include "msb_kernels.cuh"

namespace msb {

template <int V> struct VecF;
template <> struct VecF<1> { float v[1]; __device__ __forceinline__ void load(const float *p) { v[0] = *p; } };
template <> struct VecF<2> { float v[2]; __device__ __forceinline__ void load(const float *p) { const float2 t = *(const float2 *)p; v[0] = t.x; v[1] = t.y; } };
template <> struct VecF<4> { float v[4]; __device__ __forceinline__ void load(const float *p) { const float4 t = *(const float4 *)p; v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; } };

template <int V>
__device__ __forceinline__ void store_vec(float *p, const float *a) {
  if constexpr (V == 1) *p = a[0];
  else if constexpr (V == 2) *(float2 *)p = make_float2(a[0], a[1]);
  else *(float4 *)p = make_float4(a[0], a[1], a[2], a[3]);
}

// compact per-feature record kept in shared memory by the score kernel
struct FeatS {
  const void *col;
  uint32_t rowoff;  // first chunk row inside the k-tile region
  uint32_t rows;    // chunk rows
  uint32_t ncat;
  uint16_t kind;
  uint16_t coltype;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on bar
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// gp count beyond the lookup table: fp64 closed form straight from the suffstats.  Kept out of line so
// that its fp64 register pressure stays off the hot loop.
__device__ __noinline__ float gp_overflow_score(const FeatDev *f, const double *hp, const double *ss, int slot, double xv) {
  return (float)gp_score(gp_post(hp + f->hp_off, ss + f->ss_off + (size_t)slot * f->ss_w), xv);
}

// Output layouts of the N x K score matrix:
//   row-major   scores[(row - row_lo) * ld + col]                     (the API's observable output)
//   blocked     scores[((row - row_lo) / 32 * ld + col) * 32 + (row - row_lo) % 32]
//               32-row blocks, group-major inside a block: the sampler walks one row per thread and
//               reads it fully coalesced.  Internal to the sweep.
// TABLES_ONLY: every feature is a lookup table (bb / dd): no kind dispatch, no gp / nich code.
template <int V, int RW, int NW, bool BLOCKED, bool TABLES_ONLY>
__global__ void __launch_bounds__(NW * 32, 1)
score_kernel(const FeatDev *__restrict__ feats, int nfeat, const float *__restrict__ params, size_t region_rows,
             uint32_t stage_bytes, int S, const float *__restrict__ base, float *__restrict__ scores, size_t ld,
             size_t row_lo, size_t row_hi, const double *__restrict__ hp, const double *__restrict__ ss,
             const int32_t *__restrict__ col2slot, int ncols) {
  constexpr int KT = 32 * V;
  constexpr int RL = RW / 32;  // rows per lane whose x this lane loads
  static_assert(RW % 32 == 0, "RW must be a multiple of 32");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [S stages | per-warp exchange buffers (double) | mbarriers full[S], empty[S] | feature table]
  unsigned char *stages = smem_raw;
  uint32_t *xbuf_all = reinterpret_cast<uint32_t *>(smem_raw + (size_t)S * stage_bytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(xbuf_all + 2 * NW * RW);
  FeatS *ftab = reinterpret_cast<FeatS *>(bars + 2 * S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.y;
  const float *region = params + (size_t)kt * region_rows * KT;
  const size_t row0 = row_lo + ((size_t)blockIdx.x * NW + warp) * RW;

  for (int i = tid; i < nfeat; i += NW * 32) {
    const FeatDev f = feats[i];
    FeatS t;
    t.col = f.col; t.rowoff = f.rowoff; t.rows = f.rows; t.ncat = f.ncat;
    t.kind = (uint16_t)f.kind; t.coltype = (uint16_t)f.coltype;
    ftab[i] = t;
  }
  if (tid == 0) {
    for (int s = 0; s < S; s++) {
      mbar_init(smem_u32(&bars[s]), 1);
      mbar_init(smem_u32(&bars[S + s]), NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int d, int s) {  // thread 0 only: chunk of feature d -> stage s
    const FeatS t = ftab[d];
    const uint32_t bytes = t.rows * (uint32_t)(KT * sizeof(float));
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(stages + (size_t)s * stage_bytes), region + (size_t)t.rowoff * KT, bytes, bar);
  };
  if (tid == 0)
    for (int d = 0; d < S && d < nfeat; d++) issue(d, d);

  // raw value of this lane's rows for feature d (bit pattern: table index, count, or float)
  uint32_t valid = 0;
#pragma unroll
  for (int j = 0; j < RL; j++) valid |= (row0 + j * 32 + lane < row_hi ? 1u : 0u) << j;
  const size_t myrow = row0 + lane;
  auto load_x = [&](int d, uint32_t (&x)[RL]) {
    const FeatS t = ftab[d];
    const size_t gcol = __cvta_generic_to_global(t.col);
#pragma unroll
    for (int j = 0; j < RL; j++) {
      if (TABLES_ONLY) x[j] = t.ncat;
      else x[j] = t.kind == KIND_GP ? GP_SENTINEL : (t.kind == KIND_NICH ? 0x7fc00000u : t.ncat);
      if (valid & (1u << j)) {
        const size_t r = myrow + j * 32;
        uint32_t v;
        if (t.coltype == COL_U8) asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(gcol + r));
        else if (t.coltype == COL_U16) asm("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(gcol + 2 * r));
        else asm("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(gcol + 4 * r));  // u32 counts and f32 values alike
        x[j] = v;
      }
    }
  };
  // Per-warp exchange buffer: each lane publishes the table offset (x * KT, in floats) or the float
  // value of the rows it loaded; every lane then reads all RW of them back as broadcast 128-bit
  // shared loads (4 rows per wavefront).  Published one feature ahead, into the other half.
  uint32_t *xbuf = xbuf_all + warp * 2 * RW;
  // (gp counts beyond the table publish the zero row and are flagged in a per-32-row ballot mask)
  auto publish = [&](int d, const uint32_t (&x)[RL], uint32_t (&ovf)[RL]) {
    const FeatS t = ftab[d];
#pragma unroll
    for (int j = 0; j < RL; j++) {
      uint32_t pub = x[j];
      bool over = false;
      if (TABLES_ONLY) pub = x[j] * KT;
      else if (t.kind == KIND_GP) {
        const uint32_t cap = t.ncat;
        over = x[j] != GP_SENTINEL && x[j] >= cap;
        pub = (x[j] < cap ? x[j] : cap) * KT;
      } else if (t.kind == KIND_TABLE) pub = x[j] * KT;
      else {  // nich: a masked cell (NaN) is scored as x = 0 in the branch-free loop and undone afterwards
        over = pub == 0x7fc00000u || __uint_as_float(pub) != __uint_as_float(pub);
        if (over) pub = 0u;
      }
      xbuf[(d & 1) * RW + j * 32 + lane] = pub;
      if (!TABLES_ONLY) ovf[j] = t.kind != KIND_TABLE ? __ballot_sync(0xffffffffu, over) : 0u;
    }
  };

  float acc[RW][V];
#pragma unroll
  for (int r = 0; r < RW; r++)
#pragma unroll
    for (int v = 0; v < V; v++) acc[r][v] = 0.f;

  uint32_t xa[RL], xb[RL];  // xa: feature d+1 (to publish this iteration), xb: feature d+2 (in flight)
  uint32_t ovf_cur[RL], ovf_next[RL];
#pragma unroll
  for (int j = 0; j < RL; j++) ovf_cur[j] = ovf_next[j] = 0u;
  if (nfeat > 0) { load_x(0, xa); publish(0, xa, ovf_cur); }
  if (nfeat > 1) load_x(1, xa);
  int s = 0;
  uint32_t parity = 0;

  for (int d = 0; d < nfeat; d++) {
    if (d + 2 < nfeat) load_x(d + 2, xb);               // global loads two features ahead
    if (d + 1 < nfeat) publish(d + 1, xa, ovf_next);    // shared-memory publish one feature ahead
    const FeatS t = ftab[d];
    __syncwarp();  // feature d's publish (previous iteration) is visible to the whole warp
    mbar_wait(smem_u32(&bars[s]), parity);
    const float *chunk = reinterpret_cast<const float *>(stages + (size_t)s * stage_bytes) + lane * V;
    const uint32_t *xcur = xbuf + (d & 1) * RW;
    const uint4 *xq = reinterpret_cast<const uint4 *>(xcur);

    if (TABLES_ONLY || t.kind != KIND_NICH) {
#pragma unroll
      for (int r4 = 0; r4 < RW / 4; r4++) {
        const uint4 q = xq[r4];
        const uint32_t off[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          VecF<V> tv;
          tv.load(chunk + off[e]);
#pragma unroll
          for (int v = 0; v < V; v++) acc[r4 * 4 + e][v] += tv.v[v];
        }
      }
      if (!TABLES_ONLY && t.kind == KIND_GP) {  // counts beyond the table: the fp64 closed form from the suffstats (rare)
        const uint32_t *raw = reinterpret_cast<const uint32_t *>(t.col);
#pragma unroll
        for (int j = 0; j < RL; j++) {
          uint32_t m = ovf_cur[j];
          while (m) {
            const int rr = j * 32 + __ffs(m) - 1;
            m &= m - 1;
            const double xv = (double)raw[row0 + rr];
            float add[V];
#pragma unroll
            for (int v = 0; v < V; v++) {
              const int col = kt * KT + lane * V + v;
              add[v] = 0.f;
              if (col < ncols) add[v] = gp_overflow_score(feats + d, hp, ss, col2slot[col], xv);
            }
#pragma unroll
            for (int r = 0; r < RW; r++)  // static register indexing: acc must not spill to local memory
              if (r == rr)
#pragma unroll
                for (int v = 0; v < V; v++) acc[r][v] += add[v];
          }
        }
      }
    } else if (!TABLES_ONLY) {  // KIND_NICH: c1' log2(1 + ((x - mu) s)^2); sum_d c0 is already in base[]
      VecF<V> mu, sc, c1;
      mu.load(chunk + 0 * KT);
      sc.load(chunk + 1 * KT);
      c1.load(chunk + 2 * KT);
#pragma unroll
      for (int r4 = 0; r4 < RW / 4; r4++) {
        const uint4 q = xq[r4];
        const float xs[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll
        for (int e = 0; e < 4; e++)
#pragma unroll
          for (int v = 0; v < V; v++) {
            const float tt = (xs[e] - mu.v[v]) * sc.v[v];
            acc[r4 * 4 + e][v] = fmaf(c1.v[v], log2_1p_pos(tt * tt), acc[r4 * 4 + e][v]);
          }
      }
      // masked cells: undo the x = 0 term and the c0 that base[] carries for this feature (rare)
#pragma unroll
      for (int j = 0; j < RL; j++) {
        uint32_t m = ovf_cur[j];
        if (m) {
          VecF<V> c0;
          c0.load(chunk + 3 * KT);
          float undo[V];
#pragma unroll
          for (int v = 0; v < V; v++) {
            const float tt = (0.f - mu.v[v]) * sc.v[v];
            undo[v] = -fmaf(c1.v[v], log2_1p_pos(tt * tt), c0.v[v]);
          }
          while (m) {
            const int rr = j * 32 + __ffs(m) - 1;
            m &= m - 1;
#pragma unroll
            for (int r = 0; r < RW; r++)
              if (r == rr)
#pragma unroll
                for (int v = 0; v < V; v++) acc[r][v] += undo[v];
          }
        }
      }
    }
    // release the stage; thread 0 refills it with chunk d + S once every warp has released it
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[S + s]));
    if (tid == 0 && d + S < nfeat) {
      mbar_wait(smem_u32(&bars[S + s]), parity);
      issue(d + S, s);
    }
    if (++s == S) { s = 0; parity ^= 1u; }
#pragma unroll
    for (int j = 0; j < RL; j++) { xa[j] = xb[j]; ovf_cur[j] = ovf_next[j]; }
  }

  // epilogue: + log(pseudocount) (group_manager.hpp:274-283)
  VecF<V> b;
  b.load(base + (size_t)kt * KT + lane * V);
  if constexpr (!BLOCKED) {  // 128 V-byte coalesced row-major stores
#pragma unroll
    for (int r = 0; r < RW; r++) {
      const size_t row = row0 + r;
      if (row < row_hi) {
        float o[V];
#pragma unroll
        for (int v = 0; v < V; v++) o[v] = acc[r][v] + b.v[v];
        store_vec<V>(scores + (row - row_lo) * ld + (size_t)kt * KT + lane * V, o);
      }
    }
  } else {
    // transpose each 32-row x KT-group tile through shared memory (the drained stage ring) so that
    // lane = row and every store instruction writes 32 consecutive floats of one group
    __syncthreads();  // every warp has finished reading the stages
    float *tile = reinterpret_cast<float *>(stages) + (size_t)warp * 32 * (KT + 1);
#pragma unroll
    for (int j = 0; j < RL; j++) {
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 32; r++)
#pragma unroll
        for (int v = 0; v < V; v++) tile[r * (KT + 1) + lane * V + v] = acc[j * 32 + r][v] + b.v[v];
      __syncwarp();
      const size_t rb = (row0 - row_lo) / 32 + j;  // row0 - row_lo is a multiple of 32
      if (row0 + (size_t)j * 32 < row_hi) {
        float *dst = scores + (rb * ld + (size_t)kt * KT) * 32 + lane;
#pragma unroll 8
        for (int c = 0; c < KT; c++) dst[(size_t)c * 32] = tile[lane * (KT + 1) + c];
      }
    }
  }
}

}  // namespace msb
