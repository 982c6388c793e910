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

template <int V, int RW, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
score_kernel(const FeatDev *__restrict__ feats, int nfeat, const float *__restrict__ params, size_t region_rows,
             uint32_t stage_bytes, int S, const float *__restrict__ base, float *__restrict__ scores, size_t ld,
             size_t row_lo, size_t row_hi) {
  constexpr int KT = 32 * V;
  constexpr int RL = RW / 32;  // rows per lane whose x this lane loads
  static_assert(RW % 32 == 0, "RW must be a multiple of 32");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *stages = smem_raw;
  // layout: [S stages | per-warp exchange buffers | mbarriers full[S], empty[S] | feature table]
  uint32_t *xbuf_all = reinterpret_cast<uint32_t *>(smem_raw + (size_t)S * stage_bytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(xbuf_all + NW * RW);
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

  auto issue = [&](int d) {  // thread 0 only: chunk of feature d -> stage d % S
    const FeatS t = ftab[d];
    const uint32_t bytes = t.rows * (uint32_t)(KT * sizeof(float));
    const uint32_t bar = smem_u32(&bars[d % S]);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(stages + (size_t)(d % S) * stage_bytes), region + (size_t)t.rowoff * KT, bytes, bar);
  };
  if (tid == 0)
    for (int d = 0; d < S && d < nfeat; d++) issue(d);

  auto load_x = [&](int d, uint32_t (&xi)[RL], float (&xf)[RL]) {
    const FeatS t = ftab[d];
#pragma unroll
    for (int j = 0; j < RL; j++) {
      const size_t row = row0 + j * 32 + lane;
      xi[j] = t.kind == KIND_GP ? GP_SENTINEL : t.ncat;
      xf[j] = CUDART_NAN_F;
      if (row < row_hi) {
        if (t.coltype == COL_U8) xi[j] = ((const uint8_t *)t.col)[row];
        else if (t.coltype == COL_U16) xi[j] = ((const uint16_t *)t.col)[row];
        else if (t.coltype == COL_U32) xi[j] = ((const uint32_t *)t.col)[row];
        else xf[j] = ((const float *)t.col)[row];
      }
    }
  };

  float acc[RW][V];
#pragma unroll
  for (int r = 0; r < RW; r++)
#pragma unroll
    for (int v = 0; v < V; v++) acc[r][v] = 0.f;

  // Per-warp exchange buffer: each lane publishes the table offset (x * KT, in floats) or the float
  // value of the rows it loaded; every lane then reads all RW of them back as broadcast 128-bit
  // shared loads (4 rows per wavefront).  A per-row warp shuffle would cost one crossbar
  // wavefront per row -- as much as the lookup itself.
  uint32_t *xbuf = xbuf_all + warp * RW;

  uint32_t xi[RL], xi_next[RL];
  float xf[RL], xf_next[RL];
  if (nfeat > 0) load_x(0, xi, xf);

  for (int d = 0; d < nfeat; d++) {
    if (d + 1 < nfeat) load_x(d + 1, xi_next, xf_next);  // prefetch the next feature's values
    const FeatS t = ftab[d];
    const int s = d % S;
    const uint32_t parity = (uint32_t)(d / S) & 1u;
    bool overflow = false;
    __syncwarp();  // the previous feature's reads of xbuf are done
#pragma unroll
    for (int j = 0; j < RL; j++) {
      uint32_t pub;
      if (t.kind == KIND_NICH) pub = __float_as_uint(xf[j]);
      else if (t.kind == KIND_GP) {
        const uint32_t cap = t.ncat;
        const uint32_t xr = xi[j] == GP_SENTINEL ? cap : (xi[j] < cap ? xi[j] : cap + 1);
        overflow |= xr > cap;
        pub = xr * KT;
      } else pub = xi[j] * KT;
      xbuf[j * 32 + lane] = pub;
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[s]), parity);
    const float *chunk = reinterpret_cast<const float *>(stages + (size_t)s * stage_bytes) + lane * V;
    const uint4 *xq = reinterpret_cast<const uint4 *>(xbuf);

    if (t.kind == KIND_TABLE || (t.kind == KIND_GP && !__any_sync(0xffffffffu, overflow))) {
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
    } else if (t.kind == KIND_GP) {  // some row of this warp has a count beyond the table (rare)
      const uint32_t cap = t.ncat;
#pragma unroll
      for (int r = 0; r < RW; r++) {
        const uint32_t off = xbuf[r];
        if (off <= cap * KT) {
          VecF<V> tv;
          tv.load(chunk + off);
#pragma unroll
          for (int v = 0; v < V; v++) acc[r][v] += tv.v[v];
        } else {  // evaluate the closed form
          const float xv = (float)__shfl_sync(0xffffffffu, xi[r >> 5], r & 31);
          const float lgx1 = lgammaf(xv + 1.f);
#pragma unroll
          for (int v = 0; v < V; v++) {
            const float a = chunk[(size_t)(cap + 1) * KT + v];
            const float ca = chunk[(size_t)(cap + 2) * KT + v];
            const float l1pb = chunk[(size_t)(cap + 3) * KT + v];
            acc[r][v] += lgammaf(a + xv) - lgx1 + ca - xv * l1pb;
          }
        }
      }
    } else {  // KIND_NICH
      VecF<V> mu, sc, c1, c0;
      mu.load(chunk + 0 * KT);
      sc.load(chunk + 1 * KT);
      c1.load(chunk + 2 * KT);
      c0.load(chunk + 3 * KT);
#pragma unroll
      for (int r4 = 0; r4 < RW / 4; r4++) {
        const uint4 q = xq[r4];
        const float xs[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float x = xs[e];
          if (x == x) {
#pragma unroll
            for (int v = 0; v < V; v++) {
              const float tt = (x - mu.v[v]) * sc.v[v];
              acc[r4 * 4 + e][v] += fmaf(c1.v[v], log1p_pos(tt * tt), c0.v[v]);
            }
          }
        }
      }
    }
    // release the stage; thread 0 refills it with chunk d + S once every warp has released it
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[S + s]));
    if (tid == 0 && d + S < nfeat) {
      mbar_wait(smem_u32(&bars[S + s]), parity);
      issue(d + S);
    }
#pragma unroll
    for (int j = 0; j < RL; j++) { xi[j] = xi_next[j]; xf[j] = xf_next[j]; }
  }

  // epilogue: + log(pseudocount) (group_manager.hpp:274-283), coalesced 128 V-byte stores
  VecF<V> b;
  b.load(base + (size_t)kt * KT + lane * V);
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
}

}  // namespace msb
