// Fused ChebyNet forward on the Blackwell tensor cores (tcgen05 + TMEM) for graphs that fit in shared memory:
//   (perm gather) -> T_k(L~) recursion -> contraction with the taps -> bias -> ReLU -> max-pool (+ mean over filters)
// Replaces cgcnn.chebyshev5 / chebyshev2 + b1relu / b2relu + mpool1 (models_gcn.py:587-617, 558-585, 619-639).
//
// One persistent CTA per SM, 25 warps with three roles:
//   * 20 "sparse" warps run the recursion X_k = 2 L~ X_{k-1} - X_{k-2} out of shared memory.  The state of a tile
//     of windows lives in two ping-pong slabs of 128-byte rows (row = vertex, 32 floats = G windows x FP features,
//     SWIZZLE_128B): a neighbour row is one conflict-free LDS.128 per lane, 8 lanes per row, 4 rows per warp step.
//     Rows are sorted by length once per CTA and dealt to the warps in groups of four (snake order) so that the
//     lanes of a warp stay in lock step and the warps finish together.
//   * 1 MMA warp: after every order, one elected thread issues z += X_k W_k as tcgen05.mma (M = 128 vertices,
//     N = 32 filters) straight from the slab the sparse warps just wrote -- no operand copy, no fragment loads, no
//     accumulator registers.  fp32-level accuracy from a 3-term split on the tensor cores:
//       trunc_tf32(X) * tf32(W)  +  trunc_tf32(X) * tf32(W - tf32(W))  +  bf16(X - trunc_tf32(X)) * bf16(W)
//     (the hardware truncation of kind::tf32 IS the "hi" part of X; only the bf16 remainder is stored separately).
//     Accumulators stay in TMEM across all K orders, double-buffered across tiles.
//   * 4 epilogue warps drain the finished tile of the PREVIOUS iteration from TMEM while the sparse warps already
//     work on the next one: tcgen05.ld (thread = vertex, 32 filters), bias, ReLU, max over p consecutive vertices
//     with the first-maximum rule of MaxPoolGrad, y / arg-max / mean-over-filters stores.
// x is read from HBM once, y written once; the K-stack only goes to HBM when the caller asks for it (training).
#include <algorithm>
#include <cstdlib>

#include "umma.cuh"

namespace gcnb {

using namespace um;

static constexpr int kSparseWarps = 20;
static constexpr int kEpiWarps = 4;   // warps 0..3: warp w may only touch TMEM lanes 32w..32w+31
static constexpr int kMmaWarp = 4;    // warp 4
static constexpr int kThreads = (kSparseWarps + kEpiWarps + 1) * 32;
static constexpr int kBarOrder = 1;   // named barrier: sparse warps + MMA warp, once per Chebyshev order

struct UmmaFwdParams {
  const float* x;
  const int32_t* perm;
  int M_in;
  const int32_t* rowptr;
  const int32_t* col;
  const float* val;
  int nnz;
  const float* W;
  const float* bias;
  float* y;
  uint8_t* argmax;
  float* y_mean;
  float* xstack;
  int B, M, Fin, Fout, K, p, log2p, bias_mode, relu;
  int NS;        // slabs per tile
  int S;         // windows per tile = NS * G
  int MT;        // 128-row MMA tiles per slab
  int NG;        // groups of 4 rows
  int ntiles;
  int nlo;       // lo buffers per slab (2, or 1 when shared memory is short)
  int nacc;      // TMEM accumulator buffers (2, or 1)
  int acc_cols;  // TMEM columns of one accumulator buffer = S * MT * 32
  int tmem_cols; // allocated columns (power of two)
  int slab_rows; // rows of one slab incl. the zero row (multiple of 8)
  // byte offsets into dynamic shared memory
  int off_lo, off_wh, off_wl, off_wb, off_ent, off_grow, off_gslot, off_glen, off_src, off_rlen, off_sorted, off_bias,
      off_bar;
};

// ---------------------------------------------------------------------------------------------------------------
template <int FP>
__global__ void __launch_bounds__(kThreads, 1) k_cheb_fwd_umma(const UmmaFwdParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int G = 32 / FP;    // windows per 128-byte row
  constexpr int CPW = FP / 4;   // 16-byte chunks per window
  const uint32_t sb = smem_u32(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = P.M, K = P.K, NS = P.NS;
  const uint32_t slab_bytes = (uint32_t)P.slab_rows * 128u, lo_bytes = (uint32_t)P.slab_rows * 64u;
  // slab (buffer u, slab s) at sb + (u*NS + s)*slab_bytes ; lo (buffer u, slab s) at sb + off_lo + (u*NS + s)*lo_bytes
  int* grow = reinterpret_cast<int*>(smem + P.off_grow);       // [NG*4] row of (group, slot), -1 = none
  int2* gslot = reinterpret_cast<int2*>(smem + P.off_gslot);   // [NG*4] (first entry pair, row length)
  int* glen = reinterpret_cast<int*>(smem + P.off_glen);       // [NG] entry pairs of the longest row of the group
  int* src_row = reinterpret_cast<int*>(smem + P.off_src);     // [M] source row of the raw window, -1 = zero
  int* rlen = reinterpret_cast<int*>(smem + P.off_rlen);       // [NG*4]
  int* sorted = reinterpret_cast<int*>(smem + P.off_sorted);   // [NG*4]
  float* bias_s = reinterpret_cast<float*>(smem + P.off_bias); // [32]
  const uint32_t bar0 = sb + P.off_bar;
  auto bar_mma = [bar0](uint32_t i) { return bar0 + i * 8u; };          // tcgen05.mma of order n done (n & 1)
  auto bar_full = [bar0](uint32_t i) { return bar0 + 16u + i * 8u; };   // accumulator buffer i complete
  auto bar_empty = [bar0](uint32_t i) { return bar0 + 32u + i * 8u; };  // accumulator buffer i drained
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + P.off_bar + 48);

  // ---- prologue (all warps) ---------------------------------------------------------------------------------
  if ((sb & 1023u) != 0) __trap();  // the swizzled operand layouts need a 1 KB aligned base
  if (tid == 0) {
    mbar_init(bar_mma(0), 1); mbar_init(bar_mma(1), 1);
    mbar_init(bar_full(0), 1); mbar_init(bar_full(1), 1);
    mbar_init(bar_empty(0), kEpiWarps); mbar_init(bar_empty(1), kEpiWarps);
    mbar_init_fence();
  }
  if (warp == kMmaWarp) tmem_alloc(sb + P.off_bar + 48, (uint32_t)P.tmem_cols);
  // zero the state (slabs + lo): the zero row behind every slab and the rows an MMA tile reads past M stay zero
  for (uint32_t a = tid * 16u; a < (uint32_t)P.off_wh; a += kThreads * 16u) sts128(sb + a, make_float4(0.f, 0.f, 0.f, 0.f));
  // taps: tf32 hi / lo and bf16 images, contraction index kk = k*FP + f  (W row = f*K + k, models_gcn.py:611-615)
  for (int idx = tid; idx < K * FP * 32; idx += kThreads) {
    const int o = idx & 31, kk = idx >> 5, k = kk / FP, f = kk - k * FP;
    float w = 0.f;
    if (f < P.Fin && o < P.Fout) w = __ldg(P.W + ((long long)f * K + k) * P.Fout + o);
    const float hi = tf32_rna(w), lo = tf32_rna(w - hi);
    *reinterpret_cast<float*>(smem + P.off_wh + tap_off_tf32(kk, o)) = hi;
    *reinterpret_cast<float*>(smem + P.off_wl + tap_off_tf32(kk, o)) = lo;
    *reinterpret_cast<__nv_bfloat16*>(smem + P.off_wb + tap_off_bf16(kk, o)) = __float2bfloat16_rn(w);
  }
  // operator image: rows sorted by decreasing length; entries re-encoded as (gather code, value) in CSR order,
  // every row starting on an even entry so that one LDS.128 fetches two entries
  const int M4 = P.NG * 4;
  for (int r = tid; r < M4; r += kThreads) rlen[r] = r < M ? __ldg(P.rowptr + r + 1) - __ldg(P.rowptr + r) : -1;
  for (int r = tid; r < M; r += kThreads) {
    int s = r;
    if (P.perm) { s = __ldg(P.perm + r); if (s < 0 || s >= P.M_in) s = -1; }
    src_row[r] = s;
  }
  if (tid < 32) bias_s[tid] = (P.bias_mode == GCNB_BIAS_PER_FILTER && tid < P.Fout) ? __ldg(P.bias + tid) : 0.f;
  {
    const int npairs = (P.nnz + M + 2) >> 1;  // zero entries everywhere (padding entry of odd rows: zero row, 0.0)
    const uint32_t zc = gather_code(M);
    int4* e4 = reinterpret_cast<int4*>(smem + P.off_ent);
    for (int i = tid; i < npairs; i += kThreads) e4[i] = make_int4((int)zc, 0, (int)zc, 0);
  }
  __syncthreads();
  for (int r = tid; r < M4; r += kThreads) {
    const int l = rlen[r];
    int rank = 0;
    for (int o = 0; o < M4; ++o) {
      const int lo = rlen[o];
      rank += (lo > l || (lo == l && o < r)) ? 1 : 0;
    }
    sorted[rank] = r;
  }
  {
    int2* ent = reinterpret_cast<int2*>(smem + P.off_ent);
    for (int r = warp; r < M; r += kThreads / 32) {
      const int beg = __ldg(P.rowptr + r), len = rlen[r];
      const int es = (beg + r + 1) & ~1;
      for (int j = lane; j < len; j += 32)
        ent[es + j] = make_int2((int)gather_code(__ldg(P.col + beg + j)), __float_as_int(__ldg(P.val + beg + j)));
    }
  }
  __syncthreads();
  for (int i = tid; i < M4; i += kThreads) {
    const int r = sorted[i];
    if (r < M) {
      grow[i] = r;
      gslot[i] = make_int2(((__ldg(P.rowptr + r) + r + 1) & ~1) >> 1, rlen[r]);
    } else {
      grow[i] = -1;
      gslot[i] = make_int2(0, 0);
    }
    if ((i & 3) == 0) glen[i >> 2] = r < M ? (rlen[r] + 1) >> 1 : 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int NI = P.NG * NS;  // work items (row group, slab) of one order
  const int nsync = (kSparseWarps + 1) * 32;

  if (warp > kMmaWarp) {
    // =========================================== sparse warps ==================================================
    const int sw = warp - (kMmaWarp + 1);
    const int q = lane >> 3, c = lane & 7;
    const uint32_t c16 = (uint32_t)c << 4;
    const int gw = c / CPW, fc = c - gw * CPW;  // window inside the row, feature chunk inside the window
    const uint32_t ent_base = sb + P.off_ent;
    uint32_t n = 0;  // orders issued so far (all tiles)
    int base = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      // ---- order 0: the (gathered, zero padded) raw windows -----------------------------------------------
      {
        const uint32_t lo_u = P.nlo == 2 ? (n & 1u) : 0u;
        if (P.nlo == 1 && n > 0) mbar_wait(bar_mma((n - 1) & 1), ((n - 1) >> 1) & 1);
        for (int r0 = 0; r0 * kSparseWarps < NI; r0 += 4) {
          float4 v[4];
          int rows[4], slabs[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + u;
            const int ii = r * kSparseWarps + ((r & 1) ? kSparseWarps - 1 - sw : sw);
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            rows[u] = -1;
            slabs[u] = 0;
            if (ii < NI) {
              const int g = ii / NS, s = ii - g * NS;
              const int row = grow[g * 4 + q];
              rows[u] = row;
              slabs[u] = s;
              const int b = tile * P.S + s * G + gw;
              if (row >= 0 && b < P.B) {
                const int src = src_row[row];
                if (src >= 0) {
                  const float* xp = P.x + ((long long)b * P.M_in + src) * P.Fin + fc * 4;
                  const int nf = P.Fin - fc * 4;
                  if (nf > 0) v[u].x = __ldg(xp);
                  if (nf > 1) v[u].y = __ldg(xp + 1);
                  if (nf > 2) v[u].z = __ldg(xp + 2);
                  if (nf > 3) v[u].w = __ldg(xp + 3);
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (rows[u] >= 0) {
              const int s = slabs[u], row = rows[u];
              store_state(sb + (uint32_t)(base * NS + s) * slab_bytes, sb + P.off_lo + (lo_u * NS + s) * lo_bytes, row, c,
                          v[u]);
              const int b = tile * P.S + s * G + gw;
              if (P.xstack && b < P.B)
                *reinterpret_cast<float4*>(P.xstack + ((long long)b * M + row) * FP + fc * 4) = v[u];
            }
          }
        }
        fence_async_smem();
        named_bar_sync(kBarOrder, nsync);
        ++n;
      }
      // ---- orders 1 .. K-1 --------------------------------------------------------------------------------
      for (int k = 1; k < K; ++k) {
        const int cur = (base + k) & 1;
        const uint32_t lo_u = P.nlo == 2 ? (n & 1u) : 0u;
        if (P.nlo == 1) mbar_wait(bar_mma((n - 1) & 1), ((n - 1) >> 1) & 1);
        float* spill = P.xstack ? P.xstack + (long long)k * P.B * M * FP : nullptr;
        for (int r = 0;; ++r) {
          const int ii = r * kSparseWarps + ((r & 1) ? kSparseWarps - 1 - sw : sw);
          if (ii >= NI) break;
          const int g = ii / NS, s = ii - g * NS;
          const int row = grow[g * 4 + q];
          const int2 meta = gslot[g * 4 + q];
          const int len2 = glen[g];
          const uint32_t src = sb + (uint32_t)((cur ^ 1) * NS + s) * slab_bytes;
          const uint32_t dst = sb + (uint32_t)(cur * NS + s) * slab_bytes;
          const uint32_t ea = ent_base + (uint32_t)meta.x * 16u;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 2
          for (int j = 0; j < len2; ++j) {
            if (2 * j < meta.y) {
              const int4 e = lds128i(ea + (uint32_t)j * 16u);
              const float4 v0 = lds128(src + ((uint32_t)e.x ^ c16));
              const float4 v1 = lds128(src + ((uint32_t)e.z ^ c16));
              const float w0 = __int_as_float(e.y), w1 = __int_as_float(e.w);
              a0 = fmaf(w0, v0.x, a0); a1 = fmaf(w0, v0.y, a1); a2 = fmaf(w0, v0.z, a2); a3 = fmaf(w0, v0.w, a3);
              a0 = fmaf(w1, v1.x, a0); a1 = fmaf(w1, v1.y, a1); a2 = fmaf(w1, v1.z, a2); a3 = fmaf(w1, v1.w, a3);
            }
          }
          if (row >= 0) {
            float4 o;
            if (k == 1) {
              o = make_float4(a0, a1, a2, a3);
            } else {
              const float4 own = lds128(dst + slab_off(row, c));
              o = make_float4(fmaf(2.f, a0, -own.x), fmaf(2.f, a1, -own.y), fmaf(2.f, a2, -own.z), fmaf(2.f, a3, -own.w));
            }
            store_state(dst, sb + P.off_lo + (lo_u * NS + s) * lo_bytes, row, c, o);
            const int b = tile * P.S + s * G + gw;
            if (spill && b < P.B) *reinterpret_cast<float4*>(spill + ((long long)b * M + row) * FP + fc * 4) = o;
          }
        }
        fence_async_smem();
        named_bar_sync(kBarOrder, nsync);
        ++n;
      }
      base = (base + K) & 1;
    }
  } else if (warp == kMmaWarp) {
    // =========================================== MMA warp ======================================================
    uint32_t n = 0;
    int base = 0, it = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const int buf = P.nacc == 2 ? (it & 1) : 0;
      const int use = P.nacc == 2 ? (it >> 1) : it;  // earlier uses of this accumulator buffer
      if (use > 0) mbar_wait(bar_empty(buf), (uint32_t)(use - 1) & 1u);
      for (int k = 0; k < K; ++k) {
        named_bar_sync(kBarOrder, nsync);  // X_k (and its remainder) is complete in shared memory
        tc_fence_after();
        if (lane == 0) {
          const int cur = (base + k) & 1;
          const uint32_t lo_u = P.nlo == 2 ? (n & 1u) : 0u;
          const uint32_t wh = sb + P.off_wh + (uint32_t)(k * (FP / 4) * 4) * 128u;
          const uint32_t wl = sb + P.off_wl + (uint32_t)(k * (FP / 4) * 4) * 128u;
          const uint32_t wb = sb + P.off_wb + (uint32_t)(k * (FP / 8) * 4) * 128u;
          for (int s = 0; s < NS; ++s) {
            const uint32_t slab = sb + (uint32_t)(cur * NS + s) * slab_bytes;
            const uint32_t lo = sb + P.off_lo + (lo_u * NS + s) * lo_bytes;
#pragma unroll
            for (int g = 0; g < G; ++g) {
              for (int i = 0; i < P.MT; ++i) {
                const uint32_t d = tmem + (uint32_t)(buf * P.acc_cols + ((s * G + g) * P.MT + i) * 32);
                const uint32_t a = slab + (uint32_t)i * 16384u + (uint32_t)(g * FP * 4);
                const uint32_t l = lo + (uint32_t)i * 8192u + (uint32_t)(g * FP * 2);
#pragma unroll
                for (int j = 0; j < FP / 8; ++j)
                  mma_tf32(d, smem_desc(kDescSlab, a + j * 32), smem_desc(kDescTaps, wh + j * 1024), kIdescTf32,
                           (k | j) != 0);
#pragma unroll
                for (int j = 0; j < FP / 8; ++j)
                  mma_tf32(d, smem_desc(kDescSlab, a + j * 32), smem_desc(kDescTaps, wl + j * 1024), kIdescTf32, 1);
#pragma unroll
                for (int j = 0; j < FP / 16; ++j)
                  mma_bf16(d, smem_desc(kDescLo, l + j * 32), smem_desc(kDescTaps, wb + j * 1024), kIdescBf16, 1);
              }
            }
          }
          mma_commit(bar_mma(n & 1));
          if (k == K - 1) mma_commit(bar_full(buf));
        }
        __syncwarp();
        // X_k's slab and remainder may be overwritten two orders from now: only pass the next barrier once the
        // tensor cores have finished reading them
        mbar_wait(bar_mma(n & 1), (n >> 1) & 1);
        ++n;
      }
      base = (base + K) & 1;
    }
  } else {
    // =========================================== epilogue warps =================================================
    const int e = warp;  // TMEM lane quarter
    const int p = P.p, log2p = P.log2p;
    const int Mo = M >> log2p;
    const int il = lane & (p - 1);  // position inside the pooling window
    int it = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const int buf = P.nacc == 2 ? (it & 1) : 0;
      const int use = P.nacc == 2 ? (it >> 1) : it;
      mbar_wait(bar_full(buf), (uint32_t)use & 1u);
      tc_fence_after();
      for (int w = 0; w < P.S; ++w) {
        const int b = tile * P.S + w;
        for (int i = 0; i < P.MT; ++i) {
          const int row0 = i * 128 + e * 32;
          if (row0 >= M) continue;
          float a[32];
          tmem_ld32(tmem + ((uint32_t)(e * 32) << 16) + (uint32_t)(buf * P.acc_cols + (w * P.MT + i) * 32), a);
          const int row = row0 + lane;
          const bool valid = row < M && b < P.B;
          if (P.bias_mode == GCNB_BIAS_PER_FILTER) {
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias_s + c4 * 4);
              a[c4 * 4] += bv.x; a[c4 * 4 + 1] += bv.y; a[c4 * 4 + 2] += bv.z; a[c4 * 4 + 3] += bv.w;
            }
          } else if (P.bias_mode == GCNB_BIAS_PER_VERTEX) {
            if (valid) {
              const float* bp = P.bias + (long long)row * P.Fout;
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                if (c4 * 4 < P.Fout) {
                  const float4 bv = __ldg(reinterpret_cast<const float4*>(bp + c4 * 4));
                  a[c4 * 4] += bv.x; a[c4 * 4 + 1] += bv.y; a[c4 * 4 + 2] += bv.z; a[c4 * 4 + 3] += bv.w;
                }
              }
            }
          }
          if (P.relu) {
#pragma unroll
            for (int o = 0; o < 32; ++o) a[o] = fmaxf(a[o], 0.f);
          }
          uint32_t am[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          if (p > 1) {
            const uint32_t gmask = (1u << p) - 1u;
            const int gshift = lane & ~(p - 1);
#pragma unroll
            for (int o = 0; o < 32; ++o) {
              float m = a[o];
              for (int d = 1; d < p; d <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
              const uint32_t hit = (__ballot_sync(0xffffffffu, a[o] == m) >> gshift) & gmask;
              const uint32_t first = hit ? (uint32_t)(__ffs((int)hit) - 1) : 0u;  // first maximum (MaxPoolGrad)
              a[o] = m;
              am[o >> 2] |= first << ((o & 3) * 8);
            }
          }
          if (valid) {
            const long long orow = (long long)b * Mo + (row >> log2p);
            float* yp = P.y + orow * P.Fout;
            uint8_t* ap = (P.argmax && p > 1) ? P.argmax + orow * P.Fout : nullptr;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              if (c4 * 4 < P.Fout && (c4 >> (3 - log2p)) == il) {
                *reinterpret_cast<float4*>(yp + c4 * 4) = make_float4(a[c4 * 4], a[c4 * 4 + 1], a[c4 * 4 + 2], a[c4 * 4 + 3]);
                if (ap) *reinterpret_cast<uint32_t*>(ap + c4 * 4) = am[c4];
              }
            }
            if (P.y_mean && il == 0) {
              float sm = 0.f;
#pragma unroll
              for (int o = 0; o < 32; ++o)
                if (o < P.Fout) sm += a[o];
              P.y_mean[orow] = sm / (float)P.Fout;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty(buf));
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)P.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
struct UmmaFwdPlan {
  bool ok;
  int FP, G, NS, MT, NG, nlo, nacc, acc_cols, tmem_cols, slab_rows;
  size_t smem;
  int off_lo, off_wh, off_wl, off_wb, off_ent, off_grow, off_gslot, off_glen, off_src, off_rlen, off_sorted, off_bias,
      off_bar;
};

static bool umma_enabled() {
  static const bool on = [] {
    const char* v = std::getenv("GCNB_UMMA");
    return !(v && v[0] == '0');
  }();
  return on;
}

static UmmaFwdPlan plan_umma_fwd(const LayerShape& s, int sm_count, int smem_optin) {
  UmmaFwdPlan pl{};
  pl.ok = false;
  if (!umma_enabled()) return pl;
  if (s.Fin <= 8 || s.Fin > 32 || s.Fout < 4 || s.Fout > 32 || (s.Fout & 3)) return pl;
  if (s.p != 1 && s.p != 2 && s.p != 4 && s.p != 8) return pl;
  if (s.M % s.p != 0 || s.M < 1 || s.M > 512 || s.K < 1 || s.B < 1) return pl;
  pl.FP = s.Fin <= 16 ? 16 : 32;
  pl.G = 32 / pl.FP;
  pl.MT = ceil_div(s.M, 128);
  pl.NG = ceil_div(s.M, 4);
  pl.slab_rows = (int)align_up((size_t)s.M + 1, 8);
  const size_t slab = (size_t)pl.slab_rows * 128, lo = (size_t)pl.slab_rows * 64;
  const size_t taps = (size_t)s.K * pl.FP * 32 * 4;  // one tf32 image
  const size_t ent = align_up(((size_t)s.nnz + s.M + 2) * 8, 16);
  const size_t tables = (size_t)pl.NG * 4 * 4 /*grow*/ + (size_t)pl.NG * 4 * 8 /*gslot*/ + align_up((size_t)pl.NG * 4, 16) +
                        align_up((size_t)s.M * 4, 16) /*src*/ + 2 * (size_t)pl.NG * 4 * 4 /*rlen, sorted*/ + 128 /*bias*/ +
                        64 /*barriers*/;
  const size_t budget = std::min<size_t>((size_t)smem_optin, 227 * 1024);
  double best = 1e30;
  for (int ns = 1; ns <= 4; ++ns) {
    const int S = ns * pl.G;
    const int acc_cols = S * pl.MT * 32;
    if (acc_cols > 512) continue;
    for (int nlo = 2; nlo >= 1; --nlo) {
      const size_t need = 2 * ns * slab + (size_t)nlo * ns * lo + 2 * taps + taps / 2 + ent + tables;
      // the last MMA tile of a slab reads up to row MT*128: those bytes must exist inside the allocation
      const size_t overrun = (size_t)(pl.MT * 128 - pl.slab_rows) * 128;
      if (need > budget || 2 * ns * slab + overrun > need) continue;
      const int tiles = ceil_div(s.B, S);
      const double cost = std::ceil((double)tiles / sm_count) * ns * (nlo == 2 ? 1.0 : 1.05) - 0.001 * ns;
      if (cost < best) {
        best = cost;
        pl.NS = ns;
        pl.nlo = nlo;
        pl.acc_cols = acc_cols;
        pl.smem = need;
      }
      break;
    }
  }
  if (best > 1e29) return pl;
  pl.nacc = 2 * pl.acc_cols <= 512 ? 2 : 1;
  int cols = 32;
  while (cols < pl.nacc * pl.acc_cols) cols *= 2;
  pl.tmem_cols = cols;
  size_t off = 2 * (size_t)pl.NS * slab;
  pl.off_lo = (int)off; off += (size_t)pl.nlo * pl.NS * lo;
  pl.off_wh = (int)off; off += taps;
  pl.off_wl = (int)off; off += taps;
  pl.off_wb = (int)off; off += taps / 2;
  pl.off_ent = (int)off; off += ent;
  pl.off_grow = (int)off; off += (size_t)pl.NG * 16;
  pl.off_gslot = (int)off; off += (size_t)pl.NG * 32;
  pl.off_glen = (int)off; off += align_up((size_t)pl.NG * 4, 16);
  pl.off_src = (int)off; off += align_up((size_t)s.M * 4, 16);
  pl.off_rlen = (int)off; off += (size_t)pl.NG * 16;
  pl.off_sorted = (int)off; off += (size_t)pl.NG * 16;
  pl.off_bias = (int)off; off += 128;
  pl.off_bar = (int)off; off += 64;
  if (off > pl.smem) return pl;
  pl.ok = true;
  return pl;
}

bool umma_fwd_supported(const LayerShape& s) {
  DeviceInfo di;
  if (device_info(&di) != GCNB_OK) {  // no device visible (shape queries on a CPU box): assume a B200
    di.sm_count = 148;
    di.smem_optin = 227 * 1024;
  }
  return plan_umma_fwd(s, di.sm_count, di.smem_optin).ok;
}

int umma_cheb_fwd(const float* x, const int32_t* perm, int M_in, const gcnb_csr& L, const float* W, const float* bias,
                  float* y, uint8_t* argmax, float* y_mean, float* xstack, const LayerShape& s, int bias_mode, int relu,
                  cudaStream_t st) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  const UmmaFwdPlan pl = plan_umma_fwd(s, di.sm_count, di.smem_optin);
  if (!pl.ok) {
    set_error("tcgen05 forward does not support this shape");
    return GCNB_ERR_INVALID;
  }
  UmmaFwdParams P{};
  P.x = x; P.perm = perm; P.M_in = M_in;
  P.rowptr = L.rowptr; P.col = L.col; P.val = L.val; P.nnz = L.nnz;
  P.W = W; P.bias = bias; P.y = y; P.argmax = argmax; P.y_mean = y_mean; P.xstack = xstack;
  P.B = s.B; P.M = s.M; P.Fin = s.Fin; P.Fout = s.Fout; P.K = s.K; P.p = s.p; P.bias_mode = bias_mode; P.relu = relu;
  P.log2p = 0;
  while ((1 << P.log2p) < s.p) ++P.log2p;
  P.NS = pl.NS; P.S = pl.NS * pl.G; P.MT = pl.MT; P.NG = pl.NG; P.nlo = pl.nlo; P.nacc = pl.nacc;
  P.acc_cols = pl.acc_cols; P.tmem_cols = pl.tmem_cols; P.slab_rows = pl.slab_rows;
  P.ntiles = ceil_div(s.B, P.S);
  P.off_lo = pl.off_lo; P.off_wh = pl.off_wh; P.off_wl = pl.off_wl; P.off_wb = pl.off_wb; P.off_ent = pl.off_ent;
  P.off_grow = pl.off_grow; P.off_gslot = pl.off_gslot; P.off_glen = pl.off_glen; P.off_src = pl.off_src;
  P.off_rlen = pl.off_rlen; P.off_sorted = pl.off_sorted; P.off_bias = pl.off_bias; P.off_bar = pl.off_bar;
  const int grid = std::min(P.ntiles, di.sm_count);
  if (pl.FP == 16) {
    GCNB_CUDA(cudaFuncSetAttribute(k_cheb_fwd_umma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    k_cheb_fwd_umma<16><<<grid, kThreads, pl.smem, st>>>(P);
  } else {
    GCNB_CUDA(cudaFuncSetAttribute(k_cheb_fwd_umma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    k_cheb_fwd_umma<32><<<grid, kThreads, pl.smem, st>>>(P);
  }
  GCNB_LAUNCH_CHECK("k_cheb_fwd_umma");
  return GCNB_OK;
}

}  // namespace gcnb
