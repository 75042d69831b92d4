// Tensor-core contractions of the general (HBM-resident) path: vertex-level graphs, BASELINE config 5.
//
// The general path keeps the Chebyshev stack in HBM in vertex-major order, Xs[k][r][f] with r = m*B + b, so
// every [rows x Fin] tile of one order is ONE contiguous span -- a single TMA bulk copy per (tile, order).
// Three GEMM-shaped steps stream that stack; each is a warp-specialised kernel: one producer warp feeds a
// ring of shared-memory stages with cp.async.bulk + mbarrier (full/empty pairs), eight consumer warps run
// mma.sync m16n8k8 TF32 with the 3-pass error-compensated split (fp32-level accuracy), one CTA per SM.
//
//   k_stack_contract : z[r][o]   = sum_k X_k[r][:] W_k          (forward; bias/ReLU fused when p == 1)
//   k_node_dw        : dW_k[f][o] = sum_r X_k[r][f] dZ[r][o]    (weight gradient, per-CTA partials)
//   k_dz_wt          : G_k[r][f] = sum_o dZ[r][o] W_k[f][o]     (seeds of the adjoint recursion; TMA stores)
//
// Reference math: models_gcn.py:611-616 (x W with W row = fin*K + k) and its tf.gradients, :298-303.
#include "fused_common.cuh"

namespace gcnb {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;  // + one producer warp
constexpr int kMaxStages = 8;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared -> global bulk store (TMA), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
}

struct Ring {
  uint64_t* full;
  uint64_t* empty;
  int NS;
  __device__ __forceinline__ void init(unsigned char* base, int ns, int tid) {
    full = reinterpret_cast<uint64_t*>(base);
    empty = full + kMaxStages;
    NS = ns;
    if (tid == 0)
      for (int s = 0; s < ns; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, kConsumerWarps);
      }
  }
  // producer: wait until the consumers released the previous use of the stage item `it` maps to
  __device__ __forceinline__ int acquire(long long it) const {
    const int s = (int)(it % NS);
    const long long u = it / NS;
    if (u > 0) mbar_wait(empty + s, (uint32_t)((u - 1) & 1));
    return s;
  }
  __device__ __forceinline__ int wait_full(long long it) const {
    const int s = (int)(it % NS);
    mbar_wait(full + s, (uint32_t)((it / NS) & 1));
    return s;
  }
  __device__ __forceinline__ void release(int s, int lane) const {
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
};
constexpr int kRingBytes = 2 * kMaxStages * 8;  // 128

// ------------------------------------------------------------------------------------------------
// forward contraction
// ------------------------------------------------------------------------------------------------
struct ContractArgs {
  const float* Xs;
  long long slab;  // floats between orders
  const float* W;
  const float* bias;
  float* out;  // direct: y[b][m][o] (bias/ReLU applied); else Zn[r][o]
  long long R;
  int B, M, Fin, Fout, K, KC, bias_mode, relu, direct, NS, ntiles;
};

constexpr int kTileRows = 256;  // 8 consumer warps x 32 rows

template <int NT>
__global__ void __launch_bounds__(kThreads, 1) k_stack_contract(ContractArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  Ring ring;
  ring.init(smem, a.NS, tid);
  float2* wfrag = reinterpret_cast<float2*>(smem + kRingBytes);
  const int nw = a.K * a.KC * NT * 32;
  const uint32_t stage_bytes = (uint32_t)kTileRows * a.Fin * 4;
  unsigned char* stages = smem + kRingBytes + align_up((size_t)nw * 8, 128);
  for (int i = tid; i < nw; i += kThreads) {
    const int l = i & 31, nt = (i >> 5) % NT, kc = ((i >> 5) / NT) % a.KC, k = (i >> 5) / (NT * a.KC);
    const int f0 = kc * 8 + (l & 3), f1 = f0 + 4, o = nt * 8 + (l >> 2);
    float2 v;
    v.x = (f0 < a.Fin && o < a.Fout) ? __ldg(a.W + ((long long)f0 * a.K + k) * a.Fout + o) : 0.f;
    v.y = (f1 < a.Fin && o < a.Fout) ? __ldg(a.W + ((long long)f1 * a.K + k) * a.Fout + o) : 0.f;
    wfrag[i] = v;
  }
  __syncthreads();

  if (warp == kConsumerWarps) {  // ---- producer
    if (lane == 0) {
      long long it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const long long r0 = (long long)tile * kTileRows;
        const long long rows = a.R - r0 < kTileRows ? a.R - r0 : kTileRows;
        const uint32_t bytes = (uint32_t)(rows * a.Fin * 4);
        for (int k = 0; k < a.K; ++k, ++it) {
          const int s = ring.acquire(it);
          mbar_expect_tx(ring.full + s, bytes);
          bulk_g2s_ring(stages + (size_t)s * stage_bytes, a.Xs + (long long)k * a.slab + r0 * a.Fin, bytes, ring.full + s);
        }
      }
    }
    return;
  }

  // ---- consumers: warp w owns rows [32w, 32w+32) of the tile, all NT column tiles
  const int Fin = a.Fin;
  long long it = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
    for (int k = 0; k < a.K; ++k, ++it) {
      const int s = ring.wait_full(it);
      const float* T = reinterpret_cast<const float*>(stages + (size_t)s * stage_bytes);
      for (int kc = 0; kc < a.KC; ++kc) {
        const int c0 = kc * 8 + t, c1 = c0 + 4;
        const bool ok0 = c0 < Fin, ok1 = c1 < Fin;
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float* row = T + (warp * 32 + mt * 16 + g) * Fin;
          const float v0 = ok0 ? row[c0] : 0.f;
          const float v1 = ok0 ? row[8 * Fin + c0] : 0.f;
          const float v2 = ok1 ? row[c1] : 0.f;
          const float v3 = ok1 ? row[8 * Fin + c1] : 0.f;
          split_trunc(v0, ah[mt][0], al[mt][0]);
          split_trunc(v1, ah[mt][1], al[mt][1]);
          split_trunc(v2, ah[mt][2], al[mt][2]);
          split_trunc(v3, ah[mt][3], al[mt][3]);
        }
        const float2* wf = wfrag + ((long long)(k * a.KC + kc) * NT) * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float2 b = wf[nt * 32];
          uint32_t bh0, bl0, bh1, bl1;
          split_trunc(b.x, bh0, bl0);
          split_trunc(b.y, bh1, bl1);
          mma_3xtf32(acc[0][nt], ah[0], al[0], bh0, bh1, bl0, bl1);
          mma_3xtf32(acc[1][nt], ah[1], al[1], bh0, bh1, bl0, bl1);
        }
      }
      ring.release(s, lane);
    }
    // epilogue: rows r = m*B + b
    const long long r0 = (long long)tile * kTileRows + warp * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const long long r = r0 + mt * 16 + half * 8 + g;
        if (r >= a.R) continue;
        const int m = (int)(r / a.B), b = (int)(r - (long long)m * a.B);
        float* orow = a.direct ? a.out + ((long long)b * a.M + m) * a.Fout : a.out + r * a.Fout;
        const float* brow = a.bias_mode == GCNB_BIAS_PER_VERTEX ? a.bias + (long long)m * a.Fout : a.bias;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int o = nt * 8 + 2 * t;
          float v0 = acc[mt][nt][half * 2], v1 = acc[mt][nt][half * 2 + 1];
          if (a.direct) {
            if (a.bias_mode != GCNB_BIAS_NONE) {
              if (o < a.Fout) v0 += brow[o];
              if (o + 1 < a.Fout) v1 += brow[o + 1];
            }
            if (a.relu) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
          }
          if ((a.Fout & 1) == 0) {
            if (o < a.Fout) *reinterpret_cast<float2*>(orow + o) = make_float2(v0, v1);
          } else {
            if (o < a.Fout) orow[o] = v0;
            if (o + 1 < a.Fout) orow[o + 1] = v1;
          }
        }
      }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient from the vertex-major stack
// ------------------------------------------------------------------------------------------------
struct NodeDwArgs {
  const float* Xs;
  long long slab;
  const float* dZ;  // [R][ldz]
  float* part;      // [K][nblocks][Fin][Fout]
  long long R;
  int ldz, Fin, Fout, K, NS, nchunks, chunks_per_cta;
  uint32_t stage_bytes;
};

constexpr int kDwRows = 64;  // rows per stage: two halves of 32, four mma k-steps each

template <int KPW, int MT, int NT>
__global__ void __launch_bounds__(kThreads, 1) k_node_dw(NodeDwArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  Ring ring;
  ring.init(smem, a.NS, tid);
  unsigned char* stages = smem + kRingBytes;
  __syncthreads();
  const int cbeg = blockIdx.x * a.chunks_per_cta;
  const int cend = min(cbeg + a.chunks_per_cta, a.nchunks);
  const uint32_t xtile = (uint32_t)kDwRows * a.Fin * 4;  // bytes of one order's tile inside a stage

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      long long it = 0;
      for (int c = cbeg; c < cend; ++c, ++it) {
        const long long r0 = (long long)c * kDwRows;
        const long long rows = a.R - r0 < kDwRows ? a.R - r0 : kDwRows;
        const uint32_t xb = (uint32_t)(rows * a.Fin * 4), zb = (uint32_t)(rows * a.ldz * 4);
        const int s = ring.acquire(it);
        unsigned char* st = stages + (size_t)s * a.stage_bytes;
        mbar_expect_tx(ring.full + s, xb * a.K + zb);
        for (int k = 0; k < a.K; ++k) bulk_g2s_ring(st + (size_t)k * xtile, a.Xs + (long long)k * a.slab + r0 * a.Fin, xb, ring.full + s);
        bulk_g2s_ring(st + (size_t)a.K * xtile, a.dZ + r0 * a.ldz, zb, ring.full + s);
      }
    }
  } else {
    // consumer warp = (order group kg, row half rh): orders kg, kg+4, ... ; rows [32 rh, 32 rh + 32) of the chunk
    const int kg = warp & 3, rh = warp >> 2;
    float acc[KPW][MT][NT][4];
#pragma unroll
    for (int i = 0; i < KPW; ++i)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[i][mt][nt][c] = 0.f;
    long long it = 0;
    for (int c = cbeg; c < cend; ++c, ++it) {
      const long long r0 = (long long)c * kDwRows;
      const int valid = (int)(a.R - r0 < kDwRows ? a.R - r0 : kDwRows);
      const int s = ring.wait_full(it);
      const unsigned char* st = stages + (size_t)s * a.stage_bytes;
      const float* Z = reinterpret_cast<const float*>(st + (size_t)a.K * xtile);
#pragma unroll 1
      for (int ks = 0; ks < 4; ++ks) {
        const int ra = rh * 32 + ks * 8 + t, rb = ra + 4;  // the two rows this lane feeds into the k dimension
        const bool va = ra < valid, vb = rb < valid;
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int o = nt * 8 + g;
          const float z0 = (va && o < a.Fout) ? Z[ra * a.ldz + o] : 0.f;
          const float z1 = (vb && o < a.Fout) ? Z[rb * a.ldz + o] : 0.f;
          split_trunc(z0, bh[nt][0], bl[nt][0]);
          split_trunc(z1, bh[nt][1], bl[nt][1]);
        }
#pragma unroll
        for (int i = 0; i < KPW; ++i) {
          const int k = kg + 4 * i;
          if (k < a.K) {
            const float* X = reinterpret_cast<const float*>(st + (size_t)k * xtile);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const int f0 = mt * 16 + g, f1 = f0 + 8;
              uint32_t ah[4], al[4];
              split_trunc((va && f0 < a.Fin) ? X[ra * a.Fin + f0] : 0.f, ah[0], al[0]);
              split_trunc((va && f1 < a.Fin) ? X[ra * a.Fin + f1] : 0.f, ah[1], al[1]);
              split_trunc((vb && f0 < a.Fin) ? X[rb * a.Fin + f0] : 0.f, ah[2], al[2]);
              split_trunc((vb && f1 < a.Fin) ? X[rb * a.Fin + f1] : 0.f, ah[3], al[3]);
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) mma_3xtf32(acc[i][mt][nt], ah, al, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
            }
          }
        }
      }
      ring.release(s, lane);
    }
    // the pipeline is drained (every issued stage was consumed): reuse the stage memory to add the two row halves
    consumer_barrier();
    float* red = reinterpret_cast<float*>(stages);
    constexpr int kAcc = KPW * MT * NT * 4;
    if (rh == 1) {
      float* dst = red + ((size_t)kg * kAcc) * 32 + lane;
#pragma unroll
      for (int i = 0; i < KPW; ++i)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[(((i * MT + mt) * NT + nt) * 4 + c) * 32] = acc[i][mt][nt][c];
    }
    consumer_barrier();
    if (rh == 0) {
      const float* src = red + ((size_t)kg * kAcc) * 32 + lane;
#pragma unroll
      for (int i = 0; i < KPW; ++i) {
        const int k = kg + 4 * i;
        if (k >= a.K) continue;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int f = mt * 16 + g + (c >> 1) * 8, o = nt * 8 + 2 * t + (c & 1);
              const float v = acc[i][mt][nt][c] + src[(((i * MT + mt) * NT + nt) * 4 + c) * 32];
              if (f < a.Fin && o < a.Fout)
                a.part[(((long long)k * gridDim.x + blockIdx.x) * a.Fin + f) * a.Fout + o] = v;
            }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// G_k = dZ W_k^T for every order, written as a vertex-major stack Gs[k][r][f]
// ------------------------------------------------------------------------------------------------
struct DzWtArgs {
  const float* dZ;  // [R][ldz]
  const float* W;
  float* Gs;
  long long slab, R;
  int ldz, Fin, Fout, K, NS, ntiles;
};

template <int KC, int NTF>
__global__ void __launch_bounds__(kThreads, 1) k_dz_wt(DzWtArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  Ring ring;
  ring.init(smem, a.NS, tid);
  float2* wfrag = reinterpret_cast<float2*>(smem + kRingBytes);
  const int nw = a.K * KC * NTF * 32;
  // per consumer warp: two output buffers of 32 rows x Fin floats
  const uint32_t obuf = (uint32_t)align_up((size_t)32 * a.Fin * 4, 128);
  unsigned char* outs = smem + kRingBytes + align_up((size_t)nw * 8, 128);
  const uint32_t stage_bytes = (uint32_t)kTileRows * a.ldz * 4;
  unsigned char* stages = outs + (size_t)kConsumerWarps * 2 * obuf;
  for (int i = tid; i < nw; i += kThreads) {
    const int l = i & 31, nt = (i >> 5) % NTF, kc = ((i >> 5) / NTF) % KC, k = (i >> 5) / (NTF * KC);
    const int o0 = kc * 8 + (l & 3), o1 = o0 + 4, f = nt * 8 + (l >> 2);
    float2 v;
    v.x = (f < a.Fin && o0 < a.Fout) ? __ldg(a.W + ((long long)f * a.K + k) * a.Fout + o0) : 0.f;
    v.y = (f < a.Fin && o1 < a.Fout) ? __ldg(a.W + ((long long)f * a.K + k) * a.Fout + o1) : 0.f;
    wfrag[i] = v;
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      long long it = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
        const long long r0 = (long long)tile * kTileRows;
        const long long rows = a.R - r0 < kTileRows ? a.R - r0 : kTileRows;
        const uint32_t bytes = (uint32_t)(rows * a.ldz * 4);
        const int s = ring.acquire(it);
        mbar_expect_tx(ring.full + s, bytes);
        bulk_g2s_ring(stages + (size_t)s * stage_bytes, a.dZ + r0 * a.ldz, bytes, ring.full + s);
      }
    }
    return;
  }

  unsigned char* obase = outs + (size_t)(warp * 2) * obuf;
  long long it = 0;
  int flip = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
    const int s = ring.wait_full(it);
    const float* Z = reinterpret_cast<const float*>(stages + (size_t)s * stage_bytes);
    // the warp's 32 rows of dZ as pre-split A fragments, kept for all K orders
    uint32_t ah[2][KC][4], al[2][KC][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const float* row = Z + (warp * 32 + mt * 16 + g) * a.ldz;
        const int c0 = kc * 8 + t, c1 = c0 + 4;
        split_trunc(c0 < a.Fout ? row[c0] : 0.f, ah[mt][kc][0], al[mt][kc][0]);
        split_trunc(c0 < a.Fout ? row[8 * a.ldz + c0] : 0.f, ah[mt][kc][1], al[mt][kc][1]);
        split_trunc(c1 < a.Fout ? row[c1] : 0.f, ah[mt][kc][2], al[mt][kc][2]);
        split_trunc(c1 < a.Fout ? row[8 * a.ldz + c1] : 0.f, ah[mt][kc][3], al[mt][kc][3]);
      }
    ring.release(s, lane);
    const long long rw = (long long)tile * kTileRows + warp * 32;  // first row of this warp
    if (rw >= a.R) continue;
    const long long wrows = a.R - rw < 32 ? a.R - rw : 32;
    for (int k = 0; k < a.K; ++k) {
      float acc[2][NTF][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTF; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const float2* wf = wfrag + ((long long)(k * KC + kc) * NTF) * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < NTF; ++nt) {
          const float2 b = wf[nt * 32];
          uint32_t bh0, bl0, bh1, bl1;
          split_trunc(b.x, bh0, bl0);
          split_trunc(b.y, bh1, bl1);
          mma_3xtf32(acc[0][nt], ah[0][kc], al[0][kc], bh0, bh1, bl0, bl1);
          mma_3xtf32(acc[1][nt], ah[1][kc], al[1][kc], bh0, bh1, bl0, bl1);
        }
      }
      // stage the [32 x Fin] block in the warp's buffer, then one TMA store; the buffer used two orders ago
      // must have been read out by then
      float* o = reinterpret_cast<float*>(obase + (size_t)flip * obuf);
      if (lane == 0) bulk_wait_read<1>();
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTF; ++nt)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int row = mt * 16 + g + (c >> 1) * 8, f = nt * 8 + 2 * t + (c & 1);
            if (f < a.Fin) o[row * a.Fin + f] = acc[mt][nt][c];
          }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) bulk_s2g(a.Gs + (long long)k * a.slab + rw * a.Fin, o, (uint32_t)(wrows * a.Fin * 4));
      flip ^= 1;
    }
  }
  if (lane == 0) bulk_wait_all();
}

int max_dyn_smem() {
  DeviceInfo di;
  if (device_info(&di) != GCNB_OK) return 0;
  return di.smem_optin;
}
int sm_count() {
  DeviceInfo di;
  if (device_info(&di) != GCNB_OK) return 148;
  return di.sm_count;
}

template <typename Kern>
int set_smem(Kern kern, size_t bytes) {
  GCNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return GCNB_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Common precondition of the three kernels: every tile span must be a legal bulk copy.
bool node_mma_layout_ok(long long R, int Fin, long long slab) {
  return (R * Fin) % 4 == 0 && slab % 4 == 0 && Fin >= 1 && Fin <= 64;
}

int node_mma_ldz(int Fout) { return (Fout + 3) / 4 * 4 + 4; }

// ---- forward contraction ------------------------------------------------------------------------
static bool contract_plan(const LayerShape& s, int* NT, int* NS, size_t* smem) {
  if (s.Fout > 64 || s.K < 1) return false;
  *NT = s.Fout <= 8 ? 1 : s.Fout <= 16 ? 2 : s.Fout <= 32 ? 4 : 8;
  const int KC = ceil_div(s.Fin, 8);
  const size_t fixed = kRingBytes + align_up((size_t)s.K * KC * *NT * 32 * 8, 128);
  const size_t stage = (size_t)kTileRows * s.Fin * 4;
  const int cap = max_dyn_smem();
  if ((size_t)cap < fixed + 2 * stage) return false;
  *NS = (int)std::min<size_t>(kMaxStages, ((size_t)cap - fixed) / stage);
  *smem = fixed + (size_t)*NS * stage;
  return true;
}

bool node_contract_supported(const LayerShape& s) {
  int NT, NS;
  size_t smem;
  const long long R = (long long)s.M * s.B;
  return node_mma_layout_ok(R, s.Fin, R * s.Fin) && contract_plan(s, &NT, &NS, &smem);
}

// out = y (sample-major, bias/ReLU applied) when direct != 0, else Zn[r][Fout]
int node_contract(const float* Xs, long long slab, const float* W, const float* bias, float* out, const LayerShape& s,
                  int bias_mode, int relu, int direct, cudaStream_t st) {
  int NT, NS;
  size_t smem;
  if (!contract_plan(s, &NT, &NS, &smem) || !aligned16(Xs)) {
    set_error("node_contract: unsupported shape");
    return GCNB_ERR_INVALID;
  }
  ContractArgs a;
  a.Xs = Xs; a.slab = slab; a.W = W; a.bias = bias; a.out = out;
  a.R = (long long)s.M * s.B; a.B = s.B; a.M = s.M; a.Fin = s.Fin; a.Fout = s.Fout; a.K = s.K;
  a.KC = ceil_div(s.Fin, 8); a.bias_mode = bias_mode; a.relu = relu; a.direct = direct; a.NS = NS;
  a.ntiles = (int)ceil_div_ll(a.R, kTileRows);
  const int grid = std::min(a.ntiles, sm_count());
#define GCNB_CONTRACT(NTV)                                                   \
  {                                                                          \
    int rc = set_smem(k_stack_contract<NTV>, smem);                          \
    if (rc) return rc;                                                       \
    k_stack_contract<NTV><<<grid, kThreads, smem, st>>>(a);                  \
  }
  switch (NT) {
    case 1: GCNB_CONTRACT(1) break;
    case 2: GCNB_CONTRACT(2) break;
    case 4: GCNB_CONTRACT(4) break;
    default: GCNB_CONTRACT(8) break;
  }
#undef GCNB_CONTRACT
  GCNB_LAUNCH_CHECK("k_stack_contract");
  return GCNB_OK;
}

// ---- weight gradient ------------------------------------------------------------------------------
static bool node_dw_plan(const LayerShape& s, int* KPW, int* MT, int* NT, int* NS, size_t* stage, size_t* smem) {
  if (s.Fout > 64 || s.Fin > 32) return false;
  *MT = s.Fin <= 16 ? 1 : 2;
  *NT = s.Fout <= 32 ? 4 : 8;
  *KPW = s.K <= 8 ? 2 : s.K <= 16 ? 4 : s.K <= 28 ? 7 : 0;
  if (*KPW == 0 || *KPW * *MT * *NT > 32) return false;
  const int ldz = node_mma_ldz(s.Fout);
  *stage = align_up((size_t)s.K * kDwRows * s.Fin * 4 + (size_t)kDwRows * ldz * 4, 128);
  const int cap = max_dyn_smem();
  const size_t red = (size_t)4 * *KPW * *MT * *NT * 4 * 32 * 4;  // the final reduction reuses the stages
  if ((size_t)cap < kRingBytes + 2 * *stage || 2 * *stage < red) return false;
  *NS = (int)std::min<size_t>(4, ((size_t)cap - kRingBytes) / *stage);
  *smem = kRingBytes + (size_t)*NS * *stage;
  return true;
}

bool node_dw_supported(const LayerShape& s) {
  int KPW, MT, NT, NS;
  size_t stage, smem;
  const long long R = (long long)s.M * s.B;
  return node_mma_layout_ok(R, s.Fin, R * s.Fin) && node_dw_plan(s, &KPW, &MT, &NT, &NS, &stage, &smem);
}

int node_dw_blocks(const LayerShape& s) {
  const int nchunks = (int)ceil_div_ll((long long)s.M * s.B, kDwRows);
  const int per = ceil_div(nchunks, sm_count());
  return ceil_div(nchunks, per);
}

// part[K][nblocks][Fin][Fout]; the caller reduces over the blocks (k_dw_reduce)
int node_dw(const float* Xs, long long slab, const float* dZ, float* part, const LayerShape& s, cudaStream_t st) {
  int KPW, MT, NT, NS;
  size_t stage, smem;
  if (!node_dw_plan(s, &KPW, &MT, &NT, &NS, &stage, &smem) || !aligned16(Xs) || !aligned16(dZ)) {
    set_error("node_dw: unsupported shape");
    return GCNB_ERR_INVALID;
  }
  NodeDwArgs a;
  a.Xs = Xs; a.slab = slab; a.dZ = dZ; a.part = part; a.R = (long long)s.M * s.B; a.ldz = node_mma_ldz(s.Fout);
  a.Fin = s.Fin; a.Fout = s.Fout; a.K = s.K; a.NS = NS;
  a.nchunks = (int)ceil_div_ll(a.R, kDwRows);
  a.chunks_per_cta = ceil_div(a.nchunks, sm_count());
  a.stage_bytes = (uint32_t)stage;
  const int grid = ceil_div(a.nchunks, a.chunks_per_cta);
#define GCNB_NODE_DW(KP, M_, N_)                                             \
  if (KPW == KP && MT == M_ && NT == N_) {                                   \
    int rc = set_smem(k_node_dw<KP, M_, N_>, smem);                          \
    if (rc) return rc;                                                       \
    k_node_dw<KP, M_, N_><<<grid, kThreads, smem, st>>>(a);                  \
  } else
  GCNB_NODE_DW(2, 1, 4)
  GCNB_NODE_DW(2, 2, 4)
  GCNB_NODE_DW(2, 1, 8)
  GCNB_NODE_DW(2, 2, 8)
  GCNB_NODE_DW(4, 1, 4)
  GCNB_NODE_DW(4, 2, 4)
  GCNB_NODE_DW(4, 1, 8)
  GCNB_NODE_DW(7, 1, 4) {
    set_error("node_dw: no kernel for KPW=%d MT=%d NT=%d", KPW, MT, NT);
    return GCNB_ERR_INVALID;
  }
#undef GCNB_NODE_DW
  GCNB_LAUNCH_CHECK("k_node_dw");
  return GCNB_OK;
}

// ---- seeds of the adjoint recursion -------------------------------------------------------------
static bool dz_wt_plan(const LayerShape& s, int* KC, int* NTF, int* NS, size_t* smem) {
  if (s.Fout > 32 || s.Fin > 32) return false;
  *KC = s.Fout <= 16 ? 2 : 4;
  *NTF = s.Fin <= 8 ? 1 : s.Fin <= 16 ? 2 : 4;
  const int ldz = node_mma_ldz(s.Fout);
  const size_t fixed = kRingBytes + align_up((size_t)s.K * *KC * *NTF * 32 * 8, 128) +
                       (size_t)kConsumerWarps * 2 * align_up((size_t)32 * s.Fin * 4, 128);
  const size_t stage = (size_t)kTileRows * ldz * 4;
  const int cap = max_dyn_smem();
  if ((size_t)cap < fixed + 2 * stage) return false;
  *NS = (int)std::min<size_t>(4, ((size_t)cap - fixed) / stage);
  *smem = fixed + (size_t)*NS * stage;
  return true;
}

bool node_dz_wt_supported(const LayerShape& s) {
  int KC, NTF, NS;
  size_t smem;
  const long long R = (long long)s.M * s.B;
  return node_mma_layout_ok(R, s.Fin, R * s.Fin) && dz_wt_plan(s, &KC, &NTF, &NS, &smem);
}

int node_dz_wt(const float* dZ, const float* W, float* Gs, long long slab, const LayerShape& s, cudaStream_t st) {
  int KC, NTF, NS;
  size_t smem;
  if (!dz_wt_plan(s, &KC, &NTF, &NS, &smem) || !aligned16(dZ) || !aligned16(Gs)) {
    set_error("node_dz_wt: unsupported shape");
    return GCNB_ERR_INVALID;
  }
  DzWtArgs a;
  a.dZ = dZ; a.W = W; a.Gs = Gs; a.slab = slab; a.R = (long long)s.M * s.B; a.ldz = node_mma_ldz(s.Fout);
  a.Fin = s.Fin; a.Fout = s.Fout; a.K = s.K; a.NS = NS; a.ntiles = (int)ceil_div_ll(a.R, kTileRows);
  const int grid = std::min(a.ntiles, sm_count());
#define GCNB_DZ_WT(KCV, NTV)                                                 \
  if (KC == KCV && NTF == NTV) {                                             \
    int rc = set_smem(k_dz_wt<KCV, NTV>, smem);                              \
    if (rc) return rc;                                                       \
    k_dz_wt<KCV, NTV><<<grid, kThreads, smem, st>>>(a);                      \
  } else
  GCNB_DZ_WT(2, 1)
  GCNB_DZ_WT(2, 2)
  GCNB_DZ_WT(2, 4)
  GCNB_DZ_WT(4, 1)
  GCNB_DZ_WT(4, 2)
  GCNB_DZ_WT(4, 4) {
    set_error("node_dz_wt: no kernel for KC=%d NTF=%d", KC, NTF);
    return GCNB_ERR_INVALID;
  }
#undef GCNB_DZ_WT
  GCNB_LAUNCH_CHECK("k_dz_wt");
  return GCNB_OK;
}

}  // namespace gcnb
