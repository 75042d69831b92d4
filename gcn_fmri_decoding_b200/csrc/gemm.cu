// Dense fp32 GEMM on the tensor cores with fp32-level accuracy (3-pass error-compensated TF32, mma.sync m16n8k8).
//
//   C[M x N] = op(A)[M x K] * op(B)[K x N] (+ bias[N] broadcast over rows),   row-major, arbitrary leading dimensions
//
// Used for the two dense graph-Fourier transforms of the spectral layer (models_gcn.py:512-528: [M,M] x [M, B*F]) and
// for the small FC GEMMs of the training step, where a single-pass TF32 GEMM would break the 1e-4 parity bar and the
// SIMT sgemm is several times slower.  64x64x16 CTA tiles, 4 warps (2x2), 32x32 per warp, double-buffered shared
// memory filled with guarded loads (any M, N, K, any alignment, either operand transposed).
#include <algorithm>

#include <cooperative_groups.h>

#include "fused_common.cuh"

namespace gcnb {

constexpr int GBM = 64, GBN = 64, GBK = 16;
constexpr int GAS = GBK + 4;  // A tile row stride: 20 -> fragment rows hit distinct banks
constexpr int GBS = GBN + 8;  // B tile row stride: 72 -> 8t + g distinct

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int M, N, K, lda, ldb, ldc, ta, tb;
  // fused epilogue (after the bias): GCNB_EPI_RELU_DROPOUT: v = dropout(relu(v)) with the mask of k_relu_dropout_fwd;
  // GCNB_EPI_MASK: v = aux[m][n] > 0 ? v / keep : 0  (the adjoint of the former, aux = the forward activation)
  int epi;
  const float* aux;
  int ld_aux;
  float keep, inv_keep;
  unsigned seed;
  const float* step;
  int splits;  // small kernel: K is split over a thread-block cluster of this many CTAs (1 = no split)
};

struct Epilogue {
  uint32_t key, thresh;
  __device__ __forceinline__ explicit Epilogue(const GemmArgs& g) : key(0), thresh(0) {
    if (g.epi == GCNB_EPI_RELU_DROPOUT) {
      key = dropout_key(g.seed, g.step);
      thresh = dropout_threshold(g.keep);
    }
  }
  __device__ __forceinline__ float operator()(const GemmArgs& g, float v, int m, int n) const {
    if (g.epi == GCNB_EPI_RELU_DROPOUT) {
      v = fmaxf(v, 0.f);
      if (g.keep < 1.f) v = dropout_keeps((uint32_t)((long long)m * g.N + n), key, thresh) ? v * g.inv_keep : 0.f;
    } else if (g.epi == GCNB_EPI_MASK) {
      v = __ldg(g.aux + (long long)m * g.ld_aux + n) > 0.f ? v * g.inv_keep : 0.f;
    }
    return v;
  }
};

__device__ __forceinline__ float gemm_a(const GemmArgs& g, int m, int k) {
  if (m >= g.M || k >= g.K) return 0.f;
  return g.ta ? __ldg(g.A + (long long)k * g.lda + m) : __ldg(g.A + (long long)m * g.lda + k);
}

__device__ __forceinline__ float gemm_b(const GemmArgs& g, int k, int n) {
  if (k >= g.K || n >= g.N) return 0.f;
  return g.tb ? __ldg(g.B + (long long)n * g.ldb + k) : __ldg(g.B + (long long)k * g.ldb + n);
}

__global__ void __launch_bounds__(128) k_gemm_3xtf32(const GemmArgs g) {
  __shared__ float As[2][GBM * GAS];
  __shared__ float Bs[2][GBK * GBS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  // register staging of the next tile (8 A + 8 B values per thread); consecutive threads walk the operand's
  // contiguous axis
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 128;  // 0..1023
      int m, k;
      if (g.ta) { k = idx >> 6; m = idx & 63; } else { m = idx >> 4; k = idx & 15; }
      ra[i] = gemm_a(g, m0 + m, k0 + k);
      int kb, n;
      if (g.tb) { n = idx >> 4; kb = idx & 15; } else { kb = idx >> 6; n = idx & 63; }
      rb[i] = gemm_b(g, k0 + kb, n0 + n);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 128;
      int m, k;
      if (g.ta) { k = idx >> 6; m = idx & 63; } else { m = idx >> 4; k = idx & 15; }
      As[buf][m * GAS + k] = ra[i];
      int kb, n;
      if (g.tb) { n = idx >> 4; kb = idx & 15; } else { kb = idx >> 6; n = idx & 63; }
      Bs[buf][kb * GBS + n] = rb[i];
    }
  };

  const int nk = (g.K + GBK - 1) / GBK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * GBK);  // global loads in flight while the tensor cores work on `buf`
    const float* as = As[buf];
    const float* bs = Bs[buf];
#pragma unroll
    for (int ks = 0; ks < GBK; ks += 8) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float* p = as + (wm + i * 16 + gq) * GAS + ks + t;
        split_trunc(p[0], ah[i][0], al[i][0]);
        split_trunc(p[8 * GAS], ah[i][1], al[i][1]);
        split_trunc(p[4], ah[i][2], al[i][2]);
        split_trunc(p[8 * GAS + 4], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* p = bs + (ks + t) * GBS + wn + j * 8 + gq;
        uint32_t bh0, bl0, bh1, bl1;
        split_trunc(p[0], bh0, bl0);
        split_trunc(p[4 * GBS], bh1, bl1);
#pragma unroll
        for (int i = 0; i < 2; ++i) mma_3xtf32(acc[i][j], ah[i], al[i], bh0, bh1, bl0, bl1);
      }
    }
    if (kt + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }

  const Epilogue epi(g);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + wn + j * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = m0 + wm + i * 16 + gq + 8 * h;
        if (m >= g.M) continue;
        float v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
        if (g.bias != nullptr) {
          if (n < g.N) v0 += __ldg(g.bias + n);
          if (n + 1 < g.N) v1 += __ldg(g.bias + n + 1);
        }
        float* c = g.C + (long long)m * g.ldc + n;
        if (n < g.N) c[0] = epi(g, v0, m, n);
        if (n + 1 < g.N) c[1] = epi(g, v1, m, n + 1);
      }
    }
}

// ---- small / skinny problems (the FC layers at B = 512): 32x32x32 tiles so that even a 512x256 output gives 128 CTAs,
// and a 5-stage cp.async pipeline (4-byte copies: any alignment, any transpose, zero-fill out of range) so that the
// serial walk over K is not exposed to global-memory latency.
constexpr int SBM = 32, SBN = 32, SBK = 32, SST = 5;  // 5 stages: 48.6 KB static smem, covers L2 latency on the serial walk over K
constexpr int SAS = SBK + 4;  // 36: A-fragment rows on distinct banks
constexpr int SBS = SBN + 8;  // 40: 8t + g distinct

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool ok) {
  const int n = ok ? 4 : 0;  // src-size 0 => the 4 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(128) k_gemm_small_3xtf32(const GemmArgs g) {
  __shared__ float As[SST][SBM * SAS];
  __shared__ float Bs[SST][SBK * SBS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  float acc[2][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;

  // split-K over the CTAs of a cluster (blockIdx.z = rank): each takes a contiguous range of k-tiles
  const int nk_all = (g.K + SBK - 1) / SBK;
  const int per = (nk_all + g.splits - 1) / g.splits;
  const int kt0 = (int)blockIdx.z * per;
  const int nk = max(0, min(nk_all, kt0 + per) - kt0);
  auto issue = [&](int kt, int st) {
    const int k0 = (kt0 + kt) * SBK;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 128;  // 0..1023
      int m, k;
      if (g.ta) { k = idx >> 5; m = idx & 31; } else { m = idx >> 5; k = idx & 31; }
      const bool oka = (m0 + m) < g.M && (k0 + k) < g.K;
      const float* pa = g.ta ? g.A + (long long)(k0 + k) * g.lda + m0 + m : g.A + (long long)(m0 + m) * g.lda + k0 + k;
      cp_async4(&As[st][m * SAS + k], oka ? pa : g.A, oka);
      int kb, n;
      if (g.tb) { n = idx >> 5; kb = idx & 31; } else { kb = idx >> 5; n = idx & 31; }
      const bool okb = (k0 + kb) < g.K && (n0 + n) < g.N;
      const float* pb = g.tb ? g.B + (long long)(n0 + n) * g.ldb + k0 + kb : g.B + (long long)(k0 + kb) * g.ldb + n0 + n;
      cp_async4(&Bs[st][kb * SBS + n], okb ? pb : g.B, okb);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (int s = 0; s < SST - 1; ++s) {
    if (s < nk) issue(s, s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(SST - 2) : "memory");
    __syncthreads();  // tile kt landed for everyone; tile kt-1's buffer is free for the prefetch below
    if (kt + SST - 1 < nk) issue(kt + SST - 1, (kt + SST - 1) % SST);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const float* as = As[kt % SST];
    const float* bs = Bs[kt % SST];
#pragma unroll
    for (int ks = 0; ks < SBK; ks += 8) {
      uint32_t ah[4], al[4];
      const float* p = as + (wm + gq) * SAS + ks + t;
      split_trunc(p[0], ah[0], al[0]);
      split_trunc(p[8 * SAS], ah[1], al[1]);
      split_trunc(p[4], ah[2], al[2]);
      split_trunc(p[8 * SAS + 4], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float* q = bs + (ks + t) * SBS + wn + j * 8 + gq;
        uint32_t bh0, bl0, bh1, bl1;
        split_trunc(q[0], bh0, bl0);
        split_trunc(q[4 * SBS], bh1, bl1);
        mma_3xtf32(acc[j], ah, al, bh0, bh1, bl0, bl1);
      }
    }
  }
  if (g.splits > 1) {
    // Deterministic split-K reduction through distributed shared memory: every rank but 0 parks its partial tile in
    // its own shared memory, rank 0 adds them in rank order and runs the epilogue.
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // the pipeline buffers are dead: reuse the first one for the partial tile
    float* part = &As[0][0];
    const unsigned rank = cluster.block_rank();
    if (rank != 0) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) part[(j * 4 + c) * 128 + tid] = acc[j][c];
    }
    cluster.sync();
    if (rank == 0) {
      for (unsigned r = 1; r < cluster.num_blocks(); ++r) {
        const float* rp = cluster.map_shared_rank(part, r);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[j][c] += rp[(j * 4 + c) * 128 + tid];
      }
    }
    cluster.sync();  // the partials must stay alive until rank 0 has read them
    if (rank != 0) return;
  }
  const Epilogue epi(g);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int n = n0 + wn + j * 8 + 2 * t;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + wm + gq + 8 * h;
      if (m >= g.M) continue;
      float v0 = acc[j][2 * h], v1 = acc[j][2 * h + 1];
      if (g.bias != nullptr) {
        if (n < g.N) v0 += __ldg(g.bias + n);
        if (n + 1 < g.N) v1 += __ldg(g.bias + n + 1);
      }
      float* c = g.C + (long long)m * g.ldc + n;
      if (n < g.N) c[0] = epi(g, v0, m, n);
      if (n + 1 < g.N) c[1] = epi(g, v1, m, n + 1);
    }
  }
}

int launch_gemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb, int ldc,
                int ta, int tb, cudaStream_t st) {
  return launch_gemm_epi(A, B, C, bias, M, N, K, lda, ldb, ldc, ta, tb, GCNB_EPI_NONE, nullptr, 0, 1.f, 0u, nullptr, st);
}

int launch_gemm_epi(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                    int ldc, int ta, int tb, int epilogue, const float* aux, int ld_aux, float keep, unsigned seed,
                    const float* step, cudaStream_t st) {
  GemmArgs g{A, B, C, bias, M, N, K, lda, ldb, ldc, ta, tb, epilogue, aux, ld_aux, keep, keep >= 1.f ? 1.f : 1.f / keep,
             seed, step, 1};
  g.splits = 1;
  if ((long long)ceil_div(N, GBN) * ceil_div(M, GBM) < 2 * 148) {  // not enough 64x64 tiles to fill the chip
    dim3 grid((unsigned)ceil_div(N, SBN), (unsigned)ceil_div(M, SBM));
    // A handful of output tiles with a long K (weight gradients of the FC layers: [25 x 512], [256 x 22] over K = 512)
    // leaves most SMs idle behind a serial walk over K: split K over a cluster of up to 8 CTAs.
    const int ctas = (int)(grid.x * grid.y), nk = ceil_div(K, SBK);
    int splits = 1;
    while (splits < 8 && splits * 2 <= nk / 2 && ctas * splits * 2 <= 148) splits *= 2;
    if (splits > 1) {
      g.splits = splits;
      grid.z = splits;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = grid;
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 1;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = splits;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      GCNB_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_small_3xtf32, g));
      GCNB_LAUNCH_CHECK("k_gemm_small_3xtf32 (cluster split-K)");
      return GCNB_OK;
    }
    k_gemm_small_3xtf32<<<grid, 128, 0, st>>>(g);
    GCNB_LAUNCH_CHECK("k_gemm_small_3xtf32");
    return GCNB_OK;
  }
  dim3 grid((unsigned)ceil_div(N, GBN), (unsigned)ceil_div(M, GBM));
  k_gemm_3xtf32<<<grid, 128, 0, st>>>(g);
  GCNB_LAUNCH_CHECK("k_gemm_3xtf32");
  return GCNB_OK;
}

}  // namespace gcnb

extern "C" int gcnb_gemm_f32(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda,
                             int ldb, int ldc, int transA, int transB, gcnb_stream_t stream) {
  GCNB_REQUIRE(A && B && C && M >= 1 && N >= 1 && K >= 1, "gcnb_gemm_f32: bad arguments");
  GCNB_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "gcnb_gemm_f32: leading dimension too small");
  return gcnb::launch_gemm(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, static_cast<cudaStream_t>(stream));
}

extern "C" int gcnb_gemm_epilogue_f32(const float* A, const float* B, float* C, const float* bias, int M, int N, int K,
                                      int lda, int ldb, int ldc, int transA, int transB, int epilogue, const float* aux,
                                      int ld_aux, float keep, unsigned seed, const float* step, gcnb_stream_t stream) {
  GCNB_REQUIRE(A && B && C && M >= 1 && N >= 1 && K >= 1, "gcnb_gemm_epilogue_f32: bad arguments");
  GCNB_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N,
               "gcnb_gemm_epilogue_f32: leading dimension too small");
  GCNB_REQUIRE(epilogue >= GCNB_EPI_NONE && epilogue <= GCNB_EPI_MASK, "gcnb_gemm_epilogue_f32: bad epilogue %d", epilogue);
  GCNB_REQUIRE(keep > 0.f && keep <= 1.f, "gcnb_gemm_epilogue_f32: keep must be in (0, 1]");
  GCNB_REQUIRE(epilogue != GCNB_EPI_MASK || (aux != nullptr && ld_aux >= N), "gcnb_gemm_epilogue_f32: the mask epilogue needs aux");
  return gcnb::launch_gemm_epi(A, B, C, bias, M, N, K, lda, ldb, ldc, transA, transB, epilogue, aux, ld_aux, keep, seed, step,
                               static_cast<cudaStream_t>(stream));
}
