// Spectral (graph-Fourier) layer: cgcnn.fourier / filter_in_fourier (models_gcn.py:512-539).
//
//   xh = Ut x          (graph Fourier transform, dense [M,M] x [M, B*Fin])
//   yh[m] = W[m] xh[m] (one Fout x Fin mix per graph frequency m)
//   z  = Ut^T yh       (inverse transform), then the shared bias/ReLU/max-pool epilogue.
//
// Everything runs in the vertex-major layout Xn[m][b*F + f] so that both transforms are plain
// row-major GEMMs over all samples at once (no per-sample batching, no transposes of the data
// besides the one pass in and one pass out).  Backward recomputes xh.
// The two dense transforms run on the tensor cores (3-pass TF32 split, gemm.cu) -- SURVEY 8a row a3.
#include <algorithm>

#include "common.cuh"

namespace gcnb {

// C[Mr x N] = op(A)[Mr x Kd] * Bm[Kd x N];  A row-major with leading dimension lda;
// TA: op(A)[i][k] = A[k][i].  64x64x16 tiles, 256 threads, 4x4 outputs per thread.
template <bool TA>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, const float* __restrict__ Bm,
                                               float* __restrict__ C, int Mr, long long N, int Kd, int lda) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.y * 64;
  const long long col0 = (long long)blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < Kd; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * 256;
      int i, kk;
      if (TA) { kk = idx >> 6; i = idx & 63; } else { i = idx >> 4; kk = idx & 15; }
      float v = 0.f;
      if (row0 + i < Mr && k0 + kk < Kd)
        v = TA ? __ldg(A + (long long)(k0 + kk) * lda + row0 + i) : __ldg(A + (long long)(row0 + i) * lda + k0 + kk);
      As[kk][i] = v;
      const int kb = idx >> 6, j = idx & 63;
      float w = 0.f;
      if (k0 + kb < Kd && col0 + j < N) w = __ldg(Bm + (long long)(k0 + kb) * N + col0 + j);
      Bs[kb][j] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= Mr) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long c = col0 + tx * 4 + j;
      if (c < N) C[(long long)r * N + c] = acc[i][j];
    }
  }
}

// The two graph-Fourier transforms run on the tensor cores (3xTF32, gemm.cu) when the column count fits an int;
// the FFMA kernel above is the fallback for gigantic batches.
static int launch_sgemm(bool ta, const float* A, const float* Bm, float* C, int Mr, long long N, int Kd, int lda,
                        cudaStream_t st) {
  if (N < (1LL << 30)) return launch_gemm(A, Bm, C, nullptr, Mr, (int)N, Kd, lda, (int)N, (int)N, ta ? 1 : 0, 0, st);
  dim3 grid((unsigned)ceil_div_ll(N, 64), (unsigned)ceil_div(Mr, 64));
  if (ta)
    k_sgemm<true><<<grid, 256, 0, st>>>(A, Bm, C, Mr, N, Kd, lda);
  else
    k_sgemm<false><<<grid, 256, 0, st>>>(A, Bm, C, Mr, N, Kd, lda);
  GCNB_LAUNCH_CHECK("k_sgemm");
  return GCNB_OK;
}

// out[m][b*No + n] = sum_k w(m, n, k) * in[m][b*Ki + k]
//   forward  (TR=false): n = o, k = f, w = W[m][o][f]
//   backward (TR=true) : n = f, k = o, w = W[m][o][f]
template <bool TR>
__global__ void __launch_bounds__(256) k_node_mix(const float* __restrict__ W, const float* __restrict__ in,
                                                  float* __restrict__ out, int B, int Fin, int Fout) {
  extern __shared__ float Ws[];  // [No][Ki + 1]
  const int No = TR ? Fin : Fout, Ki = TR ? Fout : Fin;
  const int m = blockIdx.x;
  const float* Wm = W + (long long)m * Fout * Fin;
  for (int idx = threadIdx.x; idx < Fout * Fin; idx += blockDim.x) {
    const int o = idx / Fin, f = idx - o * Fin;
    if (TR) Ws[f * (Ki + 1) + o] = Wm[idx]; else Ws[o * (Ki + 1) + f] = Wm[idx];
  }
  __syncthreads();
  const float* row = in + (long long)m * B * Ki;
  float* orow = out + (long long)m * B * No;
  for (int idx = threadIdx.x; idx < B * No; idx += blockDim.x) {
    const int b = idx / No, n = idx - b * No;
    const float* xin = row + (long long)b * Ki;
    const float* w = Ws + n * (Ki + 1);
    float s = 0.f;
    for (int k = 0; k < Ki; ++k) s = fmaf(w[k], xin[k], s);
    orow[idx] = s;
  }
}

// dW[m][o][f] = sum_b dYh[m][b*Fout+o] * Xh[m][b*Fin+f]   (fixed order over b)
__global__ void __launch_bounds__(256) k_node_dw(const float* __restrict__ dYh, const float* __restrict__ Xh,
                                                 float* __restrict__ dW, int B, int Fin, int Fout) {
  const int m = blockIdx.x;
  const float* d = dYh + (long long)m * B * Fout;
  const float* x = Xh + (long long)m * B * Fin;
  for (int idx = threadIdx.x; idx < Fout * Fin; idx += blockDim.x) {
    const int o = idx / Fin, f = idx - o * Fin;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(d[(long long)b * Fout + o], x[(long long)b * Fin + f], s);
    dW[(long long)m * Fout * Fin + idx] = s;
  }
}

}  // namespace gcnb

using namespace gcnb;

extern "C" {

size_t gcnb_spectral_workspace_bytes(int B, int M, int Fin, int Fout, int p, int backward) {
  (void)p;
  const size_t a = (size_t)M * B * Fin * sizeof(float), c = (size_t)M * B * Fout * sizeof(float);
  size_t n = 2 * align_up(a, 256) + 2 * align_up(c, 256);
  if (backward) n += align_up(db_scratch_floats(M, Fout) * sizeof(float), 256);
  return n + 256;
}

static int check_spectral(const char* who, int B, int M, int Fin, int Fout, int p, int bias_mode) {
  GCNB_REQUIRE(B >= 1 && M >= 1 && Fin >= 1 && Fout >= 1, "%s: bad sizes", who);
  GCNB_REQUIRE(p >= 1 && (p & (p - 1)) == 0 && p <= 128, "%s: pooling size must be a power of 2 in [1,128]", who);
  GCNB_REQUIRE(bias_mode >= 0 && bias_mode <= 2, "%s: bad bias_mode", who);
  GCNB_REQUIRE((size_t)(Fin + 1) * (Fout + 1) * sizeof(float) <= 48 * 1024,
               "%s: Fin*Fout too large for the per-frequency mix kernel", who);
  return GCNB_OK;
}

int gcnb_spectral_fwd_f32(const float* x, const float* Ut, const float* W, const float* bias, float* y,
                          uint8_t* argmax, int B, int M, int Fin, int Fout, int p, int bias_mode, int relu,
                          void* workspace, size_t workspace_bytes, gcnb_stream_t stream) {
  int rc = check_spectral("gcnb_spectral_fwd_f32", B, M, Fin, Fout, p, bias_mode);
  if (rc) return rc;
  GCNB_REQUIRE(x && Ut && W && y, "gcnb_spectral_fwd_f32: NULL argument");
  GCNB_REQUIRE(bias_mode == GCNB_BIAS_NONE || bias, "gcnb_spectral_fwd_f32: bias is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws(workspace, workspace_bytes);
  const size_t na = (size_t)M * B * Fin, nc = (size_t)M * B * Fout;
  float* X0 = ws.take<float>(na);
  float* Xh = ws.take<float>(na);
  float* Yh = ws.take<float>(nc);
  float* Zn = ws.take<float>(nc);
  if (!X0 || !Xh || !Yh || !Zn) {
    set_error("gcnb_spectral_fwd_f32: workspace too small");
    return GCNB_ERR_WORKSPACE;
  }
  if ((rc = launch_to_node_major(x, nullptr, X0, B, M, M, Fin, st))) return rc;
  if ((rc = launch_sgemm(false, Ut, X0, Xh, M, (long long)B * Fin, M, M, st))) return rc;
  k_node_mix<false><<<M, 256, (size_t)Fout * (Fin + 1) * sizeof(float), st>>>(W, Xh, Yh, B, Fin, Fout);
  GCNB_LAUNCH_CHECK("k_node_mix");
  if ((rc = launch_sgemm(true, Ut, Yh, Zn, M, (long long)B * Fout, M, M, st))) return rc;
  return launch_epilogue(Zn, bias, y, argmax, B, M, Fout, p, bias_mode, relu, st);
}

int gcnb_spectral_bwd_f32(const float* x, const float* y, const uint8_t* argmax, const float* dy, const float* Ut,
                          const float* W, float* dx, float* dW, float* db, int B, int M, int Fin, int Fout, int p,
                          int bias_mode, int relu, void* workspace, size_t workspace_bytes, gcnb_stream_t stream) {
  int rc = check_spectral("gcnb_spectral_bwd_f32", B, M, Fin, Fout, p, bias_mode);
  if (rc) return rc;
  GCNB_REQUIRE(x && y && dy && Ut && W && dW, "gcnb_spectral_bwd_f32: NULL argument");
  GCNB_REQUIRE(p == 1 || argmax, "gcnb_spectral_bwd_f32: argmax is required when p > 1");
  GCNB_REQUIRE(bias_mode == GCNB_BIAS_NONE || db, "gcnb_spectral_bwd_f32: db is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Workspace ws(workspace, workspace_bytes);
  const size_t na = (size_t)M * B * Fin, nc = (size_t)M * B * Fout;
  float* bufA = ws.take<float>(na);  // X0, later dXh
  float* bufB = ws.take<float>(na);  // Xh, later dXn
  float* dZn = ws.take<float>(nc);
  float* dYh = ws.take<float>(nc);
  float* dbs = ws.take<float>(db_scratch_floats(M, Fout));
  if (!bufA || !bufB || !dZn || !dYh || !dbs) {
    set_error("gcnb_spectral_bwd_f32: workspace too small");
    return GCNB_ERR_WORKSPACE;
  }
  if ((rc = launch_dz(dy, y, argmax, dZn, Fout, B, M, Fout, p, relu, st))) return rc;
  if ((rc = launch_db(dZn, Fout, db, dbs, B, M, Fout, bias_mode, st))) return rc;
  if ((rc = launch_to_node_major(x, nullptr, bufA, B, M, M, Fin, st))) return rc;
  if ((rc = launch_sgemm(false, Ut, bufA, bufB, M, (long long)B * Fin, M, M, st))) return rc;   // xh
  if ((rc = launch_sgemm(false, Ut, dZn, dYh, M, (long long)B * Fout, M, M, st))) return rc;    // dyh = Ut dz
  k_node_dw<<<M, 256, 0, st>>>(dYh, bufB, dW, B, Fin, Fout);
  GCNB_LAUNCH_CHECK("k_node_dw");
  if (dx == nullptr) return GCNB_OK;
  k_node_mix<true><<<M, 256, (size_t)Fin * (Fout + 1) * sizeof(float), st>>>(W, dYh, bufA, B, Fin, Fout);  // dxh
  GCNB_LAUNCH_CHECK("k_node_mix<T>");
  if ((rc = launch_sgemm(true, Ut, bufA, bufB, M, (long long)B * Fin, M, M, st))) return rc;    // dx (vertex-major)
  return launch_from_node_major(bufB, dx, B, M, Fin, st);
}

}  // extern "C"
