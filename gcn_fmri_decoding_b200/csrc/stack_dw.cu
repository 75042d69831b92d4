// Weight / bias gradient of the ChebyNet layer from a saved Chebyshev basis (training, first layer).
//
// When the forward pass kept X_k for all k (`xstack`, [K][B*M][FP]), the weight gradient
//     dW[f*K+k][o] = sum_{rows} X_k[row][f] * dZ[row][o],        rows = (sample, vertex)
// needs no recursion at all: it is one tall-skinny GEMM with the rows as the reduction dimension, streamed once from
// HBM.  That trades K*|x| bytes of extra traffic (cheap at 6+ TB/s) for the shared-memory-bound recomputation of the
// basis (measured 3x slower for the 400-vertex layer).  Each CTA walks chunks of 128 rows: the X tiles arrive by
// cp.async into a double buffer, dZ (MaxPoolGrad o ReluGrad of dy) is rebuilt per chunk from the pooled tensors, the
// contraction runs on the tensor cores (3xTF32 mma.sync, accumulators live in registers for the whole kernel), and
// the result is reduced warps -> CTA -> k_dw_from_partials in a fixed order (no float atomics).
#include <algorithm>

#include "fused_common.cuh"

namespace gcnb {


struct StackDwParams {
  const float* xstack;
  const float* y;
  const uint8_t* argmax;
  const float* dy;
  float* dw_part;
  float* db_part;  // nullable
  long long R;     // B*M rows
  int B, M, Mo, Fout, FoP, FP, p, log2p, relu, dy_is_mean;
  int CR, SX, SZ, nchunks;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int K, int MT, int NT>
// light instances (<= 20 accumulator fragments) run two CTAs per SM: one CTA's loads hide behind the other's MMAs
__global__ void __launch_bounds__(256, (K * MT * NT > 20) ? 1 : 2) k_dw_from_stack(const StackDwParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int CR = P.CR, SX = P.SX, SZ = P.SZ, FP = P.FP;
  const size_t xbytes = (size_t)K * CR * SX * 4, zbytes = (size_t)CR * SZ * 4;
  float* Xs[2] = {reinterpret_cast<float*>(smem), reinterpret_cast<float*>(smem + xbytes + zbytes)};
  float* Zs[2] = {reinterpret_cast<float*>(smem + xbytes), reinterpret_cast<float*>(smem + 2 * xbytes + zbytes)};
  constexpr int NCH = MT * NT / 2;

  float acc[K][MT][NT][4];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[k][m][n][c] = 0.f;
  float dbacc = 0.f;  // column o = tid % FoP of dZ (FoP divides 256)

  const int q4 = FP >> 2;
  auto issue_x = [&](int chunk, int buf) {  // K tiles of CR rows x FP floats, dense in HBM, padded rows in smem
    const long long r0 = (long long)chunk * CR;
    const int q_shift = FP == 8 ? 1 : (FP == 16 ? 2 : 3), r_shift = CR == 128 ? 7 : (CR == 64 ? 6 : 5);
    for (int idx = tid; idx < K * CR * q4; idx += 256) {
      const int q = idx & (q4 - 1), r = (idx >> q_shift) & (CR - 1), k = idx >> (q_shift + r_shift);
      float* dst = Xs[buf] + ((size_t)k * CR + r) * SX + 4 * q;
      if (r0 + r < P.R) cp_async16(dst, P.xstack + ((size_t)k * P.R + r0 + r) * FP + 4 * q);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  // dZ rows of a chunk from the pooled tensors, split in two halves so that the global loads of the NEXT chunk are
  // in flight while the tensor cores work on the current one.  M % p == 0 and CR % p == 0, so a pooling window never
  // straddles a chunk and the flat pooled row is simply row >> log2p: no divisions, every pooled element is loaded
  // once, and the (up to) 4 x 3 loads of a thread are independent.
  const int fo_shift = P.FoP == 16 ? 4 : 5;
  const int zo = tid & (P.FoP - 1);
  const int rows_per_pass = 256 >> fo_shift, cpr = CR >> P.log2p;
  const long long npr = P.R >> P.log2p;
  float gv[4];
  int am[4];
  auto load_z = [&](int chunk) {  // requires cpr <= 4 * rows_per_pass (checked on the host)
    const long long pr0 = ((long long)chunk * CR) >> P.log2p;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = (tid >> fo_shift) + u * rows_per_pass;
      const long long pr = pr0 + j;
      gv[u] = 0.f;
      am[u] = 0;
      if (j < cpr && pr < npr && zo < P.Fout) {
        const long long gi = pr * P.Fout + zo;
        if (P.argmax != nullptr && P.p > 1) am[u] = __ldg(P.argmax + gi);
        gv[u] = P.dy_is_mean ? __ldg(P.dy + pr) / (float)P.Fout : __ldg(P.dy + gi);
        if (P.relu && !(__ldg(P.y + gi) > 0.f)) gv[u] = 0.f;
      }
    }
  };
  auto store_z = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = (tid >> fo_shift) + u * rows_per_pass;
      if (j >= cpr) continue;
      dbacc += gv[u];
      float* dst = Zs[buf] + (size_t)(j << P.log2p) * SZ + zo;
      for (int i = 0; i < P.p; ++i) dst[(size_t)i * SZ] = (i == am[u]) ? gv[u] : 0.f;
    }
  };

  int chunk = blockIdx.x, buf = 0;
  if (chunk < P.nchunks) {
    issue_x(chunk, 0);
    load_z(chunk);
    store_z(0);
  }
  for (; chunk < P.nchunks; chunk += gridDim.x, buf ^= 1) {
    const int next = chunk + gridDim.x;
    if (next < P.nchunks) {
      issue_x(next, buf ^ 1);
      load_z(next);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* xs = Xs[buf];
    const float* zs = Zs[buf];
    for (int step = warp; step < CR / 8; step += 8) {
      const int r8 = step * 8;
      uint32_t bh[NT][2], bl[NT][2];
      const float* zr = zs + (size_t)(r8 + t) * SZ + g;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        split_trunc(zr[n * 8], bh[n][0], bl[n][0]);
        split_trunc(zr[n * 8 + 4 * SZ], bh[n][1], bl[n][1]);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float* xr = xs + ((size_t)k * CR + r8 + t) * SX + g;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          uint32_t ah[4], al[4];
          const bool hi_ok = (m * 16 + g + 8) < FP;
          split_trunc(xr[m * 16], ah[0], al[0]);
          split_trunc(hi_ok ? xr[m * 16 + 8] : 0.f, ah[1], al[1]);
          split_trunc(xr[m * 16 + 4 * SX], ah[2], al[2]);
          split_trunc(hi_ok ? xr[m * 16 + 8 + 4 * SX] : 0.f, ah[3], al[3]);
#pragma unroll
          for (int n = 0; n < NT; ++n) mma_3xtf32(acc[k][m][n], ah, al, bh[n][0], bh[n][1], bl[n][0], bl[n][1]);
        }
      }
    }
    if (next < P.nchunks) store_z(buf ^ 1);  // Zs[buf^1] was last read one iteration ago (barrier below)
    __syncthreads();  // everyone done with `buf` before it is refilled two iterations later
  }

  // ---- warps -> CTA (fixed order), CTA partial in the chunk layout k_dw_from_partials expects -----------------
  float* red = reinterpret_cast<float*>(smem);  // 8 warps x K*NCH*256 floats (fits: see host-side check)
  const int per_warp = K * NCH * 256;
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          red[(size_t)warp * per_warp + (k * NCH + m * (NT / 2) + n / 2) * 256 + ((n & 1) * 4 + c) * 32 + lane] =
              acc[k][m][n][c];
  __syncthreads();
  float* out = P.dw_part + (size_t)blockIdx.x * per_warp;
  for (int e = tid; e < per_warp; e += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[(size_t)w * per_warp + e];
    out[e] = s;
  }
  if (P.db_part != nullptr) {
    __syncthreads();
    red[tid] = dbacc;
    __syncthreads();
    if (tid < P.FoP) {
      float s = 0.f;
      for (int i = tid; i < 256; i += P.FoP) s += red[i];
      P.db_part[(size_t)blockIdx.x * P.FoP + tid] = s;
    }
  }
}

struct StackPlan {
  bool ok;
  int FP, FoP, MT, NT, CR, SX, SZ, ctas_per_sm;
  size_t smem;
};

int fused_feature_pad(int Fin) { return Fin < 1 || Fin > 32 ? 0 : (Fin <= 8 ? 8 : (Fin <= 16 ? 16 : 32)); }

static StackPlan plan_stack(const LayerShape& s) {
  StackPlan pl;
  pl.ok = false;
  pl.FP = fused_feature_pad(s.Fin);
  if (pl.FP == 0 || s.Fout < 1 || s.Fout > 32 || s.K < 1 || s.K > 5) return pl;
  if (s.p > 16 || s.M % s.p != 0) return pl;
  pl.FoP = s.Fout <= 16 ? 16 : 32;
  pl.MT = pl.FP <= 16 ? 1 : 2;
  pl.NT = pl.FoP / 8;
  pl.SX = pl.FP == 8 ? 24 : pl.FP + 8;
  pl.SZ = pl.FoP + 8;
  const size_t red = (size_t)8 * s.K * (pl.MT * pl.NT / 2) * 256 * 4;
  const bool light = s.K * pl.MT * pl.NT <= 20;
  for (int pass = light ? 0 : 1; pass < 2; ++pass) {  // pass 0: two CTAs per SM (<= 110 KB each); pass 1: one
    const size_t budget = pass == 0 ? 110 * 1024 : 220 * 1024;
    for (int cr : {128, 64, 32}) {
      const size_t need = 2 * ((size_t)s.K * cr * pl.SX * 4 + (size_t)cr * pl.SZ * 4);
      if (cr / s.p > 4 * 256 / pl.FoP) continue;  // pooled rows of a chunk must fit the 4 register slots per thread
      if (std::max(need, red) <= budget) {
        pl.CR = cr;
        pl.smem = std::max(need, red);
        pl.ctas_per_sm = pass == 0 ? 2 : 1;
        pl.ok = true;
        return pl;
      }
    }
  }
  return pl;
}

bool stack_dw_supported(const LayerShape& s) { return plan_stack(s).ok; }

size_t stack_dw_workspace(const LayerShape& s) {
  const StackPlan pl = plan_stack(s);
  if (!pl.ok) return 0;
  return align_up((size_t)296 * (s.K * (pl.MT * pl.NT / 2) * 256 + 32) * 4, 256) + db_vertex_workspace(s) + 512;
}

template <int K, int MT, int NT>
static int launch_stack(const StackDwParams& P, const StackPlan& pl, int grid, cudaStream_t st) {
  auto kern = k_dw_from_stack<K, MT, NT>;
  GCNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  kern<<<grid, 256, pl.smem, st>>>(P);
  GCNB_LAUNCH_CHECK("k_dw_from_stack");
  return GCNB_OK;
}

int stack_dw(const float* xstack, const float* y, const uint8_t* argmax, const float* dy, int dy_is_mean, float* dW,
             float* db, const LayerShape& s, int bias_mode, int relu, Workspace& ws, cudaStream_t st) {
  const StackPlan pl = plan_stack(s);
  if (!pl.ok) {
    set_error("stack_dw: unsupported shape");
    return GCNB_ERR_INVALID;
  }
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  StackDwParams P;
  P.xstack = xstack; P.y = y; P.argmax = argmax; P.dy = dy;
  P.R = (long long)s.B * s.M; P.B = s.B; P.M = s.M; P.Mo = s.M / s.p; P.Fout = s.Fout; P.FoP = pl.FoP; P.FP = pl.FP;
  P.p = s.p; P.log2p = 0;
  while ((1 << P.log2p) < s.p) ++P.log2p;
  P.relu = relu; P.dy_is_mean = dy_is_mean;
  P.CR = pl.CR; P.SX = pl.SX; P.SZ = pl.SZ;
  P.nchunks = (int)ceil_div_ll(P.R, pl.CR);
  const int grid = std::min(P.nchunks, std::min(di.sm_count * pl.ctas_per_sm, 296));
  const int nch = pl.MT * pl.NT / 2;
  float* part = ws.take<float>((size_t)grid * s.K * nch * 256);
  float* dbf = ws.take<float>((size_t)grid * 32);
  if (!part || !dbf) {
    set_error("stack_dw: workspace too small");
    return GCNB_ERR_WORKSPACE;
  }
  const bool db_fused = bias_mode == GCNB_BIAS_PER_FILTER && db != nullptr;
  P.dw_part = part;
  P.db_part = db_fused ? dbf : nullptr;
  rc = GCNB_ERR_INVALID;
#define GCNB_STACK_CASE(k, mt, nt) \
  if (s.K == k && pl.MT == mt && pl.NT == nt) rc = launch_stack<k, mt, nt>(P, pl, grid, st);
#define GCNB_STACK_K(k) GCNB_STACK_CASE(k, 1, 2) GCNB_STACK_CASE(k, 1, 4) GCNB_STACK_CASE(k, 2, 2) GCNB_STACK_CASE(k, 2, 4)
  GCNB_STACK_K(1) GCNB_STACK_K(2) GCNB_STACK_K(3) GCNB_STACK_K(4) GCNB_STACK_K(5)
#undef GCNB_STACK_K
#undef GCNB_STACK_CASE
  if (rc) return rc;
  rc = launch_dw_from_partials(part, dW, grid, s.K, pl.MT, pl.NT, s.Fin, s.Fout, P.db_part, db, pl.FoP, st);
  if (rc) return rc;
  if (bias_mode == GCNB_BIAS_PER_VERTEX && db != nullptr)
    return launch_db_vertex(dy, y, argmax, db, s, relu, dy_is_mean, ws, st);
  return GCNB_OK;
}

}  // namespace gcnb
