// Training-step pieces around the conv stack ("next" rows of the scope table):
//   * sparse softmax cross-entropy, forward and backward in one pass (models_gcn.py:253-259),
//   * TensorFlow-1.x Adam over one flat parameter buffer, L2 term folded in (models_gcn.py:260-262, :294).
// They exist so that a whole training step is a handful of launches inside one CUDA graph.
#include <algorithm>

#include "common.cuh"

namespace gcnb {

// One warp per row: loss_row = logsumexp(z) - z[label]; dlogits = (softmax(z) - onehot) * scale.
// Row losses go to loss_rows[B]; a second tiny kernel averages them in a fixed order.
__global__ void k_softmax_xent(const float* __restrict__ logits, const long long* __restrict__ labels,
                               float* __restrict__ loss_rows, float* __restrict__ dlogits, int B, int C, float scale) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* z = logits + (long long)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(z[c] - mx);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
  const float lse = mx + logf(se);
  const int lab = (int)labels[row];
  if (dlogits != nullptr)
    for (int c = lane; c < C; c += 32) dlogits[(long long)row * C + c] = (expf(z[c] - lse) - (c == lab ? 1.f : 0.f)) * scale;
  if (lane == 0) loss_rows[row] = lse - z[lab];
}

__global__ void k_mean_rows(const float* __restrict__ v, float* __restrict__ out, int n) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += v[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0] / (float)n;
}

// state[0] = beta1^t, state[1] = beta2^t, state[2] = lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
__global__ void k_adam_tick(float* state, float lr, float b1, float b2) {
  const float p1 = state[0] * b1, p2 = state[1] * b2;
  state[0] = p1;
  state[1] = p2;
  state[2] = lr * sqrtf(1.f - p2) / (1.f - p1);
}

// g_total = g * gscale + reg * p (where decay[i] != 0);  m, v, p updated in place (TF "epsilon hat" form)
__global__ void k_adam_tf(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                          const uint8_t* __restrict__ decay, const float* __restrict__ state, long long n, float b1,
                          float b2, float eps, float reg, float gscale) {
  const float lr_t = state[2];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i] * gscale;
    if (decay != nullptr && decay[i]) gi = fmaf(reg, pi, gi);
    const float mi = fmaf(1.f - b1, gi - m[i], m[i]);
    const float vi = fmaf(1.f - b2, gi * gi - v[i], v[i]);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace gcnb

using namespace gcnb;

extern "C" {

int gcnb_softmax_xent_f32(const float* logits, const int64_t* labels, float* loss, float* dlogits, float* loss_rows,
                          int B, int C, gcnb_stream_t stream) {
  GCNB_REQUIRE(logits && labels && loss && loss_rows && B >= 1 && C >= 1, "gcnb_softmax_xent_f32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_softmax_xent<<<ceil_div(B * 32, 256), 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), loss_rows,
                                                       dlogits, B, C, 1.f / (float)B);
  GCNB_LAUNCH_CHECK("k_softmax_xent");
  k_mean_rows<<<1, 256, 0, st>>>(loss_rows, loss, B);
  GCNB_LAUNCH_CHECK("k_mean_rows");
  return GCNB_OK;
}

int gcnb_adam_tf_f32(float* p, const float* g, float* m, float* v, const uint8_t* decay, float* state, long long n,
                     float lr, float beta1, float beta2, float eps, float reg, float gscale, gcnb_stream_t stream) {
  GCNB_REQUIRE(p && g && m && v && state && n >= 1, "gcnb_adam_tf_f32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_adam_tick<<<1, 1, 0, st>>>(state, lr, beta1, beta2);
  GCNB_LAUNCH_CHECK("k_adam_tick");
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ceil_div_ll(n, 256), 148 * 8));
  k_adam_tf<<<grid, 256, 0, st>>>(p, g, m, v, decay, state, n, beta1, beta2, eps, reg, gscale);
  GCNB_LAUNCH_CHECK("k_adam_tf");
  return GCNB_OK;
}

}  // extern "C"
