// Data-parallel gradient averaging FUSED into the optimiser update: one cooperative kernel per step that
//   (A) publishes this rank's flat gradient in a peer-mapped staging buffer (double-buffered by step parity),
//   (B) exchanges one flag per peer over NVLink (release store into every peer's flag array, acquire spin on its own),
//   (C) reads every rank's staged gradient straight from peer memory (P2P loads through NVLink / NVSwitch), sums them in
//       RANK ORDER -- so every rank computes bit-identical averages and the replicas never drift -- and applies the
//       TF-1.x Adam step (models_gcn.py:294) in the same pass.
// It replaces `ncclAllReduce` + `k_adam_tf`: the all-reduce of this path is 632 KB once per ~0.3 ms step, i.e. pure
// latency (round 1 measured +30..50 us per step for the NCCL call at 2..8 GPUs), and a one-shot read of 7 x 632 KB per
// GPU is ~6 us of NVLink time.  The reference has no distributed code at all; the collective exists
// only because SURVEY.md 8(e) shards the batch.
//
// Memory the caller provides (host side: gcn_fmri_decoding_b200/train.py allocates it as torch symmetric memory and
// passes the peer pointers): per rank a staging area of 2*n floats followed by a flag array of 64 uint32, zeroed
// before the first step.  Flags carry the optimiser step number, which only grows: no reset, no second barrier --
// the parity slot a peer may still be reading is only rewritten two steps later, after that peer has signalled
// the step in between.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace gcnb {

constexpr int kMaxRanks = 16;

struct AdamArParams {
  float* p;
  const float* g;
  float *m, *v;
  const uint8_t* decay;
  const float* state;            // {b1^t, b2^t, lr_t, t}: t (already advanced for this step) keys parity and flags
  long long n;
  float b1, b2, eps, reg;
  float* stage[kMaxRanks];       // stage[q] = rank q's staging area (2*n floats), peer-mapped
  uint32_t* flags[kMaxRanks];    // flags[q] = rank q's flag array (one uint32 per source rank)
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) k_adam_tf_allreduce(const AdamArParams P) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t step = (uint32_t)P.state[3];
  const long long par = (long long)(step & 1u) * P.n;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  // (A) publish the local gradient
  float* mine = P.stage[P.rank] + par;
  for (long long i = tid; i < P.n; i += nth) mine[i] = P.g[i];
  __threadfence_system();
  grid.sync();
  // (B) one flag per peer, then wait for every peer's flag of this step
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < P.world) st_release_sys(P.flags[threadIdx.x] + P.rank, step);
    if ((int)threadIdx.x < P.world) {
      const uint32_t* f = P.flags[P.rank] + threadIdx.x;
      const long long t0 = clock64();
      while ((int32_t)(ld_acquire_sys(f) - step) < 0) {
        if (clock64() - t0 > 20000000000ll) __trap();  // a missing peer must not hang the GPU forever
      }
    }
  }
  grid.sync();
  // (C) sum the ranks in rank order straight from peer memory, then Adam
  const float lr_t = P.state[2];
  const float gscale = 1.f / (float)P.world;
  for (long long i = tid; i < P.n; i += nth) {
    float gs = 0.f;
    for (int q = 0; q < P.world; ++q) gs += __ldcv(P.stage[q] + par + i);  // volatile-class load: never a stale line
    const float pi = P.p[i];
    float gi = gs * gscale;
    if (P.decay != nullptr && P.decay[i]) gi = fmaf(P.reg, pi, gi);
    const float mi = fmaf(1.f - P.b1, gi - P.m[i], P.m[i]);
    const float vi = fmaf(1.f - P.b2, gi * gi - P.v[i], P.v[i]);
    P.m[i] = mi;
    P.v[i] = vi;
    P.p[i] = pi - lr_t * mi / (sqrtf(vi) + P.eps);
  }
}

}  // namespace gcnb

using namespace gcnb;

extern "C" {

size_t gcnb_adam_allreduce_stage_bytes(long long n) { return (size_t)(2 * n) * sizeof(float) + 64 * sizeof(uint32_t); }

int gcnb_adam_tf_allreduce_f32(float* p, const float* g, float* m, float* v, const uint8_t* decay, const float* state,
                               long long n, float beta1, float beta2, float eps, float reg, void* const* peer_stage,
                               int rank, int world, gcnb_stream_t stream) {
  GCNB_REQUIRE(p && g && m && v && state && n >= 1 && peer_stage, "gcnb_adam_tf_allreduce_f32: bad arguments");
  GCNB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world,
               "gcnb_adam_tf_allreduce_f32: rank %d / world %d out of range (at most %d ranks)", rank, world, kMaxRanks);
  AdamArParams P{};
  P.p = p; P.g = g; P.m = m; P.v = v; P.decay = decay; P.state = state; P.n = n;
  P.b1 = beta1; P.b2 = beta2; P.eps = eps; P.reg = reg; P.rank = rank; P.world = world;
  for (int q = 0; q < world; ++q) {
    GCNB_REQUIRE(peer_stage[q] != nullptr, "gcnb_adam_tf_allreduce_f32: peer pointer %d is NULL", q);
    P.stage[q] = static_cast<float*>(peer_stage[q]);
    P.flags[q] = reinterpret_cast<uint32_t*>(static_cast<float*>(peer_stage[q]) + 2 * n);
  }
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  const int grid = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(n, 256 * 4), di.sm_count));
  void* args[] = {(void*)&P};
  GCNB_CUDA(cudaLaunchCooperativeKernel((const void*)k_adam_tf_allreduce, dim3(grid), dim3(256), args, 0,
                                        static_cast<cudaStream_t>(stream)));
  GCNB_LAUNCH_CHECK("k_adam_tf_allreduce");
  return GCNB_OK;
}

}  // extern "C"
