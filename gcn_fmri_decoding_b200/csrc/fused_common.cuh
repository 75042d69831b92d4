// Building blocks of the fused (shared-memory resident) ChebyNet kernels.
//
// One CTA works on a tile of S windows.  Shared memory holds
//   ent[nnz]        (int2)  (byte offset of the neighbour's slab row = col*RS*4, value bits)
//   rowinfo[Mpad]   (int4)  rows sorted by decreasing length: (first entry, length, own row byte offset, row)
//   slab[Mpad][RS]  fp32 state X_k: row m = vertex, column s*FP + f (sample s of the tile, feature f
//                   padded to FP in {8,16,32}); RS = S*FP + 4, so RS/4 is odd and the 8 rows x 4 columns
//                   of an mma A-fragment load fall into 32 distinct banks.
// Warps form a grid [SG sample groups][RW]; a warp works on the CW = WS*FP columns of its sample group.
//
// Sparse recursion step (spmm_rows): each lane owns 4 consecutive columns (one LDS.128 per neighbour),
// so a row takes LPR = CW/4 lanes and a warp advances 32/LPR rows at once.  Rows are dealt to the
// warps in groups of equal length (sorted once per CTA), which keeps the lanes of a warp in lock step
// and the warps of a CTA balanced; ownership of rows in this phase is independent of the tensor-core
// phase, one __syncthreads() per Chebyshev order separates the two.
//
// Contraction (mma_tiles): warp (rw, sg) owns row tiles {rw, rw+RW, ...} of 16 vertices for its WS
// samples and multiplies them with the taps on the tensor cores (mma.sync m16n8k8 TF32, 3-pass
// error-compensated split = fp32-level accuracy), accumulating in registers across the K orders.
#pragma once
#include <algorithm>

#include "common.cuh"

namespace gcnb {

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// d += a(16x8, row) * b(8x8, col), TF32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// error-compensated product: a = ah + al, b = bh + bl (all TF32); drops only al*bl (~2^-22 relative)
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                           uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d, al[0], al[1], al[2], al[3], bh0, bh1);
  mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(v);
  lo = to_tf32(v - __uint_as_float(hi));
}

// Cheap split for the data operand (two instructions instead of two software-emulated cvt.rna):
// hi = v truncated to TF32 (exactly representable), lo = v - hi (exact in fp32; the tensor core reads its
// top 19 bits).  |v - hi - tf32(lo)| <= 2^-20 |v|, far below the 1e-4 parity bar.
__device__ __forceinline__ void split_trunc(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// Barrier over the warps of one sample group only (named barrier 1 + sg).  The slab columns of different sample
// groups are disjoint, so inside the recursion the groups need not wait for each other: they drift apart and the
// sparse phase of one overlaps the tensor-core phase of another.
__device__ __forceinline__ void group_barrier(int sg, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + sg), "r"(nthreads) : "memory");
}

// ---- mbarrier + TMA bulk copy (global -> shared), used to prefetch the next tile of windows ------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Same copy without the proxy fence, for full/empty mbarrier rings: the stage being overwritten was only READ through
// the generic proxy, and those reads are ordered before the copy by the consumers' arrive on the empty barrier and
// the producer's wait on it (no generic write to the destination to make visible to the async proxy).
__device__ __forceinline__ void bulk_g2s_ring(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Geometry shared by host (launch configuration) and device.
struct TileGeom {
  int M, Mpad, RT;  // vertices, padded to 16, row tiles of 16
  int FP, KS;       // padded feature width of the slab columns (8, 16 or 32), FP/8
  int WS, SG, S;    // samples per warp, sample groups, samples per CTA tile
  int RW, TPW;      // row warps, row tiles per warp (tensor-core phase)
  int RS;           // slab row stride in floats
  int LPR;          // lanes per row in the sparse phase = WS*FP/4 (8, 16 or 32)
  int nwarps;
};

static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

static inline size_t operator_smem_bytes(int Mpad, int nnz) {
  return align_up((size_t)std::max(nnz, 1) * 8, 16) + (size_t)Mpad * 16 + 2 * 256 * 4;
}

// Shared-memory image of the operator, built once per CTA from the CSR in global memory.
struct OperatorSmem {
  int2* ent;      // [nnz]
  int4* rowinfo;  // [Mpad]
  int* hist;      // [256] scratch
  int* start;     // [256] scratch
  __device__ void carve(unsigned char* p, int Mpad, int nnz) {
    ent = reinterpret_cast<int2*>(p);
    p += ((size_t)(nnz > 0 ? nnz : 1) * 8 + 15) / 16 * 16;
    rowinfo = reinterpret_cast<int4*>(p);
    p += (size_t)Mpad * 16;
    hist = reinterpret_cast<int*>(p);
    start = hist + 256;
  }
};

// Counting sort of the rows by decreasing length (lengths >= 255 share a bin).  All threads of the CTA.
__device__ __forceinline__ void build_operator(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                               const float* __restrict__ val, int M, int Mpad, int nnz, int RS,
                                               OperatorSmem& op) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < 256; i += nthr) op.hist[i] = 0;
  __syncthreads();
  for (int r = tid; r < Mpad; r += nthr) {
    const int len = r < M ? rowptr[r + 1] - rowptr[r] : 0;
    atomicAdd(&op.hist[min(len, 255)], 1);
  }
  for (int j = tid; j < nnz; j += nthr) op.ent[j] = make_int2(col[j] * RS * 4, __float_as_int(val[j]));
  __syncthreads();
  if (tid < 32) {  // start[bin] = number of rows in longer bins
    int carry = 0;
    for (int base = 224; base >= 0; base -= 32) {
      const int bin = base + 31 - tid;  // lane 0 takes the longest bin of the chunk
      const int h = op.hist[bin];
      int incl = h;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (tid >= d) incl += o;
      }
      op.start[bin] = carry + incl - h;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  for (int r = tid; r < Mpad; r += nthr) {
    const int beg = r < M ? rowptr[r] : 0;
    const int len = r < M ? rowptr[r + 1] - beg : 0;
    const int pos = atomicAdd(&op.start[min(len, 255)], 1);
    op.rowinfo[pos] = make_int4(beg, len, r * RS * 4, r);
  }
  __syncthreads();
}

// One recursion step for all rows dealt to this warp (columns: the warp's CW columns, 4 per lane):
//   dst[r] = alpha * (A src)[r] - (use_old ? dst[r] : 0)
// forward:  X_1 = L X_0 (alpha 1), X_k = 2 L X_{k-1} - X_{k-2};  adjoint (Clenshaw): b_k = 2 L^T b_{k+1} - b_{k+2}.
// `part`/`nparts`: this warp's share of the row groups (round robin over the length-sorted order).
template <int LPR>
__device__ __noinline__ void spmm_rows(const OperatorSmem& op, const unsigned char* __restrict__ src,
                                       unsigned char* __restrict__ dst, int Mpad, int col_byte, int part, int nparts,
                                       float alpha, bool use_old) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int q = lane / LPR, sub = lane % LPR;
  const unsigned char* s = src + col_byte + sub * 16;
  unsigned char* d = dst + col_byte + sub * 16;
  const int ngroups = Mpad / RPW;
  for (int gi = part; gi < ngroups; gi += nparts) {
    const int4 info = op.rowinfo[gi * RPW + q];
    const int len = info.y;
    const int nmin = RPW > 1 ? __reduce_min_sync(0xffffffffu, len) : len;
    const int nmax = RPW > 1 ? __reduce_max_sync(0xffffffffu, len) : len;
    const int2* e = op.ent + info.x;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    int it = 0;
    // four neighbours per trip: the four entry loads and the four data loads are independent (ILP); even entries
    // accumulate into a0, odd ones into a1 -- the same association as the two-at-a-time loops below
    for (; it + 4 <= nmin; it += 4) {
      const int2 e0 = e[it], e1 = e[it + 1], e2 = e[it + 2], e3 = e[it + 3];
      const float4 x0 = *reinterpret_cast<const float4*>(s + e0.x);
      const float4 x1 = *reinterpret_cast<const float4*>(s + e1.x);
      const float4 x2 = *reinterpret_cast<const float4*>(s + e2.x);
      const float4 x3 = *reinterpret_cast<const float4*>(s + e3.x);
      const float v0 = __int_as_float(e0.y), v1 = __int_as_float(e1.y), v2 = __int_as_float(e2.y), v3 = __int_as_float(e3.y);
      a0.x = fmaf(v0, x0.x, a0.x); a0.y = fmaf(v0, x0.y, a0.y); a0.z = fmaf(v0, x0.z, a0.z); a0.w = fmaf(v0, x0.w, a0.w);
      a1.x = fmaf(v1, x1.x, a1.x); a1.y = fmaf(v1, x1.y, a1.y); a1.z = fmaf(v1, x1.z, a1.z); a1.w = fmaf(v1, x1.w, a1.w);
      a0.x = fmaf(v2, x2.x, a0.x); a0.y = fmaf(v2, x2.y, a0.y); a0.z = fmaf(v2, x2.z, a0.z); a0.w = fmaf(v2, x2.w, a0.w);
      a1.x = fmaf(v3, x3.x, a1.x); a1.y = fmaf(v3, x3.y, a1.y); a1.z = fmaf(v3, x3.z, a1.z); a1.w = fmaf(v3, x3.w, a1.w);
    }
    for (; it + 2 <= nmin; it += 2) {
      const int2 e0 = e[it], e1 = e[it + 1];
      const float4 x0 = *reinterpret_cast<const float4*>(s + e0.x);
      const float4 x1 = *reinterpret_cast<const float4*>(s + e1.x);
      const float v0 = __int_as_float(e0.y), v1 = __int_as_float(e1.y);
      a0.x = fmaf(v0, x0.x, a0.x); a0.y = fmaf(v0, x0.y, a0.y); a0.z = fmaf(v0, x0.z, a0.z); a0.w = fmaf(v0, x0.w, a0.w);
      a1.x = fmaf(v1, x1.x, a1.x); a1.y = fmaf(v1, x1.y, a1.y); a1.z = fmaf(v1, x1.z, a1.z); a1.w = fmaf(v1, x1.w, a1.w);
    }
    // tail: same association as the main loop (even entries -> a0, odd -> a1), so a row's result does not
    // depend on which rows it was grouped with
    for (; it < nmax; it += 2) {
      if (it < len) {
        const int2 e0 = e[it];
        const float4 x0 = *reinterpret_cast<const float4*>(s + e0.x);
        const float v0 = __int_as_float(e0.y);
        a0.x = fmaf(v0, x0.x, a0.x); a0.y = fmaf(v0, x0.y, a0.y); a0.z = fmaf(v0, x0.z, a0.z); a0.w = fmaf(v0, x0.w, a0.w);
      }
      if (it + 1 < len) {
        const int2 e1 = e[it + 1];
        const float4 x1 = *reinterpret_cast<const float4*>(s + e1.x);
        const float v1 = __int_as_float(e1.y);
        a1.x = fmaf(v1, x1.x, a1.x); a1.y = fmaf(v1, x1.y, a1.y); a1.z = fmaf(v1, x1.z, a1.z); a1.w = fmaf(v1, x1.w, a1.w);
      }
    }
    a0.x += a1.x; a0.y += a1.y; a0.z += a1.z; a0.w += a1.w;
    float4* out = reinterpret_cast<float4*>(d + info.z);
    if (use_old) {
      const float4 o = *out;
      a0.x = fmaf(alpha, a0.x, -o.x); a0.y = fmaf(alpha, a0.y, -o.y); a0.z = fmaf(alpha, a0.z, -o.z); a0.w = fmaf(alpha, a0.w, -o.w);
    } else {
      a0.x *= alpha; a0.y *= alpha; a0.z *= alpha; a0.w *= alpha;
    }
    *out = a0;
  }
}

// LPR = CW/4 lanes per row with CW the number of slab columns a warp works on (8 .. 128)
__device__ __forceinline__ void spmm_dispatch(int LPR, const OperatorSmem& op, const unsigned char* src,
                                              unsigned char* dst, int Mpad, int col_byte, int part, int nparts,
                                              float alpha, bool use_old) {
  switch (LPR) {
    case 2: spmm_rows<2>(op, src, dst, Mpad, col_byte, part, nparts, alpha, use_old); break;
    case 4: spmm_rows<4>(op, src, dst, Mpad, col_byte, part, nparts, alpha, use_old); break;
    case 8: spmm_rows<8>(op, src, dst, Mpad, col_byte, part, nparts, alpha, use_old); break;
    case 16: spmm_rows<16>(op, src, dst, Mpad, col_byte, part, nparts, alpha, use_old); break;
    default: spmm_rows<32>(op, src, dst, Mpad, col_byte, part, nparts, alpha, use_old); break;
  }
}

}  // namespace gcnb
