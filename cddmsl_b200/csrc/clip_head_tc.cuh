// Large-vocabulary CLIP head on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// For LVIS-scale vocabularies (K = 1203 concepts, fast_rcnn.py:549-555 with `test_cls_score` /
// configs/LVISv1-InstanceSegmentation/CLIP_fast_rcnn_R_50_C4.yaml) the region-by-text product is a real dense
// contraction: [R x 1024] . [1024 x 1204] forward and [R x 1204] . [1204 x 1024] for dx — 3.2e11 flop at R = 65536,
// arithmetic intensity ~400 flop/B, i.e. tensor-core bound.  The parity bar (1e-5 relative on logits
// that are cosines / 0.01) rules out plain TF32 (10-bit mantissa -> ~1e-3), so every operand is split
//     x = hi + lo,  hi = x with the low 13 mantissa bits cleared (exact in TF32),  lo = x - hi
// and each K-slice issues three tcgen05.mma: hi.hi + hi.lo + lo.hi, accumulated in fp32 in TMEM (error ~2^-22).
//
// One kernel, C[M x N] = A[M x K] . B[N x K]^T, both operands K-major fp32:
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of a [128 x 32] A tile and a [128 x 32] B tile per stage,
//               128-byte swizzle, completion on an mbarrier
//   warps 2-9   split: every thread reads its row of the freshly landed tiles (= the hi operands as the tensor core
//               sees them) and writes lo = x - trunc(x) to a second tile with the same swizzled addresses; in GEMM 1
//               they also accumulate |x|^2 per row
//   warp 1      allocates TMEM, issues 3 x 4 tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8) per stage from
//               shared-memory descriptors, tcgen05.commit releases the stage / publishes the accumulator
//   warps 10-13 epilogue: tcgen05.ld 32x32b, row scaling, global stores
// 3 stages x 64 KB of shared memory, one persistent CTA per SM, two TMEM accumulators (epilogue overlaps the next tile).
//   GEMM 1 epilogue: logits = acc * (1 / max(|x|, eps)) / T  -> scores [R, K+1], norms
//   (CUDA cores)   : row softmax / CE / focal factor / statistics, G = dL/dlogits, per-row scalars
//   GEMM 2 epilogue: dx = acc * su[r] - x * sx[r]
// (included at the end of clip_head.cu: same translation unit as the prep / finish kernels it launches)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma_host.cuh"

namespace cddmsl {

// 128 x 256 output tiles: with fp32 operands the 128 x 128 version moved 5.4 GB from L2 per GEMM (5.2 TB/s, the L2
// limit; ncu round 2: tensor pipe 50 %); a 256-wide B tile cuts that to 4.0 GB and the shared-memory reads per flop
// of the three MMAs by a quarter.
constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 32;  // BK * 4 B = 128 B = one swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;        // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;        // 32 KB
constexpr int TC_OFF_ALO = TC_A_BYTES, TC_OFF_BHI = 2 * TC_A_BYTES, TC_OFF_BLO = 2 * TC_A_BYTES + TC_B_BYTES;
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // A hi, A lo, B hi, B lo = 96 KB
constexpr int TC_SPLIT_WARPS = (TC_BM + TC_BN) / 32;              // one thread per operand row
constexpr int TC_THREADS = (2 + TC_SPLIT_WARPS + 4) * 32;  // warp 0 TMA, warp 1 MMA, 12 split warps, 4 epilogue warps
constexpr uint32_t TC_TMEM_COLS = 2 * TC_BN;  // two fp32 accumulators of TC_BN columns (all 512 columns of the SM)

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_c),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// TF32 "hi" part: clear the low 13 mantissa bits (one LOP3).  cvt.rna.tf32.f32 would round to nearest, but it is
// emulated on sm_100a (~7 instructions) and made the split warps the bottleneck (ncu: 43 % issue-active, tensor pipe
// 26 %); the accumulated error is dominated by the tensor-core accumulation either way (measured 1.6e-4 vs 1.8e-4
// max abs on logits of scale 17.7).
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// K-major, 128-byte swizzle: rows of 128 B, 8-row atoms of 1024 B (SBO), LBO unused (1), descriptor version 1.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, N = 128, M = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                ((uint32_t)(TC_BM >> 4) << 24);

struct TcEpilogue {
  int mode;            // 0: logits (GEMM 1), 1: dx (GEMM 2)
  float* out;          // [M, ldo]
  int ldo;             // leading dimension of out (floats)
  int n_valid;         // columns of out that exist
  float inv_T;
  float* norms;        // mode 0: |x| per row (written by the n-tile 0 CTAs)
  const float* x;      // mode 1: [M, ldo] the un-normalised embeddings
  const float* su;     // mode 1: per-row scale of the accumulator
  const float* sx;     // mode 1: per-row scale of x
};

// Persistent: each CTA walks output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, so CTAs that run
// at the same time share their A rows in L2).  Three pipelines: smem stages (TMA -> split -> MMA -> TMA), two TMEM
// accumulators (MMA <-> epilogue) so the epilogue of tile i overlaps the main loop of tile i+1, and the tile walk.
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N,
                   int K, TcEpilogue ep) {
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[TC_STAGES], conv_bar[TC_STAGES], empty_bar[TC_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float ss_s[2][TC_BM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  const int tiles_n = (N + TC_BN - 1) / TC_BN;
  const int ntiles = tiles_n * ((M + TC_BM - 1) / TC_BM);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], TC_SPLIT_WARPS);   // one arrival per split warp
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int g = 0;  // k-blocks issued so far (ring position)
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * TC_BM, n0 = (tile % tiles_n) * TC_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % TC_STAGES;
          const uint32_t ph = (g / TC_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);  // first pass over the ring falls through (barrier parity starts at 0)
          uint8_t* st = smem + (size_t)s * TC_STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], TC_A_BYTES + TC_B_BYTES);
          tma_load_2d(st, &map_a, kb * TC_BK, m0, &full_bar[s]);               // A hi slot
          tma_load_2d(st + TC_OFF_BHI, &map_b, kb * TC_BK, n0, &full_bar[s]);  // B hi slot
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int g = 0, it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + (uint32_t)(buf * TC_BN);
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        const int s = g % TC_STAGES;
        const uint32_t ph = (g / TC_STAGES) & 1;
        // The tensor core reads fp32 words as TF32 by ignoring the low 13 mantissa bits, i.e. the freshly landed
        // tiles ARE the hi operands: hi.hi can be issued as soon as the TMA data is there, while the split warps
        // are still producing the lo tiles for the two cross terms.
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = smem_u32(smem + (size_t)s * TC_STAGE_BYTES);
        const uint64_t a_hi = make_desc_sw128(st), a_lo = make_desc_sw128(st + TC_OFF_ALO);
        const uint64_t b_hi = make_desc_sw128(st + TC_OFF_BHI), b_lo = make_desc_sw128(st + TC_OFF_BLO);
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 B per K=8 step inside the swizzle row
            umma_tf32(acc, a_hi + adv, b_hi + adv, kIdescTf32, (kb | k) != 0);
          }
        }
        __syncwarp();
        mbar_wait(&conv_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
            umma_tf32(acc, a_hi + adv, b_lo + adv, kIdescTf32, 1u);
            umma_tf32(acc, a_lo + adv, b_hi + adv, kIdescTf32, 1u);
          }
          umma_commit(&empty_bar[s]);                       // stage reusable once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(&acc_full[buf]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + TC_SPLIT_WARPS) {
    // ===================== split (hi / lo) warps =====================
    // 384 threads: thread u < 128 rewrites row u of the A tile, thread u >= 128 row u-128 of the B tile
    const int u = threadIdx.x - 64;
    const int op = u >= TC_BM ? 1 : 0;
    const int t = op ? u - TC_BM : u;
    int g = 0, it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      float ss = 0.f;
      for (int kb = 0; kb < num_kb; ++kb, ++g) {
        const int s = g % TC_STAGES;
        const uint32_t ph = (g / TC_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        uint8_t* st = smem + (size_t)s * TC_STAGE_BYTES;
        float4* hi = reinterpret_cast<float4*>(st + (op ? TC_OFF_BHI : 0)) + t * 8;
        float4* lo = reinterpret_cast<float4*>(st + (op ? TC_OFF_BLO : TC_OFF_ALO)) + t * 8;
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = hi[c ^ (t & 7)];  // swizzled visiting order: conflict-free per quarter warp
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cc = c ^ (t & 7);
          if (op == 0)
            ss = fmaf(v[c].x, v[c].x, fmaf(v[c].y, v[c].y, fmaf(v[c].z, v[c].z, fmaf(v[c].w, v[c].w, ss))));
          float4 l;  // lo = v - trunc_tf32(v): exact in fp32; the tensor core drops its low bits (2^-21 |v|)
          l.x = v[c].x - to_tf32(v[c].x);
          l.y = v[c].y - to_tf32(v[c].y);
          l.z = v[c].z - to_tf32(v[c].z);
          l.w = v[c].w - to_tf32(v[c].w);
          lo[cc] = l;
        }
        // |x|^2 of the tile's rows travels with the last stage: written before the arrive the MMA (and through its
        // commit the epilogue) synchronises on.  The slot is free again long before tile it+2 gets here: its MMAs
        // cannot start until the epilogue of tile `it` has released the accumulator.
        if (op == 0 && kb == num_kb - 1) ss_s[it & 1][t] = ss;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[s]);
      }
    }
  } else {
    // ===================== epilogue warps (the last four): a warp may touch TMEM lanes [32 * (warp % 4), +32)
    const int lg = warp & 3;
    const int row_in_tile = lg * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int m0 = (tile / tiles_n) * TC_BM, n0 = (tile % tiles_n) * TC_BN;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + row_in_tile;
      float rs = 0.f, rx = 0.f;
      if (ep.mode == 0) {
        const float nrm = sqrtf(ss_s[buf][row_in_tile]);
        rs = ep.inv_T / fmaxf(nrm, 1e-12f);
        if (row < M && n0 == 0) ep.norms[row] = nrm;
      } else if (row < M) {
        rs = ep.su[row];
        rx = ep.sx[row];
      }
#pragma unroll 1
      for (int c0 = 0; c0 < TC_BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * TC_BN + c0), r);
        if (row < M) {
          const int col0 = n0 + c0;
          float* o = ep.out + (size_t)row * ep.ldo + col0;
          const float* xr = ep.mode == 1 ? ep.x + (size_t)row * ep.ldo + col0 : nullptr;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            if (col0 + q + 3 < ep.n_valid) {
              float4 v = make_float4(__uint_as_float(r[q]) * rs, __uint_as_float(r[q + 1]) * rs,
                                     __uint_as_float(r[q + 2]) * rs, __uint_as_float(r[q + 3]) * rs);
              if (ep.mode == 1) {
                const float4 xv = *reinterpret_cast<const float4*>(xr + q);
                v.x = fmaf(-xv.x, rx, v.x);
                v.y = fmaf(-xv.y, rx, v.y);
                v.z = fmaf(-xv.z, rx, v.z);
                v.w = fmaf(-xv.w, rx, v.w);
              }
              *reinterpret_cast<float4*>(o + q) = v;
            } else {
              for (int e = 0; e < 4; ++e)
                if (col0 + q + e < ep.n_valid) {
                  float v = __uint_as_float(r[q + e]) * rs;
                  if (ep.mode == 1) v = fmaf(-xr[q + e], rx, v);
                  o[q + e] = v;
                }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);  // this warp's quarter of the accumulator is drained
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// row pass between the two GEMMs (CUDA cores, memory bound): softmax / loss / statistics / dL/dlogits
// ------------------------------------------------------------------------------------------------
struct TcRowArgs {
  const float* logits;  // [R, K1]
  const int64_t* gt;
  const float* norms;   // |x|
  const float* grad_scale;
  const float* norm_dev;
  float norm_host;
  int R, K1, K1p;
  float inv_T;
  int loss_mode;
  float gamma, bg_weight;
  int strict_nan;
  float* G;        // [R, K1p] (nullable): dL/dlogits, zero padded
  float* su;       // [R]
  float* sx;       // [R]
  float* partial;  // [grid]
  int32_t* stats;
};

__global__ void __launch_bounds__(256) clip_head_tc_rows_kernel(TcRowArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K1 - 1;
  float loss_acc = 0.f;
  int st_acc = 0, st_fg = 0, st_fgacc = 0, st_fn = 0;
  const float gscale = a.grad_scale ? *a.grad_scale : 1.f;
  const float inv_den = 1.f / (a.norm_dev ? *a.norm_dev : a.norm_host);
  for (int r = blockIdx.x * 8 + warp; r < a.R; r += gridDim.x * 8) {
    const float* lg = a.logits + (size_t)r * a.K1;
    float m = -INFINITY;
    int am = 0x7fffffff;
    for (int k = lane; k < a.K1; k += 32) {
      const float l = lg[k];
      if (l > m) {
        m = l;
        am = k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) {
        m = om;
        am = oa;
      }
    }
    float se = 0.f;
    for (int k = lane; k < a.K1; k += 32) se += expf(lg[k] - m);
    se = warp_sum(se);
    const int t = min(max((int)a.gt[r], 0), K);
    const float lt = lg[t];
    const float ce = (logf(se) + m) - lt;
    const float pt = expf(lt - m) / se;
    const bool is_bg = (t == K);
    float wi = 1.f, li, coef;
    if (a.loss_mode == CDDMSL_LOSS_FOCAL) {
      if (a.bg_weight >= 0.f && is_bg) wi = a.bg_weight;
      const float om = 1.f - pt;
      const float f = powf(om, a.gamma);
      li = ce * f;
      float t2;
      if (om > 0.f) t2 = a.gamma * ce * pt * (f / om);
      else t2 = a.strict_nan ? __int_as_float(0x7fc00000) : 0.f;
      coef = f + t2;
    } else {
      if (a.loss_mode == CDDMSL_LOSS_WEIGHTED_CE && is_bg) wi = a.bg_weight;
      li = ce;
      coef = 1.f;
    }
    if (lane == 0) {
      loss_acc += wi * li;
      st_acc += (am == t);
      if (t < K) {
        st_fg += 1;
        st_fgacc += (am == t);
        st_fn += (am == K);
      }
    }
    if (a.G) {
      const float cg = coef * wi * inv_den * gscale;
      const float inv_se = 1.f / se;
      float* g = a.G + (size_t)r * a.K1p;
      float s = 0.f;
      for (int k = lane; k < a.K1p; k += 32) {
        float gk = 0.f;
        if (k < a.K1) {
          const float l = lg[k];
          gk = cg * (expf(l - m) * inv_se - (k == t ? 1.f : 0.f));
          s = fmaf(gk, l, s);  // sum_k g_k l_k = (sum_k g_k c_k) / T
        }
        g[k] = gk;
      }
      s = warp_sum(s);
      if (lane == 0) {
        const float nrm = a.norms[r];
        const float inv_n = 1.f / fmaxf(nrm, 1e-12f);
        a.su[r] = inv_n * a.inv_T;
        a.sx[r] = nrm < 1e-12f ? 0.f : s * inv_n * inv_n;
      }
    }
  }
  __shared__ float red[8];
  if (lane == 0) red[warp] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    a.partial[blockIdx.x] = s;
  }
  if (a.stats && lane == 0) {
    if (st_acc) atomicAdd(a.stats + 0, st_acc);
    if (st_fg) atomicAdd(a.stats + 1, st_fg);
    if (st_fgacc) atomicAdd(a.stats + 2, st_fgacc);
    if (st_fn) atomicAdd(a.stats + 3, st_fn);
  }
}

// wallT[d][k] = wall[k][d], zero padded to K1p columns
__global__ void clip_head_tc_transpose_w_kernel(const float* __restrict__ wall, float* __restrict__ wallT, int K1,
                                                int K1p, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * K1p) return;
  const int d = i / K1p, k = i - d * K1p;
  wallT[i] = k < K1 ? wall[(size_t)k * D + d] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// row-major fp32 [rows, cols] (ld floats between rows), box [TC_BK cols x box_rows rows], 128-byte swizzle
static int make_map(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_rows) {
  return tma_encode_2d_f32(m, base, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld * 4,
                           TC_BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)
             ? CDDMSL_EINVAL
             : 0;
}

static int launch_gemm(const float* A, int M, int lda, const float* B, int N, int ldb, int K, TcEpilogue ep,
                       cudaStream_t stream) {
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M, K, lda, TC_BM);
  if (rc) return rc;
  rc = make_map(&mb, B, N, K, ldb, TC_BN);
  if (rc) return rc;
  const int smem = TC_STAGES * TC_STAGE_BYTES + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  const int ntiles = ceil_div(N, TC_BN) * ceil_div(M, TC_BM);
  gemm_tf32x3_kernel<<<min(ntiles, sm_count()), TC_THREADS, smem, stream>>>(ma, mb, M, N, K, ep);
  count_launch();
  return (int)cudaGetLastError();
}

int g_head_tc = 2;  // tuning knob "head_tc": 0 never, 1 whenever the shape allows, 2 automatic

bool clip_head_tc_eligible(int R, int D, int K) {
  const int forced = g_head_tc;
  if (forced == 0) return false;
  const bool shape_ok = (D % 4 == 0) && D >= TC_BK && ((K + 1) % 4 == 0) && R >= 1;
  if (forced == 1) return shape_ok;
  return shape_ok && K >= 255 && R >= 512;  // below that the region-by-text product is not a dense contraction
}

size_t clip_head_tc_workspace_bytes(int R, int D, int K) {
  const size_t K1 = K + 1, K1p = (K1 + 31) / 32 * 32;
  size_t b = 0;
  b += align_up(K1 * D * 4, 256);          // wall
  b += align_up((size_t)D * K1p * 4, 256); // wallT
  b += align_up((size_t)R * K1 * 4, 256);  // logits (when the caller does not want scores)
  b += align_up((size_t)R * K1p * 4, 256); // G
  b += 3 * align_up((size_t)R * 4, 256);   // norms, su, sx
  b += align_up(1184 * 4, 256) + 256;      // partials, norm
  return b;
}

// op: 0 scores only, 1 loss (+dx), 2 dx from given dscores
int clip_head_tc_run(int op, const float* x, const float* w, const float* w_bg, const int64_t* gt,
                     const float* dscores, int R, int D, int K, float temperature, int loss_mode, float gamma,
                     float bg_weight, const float* grad_scale, int strict_nan, float* scores, float* loss, float* dx,
                     int32_t* stats, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (clip_head_tc_workspace_bytes(R, D, K) > workspace_bytes) return CDDMSL_EWORKSPACE;
  const int K1 = K + 1, K1p = (K1 + 31) / 32 * 32;
  char* p = (char*)workspace;
  auto take = [&](size_t bytes) {
    void* q = p;
    p += align_up(bytes, 256);
    return q;
  };
  float* wall = (float*)take((size_t)K1 * D * 4);
  float* wallT = (float*)take((size_t)D * K1p * 4);
  float* lg_ws = (float*)take((size_t)R * K1 * 4);
  float* G = (float*)take((size_t)R * K1p * 4);
  float* norms = (float*)take((size_t)R * 4);
  float* su = (float*)take((size_t)R * 4);
  float* sx = (float*)take((size_t)R * 4);
  float* partial = (float*)take(1184 * 4);
  float* norm = (float*)take(4);
  float* logits = scores ? scores : lg_ws;
  const float inv_T = 1.f / temperature;

  clip_head_prep_kernel<<<ceil_div(K1 * 32, 256), 256, 0, stream>>>(w, w_bg, K, D, wall);
  count_launch();
  // GEMM 1: logits[R, K1] = x . wall^T, scaled per row by 1/(T |x|)
  TcEpilogue e1 = {};
  e1.mode = 0;
  e1.out = logits;
  e1.ldo = K1;
  e1.n_valid = K1;
  e1.inv_T = inv_T;
  e1.norms = norms;
  int rc = launch_gemm(x, R, D, wall, K1, D, D, e1, stream);
  if (rc) return rc;
  if (op == 0) return 0;

  const bool want_dx = dx != nullptr;
  if (op == 1) {
    if (stats) {
      cudaError_t e = cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), stream);
      if (e != cudaSuccess) return (int)e;
    }
    TcRowArgs a = {};
    a.logits = logits;
    a.gt = gt;
    a.norms = norms;
    a.grad_scale = grad_scale;
    a.norm_host = (float)R;
    if (loss_mode == CDDMSL_LOSS_WEIGHTED_CE) {
      clip_head_wsum_kernel<<<1, 1024, 0, stream>>>(gt, R, K, bg_weight, norm);
      count_launch();
      a.norm_dev = norm;
    }
    a.R = R;
    a.K1 = K1;
    a.K1p = K1p;
    a.inv_T = inv_T;
    a.loss_mode = loss_mode;
    a.gamma = gamma;
    a.bg_weight = bg_weight;
    a.strict_nan = strict_nan;
    a.G = want_dx ? G : nullptr;
    a.su = su;
    a.sx = sx;
    a.partial = partial;
    a.stats = stats;
    const int grid = min(ceil_div(R, 8), 1184);
    clip_head_tc_rows_kernel<<<grid, 256, 0, stream>>>(a);
    count_launch();
    clip_head_finish_kernel<<<1, 256, 0, stream>>>(partial, grid, a.norm_dev, a.norm_host, loss);
    count_launch();
    if (!want_dx) return (int)cudaGetLastError();
  } else {
    return CDDMSL_EINVAL;  // dscores-driven backward stays on the CUDA-core kernel
  }
  (void)dscores;
  // GEMM 2: dx[R, D] = G . wall (= G . wallT^T), then dx = acc * su - x * sx
  clip_head_tc_transpose_w_kernel<<<ceil_div(D * K1p, 256), 256, 0, stream>>>(wall, wallT, K1, K1p, D);
  count_launch();
  TcEpilogue e2 = {};
  e2.mode = 1;
  e2.out = dx;
  e2.ldo = D;
  e2.n_valid = D;
  e2.x = x;
  e2.su = su;
  e2.sx = sx;
  return launch_gemm(G, R, K1p, wallT, D, K1p, K1p, e2, stream);
}

}  // namespace cddmsl
