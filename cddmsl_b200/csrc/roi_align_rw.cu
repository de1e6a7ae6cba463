// ROIAlign backward, "row walk" formulation (14x14 pooler, C % 32 == 0) -- round 2.
//
// The channels-last backward of roi_align_cl.cu gives a warp two bin rows of 64 channels; every column it leaves
// is scattered into up to four map rows right away, so a RoI issues ~2x the red.global traffic its footprint needs
// (7 row pairs x their slots; ncu: L1->XBAR request port 74 % busy, L2 61 %), and the x-walk with its
// data-dependent flush runs 14 times per RoI (issue slots 70 % busy, 72 % of the instructions are control).
//
// Here ONE warp owns a (RoI, 32 channels) unit, lanes on channels as before (coalesced 128-byte REDs into the
// channels-last gradient map, warp-uniform control), but the two separable passes are ordered the other way round:
//   1. vertical pass over the 14 bin rows with the 14 bins of a row in REGISTERS: a rolling window of three
//      map rows  R[y][pw] += wy * g[ph][pw]  (28 independent FFMAs per y-sample, no flush logic at all);
//   2. whenever two map rows are complete they are emitted together: ONE x-walk over the 14 x gw samples feeds both
//      rows, the flush of a finished column is a RED pair -- every (row, column, channel) of the footprint
//      receives exactly one RED.
// Tap weights, flush flags and the emit schedule are data-independent per RoI: a plan kernel writes them once as a
// 2.6 KB record that all 32 channel groups of the RoI read (the kernel is a straight interpreter of that record).
// The grad tile [32 ch][196] streams through a per-warp RING of three TMA boxes of [32 channels][one bin-row pair =
// 112 bytes] (a lane's row starts 28 words after its neighbour's, so its reads spread over all banks), each box on its
// own mbarrier: the vertical pass consumes the pairs in order, a finished box is refilled at once with the pair three
// ahead (of this unit or the next), so loads run two to three pairs ahead of the math without a CTA-wide barrier and
// a warp needs 13 KB of shared memory -- 16 independent warps per SM.
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"
#include "roi_common.cuh"
#include "tma_host.cuh"

namespace cddmsl {

int launch_transpose(const float* in, float* out, int N, int A, int B, cudaStream_t stream);

namespace {

constexpr int kRwP = 14;
#ifndef CDDMSL_RW_WARPS
#define CDDMSL_RW_WARPS 16
#endif
constexpr int kRwWarps = CDDMSL_RW_WARPS;
constexpr int kRwMaxG = 10;                      // samples per bin and axis the record holds (RoIs up to 140 cells)
constexpr int kRwPairs = kRwP / 2;               // bin-row pairs per unit
constexpr int kRwNB = 3;                         // ring depth (boxes in flight per warp)
constexpr int kRwPairFloats = 2 * kRwP;          // one TMA box row: two bin rows of one channel, 112 bytes
constexpr int kRwBoxFloats = 32 * kRwPairFloats; // box = [32 channels][28 floats] = 3584 bytes
constexpr int kRwXs = kRwP * kRwMaxG;            // 140 x-samples
constexpr int kRwYs = kRwP * kRwMaxG + 4;        // 140 y-samples + 2 sentinels, padded
constexpr int kRwHdrBytes = 64;
constexpr int kRwOffXw = kRwHdrBytes;                         // float2[140]
constexpr int kRwOffYw = kRwOffXw + kRwXs * 8;                // float2[144]
constexpr int kRwOffXf = kRwOffYw + kRwYs * 8;                // uint8[144] (140 used)
constexpr int kRwOffYc = kRwOffXf + 144;                      // uint8[144]
constexpr int kRwRecBytes = kRwOffYc + 144;                   // 2624 = 16 * 164
static_assert(kRwRecBytes % 16 == 0, "record is moved with one bulk copy");
constexpr int kRwWarpBytes = kRwNB * kRwBoxFloats * 4 + kRwRecBytes + 64;  // 13 440 = 128 * 105
static_assert(kRwWarpBytes % 128 == 0, "TMA destinations must be 128-byte aligned");
constexpr int kRwBars = kRwNB + 1;
constexpr int kRwSmemBytes = kRwWarps * kRwWarpBytes + kRwWarps * kRwBars * 8;
constexpr int kRwNoTensorMap = -1000;  // cuTensorMapEncodeTiled unavailable: the caller falls back

enum : int { kRwSkip = 0, kRwFast = 1, kRwSlow = 2 };
// y-sample codes
enum : unsigned { kYEmit = 1u, kYTwo = 2u, kYPos1 = 4u, kYNoUpd = 8u, kYNewRow = 16u };

struct RwHdr {       // 64 bytes
  int kind;          // kRwSkip / kRwFast / kRwSlow
  int batch;
  int x0, y0;        // first column / row touched
  int gw, gh;
  int ny;            // y codes incl. sentinels
  int xtail;         // the column after the last low column receives weight too
  unsigned xmask;    // flush-before flags of the first 32 x-samples (all of them for gw <= 2)
  int pad[7];
};
static_assert(sizeof(RwHdr) == kRwHdrBytes, "header layout");

__device__ __forceinline__ uint32_t rw_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rw_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(rw_smem(bar)));
}
// barriers and copy destinations are passed as 32-bit shared-window addresses computed once per warp
__device__ __forceinline__ void rw_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rw_tma_box(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void rw_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void rw_wait(uint32_t a, uint32_t parity) {
#pragma unroll 1
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
__device__ __forceinline__ void rw_red(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void rw_red_if(bool on, float* p, float v) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.u32 q, %2, 0;\n\t"
      "@q red.global.add.f32 [%0], %1;\n\t}" ::"l"(p),
      "f"(v), "r"((unsigned)on)
      : "memory");
}

// same validity rule and weights as make_tap (roi_common.cuh); lo < 0: the sample contributes nothing
struct RwTap {
  int lo, hi;
  float wl, wh;
};
__device__ __forceinline__ RwTap rw_tap(float start, float bin, int p, int i, int g, int L, float wscale) {
  const Tap t = make_tap(start, bin, p, i, g, L, 0);
  RwTap e;
  const bool dead = (t.wl == 0.f && t.wh == 0.f);
  e.lo = dead ? -1 : t.lo;
  e.hi = t.hi;
  e.wl = t.wl * wscale;
  e.wh = t.hi != t.lo ? t.wh * wscale : 0.f;
  return e;
}

// ------------------------------------------------------------------------------------------------
// plan: one thread per RoI writes its record
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rw_plan_kernel(const float* __restrict__ rois, unsigned char* __restrict__ recs,
                                                      int N, int H, int W, int R, float scale, int sampling_ratio,
                                                      int aligned) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  unsigned char* rec = recs + (size_t)r * kRwRecBytes;
  RwHdr h;
  memset(&h, 0, sizeof(h));
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, kRwP, kRwP, sampling_ratio, H, W);
  h.batch = g.batch;
  h.gw = g.gw;
  h.gh = g.gh;
  h.kind = kRwFast;
  if (g.gw <= 0 || g.gh <= 0 || g.batch < 0 || g.batch >= N) h.kind = kRwSkip;
  else if (g.gw > kRwMaxG || g.gh > kRwMaxG) h.kind = kRwSlow;
  if (h.kind == kRwFast) {
    float2* xw = reinterpret_cast<float2*>(rec + kRwOffXw);
    unsigned char* xf = rec + kRwOffXf;
    bool mono = true;
    int x0 = -1, prev = 0, xtail = 0;
    unsigned xmask = 0;
    const int nx = kRwP * g.gw;
    for (int t = 0; t < nx; ++t) {
      const int p = t / g.gw, i = t - p * g.gw;
      const RwTap e = rw_tap(g.sw, g.bw, p, i, g.gw, W, 1.f);
      int adv = 0;
      float2 w = make_float2(0.f, 0.f);
      if (e.lo >= 0) {
        if (x0 < 0) x0 = e.lo;
        else {
          adv = e.lo - prev;
          if (adv < 0 || adv > 1) mono = false;
        }
        prev = e.lo;
        xtail = e.hi != e.lo;
        w = make_float2(e.wl, e.wh);
      }
      xw[t] = w;
      xf[t] = (unsigned char)(adv == 1);
      if (adv == 1 && t < 32) xmask |= 1u << t;
    }
    float2* yw = reinterpret_cast<float2*>(rec + kRwOffYw);
    unsigned char* yc = rec + kRwOffYc;
    int y0 = -1, prevd = 0, n = 0;
    const int nys = kRwP * g.gh;
    for (int t = 0; t < nys; ++t) {
      const int p = t / g.gh, i = t - p * g.gh;
      const RwTap e = rw_tap(g.sh, g.bh, p, i, g.gh, H, g.inv_count);
      unsigned code = i == 0 ? kYNewRow : 0u;
      float2 w = make_float2(0.f, 0.f);
      if (e.lo < 0) code |= kYNoUpd;
      else {
        int d = 0;
        if (y0 < 0) y0 = e.lo;
        else d = e.lo - y0;
        if (d - prevd < 0 || d - prevd > 1) mono = false;
        if ((d >> 1) > (prevd >> 1)) code |= kYEmit | kYTwo;  // rows cur, cur+1 are complete (both < lo <= H-1)
        if (d & 1) code |= kYPos1;
        prevd = d;
        w = make_float2(e.wl, e.wh);
      }
      yw[n] = w;
      yc[n] = (unsigned char)code;
      ++n;
    }
    if (x0 < 0 || y0 < 0) h.kind = kRwSkip;
    else if (!mono) h.kind = kRwSlow;
    else {
      // sentinels: flush what is still in the window
      const int cur = y0 + 2 * (prevd >> 1);
      yw[n] = make_float2(0.f, 0.f);
      yc[n++] = (unsigned char)(kYEmit | kYNoUpd | (cur + 1 < H ? kYTwo : 0u));
      if ((prevd & 1) && cur + 2 < H) {
        yw[n] = make_float2(0.f, 0.f);
        yc[n++] = (unsigned char)(kYEmit | kYNoUpd);
      }
      h.x0 = x0;
      h.y0 = y0;
      h.ny = n;
      h.xtail = xtail;
      h.xmask = xmask;
    }
  }
  *reinterpret_cast<RwHdr*>(rec) = h;
}

// ------------------------------------------------------------------------------------------------
// emission of two finished map rows: one x-walk, one RED per (row, column, channel)
// ------------------------------------------------------------------------------------------------
#define CDDMSL_RW_XSTEP(FLAG, WV, PW)            \
  if (FLAG) {                                    \
    rw_red(p, a0);                               \
    if (two) rw_red(p + WC, b0);                 \
    p += C;                                      \
    a0 = a1;                                     \
    b0 = b1;                                     \
    a1 = 0.f;                                    \
    b1 = 0.f;                                    \
  }                                              \
  a0 = fmaf((WV).x, Ra[PW], a0);                 \
  a1 = fmaf((WV).y, Ra[PW], a1);                 \
  b0 = fmaf((WV).x, Rb[PW], b0);                 \
  b1 = fmaf((WV).y, Rb[PW], b1);

template <int GW>
__device__ __forceinline__ void rw_emit(const float (&Ra)[kRwP], const float (&Rb)[kRwP], float* p, bool two, int WC,
                                        int C, int gw, const float2* __restrict__ xw,
                                        const unsigned char* __restrict__ xf, unsigned xmask, bool xtail) {
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  if (GW == 1) {
#pragma unroll
    for (int pw = 0; pw < kRwP; ++pw) {
      const float2 w = xw[pw];
      CDDMSL_RW_XSTEP(xmask & (1u << pw), w, pw)
    }
  } else if (GW == 2) {
#pragma unroll
    for (int pw = 0; pw < kRwP; ++pw) {
      const float4 w = reinterpret_cast<const float4*>(xw)[pw];
      const float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
      CDDMSL_RW_XSTEP(xmask & (1u << (2 * pw)), w0, pw)
      CDDMSL_RW_XSTEP(xmask & (1u << (2 * pw + 1)), w1, pw)
    }
  } else {
    int t = 0;
#pragma unroll
    for (int pw = 0; pw < kRwP; ++pw) {
#pragma unroll 1
      for (int s = 0; s < gw; ++s, ++t) {
        const float2 w = xw[t];
        const unsigned f = xf[t];
        CDDMSL_RW_XSTEP(f, w, pw)
      }
    }
  }
  rw_red(p, a0);
  if (two) rw_red(p + WC, b0);
  if (xtail) {
    rw_red(p + C, a1);
    if (two) rw_red(p + C + WC, b1);
  }
}

// ------------------------------------------------------------------------------------------------
// main kernel: persistent, 16 independent warps per CTA
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRwWarps * 32, 1)
roi_align_bwd_rw_kernel(const __grid_constant__ CUtensorMap gmap, const float* __restrict__ rois,
                        const unsigned char* __restrict__ recs, float* __restrict__ gt, unsigned* __restrict__ counter,
                        int N, int C, int H, int W, int R, float scale, int sampling_ratio, int aligned, int vpr,
                        int kgroups) {
  extern __shared__ __align__(128) unsigned char rw_smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(rw_smem_raw + (size_t)warp * kRwWarpBytes);  // [kRwNB][32][28]
  unsigned char* rec = reinterpret_cast<unsigned char*>(ring + kRwNB * kRwBoxFloats);
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(rw_smem_raw + (size_t)kRwWarps * kRwWarpBytes) + warp * kRwBars;
  const uint32_t bars = rw_smem(bar_ptr), barR = bars + kRwNB * 8u;   // slot b: bars + 8 b
  const uint32_t ring_s = rw_smem(ring), rec_s = rw_smem(rec);
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kRwBars; ++b) rw_mbar_init(bar_ptr + b);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const unsigned nvisits = (unsigned)R * (unsigned)vpr;
  const int ngroups = C / 32;
  const int WC = W * C;
  uint32_t pr = 0;
  int cslot = 0;        // ring slot of the pair consumed next
  uint32_t cpar = 0;    // its phase parity

  auto fetch = [&]() -> unsigned {
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(counter, 1u);
    return __shfl_sync(0xffffffffu, v, 0);
  };
  auto issue_rec = [&](unsigned v) {
    if (lane == 0) {
      rw_expect(barR, kRwRecBytes);
      rw_bulk(rec_s, recs + (size_t)(v / (unsigned)vpr) * kRwRecBytes, kRwRecBytes, barR);
    }
  };
  auto issue_box = [&](int slot, int r, int cg, int pair) {
    if (lane == 0) {
      rw_expect(bars + 8u * slot, (uint32_t)(kRwBoxFloats * 4));
      rw_tma_box(ring_s + (uint32_t)(slot * kRwBoxFloats * 4), &gmap, pair * kRwPairFloats, r * C + cg * 32,
                 bars + 8u * slot);
    }
  };

  unsigned v = fetch();
  if (v < nvisits) {
    const int r = (int)(v / (unsigned)vpr), cg = (int)(v % (unsigned)vpr) * kgroups;
    issue_rec(v);
#pragma unroll
    for (int b = 0; b < kRwNB; ++b) issue_box(b, r, cg, b);
  }
  while (v < nvisits) {
    const unsigned vn = fetch();
    const int r = (int)(v / (unsigned)vpr), cgbase = (int)(v % (unsigned)vpr) * kgroups;
    const int nk = min(kgroups, ngroups - cgbase);
    rw_wait(barR, pr);
    pr ^= 1u;
    const RwHdr* hp = reinterpret_cast<const RwHdr*>(rec);
    const int kind = hp->kind, batch = hp->batch, x0 = hp->x0, y0 = hp->y0, gw = hp->gw, ny = hp->ny;
    const bool xtail = hp->xtail != 0;
    const unsigned xmask = hp->xmask;
    const float2* xw = reinterpret_cast<const float2*>(rec + kRwOffXw);
    const float2* yw = reinterpret_cast<const float2*>(rec + kRwOffYw);
    const unsigned char* xf = rec + kRwOffXf;
    const unsigned char* yc = rec + kRwOffYc;
    for (int k = 0; k < nk; ++k) {
      const int cg = cgbase + k;
      bool has_next = true;
      int rn = r, cgn = cg + 1;
      if (k + 1 >= nk) {
        has_next = vn < nvisits;
        rn = (int)(vn / (unsigned)vpr);
        cgn = (int)(vn % (unsigned)vpr) * kgroups;
      }
      // pair `p` of this unit has been consumed by every lane: refill its slot with the pair kRwNB ahead
      auto release = [&](int p) {
        __syncwarp();
        const int np = p + kRwNB;
        if (np < kRwPairs) issue_box(cslot, r, cg, np);
        else if (has_next) issue_box(cslot, rn, cgn, np - kRwPairs);
        if (++cslot == kRwNB) {
          cslot = 0;
          cpar ^= 1u;
        }
      };
      float* img = gt + (size_t)batch * H * WC + (size_t)cg * 32 + lane;
      if (kind == kRwFast) {
        float R0[kRwP], R1[kRwP], R2[kRwP], g[kRwP];
#pragma unroll
        for (int j = 0; j < kRwP; ++j) R0[j] = R1[j] = R2[j] = g[j] = 0.f;
        float* rowp = img + ((size_t)y0 * W + x0) * C;
        int brow = -1;
#pragma unroll 1
        for (int i = 0; i < ny; ++i) {
          const unsigned code = yc[i];
          const float2 w = yw[i];
          if (code & (kYNewRow | kYEmit)) {
            if (code & kYNewRow) {
              ++brow;
              if (!(brow & 1)) {
                if (brow) release((brow >> 1) - 1);
                rw_wait(bars + 8u * cslot, cpar);
              }
              const float2* gp = reinterpret_cast<const float2*>(ring + cslot * kRwBoxFloats + lane * kRwPairFloats +
                                                                 (brow & 1) * kRwP);
#pragma unroll
              for (int j = 0; j < kRwP / 2; ++j) {
                const float2 t = gp[j];
                g[2 * j] = t.x;
                g[2 * j + 1] = t.y;
              }
            }
            if (code & kYEmit) {
              const bool two = (code & kYTwo) != 0;
              if (gw == 1) rw_emit<1>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
              else if (gw == 2) rw_emit<2>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
              else rw_emit<0>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
#pragma unroll
              for (int j = 0; j < kRwP; ++j) {
                R0[j] = R2[j];
                R1[j] = 0.f;
                R2[j] = 0.f;
              }
              rowp += 2 * (size_t)WC;
            }
          }
          // samples outside the map and the sentinels carry zero weights
          if (code & kYPos1) {
#pragma unroll
            for (int j = 0; j < kRwP; ++j) {
              R1[j] = fmaf(w.x, g[j], R1[j]);
              R2[j] = fmaf(w.y, g[j], R2[j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < kRwP; ++j) {
              R0[j] = fmaf(w.x, g[j], R0[j]);
              R1[j] = fmaf(w.y, g[j], R1[j]);
            }
          }
        }
        release(kRwPairs - 1);
      } else if (kind == kRwSlow) {
        // reference-shaped scatter (huge or non-monotone sampling grids), lanes on channels
        const RoiGeom gm = roi_geom(rois + (size_t)r * 5, scale, aligned, kRwP, kRwP, sampling_ratio, H, W);
        for (int pair = 0; pair < kRwPairs; ++pair) {
          rw_wait(bars + 8u * cslot, cpar);
          const float* src = ring + cslot * kRwBoxFloats + lane * kRwPairFloats;
          for (int e = 0; e < kRwPairFloats; ++e) {
            const int ph = 2 * pair + e / kRwP, pw = e % kRwP;
            const float go = src[e] * gm.inv_count;
            for (int iy = 0; iy < gm.gh; ++iy) {
              const Tap ty = make_tap(gm.sh, gm.bh, ph, iy, gm.gh, H, 0);
              if (ty.wl == 0.f && ty.wh == 0.f) continue;
              for (int ix = 0; ix < gm.gw; ++ix) {
                const Tap tx = make_tap(gm.sw, gm.bw, pw, ix, gm.gw, W, 0);
                if (tx.wl == 0.f && tx.wh == 0.f) continue;
                rw_red(img + ((size_t)ty.lo * W + tx.lo) * C, go * ty.wl * tx.wl);
                rw_red(img + ((size_t)ty.lo * W + tx.hi) * C, go * ty.wl * tx.wh);
                rw_red(img + ((size_t)ty.hi * W + tx.lo) * C, go * ty.wh * tx.wl);
                rw_red(img + ((size_t)ty.hi * W + tx.hi) * C, go * ty.wh * tx.wh);
              }
            }
          }
          release(pair);
        }
      } else {
        for (int pair = 0; pair < kRwPairs; ++pair) {
          rw_wait(bars + 8u * cslot, cpar);
          release(pair);
        }
      }
    }
    __syncwarp();
    if (vn < nvisits) issue_rec(vn);
    v = vn;
  }
}

int g_roi_rw = 1;      // tuning knob "roi_rw": 1 = use this backward where eligible
int g_roi_rw_k = 4;    // channel groups (of 32) per visit of a RoI ("roi_rw_k")

}  // namespace

int tune_roi_rw(const char* key, int value) {
  if (!strcmp(key, "roi_rw")) g_roi_rw = value;
  else if (!strcmp(key, "roi_rw_k")) g_roi_rw_k = value > 0 ? value : 1;
  else return 0;
  return 1;
}

bool roi_rw_eligible(int N, int C, int H, int W, int R, int P) {
  (void)N;
  return g_roi_rw && P == kRwP && C > 0 && C % 32 == 0 && (long long)H * W * C < 0x7fffffffLL &&
         (long long)R * (C / 32) < 0x7fffffffLL;
}

size_t roi_rw_workspace_bytes(int N, int C, int H, int W, int R) {
  return align_up((size_t)N * C * H * W * 4, 256) + align_up((size_t)R * kRwRecBytes, 256) + 256;
}

int roi_align_bwd_rw(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, void* ws, cudaStream_t stream) {
  (void)P;
  CUtensorMap gmap;  // gout as [R*C rows][196]: box = 32 channels x one bin-row pair (112 bytes)
  if (tma_encode_2d_f32(&gmap, gout, kRwP * kRwP, (unsigned long long)R * C, kRwP * kRwP * 4, kRwPairFloats, 32,
                        CU_TENSOR_MAP_SWIZZLE_NONE))
    return kRwNoTensorMap;
  float* gt = reinterpret_cast<float*>(ws);
  const size_t map_bytes = align_up((size_t)N * C * H * W * 4, 256);
  unsigned* counter = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(ws) + map_bytes);
  unsigned char* recs = reinterpret_cast<unsigned char*>(counter) + 256;
  cudaError_t e = cudaMemsetAsync(gt, 0, map_bytes + 256, stream);  // gradient map and the work counter
  if (e != cudaSuccess) return (int)e;
  rw_plan_kernel<<<ceil_div(R, 128), 128, 0, stream>>>(rois, recs, N, H, W, R, scale, sampling_ratio, aligned);
  count_launch();
  CDDMSL_CHECK_LAUNCH();
  const int ngroups = C / 32;
  const int kgroups = min(g_roi_rw_k, ngroups);
  const int vpr = ceil_div(ngroups, kgroups);
  auto k = roi_align_bwd_rw_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kRwSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long nvisits = (long long)R * vpr;
  const long long want = (nvisits + kRwWarps - 1) / kRwWarps;
  const int grid = (int)(want < (long long)sm_count() ? want : (long long)sm_count());
  k<<<grid, kRwWarps * 32, kRwSmemBytes, stream>>>(gmap, rois, recs, gt, counter, N, C, H, W, R, scale,
                                                   sampling_ratio, aligned, vpr, kgroups);
  count_launch();
  CDDMSL_CHECK_LAUNCH();
  return launch_transpose(gt, gin, N, H * W, C, stream);  // NHWC -> NCHW (overwrites gin completely)
}

}  // namespace cddmsl
