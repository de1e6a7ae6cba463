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
// CPL channels per lane.  CPL = 1 (default): 16 warps x ring of 3 boxes (128 registers; measured alternatives on the
// VOC shape: 12 warps x ring of 4 at 141 registers 2.23 ms, 20 warps x ring of 2 at 96 registers 2.16 ms, this 1.97 ms);  CPL = 2 (C % 64 == 0, knob):
// every tap, flag test and branch serves two channels -- 9 warps x ring of 3 boxes of 64 channels (168 registers).
template <int CPL>
struct RwCfg;
constexpr int kRwMaxG = 10;                      // samples per bin and axis the record holds (RoIs up to 140 cells)
constexpr int kRwPairs = kRwP / 2;               // bin-row pairs per unit
constexpr int kRwPairFloats = 2 * kRwP;          // one TMA box row: two bin rows of one channel, 112 bytes
constexpr int kRwBoxFloats = 32 * kRwPairFloats; // box = [32 channels][28 floats] = 3584 bytes (per channel of a lane)
constexpr int kRwXs = kRwP * kRwMaxG;            // 140 x-samples
constexpr int kRwYs = kRwP * kRwMaxG + 4;        // 140 y-samples + 2 sentinels, padded
constexpr int kRwHdrBytes = 64;
constexpr int kRwOffXw = kRwHdrBytes;                         // float2[140]
constexpr int kRwOffYw = kRwOffXw + kRwXs * 8;                // float2[144]
constexpr int kRwOffXf = kRwOffYw + kRwYs * 8;                // uint8[144] (140 used)
constexpr int kRwOffYc = kRwOffXf + 144;                      // uint8[144]
constexpr int kRwRecBytes = kRwOffYc + 144;                   // 2624 = 16 * 164
static_assert(kRwRecBytes % 16 == 0, "record is moved with one bulk copy");
template <>
struct RwCfg<1> {
  static constexpr int kWarps = 16, kNB = 3;
};
template <>
struct RwCfg<2> {
  static constexpr int kWarps = 9, kNB = 3;
};
template <int CPL>
struct RwLayout {
  static constexpr int kWarps = RwCfg<CPL>::kWarps, kNB = RwCfg<CPL>::kNB;
  static constexpr int kBoxFloats = CPL * kRwBoxFloats;                        // [32 * CPL channels][28]
  static constexpr int kWarpBytes = kNB * kBoxFloats * 4 + kRwRecBytes + 64;   // multiple of 128
  static constexpr int kBars = kNB + 1;
  static constexpr int kSmemBytes = kWarps * kWarpBytes + kWarps * kBars * 8;
  static_assert(kWarpBytes % 128 == 0, "TMA destinations must be 128-byte aligned");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};
constexpr int kRwNoTensorMap = -1000;  // cuTensorMapEncodeTiled unavailable: the caller falls back

enum : int { kRwSkip = 0, kRwFast = 1, kRwSlow = 2 };
// y-sample codes
enum : unsigned { kYEmit = 1u, kYTwo = 2u, kYPos1 = 4u, kYNoUpd = 8u, kYNewRow = 16u, kYSkip = 32u };

struct RwHdr {       // 64 bytes
  int kind;          // kRwSkip / kRwFast / kRwSlow
  int batch;
  int x0, y0;        // first column / row touched
  int gw, gh;
  int ny;            // y codes incl. sentinels
  int xtail;         // the column after the last low column receives weight too
  unsigned xmask;    // flush-before flags of the first 32 x-samples (all of them for gw <= 2)
  int xsimple;       // every x-sample advances by at most one column (always true with adaptive sampling)
  int pad[6];
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
// pull a box into L2 ahead of the copy that will need it (no shared memory, no barrier)
__device__ __forceinline__ void rw_tma_prefetch(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
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
    bool ok = true, xsimple = true;
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
          if (adv < 0 || adv > 255) ok = false;
          if (adv > 1) xsimple = false;
        }
        prev = e.lo;
        xtail = e.hi != e.lo;
        w = make_float2(e.wl, e.wh);
      }
      xw[t] = w;
      xf[t] = (unsigned char)adv;  // columns to advance before this sample (fixed sampling grids: may exceed 1)
      if (adv == 1 && t < 32) xmask |= 1u << t;
    }
    float2* yw = reinterpret_cast<float2*>(rec + kRwOffYw);
    unsigned char* yc = rec + kRwOffYc;
    int y0 = -1, wb = 0, lastd = 0, n = 0;   // wb: row (relative to y0, even) the window starts at
    const int nys = kRwP * g.gh;
    for (int t = 0; t < nys && ok; ++t) {
      const int p = t / g.gh, i = t - p * g.gh;
      const RwTap e = rw_tap(g.sh, g.bh, p, i, g.gh, H, g.inv_count);
      unsigned code = i == 0 ? kYNewRow : 0u;
      float2 w = make_float2(0.f, 0.f);
      if (n + 4 > kRwYs - 2) {  // keep room for the sentinels
        ok = false;
        break;
      }
      if (e.lo < 0) code |= kYNoUpd;
      else {
        int d = 0;
        if (y0 < 0) y0 = e.lo;
        else d = e.lo - y0;
        if (d < wb) ok = false;
        // the window [wb, wb+2] has to reach row d: k two-row shifts.  The first two may still hold weight and are
        // emitted (rows below lo <= H-1, so both exist); any further ones only move the row pointer.
        const int k = (d >> 1) - (wb >> 1);
        if (k == 1) code |= kYEmit | kYTwo;
        else if (k >= 2) {
          yw[n] = make_float2(0.f, 0.f);
          yc[n++] = (unsigned char)(kYEmit | kYTwo | kYNoUpd);
          if (k == 2) code |= kYEmit | kYTwo;
          else {
            yw[n] = make_float2(0.f, 0.f);
            yc[n++] = (unsigned char)(kYEmit | kYTwo | kYNoUpd);
            yw[n] = make_float2(__int_as_float(2 * (k - 2)), 0.f);
            yc[n++] = (unsigned char)(kYSkip | kYNoUpd);
          }
        }
        wb += 2 * max(k, 0);
        if (d & 1) code |= kYPos1;
        lastd = d;
        w = make_float2(e.wl, e.wh);
      }
      yw[n] = w;
      yc[n] = (unsigned char)code;
      ++n;
    }
    if (x0 < 0 || y0 < 0) h.kind = kRwSkip;
    else if (!ok) h.kind = kRwSlow;
    else {
      // sentinels: flush what is still in the window
      const int cur = y0 + wb;
      yw[n] = make_float2(0.f, 0.f);
      yc[n++] = (unsigned char)(kYEmit | kYNoUpd | (cur + 1 < H ? kYTwo : 0u));
      if ((lastd & 1) && cur + 2 < H) {
        yw[n] = make_float2(0.f, 0.f);
        yc[n++] = (unsigned char)(kYEmit | kYNoUpd);
      }
      h.x0 = x0;
      h.y0 = y0;
      h.ny = n;
      h.xtail = xtail;
      h.xmask = xmask;
      h.xsimple = xsimple;
    }
  }
  *reinterpret_cast<RwHdr*>(rec) = h;
}

// ------------------------------------------------------------------------------------------------
// emission of two finished map rows: one x-walk, one RED per (row, column, channel)
// ------------------------------------------------------------------------------------------------
#define CDDMSL_RW_FLUSH()                                  \
  {                                                        \
    _Pragma("unroll") for (int q = 0; q < CPL; ++q) {      \
      rw_red(p + 32 * q, a0[q]);                           \
      if (two) rw_red(p + WC + 32 * q, b0[q]);             \
      a0[q] = a1[q];                                       \
      b0[q] = b1[q];                                       \
      a1[q] = 0.f;                                         \
      b1[q] = 0.f;                                         \
    }                                                      \
    p += C;                                                \
  }
#define CDDMSL_RW_FMAS(WV, PW)                             \
  _Pragma("unroll") for (int q = 0; q < CPL; ++q) {        \
    a0[q] = fmaf((WV).x, Ra[q][PW], a0[q]);                \
    a1[q] = fmaf((WV).y, Ra[q][PW], a1[q]);                \
    b0[q] = fmaf((WV).x, Rb[q][PW], b0[q]);                \
    b1[q] = fmaf((WV).y, Rb[q][PW], b1[q]);                \
  }
#define CDDMSL_RW_XSTEP(FLAG, WV, PW) \
  if (FLAG) CDDMSL_RW_FLUSH()         \
  CDDMSL_RW_FMAS(WV, PW)

template <int GW, int CPL>
__device__ __forceinline__ void rw_emit(const float (&Ra)[CPL][kRwP], const float (&Rb)[CPL][kRwP], float* p, bool two,
                                        int WC, int C, int gw, const float2* __restrict__ xw,
                                        const unsigned char* __restrict__ xf, unsigned xmask, bool xtail) {
  float a0[CPL], a1[CPL], b0[CPL], b1[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) a0[q] = a1[q] = b0[q] = b1[q] = 0.f;
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
  } else if (GW == 0) {
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
  } else {
    // fixed sampling grids on wide RoIs: a sample may advance by several columns (kept apart from the walk above:
    // the second flush costs the common case 25 % when it sits in the same loop)
    int t = 0;
#pragma unroll
    for (int pw = 0; pw < kRwP; ++pw) {
#pragma unroll 1
      for (int s = 0; s < gw; ++s, ++t) {
        const float2 w = xw[t];
        const unsigned f = xf[t];
        if (f) {
          CDDMSL_RW_FLUSH()
          if (f > 1) {             // the high column of the previous sample, then f - 2 untouched columns
            CDDMSL_RW_FLUSH()
            p += (size_t)(f - 2) * C;
          }
        }
        CDDMSL_RW_FMAS(w, pw)
      }
    }
  }
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    rw_red(p + 32 * q, a0[q]);
    if (two) rw_red(p + WC + 32 * q, b0[q]);
    if (xtail) {
      rw_red(p + C + 32 * q, a1[q]);
      if (two) rw_red(p + C + WC + 32 * q, b1[q]);
    }
  }
}

// one bin-row pair of a RoI whose sampling grid does not fit the record: reference-shaped scatter, lanes on channels
template <int CPL>
__device__ __forceinline__ void rw_slow_pair(const float* box, float* img, const RoiGeom& gm, int pair, int H, int W,
                                          int C) {
  const int lane = threadIdx.x & 31;
  for (int eq = 0; eq < kRwPairFloats * CPL; ++eq) {
    const int q = eq / kRwPairFloats, e = eq - q * kRwPairFloats;
    const float* src = box + (lane + 32 * q) * kRwPairFloats;
    float* img_q = img + 32 * q;
    const int ph = 2 * pair + e / kRwP, pw = e % kRwP;
    const float go = src[e] * gm.inv_count;
    for (int iy = 0; iy < gm.gh; ++iy) {
      const Tap ty = make_tap(gm.sh, gm.bh, ph, iy, gm.gh, H, 0);
      if (ty.wl == 0.f && ty.wh == 0.f) continue;
      for (int ix = 0; ix < gm.gw; ++ix) {
        const Tap tx = make_tap(gm.sw, gm.bw, pw, ix, gm.gw, W, 0);
        if (tx.wl == 0.f && tx.wh == 0.f) continue;
        rw_red(img_q + ((size_t)ty.lo * W + tx.lo) * C, go * ty.wl * tx.wl);
        rw_red(img_q + ((size_t)ty.lo * W + tx.hi) * C, go * ty.wl * tx.wh);
        rw_red(img_q + ((size_t)ty.hi * W + tx.lo) * C, go * ty.wh * tx.wl);
        rw_red(img_q + ((size_t)ty.hi * W + tx.hi) * C, go * ty.wh * tx.wh);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// main kernel: persistent, independent warps
// ------------------------------------------------------------------------------------------------
// FIXED: the call uses a fixed sampling grid (sampling_ratio > 0), where a sample may advance by several columns; the
// walk for that lives only in this instantiation -- the adaptive-sampling kernel (the reference's setting) has to keep
// its hot code inside the 32 KB instruction cache (measured: 2.10 ms without, 2.34 ms with that walk compiled in;
// likewise an unrolled walk for three samples per bin: 1.96 -> 2.39 ms; dropping the two-sample one: no change).
template <int CPL, bool FIXED>
__global__ void __launch_bounds__(RwCfg<CPL>::kWarps * 32, 1)
roi_align_bwd_rw_kernel(const __grid_constant__ CUtensorMap gmap, const float* __restrict__ rois,
                        const unsigned char* __restrict__ recs, float* __restrict__ gt, unsigned* __restrict__ counter,
                        int N, int C, int H, int W, int R, float scale, int sampling_ratio, int aligned, int vpr,
                        int kgroups) {
  using L = RwLayout<CPL>;
  constexpr int kNB = L::kNB, kBoxFloats = L::kBoxFloats, GC = 32 * CPL;
  extern __shared__ __align__(128) unsigned char rw_smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(rw_smem_raw + (size_t)warp * L::kWarpBytes);  // [kNB][32 * CPL][28]
  unsigned char* rec = reinterpret_cast<unsigned char*>(ring + kNB * kBoxFloats);
  uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(rw_smem_raw + (size_t)L::kWarps * L::kWarpBytes) + warp * L::kBars;
  const uint32_t bars = rw_smem(bar_ptr), barR = bars + kNB * 8u;   // slot b: bars + 8 b
  const uint32_t ring_s = rw_smem(ring), rec_s = rw_smem(rec);
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < L::kBars; ++b) rw_mbar_init(bar_ptr + b);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const unsigned nvisits = (unsigned)R * (unsigned)vpr;
  const int ngroups = C / GC;
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
      rw_expect(bars + 8u * slot, (uint32_t)(kBoxFloats * 4));
      rw_tma_box(ring_s + (uint32_t)(slot * kBoxFloats * 4), &gmap, pair * kRwPairFloats, r * C + cg * GC,
                 bars + 8u * slot);
    }
  };

  unsigned v = fetch();
  if (v < nvisits) {
    const int r = (int)(v / (unsigned)vpr), cg = (int)(v % (unsigned)vpr) * kgroups;
    issue_rec(v);
#pragma unroll
    for (int b = 0; b < kNB; ++b) issue_box(b, r, cg, b);
  }
  while (v < nvisits) {
    const unsigned vn = fetch();
    const int r = (int)(v / (unsigned)vpr), cgbase = (int)(v % (unsigned)vpr) * kgroups;
    const int nk = min(kgroups, ngroups - cgbase);
    rw_wait(barR, pr);
    pr ^= 1u;
    const RwHdr* hp = reinterpret_cast<const RwHdr*>(rec);
    const int kind = hp->kind, batch = hp->batch, x0 = hp->x0, y0 = hp->y0, gw = hp->gw, ny = hp->ny;
    const bool xtail = hp->xtail != 0, xsimple = hp->xsimple != 0;
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
      // pair `p` of this unit has been consumed by every lane: refill its slot with the pair kNB ahead
      auto release = [&](int p) {
        __syncwarp();
        const int np = p + kNB;
        if (np < kRwPairs) issue_box(cslot, r, cg, np);
        else if (has_next) issue_box(cslot, rn, cgn, np - kRwPairs);
        if (kNB < 3 && lane == 0) {  // shallow ring: the pair after that starts its trip from DRAM to L2 now
          const int pp = np + 1;
          if (pp < kRwPairs) rw_tma_prefetch(&gmap, pp * kRwPairFloats, r * C + cg * GC);
          else if (has_next) rw_tma_prefetch(&gmap, (pp - kRwPairs) * kRwPairFloats, rn * C + cgn * GC);
        }
        if (++cslot == kNB) {
          cslot = 0;
          cpar ^= 1u;
        }
      };
      float* img = gt + (size_t)batch * H * WC + (size_t)cg * GC + lane;
      if (kind == kRwFast && (FIXED || xsimple)) {
        float R0[CPL][kRwP], R1[CPL][kRwP], R2[CPL][kRwP], g[CPL][kRwP];
#pragma unroll
        for (int q = 0; q < CPL; ++q)
#pragma unroll
          for (int j = 0; j < kRwP; ++j) R0[q][j] = R1[q][j] = R2[q][j] = g[q][j] = 0.f;
        float* rowp = img + ((size_t)y0 * W + x0) * C;
        int brow = -1;
#pragma unroll 1
        for (int i = 0; i < ny; ++i) {
          const unsigned code = yc[i];
          float2 w = yw[i];
          if (code & (kYNewRow | kYEmit | kYSkip)) {
            if (code & kYSkip) {  // fixed sampling grid with a gap of more than two rows: the window is empty
              rowp += (size_t)__float_as_int(w.x) * WC;
              w = make_float2(0.f, 0.f);   // (the update below then adds nothing: no second exit from the loop body)
            }
            if (code & kYNewRow) {
              ++brow;
              if (!(brow & 1)) {
                if (brow) release((brow >> 1) - 1);
                rw_wait(bars + 8u * cslot, cpar);
              }
              const float2* gp = reinterpret_cast<const float2*>(ring + cslot * kBoxFloats + lane * kRwPairFloats +
                                                                 (brow & 1) * kRwP);
#pragma unroll
              for (int q = 0; q < CPL; ++q)
#pragma unroll
                for (int j = 0; j < kRwP / 2; ++j) {
                  const float2 t = gp[q * (kRwBoxFloats / 2) + j];   // channel lane + 32 q: 32 box rows further
                  g[q][2 * j] = t.x;
                  g[q][2 * j + 1] = t.y;
                }
            }
            if (code & kYEmit) {
              const bool two = (code & kYTwo) != 0;
              // (CPL = 2: no unrolled gw = 2 walk -- the hot code has to stay inside the 32 KB L1.5 I-cache;
              //  ncu with it: 32 % of the stall samples were instruction fetch)
              if (FIXED) {
                if (!xsimple) rw_emit<3, CPL>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
                else rw_emit<0, CPL>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
              } else {
                if (gw == 1) rw_emit<1, CPL>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
                else if (CPL == 1 && gw == 2) rw_emit<2, CPL>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
                else rw_emit<0, CPL>(R0, R1, rowp, two, WC, C, gw, xw, xf, xmask, xtail);
              }
#pragma unroll
              for (int q = 0; q < CPL; ++q)
#pragma unroll
                for (int j = 0; j < kRwP; ++j) {
                  R0[q][j] = R2[q][j];
                  R1[q][j] = 0.f;
                  R2[q][j] = 0.f;
                }
              rowp += 2 * (size_t)WC;
            }
          }
          // samples outside the map and the sentinels carry zero weights
          // packed fp32x2 FMAs (FFMA2): two bins per instruction
          const float2 wl2 = make_float2(w.x, w.x), wh2 = make_float2(w.y, w.y);
#define CDDMSL_RW_UPD(RA, RB)                                                              \
  _Pragma("unroll") for (int q = 0; q < CPL; ++q) _Pragma("unroll") for (int j = 0; j < kRwP / 2; ++j) { \
    const float2 gg = make_float2(g[q][2 * j], g[q][2 * j + 1]);                           \
    const float2 ra = __ffma2_rn(wl2, gg, make_float2(RA[q][2 * j], RA[q][2 * j + 1]));    \
    const float2 rb = __ffma2_rn(wh2, gg, make_float2(RB[q][2 * j], RB[q][2 * j + 1]));    \
    RA[q][2 * j] = ra.x;                                                                   \
    RA[q][2 * j + 1] = ra.y;                                                               \
    RB[q][2 * j] = rb.x;                                                                   \
    RB[q][2 * j + 1] = rb.y;                                                               \
  }
          if (code & kYPos1) {
            CDDMSL_RW_UPD(R1, R2)
          } else {
            CDDMSL_RW_UPD(R0, R1)
          }
#undef CDDMSL_RW_UPD
        }
        release(kRwPairs - 1);
      } else if (kind != kRwSkip) {
        // reference-shaped scatter (huge sampling grids), lanes on channels
        const RoiGeom gm = roi_geom(rois + (size_t)r * 5, scale, aligned, kRwP, kRwP, sampling_ratio, H, W);
        for (int pair = 0; pair < kRwPairs; ++pair) {
          rw_wait(bars + 8u * cslot, cpar);
          rw_slow_pair<CPL>(ring + cslot * kBoxFloats, img, gm, pair, H, W, C);
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
int g_roi_rw_k = 0;    // channel groups per visit of a RoI ("roi_rw_k"; 0 = automatic: 2, or 1 on maps of >= 4096 cells whose
                       // RoIs are large).  Fine-grained visits balance the persistent warps:
                       // VOC shape 1 -> 1.98, 2 -> 1.97, 4 -> 2.09, 8 -> 2.68 ms; Cityscapes shape 1 -> 3.73, 2 -> 3.96,
                       // 4 -> 5.00 ms.  (An L2 evict_last hint on the REDs changes nothing, evict_first on the tile loads
                       // costs 0.1 ms.)
int g_roi_rw_min_units = 65536;  // below this many (RoI, 32-channel) units the channels-last kernel is faster: the
                                 // persistent grid does not fill and the plan kernel's latency shows
                                 // ("roi_rw_min_units"; 1024 RoIs x 1024 channels: 0.45 ms there, 0.77 ms here)
int g_roi_rw_cpl = 1;  // channels per lane ("roi_rw_cpl"; 2 needs C % 64 == 0).  Measured on the VOC shape: CPL = 2 executes
                       // 23 % fewer instructions (0.90 G vs 1.17 G) but its 7 KB boxes leave room for fewer warps: 12 warps
                       // with a 2-deep ring 2.37 ms (12 % of the stall samples wait for the next box), 9 warps with a
                       // 3-deep ring 1.97 ms at one channel group per visit -- the same as CPL = 1 (1.96 ms), so no default

template <int CPL, bool FIXED>
int rw_launch(const float* gout, const float* rois, const unsigned char* recs, float* gt, unsigned* counter, int N,
              int C, int H, int W, int R, float scale, int sampling_ratio, int aligned, cudaStream_t stream) {
  using L = RwLayout<CPL>;
  CUtensorMap gmap;  // gout as [R*C rows][196]: box = 32*CPL channels x one bin-row pair (112 bytes)
  if (tma_encode_2d_f32(&gmap, gout, kRwP * kRwP, (unsigned long long)R * C, kRwP * kRwP * 4, kRwPairFloats, 32 * CPL,
                        CU_TENSOR_MAP_SWIZZLE_NONE))
    return kRwNoTensorMap;
  const int ngroups = C / (32 * CPL);
  const int kgroups = min(g_roi_rw_k > 0 ? g_roi_rw_k : (H * W >= 4096 ? 1 : 2), ngroups);
  const int vpr = ceil_div(ngroups, kgroups);
  auto k = roi_align_bwd_rw_kernel<CPL, FIXED>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long nvisits = (long long)R * vpr;
  const long long want = (nvisits + L::kWarps - 1) / L::kWarps;
  const int grid = (int)(want < (long long)sm_count() ? want : (long long)sm_count());
  k<<<grid, L::kWarps * 32, L::kSmemBytes, stream>>>(gmap, rois, recs, gt, counter, N, C, H, W, R, scale,
                                                     sampling_ratio, aligned, vpr, kgroups);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace

int tune_roi_rw(const char* key, int value) {
  if (!strcmp(key, "roi_rw")) g_roi_rw = value;
  else if (!strcmp(key, "roi_rw_k")) g_roi_rw_k = value > 0 ? value : 0;
  else if (!strcmp(key, "roi_rw_min_units")) g_roi_rw_min_units = value;
  else if (!strcmp(key, "roi_rw_cpl")) g_roi_rw_cpl = value == 1 ? 1 : 2;
  else return 0;
  return 1;
}

bool roi_rw_eligible(int N, int C, int H, int W, int R, int P) {
  (void)N;
  return g_roi_rw && P == kRwP && C > 0 && C % 32 == 0 && (long long)H * W * C < 0x7fffffffLL &&
         (long long)R * C < 0x7fffffffLL /* TMA row coordinate r * C + c */ &&
         (long long)R * (C / 32) >= g_roi_rw_min_units;
}

size_t roi_rw_workspace_bytes(int N, int C, int H, int W, int R) {
  return align_up((size_t)N * C * H * W * 4, 256) + align_up((size_t)R * kRwRecBytes, 256) + 256;
}

int roi_align_bwd_rw(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, void* ws, cudaStream_t stream) {
  (void)P;
  float* gt = reinterpret_cast<float*>(ws);
  const size_t map_bytes = align_up((size_t)N * C * H * W * 4, 256);
  unsigned* counter = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(ws) + map_bytes);
  unsigned char* recs = reinterpret_cast<unsigned char*>(counter) + 256;
  {  // probe the tensor-map encoder before anything is launched
    CUtensorMap probe;
    if (tma_encode_2d_f32(&probe, gout, kRwP * kRwP, (unsigned long long)R * C, kRwP * kRwP * 4, kRwPairFloats, 32,
                          CU_TENSOR_MAP_SWIZZLE_NONE))
      return kRwNoTensorMap;
  }
  cudaError_t e = cudaMemsetAsync(gt, 0, map_bytes + 256, stream);  // gradient map and the work counter
  if (e != cudaSuccess) return (int)e;
  rw_plan_kernel<<<ceil_div(R, 128), 128, 0, stream>>>(rois, recs, N, H, W, R, scale, sampling_ratio, aligned);
  count_launch();
  CDDMSL_CHECK_LAUNCH();
#define CDDMSL_RW_LAUNCH(CPLV, FIXEDV) \
  rw_launch<CPLV, FIXEDV>(gout, rois, recs, gt, counter, N, C, H, W, R, scale, sampling_ratio, aligned, stream)
  const bool two_per_lane = C % 64 == 0 && g_roi_rw_cpl == 2, fixed = sampling_ratio > 0;
  const int rc = two_per_lane ? (fixed ? CDDMSL_RW_LAUNCH(2, true) : CDDMSL_RW_LAUNCH(2, false))
                              : (fixed ? CDDMSL_RW_LAUNCH(1, true) : CDDMSL_RW_LAUNCH(1, false));
#undef CDDMSL_RW_LAUNCH
  if (rc) return rc;
  return launch_transpose(gt, gin, N, H * W, C, stream);  // NHWC -> NCHW (overwrites gin completely)
}

}  // namespace cddmsl
