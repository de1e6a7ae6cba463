// ROIAlign forward / backward, channels-last formulation (the fast path for the 14x14 pooler of the
// reference's configs: config/defaults.py:423-426).
//
// Why: the first version of these kernels (roi_align.cu, kept as the generic path) turned out to be
// instruction-issue bound (ncu: 73-81 % issue-active, 4-11 warp instructions per element) because with an NCHW
// map the lanes of a warp must be spread over bins, so every lane needs its own taps, its own addresses and a
// staged copy of the footprint.  With the 32 lanes of a warp on 32 CHANNELS instead:
//   * a bilinear tap is one fully coalesced 128-byte load from a channels-last copy of the map — no staging,
//   * taps, weights and all control flow are warp-uniform (read from a tiny per-RoI table in shared memory),
//   * a warp owns two output rows (ph, ph+1) of 32 channels and walks the footprint columns once with a
//     register sliding window (separable form out = Ay F Ax^T: vertical blend per column, reused by every
//     sample that touches the column),
//   * the 2x14 results per lane are 28 contiguous floats of the [R,C,14,14] output: seven conflict-free
//     STS.128 into the CTA's output tile, which leaves as ONE TMA bulk store (full 128-byte lines).
// The NCHW <-> NHWC copies of the feature / gradient map are 2 x 157 MB against 6.6 GB of pooled tensor.
// The backward is the exact transpose: the grad tile comes in through shared memory, the window accumulates
// T = G Ax per column and flushes dF[y][x][32ch] += Ay^T T with coalesced 128-byte red.global.add.f32.
#include <cuda_runtime.h>

#include "common.cuh"
#include "roi_common.cuh"
#include "tma_host.cuh"

namespace cddmsl {

constexpr int kMaxG = 10;  // samples per bin and axis that fit the tap tables (RoIs up to 140 cells); sized so that
                            // 4 CTAs of the 2-channel-per-lane kernels share one SM's 227 KB of shared memory

struct __align__(16) TapE {
  int lo;    // absolute column / row index of the low tap, -1: sample contributes nothing
  int hi;
  float wl;  // weights (x table: already divided by the sample count)
  float wh;
};

// ------------------------------------------------------------------------------------------------
// [N][A][B] -> [N][B][A] tiled transpose (NCHW <-> NHWC with A=C,B=H*W or the reverse)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int A,
                                                        int B) {
  __shared__ float tile[32][33];
  const size_t img = (size_t)blockIdx.z * A * B;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int a = a0 + ty + k, b = b0 + tx;
    if (a < A && b < B) tile[ty + k][tx] = __ldg(in + img + (size_t)a * B + b);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int b = b0 + ty + k, a = a0 + tx;
    if (a < A && b < B) out[img + (size_t)b * A + a] = tile[tx][ty + k];
  }
}

int launch_transpose(const float* in, float* out, int N, int A, int B, cudaStream_t stream) {
  if (N == 0 || A == 0 || B == 0) return 0;
  dim3 grid(ceil_div(B, 32), ceil_div(A, 32), N);
  transpose_kernel<<<grid, 256, 0, stream>>>(in, out, A, B);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// shared prologue: per-RoI tap tables
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ TapE make_tap_entry(float start, float bin, int p, int i, int g, int L, float wscale) {
  const Tap tp = make_tap(start, bin, p, i, g, L, 0);
  TapE e;
  e.lo = (tp.wl == 0.f && tp.wh == 0.f) ? -1 : tp.lo;
  e.hi = tp.hi;
  e.wl = tp.wl * wscale;
  e.wh = tp.wh * wscale;
  return e;
}

__device__ __forceinline__ TapE null_tap() {
  TapE e;
  e.lo = -1;
  e.hi = 0;
  e.wl = e.wh = 0.f;
  return e;
}

// Tables cover PE = P rounded up to even bins (rows / columns are processed in pairs); a bin >= P contributes nothing.
template <int P, int NT>
__device__ __forceinline__ void build_tables(TapE* xtab, TapE* ytab, const RoiGeom& g, int H, int W) {
  constexpr int PE = (P + 1) & ~1;
  for (int t = threadIdx.x; t < PE * g.gw; t += NT) {
    const int p = t / g.gw, i = t - p * g.gw;
    xtab[t] = p < P ? make_tap_entry(g.sw, g.bw, p, i, g.gw, W, g.inv_count) : null_tap();
  }
  if (ytab) {
    for (int t = threadIdx.x; t < PE * g.gh; t += NT) {
      const int p = t / g.gh, i = t - p * g.gh;
      ytab[t] = p < P ? make_tap_entry(g.sh, g.bh, p, i, g.gh, H, 1.f) : null_tap();
    }
  }
}

// Monotone check of the x-table + rewrite of its invalid samples, used by the fast sample loops.
// With adaptive sampling (sampling_ratio == 0) the x-samples of a row are at most one cell apart, so the low column of
// consecutive samples advances by 0 or 1 and the sliding window never restarts.  This verifies that on the ACTUAL
// table (rounding could break it) and, if it holds, rewrites the invalid samples at both ends (including the padded
// bins) as zero-weight copies of their nearest valid neighbour, so a sample step needs ONE test (did the column
// change?) instead of nested validity / slide / restart tests.  Call with the table written but not yet synchronised;
// on return the (possibly rewritten) table still needs a barrier before it is read.  scratch: 3 ints of shared memory.
template <int NT>
__device__ __forceinline__ bool monotone_x_table(TapE* xtab, int ns, int* scratch) {
  if (threadIdx.x == 0) {
    scratch[0] = 0x7fffffff;  // first valid sample
    scratch[1] = -1;          // last valid sample
    scratch[2] = 1;           // monotone so far
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ns; t += NT)
    if (xtab[t].lo >= 0) {
      atomicMin(&scratch[0], t);
      atomicMax(&scratch[1], t);
    }
  __syncthreads();
  const int sf = scratch[0], sl = scratch[1];
  for (int t = threadIdx.x; t < ns; t += NT)
    if (t >= sf && t < sl) {  // valid samples are contiguous in x; anything else clears the flag
      const int a = xtab[t].lo, b = xtab[t + 1].lo;
      if (a < 0 || b < a || (b != a && b != xtab[t].hi)) atomicAnd(&scratch[2], 0);
    }
  __syncthreads();
  const bool mono = sl >= 0 && scratch[2] != 0;
  if (mono)
    for (int t = threadIdx.x; t < ns; t += NT)
      if (t < sf || t > sl) {
        TapE e = xtab[t < sf ? sf : sl];
        e.wl = e.wh = 0.f;
        xtab[t] = e;
      }
  return mono;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// CPL channels per lane: lane l owns channels l, l+32, ... of the CTA's 32*CPL-channel group (every tap is CPL
// coalesced 128-byte loads at immediate offsets +128 B, so address arithmetic and control flow are shared).
template <int CPL>
struct Vec {
  float v[CPL];
};

template <int CPL>
__device__ __forceinline__ Vec<CPL> vzero() {
  Vec<CPL> r;
#pragma unroll
  for (int k = 0; k < CPL; ++k) r.v[k] = 0.f;
  return r;
}

template <int CPL>
__device__ __forceinline__ Vec<CPL> vload(const float* __restrict__ p) {
  Vec<CPL> r;
#pragma unroll
  for (int k = 0; k < CPL; ++k) r.v[k] = __ldg(p + 32 * k);
  return r;
}

// row pointer (64-bit, per lane) + column byte offset (32-bit, warp-uniform): two integer instructions per
// address instead of the four (IADD3/IMAD.X/LEA/LEA.HI.X) the compiler emits for `ptr + idx`.
template <int CPL>
__device__ __forceinline__ Vec<CPL> vload_off(const float* __restrict__ rowp, unsigned xb) {
  Vec<CPL> r;
  unsigned long long a;
  asm("{\n\t.reg .u64 t;\n\tcvt.u64.u32 t, %1;\n\tadd.u64 %0, %2, t;\n\t}" : "=l"(a) : "r"(xb), "l"(rowp));
  asm("ld.global.nc.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(a));
  if (CPL > 1) asm("ld.global.nc.f32 %0, [%1+128];" : "=f"(r.v[CPL > 1 ? 1 : 0]) : "l"(a));
  return r;
}

template <int CPL>
__device__ __forceinline__ void vfma(Vec<CPL>& acc, float w, const Vec<CPL>& x) {
#pragma unroll
  for (int k = 0; k < CPL; ++k) acc.v[k] = fmaf(w, x.v[k], acc.v[k]);
}

// Vertical blend of one footprint column for the warp's two rows: v = sum_i wl_i F[lo_i][x] + wh_i F[hi_i][x].
// Addressing is "per-lane 64-bit row pointer + 32-bit element offset of the column".
template <int GH>
struct RowTaps {  // GH > 0: row pointers + weights in registers;  GH == 0: read from the table every time
  const float* plo[GH > 0 ? GH : 1];
  const float* phi[GH > 0 ? GH : 1];
  float wl[GH > 0 ? GH : 1], wh[GH > 0 ? GH : 1];
};

template <int GH>
__device__ __forceinline__ void load_row_taps(RowTaps<GH>& rt, const float* __restrict__ base, const TapE* ytab_row,
                                              int WC) {
#pragma unroll
  for (int i = 0; i < (GH > 0 ? GH : 0); ++i) {
    const TapE e = ytab_row[i];
    const bool ok = e.lo >= 0;
    rt.plo[i] = base + (ok ? e.lo * WC : 0);
    rt.phi[i] = base + (ok ? e.hi * WC : 0);
    rt.wl[i] = ok ? e.wl : 0.f;
    rt.wh[i] = ok ? e.wh : 0.f;
  }
}

template <int GH, int CPL>
__device__ __forceinline__ void fwd_column(const float* __restrict__ base, unsigned xo, const RowTaps<GH>& ra,
                                           const RowTaps<GH>& rb, const TapE* ya, const TapE* yb, int gh, int WC,
                                           Vec<CPL>& va, Vec<CPL>& vb) {
  va = vzero<CPL>();
  vb = vzero<CPL>();
  if (GH > 0) {
    Vec<CPL> l[4 * (GH > 0 ? GH : 1)];
#pragma unroll
    for (int i = 0; i < (GH > 0 ? GH : 0); ++i) {  // issue every load of the column before the first use
      l[4 * i + 0] = vload_off<CPL>(ra.plo[i], xo);
      l[4 * i + 1] = vload_off<CPL>(ra.phi[i], xo);
      l[4 * i + 2] = vload_off<CPL>(rb.plo[i], xo);
      l[4 * i + 3] = vload_off<CPL>(rb.phi[i], xo);
    }
#pragma unroll
    for (int i = 0; i < (GH > 0 ? GH : 0); ++i) {
      vfma<CPL>(va, ra.wl[i], l[4 * i + 0]);
      vfma<CPL>(va, ra.wh[i], l[4 * i + 1]);
      vfma<CPL>(vb, rb.wl[i], l[4 * i + 2]);
      vfma<CPL>(vb, rb.wh[i], l[4 * i + 3]);
    }
  } else {
    for (int i = 0; i < gh; ++i) {
      const TapE ea = ya[i], eb = yb[i];
      if (ea.lo >= 0) {
        vfma<CPL>(va, ea.wl, vload_off<CPL>(base, xo + (unsigned)(ea.lo * WC) * 4u));
        vfma<CPL>(va, ea.wh, vload_off<CPL>(base, xo + (unsigned)(ea.hi * WC) * 4u));
      }
      if (eb.lo >= 0) {
        vfma<CPL>(vb, eb.wl, vload_off<CPL>(base, xo + (unsigned)(eb.lo * WC) * 4u));
        vfma<CPL>(vb, eb.wh, vload_off<CPL>(base, xo + (unsigned)(eb.hi * WC) * 4u));
      }
    }
  }
}

// One sample of the sliding window (warp-uniform control flow).  The column loader appears at ONE site (the
// t-loop runs once when the window just slides, twice when it (re)starts): the first version of this kernel,
// fully unrolled over the 14 bins, was 21 K SASS instructions and lost 27 % of its issue slots to
// instruction-cache misses (ncu stall_no_inst).
#define CDDMSL_FWD_SAMPLE(e, SA, SB)                                                        \
  if ((e).lo >= 0) {                                                                        \
    if ((e).lo != cur) {                                                                    \
      if ((e).lo == cur + 1) {                                                              \
        va0 = va1;                                                                          \
        vb0 = vb1;                                                                          \
      } else {                                                                              \
        fwd_column<GH, CPL>(base, (unsigned)(e).lo * Cb, ra, rb, ya, yb, gh, WC, va0, vb0); \
      }                                                                                     \
      cur = (e).lo;                                                                         \
      /* hi == min(lo+1, W-1): always a valid column (weight 0 when clamped) */             \
      fwd_column<GH, CPL>(base, (unsigned)(e).hi * Cb, ra, rb, ya, yb, gh, WC, va1, vb1);   \
    }                                                                                       \
    vfma<CPL>(SA, (e).wl, va0);                                                             \
    vfma<CPL>(SA, (e).wh, va1);                                                             \
    vfma<CPL>(SB, (e).wl, vb0);                                                             \
    vfma<CPL>(SB, (e).wh, vb1);                                                             \
  }

template <int P, int GH, int CPL>
__device__ __forceinline__ void fwd_rows(const float* __restrict__ base /* image + first channel of this lane */,
                                         const TapE* __restrict__ xtab, const TapE* __restrict__ ytab, int gw, int gh,
                                         int W, int C, int row_a, float* __restrict__ orow /* tile row of channel lane */,
                                         int nch /* channels of this lane that exist */,
                                         int kstride = 32 * P * P /* floats between channels l and l+32 in the tile */) {
  const int WC = W * C;
  const unsigned Cb = (unsigned)C * 4u;  // bytes between two columns of the channels-last map
  RowTaps<GH> ra, rb;
  const TapE* ya = ytab + row_a * gh;
  const TapE* yb = ya + gh;
  load_row_taps<GH>(ra, base, ya, WC);
  load_row_taps<GH>(rb, base, yb, WC);
  int cur = -4;
  Vec<CPL> va0 = vzero<CPL>(), va1 = vzero<CPL>(), vb0 = vzero<CPL>(), vb1 = vzero<CPL>();
  const TapE* xt = xtab;
#pragma unroll 1
  for (int pw = 0; pw < P; pw += 2) {
    Vec<CPL> sa0 = vzero<CPL>(), sb0 = vzero<CPL>(), sa1 = vzero<CPL>(), sb1 = vzero<CPL>();
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_FWD_SAMPLE(e, sa0, sb0)
    }
    xt += gw;
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_FWD_SAMPLE(e, sa1, sb1)
    }
    xt += gw;
    // rows a and b of a channel are 2*P contiguous floats of the output tile; channel k of this lane is 32*k
    // tile rows further on
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      if (k < nch) {
        float* o = orow + k * kstride;
        if (P % 2 == 0) {
          *reinterpret_cast<float2*>(o + pw) = make_float2(sa0.v[k], sa1.v[k]);
          *reinterpret_cast<float2*>(o + P + pw) = make_float2(sb0.v[k], sb1.v[k]);
        } else {  // odd pooled size (7x7): rows are not 8-byte aligned and the padded bin / row does not exist
          o[pw] = sa0.v[k];
          if (pw + 1 < P) o[pw + 1] = sa1.v[k];
          if (row_a + 1 < P) {
            o[P + pw] = sb0.v[k];
            if (pw + 1 < P) o[P + pw + 1] = sb1.v[k];
          }
        }
      }
    }
  }
}

// Monotone variant of the sample step (table prepared by monotone_x_table): the window only ever slides.
#define CDDMSL_FWD_SAMPLE_MONO(e, SA, SB)                                                     \
  if ((e).lo != cur) {                                                                        \
    va0 = va1;                                                                                \
    vb0 = vb1;                                                                                \
    cur = (e).lo;                                                                             \
    fwd_column<GH, CPL>(base, (unsigned)(e).hi * Cb, ra, rb, ya, yb, gh, WC, va1, vb1);       \
  }                                                                                           \
  vfma<CPL>(SA, (e).wl, va0);                                                                 \
  vfma<CPL>(SA, (e).wh, va1);                                                                 \
  vfma<CPL>(SB, (e).wl, vb0);                                                                 \
  vfma<CPL>(SB, (e).wh, vb1);

template <int P, int GH, int CPL, int GW>
__device__ __forceinline__ void fwd_rows_mono(const float* __restrict__ base, const TapE* __restrict__ xtab,
                                              const TapE* __restrict__ ytab, int gw_, int gh, int W, int C, int row_a,
                                              float* __restrict__ orow, int kstride) {
  static_assert(P % 2 == 0, "monotone path: even pooled sizes only");
  const int gw = GW > 0 ? GW : gw_;  // 1 or 2 samples per bin known at compile time: the sample loops unroll away
  const int WC = W * C;
  const unsigned Cb = (unsigned)C * 4u;
  RowTaps<GH> ra, rb;
  const TapE* ya = ytab + row_a * gh;
  const TapE* yb = ya + gh;
  load_row_taps<GH>(ra, base, ya, WC);
  load_row_taps<GH>(rb, base, yb, WC);
  // the window starts one column to the left of the first sample with that sample's low column already in slot 1:
  // the first step slides it into slot 0 and fetches the high column
  const int first = xtab[0].lo;
  int cur = first - 1;
  Vec<CPL> va0 = vzero<CPL>(), vb0 = vzero<CPL>(), va1, vb1;
  fwd_column<GH, CPL>(base, (unsigned)first * Cb, ra, rb, ya, yb, gh, WC, va1, vb1);
  const TapE* xt = xtab;
#pragma unroll 1
  for (int pw = 0; pw < P; pw += 2) {
    Vec<CPL> sa0 = vzero<CPL>(), sb0 = vzero<CPL>(), sa1 = vzero<CPL>(), sb1 = vzero<CPL>();
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_FWD_SAMPLE_MONO(e, sa0, sb0)
    }
    xt += gw;
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_FWD_SAMPLE_MONO(e, sa1, sb1)
    }
    xt += gw;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      float* o = orow + k * kstride;
      *reinterpret_cast<float2*>(o + pw) = make_float2(sa0.v[k], sa1.v[k]);
      *reinterpret_cast<float2*>(o + P + pw) = make_float2(sb0.v[k], sb1.v[k]);
    }
  }
}

// gh == 1 monotone path with the NEXT footprint column's raw loads kept in flight: the kernel is bound by the latency
// of first-touch lines (long-scoreboard 6.5 cycles per issued instruction), and with one row sample per bin a column is
// only four loads per channel -- room for eight more registers.
template <int P, int CPL, int GW>
__device__ __forceinline__ void fwd_rows_mono_pf(const float* __restrict__ base, const TapE* __restrict__ xtab,
                                                 const TapE* __restrict__ ytab, int gw_, int W, int C, int row_a,
                                                 float* __restrict__ orow, int kstride) {
  static_assert(P % 2 == 0, "monotone path: even pooled sizes only");
  const int gw = GW > 0 ? GW : gw_;
  const int WC = W * C;
  const unsigned Cb = (unsigned)C * 4u;
  RowTaps<1> ra, rb;
  load_row_taps<1>(ra, base, ytab + row_a, WC);
  load_row_taps<1>(rb, base, ytab + row_a + 1, WC);
  Vec<CPL> p0, p1, p2, p3;  // raw rows (a.lo, a.hi, b.lo, b.hi) of the prefetched column
#define CDDMSL_PF_ISSUE(col)                      \
  {                                               \
    const unsigned xo_ = (unsigned)(col) * Cb;    \
    p0 = vload_off<CPL>(ra.plo[0], xo_);          \
    p1 = vload_off<CPL>(ra.phi[0], xo_);          \
    p2 = vload_off<CPL>(rb.plo[0], xo_);          \
    p3 = vload_off<CPL>(rb.phi[0], xo_);          \
  }
#define CDDMSL_PF_BLEND(va, vb)       \
  {                                   \
    va = vzero<CPL>();                \
    vb = vzero<CPL>();                \
    vfma<CPL>(va, ra.wl[0], p0);      \
    vfma<CPL>(va, ra.wh[0], p1);      \
    vfma<CPL>(vb, rb.wl[0], p2);      \
    vfma<CPL>(vb, rb.wh[0], p3);      \
  }
  const int first = xtab[0].lo;
  int cur = first - 1;
  Vec<CPL> va0 = vzero<CPL>(), vb0 = vzero<CPL>(), va1, vb1;
  CDDMSL_PF_ISSUE(first)
  CDDMSL_PF_BLEND(va1, vb1)
  CDDMSL_PF_ISSUE(min(first + 1, W - 1))
  const TapE* xt = xtab;
#pragma unroll 1
  for (int pw = 0; pw < P; pw += 2) {
    Vec<CPL> s[2][2] = {{vzero<CPL>(), vzero<CPL>()}, {vzero<CPL>(), vzero<CPL>()}};  // [bin][row]
#pragma unroll
    for (int bin = 0; bin < 2; ++bin) {
      for (int ix = 0; ix < gw; ++ix) {
        const TapE e = xt[ix];
        if (e.lo != cur) {  // slide: the prefetched column is exactly e.hi
          va0 = va1;
          vb0 = vb1;
          cur = e.lo;
          CDDMSL_PF_BLEND(va1, vb1)
          CDDMSL_PF_ISSUE(min(e.hi + 1, W - 1))
        }
        vfma<CPL>(s[bin][0], e.wl, va0);
        vfma<CPL>(s[bin][0], e.wh, va1);
        vfma<CPL>(s[bin][1], e.wl, vb0);
        vfma<CPL>(s[bin][1], e.wh, vb1);
      }
      xt += gw;
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      float* o = orow + k * kstride;
      *reinterpret_cast<float2*>(o + pw) = make_float2(s[0][0].v[k], s[1][0].v[k]);
      *reinterpret_cast<float2*>(o + P + pw) = make_float2(s[0][1].v[k], s[1][1].v[k]);
    }
  }
#undef CDDMSL_PF_ISSUE
#undef CDDMSL_PF_BLEND
}

__device__ void fwd_direct_any(const float* __restrict__ in, float* __restrict__ out_roi, int c0, int nch, int C, int H,
                               int W, int PH, int PW, const RoiGeom& g, int nthreads) {
  const int per = PH * PW;
  for (int e = threadIdx.x; e < nch * per; e += nthreads) {
    const int c = e / per, b = e - c * per;
    const int ph = b / PW, pw = b - ph * PW;
    const float* plane = in + ((size_t)g.batch * C + c0 + c) * H * W;
    float acc = 0.f;
    for (int iy = 0; iy < g.gh; ++iy) {
      const Tap ty = make_tap(g.sh, g.bh, ph, iy, g.gh, H, 0);
      for (int ix = 0; ix < g.gw; ++ix) {
        const Tap tx = make_tap(g.sw, g.bw, pw, ix, g.gw, W, 0);
        acc += ty.wl * tx.wl * plane[ty.lo * W + tx.lo] + ty.wl * tx.wh * plane[ty.lo * W + tx.hi] +
               ty.wh * tx.wl * plane[ty.hi * W + tx.lo] + ty.wh * tx.wh * plane[ty.hi * W + tx.hi];
      }
    }
    out_roi[(size_t)(c0 + c) * per + b] = acc * g.inv_count;
  }
}

template <int P, int CPL>
__global__ void __launch_bounds__(((P + 1) / 2) * 32, 4)
roi_align_fwd_cl_kernel(const float* __restrict__ ft, const float* __restrict__ in_nchw,
                        const float* __restrict__ rois, float* __restrict__ out, int N, int C, int H, int W, int R,
                        float scale, int sampling_ratio, int aligned, int ngroups, int gpc) {
  constexpr int NW = (P + 1) / 2, NT = NW * 32, PER = P * P, GC = 32 * CPL, PE = (P + 1) & ~1;
  extern __shared__ __align__(128) float dyn_smem[];
  float* O_s = dyn_smem;                                        // [GC][PER]
  TapE* xtab = reinterpret_cast<TapE*>(O_s + GC * PER);         // [PE * kMaxG]
  TapE* ytab = xtab + PE * kMaxG;
  // One CTA = one RoI x `gpc` consecutive 64-channel groups: geometry and tap tables are built once and reused.
  const int r = blockIdx.x / ngroups;
  const int cbeg = (blockIdx.x - r * ngroups) * (GC * gpc);
  const int cend = min(C, cbeg + GC * gpc);
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, P, P, sampling_ratio, H, W);
  if (g.gw <= 0 || g.gh <= 0 || g.batch < 0 || g.batch >= N) {
    float* o = out + ((size_t)r * C + cbeg) * PER;
    for (int e = threadIdx.x; e < (cend - cbeg) * PER; e += NT) o[e] = 0.f;
    return;
  }
  if (g.gw > kMaxG || g.gh > kMaxG) {  // more samples per bin than the tables hold: reference-order loop
    fwd_direct_any(in_nchw, out + (size_t)r * C * PER, cbeg, cend - cbeg, C, H, W, P, P, g, NT);
    return;
  }
  build_tables<P, NT>(xtab, ytab, g, H, W);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = cbeg; c0 < cend; c0 += GC) {
    const int nc = min(GC, cend - c0);
    float* out_tile = out + ((size_t)r * C + c0) * PER;
    // channels of this lane: c0 + lane + 32k; a ragged last group masks the missing ones
    const int nch = lane < nc ? min(CPL, (nc - lane + 31) / 32) : 0;
    if (nch > 0) {
      const float* base = ft + (size_t)g.batch * H * W * C + c0 + lane;
      float* orow = O_s + lane * PER + (2 * warp) * P;
      if (nch == CPL) {
        if (g.gh == 1) fwd_rows<P, 1, CPL>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, nch);
        else if (g.gh == 2) fwd_rows<P, 2, CPL>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, nch);
        else fwd_rows<P, 0, CPL>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, nch);
      } else {
        fwd_rows<P, 0, 1>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, 1);  // ragged tail: one channel
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // the bulk copy needs a 16-byte aligned destination and size (always true for 14x14; for 7x7 when the tile
    // starts at a multiple of 4 channels)
    const bool bulk_ok = ((nc * PER) % 4 == 0) && ((reinterpret_cast<uintptr_t>(out_tile) & 15) == 0);
    if (bulk_ok) {
      if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(O_s);
        const uint32_t bytes = (uint32_t)(nc * PER) * 4u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out_tile), "r"(s),
                     "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // O_s may be rewritten (or the CTA retire) once the copy engine has READ the tile
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else {
      for (int e = threadIdx.x; e < nc * PER; e += NT) out_tile[e] = O_s[e];
    }
    if (c0 + GC < cend) __syncthreads();  // next group overwrites the tile
  }
}

// ---- mbarrier + 1-D TMA bulk load (global -> shared), used to bring the contiguous grad tile in -----------------
__device__ __forceinline__ uint32_t roi_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void roi_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(roi_smem_u32(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void roi_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(roi_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   roi_smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(roi_smem_u32(bar))
               : "memory");
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void roi_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = roi_smem_u32(bar);
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

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct __align__(16) YSlots {  // up to 4 distinct rows touched by (sample i of row a, sample i of row b), merged
  int off[4];    // y * W * C, -1: inactive
  float wa[4], wb[4];
};

__device__ __forceinline__ YSlots make_slots(const TapE ea, const TapE eb, int WC) {
  YSlots s;
  const int y0 = ea.lo >= 0 ? ea.lo : -10, y1 = ea.lo >= 0 ? ea.hi : -11;
  const int y2 = eb.lo >= 0 ? eb.lo : -12, y3 = eb.lo >= 0 ? eb.hi : -13;
  const float a0 = ea.lo >= 0 ? ea.wl : 0.f, a1 = ea.lo >= 0 ? ea.wh : 0.f;
  const float b2 = eb.lo >= 0 ? eb.wl : 0.f, b3 = eb.lo >= 0 ? eb.wh : 0.f;
  s.off[0] = y0 >= 0 ? y0 * WC : -1;
  s.wa[0] = a0 + (y1 == y0 ? a1 : 0.f);
  s.wb[0] = (y2 == y0 ? b2 : 0.f) + (y3 == y0 ? b3 : 0.f);
  const bool n1 = y1 >= 0 && y1 != y0;
  s.off[1] = n1 ? y1 * WC : -1;
  s.wa[1] = a1;
  s.wb[1] = (y2 == y1 ? b2 : 0.f) + (y3 == y1 ? b3 : 0.f);
  const bool n2 = y2 >= 0 && y2 != y0 && y2 != y1;
  s.off[2] = n2 ? y2 * WC : -1;
  s.wa[2] = 0.f;
  s.wb[2] = b2 + (y3 == y2 ? b3 : 0.f);
  const bool n3 = y3 >= 0 && y3 != y0 && y3 != y1 && y3 != y2;
  s.off[3] = n3 ? y3 * WC : -1;
  s.wa[3] = 0.f;
  s.wb[3] = b3;
  return s;
}

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// lane base pointer (64-bit) + byte offset (32-bit): CPL coalesced 128-byte reductions 128 B apart
template <int CPL>
__device__ __forceinline__ void red_add_off(float* base, unsigned ob, const float* v, int nch) {
  unsigned long long a;
  asm("{\n\t.reg .u64 t;\n\tcvt.u64.u32 t, %1;\n\tadd.u64 %0, %2, t;\n\t}" : "=l"(a) : "r"(ob), "l"(base));
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(a), "f"(v[0]) : "memory");
  if (CPL > 1 && nch > 1) asm volatile("red.global.add.f32 [%0+128], %1;" ::"l"(a), "f"(v[CPL > 1 ? 1 : 0]) : "memory");
}

// dF[y][x][c] += Ay^T (ta, tb) for one footprint column (xb = column byte offset; slot offsets are in elements)
template <bool GH1, int CPL>
__device__ __forceinline__ void bwd_flush(float* __restrict__ base, unsigned xb, const YSlots& s1,
                                          const YSlots* __restrict__ slots, int gh, const Vec<CPL>& ta,
                                          const Vec<CPL>& tb, int nch) {
  if (GH1) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (s1.off[k] >= 0) {
        float v[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) v[q] = fmaf(s1.wa[k], ta.v[q], s1.wb[k] * tb.v[q]);
        red_add_off<CPL>(base, xb + (unsigned)s1.off[k] * 4u, v, nch);
      }
  } else {
    for (int i = 0; i < gh; ++i) {
      const YSlots s = slots[i];  // 3 x LDS.128, warp-uniform
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (s.off[k] >= 0) {
          float v[CPL];
#pragma unroll
          for (int q = 0; q < CPL; ++q) v[q] = fmaf(s.wa[k], ta.v[q], s.wb[k] * tb.v[q]);
          red_add_off<CPL>(base, xb + (unsigned)s.off[k] * 4u, v, nch);
        }
    }
  }
}

#define CDDMSL_BWD_SAMPLE(e, COMP)                                                            \
  if ((e).lo >= 0) {                                                                          \
    if ((e).lo != cur) {                                                                      \
      if (cur >= 0) {                                                                         \
        bwd_flush<GH1, CPL>(base, (unsigned)cur * Cb, s1, slots, gh, ta0, tb0, nch);          \
        if ((e).lo == cur + 1) {                                                              \
          ta0 = ta1;                                                                          \
          tb0 = tb1;                                                                          \
        } else {                                                                              \
          if (curhi != cur) bwd_flush<GH1, CPL>(base, (unsigned)curhi * Cb, s1, slots, gh, ta1, tb1, nch); \
          ta0 = vzero<CPL>();                                                                 \
          tb0 = vzero<CPL>();                                                                 \
        }                                                                                     \
      }                                                                                       \
      ta1 = vzero<CPL>();                                                                     \
      tb1 = vzero<CPL>();                                                                     \
      cur = (e).lo;                                                                           \
      curhi = (e).hi;                                                                         \
    }                                                                                         \
    _Pragma("unroll") for (int q = 0; q < CPL; ++q) {                                         \
      ta0.v[q] = fmaf((e).wl, ga[q].COMP, ta0.v[q]);                                          \
      ta1.v[q] = fmaf((e).wh, ga[q].COMP, ta1.v[q]);                                          \
      tb0.v[q] = fmaf((e).wl, gb[q].COMP, tb0.v[q]);                                          \
      tb1.v[q] = fmaf((e).wh, gb[q].COMP, tb1.v[q]);                                          \
    }                                                                                         \
  }

template <int P, bool GH1, int CPL>
__device__ __forceinline__ void bwd_rows(float* __restrict__ base, const TapE* __restrict__ xtab,
                                         const YSlots* __restrict__ slots, int gw, int gh, int W, int C,
                                         const float* __restrict__ grow, int nch, bool row_b_exists,
                                         int kstride = 32 * P * P) {
  YSlots s1;
  if (GH1) s1 = slots[0];
  const unsigned Cb = (unsigned)C * 4u;
  int cur = -4, curhi = -4;
  Vec<CPL> ta0 = vzero<CPL>(), ta1 = vzero<CPL>(), tb0 = vzero<CPL>(), tb1 = vzero<CPL>();
  const TapE* xt = xtab;
#pragma unroll 1
  for (int pw = 0; pw < P; pw += 2) {
    float2 ga[CPL], gb[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const float* gq = grow + (q < nch ? q : 0) * kstride;
      if (P % 2 == 0) {
        ga[q] = *reinterpret_cast<const float2*>(gq + pw);
        gb[q] = *reinterpret_cast<const float2*>(gq + P + pw);
      } else {  // odd pooled size: unaligned rows; the padded bin / row carries no gradient (its taps are null)
        ga[q].x = gq[pw];
        ga[q].y = pw + 1 < P ? gq[pw + 1] : 0.f;
        gb[q].x = row_b_exists ? gq[P + pw] : 0.f;
        gb[q].y = (row_b_exists && pw + 1 < P) ? gq[P + pw + 1] : 0.f;
      }
    }
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_BWD_SAMPLE(e, x)
    }
    xt += gw;
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_BWD_SAMPLE(e, y)
    }
    xt += gw;
  }
  if (cur >= 0) {
    bwd_flush<GH1, CPL>(base, (unsigned)cur * Cb, s1, slots, gh, ta0, tb0, nch);
    if (curhi != cur) bwd_flush<GH1, CPL>(base, (unsigned)curhi * Cb, s1, slots, gh, ta1, tb1, nch);
  }
}

// ---- monotone fast path ------------------------------------------------------------------------------------
// With adaptive sampling (sampling_ratio == 0) the x-samples of a row are at most one cell apart, so the low column of
// consecutive samples advances by 0 or 1: the window never restarts.  When the prologue has verified that on the
// actual table (and rewritten the invalid samples at both ends as zero-weight copies of their nearest valid
// neighbour), the per-sample step needs ONE test (did the column change?) instead of four nested ones.
#define CDDMSL_BWD_SAMPLE_MONO(e, COMP)                                                       \
  if ((e).lo != cur) {                                                                        \
    bwd_flush<GH1, CPL>(base, (unsigned)cur * Cb, s1, slots, gh, ta0, tb0, nch);              \
    ta0 = ta1;                                                                                \
    tb0 = tb1;                                                                                \
    ta1 = vzero<CPL>();                                                                       \
    tb1 = vzero<CPL>();                                                                       \
    cur = (e).lo;                                                                             \
    curhi = (e).hi;                                                                           \
  }                                                                                           \
  _Pragma("unroll") for (int q = 0; q < CPL; ++q) {                                           \
    ta0.v[q] = fmaf((e).wl, ga[q].COMP, ta0.v[q]);                                            \
    ta1.v[q] = fmaf((e).wh, ga[q].COMP, ta1.v[q]);                                            \
    tb0.v[q] = fmaf((e).wl, gb[q].COMP, tb0.v[q]);                                            \
    tb1.v[q] = fmaf((e).wh, gb[q].COMP, tb1.v[q]);                                            \
  }

template <int P, bool GH1, int CPL, int GW>
__device__ __forceinline__ void bwd_rows_mono(float* __restrict__ base, const TapE* __restrict__ xtab,
                                              const YSlots* __restrict__ slots, int gw_, int gh, int W, int C,
                                              const float* __restrict__ grow, int nch) {
  static_assert(P % 2 == 0, "monotone path: even pooled sizes only");
  const int gw = GW > 0 ? GW : gw_;
  YSlots s1;
  if (GH1) s1 = slots[0];
  const unsigned Cb = (unsigned)C * 4u;
  int cur = xtab[0].lo, curhi = xtab[0].hi;  // entry 0 is valid or a copy of the first valid sample
  Vec<CPL> ta0 = vzero<CPL>(), ta1 = vzero<CPL>(), tb0 = vzero<CPL>(), tb1 = vzero<CPL>();
  const TapE* xt = xtab;
#pragma unroll 1
  for (int pw = 0; pw < P; pw += 2) {
    float2 ga[CPL], gb[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const float* gq = grow + (q < nch ? q : 0) * 32 * P * P;
      ga[q] = *reinterpret_cast<const float2*>(gq + pw);
      gb[q] = *reinterpret_cast<const float2*>(gq + P + pw);
    }
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_BWD_SAMPLE_MONO(e, x)
    }
    xt += gw;
    for (int ix = 0; ix < gw; ++ix) {
      const TapE e = xt[ix];
      CDDMSL_BWD_SAMPLE_MONO(e, y)
    }
    xt += gw;
  }
  bwd_flush<GH1, CPL>(base, (unsigned)cur * Cb, s1, slots, gh, ta0, tb0, nch);
  if (curhi != cur) bwd_flush<GH1, CPL>(base, (unsigned)curhi * Cb, s1, slots, gh, ta1, tb1, nch);
}

template <int P, int CPL>
__global__ void __launch_bounds__(((P + 1) / 2) * 32, 4)
roi_align_bwd_cl_kernel(const float* __restrict__ gout, const float* __restrict__ rois, float* __restrict__ gt, int N,
                        int C, int H, int W, int R, float scale, int sampling_ratio, int aligned, int ngroups, int gpc) {
  constexpr int NW = (P + 1) / 2, NT = NW * 32, PER = P * P, GC = 32 * CPL, PE = (P + 1) & ~1;
  extern __shared__ __align__(128) float dyn_smem[];
  float* G_s = dyn_smem;                                   // [GC][PER] (+ pad to 16 B)
  TapE* xtab = reinterpret_cast<TapE*>(G_s + ((GC * PER + 3) & ~3));    // [PE * kMaxG]
  YSlots* yslots = reinterpret_cast<YSlots*>(xtab + PE * kMaxG);  // [NW][kMaxG]: [row pair][sample]
  const int r = blockIdx.x / ngroups;
  const int cbeg = (blockIdx.x - r * ngroups) * (GC * gpc);
  const int cend = min(C, cbeg + GC * gpc);
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, P, P, sampling_ratio, H, W);
  if (g.gw <= 0 || g.gh <= 0 || g.batch < 0 || g.batch >= N) return;
  if (g.gw > kMaxG || g.gh > kMaxG) {  // reference-style scatter into the channels-last map
    const float* g_tile = gout + ((size_t)r * C + cbeg) * PER;
    float* img = gt + (size_t)g.batch * H * W * C + cbeg;
    for (int e = threadIdx.x; e < (cend - cbeg) * PER; e += NT) {
      const int c = e / PER, b = e - c * PER;
      const int ph = b / P, pw = b - ph * P;
      const float go = g_tile[e] * g.inv_count;
      for (int iy = 0; iy < g.gh; ++iy) {
        const Tap ty = make_tap(g.sh, g.bh, ph, iy, g.gh, H, 0);
        if (ty.wl == 0.f && ty.wh == 0.f) continue;
        for (int ix = 0; ix < g.gw; ++ix) {
          const Tap tx = make_tap(g.sw, g.bw, pw, ix, g.gw, W, 0);
          if (tx.wl == 0.f && tx.wh == 0.f) continue;
          red_add(img + ((size_t)ty.lo * W + tx.lo) * C + c, go * ty.wl * tx.wl);
          red_add(img + ((size_t)ty.lo * W + tx.hi) * C + c, go * ty.wl * tx.wh);
          red_add(img + ((size_t)ty.hi * W + tx.lo) * C + c, go * ty.wh * tx.wl);
          red_add(img + ((size_t)ty.hi * W + tx.hi) * C + c, go * ty.wh * tx.wh);
        }
      }
    }
    return;
  }
  // The grad tile of a channel group is one contiguous [GC][PER] block: a single TMA bulk load per group, issued
  // BEFORE the tap tables are built so the copy engine works under the prologue (was: LDG -> STS by every thread,
  // 17 % of the kernel's stall samples sat on those STS).
  __shared__ __align__(8) uint64_t tile_bar;
  if (threadIdx.x == 0) roi_mbar_init(&tile_bar);
  __syncthreads();
  uint32_t tile_phase = 0;
  auto tile_is_bulk = [&](int c0) {
    const int nc = min(GC, cend - c0);
    return ((nc * PER) % 4 == 0) && ((reinterpret_cast<uintptr_t>(gout + ((size_t)r * C + c0) * PER) & 15) == 0);
  };
  auto issue_tile = [&](int c0) {
    if (threadIdx.x == 0 && tile_is_bulk(c0))
      roi_bulk_load(G_s, gout + ((size_t)r * C + c0) * PER, (uint32_t)(min(GC, cend - c0) * PER) * 4u, &tile_bar);
  };
  issue_tile(cbeg);
  build_tables<P, NT>(xtab, nullptr, g, H, W);
  for (int t = threadIdx.x; t < NW * g.gh; t += NT) {  // merged row slots of every (row pair, sample)
    const int j = t / g.gh, i = t - j * g.gh;
    yslots[j * kMaxG + i] =
        make_slots(make_tap_entry(g.sh, g.bh, 2 * j, i, g.gh, H, 1.f),
                   2 * j + 1 < P ? make_tap_entry(g.sh, g.bh, 2 * j + 1, i, g.gh, H, 1.f) : null_tap(), W * C);
  }
  // Monotone fast path (see monotone_x_table).  Its barriers cost nothing here: the grad tile is still in flight.
  __shared__ int mono_s[3];
  bool mono = false;
  if (P % 2 == 0 && sampling_ratio <= 0) mono = monotone_x_table<NT>(xtab, PE * g.gw, mono_s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool row_b_exists = 2 * warp + 1 < P;
  for (int c0 = cbeg; c0 < cend; c0 += GC) {
    const int nc = min(GC, cend - c0);
    const float* g_tile = gout + ((size_t)r * C + c0) * PER;
    float* img = gt + (size_t)g.batch * H * W * C + c0;
    if (tile_is_bulk(c0)) {
      roi_mbar_wait(&tile_bar, tile_phase);
      tile_phase ^= 1u;
    } else {  // unaligned / ragged tile (7x7 with an odd channel offset): plain loads
      for (int e = threadIdx.x; e < nc * PER; e += NT) G_s[e] = __ldcs(g_tile + e);
    }
    __syncthreads();  // tile (and, the first time round, the tables) visible
    const int nch = lane < nc ? min(CPL, (nc - lane + 31) / 32) : 0;
    if (nch > 0) {
      float* base = img + lane;
      const float* grow = G_s + lane * PER + (2 * warp) * P;
      bool took_mono_path = false;
      if constexpr (P % 2 == 0) {
        if (mono) {
#define CDDMSL_BWD_MONO(GH1V, GWV) \
  bwd_rows_mono<P, GH1V, CPL, GWV>(base, xtab, yslots + warp * kMaxG, g.gw, g.gh, W, C, grow, nch)
          // (more compile-time sample counts were measured on the backward: slower -- the kernel is issue-bound and
          //  the extra variants cost more instruction-cache misses than their unrolled loops save)
          if (g.gh == 1 && g.gw == 1) CDDMSL_BWD_MONO(true, 1);
          else if (g.gh == 1) CDDMSL_BWD_MONO(true, 0);
          else CDDMSL_BWD_MONO(false, 0);
#undef CDDMSL_BWD_MONO
          took_mono_path = true;
        }
      }
      if (!took_mono_path) {
        if (g.gh == 1)
          bwd_rows<P, true, CPL>(base, xtab, yslots + warp * kMaxG, g.gw, g.gh, W, C, grow, nch, row_b_exists);
        else
          bwd_rows<P, false, CPL>(base, xtab, yslots + warp * kMaxG, g.gw, g.gh, W, C, grow, nch, row_b_exists);
      }
    }
    if (c0 + GC < cend) {
      __syncthreads();  // every warp is done reading the tile: the next group's copy may overwrite it
      issue_tile(c0 + GC);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 14x14, C % 64 == 0: warps decoupled through per-warp TMA tensor copies
// ------------------------------------------------------------------------------------------------
// The kernels above move the [64][196] tile of a channel group with ONE bulk copy, so every group costs CTA-wide
// barriers (ncu: barrier = 14 % of the forward's stall cycles, the mbarrier spin 12 % of the backward's samples).
// Warp j only ever touches rows 2j, 2j+1 of the 64 channels: a [64 channels][28 floats] box of the [R*C][196] view
// of the pooled tensor.  That box is exactly what a 2-D TMA tensor copy moves, so here each warp owns a private
// [64][28] staging buffer (same 49 KB per CTA in total) and issues its own cp.async.bulk.tensor: no barrier after
// the tap tables are built, warps drift freely through the channel groups.
constexpr int kRowsPerWarp = 2;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(roi_smem_u32(smem_src)), "r"(x), "r"(y)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_warp(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar,
                                                 uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(roi_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          roi_smem_u32(smem_dst)),
      "l"(map), "r"(roi_smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

template <int CPL>
__global__ void __launch_bounds__(7 * 32, 4)
roi_align_fwd_cl_tma_kernel(const float* __restrict__ ft, const float* __restrict__ in_nchw,
                            const float* __restrict__ rois, float* __restrict__ out,
                            const __grid_constant__ CUtensorMap out_map, int N, int C, int H, int W, int R, float scale,
                            int sampling_ratio, int aligned, int ngroups, int gpc) {
  constexpr int P = 14, NW = 7, NT = NW * 32, PER = P * P, GC = 32 * CPL, PE = 14, WROW = kRowsPerWarp * P;
  extern __shared__ __align__(128) float dyn_smem[];
  float* stage = dyn_smem;                                       // [NW][GC][WROW]
  TapE* xtab = reinterpret_cast<TapE*>(stage + NW * GC * WROW);  // [PE * kMaxG]
  TapE* ytab = xtab + PE * kMaxG;
  const int r = blockIdx.x / ngroups;
  const int cbeg = (blockIdx.x - r * ngroups) * (GC * gpc);
  const int cend = min(C, cbeg + GC * gpc);
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, P, P, sampling_ratio, H, W);
  if (g.gw <= 0 || g.gh <= 0 || g.batch < 0 || g.batch >= N) {
    float* o = out + ((size_t)r * C + cbeg) * PER;
    for (int e = threadIdx.x; e < (cend - cbeg) * PER; e += NT) o[e] = 0.f;
    return;
  }
  if (g.gw > kMaxG || g.gh > kMaxG) {
    fwd_direct_any(in_nchw, out + (size_t)r * C * PER, cbeg, cend - cbeg, C, H, W, P, P, g, NT);
    return;
  }
  build_tables<P, NT>(xtab, ytab, g, H, W);
  __shared__ int mono_s[3];
  const bool mono = sampling_ratio <= 0 && monotone_x_table<NT>(xtab, PE * g.gw, mono_s);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wst = stage + warp * (GC * WROW);
  float* orow = wst + lane * WROW;
  bool pending = false;
  for (int c0 = cbeg; c0 < cend; c0 += GC) {  // C % GC == 0: every group is full
    const float* base = ft + (size_t)g.batch * H * W * C + c0 + lane;
    if (pending) {  // the copy engine must have read this warp's previous box before it is overwritten
      // (deferring this wait behind the first bins' loads was measured: no gain)
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
#define CDDMSL_FWD_MONO(GHV, GWV) \
  fwd_rows_mono<P, GHV, CPL, GWV>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, 32 * WROW)
    if (mono && g.gh == 1) {
#define CDDMSL_FWD_MONO_PF(GWV) \
  fwd_rows_mono_pf<P, CPL, GWV>(base, xtab, ytab, g.gw, W, C, 2 * warp, orow, 32 * WROW)
      if (g.gw == 1) CDDMSL_FWD_MONO_PF(1);
      else if (g.gw == 2) CDDMSL_FWD_MONO_PF(2);
      else CDDMSL_FWD_MONO_PF(0);
#undef CDDMSL_FWD_MONO_PF
    } else if (mono && g.gh == 2) {
      if (g.gw == 1) CDDMSL_FWD_MONO(2, 1);
      else if (g.gw == 2) CDDMSL_FWD_MONO(2, 2);
      else CDDMSL_FWD_MONO(2, 0);
    }
#undef CDDMSL_FWD_MONO
    else if (g.gh == 1) fwd_rows<P, 1, CPL>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, CPL, 32 * WROW);
    else if (g.gh == 2) fwd_rows<P, 2, CPL>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, CPL, 32 * WROW);
    else fwd_rows<P, 0, CPL>(base, xtab, ytab, g.gw, g.gh, W, C, 2 * warp, orow, CPL, 32 * WROW);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) tma_store_2d(&out_map, wst, WROW * warp, r * C + c0);
    pending = true;
  }
  if (pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the read
}

template <int CPL>
__global__ void __launch_bounds__(7 * 32, 4)
roi_align_bwd_cl_tma_kernel(const __grid_constant__ CUtensorMap gout_map, const float* __restrict__ gout,
                            const float* __restrict__ rois, float* __restrict__ gt, int N, int C, int H, int W, int R,
                            float scale, int sampling_ratio, int aligned, int ngroups, int gpc) {
  constexpr int P = 14, NW = 7, NT = NW * 32, PER = P * P, GC = 32 * CPL, PE = 14, WROW = kRowsPerWarp * P;
  extern __shared__ __align__(128) float dyn_smem[];
  float* stage = dyn_smem;                                          // [NW][GC][WROW]
  TapE* xtab = reinterpret_cast<TapE*>(stage + NW * GC * WROW);     // [PE * kMaxG]
  YSlots* yslots = reinterpret_cast<YSlots*>(xtab + PE * kMaxG);    // [NW][kMaxG]
  __shared__ __align__(8) uint64_t bars[NW];
  const int r = blockIdx.x / ngroups;
  const int cbeg = (blockIdx.x - r * ngroups) * (GC * gpc);
  const int cend = min(C, cbeg + GC * gpc);
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, P, P, sampling_ratio, H, W);
  if (g.gw <= 0 || g.gh <= 0 || g.batch < 0 || g.batch >= N) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (g.gw > kMaxG || g.gh > kMaxG) {  // reference-style scatter into the channels-last map
    const float* g_tile = gout + ((size_t)r * C + cbeg) * PER;
    float* img = gt + (size_t)g.batch * H * W * C + cbeg;
    for (int e = threadIdx.x; e < (cend - cbeg) * PER; e += NT) {
      const int c = e / PER, b = e - c * PER;
      const int ph = b / P, pw = b - ph * P;
      const float go = g_tile[e] * g.inv_count;
      for (int iy = 0; iy < g.gh; ++iy) {
        const Tap ty = make_tap(g.sh, g.bh, ph, iy, g.gh, H, 0);
        if (ty.wl == 0.f && ty.wh == 0.f) continue;
        for (int ix = 0; ix < g.gw; ++ix) {
          const Tap tx = make_tap(g.sw, g.bw, pw, ix, g.gw, W, 0);
          if (tx.wl == 0.f && tx.wh == 0.f) continue;
          red_add(img + ((size_t)ty.lo * W + tx.lo) * C + c, go * ty.wl * tx.wl);
          red_add(img + ((size_t)ty.lo * W + tx.hi) * C + c, go * ty.wl * tx.wh);
          red_add(img + ((size_t)ty.hi * W + tx.lo) * C + c, go * ty.wh * tx.wl);
          red_add(img + ((size_t)ty.hi * W + tx.hi) * C + c, go * ty.wh * tx.wh);
        }
      }
    }
    return;
  }
  float* wst = stage + warp * (GC * WROW);
  uint64_t* bar = &bars[warp];
  constexpr uint32_t kBoxBytes = GC * WROW * 4;
  if (lane == 0) {  // this warp's first box flies while the tables are built
    roi_mbar_init(bar);
    tma_load_2d_warp(wst, &gout_map, WROW * warp, r * C + cbeg, bar, kBoxBytes);
  }
  build_tables<P, NT>(xtab, nullptr, g, H, W);
  for (int t = threadIdx.x; t < NW * g.gh; t += NT) {
    const int j = t / g.gh, i = t - j * g.gh;
    yslots[j * kMaxG + i] = make_slots(make_tap_entry(g.sh, g.bh, 2 * j, i, g.gh, H, 1.f),
                                       make_tap_entry(g.sh, g.bh, 2 * j + 1, i, g.gh, H, 1.f), W * C);
  }
  __syncthreads();  // tables (and every warp's mbarrier init) visible
  uint32_t phase = 0;
  const float* grow = wst + lane * WROW;
  for (int c0 = cbeg; c0 < cend; c0 += GC) {
    float* img = gt + (size_t)g.batch * H * W * C + c0;
    roi_mbar_wait(bar, phase);
    phase ^= 1u;
    if (g.gh == 1)
      bwd_rows<P, true, CPL>(img + lane, xtab, yslots + warp * kMaxG, g.gw, g.gh, W, C, grow, CPL, true, 32 * WROW);
    else
      bwd_rows<P, false, CPL>(img + lane, xtab, yslots + warp * kMaxG, g.gw, g.gh, W, C, grow, CPL, true, 32 * WROW);
    if (c0 + GC < cend) {
      __syncwarp();  // every lane is done reading the box
      if (lane == 0) tma_load_2d_warp(wst, &gout_map, WROW * warp, r * C + c0 + GC, bar, kBoxBytes);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool roi_cl_eligible(int N, int C, int H, int W, int PH, int PW) {
  (void)N;
  return PH == PW && (PH == 14 || PH == 7) && (long long)H * W * C < 0x7fffffffLL;
}

size_t roi_cl_workspace_bytes(int N, int C, int H, int W) { return align_up((size_t)N * C * H * W * 4, 256); }

int g_fwd_cpl = 2, g_bwd_cpl = 2;
int g_roi_gpc = 2;  // 64-channel groups per CTA (tuning knob "roi_gpc")  // channels per lane (tuning knobs "roi_fwd_cpl" / "roi_bwd_cpl")

template <int P, int CPL>
static int launch_fwd_cl(const float* ft, const float* in, const float* rois, float* out, int N, int C, int H, int W,
                         int R, float scale, int sampling_ratio, int aligned, cudaStream_t stream) {
  constexpr int PE = (P + 1) & ~1, NW = (P + 1) / 2;
  const int gpc = max(1, g_roi_gpc);
  const int ngroups = ceil_div(C, 32 * CPL * gpc);
  if ((long long)R * ngroups > 0x7fffffffLL) return CDDMSL_EINVAL;
  const int smem = 32 * CPL * P * P * 4 + 2 * PE * kMaxG * (int)sizeof(TapE);
  auto k = roi_align_fwd_cl_kernel<P, CPL>;
  cudaError_t ea = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ea != cudaSuccess) return (int)ea;
  k<<<(unsigned)((long long)R * ngroups), NW * 32, smem, stream>>>(ft, in, rois, out, N, C, H, W, R, scale,
                                                                   sampling_ratio, aligned, ngroups, gpc);
  count_launch();
  return (int)cudaGetLastError();
}

template <int P, int CPL>
static int launch_bwd_cl(const float* gout, const float* rois, float* gt, int N, int C, int H, int W, int R,
                         float scale, int sampling_ratio, int aligned, cudaStream_t stream) {
  constexpr int PE = (P + 1) & ~1, NW = (P + 1) / 2;
  const int gpc = max(1, g_roi_gpc);
  const int ngroups = ceil_div(C, 32 * CPL * gpc);
  if ((long long)R * ngroups > 0x7fffffffLL) return CDDMSL_EINVAL;
  const int smem = ((32 * CPL * P * P + 3) & ~3) * 4 + PE * kMaxG * (int)sizeof(TapE) + NW * kMaxG * (int)sizeof(YSlots);
  auto k = roi_align_bwd_cl_kernel<P, CPL>;
  cudaError_t ea = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ea != cudaSuccess) return (int)ea;
  k<<<(unsigned)((long long)R * ngroups), NW * 32, smem, stream>>>(gout, rois, gt, N, C, H, W, R, scale,
                                                                   sampling_ratio, aligned, ngroups, gpc);
  count_launch();
  return (int)cudaGetLastError();
}

// tuning knob "roi_tma": per-warp TMA tensor copies for 14x14 when C % 64 == 0; bit 0 = forward, bit 1 = backward.
// Measured (cfg #2): forward 2.66 -> 2.52 ms; backward 2.57 -> 3.33 ms (sixty-four 112-byte row fetches per box cost
// more than the barrier they remove) -- so the default enables the forward only.
int g_roi_tma = 1;
constexpr int kTmaEncodeFailed = -1000;  // internal: cuTensorMapEncodeTiled unavailable / failed -> non-TMA kernels

static bool roi_tma_ok(int C, int R, int P, int bit) {
  return (g_roi_tma & bit) && P == 14 && C % 64 == 0 && (long long)R * C < 0x7fffffffLL;
}

static int launch_fwd_cl_tma(const float* ft, const float* in, const float* rois, float* out, int N, int C, int H,
                             int W, int R, float scale, int sampling_ratio, int aligned, cudaStream_t stream) {
  constexpr int CPL = 2, P = 14;
  CUtensorMap map;
  if (tma_encode_2d_f32(&map, out, P * P, (unsigned long long)R * C, P * P * 4, kRowsPerWarp * P, 32 * CPL,
                        CU_TENSOR_MAP_SWIZZLE_NONE))
    return kTmaEncodeFailed;
  const int gpc = max(1, g_roi_gpc);
  const int ngroups = ceil_div(C, 32 * CPL * gpc);
  if ((long long)R * ngroups > 0x7fffffffLL) return CDDMSL_EINVAL;
  const int smem = 32 * CPL * P * P * 4 + 2 * P * kMaxG * (int)sizeof(TapE);
  auto k = roi_align_fwd_cl_tma_kernel<CPL>;
  cudaError_t ea = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ea != cudaSuccess) return (int)ea;
  k<<<(unsigned)((long long)R * ngroups), 7 * 32, smem, stream>>>(ft, in, rois, out, map, N, C, H, W, R, scale,
                                                                 sampling_ratio, aligned, ngroups, gpc);
  count_launch();
  return (int)cudaGetLastError();
}

static int launch_bwd_cl_tma(const float* gout, const float* rois, float* gt, int N, int C, int H, int W, int R,
                             float scale, int sampling_ratio, int aligned, cudaStream_t stream) {
  constexpr int CPL = 2, P = 14;
  CUtensorMap map;
  if (tma_encode_2d_f32(&map, gout, P * P, (unsigned long long)R * C, P * P * 4, kRowsPerWarp * P, 32 * CPL,
                        CU_TENSOR_MAP_SWIZZLE_NONE))
    return kTmaEncodeFailed;
  const int gpc = max(1, g_roi_gpc);
  const int ngroups = ceil_div(C, 32 * CPL * gpc);
  if ((long long)R * ngroups > 0x7fffffffLL) return CDDMSL_EINVAL;
  const int smem = 32 * CPL * P * P * 4 + P * kMaxG * (int)sizeof(TapE) + 7 * kMaxG * (int)sizeof(YSlots);
  auto k = roi_align_bwd_cl_tma_kernel<CPL>;
  cudaError_t ea = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ea != cudaSuccess) return (int)ea;
  k<<<(unsigned)((long long)R * ngroups), 7 * 32, smem, stream>>>(map, gout, rois, gt, N, C, H, W, R, scale,
                                                                 sampling_ratio, aligned, ngroups, gpc);
  count_launch();
  return (int)cudaGetLastError();
}

int roi_align_fwd_cl(const float* in, const float* rois, float* out, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, float* ft, cudaStream_t stream) {
  int rc = launch_transpose(in, ft, N, C, H * W, stream);  // NCHW -> NHWC
  if (rc) return rc;
  if (R > 0 && roi_tma_ok(C, R, P, 1) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    rc = launch_fwd_cl_tma(ft, in, rois, out, N, C, H, W, R, scale, sampling_ratio, aligned, stream);
    if (rc != kTmaEncodeFailed) return rc;  // no tensor map: the bulk-store kernels handle the same shapes
  }
  if (P == 7) return launch_fwd_cl<7, 2>(ft, in, rois, out, N, C, H, W, R, scale, sampling_ratio, aligned, stream);
  return g_fwd_cpl == 2 ? launch_fwd_cl<14, 2>(ft, in, rois, out, N, C, H, W, R, scale, sampling_ratio, aligned, stream)
                        : launch_fwd_cl<14, 1>(ft, in, rois, out, N, C, H, W, R, scale, sampling_ratio, aligned, stream);
}

int roi_align_bwd_cl(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, float* gt, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(gt, 0, (size_t)N * C * H * W * sizeof(float), stream);
  if (e != cudaSuccess) return (int)e;
  int rc = kTmaEncodeFailed;
  if (R > 0 && roi_tma_ok(C, R, P, 2) && (reinterpret_cast<uintptr_t>(gout) & 15) == 0)
    rc = launch_bwd_cl_tma(gout, rois, gt, N, C, H, W, R, scale, sampling_ratio, aligned, stream);
  if (rc == kTmaEncodeFailed)
    rc = P == 7 ? launch_bwd_cl<7, 2>(gout, rois, gt, N, C, H, W, R, scale, sampling_ratio, aligned, stream)
           : g_bwd_cpl == 2
               ? launch_bwd_cl<14, 2>(gout, rois, gt, N, C, H, W, R, scale, sampling_ratio, aligned, stream)
               : launch_bwd_cl<14, 1>(gout, rois, gt, N, C, H, W, R, scale, sampling_ratio, aligned, stream);
  if (rc) return rc;
  return launch_transpose(gt, gin, N, H * W, C, stream);  // NHWC -> NCHW (overwrites gin completely)
}

}  // namespace cddmsl
