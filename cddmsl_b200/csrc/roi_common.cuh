// Geometry / tap arithmetic shared by the ROIAlign kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace cddmsl {

constexpr int kThreads = 256;

struct RoiGeom {
  int batch;
  float sw, sh, bw, bh;   // roi start (feature coords) and bin size
  int gw, gh;             // sampling grid per bin
  float inv_count;
  int fx0, fw, fy0, fh;   // footprint (origin, extent) in the feature plane
};

// Same arithmetic as the reference op (see oracle/c/roi_align_ref.c:roi_params).
__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int aligned, int PH,
                                            int PW, int sampling_ratio, int H, int W) {
  RoiGeom g;
  g.batch = (int)roi[0];
  // Unfused (non-FMA) arithmetic, term for term as the CPU kernel the oracle runs: on a 300-cell-wide map one
  // ulp of a coordinate is 3e-5, which a contracted multiply-add would turn into a 1e-5 relative output error.
  const float offset = aligned ? 0.5f : 0.0f;
  g.sw = __fsub_rn(__fmul_rn(roi[1], scale), offset);
  g.sh = __fsub_rn(__fmul_rn(roi[2], scale), offset);
  const float ew = __fsub_rn(__fmul_rn(roi[3], scale), offset);
  const float eh = __fsub_rn(__fmul_rn(roi[4], scale), offset);
  float rw = __fsub_rn(ew, g.sw), rh = __fsub_rn(eh, g.sh);
  if (!aligned) {
    rw = fmaxf(rw, 1.0f);
    rh = fmaxf(rh, 1.0f);
  }
  g.bh = __fdiv_rn(rh, (float)PH);
  g.bw = __fdiv_rn(rw, (float)PW);
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  const int cnt = g.gh * g.gw;
  g.inv_count = 1.0f / (float)(cnt > 1 ? cnt : 1);
  // conservative footprint of all valid samples (one cell of slack for rounding on either side)
  const float xa = g.sw, xb = g.sw + (float)PW * g.bw;
  const float ya = g.sh, yb = g.sh + (float)PH * g.bh;
  const float xlo = fminf(fmaxf(fminf(xa, xb), -2.f), (float)W + 2.f);
  const float xhi = fminf(fmaxf(fmaxf(xa, xb), -2.f), (float)W + 2.f);
  const float ylo = fminf(fmaxf(fminf(ya, yb), -2.f), (float)H + 2.f);
  const float yhi = fminf(fmaxf(fmaxf(ya, yb), -2.f), (float)H + 2.f);
  g.fx0 = min(max((int)floorf(xlo) - 1, 0), W - 1);
  g.fy0 = min(max((int)floorf(ylo) - 1, 0), H - 1);
  const int fx1 = min(max((int)floorf(xhi) + 2, 0), W - 1);
  const int fy1 = min(max((int)floorf(yhi) + 2, 0), H - 1);
  g.fw = fx1 - g.fx0 + 1;
  g.fh = fy1 - g.fy0 + 1;
  return g;
}

struct Tap {
  int lo, hi;      // indices relative to the footprint origin
  float wl, wh;    // weights of lo / hi (both 0 for a sample outside [-1, L])
};

// One 1-D bilinear tap pair; `p` bin index, `i` sample index inside the bin.  The coordinate expression is
// the reference's, term for term: start + p*bin + (i + .5)*bin/g.
__device__ __forceinline__ Tap make_tap(float start, float bin, int p, int i, int g, int L, int f0) {
  float v = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                      __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)g));  // no FMA contraction
  Tap t;
  if (v < -1.0f || v > (float)L) {
    t.lo = t.hi = 0;
    t.wl = t.wh = 0.f;
    return t;
  }
  if (v <= 0.f) v = 0.f;
  int lo = (int)v, hi;
  if (lo >= L - 1) {
    hi = lo = L - 1;
    v = (float)lo;
  } else {
    hi = lo + 1;
  }
  const float l = __fsub_rn(v, (float)lo);
  t.lo = lo - f0;
  t.hi = hi - f0;
  t.wl = 1.f - l;
  t.wh = l;
  return t;
}

__device__ __forceinline__ uint32_t magic_of(int d) { return (uint32_t)(0xFFFFFFFFu / (uint32_t)d) + 1u; }
// q = n / d for n*d < 2^32 (all uses here are < 2^24)
__device__ __forceinline__ int fast_div(int n, uint32_t magic, int d) {
  return d == 1 ? n : (int)__umulhi((uint32_t)n, magic);
}


}  // namespace cddmsl
