// ROIAlign forward / backward for sm_100a.
//
// Replaces torchvision::roi_align / ::_roi_align_backward as the reference calls them from
// detectron2/layers/roi_align.py:49-65 (aligned flag, adaptive sampling grid when sampling_ratio == 0).
//
// Design: bilinear sampling on a regular (ph,iy) x (pw,ix) grid is separable,
//   out[c] = Ay . F[c] . Ax^T / count,
// so one CTA takes (RoI, chunk of channels), stages the RoI's footprint of the NCHW feature planes in
// shared memory with coalesced loads, runs a vertical pass (lanes <-> ph, taps held in registers) and a
// horizontal pass (lanes <-> pw) entirely out of shared memory, builds the [channels,PH,PW] output tile in
// its exact global layout and writes it with one TMA bulk store (cp.async.bulk.global.shared::cta), i.e.
// full 128-byte lines and no per-thread store instructions.  The backward is the transpose: stage the
// grad tile, T = G . Ax, dF = Ay^T . T with band-limited dense tap tables, then accumulate.
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"
#include "roi_common.cuh"

namespace cddmsl {

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------

// Vertical pass: V[c][ph][x] (+)= sum_i wl_i * F[c][lo_i][x] + wh_i * F[c][hi_i][x], lanes <-> ph.
template <int G, int LPR>
__device__ __forceinline__ void fwd_pass_v(const float* __restrict__ F_s, float* __restrict__ V_s, int nch,
                                           int fh, int fw, int fws, int PH, const RoiGeom& g, int i0, bool accum,
                                           int H) {
  constexpr int NS = kThreads / LPR;
  const int lane = threadIdx.x % LPR, slot = threadIdx.x / LPR;
  int olo[G], ohi[G];
  float wl[G], wh[G];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    Tap t;
    if (lane < PH && i0 + i < g.gh) {
      t = make_tap(g.sh, g.bh, lane, i0 + i, g.gh, H, g.fy0);
    } else {
      t.lo = t.hi = 0;
      t.wl = t.wh = 0.f;
    }
    olo[i] = t.lo * fws;
    ohi[i] = t.hi * fws;
    wl[i] = t.wl;
    wh[i] = t.wh;
  }
  const int ncols = nch * fw;
  const uint32_t mg = magic_of(fw);
  const int plane = fh * fws;
  for (int col = slot; col < ncols; col += NS) {
    const int c = fast_div(col, mg, fw);
    const int x = col - c * fw;
    const float* Fc = F_s + c * plane + x;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      acc = fmaf(wl[i], Fc[olo[i]], acc);
      acc = fmaf(wh[i], Fc[ohi[i]], acc);
    }
    if (lane < PH) {
      float* v = V_s + (c * PH + lane) * fws + x;
      *v = accum ? (*v + acc) : acc;
    }
  }
}

// Horizontal pass: O[c][ph][pw] (+)= sum_i wl_i * V[c][ph][lo_i] + wh_i * V[c][ph][hi_i], lanes <-> pw.
template <int G, int LPR>
__device__ __forceinline__ void fwd_pass_h(const float* __restrict__ V_s, float* __restrict__ O_s, int nch,
                                           int fws, int PH, int PW, const RoiGeom& g, int i0, bool accum,
                                           bool last, int W) {
  constexpr int NS = kThreads / LPR;
  const int lane = threadIdx.x % LPR, slot = threadIdx.x / LPR;
  int olo[G], ohi[G];
  float wl[G], wh[G];
#pragma unroll
  for (int i = 0; i < G; ++i) {
    Tap t;
    if (lane < PW && i0 + i < g.gw) {
      t = make_tap(g.sw, g.bw, lane, i0 + i, g.gw, W, g.fx0);
    } else {
      t.lo = t.hi = 0;
      t.wl = t.wh = 0.f;
    }
    olo[i] = t.lo;
    ohi[i] = t.hi;
    wl[i] = t.wl;
    wh[i] = t.wh;
  }
  const int nrows = nch * PH;
  const float inv = g.inv_count;
  for (int row = slot; row < nrows; row += NS) {
    const float* Vr = V_s + row * fws;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      acc = fmaf(wl[i], Vr[olo[i]], acc);
      acc = fmaf(wh[i], Vr[ohi[i]], acc);
    }
    if (lane < PW) {
      float* o = O_s + row * PW + lane;
      if (accum) acc += *o;
      *o = last ? acc * inv : acc;
    }
  }
}

template <int LPR>
__device__ __forceinline__ void fwd_run_v(const float* F_s, float* V_s, int nch, int fh, int fw, int fws, int PH,
                                          const RoiGeom& g, int H) {
  switch (g.gh) {
    case 1: fwd_pass_v<1, LPR>(F_s, V_s, nch, fh, fw, fws, PH, g, 0, false, H); break;
    case 2: fwd_pass_v<2, LPR>(F_s, V_s, nch, fh, fw, fws, PH, g, 0, false, H); break;
    case 3: fwd_pass_v<3, LPR>(F_s, V_s, nch, fh, fw, fws, PH, g, 0, false, H); break;
    case 4: fwd_pass_v<4, LPR>(F_s, V_s, nch, fh, fw, fws, PH, g, 0, false, H); break;
    default:
      // more than 4 samples per bin: chunks of 4 accumulated in shared memory (each thread re-reads
      // only the V cells it wrote itself, so no barrier is needed between chunks)
      for (int i0 = 0; i0 < g.gh; i0 += 4) fwd_pass_v<4, LPR>(F_s, V_s, nch, fh, fw, fws, PH, g, i0, i0 > 0, H);
  }
}

template <int LPR>
__device__ __forceinline__ void fwd_run_h(const float* V_s, float* O_s, int nch, int fws, int PH, int PW,
                                          const RoiGeom& g, int W) {
  switch (g.gw) {
    case 1: fwd_pass_h<1, LPR>(V_s, O_s, nch, fws, PH, PW, g, 0, false, true, W); break;
    case 2: fwd_pass_h<2, LPR>(V_s, O_s, nch, fws, PH, PW, g, 0, false, true, W); break;
    case 3: fwd_pass_h<3, LPR>(V_s, O_s, nch, fws, PH, PW, g, 0, false, true, W); break;
    case 4: fwd_pass_h<4, LPR>(V_s, O_s, nch, fws, PH, PW, g, 0, false, true, W); break;
    default:
      for (int i0 = 0; i0 < g.gw; i0 += 4)
        fwd_pass_h<4, LPR>(V_s, O_s, nch, fws, PH, PW, g, i0, i0 > 0, i0 + 4 >= g.gw, W);
  }
}

// Reference-order evaluation straight from global memory; used when a footprint does not fit the
// shared-memory budget (very large feature maps) and for pooled sizes above 16.
__device__ void fwd_direct(const float* __restrict__ in, float* __restrict__ out_roi, int c0, int nch, int C,
                           int H, int W, int PH, int PW, const RoiGeom& g) {
  const int per = PH * PW;
  for (int e = threadIdx.x; e < nch * per; e += kThreads) {
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
    out_roi[(size_t)(c0 + c) * per + e - c * per] = acc * g.inv_count;
  }
}

template <int LPR>
__global__ void __launch_bounds__(kThreads, 3)
roi_align_fwd_kernel(const float* __restrict__ in, const float* __restrict__ rois, float* __restrict__ out, int N,
                     int C, int H, int W, int R, int PH, int PW, float scale, int sampling_ratio, int aligned,
                     int CC, int nchunks, int smem_floats) {
  extern __shared__ __align__(128) float smem[];
  const int r = blockIdx.x / nchunks;
  const int c0 = (blockIdx.x - r * nchunks) * CC;
  const int nc_cta = min(CC, C - c0);
  const int per = PH * PW;
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio, H, W);
  float* out_roi = out + (size_t)r * C * per;

  if (g.gw <= 0 || g.gh <= 0 || g.batch < 0 || g.batch >= N) {  // empty sampling grid -> zeros
    for (int e = threadIdx.x; e < nc_cta * per; e += kThreads) out_roi[(size_t)c0 * per + e] = 0.f;
    return;
  }
  const int fw = g.fw, fh = g.fh;
  const int fws = fw | 1;  // odd row stride: lanes that differ in their row hit different banks
  const int per_pad = (per + 3) & ~3;
  const int per_ch = fh * fws + PH * fws + per_pad;
  int ccs = min(nc_cta, smem_floats / per_ch);
  if (ccs <= 0) {
    fwd_direct(in, out_roi, c0, nc_cta, C, H, W, PH, PW, g);
    return;
  }
  // O_s first so that it is 16-byte aligned for the bulk store
  float* O_s = smem;
  float* F_s = O_s + ((ccs * per + 3) & ~3);
  float* V_s = F_s + ccs * fh * fws;
  const bool bulk_ok = (per % 4 == 0);
  const uint32_t mg_fw = magic_of(fw), mg_fh = magic_of(fh);
  bool pending = false;

  for (int cs = 0; cs < nc_cta; cs += ccs) {
    const int nch = min(ccs, nc_cta - cs);
    // ---- stage the footprint: coalesced along x, rows of the plane are W floats apart
    {
      const float* src = in + (((size_t)g.batch * C + c0 + cs) * H + g.fy0) * W + g.fx0;
      const int total = nch * fh * fw;
      for (int e = threadIdx.x; e < total; e += kThreads) {
        const int row = fast_div(e, mg_fw, fw);
        const int x = e - row * fw;
        const int c = fast_div(row, mg_fh, fh);
        const int y = row - c * fh;
        F_s[row * fws + x] = __ldg(src + ((size_t)c * H + y) * W + x);
      }
    }
    __syncthreads();
    fwd_run_v<LPR>(F_s, V_s, nch, fh, fw, fws, PH, g, H);
    if (pending && threadIdx.x == 0) {  // previous bulk store must have finished reading O_s
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      pending = false;
    }
    __syncthreads();
    fwd_run_h<LPR>(V_s, O_s, nch, fws, PH, PW, g, W);
    float* dst = out_roi + (size_t)(c0 + cs) * per;
    if (bulk_ok) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(O_s);
        const uint32_t bytes = (uint32_t)(nch * per) * 4u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        pending = true;
      }
    } else {
      __syncthreads();
      for (int e = threadIdx.x; e < nch * per; e += kThreads) dst[e] = O_s[e];
      __syncthreads();
    }
  }
  if (pending && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Dense, band-limited tap tables of one RoI in shared memory:
//   Ax[pw][x] = sum_ix weight of column x in bin pw (times 1/count),  band [pf[x], pl[x]] of bins touching x
//   Ay[ph][y] likewise, band [qf[y], ql[y]].
struct BwdTables {
  float* Ax;   // [PW][fws]
  float* Ay;   // [PH][fhs]
  int* pf;     // [fw]
  int* pl;
  int* qf;     // [fh]
  int* ql;
};

__device__ __forceinline__ void bwd_build_tables(const BwdTables& t, const RoiGeom& g, int PH, int PW, int fws,
                                                 int fhs, int H, int W) {
  for (int e = threadIdx.x; e < PW * fws; e += kThreads) t.Ax[e] = 0.f;
  for (int e = threadIdx.x; e < PH * fhs; e += kThreads) t.Ay[e] = 0.f;
  __syncthreads();
  if (threadIdx.x < PW) {  // thread owns row pw of Ax: no races
    float* row = t.Ax + threadIdx.x * fws;
    for (int i = 0; i < g.gw; ++i) {
      const Tap tp = make_tap(g.sw, g.bw, threadIdx.x, i, g.gw, W, g.fx0);
      row[tp.lo] += tp.wl * g.inv_count;
      row[tp.hi] += tp.wh * g.inv_count;
    }
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + PH) {
    const int ph = threadIdx.x - 32;
    float* row = t.Ay + ph * fhs;
    for (int i = 0; i < g.gh; ++i) {
      const Tap tp = make_tap(g.sh, g.bh, ph, i, g.gh, H, g.fy0);
      row[tp.lo] += tp.wl;
      row[tp.hi] += tp.wh;
    }
  }
  __syncthreads();
  for (int x = threadIdx.x; x < g.fw; x += kThreads) {
    int f = PW, l = -1;
    for (int p = 0; p < PW; ++p)
      if (t.Ax[p * fws + x] != 0.f) {
        f = min(f, p);
        l = p;
      }
    t.pf[x] = f;
    t.pl[x] = l;
  }
  for (int y = threadIdx.x; y < g.fh; y += kThreads) {
    int f = PH, l = -1;
    for (int p = 0; p < PH; ++p)
      if (t.Ay[p * fhs + y] != 0.f) {
        f = min(f, p);
        l = p;
      }
    t.qf[y] = f;
    t.ql[y] = l;
  }
  __syncthreads();
}

// Scatter variant: one CTA per (RoI, channel chunk); footprint gradient built without atomics in shared
// memory, then added to gin with one red.global.add.f32 per footprint cell.
__global__ void __launch_bounds__(kThreads)
roi_align_bwd_red_kernel(const float* __restrict__ gout, const float* __restrict__ rois, float* __restrict__ gin,
                         int N, int C, int H, int W, int R, int PH, int PW, float scale, int sampling_ratio,
                         int aligned, int CC, int nchunks, int smem_floats) {
  extern __shared__ __align__(128) float smem[];
  const int r = blockIdx.x / nchunks;
  const int c0 = (blockIdx.x - r * nchunks) * CC;
  const int nc_cta = min(CC, C - c0);
  const int per = PH * PW;
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, PH, PW, sampling_ratio, H, W);
  if (g.gw <= 0 || g.gh <= 0 || g.batch < 0 || g.batch >= N) return;
  const int fw = g.fw, fh = g.fh, fws = fw | 1, fhs = fh | 1;

  BwdTables t;
  t.Ax = smem;
  t.Ay = t.Ax + PW * fws;
  t.pf = (int*)(t.Ay + PH * fhs);
  t.pl = t.pf + fw;
  t.qf = t.pl + fw;
  t.ql = t.qf + fh;
  float* dyn = (float*)(t.ql + fh);
  dyn = (float*)(((uintptr_t)dyn + 15) & ~(uintptr_t)15);
  const int used = (int)(dyn - smem);
  const int per_pad = (per + 3) & ~3;
  const int per_ch = per_pad + PH * fws;
  const int ccs = min(nc_cta, (smem_floats - used) / per_ch);
  const float* g_roi = gout + (size_t)r * C * per;

  if (ccs <= 0) {
    // footprint too large for shared memory: reference-style scatter, 4 atomics per sample
    for (int e = threadIdx.x; e < nc_cta * per; e += kThreads) {
      const int c = e / per, b = e - c * per;
      const int ph = b / PW, pw = b - ph * PW;
      float* plane = gin + ((size_t)g.batch * C + c0 + c) * H * W;
      const float go = g_roi[(size_t)(c0 + c) * per + b] * g.inv_count;
      for (int iy = 0; iy < g.gh; ++iy) {
        const Tap ty = make_tap(g.sh, g.bh, ph, iy, g.gh, H, 0);
        for (int ix = 0; ix < g.gw; ++ix) {
          const Tap tx = make_tap(g.sw, g.bw, pw, ix, g.gw, W, 0);
          if (ty.wl == 0.f && ty.wh == 0.f) continue;
          if (tx.wl == 0.f && tx.wh == 0.f) continue;
          atomicAdd(plane + ty.lo * W + tx.lo, go * ty.wl * tx.wl);
          atomicAdd(plane + ty.lo * W + tx.hi, go * ty.wl * tx.wh);
          atomicAdd(plane + ty.hi * W + tx.lo, go * ty.wh * tx.wl);
          atomicAdd(plane + ty.hi * W + tx.hi, go * ty.wh * tx.wh);
        }
      }
    }
    return;
  }
  bwd_build_tables(t, g, PH, PW, fws, fhs, H, W);
  float* G_s = dyn;
  float* T_s = G_s + ccs * per_pad;  // (per_pad only to keep T_s 16-byte aligned)

  // lanes per row: smallest power of two >= fw, within [8, 32]
  const int lr_shift = fw <= 8 ? 3 : (fw <= 16 ? 4 : 5);
  const int LR = 1 << lr_shift;
  const int lane = threadIdx.x & (LR - 1);
  const int slot = threadIdx.x >> lr_shift;
  const int nslots = kThreads >> lr_shift;
  const uint32_t mg_fh = magic_of(fh);

  for (int cs = 0; cs < nc_cta; cs += ccs) {
    const int nch = min(ccs, nc_cta - cs);
    {  // stage the contiguous grad tile
      const float* src = g_roi + (size_t)(c0 + cs) * per;
      const int total = nch * per;
      if ((per & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(G_s);
        for (int e = threadIdx.x; e < total / 4; e += kThreads) d4[e] = __ldg(s4 + e);
      } else {
        for (int e = threadIdx.x; e < total; e += kThreads) G_s[e] = __ldg(src + e);
      }
    }
    __syncthreads();
    // T[c][ph][x] = sum_{pw in band(x)} G[c][ph][pw] * Ax[pw][x]
    for (int row = slot; row < nch * PH; row += nslots) {
      const float* Gr = G_s + row * PW;
      for (int x = lane; x < fw; x += LR) {
        float acc = 0.f;
        const int p1 = t.pl[x];
        for (int p = t.pf[x]; p <= p1; ++p) acc = fmaf(Gr[p], t.Ax[p * fws + x], acc);
        T_s[row * fws + x] = acc;
      }
    }
    __syncthreads();
    // dF[c][y][x] = sum_{ph in band(y)} Ay[ph][y] * T[c][ph][x]
    for (int row = slot; row < nch * fh; row += nslots) {
      const int c = fast_div(row, mg_fh, fh);
      const int y = row - c * fh;
      const int q0 = t.qf[y], q1 = t.ql[y];
      if (q1 < q0) continue;
      float* dst = gin + (((size_t)g.batch * C + c0 + cs + c) * H + g.fy0 + y) * W + g.fx0;
      const float* Tc = T_s + c * PH * fws;
      for (int x = lane; x < fw; x += LR) {
        if (t.pl[x] < t.pf[x]) continue;
        float acc = 0.f;
        for (int p = q0; p <= q1; ++p) acc = fmaf(t.Ay[p * fhs + y], Tc[p * fws + x], acc);
        atomicAdd(dst + x, acc);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_fwd_smem_kb = 56;
static int g_fwd_cc = 32;
static int g_bwd_smem_kb = 56;
static int g_bwd_cc = 32;
static int g_use_cl = 1;
extern int g_fwd_cpl, g_bwd_cpl, g_roi_gpc, g_roi_tma;  // channels-last fast path (roi_align_cl.cu) for the 14x14 pooler

bool roi_cl_eligible(int N, int C, int H, int W, int PH, int PW);
size_t roi_cl_workspace_bytes(int N, int C, int H, int W);
int roi_align_fwd_cl(const float* in, const float* rois, float* out, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, float* ft, cudaStream_t stream);
// plane-resident path (roi_align_pr.cu): map slices stay in shared memory, lanes are output columns
bool roi_pr_eligible(int N, int C, int H, int W, int R, int P, int bit);
size_t roi_pr_workspace_bytes(int N, int R);
int roi_align_fwd_pr(const float* in, const float* rois, float* out, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, void* ws, cudaStream_t stream,
                     const float* in2 = nullptr, float* out2 = nullptr);
int roi_align_bwd_pr(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, void* ws, cudaStream_t stream);
bool roi_pr_bwd_eligible(int N, int C, int H, int W, int R, int P);
int tune_roi_pr(const char* key, int value);
// row-walk backward (roi_align_rw.cu): one warp per (RoI, 32 channels), one RED per footprint cell and channel
bool roi_rw_eligible(int N, int C, int H, int W, int R, int P);
size_t roi_rw_workspace_bytes(int N, int C, int H, int W, int R);
int roi_align_bwd_rw(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, void* ws, cudaStream_t stream);
int tune_roi_rw(const char* key, int value);
int roi_align_bwd_cl(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, float* gt, cudaStream_t stream);

int tune_roi(const char* key, int value) {
  if (!strcmp(key, "roi_fwd_smem_kb")) g_fwd_smem_kb = value;
  else if (!strcmp(key, "roi_fwd_cc")) g_fwd_cc = value;
  else if (!strcmp(key, "roi_bwd_smem_kb")) g_bwd_smem_kb = value;
  else if (!strcmp(key, "roi_bwd_cc")) g_bwd_cc = value;
  else if (!strcmp(key, "roi_use_cl")) g_use_cl = value;
  else if (!strcmp(key, "roi_fwd_cpl")) g_fwd_cpl = value;
  else if (!strcmp(key, "roi_bwd_cpl")) g_bwd_cpl = value;
  else if (!strcmp(key, "roi_gpc")) g_roi_gpc = value;
  else if (!strcmp(key, "roi_tma")) g_roi_tma = value;
  else return tune_roi_pr(key, value) || tune_roi_rw(key, value);
  return 1;
}

static size_t roi_workspace_bytes(int N, int C, int H, int W, int R) {
  N = N > 0 ? N : 0, C = C > 0 ? C : 0, H = H > 0 ? H : 0, W = W > 0 ? W : 0, R = R > 0 ? R : 0;
  const size_t a = roi_cl_workspace_bytes(N, C, H, W), b = roi_pr_workspace_bytes(N, R);
  const size_t c = roi_rw_workspace_bytes(N, C, H, W, R);
  return (a > b ? (a > c ? a : c) : (b > c ? b : c)) + 256;
}

}  // namespace cddmsl

using namespace cddmsl;

extern "C" size_t cddmsl_roi_align_fwd_workspace_bytes(int N, int C, int H, int W, int R) {
  return roi_workspace_bytes(N, C, H, W, R);
}

extern "C" int cddmsl_roi_align_fwd(const float* in, const float* rois, float* out, int N, int C, int H, int W,
                                    int R, int PH, int PW, float spatial_scale, int sampling_ratio, int aligned,
                                    void* workspace, size_t workspace_bytes, cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || C < 0 || H < 0 || W < 0 || R < 0 || PH <= 0 || PW <= 0) return CDDMSL_EINVAL;
  if (R == 0 || C == 0) return CDDMSL_OK;
  if (!rois || !out) return CDDMSL_EINVAL;
  if (N == 0 || H == 0 || W == 0) {  // no image to sample from: defined as zeros
    CDDMSL_CUDA(cudaMemsetAsync(out, 0, (size_t)R * C * PH * PW * sizeof(float), stream));
    return CDDMSL_OK;
  }
  if (!in) return CDDMSL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return CDDMSL_EALIGN;
  const bool ws_ok = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0;
  if (PH == PW && ws_ok && workspace_bytes >= roi_pr_workspace_bytes(N, R) && roi_pr_eligible(N, C, H, W, R, PH, 1))
    return roi_align_fwd_pr(in, rois, out, N, C, H, W, R, PH, spatial_scale, sampling_ratio, aligned, workspace,
                            stream);
  if (g_use_cl && roi_cl_eligible(N, C, H, W, PH, PW) && workspace &&
      workspace_bytes >= roi_cl_workspace_bytes(N, C, H, W) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0)
    return roi_align_fwd_cl(in, rois, out, N, C, H, W, R, PH, spatial_scale, sampling_ratio, aligned,
                            (float*)workspace, stream);
  const int CC = min(g_fwd_cc, C);
  const int nchunks = ceil_div(C, CC);
  if ((long long)R * nchunks > 0x7fffffffLL) return CDDMSL_EINVAL;
  const int smem_bytes = g_fwd_smem_kb * 1024;
  const int smem_floats = (PH > 16 || PW > 16) ? 0 : smem_bytes / 4;  // > 16 bins per side: direct path
  dim3 grid((unsigned)((long long)R * nchunks)), block(kThreads);
  if (PH <= 8 && PW <= 8) {
    auto k = roi_align_fwd_kernel<8>;
    CDDMSL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    k<<<grid, block, smem_bytes, stream>>>(in, rois, out, N, C, H, W, R, PH, PW, spatial_scale, sampling_ratio,
                                           aligned, CC, nchunks, smem_floats);
  } else {
    auto k = roi_align_fwd_kernel<16>;
    CDDMSL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    k<<<grid, block, smem_bytes, stream>>>(in, rois, out, N, C, H, W, R, PH, PW, spatial_scale, sampling_ratio,
                                           aligned, CC, nchunks, smem_floats);
  }
  count_launch();
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

extern "C" size_t cddmsl_roi_align_bwd_workspace_bytes(int N, int C, int H, int W, int R) {
  return roi_workspace_bytes(N, C, H, W, R);
}

extern "C" int cddmsl_roi_align_bwd(const float* gout, const float* rois, float* gin, int N, int C, int H, int W,
                                    int R, int PH, int PW, float spatial_scale, int sampling_ratio, int aligned,
                                    void* workspace, size_t workspace_bytes, cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (N < 0 || C < 0 || H < 0 || W < 0 || R < 0 || PH <= 0 || PW <= 0) return CDDMSL_EINVAL;
  const size_t gin_bytes = (size_t)N * C * H * W * sizeof(float);
  if (gin_bytes == 0) return CDDMSL_OK;
  if (!gin) return CDDMSL_EINVAL;
  if (R == 0) {
    CDDMSL_CUDA(cudaMemsetAsync(gin, 0, gin_bytes, stream));
    return CDDMSL_OK;
  }
  if (!gout || !rois) return CDDMSL_EINVAL;
  if ((reinterpret_cast<uintptr_t>(gout) & 15) != 0) return CDDMSL_EALIGN;
  if (PH == PW && workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
      workspace_bytes >= roi_pr_workspace_bytes(N, R) && roi_pr_bwd_eligible(N, C, H, W, R, PH))
    return roi_align_bwd_pr(gout, rois, gin, N, C, H, W, R, PH, spatial_scale, sampling_ratio, aligned, workspace,
                            stream);
  if (PH == PW && workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 &&
      workspace_bytes >= roi_rw_workspace_bytes(N, C, H, W, R) && roi_rw_eligible(N, C, H, W, R, PH)) {
    const int rc = roi_align_bwd_rw(gout, rois, gin, N, C, H, W, R, PH, spatial_scale, sampling_ratio, aligned,
                                    workspace, stream);
    if (rc != -1000) return rc;  // -1000: no tensor map available, nothing was launched
  }
  if (g_use_cl && roi_cl_eligible(N, C, H, W, PH, PW) && workspace &&
      workspace_bytes >= roi_cl_workspace_bytes(N, C, H, W) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0)
    return roi_align_bwd_cl(gout, rois, gin, N, C, H, W, R, PH, spatial_scale, sampling_ratio, aligned,
                            (float*)workspace, stream);
  CDDMSL_CUDA(cudaMemsetAsync(gin, 0, gin_bytes, stream));
  const int CC = min(g_bwd_cc, C);
  const int nchunks = ceil_div(C, CC);
  if ((long long)R * nchunks > 0x7fffffffLL) return CDDMSL_EINVAL;
  const int smem_bytes = g_bwd_smem_kb * 1024;
  const int smem_floats = (PH > 32 || PW > 32) ? 0 : smem_bytes / 4;
  dim3 grid((unsigned)((long long)R * nchunks)), block(kThreads);
  CDDMSL_CUDA(cudaFuncSetAttribute(roi_align_bwd_red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   smem_bytes));
  roi_align_bwd_red_kernel<<<grid, block, smem_bytes, stream>>>(gout, rois, gin, N, C, H, W, R, PH, PW,
                                                                 spatial_scale, sampling_ratio, aligned, CC, nchunks,
                                                                 smem_floats);
  count_launch();
  CDDMSL_CHECK_LAUNCH();
  return CDDMSL_OK;
}

// ---- dual-map variants: the same RoIs on two feature maps of the same shape ---------------------------------------
extern "C" int cddmsl_roi_align_fwd2(const float* in_a, const float* in_b, const float* rois, float* out_a,
                                     float* out_b, int N, int C, int H, int W, int R, int PH, int PW,
                                     float spatial_scale, int sampling_ratio, int aligned, void* workspace,
                                     size_t workspace_bytes, cddmsl_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool ws_ok = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0;
  if (N > 0 && C > 0 && H > 0 && W > 0 && R > 0 && PH == PW && in_a && in_b && rois && out_a && out_b && ws_ok &&
      (reinterpret_cast<uintptr_t>(out_a) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_b) & 15) == 0 &&
      workspace_bytes >= roi_pr_workspace_bytes(N, R) && roi_pr_eligible(N, C, H, W, R, PH, 1))
    return roi_align_fwd_pr(in_a, rois, out_a, N, C, H, W, R, PH, spatial_scale, sampling_ratio, aligned, workspace,
                            stream, in_b, out_b);
  int rc = cddmsl_roi_align_fwd(in_a, rois, out_a, N, C, H, W, R, PH, PW, spatial_scale, sampling_ratio, aligned,
                                workspace, workspace_bytes, stream_);
  if (rc) return rc;
  return cddmsl_roi_align_fwd(in_b, rois, out_b, N, C, H, W, R, PH, PW, spatial_scale, sampling_ratio, aligned,
                              workspace, workspace_bytes, stream_);
}

extern "C" int cddmsl_roi_align_bwd2(const float* gout_a, const float* gout_b, const float* rois, float* gin_a,
                                     float* gin_b, int N, int C, int H, int W, int R, int PH, int PW,
                                     float spatial_scale, int sampling_ratio, int aligned, void* workspace,
                                     size_t workspace_bytes, cddmsl_stream_t stream_) {
  int rc = cddmsl_roi_align_bwd(gout_a, rois, gin_a, N, C, H, W, R, PH, PW, spatial_scale, sampling_ratio, aligned,
                                workspace, workspace_bytes, stream_);
  if (rc) return rc;
  return cddmsl_roi_align_bwd(gout_b, rois, gin_b, N, C, H, W, R, PH, PW, spatial_scale, sampling_ratio, aligned,
                              workspace, workspace_bytes, stream_);
}
