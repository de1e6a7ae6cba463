// ROIAlign forward / backward, PLANE-RESIDENT formulation (round 2; the default for the 14x14 / 7x7 poolers of the
// reference's configs whenever a few channel planes of one image fit into shared memory).
//
// Why (evidence: profiles/ncu_full_r1_summary.csv): the RoI-centric channels-last kernels of roi_align_cl.cu moved
// exactly the algorithmic DRAM bytes, yet every (RoI, channel) re-read its footprint from L2 -- 6.1 GB of L2->SM
// traffic in the forward, 12 GB of L2 reductions in the backward -- and sat at 43-46 % of the HBM roofline on exposed
// L2 latency (forward) and instruction issue (backward).  The RoIs of an image overlap ~40x on the res4 map, so here
// the MAP is what stays put:
//   * a persistent CTA (one per SM) takes a unit = (image, group of CH channels, chunk of that image's RoIs), copies
//     the CH planes [H x W] of the NCHW input into shared memory ONCE (no channels-last transpose, 153 KB for 16
//     planes of a 38x63 map) and then serves every tap of every RoI of the chunk from shared memory (29 cycles, no
//     L2 traffic: the forward's only global traffic is the compulsory output stream);
//   * the lanes of a warp are the OUTPUT COLUMNS (14 bins x 2 channel slots, 7 bins x 4): the x-taps of a lane are
//     fixed per RoI and live in registers, the y-taps are warp-uniform, consecutive lanes write consecutive floats
//     of out[r][c][ph][:] -- no staging tile, no bank-conflicting transposition;
//   * bilinear taps of all samples of a bin are merged per cell into a short dense band (<= g+1 cells per axis
//     instead of 2g taps); a tiny planning kernel builds the bands once per RoI with the reference's exact,
//     unfused coordinate arithmetic (roi_common.cuh) and buckets the RoIs by image;
//   * one sample per bin vertically (three quarters of the RoIs): a register window over the two blended rows,
//     so a footprint row is read once per RoI and channel.
// The backward is the transpose on a zero-initialised gradient slice in shared memory (no global atomics at all).
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"
#include "roi_common.cuh"

namespace cddmsl {

constexpr int kPrNW = 14;       // longest band (cells) a record holds; longer bands take the per-sample path
constexpr int kPrMaxG = 32;     // samples per bin and axis the planner merges; beyond: per-sample path
constexpr int kPrThreads = 576;
constexpr int kPrWarps = kPrThreads / 32;
constexpr int kPrNB = 4;        // cost buckets per image (heaviest first): units start with the expensive RoIs and
                                // concurrently running warps mostly execute the same code variant

// One bin of one axis = the merged taps of its samples: first cell of the band, number of cells (0: the bin
// contributes nothing) and one weight per cell (x bins: already divided by the sample count).  Stored as two
// 32-byte halves: A = {start, n, w[0..5]} -- all the reference's own configurations ever need and what the warps
// prefetch into shared memory -- and B = {w[6..13]}, read from global memory by the rare wider bands.
struct __align__(16) PrRecA {
  int start, n;
  float w[6];
};
struct __align__(16) PrRecB {
  float w[8];
};
static_assert(sizeof(PrRecA) == 32 && sizeof(PrRecB) == 32, "record halves are moved in 16-byte pieces");

enum { PR_ZERO = 0, PR_BAND = 1, PR_DIRECT = 3 };
struct __align__(16) PrHdr {  // slot 0 of a block
  int cls, nxmax, nymax, roi;
  int x0, fw, nimax, pad;  // backward: first footprint column, footprint width, most bins sharing one column
};
static_assert(sizeof(PrHdr) == 32, "header fills slot 0");
// One RoI = one block of 29 slots: [header][x bins 0..13][y bins 0..13]; blocks are stored in (image, cost bucket)
// order, so a chunk of an image's RoIs is a contiguous run of blocks.
constexpr int kPrBlockSlots = 29;
constexpr int kPrBlockBytes = kPrBlockSlots * 32;   // A half: 928 bytes
constexpr int kPrBlockVec = kPrBlockBytes / 16;     // 58 16-byte pieces
constexpr int kPrBlockFloats = kPrBlockBytes / 4;   // 232

// Inverse x table of the backward (per RoI, kPrInvBytes): for every footprint column (or column part, when several
// lanes share a column of a narrow footprint) the contiguous window of bins whose band covers it and their weights.
//   fmt 1 (fw <= 16): 16 entries of 32 bytes {int first_bin; float w[7]}, 2^logl lanes per column, nps bins per lane
//   fmt 2 (fw <= 128, at most 3 bins per column): fw entries of 16 bytes {int first_bin; float w[3]}
// hdr.pad = fmt | logl << 8 | nps << 16 (fmt 0: no table -> per-sample path).
constexpr int kPrInvBytes = 2048, kPrInvFloats = kPrInvBytes / 4;
constexpr int kPrInvPrefetch = 1024;  // bytes fetched ahead for every RoI (all of fmt 1, fmt 2 up to 64 columns)

struct __align__(16) PrChunk {
  int img, begin, end, pad;
};
struct PrPlan {
  int* counter;  // [0] unit counter (fwd), [1] number of chunks, [2] unit counter (bwd)
  int* counts;   // [N * kPrNB] RoIs per (image, bucket)
  int* starts;   // [N * kPrNB + 1]
  int* cursor;   // [N * kPrNB]
  PrChunk* chunks;
  PrRecA* blocksA;  // [R][29]
  PrRecB* blocksB;  // [R][29]
  float* inv;       // [R][kPrInvFloats]: inverse x tables of the backward (NULL: not wanted)
};

// ------------------------------------------------------------------------------------------------
// planning: RoIs per (image, cost bucket), chunk list, merged bands per RoI
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pr_bucket(const RoiGeom& g, bool valid) {
  if (!valid) return kPrNB - 1;
  const long long s = (long long)g.gw * g.gh;
  return s >= 6 ? 0 : s >= 3 ? 1 : s == 2 ? 2 : 3;
}

template <int P>
__global__ void __launch_bounds__(256) pr_plan_count_kernel(const float* __restrict__ rois, int R, int N, int H, int W,
                                                            float scale, int sampling_ratio, int aligned,
                                                            int* __restrict__ counts) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= R) return;
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, P, P, sampling_ratio, H, W);
  const bool valid = g.gw > 0 && g.gh > 0 && g.batch >= 0 && g.batch < N;
  atomicAdd(&counts[(valid ? g.batch : 0) * kPrNB + pr_bucket(g, valid)], 1);
}

// Block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, *total = sum.
__device__ __forceinline__ int pr_block_scan(int v, int* s_warp /* [32] */, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += a;
  }
  __syncthreads();  // s_warp may still be read from a previous call
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int a = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int x = __shfl_up_sync(0xffffffffu, a, o);
      if (lane >= o) a += x;
    }
    s_warp[lane] = a;
  }
  __syncthreads();
  *total = s_warp[31];
  return (warp ? s_warp[warp - 1] : 0) + inc - v;
}

// One CTA: exclusive scan of the (image, bucket) counts, the chunk list, and the reset of the work counters.
__global__ void __launch_bounds__(1024) pr_plan_scan_kernel(PrPlan plan, int N, int chunk) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int E = N * kPrNB;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < E; base += 1024) {
    const int e = base + threadIdx.x;
    const int c = e < E ? plan.counts[e] : 0;
    int total;
    const int off = pr_block_scan(c, s_warp, &total) + s_carry;
    if (e < E) {
      plan.starts[e] = off;
      plan.cursor[e] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) plan.starts[E] = s_carry;
  __syncthreads();
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {  // chunks of each image's (bucket-ordered) run of RoIs
    const int i = base + threadIdx.x;
    const int b0 = i < N ? plan.starts[i * kPrNB] : 0;
    const int c = i < N ? plan.starts[(i + 1) * kPrNB] - b0 : 0;
    const int k = (c + chunk - 1) / chunk;
    int total;
    const int off = pr_block_scan(k, s_warp, &total) + s_carry;
    for (int q = 0; q < k; ++q) {
      PrChunk ck;
      ck.img = i;
      ck.begin = b0 + q * chunk;
      ck.end = min(b0 + c, ck.begin + chunk);
      ck.pad = 0;
      plan.chunks[off + q] = ck;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    plan.counter[0] = 0;
    plan.counter[1] = s_carry;
    plan.counter[2] = 0;
  }
}

// One warp per RoI: merged bands with the reference's coordinate arithmetic, written at the RoI's sorted position.
template <int P>
__global__ void __launch_bounds__(256) pr_plan_fill_kernel(const float* __restrict__ rois, int R, int N, int H, int W,
                                                           float scale, int sampling_ratio, int aligned, PrPlan plan) {
  static_assert(P <= 14, "two bins axes in one warp");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= R) return;
  const RoiGeom g = roi_geom(rois + (size_t)r * 5, scale, aligned, P, P, sampling_ratio, H, W);
  const bool valid = g.gw > 0 && g.gh > 0 && g.batch >= 0 && g.batch < N;
  const bool mergeable = g.gw <= kPrMaxG && g.gh <= kPrMaxG;
  int pos = 0;
  if (lane == 0) {
    const int e = (valid ? g.batch : 0) * kPrNB + pr_bucket(g, valid);
    pos = plan.starts[e] + atomicAdd(&plan.cursor[e], 1);
  }
  pos = __shfl_sync(0xffffffffu, pos, 0);
  PrRecA* blkA = plan.blocksA + (size_t)pos * kPrBlockSlots;
  PrRecB* blkB = plan.blocksB + (size_t)pos * kPrBlockSlots;
  const int axis = lane >= P ? 1 : 0;  // 0: x bins (lanes 0..P-1), 1: y bins (lanes P..2P-1)
  const int p = lane - axis * P;
  int n = 0, rstart_l = 0;
  bool overflow = false;
  if (lane < 2 * P) {
    int rstart = 0, rn = 0;
    float rw[kPrNW];
#pragma unroll
    for (int k = 0; k < kPrNW; ++k) rw[k] = 0.f;
    if (valid && mergeable) {
      const float start = axis ? g.sh : g.sw, bin = axis ? g.bh : g.bw;
      const int gs = axis ? g.gh : g.gw, L = axis ? H : W;
      const float sc = axis ? 1.f : g.inv_count;
      int first = 0x7fffffff, last = -1;
      for (int i = 0; i < gs; ++i) {
        const Tap t = make_tap(start, bin, p, i, gs, L, 0);
        if (t.wl == 0.f && t.wh == 0.f) continue;
        first = min(first, t.lo);
        last = max(last, t.hi);
      }
      if (last >= 0) {
        n = last - first + 1;
        if (n > kPrNW) {
          overflow = true;
        } else {
          for (int i = 0; i < gs; ++i) {
            const Tap t = make_tap(start, bin, p, i, gs, L, 0);
            if (t.wl == 0.f && t.wh == 0.f) continue;
            rw[t.lo - first] += t.wl * sc;
            rw[t.hi - first] += t.wh * sc;
          }
          rstart = first;
          rstart_l = first;
          rn = n;
        }
      }
    }
    PrRecA ra;
    PrRecB rb;
    ra.start = rstart;
    ra.n = rn;
#pragma unroll
    for (int k = 0; k < 6; ++k) ra.w[k] = rw[k];
#pragma unroll
    for (int k = 0; k < 8; ++k) rb.w[k] = rw[6 + k];
    blkA[1 + axis * 14 + p] = ra;  // x bins in slots 1.., y bins in slots 15.. (also for P = 7)
    blkB[1 + axis * 14 + p] = rb;
  }
  const unsigned ovf = __ballot_sync(0xffffffffu, overflow);
  int nx = (lane < P) ? n : 0, ny = (lane >= P && lane < 2 * P) ? n : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nx = max(nx, __shfl_xor_sync(0xffffffffu, nx, o));
    ny = max(ny, __shfl_xor_sync(0xffffffffu, ny, o));
  }
  // backward facts of the x bands: footprint [x0, x0 + fw) and the largest number of bins that touch one column
  const bool xl = lane < P && n > 0 && !overflow;
  const int my_s = xl ? rstart_l : 0x7fffffff, my_e = xl ? rstart_l + n : 0;
  int x0 = my_s, x1 = my_e, ni = 0;
  for (int j = 0; j < min(nx, kPrNW); ++j) {  // (warp-uniform trip count) how many bins cover column rstart_l + j
    int c = 0;
    for (int q = 0; q < P; ++q) {
      const int s2 = __shfl_sync(0xffffffffu, my_s, q), e2 = __shfl_sync(0xffffffffu, my_e, q);
      c += (rstart_l + j >= s2 && rstart_l + j < e2) ? 1 : 0;
    }
    if (xl && j < n) ni = max(ni, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    ni = max(ni, __shfl_xor_sync(0xffffffffu, ni, o));
  }
  if (lane == 31) {
    PrHdr h;
    h.roi = r;
    h.nxmax = nx;
    h.nymax = ny;
    h.cls = !valid ? PR_ZERO : (!mergeable || ovf) ? PR_DIRECT : PR_BAND;
    h.x0 = x1 > 0 ? x0 : 0;
    h.fw = x1 > 0 ? x1 - x0 : 0;
    h.nimax = ni;
    h.pad = 0;
    *reinterpret_cast<PrHdr*>(blkA) = h;
  }
  // ---- inverse x table for the backward ------------------------------------------------------------------------
  if (!plan.inv || !valid || !mergeable || ovf || x1 <= 0) return;
  const int fw = x1 - x0;
  int fmt = 0, logl = 0, nps = 0;
  if (fw <= 16) {
    logl = fw <= 2 ? 3 : fw <= 4 ? 2 : fw <= 8 ? 1 : 0;
    nps = (ni + (1 << logl) - 1) >> logl;
    fmt = (nps >= 1 && nps <= 7) ? 1 : 0;
  } else if (fw <= 128 && ni <= 3) {
    fmt = 2;
    nps = ni;
  }
  if (!fmt) return;
  __syncwarp();  // the x records written above are read back by other lanes
  float* inv = plan.inv + (size_t)pos * kPrInvFloats;
  const int nv = fmt == 1 ? 16 : fw, esz = fmt == 1 ? 8 : 4, wmax = esz - 1;
  for (int v = lane; v < nv; v += 32) {
    const int col = v >> logl, part = v & ((1 << logl) - 1);
    float w[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) w[q] = 0.f;
    int pa = -1, seen = 0;
    if (col < fw) {
      const int x = x0 + col;
      for (int pw = 0; pw < P; ++pw) {
        const int xs = blkA[1 + pw].start, nn = blkA[1 + pw].n;
        const int jx = x - xs;
        if (jx >= 0 && jx < nn) {
          if (seen >= part * nps && seen < (part + 1) * nps) {
            if (pa < 0) pa = min(pw, P - nps);
            const float wv = jx < 6 ? blkA[1 + pw].w[jx] : blkB[1 + pw].w[jx - 6];
            const int q = pw - pa;
#pragma unroll
            for (int qq = 0; qq < 7; ++qq)
              if (qq == q) w[qq] = wv;
          }
          ++seen;
        }
      }
    }
    float* e = inv + v * esz;
    e[0] = __int_as_float(pa < 0 ? 0 : pa);
    for (int q = 0; q < wmax; ++q) e[1 + q] = w[q];
  }
  if (lane == 31) reinterpret_cast<PrHdr*>(blkA)->pad = fmt | (logl << 8) | (nps << 16);
}

// ------------------------------------------------------------------------------------------------
// shared pieces of the main kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Copy the planes [cbase, cbase + nch) of image `img` into shared memory: plane pl at pl * PS floats, rows of RS
// floats; column W duplicates column W-1 and row H duplicates row H-1 (the clamped high taps of the reference then
// are ordinary "+1" neighbours), everything else of the slice is zero.  4-byte cp.async (LDGSTS): no registers, every
// element of the slice in flight at once (the NCHW rows of a 63-wide map are not 16-byte aligned, so neither wider
// cp.async nor TMA applies); the caller commits and waits.
__device__ __forceinline__ void pr_load_planes(float* __restrict__ planes, const float* __restrict__ src /* plane 0 */,
                                               int nch, int CH, int H, int W, int RS, int PS) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int rows = CH * (H + 1);
  for (int q = warp; q < rows; q += nwarps) {
    const int pl = q / (H + 1), y = q - pl * (H + 1);
    float* drow = planes + (size_t)pl * PS + y * RS;
    if (pl < nch) {
      const float* srow = src + (size_t)pl * H * W + (size_t)min(y, H - 1) * W;
      for (int x = lane; x <= W; x += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(pr_smem_u32(drow + x)), "l"(srow + min(x, W - 1))
                     : "memory");
    } else {
      for (int x = lane; x <= W; x += 32) drow[x] = 0.f;
    }
  }
}

// The warp's next RoI block (A half): 58 16-byte pieces straight into the warp's record buffer (cp.async, no
// registers).
__device__ __forceinline__ void pr_issue_block(float* wbuf, const PrRecA* __restrict__ blk, int lane) {
  const uint32_t d = pr_smem_u32(wbuf) + lane * 16;
  const char* s = reinterpret_cast<const char*>(blk) + lane * 16;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(s) : "memory");
  if (32 + lane < kPrBlockVec)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512), "l"(s + 512) : "memory");
}
__device__ __forceinline__ void pr_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pr_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// Lane layout: PWP = 16 (P = 14) or 8 (P = 7) lanes per channel slot, SLOTS = 32 / PWP slots.  Pass k of slot s works
// on channel cbase + chan(k, s) = SLOTS*D*(k / D) + D*s + k % D with D = min(4, CPL): the slots of one instruction
// are D channels apart, which puts them 16 (8) banks apart both in the planes (D*PS = 32/SLOTS mod 32) and in the
// [channel][196] staging tile of the 14x14 pooler (4 * 196 = 16 mod 32).
template <int P, int CPL>
struct PrMap {
  static constexpr int PWP = P > 8 ? 16 : 8, SLOTS = 32 / PWP, D = CPL < 4 ? CPL : 4, CH = SLOTS * CPL, PER = P * P;
  static constexpr int STG = (CH * PER + 31) / 32 * 32;  // floats of one warp's staging tile (128-byte multiple)
  __host__ __device__ static constexpr int koff(int k) { return SLOTS * D * (k / D) + k % D; }  // + D * slot
};

// Output sink.  Measured on the [R][C][14][14] tensor (tools/ubench/store_patterns.cu): fragments of 56 / 112 bytes
// per store instruction cap at 2.5 / 4.1 TB/s whether they are written by STG or by TMA tensor stores, whereas runs
// of >= 784 contiguous bytes reach 7 TB/s.  So a warp stages the WHOLE [CH][P*P] tile of its RoI -- CH consecutive
// channels = one contiguous, 16-byte aligned run of the pooled tensor (6272 bytes = 49 full lines for 8 channels of
// a 14x14 pooler) -- in its own shared memory and sends it off with ONE 1-D bulk copy (cp.async.bulk, SASS UBLKCP).
// Single-buffered: the copy engine has long read the tile when the next RoI's first row is ready.  The idle lanes
// of a slot (bins 14, 15) mirror the slot's last bin and store the same value to the same address: no predicates.
template <int P, int CPL>
struct PrTileSink {
  using M = PrMap<P, CPL>;
  float* tp;        // this lane's (slot, bin) position in row 0 of the staging tile
  float* op;        // ... in the current row
  float* stage;
  float* out;
  float* gdst;
  int C, cbase, nch, lane;
  bool pending, off;
  __device__ __forceinline__ void init(float* stage_, float* out_, int C_, int cbase_, int nch_, int lane_, int slot,
                                       int pe, int dbg) {
    stage = stage_;
    tp = stage_ + (M::D * slot) * M::PER + pe;
    out = out_;
    C = C_;
    cbase = cbase_;
    nch = nch_;
    lane = lane_;
    pending = false;
    off = dbg & 1;
  }
  __device__ __forceinline__ void start(int roi) {
    op = tp;
    gdst = out + ((size_t)roi * C + cbase) * M::PER;
  }
  // one output row: v(k) yields the value of channel pass k
  template <class F>
  __device__ __forceinline__ void row(F&& v) {
    if (pending) {  // first row of a RoI: the tile is about to be rewritten, the copy engine must have read it
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      pending = false;
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) op[M::koff(k) * M::PER] = v(k);
    op += P;
  }
  __device__ __forceinline__ void flush() {
    const uint32_t bytes = (uint32_t)(nch * M::PER) * 4u;
    if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(gdst) & 15) == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && !off) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                     "r"(pr_smem_u32(stage)), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending = true;
    } else {  // ragged 7x7 channel group or an unaligned tensor: plain coalesced copy
      __syncwarp();
      if (!off)
        for (int e = lane; e < nch * M::PER; e += 32) gdst[e] = stage[e];
      __syncwarp();
    }
  }
  __device__ __forceinline__ void finish() {
    if (pending) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      pending = false;
    }
  }
};

// Horizontal blend of one map row for this lane's bin: NXT cells from the lane's band start.
template <int NXT>
__device__ __forceinline__ float pr_hblend(const float* a /* shared */, const float (&wx)[NXT]) {
  float h = wx[0] * a[0];
#pragma unroll
  for (int j = 1; j < NXT; ++j) h = fmaf(wx[j], a[j], h);
  return h;
}

// The same blend for two channel slots at once with packed fp32x2 arithmetic (FMUL2 / FFMA2): per element the very
// operations of pr_hblend, so the results are bit-identical; the weight pairs (w, w) are built once per RoI.
template <int NXT>
__device__ __forceinline__ float2 pr_hblend2(const float* a0 /* shared */, const float* a1, const float2 (&wx2)[NXT]) {
  float2 h = __fmul2_rn(wx2[0], make_float2(a0[0], a1[0]));
#pragma unroll
  for (int j = 1; j < NXT; ++j) h = __ffma2_rn(wx2[j], make_float2(a0[j], a1[j]), h);
  return h;
}
// all CPL slots of one map row
template <int P, int CPL, int NXT>
__device__ __forceinline__ void pr_hblend_all(float (&h)[CPL], const float* ra, int PS, const float (&wx)[NXT],
                                              const float2 (&wx2)[NXT]) {
  using M = PrMap<P, CPL>;
  if constexpr (CPL % 2 == 0 && NXT <= 6) {
#pragma unroll
    for (int k = 0; k < CPL; k += 2) {
      const float2 v = pr_hblend2<NXT>(ra + M::koff(k) * PS, ra + M::koff(k + 1) * PS, wx2);
      h[k] = v.x;
      h[k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < CPL; ++k) h[k] = pr_hblend<NXT>(ra + M::koff(k) * PS, wx);
  }
}
template <int NXT>
__device__ __forceinline__ void pr_pair_wx(float2 (&wx2)[NXT], const float (&wx)[NXT]) {
#pragma unroll
  for (int j = 0; j < NXT; ++j) wx2[j] = make_float2(wx[j], wx[j]);
}

// x weights of this lane's bin: the first six from the prefetched record, wider bands from the B half in global memory
template <int NXT>
__device__ __forceinline__ void pr_load_wx(float (&wx)[NXT], const float* xrec /* shared */, const PrRecB* xrecB) {
#pragma unroll
  for (int j = 0; j < NXT; ++j) wx[j] = j < 6 ? xrec[2 + j] : __ldg(&xrecB->w[j < 6 ? 0 : j - 6]);
}

// Two rows per band (one sample per bin vertically, three quarters of the RoIs): the two blended rows live in
// registers A / B whose roles (top / bottom) swap on every advance -- no register moves; every map row of the
// footprint is blended exactly once per RoI and channel.
template <int P, int CPL, int NXT, class Sink>
__device__ __forceinline__ void pr_fwd_win2(const float* a0, const float (&wx)[NXT], const float* yrec, int RS, int PS,
                                            Sink& sk) {
  using M = PrMap<P, CPL>;
  float hA[CPL], hB[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) hA[k] = hB[k] = 0.f;
  float2 wx2[NXT];
  pr_pair_wx<NXT>(wx2, wx);
  int cy = -0x40000000;
  bool flip = false;  // false: A = row cy, B = row cy + 1
  int4 yn = *reinterpret_cast<const int4*>(yrec);  // warp-uniform; the next record is fetched a row ahead
#pragma unroll 1
  for (int ph = 0; ph < P; ++ph) {
    const int4 yr = yn;
    yn = *reinterpret_cast<const int4*>(yrec + min(ph + 1, P - 1) * 8);
    const int ys = yr.x;
    if (yr.y > 0 && ys != cy) {
      const float* ra = a0 + (ys + 1) * RS;
      if (ys == cy + 1) {  // slide: the new bottom row replaces the old top row
        if (!flip) pr_hblend_all<P, CPL, NXT>(hA, ra, PS, wx, wx2);
        else pr_hblend_all<P, CPL, NXT>(hB, ra, PS, wx, wx2);
        flip = !flip;
      } else {
        pr_hblend_all<P, CPL, NXT>(hA, ra - RS, PS, wx, wx2);
        pr_hblend_all<P, CPL, NXT>(hB, ra, PS, wx, wx2);
        flip = false;
      }
      cy = ys;
    }
    const float w0 = __int_as_float(yr.z), w1 = __int_as_float(yr.w);
    const float wa = flip ? w1 : w0, wb = flip ? w0 : w1;
    if constexpr (CPL % 2 == 0) {  // vertical blend of two channel slots per instruction (same operations per element)
      float o[CPL];
      const float2 wa2 = make_float2(wa, wa), wb2 = make_float2(wb, wb);
#pragma unroll
      for (int k = 0; k < CPL; k += 2) {
        const float2 t = __ffma2_rn(wb2, make_float2(hB[k], hB[k + 1]),
                                    __fmul2_rn(wa2, make_float2(hA[k], hA[k + 1])));
        o[k] = t.x;
        o[k + 1] = t.y;
      }
      sk.row([&](int k) { return o[k]; });
    } else {
      sk.row([&](int k) { return fmaf(wb, hB[k], wa * hA[k]); });
    }
  }
}

// Rolling window of WIN blended rows (rows cy .. cy+WIN-1) for taller bands (WIN <= 6: weights all in the A half).
template <int P, int CPL, int WIN, int NXT, class Sink>
__device__ __forceinline__ void pr_fwd_win(const float* a0, const float (&wx)[NXT], const float* yrec, int H, int RS,
                                           int PS, Sink& sk) {
  using M = PrMap<P, CPL>;
  float hw[WIN][CPL];
#pragma unroll
  for (int r = 0; r < WIN; ++r)
#pragma unroll
    for (int k = 0; k < CPL; ++k) hw[r][k] = 0.f;
  float2 wx2[NXT];
  pr_pair_wx<NXT>(wx2, wx);
  int cy = -0x40000000;
  int4 na = *reinterpret_cast<const int4*>(yrec);
  float4 nb = *reinterpret_cast<const float4*>(yrec + 4);
#pragma unroll 1
  for (int ph = 0; ph < P; ++ph) {
    const int4 ca = na;
    const float4 cb = nb;
    na = *reinterpret_cast<const int4*>(yrec + min(ph + 1, P - 1) * 8);
    nb = *reinterpret_cast<const float4*>(yrec + min(ph + 1, P - 1) * 8 + 4);
    const int ys = ca.x;
    if (ca.y > 0 && ys != cy) {
      int d = ys - cy;
      if (d < 0 || d > WIN) {  // (re)start: fill the whole window
        cy = ys - WIN;
        d = WIN;
      }
      for (int s = 0; s < d; ++s) {
        const float* ra = a0 + min(cy + WIN + s, H) * RS;
#pragma unroll
        for (int k = 0; k < CPL; ++k)
#pragma unroll
          for (int r = 0; r + 1 < WIN; ++r) hw[r][k] = hw[r + 1][k];
        pr_hblend_all<P, CPL, NXT>(hw[WIN - 1], ra, PS, wx, wx2);
      }
      cy = ys;
    }
    float wy[WIN];
    wy[0] = __int_as_float(ca.z);
    wy[1] = __int_as_float(ca.w);
    if (WIN > 2) wy[2] = cb.x;
    if (WIN > 3) wy[WIN > 3 ? 3 : 0] = cb.y;
    if (WIN > 4) wy[WIN > 4 ? 4 : 0] = cb.z;
    if (WIN > 5) wy[WIN > 5 ? 5 : 0] = cb.w;
    if constexpr (CPL % 2 == 0) {
      float o[CPL];
#pragma unroll
      for (int k = 0; k < CPL; k += 2) {
        float2 t = __fmul2_rn(make_float2(wy[0], wy[0]), make_float2(hw[0][k], hw[0][k + 1]));
#pragma unroll
        for (int r = 1; r < WIN; ++r)
          t = __ffma2_rn(make_float2(wy[r], wy[r]), make_float2(hw[r][k], hw[r][k + 1]), t);
        o[k] = t.x;
        o[k + 1] = t.y;
      }
      sk.row([&](int k) { return o[k]; });
    } else {
      sk.row([&](int k) {
        float o = wy[0] * hw[0][k];
#pragma unroll
        for (int r = 1; r < WIN; ++r) o = fmaf(wy[r], hw[r][k], o);
        return o;
      });
    }
  }
}

// No vertical reuse (bands taller than the window templates): every output row blends its own rows.
template <int P, int CPL, int NXT, class Sink>
__device__ __forceinline__ void pr_fwd_gen(const float* a0, const float (&wx)[NXT], const float* yrec,
                                           const PrRecB* yrecB, int RS, int PS, Sink& sk) {
  float2 wx2[NXT];
  pr_pair_wx<NXT>(wx2, wx);
#pragma unroll 1
  for (int ph = 0; ph < P; ++ph) {
    const float* yr = yrec + ph * 8;
    const int ys = __float_as_int(yr[0]), ny = __float_as_int(yr[1]);
    float acc[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
    const float* ra = a0 + ys * RS;
#pragma unroll 1
    for (int r = 0; r < ny; ++r, ra += RS) {
      const float wy = r < 6 ? yr[2 + r] : __ldg(&yrecB[ph].w[r - 6]);
      float h[CPL];
      pr_hblend_all<P, CPL, NXT>(h, ra, PS, wx, wx2);
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[k] = fmaf(wy, h[k], acc[k]);
    }
    sk.row([&](int k) { return acc[k]; });
  }
}

template <int P, int CPL, int NXT, class Sink>
__device__ __forceinline__ void pr_fwd_band(const float* a0, const float* xrec, const PrRecB* xrecB, const float* yrec,
                                            const PrRecB* yrecB, int ny, int H, int RS, int PS, Sink& sk) {
  float wx[NXT];
  pr_load_wx<NXT>(wx, xrec, xrecB);
  if (ny <= 2) pr_fwd_win2<P, CPL, NXT>(a0, wx, yrec, RS, PS, sk);
  else if (ny <= 4) pr_fwd_win<P, CPL, 4, NXT>(a0, wx, yrec, H, RS, PS, sk);
  else pr_fwd_gen<P, CPL, NXT>(a0, wx, yrec, yrecB, RS, PS, sk);
}

template <int P, int CPL>
__global__ void __launch_bounds__(kPrThreads, 1)
roi_align_fwd_pr_kernel(const float* __restrict__ in, const float* __restrict__ rois, float* __restrict__ out,
                        PrPlan plan, int N, int C, int H, int W, int RS, int PS, int ngroups, float scale,
                        int sampling_ratio, int aligned, int dbg) {
  using M = PrMap<P, CPL>;
  constexpr int CH = M::CH, D = M::D;
  extern __shared__ __align__(128) float smem[];
  float* wrec = smem;                                                    // [warps][2][232]: RoI blocks (A halves)
  float* stage = smem + kPrWarps * 2 * kPrBlockFloats;                   // [warps][CH][P*P]: output tiles
  float* planes = stage + kPrWarps * M::STG;                             // [CH][PS] + 32 floats of zero tail
  __shared__ int s_unit, s_next;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / M::PWP;
  const int p = min(lane - slot * M::PWP, P - 1);  // idle lanes mirror the slot's last bin
  float* wbuf = wrec + warp * (2 * kPrBlockFloats);
  // pads between / after the planes: written once, never touched by the loaders
  for (int e = threadIdx.x; e < CH * PS + 32; e += blockDim.x) planes[e] = 0.f;
  const int nunits = plan.counter[1] * ngroups;
  PrTileSink<P, CPL> sk;
  for (;;) {
    __syncthreads();  // previous unit fully consumed (planes, s_unit, s_next)
    if (threadIdx.x == 0) {
      s_unit = atomicAdd(&plan.counter[0], 1);
      s_next = 0;
    }
    __syncthreads();
    const int u = s_unit;
    if (u >= nunits) break;
    const int j = u / ngroups, cg = u - j * ngroups;
    const PrChunk ck = plan.chunks[j];
    const int cbase = cg * CH, nch = min(CH, C - cbase), nroi = ck.end - ck.begin;
    // this warp's first RoI block flies while the planes are loaded
    int i = 0;
    if (lane == 0) i = atomicAdd(&s_next, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i < nroi) pr_issue_block(wbuf, plan.blocksA + (size_t)(ck.begin + i) * kPrBlockSlots, lane);
    pr_commit();
    pr_load_planes(planes, in + ((size_t)ck.img * C + cbase) * H * W, nch, CH, H, W, RS, PS);
    pr_commit();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    sk.init(stage + warp * M::STG, out, C, cbase, nch, lane, slot, p, dbg);
    int buf = 0;
    while (i < nroi) {
      int inext = 0;
      if (lane == 0) inext = atomicAdd(&s_next, 1);
      inext = __shfl_sync(0xffffffffu, inext, 0);
      if (inext < nroi)
        pr_issue_block(wbuf + (buf ^ 1) * kPrBlockFloats, plan.blocksA + (size_t)(ck.begin + inext) * kPrBlockSlots,
                       lane);
      pr_commit();
      pr_wait_prev();
      __syncwarp();
      const float* blk = wbuf + buf * kPrBlockFloats;
      const PrHdr hd = *reinterpret_cast<const PrHdr*>(blk);
      const float* xrec = blk + (1 + p) * 8;
      const float* yrec = blk + 15 * 8;
      const PrRecB* blkB = plan.blocksB + (size_t)(ck.begin + i) * kPrBlockSlots;
      sk.start(hd.roi);
      const bool small = hd.nxmax <= 2 && hd.nymax <= 2;
      if (hd.cls == PR_BAND && !((dbg & 2) && !small) && !((dbg & 4) && small)) {
        const int xs = __float_as_int(xrec[0]);
        const float* a0 = planes + (D * slot) * PS + xs;
        const int nx = hd.nxmax, ny = hd.nymax;
#define CDDMSL_PR_NX(NXT) \
  pr_fwd_band<P, CPL, NXT>(a0, xrec, blkB + 1 + p, yrec, blkB + 15, ny, H, RS, PS, sk)
        if (nx <= 2) CDDMSL_PR_NX(2);
        else if (nx == 3) CDDMSL_PR_NX(3);
        else if (nx == 4) CDDMSL_PR_NX(4);
        else if (nx <= 6) CDDMSL_PR_NX(6);
        else CDDMSL_PR_NX(14);
#undef CDDMSL_PR_NX
      } else if (hd.cls == PR_DIRECT) {
        // ---- per-sample taps in the reference's order (huge sampling grids / sparse fixed grids): rare ----------
        const RoiGeom g = roi_geom(rois + (size_t)hd.roi * 5, scale, aligned, P, P, sampling_ratio, H, W);
        const float* a0 = planes + (D * slot) * PS;
#pragma unroll 1
        for (int ph = 0; ph < P; ++ph) {
          float acc[CPL];
#pragma unroll
          for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
          for (int iy = 0; iy < g.gh; ++iy) {
            const Tap ty = make_tap(g.sh, g.bh, ph, iy, g.gh, H, 0);
            if (ty.wl == 0.f && ty.wh == 0.f) continue;
            for (int ix = 0; ix < g.gw; ++ix) {
              const Tap tx = make_tap(g.sw, g.bw, p, ix, g.gw, W, 0);
              if (tx.wl == 0.f && tx.wh == 0.f) continue;
#pragma unroll
              for (int k = 0; k < CPL; ++k) {
                const float* b = a0 + M::koff(k) * PS;
                const float v00 = b[ty.lo * RS + tx.lo], v01 = b[ty.lo * RS + tx.hi];
                const float v10 = b[ty.hi * RS + tx.lo], v11 = b[ty.hi * RS + tx.hi];
                acc[k] += ty.wl * tx.wl * v00 + ty.wl * tx.wh * v01 + ty.wh * tx.wl * v10 + ty.wh * tx.wh * v11;
              }
            }
          }
          sk.row([&](int k) { return acc[k] * g.inv_count; });
        }
      } else {  // PR_ZERO (empty / inverted box, batch index outside the batch) or a class skipped by roi_pr_dbg
#pragma unroll 1
        for (int ph = 0; ph < P; ++ph) sk.row([&](int) { return 0.f; });
      }
      sk.flush();
      __syncwarp();  // every lane is done with this record buffer before it is refilled
      i = inext;
      buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    sk.finish();
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// Transpose of the forward.  Global traffic: the compulsory stream of the pooled gradient (one 1-D bulk load of the
// RoI's contiguous [CH][196] tile per visit, mbarrier completion, fetched one RoI ahead) and ONE coalesced
// red.global.add per (footprint row, channel) into the zeroed NCHW gradient map: fh*fw reductions per RoI and
// channel in 64-byte runs on an L2-resident image (the round-1 kernel issues ~(fh+14)*fw and moves 12 GB through the
// L2 atomic units).  Per RoI a warp
//   1. accumulates U = Ay^T G in registers with the lanes on the BINS (the y bands are warp-uniform): a window of WIN
//      footprint rows per channel pass; a row leaves the window complete;
//   2. transposes the completed row: the 4 channel passes of a bin go to a 16-entry scratch line as one float4, and
//      lane (column x) sums w * u4 over the contiguous window of bins whose band covers x.  The windows come from the
//      INVERSE x table the planning kernel built (pr_plan_fill_kernel) and the warp prefetched with cp.async; narrow
//      footprints (most RoIs are a few cells wide, so a column is shared by up to 14 bins) use 2-8 lanes per column
//      and a shuffle reduction, wide ones take several 16-column passes.
constexpr int kPrBwdWarps = 16;
constexpr int kPrBwdThreads = kPrBwdWarps * 32;
constexpr int kPrScratch = 2 * 16 * 4;            // [SLOTS][16 bins][4 passes]

__device__ __forceinline__ void pr_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pr_smem_u32(bar)));
}
__device__ __forceinline__ void pr_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pr_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   pr_smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(pr_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void pr_mbar_wait(uint64_t* bar, uint32_t parity) {  // bounded: a protocol bug must trap
  const uint32_t a = pr_smem_u32(bar);
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
__device__ __forceinline__ void pr_red_global(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Per-warp state of the transposed (horizontal) pass of one visit.
struct PrBwdFlush {
  float* sline;        // this slot's scratch line [16 bins][4 passes]
  const float* inv;    // the RoI's inverse table (shared memory)
  char* grow;          // gin + slot's first channel plane + row cy + x0 + this lane's column (bytes; advanced per row)
  int fmt, logl, nps, npass, fw;
  unsigned hw4, w4;    // bytes between channel planes / rows of gin
  int xi, p;
  bool red0;           // fmt 1: this lane is the first lane of a column inside the footprint
  // fmt 1: the lane's window of bins, fixed for the visit
  const float* sc1;    // sline + first_bin * 4
  float w1[7];
};

template <int CPL>
__device__ __forceinline__ void pr_red4(char* g, unsigned hw4, const float (&t)[CPL], const bool (&act)[CPL]) {
  using M = PrMap<14, CPL>;
#pragma unroll
  for (int k = 0; k < CPL; ++k)
    if (act[k]) pr_red_global(reinterpret_cast<float*>(g + (size_t)M::koff(k) * hw4), t[k]);
}

// Row cy of U is complete: gin[cy][x0 + col] += sum over the column's bins; then the row pointer moves on.
template <int CPL>
__device__ __forceinline__ void pr_bwd_flush(PrBwdFlush& f, bool in_map, const float (&u)[CPL],
                                             const bool (&act)[CPL]) {
  if (in_map) {  // rows beyond the map only ever collect zero weights
    if (CPL == 4) {
      *reinterpret_cast<float4*>(f.sline + f.p * 4) = make_float4(u[0], u[1 % CPL], u[2 % CPL], u[3 % CPL]);
    } else {
#pragma unroll
      for (int k = 0; k < CPL; ++k) f.sline[f.p * 4 + k] = u[k];
    }
    __syncwarp();
    if (f.fmt == 1) {
      float t[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) t[k] = 0.f;
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        if (i < f.nps) {  // warp-uniform
          if (CPL == 4) {
            const float4 q = *reinterpret_cast<const float4*>(f.sc1 + 4 * i);
            t[0] = fmaf(f.w1[i], q.x, t[0]);
            t[1 % CPL] = fmaf(f.w1[i], q.y, t[1 % CPL]);
            t[2 % CPL] = fmaf(f.w1[i], q.z, t[2 % CPL]);
            t[3 % CPL] = fmaf(f.w1[i], q.w, t[3 % CPL]);
          } else {
#pragma unroll
            for (int k = 0; k < CPL; ++k) t[k] = fmaf(f.w1[i], f.sc1[4 * i + k], t[k]);
          }
        }
      }
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (s < f.logl) {  // the lanes that share a column
#pragma unroll
          for (int k = 0; k < CPL; ++k) t[k] += __shfl_xor_sync(0xffffffffu, t[k], 1 << s);
        }
      }
      if (f.red0) pr_red4<CPL>(f.grow, f.hw4, t, act);
    } else {
      for (int ps = 0; ps < f.npass; ++ps) {
        const int v = f.xi + 16 * ps;
        const float4 e = *reinterpret_cast<const float4*>(f.inv + 4 * min(v, f.fw - 1));
        const float* sc = f.sline + __float_as_int(e.x) * 4;
        const float we[3] = {e.y, e.z, e.w};
        float t[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) t[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i < f.nps) {
            if (CPL == 4) {
              const float4 q = *reinterpret_cast<const float4*>(sc + 4 * i);
              t[0] = fmaf(we[i], q.x, t[0]);
              t[1 % CPL] = fmaf(we[i], q.y, t[1 % CPL]);
              t[2 % CPL] = fmaf(we[i], q.z, t[2 % CPL]);
              t[3 % CPL] = fmaf(we[i], q.w, t[3 % CPL]);
            } else {
#pragma unroll
              for (int k = 0; k < CPL; ++k) t[k] = fmaf(we[i], sc[4 * i + k], t[k]);
            }
          }
        }
        if (v < f.fw) pr_red4<CPL>(f.grow + 64 * ps, f.hw4, t, act);
      }
    }
    __syncwarp();  // scratch is rewritten by the next flush
  }
  f.grow += f.w4;
}

// Vertical pass: window of WIN footprint rows (rows cy .. cy+WIN-1) of U per channel pass; WIN >= the tallest band.
template <int CPL, int WIN>
__device__ __forceinline__ void pr_bwd_win(PrBwdFlush& f, const float* gp /* tile + slot/bin */, const float* yrec,
                                           int H, const bool (&act)[CPL]) {
  using M = PrMap<14, CPL>;
  float acc[WIN][CPL];
#pragma unroll
  for (int r = 0; r < WIN; ++r)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[r][k] = 0.f;
  int cy = 0;
  bool open = false;
#pragma unroll 1
  for (int ph = 0; ph < 14; ++ph, gp += 14) {
    const int4 ca = *reinterpret_cast<const int4*>(yrec + ph * 8);
    if (ca.y <= 0) continue;
    float g[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) g[k] = gp[M::koff(k) * 196];
    if (!open) {  // first band: the row pointer starts at its first row (may be negative only in theory: bands
      cy = ca.x;  // hold cells inside the map)
      f.grow += (size_t)cy * f.w4;
      open = true;
    }
    while (cy < ca.x) {  // row cy is complete (a gap wider than the window just emits zero rows: never happens with
      pr_bwd_flush<CPL>(f, cy < H, acc[0], act);  // adaptive sampling)
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
#pragma unroll
        for (int r = 0; r + 1 < WIN; ++r) acc[r][k] = acc[r + 1][k];
        acc[WIN - 1][k] = 0.f;
      }
      ++cy;
    }
    float wy[WIN];
    wy[0] = __int_as_float(ca.z);
    wy[1] = __int_as_float(ca.w);
    if (WIN > 2) {
      const float4 cb = *reinterpret_cast<const float4*>(yrec + ph * 8 + 4);
      wy[2 % WIN] = cb.x;
      wy[3 % WIN] = cb.y;
    }
#pragma unroll
    for (int r = 0; r < WIN; ++r)
#pragma unroll
      for (int k = 0; k < CPL; ++k) acc[r][k] = fmaf(wy[r], g[k], acc[r][k]);
  }
  if (open) {
#pragma unroll
    for (int r = 0; r < WIN; ++r) pr_bwd_flush<CPL>(f, cy + r < H, acc[r], act);
  }
}

template <int CPL>
__global__ void __launch_bounds__(kPrBwdThreads, 1)
roi_align_bwd_pr_kernel(const float* __restrict__ gout, const float* __restrict__ rois, float* __restrict__ gin,
                        PrPlan plan, int N, int C, int H, int W, int ngroups, float scale, int sampling_ratio,
                        int aligned, int dbg) {
  constexpr int P = 14;
  using M = PrMap<P, CPL>;
  constexpr int CH = M::CH, D = M::D, PER = 196;
  extern __shared__ __align__(128) float smem[];
  float* wrec = smem;                                                      // [warps][2][232]: RoI blocks (A halves)
  float* tiles = wrec + kPrBwdWarps * 2 * kPrBlockFloats;                  // [warps][CH][196]: pooled-gradient tiles
  float* invs = tiles + kPrBwdWarps * M::STG;                              // [warps][2][kPrInvFloats]
  float* scr = invs + kPrBwdWarps * 2 * kPrInvFloats;                      // [warps][kPrScratch]
  __shared__ int s_unit, s_next;
  __shared__ __align__(8) uint64_t bars[kPrBwdWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane >> 4;
  const int p = min(lane & 15, P - 1);  // idle lanes mirror the slot's last bin
  float* wbuf = wrec + warp * (2 * kPrBlockFloats);
  float* ibuf = invs + warp * (2 * kPrInvFloats);
  float* tile = tiles + warp * M::STG;
  uint64_t* bar = &bars[warp];
  if (lane == 0) pr_mbar_init(bar);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase = 0;
  const int HW = H * W;
  PrBwdFlush f;
  f.sline = scr + warp * kPrScratch + slot * 64;
  f.hw4 = (unsigned)HW * 4u;
  f.w4 = (unsigned)W * 4u;
  f.xi = lane & 15;
  f.p = p;
  auto issue_next = [&](int b, int idx, int begin) {  // RoI block + the head of its inverse table, no registers
    pr_issue_block(wbuf + b * kPrBlockFloats, plan.blocksA + (size_t)(begin + idx) * kPrBlockSlots, lane);
    const char* src = reinterpret_cast<const char*>(plan.inv + (size_t)(begin + idx) * kPrInvFloats) + lane * 16;
    const uint32_t dst = pr_smem_u32(ibuf + b * kPrInvFloats) + lane * 16;
#pragma unroll
    for (int q = 0; q < kPrInvPrefetch / 512; ++q)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * 512), "l"(src + q * 512) : "memory");
  };
  const int nunits = plan.counter[1] * ngroups;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      s_unit = atomicAdd(&plan.counter[2], 1);
      s_next = 0;
    }
    __syncthreads();
    const int u = s_unit;
    if (u >= nunits) break;
    const int j = u / ngroups, cg = u - j * ngroups;
    const PrChunk ck = plan.chunks[j];
    const int cbase = cg * CH, nch = min(CH, C - cbase), nroi = ck.end - ck.begin;
    const uint32_t tile_bytes = (uint32_t)nch * PER * 4u;
    char* gimg = reinterpret_cast<char*>(gin + ((size_t)ck.img * C + cbase + D * slot) * HW);
    bool act[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) act[k] = (M::koff(k) + D * slot) < nch && !(dbg & 1);
    int i = 0;
    if (lane == 0) i = atomicAdd(&s_next, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i < nroi) issue_next(0, i, ck.begin);
    pr_commit();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (i < nroi && lane == 0) {
      const int roi = reinterpret_cast<const PrHdr*>(wbuf)->roi;
      pr_bulk_load(tile, gout + ((size_t)roi * C + cbase) * PER, tile_bytes, bar);
    }
    int buf = 0;
    while (i < nroi) {
      int inext = 0;
      if (lane == 0) inext = atomicAdd(&s_next, 1);
      inext = __shfl_sync(0xffffffffu, inext, 0);
      if (inext < nroi) issue_next(buf ^ 1, inext, ck.begin);
      pr_commit();
      const float* blk = wbuf + buf * kPrBlockFloats;
      const PrHdr hd = *reinterpret_cast<const PrHdr*>(blk);
      const float* yrec = blk + 15 * 8;
      float* inv = ibuf + buf * kPrInvFloats;
      const int fmt = hd.pad & 0xff;
      const bool small = hd.nxmax <= 2 && hd.nymax <= 2;
      const bool band = hd.cls == PR_BAND && fmt != 0 && hd.nymax <= 4 && !((dbg & 2) && !small) &&
                        !((dbg & 4) && small);
      if (band) {
        f.fmt = fmt;
        f.logl = (hd.pad >> 8) & 0xff;
        f.nps = (hd.pad >> 16) & 0xff;
        f.fw = hd.fw;
        f.inv = inv;
        f.npass = (hd.fw + 15) >> 4;
        if (fmt == 2 && hd.fw * 16 > kPrInvPrefetch) {  // very wide footprint: the tail of its table (exposed)
          const float4* src = reinterpret_cast<const float4*>(plan.inv + (size_t)(ck.begin + i) * kPrInvFloats);
          for (int q = kPrInvPrefetch / 16 + lane; q < hd.fw; q += 32) reinterpret_cast<float4*>(inv)[q] = __ldg(src + q);
          __syncwarp();
        }
        const int col = fmt == 1 ? (f.xi >> f.logl) : f.xi;
        f.grow = gimg + (size_t)(hd.x0 + col) * 4;
        if (fmt == 1) {
          const float* e = inv + f.xi * 8;
          f.sc1 = f.sline + __float_as_int(e[0]) * 4;
#pragma unroll
          for (int q = 0; q < 7; ++q) f.w1[q] = e[1 + q];
          f.red0 = (f.xi & ((1 << f.logl) - 1)) == 0 && col < hd.fw;
        }
      }
      pr_mbar_wait(bar, phase);  // this RoI's gradient tile has landed
      phase ^= 1u;
      __syncwarp();
      const float* gp = tile + (D * slot) * PER + p;
      if (band) {
        if (hd.nymax <= 2) pr_bwd_win<CPL, 2>(f, gp, yrec, H, act);
        else pr_bwd_win<CPL, 4>(f, gp, yrec, H, act);
      } else if (hd.cls == PR_DIRECT || (hd.cls == PR_BAND && !(dbg & 6))) {
        // ---- per-sample scatter in the reference's order (no inverse table: sparse / huge sampling grids, bands
        //      taller than four rows, more bins per column than the table formats hold): rare ---------------------
        const RoiGeom g = roi_geom(rois + (size_t)hd.roi * 5, scale, aligned, P, P, sampling_ratio, H, W);
        if ((lane & 15) < P) {
#pragma unroll 1
          for (int ph = 0; ph < P; ++ph) {
            float go[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) go[k] = gp[M::koff(k) * PER + ph * P] * g.inv_count;
            for (int iy = 0; iy < g.gh; ++iy) {
              const Tap ty = make_tap(g.sh, g.bh, ph, iy, g.gh, H, 0);
              if (ty.wl == 0.f && ty.wh == 0.f) continue;
              for (int ix = 0; ix < g.gw; ++ix) {
                const Tap tx = make_tap(g.sw, g.bw, p, ix, g.gw, W, 0);
                if (tx.wl == 0.f && tx.wh == 0.f) continue;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                  if (!act[k]) continue;
                  float* b = reinterpret_cast<float*>(gimg) + (size_t)M::koff(k) * HW;
                  pr_red_global(b + ty.lo * W + tx.lo, go[k] * ty.wl * tx.wl);
                  pr_red_global(b + ty.lo * W + tx.hi, go[k] * ty.wl * tx.wh);
                  pr_red_global(b + ty.hi * W + tx.lo, go[k] * ty.wh * tx.wl);
                  pr_red_global(b + ty.hi * W + tx.hi, go[k] * ty.wh * tx.wh);
                }
              }
            }
          }
        }
      }
      // the tile is consumed: fetch the next RoI's (its block has certainly landed by now)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      if (inext < nroi && lane == 0) {
        const int roi = reinterpret_cast<const PrHdr*>(wbuf + (buf ^ 1) * kPrBlockFloats)->roi;
        pr_bulk_load(tile, gout + ((size_t)roi * C + cbase) * PER, tile_bytes, bar);
      }
      i = inext;
      buf ^= 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int g_roi_pr = 1;         // tuning knob "roi_pr": bit 0 forward (default), bit 1 backward (experimental: 5.4 ms vs 2.4 ms)
int g_roi_pr_chunk = 0;   // tuning knob "roi_pr_chunk": RoIs per unit (0: automatic)
int g_roi_pr_dbg = 0;     // diagnostic knob "roi_pr_dbg": 1 no stores, 2 skip multi-sample RoIs, 4 skip single-sample RoIs
int g_roi_pr_cpl = 0;     // tuning knob "roi_pr_cpl": channel passes per lane (0: as many as fit, <= 8)

static int pr_smem_budget() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0)
      cached = v;
    else
      cached = 227 * 1024;
  }
  return cached - 1024;  // static shared memory + slack
}

struct PrGeom {
  int RS, PS, CPL, CH, smem;
};

static bool pr_geometry(int C, int H, int W, int P, PrGeom* out) {
  if (P != 14 && P != 7) return false;
  const int slots = P == 14 ? 2 : 4;
  const int RS = W + 1;
  const long long fixed = 256 + (long long)kPrWarps * 2 * kPrBlockBytes;  // tail pad + record buffers
  for (int cpl : {4, 2, 1}) {
    if (g_roi_pr_cpl && cpl > g_roi_pr_cpl) continue;
    if (cpl > 1 && slots * (cpl / 2) >= C) continue;  // few channels: do not carry empty planes
    const int d = cpl < 4 ? cpl : 4, ch = slots * cpl;
    // D * PS = 32/slots (mod 32)  <=>  PS = (32/slots)/D (mod 32/D)
    const int mod = 32 / d, want = (32 / slots) / d;
    int PS = (H + 1) * RS;
    PS += ((want - PS % mod) % mod + mod) % mod;
    const long long smem = fixed + (long long)ch * PS * 4 + (long long)kPrWarps * ((ch * P * P + 31) / 32 * 32) * 4;
    if (smem > pr_smem_budget()) continue;
    out->RS = RS;
    out->PS = PS;
    out->CPL = cpl;
    out->CH = ch;
    out->smem = (int)smem;
    return true;
  }
  return false;
}

bool roi_pr_eligible(int N, int C, int H, int W, int R, int P, int bit) {
  PrGeom g;
  if (!(g_roi_pr & bit) || N <= 0 || R <= 0) return false;
  if ((long long)R * C * P * P >= (1LL << 40) || (long long)R >= (1LL << 31) - 1) return false;
  return pr_geometry(C, H, W, P, &g);
}

static int pr_chunk_size(int R, int ngroups) {
  if (g_roi_pr_chunk > 0) return g_roi_pr_chunk < 32 ? 32 : g_roi_pr_chunk;
  long long c = ((long long)R * ngroups + 12LL * sm_count() - 1) / (12LL * sm_count());
  c = (c + 31) / 32 * 32;
  if (c < 64) c = 64;
  if (c > 4096) c = 4096;
  return (int)c;
}

// bytes of planning state for R RoIs over N images (chunk list sized for the smallest chunk, 32)
size_t roi_pr_workspace_bytes(int N, int R) {
  size_t b = 256;                                            // counters
  b += align_up((size_t)(3 * N * kPrNB + 1) * 4, 256);       // counts, starts, cursor
  b += align_up(((size_t)N + (size_t)R / 32 + 1) * sizeof(PrChunk), 256);
  b += 2 * align_up((size_t)R * kPrBlockBytes, 256);         // A and B halves
  b += align_up((size_t)R * kPrInvBytes, 256);               // inverse x tables (backward)
  return b;
}

static PrPlan pr_carve(void* ws, int N, int R) {
  char* b = (char*)ws;
  PrPlan p;
  p.counter = (int*)b;
  b += 256;
  p.counts = (int*)b;
  p.starts = p.counts + N * kPrNB;
  p.cursor = p.starts + N * kPrNB + 1;
  b += align_up((size_t)(3 * N * kPrNB + 1) * 4, 256);
  p.chunks = (PrChunk*)b;
  b += align_up(((size_t)N + (size_t)R / 32 + 1) * sizeof(PrChunk), 256);
  p.blocksA = (PrRecA*)b;
  b += align_up((size_t)R * kPrBlockBytes, 256);
  p.blocksB = (PrRecB*)b;
  b += align_up((size_t)R * kPrBlockBytes, 256);
  p.inv = (float*)b;
  return p;
}

static int pr_build_plan(PrPlan plan, const float* rois, int N, int H, int W, int R, int P, float scale,
                         int sampling_ratio, int aligned, int chunk, cudaStream_t stream, bool want_inv = false) {
  if (!want_inv) plan.inv = nullptr;
  CDDMSL_CUDA(cudaMemsetAsync(plan.counts, 0, (size_t)N * kPrNB * 4, stream));
  if (P == 14) {
    pr_plan_count_kernel<14><<<ceil_div(R, 256), 256, 0, stream>>>(rois, R, N, H, W, scale, sampling_ratio, aligned,
                                                                   plan.counts);
    pr_plan_scan_kernel<<<1, 1024, 0, stream>>>(plan, N, chunk);
    pr_plan_fill_kernel<14><<<ceil_div(R, 8), 256, 0, stream>>>(rois, R, N, H, W, scale, sampling_ratio, aligned, plan);
  } else {
    pr_plan_count_kernel<7><<<ceil_div(R, 256), 256, 0, stream>>>(rois, R, N, H, W, scale, sampling_ratio, aligned,
                                                                  plan.counts);
    pr_plan_scan_kernel<<<1, 1024, 0, stream>>>(plan, N, chunk);
    pr_plan_fill_kernel<7><<<ceil_div(R, 8), 256, 0, stream>>>(rois, R, N, H, W, scale, sampling_ratio, aligned, plan);
  }
  count_launch(3);
  return (int)cudaGetLastError();
}

template <int P, int CPL>
static int pr_launch_fwd(const float* in, const float* rois, float* out, const PrPlan& plan, const PrGeom& g, int N,
                         int C, int H, int W, int ngroups, int grid, float scale, int sampling_ratio, int aligned,
                         cudaStream_t stream) {
  auto k = roi_align_fwd_pr_kernel<P, CPL>;
  CDDMSL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem));
  k<<<grid, kPrThreads, g.smem, stream>>>(in, rois, out, plan, N, C, H, W, g.RS, g.PS, ngroups, scale,
                                          sampling_ratio, aligned, g_roi_pr_dbg);
  count_launch();
  return (int)cudaGetLastError();
}

// in2 / out2 (nullable): a second feature map of the same shape pooled with the SAME RoIs (the source / target pair
// of the region-level consistency branch, clip_roi_heads.py:117-132): the plan -- bands, image buckets, chunk list --
// is built once and serves both launches.
static int pr_fwd_launch(const float* in, const float* rois, float* out, const PrPlan& plan, const PrGeom& g, int N,
                         int C, int H, int W, int P, int ngroups, int grid, float scale, int sampling_ratio,
                         int aligned, cudaStream_t stream);

int roi_align_fwd_pr(const float* in, const float* rois, float* out, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, void* ws, cudaStream_t stream,
                     const float* in2 = nullptr, float* out2 = nullptr) {
  PrGeom g;
  if (!pr_geometry(C, H, W, P, &g)) return CDDMSL_EINVAL;
  const PrPlan plan = pr_carve(ws, N, R);
  const int ngroups = ceil_div(C, g.CH);
  const int chunk = pr_chunk_size(R, ngroups);
  int rc = pr_build_plan(plan, rois, N, H, W, R, P, scale, sampling_ratio, aligned, chunk, stream);
  if (rc) return rc;
  const long long max_units = ((long long)N + R / chunk + 1) * ngroups;
  const int grid = (int)min((long long)sm_count(), max_units);
  rc = pr_fwd_launch(in, rois, out, plan, g, N, C, H, W, P, ngroups, grid, scale, sampling_ratio, aligned, stream);
  if (rc || !in2) return rc;
  CDDMSL_CUDA(cudaMemsetAsync(plan.counter, 0, sizeof(int), stream));  // the unit counter; the plan stays
  return pr_fwd_launch(in2, rois, out2, plan, g, N, C, H, W, P, ngroups, grid, scale, sampling_ratio, aligned, stream);
}

static int pr_fwd_launch(const float* in, const float* rois, float* out, const PrPlan& plan, const PrGeom& g, int N,
                         int C, int H, int W, int P, int ngroups, int grid, float scale, int sampling_ratio,
                         int aligned, cudaStream_t stream) {
#define CDDMSL_PR_FWD(PV, CV)                                                                                  \
  return pr_launch_fwd<PV, CV>(in, rois, out, plan, g, N, C, H, W, ngroups, grid, scale, sampling_ratio, aligned, \
                               stream)
  if (P == 14) {
    if (g.CPL == 4) CDDMSL_PR_FWD(14, 4);
    if (g.CPL == 2) CDDMSL_PR_FWD(14, 2);
    CDDMSL_PR_FWD(14, 1);
  }
  if (g.CPL == 4) CDDMSL_PR_FWD(7, 4);
  if (g.CPL == 2) CDDMSL_PR_FWD(7, 2);
  CDDMSL_PR_FWD(7, 1);
#undef CDDMSL_PR_FWD
}

static bool pr_geometry_bwd(int C, int H, int W, PrGeom* out) {
  (void)H;
  (void)W;
  const long long per_warp = 2LL * kPrBlockBytes + 2LL * kPrInvBytes + kPrScratch * 4LL;
  for (int cpl : {4, 2, 1}) {
    if (g_roi_pr_cpl && cpl > g_roi_pr_cpl) continue;
    if (cpl > 1 && 2 * (cpl / 2) >= C) continue;
    const int ch = 2 * cpl;
    const long long smem = 256 + kPrBwdWarps * (per_warp + ((ch * 196 + 31) / 32 * 32) * 4LL);
    if (smem > pr_smem_budget()) continue;
    out->RS = W;
    out->PS = 0;
    out->CPL = cpl;
    out->CH = ch;
    out->smem = (int)smem;
    return true;
  }
  return false;
}

bool roi_pr_bwd_eligible(int N, int C, int H, int W, int R, int P) {
  PrGeom g;
  if (!(g_roi_pr & 2) || N <= 0 || R <= 0 || P != 14) return false;
  if ((long long)R * C * P * P >= (1LL << 40) || (long long)R >= (1LL << 31) - 1) return false;
  return pr_geometry_bwd(C, H, W, &g);
}

template <int CPL>
static int pr_launch_bwd(const float* gout, const float* rois, float* gin, const PrPlan& plan, const PrGeom& g, int N,
                         int C, int H, int W, int ngroups, int grid, float scale, int sampling_ratio, int aligned,
                         cudaStream_t stream) {
  auto k = roi_align_bwd_pr_kernel<CPL>;
  CDDMSL_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem));
  k<<<grid, kPrBwdThreads, g.smem, stream>>>(gout, rois, gin, plan, N, C, H, W, ngroups, scale, sampling_ratio,
                                             aligned, g_roi_pr_dbg);
  count_launch();
  return (int)cudaGetLastError();
}

int roi_align_bwd_pr(const float* gout, const float* rois, float* gin, int N, int C, int H, int W, int R, int P,
                     float scale, int sampling_ratio, int aligned, void* ws, cudaStream_t stream) {
  PrGeom g;
  if (P != 14 || !pr_geometry_bwd(C, H, W, &g)) return CDDMSL_EINVAL;
  const PrPlan plan = pr_carve(ws, N, R);
  const int ngroups = ceil_div(C, g.CH);
  const int chunk = pr_chunk_size(R, ngroups);
  CDDMSL_CUDA(cudaMemsetAsync(gin, 0, (size_t)N * C * H * W * sizeof(float), stream));
  int rc = pr_build_plan(plan, rois, N, H, W, R, P, scale, sampling_ratio, aligned, chunk, stream, true);
  if (rc) return rc;
  const long long max_units = ((long long)N + R / chunk + 1) * ngroups;
  const int grid = (int)min((long long)sm_count(), max_units);
  if (g.CPL == 4) return pr_launch_bwd<4>(gout, rois, gin, plan, g, N, C, H, W, ngroups, grid, scale, sampling_ratio, aligned, stream);
  if (g.CPL == 2) return pr_launch_bwd<2>(gout, rois, gin, plan, g, N, C, H, W, ngroups, grid, scale, sampling_ratio, aligned, stream);
  return pr_launch_bwd<1>(gout, rois, gin, plan, g, N, C, H, W, ngroups, grid, scale, sampling_ratio, aligned, stream);
}

int tune_roi_pr(const char* key, int value) {
  if (!strcmp(key, "roi_pr")) g_roi_pr = value;
  else if (!strcmp(key, "roi_pr_chunk")) g_roi_pr_chunk = value;
  else if (!strcmp(key, "roi_pr_cpl")) g_roi_pr_cpl = value;
  else if (!strcmp(key, "roi_pr_dbg")) g_roi_pr_dbg = value;
  else return 0;
  return 1;
}

}  // namespace cddmsl
