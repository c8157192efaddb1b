// roi_align_rotated.cu — rotated RoIAlign forward (v0 and v1 conventions) for sm_100a.
//
// Replaces ROIAlignRotatedForward + launch snippets:
//   version 1: /root/reference/python/jdet/ops/roi_align_rotated_v1.py:23-147, 300-326
//   version 0: /root/reference/python/jdet/ops/roi_align_rotated.py:21-127, 257-283
// The four convention differences (centre -0.5, rotation sign, "<0" vs "<=0" clamp, count max)
// are template branches of one kernel family.
//
// Reference shape of work: one thread per output element; every thread re-derives the RoI
// geometry (sincos included) and issues 16 scattered 4-byte loads into one NCHW plane, i.e. two
// 32-B sectors per sample for 16 useful bytes.  Here the per-RoI work is hoisted:
//
//   sample table   per RoI, once: PH*PW*gh*gw sample points -> 4 tap offsets + 4 weights, in smem,
//                  reused by every channel (256x at the bench shape).
//   staged path    (dense RoI sets) the NCHW map is re-laid once as channel-last (B,H,W,C) in the
//                  caller's workspace (tiled smem transpose, both sides coalesced); the gather
//                  kernel then reads every tap as 16-B vectors over channels: a warp covers 512
//                  contiguous bytes of one pixel.  Output slab (128 ch x PH*PW) is assembled
//                  in smem and written with coalesced 16-B streaming stores.
//   direct path    (few RoIs on a big map, where re-laying the map would cost more than it saves)
//                  NCHW gathers with the hoisted table; lanes run along the bins of one channel.
//
// Layout in HBM: input (B,C,H,W) fp32; rois (R,6) = [batch, cx, cy, w, h, theta]; output
// (R,C,PH,PW) fp32; workspace: channel-last copy (B*H*W*C fp32) for the staged path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace jdet {

constexpr int kMaxSamples = 1024;       // PH*PW*gh*gw held in smem per RoI; larger grids take the loop path

struct SampleTap {
  int o00, o01, o10, o11;               // pixel offsets (y*W + x), NOT scaled by channels; -1 => sample is out of range
  float w1, w2, w3, w4;
};

struct RoiGeom {
  int batch, gh, gw;
  float cw, ch, bin_h, bin_w, start_h, start_w, ct, st, inv_count;
};

template <int VERSION>
__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float spatial_scale, int sample_num,
                                            int PH, int PW) {
  RoiGeom g;
  g.batch = (int)roi[0];
  if (VERSION == 1) {   // roi_align_rotated_v1.py:89-90
    g.cw = __fsub_rn(__fmul_rn(roi[1], spatial_scale), 0.5f);
    g.ch = __fsub_rn(__fmul_rn(roi[2], spatial_scale), 0.5f);
  } else {              // roi_align_rotated.py:77-78
    g.cw = __fmul_rn(roi[1], spatial_scale);
    g.ch = __fmul_rn(roi[2], spatial_scale);
  }
  float rw = __fmul_rn(roi[3], spatial_scale), rh = __fmul_rn(roi[4], spatial_scale);
  const float theta = roi[5];
  rw = fmaxf(rw, 1.f);
  rh = fmaxf(rh, 1.f);
  g.bin_h = __fdiv_rn(rh, (float)PH);
  g.bin_w = __fdiv_rn(rw, (float)PW);
  g.gh = sample_num > 0 ? sample_num : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.gw = sample_num > 0 ? sample_num : (int)ceilf(__fdiv_rn(rw, (float)PW));
  g.start_h = -rh * 0.5f;
  g.start_w = -rw * 0.5f;
  // one sincos per RoI instead of one per output element; correctly-rounded via double
  const double th = (double)theta;
  g.ct = (float)cos(th);
  g.st = (float)sin(th);
  const int cnt = g.gh * g.gw;
  g.inv_count = (VERSION == 1) ? (float)max(cnt, 1) : (float)cnt;   // divisor, applied with a true division
  return g;
}

// sample point -> taps/weights, following bilinear_interpolate (v1.py:23-68 / .py:21-56)
// PACK: the taps name pixels as (y << 16) | x instead of y * W + x (the staged path splits them again without a division)
template <int VERSION, bool PACK = false>
__device__ __forceinline__ SampleTap make_tap(const RoiGeom& g, int ph, int pw, int iy, int ix, int H, int W) {
  // same operation order as the reference; explicit _rn ops so no FMA contraction moves a
  // sample across a pixel or validity boundary
  const float yy = __fadd_rn(__fadd_rn(g.start_h, __fmul_rn((float)ph, g.bin_h)),
                             __fdiv_rn(__fmul_rn((float)iy + .5f, g.bin_h), (float)g.gh));
  const float xx = __fadd_rn(__fadd_rn(g.start_w, __fmul_rn((float)pw, g.bin_w)),
                             __fdiv_rn(__fmul_rn((float)ix + .5f, g.bin_w), (float)g.gw));
  float x, y;
  if (VERSION == 1) {
    x = __fadd_rn(__fadd_rn(__fmul_rn(xx, g.ct), __fmul_rn(yy, g.st)), g.cw);
    y = __fadd_rn(__fsub_rn(__fmul_rn(yy, g.ct), __fmul_rn(xx, g.st)), g.ch);
  } else {
    x = __fadd_rn(__fsub_rn(__fmul_rn(xx, g.ct), __fmul_rn(yy, g.st)), g.cw);
    y = __fadd_rn(__fadd_rn(__fmul_rn(xx, g.st), __fmul_rn(yy, g.ct)), g.ch);
  }
  SampleTap t;
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W || !(y == y) || !(x == x)) {
    // (NaN coordinates: the reference's comparisons are all false and it then indexes with
    //  (int)NaN; that is undefined behaviour there — here such samples contribute 0.)
    t.o00 = t.o01 = t.o10 = t.o11 = -1;
    t.w1 = t.w2 = t.w3 = t.w4 = 0.f;
    return t;
  }
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = y - (float)yl, lx = x - (float)xl;
  const float hy = 1.f - ly, hx = 1.f - lx;
  if (PACK) { t.o00 = (yl << 16) | xl; t.o01 = (yl << 16) | xh; t.o10 = (yh << 16) | xl; t.o11 = (yh << 16) | xh; }
  else { t.o00 = yl * W + xl; t.o01 = yl * W + xh; t.o10 = yh * W + xl; t.o11 = yh * W + xh; }
  t.w1 = __fmul_rn(hy, hx); t.w2 = __fmul_rn(hy, lx); t.w3 = __fmul_rn(ly, hx); t.w4 = __fmul_rn(ly, lx);
  return t;
}

void launch_nchw_to_nhwc(const float* in, float* out, int B, int C, int HW, cudaStream_t st);   // relayout.cu

// ---- per-RoI tap table ---------------------------------------------------------------------------
// A bin averages gh*gw samples of 4 taps each, but at DOTA RoI sizes the samples of one bin sit closer
// than a pixel, so the 16 taps of a bin name only ~8 distinct pixels (cfg2: 767 taps -> 382 distinct
// per RoI).  The table is built in two parallel passes:
//   pass 1  one thread per sample: raw[4*sample + tap] = (pixel | -1, weight)
//   pass 2  one lane per slot, a bin's slots in one lane group (tpb = 4*gh*gw a power of two <= 32):
//           match.any finds the lanes naming the same pixel; the lowest becomes the pixel's leader and
//           sums the group's weights in ascending slot order (deterministic); a ballot compacts the
//           leaders to the front of fin[bin*tpb ..] and cnt[bin] counts them.  The tail is (-1, 0).
//           Other grid sizes only drop the out-of-range samples (no merging).
// Ends with a barrier; raw may be reused afterwards.
template <int VERSION, bool PACK = false>
__device__ __forceinline__ void build_tap_table(const RoiGeom& g, int nbins, int PW, int H, int W, int2* raw, int2* fin,
                                                int fstride, int* cnt, unsigned scale = 1u) {
  // fin holds fstride >= tpb (a multiple of 8) entries per bin: cnt[bin] merged taps (pixel index * scale, weight),
  // then padding entries (the bin's LAST pixel again, weight 0.f; (0, 0.f) for a bin without taps): a reader may run
  // fixed-width steps over the padding with plain loads — they hit the line the previous tap just fetched and add
  // 0 * value — or predicate them off by count.
  const int spb = g.gh * g.gw, tpb = 4 * spb;
  for (int s = threadIdx.x; s < nbins * spb; s += blockDim.x) {
    const int bin = s / spb, k = s - bin * spb;
    const SampleTap t = make_tap<VERSION, PACK>(g, bin / PW, bin % PW, k / g.gw, k % g.gw, H, W);
    int4* dst = reinterpret_cast<int4*>(raw + 4 * s);
    dst[0] = make_int4(t.o00, __float_as_int(t.w1), t.o01, __float_as_int(t.w2));
    dst[1] = make_int4(t.o10, __float_as_int(t.w3), t.o11, __float_as_int(t.w4));
  }
  __syncthreads();
  if (tpb <= 32 && (tpb & (tpb - 1)) == 0 && (blockDim.x & 31) == 0) {
    const int lane = threadIdx.x & 31, gpw = 32 / tpb;            // lane groups (bins) per warp
    const int sl = lane & (tpb - 1), grp = lane / tpb;
    const unsigned gm = (tpb == 32 ? 0xffffffffu : ((1u << tpb) - 1u)) << (grp * tpb);
    const unsigned below = (1u << lane) - 1u;
    for (int bin0 = (threadIdx.x >> 5) * gpw; bin0 < nbins; bin0 += (blockDim.x >> 5) * gpw) {
      const int bin = bin0 + grp;
      const bool in = bin < nbins;
      const int2* rb = raw + (in ? bin : 0) * tpb;
      const int2 me = in ? rb[sl] : make_int2(-1, 0);
      const unsigned m = __match_any_sync(0xffffffffu, me.x) & gm;
      const bool leader = me.x >= 0 && (__ffs(m) - 1) == lane;
      float wsum = 0.f;
      if (leader)
        for (unsigned mm = m; mm; mm &= mm - 1) wsum += __int_as_float(rb[(__ffs(mm) - 1) & (tpb - 1)].y);
      const unsigned lb = __ballot_sync(0xffffffffu, leader) & gm;
      const int nlead = __popc(lb);
      const int pos = leader ? __popc(lb & below) : nlead + __popc(~lb & gm & below);
      const int last = __shfl_sync(0xffffffffu, me.x, lb ? 31 - __clz(lb) : 0);       // pixel of the bin's last merged tap
      const int2 pad = make_int2(nlead ? (int)((unsigned)last * scale) : 0, 0);
      if (in) {
        int2* fb = fin + bin * fstride;
        fb[pos] = leader ? make_int2((int)((unsigned)me.x * scale), __float_as_int(wsum)) : pad;
        for (int i = tpb + sl; i < fstride; i += tpb) fb[i] = pad;
        if (sl == 0) cnt[bin] = nlead;
      }
    }
  } else {
    for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
      int2* fb = fin + bin * fstride;
      int n = 0;
      int2 pad = make_int2(0, 0);
      for (int j = 0; j < tpb; j++) {
        const int2 e = raw[bin * tpb + j];
        if (e.x >= 0) { pad = make_int2((int)((unsigned)e.x * scale), 0); fb[n++] = make_int2(pad.x, e.y); }
      }
      cnt[bin] = n;
      for (; n < fstride; n++) fb[n] = pad;
    }
  }
  __syncthreads();
}

// FPN levels of one call, passed by value to the kernels.  n == 1: the plain single-map op.  n > 1: the fused
// single-level-per-RoI extractor (OrientedSingleRoIExtractor / RboxSingleRoIExtractor.execute): a RoI is stretched
// by (ext_w, ext_h), assigned to level clamp(floor(log2(sqrt(w*h) / finest + 1e-6)), 0, n-1), stretched again by
// (rs_w, rs_h) and pooled from that level with the level's spatial scale.
constexpr int kMaxLevels = 8;
struct RoiLevel { const float* nhwc; int H, W; float scale; };
struct RoiLevels { RoiLevel lv[kMaxLevels]; int n; float ext_w, ext_h, rs_w, rs_h, finest; };

// ---- prologue kernel: per-RoI tap tables (+ the NCHW -> channel-last re-layout) -------------------------
// Table record of one RoI in the caller's workspace (16-B granular):
//   int4 {batch, float bits of the divisor, fin entries per bin, FPN level} | int cnt[nbins rounded up to 4] | int2 fin[nbins][fstride]
// Building it is latency-bound ALU work (double sincos, two barriers, match.any); the re-layout is an HBM
// stream.  One launch does both, interleaved (see the kernel), so the gather CTAs start with nothing but loads
// to do.
__host__ __device__ inline int roi_table_fstride(int sampling_ratio) { return (4 * sampling_ratio * sampling_ratio + 7) & ~7; }
__host__ __device__ inline int roi_table_layout_bytes(int nbins, int sampling_ratio);   // record pitch: roi_gather_tma.cuh's rec_layout
__host__ __device__ inline size_t roi_table_stride(int nbins, int sampling_ratio) { return (size_t)roi_table_layout_bytes(nbins, sampling_ratio); }

// Work-queue block behind the tables: int ctr[64] (ctr[0] = the gather's item counter, ctr[16 + b] = RoIs in cost
// bucket b) and int order[kRoiBuckets][R] (the RoIs of bucket b in arrival order).  The gather walks the buckets from
// the most expensive RoIs to the cheapest (longest-processing-time-first): with RoIs taken in input order the last
// work items of the persistent CTAs are as likely large as small and 14 % of the kernel was tail (ncu: SM active
// 127k of 148k cycles; handing the same RoIs over largest-first: 115.7 -> 107.5 us for the whole op).
constexpr int kRoiBuckets = 8;
// (+ behind the bucket lists: int staged[R], the RoIs of the TMA-staged path in arrival order, ctr[2] of them, ctr[1] its item
//  counter; and int rec_bytes[R], the size of each RoI's table record)
__host__ __device__ inline size_t roi_queue_bytes(int R) { return 256 + (((size_t)(kRoiBuckets + 2) * (R > 0 ? R : 0) * 4 + 255) & ~(size_t)255); }
__device__ __forceinline__ int roi_cost_bucket(int taps) { return max(0, kRoiBuckets - 1 - taps / 96); }   // 0 = most taps (<= 784)
// rank-th RoI in bucket order (rank < R)
__device__ __forceinline__ int roi_by_rank(const int* __restrict__ ctr, int R, int rank) {
  const int* order = ctr + 64;
#pragma unroll
  for (int b = 0; b < kRoiBuckets; b++) {
    const int n = ctr[16 + b];
    if (rank < n) return order[(size_t)b * R + rank];
    rank -= n;
  }
  return 0;   // unreachable when the buckets hold all R RoIs
}

}  // namespace jdet
#include "roi_gather_tma.cuh"
namespace jdet {
__host__ __device__ inline int roi_table_layout_bytes(int nbins, int sampling_ratio) {
  return g4::rec_layout(nbins, roi_table_fstride(sampling_ratio), 4 * sampling_ratio * sampling_ratio).max_bytes;
}

__device__ __forceinline__ void relayout_tile(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int b,
                                              int p0, int c0, float (*tile)[33]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* src = in + (size_t)b * C * HW;
  float* dst = out + (size_t)b * C * HW;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int c = c0 + ty + 8 * k, p = p0 + tx;
    if (c < C && p < HW) tile[ty + 8 * k][tx] = __ldg(src + (size_t)c * HW + p);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int p = p0 + ty + 8 * k, c = c0 + tx;
    if (c < C && p < HW) dst[(size_t)p * C + c] = tile[tx][ty + 8 * k];
  }
}

// 32 channels x 128 pixels per CTA with 16-B global accesses on both sides (HW % 4 == 0, C % 4 == 0, 16-B aligned
// bases): 4 independent loads per thread in flight and a quarter of the load/store instructions of the scalar
// 32x32 tile, which is issue-bound (72 % issue slots busy at 5 TB/s).
constexpr int kTileW = 128;
__device__ __forceinline__ void relayout_tile_v4(const float* __restrict__ in, float* __restrict__ out, int C, int HW, int b,
                                                 int p0, int c0, float (*tile)[kTileW + 1]) {
  const float* src = in + (size_t)b * C * HW;
  float* dst = out + (size_t)b * C * HW;
  {
    const int row = threadIdx.x >> 3, c = c0 + row;                // channel row
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int p = p0 + 4 * ((threadIdx.x & 7) + 8 * i);
      v[i] = (c < C && p < HW) ? __ldg(reinterpret_cast<const float4*>(src + (size_t)c * HW + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float* t = &tile[row][4 * ((threadIdx.x & 7) + 8 * i)];
      t[0] = v[i].x; t[1] = v[i].y; t[2] = v[i].z; t[3] = v[i].w;
    }
  }
  __syncthreads();
  {
    const int cq = threadIdx.x & 7, c = c0 + 4 * cq;               // channel quad
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int px = (threadIdx.x >> 3) + 32 * i, p = p0 + px;
      if (p < HW && c < C)
        *reinterpret_cast<float4*>(dst + (size_t)p * C + c) =
            make_float4(tile[4 * cq + 0][px], tile[4 * cq + 1][px], tile[4 * cq + 2][px], tile[4 * cq + 3][px]);
    }
  }
}

// grid = (tiles_x + tables_per_row, tiles_y * B): in every grid row the first tiles_x CTAs transpose a tile, the
// rest build tables (tables_per_row = ceil(R / rows)), so both kinds of work are in flight throughout and no
// index needs a division.  tiles_x == 0: tables only (channel-last input).
// STAGED: the opt-in TMA path (packed taps, chunk_record); false compiles the plain LSU-table build only
template <int VERSION, bool STAGED>
__global__ void __launch_bounds__(256) roi_prologue_kernel(const float* __restrict__ in, float* __restrict__ nhwc, int C, int H,
                                                            int W, int tiles_x, int tiles_y, const float* __restrict__ rois,
                                                            unsigned R, int PH, int PW, int sample_num,
                                                            unsigned char* __restrict__ tables, bool vec,
                                                            int* __restrict__ work_counter, int pcap, int pxbytes, int rec_pitch,
                                                            const __grid_constant__ RoiLevels L) {
  extern __shared__ __align__(128) unsigned char smem[];
  if ((int)blockIdx.x < tiles_x) {
    const int b = tiles_y > 0 ? blockIdx.y / tiles_y : 0, ty = blockIdx.y - b * tiles_y;
    if (vec) relayout_tile_v4(in, nhwc, C, H * W, b, blockIdx.x * kTileW, ty * 32, reinterpret_cast<float(*)[kTileW + 1]>(smem));
    else relayout_tile(in, nhwc, C, H * W, b, blockIdx.x * 32, ty * 32, reinterpret_cast<float(*)[33]>(smem));
    return;
  }
  const unsigned idx = blockIdx.y * (gridDim.x - tiles_x) + (blockIdx.x - tiles_x);
  if (idx >= R) return;
  const int nbins = PH * PW;
  __shared__ RoiGeom g;
  __shared__ int s_level;
  if (threadIdx.x == 0) {
    float roi[6];
#pragma unroll
    for (int i = 0; i < 6; i++) roi[i] = rois[(size_t)idx * 6 + i];
    int level = 0;
    if (L.n > 1) {   // map_roi_levels (oriented_single_level.py:67-70, 100-102), same fp32 operations as the torch mirror
      roi[3] = __fmul_rn(roi[3], L.ext_w);
      roi[4] = __fmul_rn(roi[4], L.ext_h);
      const float side = sqrtf(__fmul_rn(roi[3], roi[4]));
      const float lf = floorf(log2f(__fadd_rn(__fdiv_rn(side, L.finest), 1e-6f)));
      level = lf <= 0.f ? 0 : (lf >= (float)(L.n - 1) ? L.n - 1 : (int)lf);   // NaN -> 0
      roi[3] = __fmul_rn(roi[3], L.rs_w);
      roi[4] = __fmul_rn(roi[4], L.rs_h);
    }
    s_level = level;
    g = roi_geom<VERSION>(roi, L.lv[level].scale, sample_num, PH, PW);
  }
  __syncthreads();
  const int level = s_level;
  const int fstride = roi_table_fstride(sample_num), tpb = 4 * sample_num * sample_num;
  const g4::RecLayout RL = g4::rec_layout(nbins, fstride, tpb);
  int4* hdr = reinterpret_cast<int4*>(smem);                       // the record, in its final layout (roi_gather_tma.cuh)
  int* cnt = reinterpret_cast<int*>(smem + RL.cnt_off);
  int2* fin = reinterpret_cast<int2*>(smem + RL.fin_off);
  int2* raw = reinterpret_cast<int2*>(smem + rec_pitch);   // rec_pitch: bytes per record in `tables` (without the pixel lists when nothing is staged)
  g4::ChunkScratch* scratch = reinterpret_cast<g4::ChunkScratch*>(raw + nbins * tpb);
  if (threadIdx.x == 0) hdr[0] = make_int4(g.batch, __float_as_int(g.inv_count), fstride, level);
  const int Hl = L.lv[level].H, Wl = L.lv[level].W;
  if constexpr (STAGED) build_tap_table<VERSION, true>(g, nbins, PW, Hl, Wl, raw, fin, fstride, cnt);     // entries: ((y << 16) | x, weight)
  else build_tap_table<VERSION, false>(g, nbins, PW, Hl, Wl, raw, fin, fstride, cnt, 4u * (unsigned)C);   // LSU path only: byte offsets at once
  // One staged chunk (roi_gather_tma.cuh) when the RoI's distinct pixels fit pcap — guess first from the geometry, distinct
  // pixels ~ area + 1.5 * (w + h) of the scaled box (cfg2 at 128 pixels: 3 wasted attempts, 15 missed of 2048) —, else the
  // LSU-path record: entries become byte offsets into one image of the channel-last map.
  int bytes = RL.cnpx_off;
  bool staged = false;
  if constexpr (STAGED) {
    const float rw = g.bin_w * (float)PW, rh = g.bin_h * (float)PH;
    if (rw * rh + 1.5f * (rw + rh) + 2.f <= (float)pcap)
      staged = g4::chunk_record(smem, nbins, fstride, tpb, pcap, pxbytes, Wl, g.batch * Hl * Wl, *scratch, &bytes);
    if (!staged) {
      const unsigned px4c = 4u * (unsigned)C;
      for (int e = threadIdx.x; e < nbins * fstride; e += blockDim.x) {
        const int p = fin[e].x;
        fin[e].x = (int)((unsigned)((p >> 16) * Wl + (p & 0xffff)) * px4c);
      }
    }
  }
  if (!staged) {
    if (threadIdx.x == 0) hdr[1] = make_int4(0, 0, 0, bytes);
    __syncthreads();
  }
  int4* dst = reinterpret_cast<int4*>(tables + (size_t)idx * (size_t)rec_pitch);
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) dst[i] = hdr[i];
  if constexpr (STAGED) {
    if (threadIdx.x == 0) work_counter[64 + (size_t)(kRoiBuckets + 1) * R + idx] = bytes;
  }
  if (staged) {
    if (threadIdx.x == 0) work_counter[64 + (size_t)kRoiBuckets * R + atomicAdd(work_counter + 2, 1)] = (int)idx;
    return;
  }
  if (threadIdx.x < 32) {   // cost = merged taps of the RoI -> its bucket of the gather's work queue (counters zeroed by the host side's memset)
    int taps = 0;
    for (int b = threadIdx.x; b < nbins; b += 32) taps += cnt[b];
#pragma unroll
    for (int d = 16; d; d >>= 1) taps += __shfl_xor_sync(0xffffffffu, taps, d);
    if (threadIdx.x == 0) {
      const int b = roi_cost_bucket(taps);
      const int pos = atomicAdd(work_counter + 16 + b, 1);
      work_counter[64 + (size_t)b * R + pos] = (int)idx;
    }
  }
}

// ---- LSU gather kernel: the RoIs the prologue did not stage (more distinct pixels than a chunk holds) ------------
// grid = (R, C / SLAB), SLAB = 4*QL channels.  QL lanes span a bin's channels (QL = 32: one bin per warp, every
// tap is 512 contiguous bytes of one channel-last pixel).  The CTA copies its RoI's table record (6.7 KB at
// 7x7x2x2) into smem and then does nothing but loads and FMAs: bins are handed to warps through an smem counter
// (their tap counts differ).  The output slab (SLAB x PH*PW, contiguous in (R,C,PH,PW)) is assembled in smem
// and leaves as ONE cp.async.bulk store.
// (256, 4) / (256, 3): without a blocks-per-SM hint ptxas squeezes the kernel into 32 registers by sinking every load
// next to its FMAs; 64 registers (80 for the two-bins-per-warp form) keep the 8 loads of a step in flight.
#ifndef JDET_ROI_PAIR_MINB
#define JDET_ROI_PAIR_MINB 3
#endif
template <int QL, bool PAIR>
__global__ void __launch_bounds__(256, PAIR ? JDET_ROI_PAIR_MINB : 4) roi_gather_kernel(const __grid_constant__ RoiLevels L,
                                                             const unsigned char* __restrict__ tables, size_t stride, int C,
                                                             int nbins, int nslabs, int R, uint32_t rec_bytes, int fin_off, int items_known,
                                                             int* __restrict__ work_counter, float* __restrict__ out) {
  constexpr int SLAB = PAIR ? 8 * QL : 4 * QL;                      // PAIR: a lane owns two channel quads 4*QL apart
  extern __shared__ __align__(128) unsigned char smem[];
  const int S = nbins | 1;                                         // odd row stride of s_out: see the store below
  float* s_out = reinterpret_cast<float*>(smem);                   // [SLAB][S]
  unsigned char* rec_base = reinterpret_cast<unsigned char*>(s_out + ((SLAB * S + 3) & ~3));   // 2 table records
  __shared__ uint64_t full[2];
  __shared__ int next_bin, s_next_item;
  const int lane = threadIdx.x & 31, q = lane % QL;                // channel quad within a 128-channel group
  const unsigned gmask = QL == 32 ? 0xffffffffu : (((1u << QL) - 1u) << (lane & ~(QL - 1)));
  const int rot = (q >> 3) & 3;
  const int ro0 = (4 * q + ((0 + rot) & 3)) * S, ro1 = (4 * q + ((1 + rot) & 3)) * S;
  const int ro2 = (4 * q + ((2 + rot) & 3)) * S, ro3 = (4 * q + ((3 + rot) & 3)) * S;

  // Persistent CTA: work items (RoI, slab) come from a global counter (RoIs differ 4x in taps); the table of the
  // NEXT item is fetched by a bulk-async copy while the current one is gathered, and the slab store of the
  // PREVIOUS item drains while the current one runs — a CTA never sits waiting for its 6.7 KB table or its store.
  int cur = blockIdx.x;                                            // first item: static; later ones from the counter
  int items = items_known;                                         // >= 0: every RoI takes this path (nothing staged): known on the host
  if (items < 0) {                                                 // else the RoIs the prologue left to this path = the LPT buckets
    items = 0;
#pragma unroll
    for (int b = 0; b < kRoiBuckets; b++) items += work_counter[16 + b];
    items *= nslabs;
  }
  __shared__ int s_roi[2];                                         // RoI of the item whose table sits in rec buffer 0 / 1
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init_fence();
    if (cur < items) {
      const int roi = roi_by_rank(work_counter, R, cur / nslabs);
      s_roi[0] = roi;
      mbar_expect_tx(&full[0], rec_bytes);
      bulk_g2s(rec_base, tables + (size_t)roi * stride, rec_bytes, &full[0]);
    }
  }
  __syncthreads();
  uint32_t buf = 0, phases = 0;                                    // bit b: parity to wait for on full[b]
  while (cur < items) {
    if (threadIdx.x == 0) {
      next_bin = 0;
      bulk_s2g_wait_read();                                        // the previous slab has left s_out
    }
    mbar_wait(&full[buf], (phases >> buf) & 1u);                   // this item's table has landed
    phases ^= 1u << buf;
    __syncthreads();
    if (threadIdx.x == 0) {                                        // after the barrier: nobody waits on the atomic's round trip
      const int nx = atomicAdd(work_counter, 1) + (int)gridDim.x;
      s_next_item = nx;                                            // read by everyone after the barrier before the store
      if (nx < items) {
        const int roi = roi_by_rank(work_counter, R, nx / nslabs);
        s_roi[buf ^ 1] = roi;                                      // (slot buf ^ 1 was last read before the barrier above)
        mbar_expect_tx(&full[buf ^ 1], rec_bytes);
        bulk_g2s(rec_base + (buf ^ 1) * rec_bytes, tables + (size_t)roi * stride, rec_bytes, &full[buf ^ 1]);
      }
    }
    const int r = s_roi[buf], c0 = (cur % nslabs) * SLAB;
    const int4* rec = reinterpret_cast<const int4*>(rec_base + buf * rec_bytes);
    const int batch = rec[0].x, fstride = rec[0].z;
    const float count = __int_as_float(rec[0].y);
    const int* cnt = reinterpret_cast<const int*>(rec + 2);        // record layout: roi_gather_tma.cuh (two header words, cnt, fin)
    const int2* fin = reinterpret_cast<const int2*>(reinterpret_cast<const unsigned char*>(rec) + fin_off);
    const RoiLevel& lv = L.lv[rec[0].w];
    const float* base = lv.nhwc + (size_t)batch * lv.H * lv.W * C + c0 + 4 * q;
    const int icnt = (int)count;
    const bool pow2 = (icnt & (icnt - 1)) == 0;                    // x / 2^k == x * 2^-k exactly
    const float rcnt = 1.f / count;
  if constexpr (PAIR) {
    // (A software-pipelined form of this loop — double-buffered load registers, static pair order, the next step's loads
    //  issued before the current step's FMAs, also across pair boundaries; 128 registers, 2 CTAs/SM — measured 127 us
    //  against 115 us for this one on the same box, same output hash: occupancy beats per-warp load depth here.)
    // Two bins per warp (one per half-warp of 16 lanes, 8 channels per lane): the per-bin bookkeeping, the table
    // reads and the address arithmetic of a step are issued once for both bins.  Control flow is uniform across the
    // warp (steps run to the larger tap count, loads are predicated per lane), 4 taps x 2 quads = 8 loads per step.
    const int half = lane >> 4;
    for (;;) {
      int p = 0;
      if (lane == 0) p = atomicAdd(&next_bin, 1);
      p = __shfl_sync(0xffffffffu, p, 0);
      if (2 * p >= nbins) break;
      const int bin = 2 * p + half;
      const bool valid = bin < nbins;
      const int n = valid ? cnt[bin] : 0;
      const int nmax = max(n, __shfl_xor_sync(0xffffffffu, n, 16));
      const int2* e = fin + (valid ? bin : 0) * fstride;
      float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
      // A step = 4 table entries x 2 channel quads = 8 loads, all issued before the first FMA.  Entries past this
      // bin's count are padding (its last pixel again, weight 0): plain loads that hit the line just fetched, so a
      // step carries no per-tap predicates or zero-fills; only a bin without any tap (n == 0, paired with a live
      // one) keeps its loads switched off, and its registers stay at the zeros set here.
      const bool live = n > 0;
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < nmax; k += 4) {
        const int4 t01 = *reinterpret_cast<const int4*>(e + k), t23 = *reinterpret_cast<const int4*>(e + k + 2);
        const float* p0 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (unsigned)t01.x);
        const float* p1 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (unsigned)t01.z);
        const float* p2 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (unsigned)t23.x);
        const float* p3 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (unsigned)t23.z);
        ldg_v4_keep(v[0], p0, live); ldg_v4_keep(v[1], p0 + 4 * QL, live);
        ldg_v4_keep(v[2], p1, live); ldg_v4_keep(v[3], p1 + 4 * QL, live);
        ldg_v4_keep(v[4], p2, live); ldg_v4_keep(v[5], p2 + 4 * QL, live);
        ldg_v4_keep(v[6], p3, live); ldg_v4_keep(v[7], p3 + 4 * QL, live);
#pragma unroll
        for (int j = 0; j < 8; j++)   // scheduling fence: no FMA may be hoisted between the loads
          asm volatile("" : "+f"(v[j].x), "+f"(v[j].y), "+f"(v[j].z), "+f"(v[j].w));
        const float w0 = __int_as_float(t01.y), w1 = __int_as_float(t01.w);   // 0 on padding
        const float w2 = __int_as_float(t23.y), w3 = __int_as_float(t23.w);
        // packed FFMA2: same fma chain per component (tap 0, 1, 2, 3), two components per issue slot
#define JDET_ACC(a, i0)                                                                                                \
  do {                                                                                                                 \
    g4::fma2(a.x, a.y, w0, v[i0].x, v[i0].y);         g4::fma2(a.z, a.w, w0, v[i0].z, v[i0].w);                         \
    g4::fma2(a.x, a.y, w1, v[i0 + 2].x, v[i0 + 2].y); g4::fma2(a.z, a.w, w1, v[i0 + 2].z, v[i0 + 2].w);                 \
    g4::fma2(a.x, a.y, w2, v[i0 + 4].x, v[i0 + 4].y); g4::fma2(a.z, a.w, w2, v[i0 + 4].z, v[i0 + 4].w);                 \
    g4::fma2(a.x, a.y, w3, v[i0 + 6].x, v[i0 + 6].y); g4::fma2(a.z, a.w, w3, v[i0 + 6].z, v[i0 + 6].w);                 \
  } while (0)
#ifdef JDET_ROI_SCALAR_FMA   // A/B switch (tools/ab_libs.py): the round-1 scalar chain
#define JDET_ACC1(a, c, i0) a.c = fmaf(w3, v[i0 + 6].c, fmaf(w2, v[i0 + 4].c, fmaf(w1, v[i0 + 2].c, fmaf(w0, v[i0].c, a.c))))
        JDET_ACC1(acc0, x, 0); JDET_ACC1(acc0, y, 0); JDET_ACC1(acc0, z, 0); JDET_ACC1(acc0, w, 0);
        JDET_ACC1(acc1, x, 1); JDET_ACC1(acc1, y, 1); JDET_ACC1(acc1, z, 1); JDET_ACC1(acc1, w, 1);
#undef JDET_ACC1
#else
        JDET_ACC(acc0, 0);
        JDET_ACC(acc1, 1);
#endif
#undef JDET_ACC
      }
      if (valid) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          float4 a = u ? acc1 : acc0;
          if (pow2) { a.x *= rcnt; a.y *= rcnt; a.z *= rcnt; a.w *= rcnt; }
          else { a.x /= count; a.y /= count; a.z /= count; a.w /= count; }
          const float b0 = rot & 1 ? a.y : a.x, b1 = rot & 1 ? a.z : a.y, b2 = rot & 1 ? a.w : a.z, b3 = rot & 1 ? a.x : a.w;
          const float d0 = rot & 2 ? b2 : b0, d1 = rot & 2 ? b3 : b1, d2 = rot & 2 ? b0 : b2, d3 = rot & 2 ? b1 : b3;
          float* row = s_out + u * 4 * QL * S + bin;
          row[ro0] = d0; row[ro1] = d1; row[ro2] = d2; row[ro3] = d3;
        }
      }
    }
  } else {
  // bins -> warps: round-robin when the bin count is a multiple of the warp count, else handed out through an smem counter
  const bool static_bins = QL == 32 && nbins % (int)(blockDim.x >> 5) == 0;
  for (int sbin = threadIdx.x >> 5;; sbin += blockDim.x >> 5) {
    int bin = sbin;
    if (!static_bins) {
      if (q == 0) bin = atomicAdd(&next_bin, 1);
      bin = __shfl_sync(gmask, bin, lane & ~(QL - 1));
    }
    if (bin >= nbins) break;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int2* e = fin + bin * fstride;
    const int n = cnt[bin];
    // One latency round per 8 taps: the loads of a step are all issued before its first FMA (a load costs ~1000
    // cycles under load; two thirds of the bins have <= 8 taps).  Full steps carry no predicates; the ragged tail
    // is one more step of (4 plain +) up to 3 predicated loads — never dummy loads: every 512-B tap costs ~8
    // cycles of the SM's L1 data path whether it hits or not, and that path and the issue slots, not HBM or L2,
    // are what this kernel runs out of.
#define JDET_TAP_ADDR(j) reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + (unsigned)t[j].x)
#define JDET_TAP_FMA(j)                                                                                                \
  do {                                                                                                                 \
    const float w_ = __int_as_float(t[j].y);                                                                           \
    acc.x += w_ * v[j].x; acc.y += w_ * v[j].y; acc.z += w_ * v[j].z; acc.w += w_ * v[j].w;                            \
  } while (0)
    int k = 0;
    for (; n - k >= 8; k += 8) {
      int2 t[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const int4 two = *reinterpret_cast<const int4*>(e + k + j);
        t[j] = make_int2(two.x, two.y);
        t[j + 1] = make_int2(two.z, two.w);
      }
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = __ldg(reinterpret_cast<const float4*>(JDET_TAP_ADDR(j)));
#pragma unroll
      for (int j = 0; j < 8; j++) JDET_TAP_FMA(j);
    }
    const int rem = n - k;                                           // 0..7 taps left; the table is readable (and zero-weighted) up to k + 8
    if (rem >= 4) {
      int2 t[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const int4 two = *reinterpret_cast<const int4*>(e + k + j);
        t[j] = make_int2(two.x, two.y);
        t[j + 1] = make_int2(two.z, two.w);
      }
      float4 v[7];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = __ldg(reinterpret_cast<const float4*>(JDET_TAP_ADDR(j)));
#pragma unroll
      for (int j = 4; j < 7; j++) v[j] = ldg_v4_if(JDET_TAP_ADDR(j), j < rem);
#pragma unroll
      for (int j = 0; j < 7; j++) JDET_TAP_FMA(j);
    } else if (rem > 0) {
      int2 t[4];
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int4 two = *reinterpret_cast<const int4*>(e + k + j);
        t[j] = make_int2(two.x, two.y);
        t[j + 1] = make_int2(two.z, two.w);
      }
      float4 v[3];
#pragma unroll
      for (int j = 0; j < 3; j++) v[j] = ldg_v4_if(JDET_TAP_ADDR(j), j < rem);
#pragma unroll
      for (int j = 0; j < 3; j++) JDET_TAP_FMA(j);
    }
#undef JDET_TAP_ADDR
#undef JDET_TAP_FMA
    // s_out[channel][bin], row stride S odd.  A lane's 4 channels are 4 rows apart from its neighbour's, so
    // a plain store would hit 8 banks 4 lanes deep; lanes 8 apart instead store their components in a
    // rotated order (rot = lane/8), which spreads every store instruction over all 32 banks.
    {
      float4 a = acc;
      if (pow2) { a.x *= rcnt; a.y *= rcnt; a.z *= rcnt; a.w *= rcnt; }
      else { a.x /= count; a.y /= count; a.z /= count; a.w /= count; }
      const float b0 = rot & 1 ? a.y : a.x, b1 = rot & 1 ? a.z : a.y, b2 = rot & 1 ? a.w : a.z, b3 = rot & 1 ? a.x : a.w;
      const float d0 = rot & 2 ? b2 : b0, d1 = rot & 2 ? b3 : b1, d2 = rot & 2 ? b0 : b2, d3 = rot & 2 ? b1 : b3;
      float* row = s_out + bin;                                     // d_j is component (j + rot) & 3 of this lane's channel quad
      row[ro0] = d0;
      row[ro1] = d1;
      row[ro2] = d2;
      row[ro3] = d3;
    }
  }
  }   // !PAIR
    // out[r][c0 .. c0+SLAB-1][bins] is one contiguous run of SLAB*nbins floats
    float* dst = out + ((size_t)r * C + c0) * nbins;
    const int total = SLAB * nbins;
    if (S == nbins && (total & 3) == 0 && (((uintptr_t)dst) & 15) == 0) {
      fence_proxy_async_smem();                                    // generic-proxy smem writes -> visible to the bulk copy
      __syncthreads();
      if (threadIdx.x == 0) bulk_s2g(dst, s_out, (uint32_t)total * 4u);
    } else {
      __syncthreads();
      for (int i = threadIdx.x; i < total; i += blockDim.x) st_stream(dst + i, s_out[(i / nbins) * S + i % nbins]);
    }
    cur = s_next_item;                                             // (rewritten only after the next item's first barrier)
    buf ^= 1u;
  }
  if (threadIdx.x == 0) bulk_s2g_wait_read();                      // smem must outlive the last store's read
}


// ---- direct NCHW kernel --------------------------------------------------------------------------
// grid = (R, channel chunks); thread task = (channel, bin); taps from the per-RoI table when it fits,
// otherwise recomputed per element (adaptive sampling grids of huge RoIs).
template <int VERSION>
__global__ void __launch_bounds__(256) roi_align_nchw_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                                                              int C, int H, int W, int PH, int PW, float spatial_scale,
                                                              int sample_num, int ch_per_cta, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nbins = PH * PW;
  const int r = blockIdx.x, c0 = blockIdx.y * ch_per_cta;
  const int c1 = min(C, c0 + ch_per_cta);
  __shared__ RoiGeom g;
  if (threadIdx.x == 0) g = roi_geom<VERSION>(rois + (size_t)r * 6, spatial_scale, sample_num, PH, PW);
  __syncthreads();
  const int spb = g.gh * g.gw;
  const bool tabled = (long long)nbins * spb <= kMaxSamples;
  SampleTap* taps = reinterpret_cast<SampleTap*>(smem);
  if (tabled) {
    for (int s = threadIdx.x; s < nbins * spb; s += blockDim.x) {
      const int bin = s / spb, k = s - bin * spb;
      taps[s] = make_tap<VERSION>(g, bin / PW, bin % PW, k / g.gw, k % g.gw, H, W);
    }
    __syncthreads();
  }
  const size_t plane = (size_t)H * W;
  const float* fb = feat + (size_t)g.batch * C * plane;
  const int total = (c1 - c0) * nbins;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = c0 + i / nbins, bin = i % nbins;
    const float* p = fb + (size_t)c * plane;
    float acc = 0.f;
    if (tabled) {
      const SampleTap* tp = taps + bin * spb;
      for (int k = 0; k < spb; k++) {
        const SampleTap t = tp[k];
        if (t.o00 < 0) continue;
        acc += t.w1 * __ldg(p + t.o00) + t.w2 * __ldg(p + t.o01) + t.w3 * __ldg(p + t.o10) + t.w4 * __ldg(p + t.o11);
      }
    } else {
      const int ph = bin / PW, pw = bin % PW;
      for (int iy = 0; iy < g.gh; iy++)
        for (int ix = 0; ix < g.gw; ix++) {
          const SampleTap t = make_tap<VERSION>(g, ph, pw, iy, ix, H, W);
          if (t.o00 < 0) continue;
          acc += t.w1 * __ldg(p + t.o00) + t.w2 * __ldg(p + t.o01) + t.w3 * __ldg(p + t.o10) + t.w4 * __ldg(p + t.o11);
        }
    }
    out[((size_t)r * C + c) * nbins + bin] = acc / g.inv_count;
  }
}

// ---- backward ------------------------------------------------------------------------------------
// Replaces ROIAlignBackward (roi_align_rotated_v1.py:192-298 / roi_align_rotated.py:164-255):
// grad_input[b,c,tap] += grad_out[r,c,bin] * w / (gh*gw).  Same per-RoI sample table as the forward.
//   staged: gradients are accumulated in a channel-last scratch with 16-B vector atomics
//           (red.global.add.v4.f32: one L2 atomic per 4 channels instead of 4), then re-laid to NCHW;
//   direct: scalar atomics into the NCHW gradient (few RoIs / odd channel counts).
// Float atomics make the last bits order-dependent, exactly as in the reference.
template <int VERSION, int SLAB>
__global__ void __launch_bounds__(256) roi_align_bwd_nhwc_kernel(const float* __restrict__ grad_out,
                                                                  const float* __restrict__ rois, int C, int H, int W,
                                                                  int PH, int PW, float spatial_scale, int sample_num,
                                                                  float* __restrict__ grad_nhwc) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nbins = PH * PW;
  const int r = blockIdx.x, c0 = blockIdx.y * SLAB;
  constexpr int QL = SLAB / 4;
  __shared__ RoiGeom g;
  if (threadIdx.x == 0) g = roi_geom<VERSION>(rois + (size_t)r * 6, spatial_scale, sample_num, PH, PW);
  __syncthreads();
  const int spb = g.gh * g.gw;
  const int tpb = 4 * spb, nslots = nbins * tpb;
  int2* fin = reinterpret_cast<int2*>(smem);
  int2* raw = fin + nslots;
  float* s_go = reinterpret_cast<float*>(raw + nslots);         // [SLAB][nbins], as laid out in grad_out
  int* cnt = reinterpret_cast<int*>(s_go + SLAB * nbins);
  const float* src = grad_out + ((size_t)r * C + c0) * nbins;
  for (int i = threadIdx.x; i < SLAB * nbins; i += blockDim.x) s_go[i] = __ldg(src + i);
  build_tap_table<VERSION>(g, nbins, PW, H, W, raw, fin, tpb, cnt);  // one atomic per distinct pixel of a bin
  float* base = grad_nhwc + (size_t)g.batch * H * W * C + c0;
  const int q = threadIdx.x % QL;
  const float cntf = (float)spb;                     // the backward divides by gh*gw in both versions
  // (grad * w) / count as the reference writes it; for a power-of-two count (sampling_ratio 2: 4) the division is the exact
  // multiplication by 2^-k — four full-precision divisions per tap were most of this loop's instructions
  const bool pow2 = (spb & (spb - 1)) == 0;
  const float rcnt = 1.f / cntf;
  for (int bin = threadIdx.x / QL; bin < nbins; bin += blockDim.x / QL) {
    const float4 go = make_float4(s_go[(4 * q + 0) * nbins + bin], s_go[(4 * q + 1) * nbins + bin],
                                  s_go[(4 * q + 2) * nbins + bin], s_go[(4 * q + 3) * nbins + bin]);
    const int2* e = fin + bin * tpb;
    const int n = cnt[bin];
    if (pow2) {
      for (int k = 0; k < n; k++) {
        const int2 t = e[k];
        const float w = __int_as_float(t.y);
        const float4 v = make_float4(__fmul_rn(__fmul_rn(go.x, w), rcnt), __fmul_rn(__fmul_rn(go.y, w), rcnt),
                                     __fmul_rn(__fmul_rn(go.z, w), rcnt), __fmul_rn(__fmul_rn(go.w, w), rcnt));
        atomicAdd(reinterpret_cast<float4*>(base + (size_t)t.x * C) + q, v);
      }
    } else {
      for (int k = 0; k < n; k++) {
        const int2 t = e[k];
        const float w = __int_as_float(t.y);
        const float4 v = make_float4(go.x * w / cntf, go.y * w / cntf, go.z * w / cntf, go.w * w / cntf);
        atomicAdd(reinterpret_cast<float4*>(base + (size_t)t.x * C) + q, v);
      }
    }
  }
}

template <int VERSION>
__global__ void __launch_bounds__(256) roi_align_bwd_nchw_kernel(const float* __restrict__ grad_out,
                                                                  const float* __restrict__ rois, int C, int H, int W,
                                                                  int PH, int PW, float spatial_scale, int sample_num,
                                                                  int ch_per_cta, float* __restrict__ grad_in) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nbins = PH * PW;
  const int r = blockIdx.x, c0 = blockIdx.y * ch_per_cta;
  const int c1 = min(C, c0 + ch_per_cta);
  __shared__ RoiGeom g;
  if (threadIdx.x == 0) g = roi_geom<VERSION>(rois + (size_t)r * 6, spatial_scale, sample_num, PH, PW);
  __syncthreads();
  const int spb = g.gh * g.gw;
  const bool tabled = (long long)nbins * spb <= kMaxSamples;
  SampleTap* taps = reinterpret_cast<SampleTap*>(smem);
  if (tabled) {
    for (int s = threadIdx.x; s < nbins * spb; s += blockDim.x) {
      const int bin = s / spb, k = s - bin * spb;
      taps[s] = make_tap<VERSION>(g, bin / PW, bin % PW, k / g.gw, k % g.gw, H, W);
    }
    __syncthreads();
  }
  const size_t plane = (size_t)H * W;
  float* gb = grad_in + (size_t)g.batch * C * plane;
  const float cnt = (float)spb;
  const int total = (c1 - c0) * nbins;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = c0 + i / nbins, bin = i % nbins;
    float* p = gb + (size_t)c * plane;
    const float go = __ldg(grad_out + ((size_t)r * C + c) * nbins + bin);
    for (int k = 0; k < spb; k++) {
      const SampleTap t = tabled ? taps[bin * spb + k] : make_tap<VERSION>(g, bin / PW, bin % PW, k / g.gw, k % g.gw, H, W);
      if (t.o00 < 0) continue;
      atomicAdd(p + t.o00, go * t.w1 / cnt);
      atomicAdd(p + t.o01, go * t.w2 / cnt);
      atomicAdd(p + t.o10, go * t.w3 / cnt);
      atomicAdd(p + t.o11, go * t.w4 / cnt);
    }
  }
}

// ---- configuration of the staged path (roi_gather_tma.cuh) ------------------------------------------------------
// slab = channels staged per pixel (128, or 64 when C % 128 != 0); ring = bytes of the staging ring, the largest
// power of two that leaves room for two output slabs and three table records in the 227 KB of an SM; pcap = the most
// pixels a whole-RoI chunk may hold (a quarter of the ring: chunks must be small against the ring, see roi_gather_tma.cuh).  ok == false: the shape does not fit (huge pooled grids) -> direct kernel.
struct StagedCfg { bool ok; int slab, pxbytes, pcap; unsigned ring; size_t smem, stride; };
// The TMA-staged gather (roi_gather_tma.cuh) is opt-in: JDET_ROI_TMA=1.  Measured on B200 (profiles/r02_roi_tma_*.txt,
// DESIGN.md section 4.3): it loses to the L1-path kernel on cfg2 — 46 us for the 1244 small RoIs it stages against ~30 us —
// so by default every RoI takes the LSU path; the staged path stays parity-tested (tests/test_gpu_parity.py).
static bool roi_tma_enabled() { const char* e = getenv("JDET_ROI_TMA"); return e && e[0] == '1'; }
static StagedCfg staged_cfg(int C, int PH, int PW, int sample_num) {
  StagedCfg c{};
  if (sample_num <= 0 || C % 64 != 0 || C > 64 * g4::kMaxSlabs) return c;
  const long long nbins_ll = (long long)PH * PW;
  if (nbins_ll * sample_num * sample_num > kMaxSamples) return c;
  const int nbins = (int)nbins_ll, tpb = 4 * sample_num * sample_num;
  c.slab = (C % 128 == 0) ? 128 : 64;
  c.pxbytes = c.slab * 4;
  c.stride = roi_tma_enabled() ? roi_table_stride(nbins, sample_num)      // record pitch: with / without the staged path's pixel lists
                               : (size_t)g4::rec_layout(nbins, roi_table_fstride(sample_num), tpb).cnpx_off;
  const size_t fixed = 2 * ((((size_t)c.slab * (nbins | 1) + 3) & ~(size_t)3) * 4) + g4::kTB * c.stride + 2560;   // + the kernel's static shared memory
  for (unsigned ring = 128u << 10; ring >= (32u << 10); ring >>= 1) {
    if (fixed + ring > 227u * 1024) continue;
    const int pcap = (int)(ring / 2 / c.pxbytes);       // a staged RoI is one chunk of at most half the ring
    if ((int)(ring / 2 / c.pxbytes) < ((tpb + 3) & ~3)) break;   // one bin must always fit a chunk
    c.ok = true; c.ring = ring; c.pcap = pcap; c.smem = fixed - 2560 + ring;
    break;
  }
  return c;
}

static bool use_staged(int B, int C, int H, int W, int R, int PH, int PW, int sample_num) {
  if (!staged_cfg(C, PH, PW, sample_num).ok) return false;
  if (H >= 32768 || W >= 32768 || (long long)B * H * W >= (1ll << 31)) return false;   // packed (y, x) taps, 32-bit gather rows
  // re-laying the map moves 8*B*C*H*W bytes; it pays once the RoI set samples the map densely
  return (long long)R * PH * PW * sample_num * sample_num * 8 >= (long long)B * H * W;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (TensorMapEncodeFn)p;
  }();
  return fn;
}
// (B*H*W, C) fp32 view of a channel-last map; one gather row = one pixel, box = {slab channels, 1 row}
static cudaError_t make_pixel_map(CUtensorMap* map, const float* nhwc, long long pixels, int C, int slab) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return cudaErrorNotSupported;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)pixels};
  cuuint64_t gstr[1] = {(cuuint64_t)C * 4};
  cuuint32_t box[2] = {(cuuint32_t)slab, 1};
  cuuint32_t es[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(nhwc), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// one prologue launch: re-layout of ONE NCHW map (input_nchw may be null) and, if with_tables, the tap tables of all RoIs
static cudaError_t launch_prologue(int version, const float* input_nchw, float* nhwc_scratch, int B, int C, int H, int W,
                                   bool with_tables, unsigned char* tables, int* work_counter, const float* rois, int R, int PH,
                                   int PW, int sampling_ratio, const RoiLevels& L, cudaStream_t st) {
  const int nbins = PH * PW;
  const StagedCfg cfg = staged_cfg(C, PH, PW, sampling_ratio);
  if (!cfg.ok) return cudaErrorInvalidConfiguration;
  const int Rt = with_tables ? R : 0;
  for (int l = 0; l < L.n; l++) {  // taps are packed (y << 16) | x; gather rows are 32-bit pixel indices, LSU entries 32-bit byte offsets
    if ((unsigned long long)L.lv[l].H * L.lv[l].W * C >= (1ull << 30)) return cudaErrorInvalidConfiguration;
    if (L.lv[l].H >= 32768 || L.lv[l].W >= 32768 || (long long)B * L.lv[l].H * L.lv[l].W >= (1ll << 31)) return cudaErrorInvalidConfiguration;
  }
  const bool vec = input_nchw && ((H * W) & 3) == 0 && (C & 3) == 0 && ((((uintptr_t)input_nchw) | ((uintptr_t)nhwc_scratch)) & 15) == 0;
  const int tiles_x = input_nchw ? jdet_ceil_div(H * W, vec ? kTileW : 32) : 0, tiles_y = input_nchw ? jdet_ceil_div(C, 32) : 0;
  const int rows = input_nchw ? tiles_y * B : jdet_ceil_div(Rt, 1024);
  if (rows > 65535) return cudaErrorInvalidConfiguration;
  if (rows == 0) return cudaSuccess;
  dim3 pgrid(tiles_x + jdet_ceil_div(Rt, rows), rows);
  const size_t smem = std::max(cfg.stride + (size_t)nbins * sampling_ratio * sampling_ratio * 4 * sizeof(int2) + sizeof(g4::ChunkScratch) + 16,
                               sizeof(float) * 32 * (kTileW + 1));
  const bool tma = roi_tma_enabled();
#define JDET_LAUNCH_PRO(V, S)                                                                                          \
  do {                                                                                                                 \
    if (smem > 48 * 1024) { cudaError_t e_ = cudaFuncSetAttribute(roi_prologue_kernel<V, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e_ != cudaSuccess) return e_; } \
    roi_prologue_kernel<V, S><<<pgrid, 256, smem, st>>>(input_nchw, nhwc_scratch, C, H, W, tiles_x, tiles_y, rois, (unsigned)Rt, PH, PW, sampling_ratio, tables, vec, work_counter, tma ? cfg.pcap : 0, cfg.pxbytes, (int)cfg.stride, L); \
  } while (0)
  if (version == 1) { if (tma) JDET_LAUNCH_PRO(1, true); else JDET_LAUNCH_PRO(1, false); }
  else              { if (tma) JDET_LAUNCH_PRO(0, true); else JDET_LAUNCH_PRO(0, false); }
#undef JDET_LAUNCH_PRO
  return cudaGetLastError();
}

// two gather launches over the tables of one prologue: the TMA-staged RoIs, then the rest on the LSU path.  Both read their
// work lists (and how many RoIs they hold) from device memory; a launch without work exits at once.
static cudaError_t launch_gather(const RoiLevels& L, int B, const unsigned char* tables, int* work_counter, int C, int R, int PH, int PW,
                                 int sampling_ratio, float* output, cudaStream_t st) {
  const int nbins = PH * PW;
  const StagedCfg cfg = staged_cfg(C, PH, PW, sampling_ratio);
  if (!cfg.ok) return cudaErrorInvalidConfiguration;
  const int nslabs = C / cfg.slab;
  if (nslabs > g4::kMaxSlabs) return cudaErrorInvalidConfiguration;
  if ((long long)R * nslabs > 0x7fffffffLL - 148 * 8) return cudaErrorInvalidConfiguration;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int fstride = roi_table_fstride(sampling_ratio), tpb = 4 * sampling_ratio * sampling_ratio;
  const g4::RecLayout RL = g4::rec_layout(nbins, fstride, tpb);
  // ---- staged RoIs: cp.async.bulk.tensor gather4 into the ring, gather from shared memory
  if (roi_tma_enabled()) {
  g4::GatherMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int l = 0; l < L.n; l++) {
    cudaError_t e = make_pixel_map(&maps.map[l], L.lv[l].nhwc, (long long)B * L.lv[l].H * L.lv[l].W, C, cfg.slab);
    if (e != cudaSuccess) return e;
  }
  g4::GatherArgs a{};
  a.tables = tables; a.rec_bytes = work_counter + 64 + (size_t)(kRoiBuckets + 1) * R; a.stride = cfg.stride;
  a.work_counter = work_counter; a.out = output;
  a.C = C; a.nbins = nbins; a.nslabs = nslabs; a.items = R;
  a.fstride = fstride; a.tpb = tpb;
  a.ring = cfg.ring;
#ifdef JDET_G4_STATS
  a.stats = g_jdet_g4_stats;
#endif
  const int grid = std::min(R, sms);
  if (cfg.slab == 128) {
    cudaError_t e_ = cudaFuncSetAttribute(g4::roi_gather4_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem); if (e_ != cudaSuccess) return e_;
    g4::roi_gather4_kernel<128><<<grid, g4::kThreads, cfg.smem, st>>>(maps, a);
  } else {
    cudaError_t e_ = cudaFuncSetAttribute(g4::roi_gather4_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem); if (e_ != cudaSuccess) return e_;
    g4::roi_gather4_kernel<64><<<grid, g4::kThreads, cfg.smem, st>>>(maps, a);
  }
  { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; }
  }
  // ---- the other RoIs: L1-path gathers straight from the channel-last map (round-1 kernel, now with packed FFMA2)
  const uint32_t rec_bytes = (uint32_t)RL.cnpx_off;                  // header, cnt, fin
  const size_t smem = (((size_t)cfg.slab * (nbins | 1) + 3) & ~(size_t)3) * 4 + 2 * (size_t)rec_bytes;
  const bool pair = cfg.slab == 128;   // two bins per warp (A/B on one box, cfg2 whole op: 116.8 vs 122.1 us)
  const int lgrid = (int)std::min<long long>((long long)R * nslabs, (long long)sms * (pair ? JDET_ROI_PAIR_MINB : 4));
#define JDET_LAUNCH_ROI(QL_, PAIR_)                                                                                    \
  do {                                                                                                                 \
    if (smem > 48 * 1024) { cudaError_t e_ = cudaFuncSetAttribute(roi_gather_kernel<QL_, PAIR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e_ != cudaSuccess) return e_; } \
    roi_gather_kernel<QL_, PAIR_><<<lgrid, 256, smem, st>>>(L, tables, cfg.stride, C, nbins, nslabs, R, rec_bytes, RL.fin_off, roi_tma_enabled() ? -1 : R * nslabs, work_counter, output); \
  } while (0)
  if (pair) JDET_LAUNCH_ROI(16, true); else JDET_LAUNCH_ROI(16, false);
#undef JDET_LAUNCH_ROI
  return cudaGetLastError();
}

// single map: prologue (tables, + re-layout when `input_nchw` is given) and gather from the channel-last map
static cudaError_t launch_staged(int version, const float* input_nchw, const float* nhwc_in, float* nhwc_scratch,
                                 unsigned char* tables, const float* rois, int B, int C, int H, int W, int R, int PH, int PW,
                                 float spatial_scale, int sampling_ratio, float* output, cudaStream_t st) {
  const size_t stride = roi_table_stride(PH * PW, sampling_ratio);
  int* work_counter = reinterpret_cast<int*>(tables + jdet_align_up((size_t)R * stride, 256));
  { cudaError_t e0 = cudaMemsetAsync(work_counter, 0, 256, st); if (e0 != cudaSuccess) return e0; }   // item counter + bucket counts
  RoiLevels L{};
  L.n = 1;
  L.lv[0] = RoiLevel{input_nchw ? nhwc_scratch : nhwc_in, H, W, spatial_scale};
  cudaError_t e = launch_prologue(version, input_nchw, nhwc_scratch, B, C, H, W, true, tables, work_counter, rois, R, PH, PW,
                                  sampling_ratio, L, st);
  if (e != cudaSuccess) return e;
  return launch_gather(L, B, tables, work_counter, C, R, PH, PW, sampling_ratio, output, st);
}

}  // namespace jdet

JDET_API size_t jdet_roi_align_rotated_workspace_bytes(int B, int C, int H, int W, int R, int PH, int PW,
                                                       int sampling_ratio) {
  if (!jdet::use_staged(B, C, H, W, R, PH, PW, sampling_ratio)) return 256;
  return jdet_align_up((size_t)B * C * H * W * sizeof(float), 256) +          // channel-last copy
         jdet_align_up((size_t)R * jdet::roi_table_stride(PH * PW, sampling_ratio), 256) + jdet::roi_queue_bytes(R);   // tap tables + work queue
}

JDET_API size_t jdet_roi_align_rotated_nhwc_workspace_bytes(int R, int PH, int PW, int sampling_ratio) {
  if (sampling_ratio <= 0 || R <= 0 || PH <= 0 || PW <= 0) return 256;
  return jdet_align_up((size_t)R * jdet::roi_table_stride(PH * PW, sampling_ratio), 256) + jdet::roi_queue_bytes(R);
}

// version 1: ROIAlignRotated_v1 / roi_align (ops/roi_align_rotated_v1.py:300-326,355-365)
// version 0: ROIAlignRotated    / roi_align (ops/roi_align_rotated.py:257-283,312-322)
// input (B,C,H,W), rois (R,6), output (R,C,PH,PW): device fp32 contiguous.  sampling_ratio is the
// integer the reference's int parameter receives (Python side truncates the float like C does).
JDET_API int jdet_roi_align_rotated(int version, const float* input, int B, int C, int H, int W, const float* rois,
                                    int R, int PH, int PW, float spatial_scale, int sampling_ratio, float* output,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  using namespace jdet;
  if ((version != 0 && version != 1) || B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || PH <= 0 || PW <= 0)
    return JDET_ERR_BAD_ARG;
  if (R == 0 || C == 0) return 0;
  if (!input || !rois || !output || B == 0 || H == 0 || W == 0) return JDET_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_staged(B, C, H, W, R, PH, PW, sampling_ratio)) {
    const size_t map_bytes = jdet_align_up((size_t)B * C * H * W * sizeof(float), 256);
    if (!workspace || workspace_bytes < jdet_roi_align_rotated_workspace_bytes(B, C, H, W, R, PH, PW, sampling_ratio))
      return JDET_ERR_WORKSPACE;
    JDET_RETURN_IF_CUDA(launch_staged(version, input, nullptr, (float*)workspace, (unsigned char*)workspace + map_bytes, rois, B, C,
                                      H, W, R, PH, PW, spatial_scale, sampling_ratio, output, st));
  } else {
    const int ch_per_cta = 32;
    const size_t smem = (size_t)kMaxSamples * sizeof(SampleTap);
    dim3 grid(R, jdet_ceil_div(C, ch_per_cta));
    if (version == 1)
      roi_align_nchw_kernel<1><<<grid, 256, smem, st>>>(input, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, ch_per_cta, output);
    else
      roi_align_nchw_kernel<0><<<grid, 256, smem, st>>>(input, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, ch_per_cta, output);
  }
  return (int)cudaGetLastError();
}

// channel-last input: no re-layout (include/jdet_b200.h)
JDET_API int jdet_roi_align_rotated_nhwc(int version, const float* input_nhwc, int B, int C, int H, int W, const float* rois,
                                         int R, int PH, int PW, float spatial_scale, int sampling_ratio, float* output,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  using namespace jdet;
  if ((version != 0 && version != 1) || B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || PH <= 0 || PW <= 0)
    return JDET_ERR_BAD_ARG;
  if (R == 0 || C == 0) return 0;
  if (!input_nhwc || !rois || !output || B == 0 || H == 0 || W == 0) return JDET_ERR_BAD_ARG;
  if (!staged_cfg(C, PH, PW, sampling_ratio).ok) return JDET_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < jdet_roi_align_rotated_nhwc_workspace_bytes(R, PH, PW, sampling_ratio))
    return JDET_ERR_WORKSPACE;
  return (int)launch_staged(version, nullptr, input_nhwc, nullptr, (unsigned char*)workspace, rois, B, C, H, W, R, PH, PW,
                            spatial_scale, sampling_ratio, output, (cudaStream_t)stream);
}

// Fused single-level-per-RoI extractor over FPN maps (OrientedSingleRoIExtractor / RboxSingleRoIExtractor.execute,
// models/roi_extractors/oriented_single_level.py:91-114, rbox_single_level.py): level choice on the device, one
// re-layout launch per level (the tap tables ride in the first), ONE gather over all RoIs writing final rows —
// no boolean masks, no zero-filled output, no scatter-add.  feats/Hs/Ws/scales are HOST arrays of nlevels entries.
JDET_API size_t jdet_roi_align_rotated_fpn_workspace_bytes(int nlevels, int B, int C, const int* Hs, const int* Ws, int R, int PH,
                                                           int PW, int sampling_ratio) {
  if (nlevels <= 0 || !Hs || !Ws || sampling_ratio <= 0 || PH <= 0 || PW <= 0) return 256;
  size_t total = 0;
  for (int l = 0; l < nlevels; l++) total += jdet_align_up((size_t)B * C * Hs[l] * Ws[l] * sizeof(float), 256);
  return total + jdet_align_up((size_t)(R > 0 ? R : 0) * jdet::roi_table_stride(PH * PW, sampling_ratio), 256) + jdet::roi_queue_bytes(R);
}

JDET_API int jdet_roi_align_rotated_fpn(int version, const float* const* feats, int nlevels, int B, int C, const int* Hs,
                                        const int* Ws, const float* scales, const float* rois, int R, int PH, int PW,
                                        int sampling_ratio, float ext_w, float ext_h, float rs_w, float rs_h, float finest_scale,
                                        float* output, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace jdet;
  if ((version != 0 && version != 1) || nlevels <= 0 || nlevels > kMaxLevels || !feats || !Hs || !Ws || !scales || B <= 0 || C <= 0 ||
      R < 0 || PH <= 0 || PW <= 0)
    return JDET_ERR_BAD_ARG;
  if (R == 0) return 0;
  if (!rois || !output) return JDET_ERR_BAD_ARG;
  if (!staged_cfg(C, PH, PW, sampling_ratio).ok) return JDET_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < jdet_roi_align_rotated_fpn_workspace_bytes(nlevels, B, C, Hs, Ws, R, PH, PW, sampling_ratio))
    return JDET_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  RoiLevels L{};
  L.n = nlevels; L.ext_w = ext_w; L.ext_h = ext_h; L.rs_w = rs_w; L.rs_h = rs_h; L.finest = finest_scale;
  unsigned char* p = (unsigned char*)workspace;
  for (int l = 0; l < nlevels; l++) {
    if (!feats[l] || Hs[l] <= 0 || Ws[l] <= 0) return JDET_ERR_BAD_ARG;
    L.lv[l] = RoiLevel{(const float*)p, Hs[l], Ws[l], scales[l]};
    p += jdet_align_up((size_t)B * C * Hs[l] * Ws[l] * sizeof(float), 256);
  }
  unsigned char* tables = p;
  int* work_counter = reinterpret_cast<int*>(tables + jdet_align_up((size_t)R * roi_table_stride(PH * PW, sampling_ratio), 256));
  JDET_RETURN_IF_CUDA(cudaMemsetAsync(work_counter, 0, 256, st));   // item counter + bucket counts
  for (int l = 0; l < nlevels; l++)
    JDET_RETURN_IF_CUDA(launch_prologue(version, feats[l], const_cast<float*>(L.lv[l].nhwc), B, C, Hs[l], Ws[l], l == 0, tables,
                                        work_counter, rois, R, PH, PW, sampling_ratio, L, st));
  JDET_RETURN_IF_CUDA(launch_gather(L, B, tables, work_counter, C, R, PH, PW, sampling_ratio, output, st));
  return 0;
}

// backward of jdet_roi_align_rotated w.r.t. input: _RotatedROIAlign[_v1].grad (roi_align_rotated_v1.py:328-351,
// roi_align_rotated.py:285-308).  grad_output (R,C,PH,PW) -> grad_input (B,C,H,W), written in full.
JDET_API size_t jdet_roi_align_rotated_backward_workspace_bytes(int B, int C, int H, int W, int R, int PH, int PW,
                                                                int sampling_ratio) {
  if (!jdet::use_staged(B, C, H, W, R, PH, PW, sampling_ratio)) return 256;
  return jdet_align_up((size_t)B * C * H * W * sizeof(float), 256);
}

JDET_API int jdet_roi_align_rotated_backward(int version, const float* grad_output, const float* rois, int R, int B, int C,
                                             int H, int W, int PH, int PW, float spatial_scale, int sampling_ratio,
                                             float* grad_input, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace jdet;
  if ((version != 0 && version != 1) || B < 0 || C < 0 || H < 0 || W < 0 || R < 0 || PH <= 0 || PW <= 0)
    return JDET_ERR_BAD_ARG;
  const size_t in_bytes = (size_t)B * C * H * W * sizeof(float);
  if (in_bytes == 0) return 0;
  if (!grad_input || (R > 0 && (!grad_output || !rois))) return JDET_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int nbins = PH * PW;
  if (R == 0) { JDET_RETURN_IF_CUDA(cudaMemsetAsync(grad_input, 0, in_bytes, st)); return 0; }
  if (use_staged(B, C, H, W, R, PH, PW, sampling_ratio)) {
    if (!workspace || workspace_bytes < jdet_align_up(in_bytes, 256)) return JDET_ERR_WORKSPACE;
    float* nhwc = (float*)workspace;
    JDET_RETURN_IF_CUDA(cudaMemsetAsync(nhwc, 0, in_bytes, st));
    const int slab = (C % 128 == 0) ? 128 : 64;
    const size_t smem = 2 * (size_t)nbins * sampling_ratio * sampling_ratio * 4 * sizeof(int2) + (size_t)slab * nbins * 4 + (size_t)nbins * 4;
    dim3 grid(R, C / slab);
#define JDET_LAUNCH_BWD(V, S)                                                                                          \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(roi_align_bwd_nhwc_kernel<V, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    roi_align_bwd_nhwc_kernel<V, S><<<grid, 256, smem, st>>>(grad_output, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, nhwc);   \
  } while (0)
    if (version == 1) { if (slab == 128) JDET_LAUNCH_BWD(1, 128); else JDET_LAUNCH_BWD(1, 64); }
    else              { if (slab == 128) JDET_LAUNCH_BWD(0, 128); else JDET_LAUNCH_BWD(0, 64); }
#undef JDET_LAUNCH_BWD
    launch_nchw_to_nhwc(nhwc, grad_input, B, H * W, C, st);   // (B, HW, C) -> (B, C, HW): same tiled transpose, roles swapped
  } else {
    JDET_RETURN_IF_CUDA(cudaMemsetAsync(grad_input, 0, in_bytes, st));
    const int ch_per_cta = 32;
    const size_t smem = (size_t)kMaxSamples * sizeof(SampleTap);
    dim3 grid(R, jdet_ceil_div(C, ch_per_cta));
    if (version == 1)
      roi_align_bwd_nchw_kernel<1><<<grid, 256, smem, st>>>(grad_output, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, ch_per_cta, grad_input);
    else
      roi_align_bwd_nchw_kernel<0><<<grid, 256, smem, st>>>(grad_output, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, ch_per_cta, grad_input);
  }
  return (int)cudaGetLastError();
}
