// feature_refine.cu — rotated_feature_align ("feature_refine" in JDet) forward for sm_100a.
//
// Replaces feature_refine_forward_kernel + launch (/root/reference/python/jdet/ops/fr.py:114-165,
// 234-240):  out[n,c,h,w] = in[n,c,h,w] + sum_{i<points} bilinear(in[n,c], py_i, px_i),
// points in {1,5}; bbox (n,h,w,5): bbox[0]*scale is the ROW, bbox[1]*scale the COLUMN (:133-134).
//
// Reference shape of work: one thread per (n,c,h,w): the box is re-read and the 1/5 sample points
// (cosf/sinf included) re-derived for every channel, 256x redundantly at the bench shape.
// Here a thread owns one pixel: it decodes its box once into taps+weights held in registers, then
// walks a slab of channels.  Consecutive lanes are consecutive pixels, so the centre read and the
// store are 128-B coalesced and the taps of neighbouring pixels fall in the same few lines of the
// plane (L1 hits).  HBM traffic ~ one read + one write of the feature map: an HBM-bound stream.
//
// Layout in HBM: features (N,C,H,W) fp32, boxes (N,H,W,5) fp32, out (N,C,H,W) fp32.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace jdet {

struct Tap4 {
  int o00, o01, o10, o11;
  float w1, w2, w3, w4;
};

// fr.py:18-67: the sample's cell, its lerp fractions and whether it contributes at all
struct TapGeo { int yl, xl, yh, xh; float ly, lx; bool valid; };
__device__ __forceinline__ TapGeo fr_tap_geo(float y, float x, int H, int W) {
  TapGeo g;
  g.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W || !(y == y) || !(x == x));   // fr.py:24-26: else exactly 0
  g.yl = g.xl = g.yh = g.xh = 0; g.ly = g.lx = 0.f;
  if (!g.valid) return g;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  g.yl = yl; g.xl = xl; g.yh = yh; g.xh = xh;
  g.ly = y - (float)yl; g.lx = x - (float)xl;
  return g;
}
__device__ __forceinline__ Tap4 fr_tap(float y, float x, int H, int W) {
  Tap4 t;
  const TapGeo g = fr_tap_geo(y, x, H, W);
  if (!g.valid) {
    t.o00 = -1; t.o01 = t.o10 = t.o11 = 0;          // sample contributes exactly 0 (fr.py:24-26)
    t.w1 = t.w2 = t.w3 = t.w4 = 0.f;
    return t;
  }
  const float hy = 1.f - g.ly, hx = 1.f - g.lx;
  t.o00 = g.yl * W + g.xl; t.o01 = g.yl * W + g.xh; t.o10 = g.yh * W + g.xl; t.o11 = g.yh * W + g.xh;
  t.w1 = __fmul_rn(hy, hx); t.w2 = __fmul_rn(hy, g.lx); t.w3 = __fmul_rn(g.ly, hx); t.w4 = __fmul_rn(g.ly, g.lx);
  return t;
}

// grid = (pixel tiles of 256, channel slabs, N)
template <int POINTS>
__global__ void __launch_bounds__(256, 3) feature_refine_kernel(const float* __restrict__ feat, const float* __restrict__ boxes,
                                                              int C, int H, int W, float spatial_scale, int ch_per_cta,
                                                              float* __restrict__ out) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  if (p >= HW) return;
  const float* bb = boxes + ((size_t)n * HW + p) * 5;
  const float roi_y = __fmul_rn(bb[0], spatial_scale), roi_x = __fmul_rn(bb[1], spatial_scale);
  Tap4 taps[POINTS];
  taps[0] = fr_tap(roi_y, roi_x, H, W);
  if (POINTS > 1) {   // fr.py:139-153
    const float rw = __fmul_rn(bb[2], spatial_scale), rh = __fmul_rn(bb[3], spatial_scale), ra = bb[4];
    const float w2 = rw * 0.5f, h2 = rh * 0.5f;
    const float ca = cosf(ra), sa = sinf(ra);
    const float wx = __fmul_rn(ca, w2), wy = __fmul_rn(sa, w2), hx = __fmul_rn(-sa, h2), hy = __fmul_rn(ca, h2);
    taps[1 % POINTS] = fr_tap(__fadd_rn(__fadd_rn(roi_y, wy), hy), __fadd_rn(__fadd_rn(roi_x, wx), hx), H, W);
    taps[2 % POINTS] = fr_tap(__fadd_rn(__fsub_rn(roi_y, wy), hy), __fadd_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[3 % POINTS] = fr_tap(__fsub_rn(__fsub_rn(roi_y, wy), hy), __fsub_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[4 % POINTS] = fr_tap(__fsub_rn(__fadd_rn(roi_y, wy), hy), __fsub_rn(__fadd_rn(roi_x, wx), hx), H, W);
  }
  const float* plane = feat + ((size_t)n * C + c0) * HW;
  float* dst = out + ((size_t)n * C + c0) * HW + p;
#pragma unroll 2
  for (int c = c0; c < c1; c++) {
    float v = __ldg(plane + p);
#pragma unroll
    for (int i = 0; i < POINTS; i++) {
      const Tap4& t = taps[i];
      if (t.o00 >= 0)
        v += t.w1 * __ldg(plane + t.o00) + t.w2 * __ldg(plane + t.o01) + t.w3 * __ldg(plane + t.o10) +
           t.w4 * __ldg(plane + t.o11);
    }
    st_stream(dst, v);
    plane += HW;
    dst += HW;
  }
}

// points == 1, W % 4 == 0: a thread owns 4 consecutive pixels, so the centre read and the store are
// 16-B vectors (4x the bytes in flight per load instruction: the op is an HBM stream whose speed is
// set by memory-level parallelism); the 4 x 4 taps stay scalar L1 gathers.
__global__ void __launch_bounds__(256) feature_refine_p1_vec4_kernel(const float* __restrict__ feat,
                                                                     const float* __restrict__ boxes, int C, int H, int W,
                                                                     float spatial_scale, int ch_per_cta,
                                                                     float* __restrict__ out) {
  const int HW = H * W;
  const int p = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  if (p >= HW) return;
  Tap4 t[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float* bb = boxes + ((size_t)n * HW + p + j) * 5;
    t[j] = fr_tap(__fmul_rn(__ldg(bb), spatial_scale), __fmul_rn(__ldg(bb + 1), spatial_scale), H, W);
  }
  const float* plane = feat + ((size_t)n * C + c0) * HW;
  float* dst = out + ((size_t)n * C + c0) * HW + p;
#pragma unroll 2
  for (int c = c0; c < c1; c++) {
    const float4 ctr = __ldg(reinterpret_cast<const float4*>(plane + p));
    float v[4] = {ctr.x, ctr.y, ctr.z, ctr.w};
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (t[j].o00 >= 0)
        v[j] += t[j].w1 * __ldg(plane + t[j].o00) + t[j].w2 * __ldg(plane + t[j].o01) + t[j].w3 * __ldg(plane + t[j].o10) +
                t[j].w4 * __ldg(plane + t[j].o11);
    st_stream_v4(dst, v[0], v[1], v[2], v[3]);
    plane += HW;
    dst += HW;
  }
}

// backward (fr.py:167-232): grad_in = grad_out + scatter(grad_out * w) — thread = pixel, channel walk.
// grad_in must already hold a copy of grad_out (the identity term); taps are float atomics as in the reference.
template <int POINTS>
__global__ void __launch_bounds__(256, 2) feature_refine_bwd_kernel(const float* __restrict__ grad_out,
                                                                     const float* __restrict__ boxes, int C, int H, int W,
                                                                     float spatial_scale, int ch_per_cta,
                                                                     float* __restrict__ grad_in) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  if (p >= HW) return;
  const float* bb = boxes + ((size_t)n * HW + p) * 5;
  const float roi_y = __fmul_rn(bb[0], spatial_scale), roi_x = __fmul_rn(bb[1], spatial_scale);
  Tap4 taps[POINTS];
  taps[0] = fr_tap(roi_y, roi_x, H, W);
  if (POINTS > 1) {
    const float rw = __fmul_rn(bb[2], spatial_scale), rh = __fmul_rn(bb[3], spatial_scale), ra = bb[4];
    const float w2 = rw * 0.5f, h2 = rh * 0.5f;
    const float ca = cosf(ra), sa = sinf(ra);
    const float wx = __fmul_rn(ca, w2), wy = __fmul_rn(sa, w2), hx = __fmul_rn(-sa, h2), hy = __fmul_rn(ca, h2);
    taps[1 % POINTS] = fr_tap(__fadd_rn(__fadd_rn(roi_y, wy), hy), __fadd_rn(__fadd_rn(roi_x, wx), hx), H, W);
    taps[2 % POINTS] = fr_tap(__fadd_rn(__fsub_rn(roi_y, wy), hy), __fadd_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[3 % POINTS] = fr_tap(__fsub_rn(__fsub_rn(roi_y, wy), hy), __fsub_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[4 % POINTS] = fr_tap(__fsub_rn(__fadd_rn(roi_y, wy), hy), __fsub_rn(__fadd_rn(roi_x, wx), hx), H, W);
  }
  const float* src = grad_out + ((size_t)n * C + c0) * HW + p;
  float* plane = grad_in + ((size_t)n * C + c0) * HW;
  for (int c = c0; c < c1; c++) {
    const float g = __ldg(src);
#pragma unroll
    for (int i = 0; i < POINTS; i++) {
      const Tap4& t = taps[i];
      if (t.o00 >= 0) {
        atomicAdd(plane + t.o00, g * t.w1);
        atomicAdd(plane + t.o01, g * t.w2);
        atomicAdd(plane + t.o10, g * t.w3);
        atomicAdd(plane + t.o11, g * t.w4);
      }
    }
    src += HW;
    plane += HW;
  }
}

// ---- TMA-staged kernel --------------------------------------------------------------------------
// Full-width row bands of an NCHW plane are contiguous, so the band a CTA works on (its rows plus kHalo
// rows above and below) is ONE cp.async.bulk (UBLKCP) per channel into a shared-memory ring guarded by
// mbarriers: kStages channels are in flight per CTA with no registers tied up, which is what an HBM
// stream with a short gather needs.  Centre reads and taps are then served from shared memory (lanes are
// consecutive pixels: conflict-light); a sample whose taps leave the band falls back to global loads.
// CTA = (row band, channel slab, image), one thread per pixel of the band.
// the 1 or 5 sample points of a box (fr.py:128-153), decoded into taps
template <int POINTS>
__device__ __forceinline__ void fr_points(const float* __restrict__ bb, float spatial_scale, float (&py)[POINTS], float (&px)[POINTS]) {
  const float roi_y = __fmul_rn(__ldg(bb), spatial_scale), roi_x = __fmul_rn(__ldg(bb + 1), spatial_scale);
  py[0] = roi_y; px[0] = roi_x;
  if (POINTS > 1) {   // fr.py:139-153
    const float rw = __fmul_rn(__ldg(bb + 2), spatial_scale), rh = __fmul_rn(__ldg(bb + 3), spatial_scale), ra = __ldg(bb + 4);
    const float w2 = rw * 0.5f, h2 = rh * 0.5f;
    const float ca = cosf(ra), sa = sinf(ra);
    const float wx = __fmul_rn(ca, w2), wy = __fmul_rn(sa, w2), hx = __fmul_rn(-sa, h2), hy = __fmul_rn(ca, h2);
    py[1 % POINTS] = __fadd_rn(__fadd_rn(roi_y, wy), hy); px[1 % POINTS] = __fadd_rn(__fadd_rn(roi_x, wx), hx);
    py[2 % POINTS] = __fadd_rn(__fsub_rn(roi_y, wy), hy); px[2 % POINTS] = __fadd_rn(__fsub_rn(roi_x, wx), hx);
    py[3 % POINTS] = __fsub_rn(__fsub_rn(roi_y, wy), hy); px[3 % POINTS] = __fsub_rn(__fsub_rn(roi_x, wx), hx);
    py[4 % POINTS] = __fsub_rn(__fadd_rn(roi_y, wy), hy); px[4 % POINTS] = __fsub_rn(__fadd_rn(roi_x, wx), hx);
  }
}
template <int POINTS>
__device__ __forceinline__ void fr_decode(const float* __restrict__ bb, float spatial_scale, int H, int W, Tap4 (&taps)[POINTS]) {
  float py[POINTS], px[POINTS];
  fr_points<POINTS>(bb, spatial_scale, py, px);
#pragma unroll
  for (int k = 0; k < POINTS; k++) taps[k] = fr_tap(py[k], px[k], H, W);
}

namespace fr_tma {

constexpr int kMaxStages = 8;

// the geometry of one instantiation: CW consumer warps, PPT pixels per consumer thread, HALO staged rows either side of the band
template <int POINTS> struct Geo;
#ifndef JDET_FR1_CW          // A/B switches, as for points = 5 below
#define JDET_FR1_CW 8
#endif
#ifndef JDET_FR1_PPT
#define JDET_FR1_PPT 4
#endif
#ifndef JDET_FR1_HALO
#define JDET_FR1_HALO 4
#endif
#ifndef JDET_FR1_MINB
#define JDET_FR1_MINB 3      // 3 CTAs per SM: 72 registers
#endif
#ifndef JDET_FR1_SMEM_KB
#define JDET_FR1_SMEM_KB 64
#endif
template <> struct Geo<1> { static constexpr int CW = JDET_FR1_CW, PPT = JDET_FR1_PPT, HALO = JDET_FR1_HALO, MINB = JDET_FR1_MINB; };
// points = 5: HALO bounds the stage (rows_per_band + 2 * HALO rows); the rows actually staged are those the band's samples read
// (fr_tma_body<5>: 21 of the 32-row capacity on average at cfg4's 128 x 128 level, whose 4-cell anchors' corners reach +-9 rows)
#ifndef JDET_FR5_CW          // A/B switches (tools/ab_libs.py build NAME=flags:-DJDET_FR5_CW=8 -DJDET_FR5_PPT=4 ...)
#define JDET_FR5_CW 16
#endif
#ifndef JDET_FR5_PPT
#define JDET_FR5_PPT 2
#endif
#ifndef JDET_FR5_HALO
#define JDET_FR5_HALO 12
#endif
#ifndef JDET_FR5_SMEM_KB
#define JDET_FR5_SMEM_KB 104
#endif
#ifndef JDET_FR5_MINB
#define JDET_FR5_MINB 1
#endif
template <> struct Geo<5> { static constexpr int CW = JDET_FR5_CW, PPT = JDET_FR5_PPT, HALO = JDET_FR5_HALO, MINB = JDET_FR5_MINB; };
// floats per stage: the band, then a zero pad of one row + 8 words that no copy ever writes: a sample outside the map reads its
// 2 x 2 taps there (see fr_tma_body)
template <int POINTS>
__host__ __device__ constexpr int fr_zero_pad(int W) { return W + 8; }
template <int POINTS>
__host__ __device__ constexpr int fr_stage_elems(int rows_per_band, int W) {
  return (rows_per_band + 2 * Geo<POINTS>::HALO) * W + fr_zero_pad<POINTS>(W);
}

// CTA = CW consumer warps + 1 producer warp; work item = (row band, channel chunk, image).  A band is CW*32*PPT / W full-width
// rows plus the rows above and below that its samples read (at most HALO either side): contiguous in an NCHW plane, so ONE
// cp.async.bulk per channel stages it.  The producer lane only waits on empty[] and issues copies; the consumers never block
// on a refill, and `stages` channels per CTA x the CTAs of the SM are in flight.  Each consumer thread owns PPT pixels (rows
// CW*32/W apart) and keeps their samples in registers for the whole channel walk (fr_tma_body).
// shared-memory ring of one CTA: `stages` bands of stage_elems floats, then the full[] / empty[] barriers and two words
// (the first / last row the band's samples touch)
template <int POINTS>
struct FrRing {
  float* ring; uint64_t* full; uint64_t* empty; int* span;
  int lo, hi, zoff, stage_elems;                                                 // staged rows [lo, hi)
  __device__ __forceinline__ FrRing(unsigned char* smem, int H, int W, int r0, int rows_per_band, int stages) {
    constexpr int kHalo = Geo<POINTS>::HALO;
    lo = max(0, r0 - kHalo); hi = min(H, r0 + rows_per_band + kHalo);
    zoff = (rows_per_band + 2 * kHalo) * W;                                      // the stage's zero pad
    stage_elems = fr_stage_elems<POINTS>(rows_per_band, W);
    ring = reinterpret_cast<float*>(smem);
    full = reinterpret_cast<uint64_t*>(ring + (size_t)stages * stage_elems);
    empty = full + kMaxStages;
    span = reinterpret_cast<int*>(empty + kMaxStages);
    const int tid = threadIdx.x;
    if (tid == 0) {
      for (int s = 0; s < stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], Geo<POINTS>::CW); }
      mbar_init_fence();
      span[0] = 0x7fffffff; span[1] = -1;
    }
    const int pad = fr_zero_pad<POINTS>(W);                                      // no copy ever writes it
    for (int i = tid; i < stages * pad; i += blockDim.x) ring[(size_t)(i / pad) * stage_elems + zoff + i % pad] = 0.f;
    __syncthreads();
  }
  // the producer lane: one bulk copy per channel, `stages` in flight
  __device__ __forceinline__ void produce(const float* plane0, int HW, int W, int nch, int stages) const {
    const uint32_t band_bytes = (uint32_t)((hi - lo) * W) * 4u;
    const float* src0 = plane0 + (size_t)lo * W;
    uint32_t stage = 0, phase = 0;
    for (int c = 0; c < nch; c++) {
      if (c >= stages) mbar_wait(&empty[stage], phase ^ 1);        // consumers released the previous occupant
      mbar_expect_tx(&full[stage], band_bytes);
      bulk_g2s(ring + (size_t)stage * stage_elems, src0 + (size_t)c * HW, band_bytes, &full[stage]);
      if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
    }
  }
};
constexpr size_t kRingTail = 2 * kMaxStages * sizeof(uint64_t) + 16;             // barriers + span

// points = 1, first form (kept for A/B, -DJDET_FR1_LEGACY): four tap offsets + four weights per pixel in registers, a
// staged-or-global branch per sample — 42 instructions per pixel-channel, 71 % of the issue slots at 4.3 TB/s
__device__ __forceinline__ void fr_tma_body1_legacy(const float* __restrict__ feat, const float* __restrict__ boxes, int C, int H, int W,
                                               float spatial_scale, int rows_per_band, int ch_per_cta, int stages,
                                               float* __restrict__ out, int band, int chunk, int n) {
  constexpr int POINTS = 1, PPT = Geo<1>::PPT, kConsumers = Geo<1>::CW * 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = H * W;
  const int c0 = chunk * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  const int r0 = band * rows_per_band;                             // first output row of this band
  const FrRing<1> R(smem_raw, H, W, r0, rows_per_band, stages);
  const int lo = R.lo, hi = R.hi, stage_elems = R.stage_elems;
  const float* ring = R.ring;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nch = c1 - c0;

  if (tid >= kConsumers) {                                         // ---- producer warp
    if (lane == 0) R.produce(feat + ((size_t)n * C + c0) * HW, HW, W, nch, stages);
    return;
  }

  // ---- consumers: this thread's pixels and their taps (band-relative offsets when every corner is staged)
  Tap4 taps[PPT][POINTS];
  int pc[PPT];                                                     // centre, band-relative; -1: pixel not in the map
  unsigned staged[PPT];                                            // bit i: sample i reads the smem band
  const int blo = lo * W, bhi = hi * W;
  const int band_px = min(rows_per_band, H - r0) * W;
#pragma unroll
  for (int j = 0; j < PPT; j++) {
    const int i = tid + j * kConsumers;                            // pixel index inside the band
    pc[j] = -1;
    staged[j] = 0u;
#pragma unroll
    for (int k = 0; k < POINTS; k++) taps[j][k].o00 = -1;
    if (i < band_px) {
      const int p = r0 * W + i;
      fr_decode<POINTS>(boxes + ((size_t)n * HW + p) * 5, spatial_scale, H, W, taps[j]);
#pragma unroll
      for (int k = 0; k < POINTS; k++) {
        Tap4& t = taps[j][k];
        if (t.o00 >= blo && t.o11 < bhi) {                         // o00 is the smallest, o11 the largest offset
          staged[j] |= 1u << k;
          t.o00 -= blo; t.o01 -= blo; t.o10 -= blo; t.o11 -= blo;
        }
      }
      pc[j] = p - blo;
    }
  }
  float* dst = out + ((size_t)n * C + c0) * HW + blo;
  const float* gplane = feat + ((size_t)n * C + c0) * HW;

  uint32_t stage = 0, phase = 0;
  for (int c = 0; c < nch; c++) {
    mbar_wait(&R.full[stage], phase);
    const float* sp = ring + (size_t)stage * stage_elems;
    float v[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j++) {
      v[j] = 0.f;
      if (pc[j] >= 0) {
        v[j] = sp[pc[j]];
#pragma unroll
        for (int k = 0; k < POINTS; k++) {
          const Tap4& t = taps[j][k];
          if (staged[j] >> k & 1u) v[j] += t.w1 * sp[t.o00] + t.w2 * sp[t.o01] + t.w3 * sp[t.o10] + t.w4 * sp[t.o11];
          else if (t.o00 >= 0)
            v[j] += t.w1 * __ldg(gplane + t.o00) + t.w2 * __ldg(gplane + t.o01) + t.w3 * __ldg(gplane + t.o10) + t.w4 * __ldg(gplane + t.o11);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PPT; j++)
      if (pc[j] >= 0) st_stream(dst + pc[j], v[j]);
    __syncwarp();                                                  // every lane issued its stores, so its smem reads returned
    if (lane == 0) mbar_arrive(&R.empty[stage]);
    dst += HW;
    gplane += HW;
    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
  }
}

// The consumer body for both point counts.  A sample is kept as ONE word (byte offset of its top-left tap inside the stage) plus
// its two lerp fractions; the 2 x 2 taps are read at +0, +4, +row, +row + 4 and the four weights formed as fr.py:57-63 forms them.
//   * Where the reference clamps the lower / right neighbour onto the tap itself (last row / column: fraction exactly 0) the cell
//     is moved one row / column back with fraction exactly 1 — the same two products in the same order, every read inside the
//     map.  A sample outside the map points at the stage's zero pad and contributes the reference's exact 0.
//   * The rows staged per band are those its samples touch (CTA-wide min / max before the first copy), up to the stage's
//     capacity; a thread with a sample beyond that (none at cfg4) reads its taps from global memory out of a local-memory table.
//   * points = 5 (21 shared-memory reads per pixel-channel) is bound by shared-memory wavefronts — the corners of independent
//     boxes land in arbitrary banks, ~2.5 wavefronts per read — so issue slots are plentiful and registers scarce: the weights
//     are RE-formed per channel (an empty asm stops the optimiser hoisting them: 15 registers per pixel instead of 40), which
//     lets a pixel's 21 reads issue back to back.  With 40 registers per pixel and a staged-or-global branch per sample the first
//     version of this path serialised read -> FMA -> read and ran at 1.14 ms where the L1 gather takes 0.68 ms; now 0.35 ms.
//   * points = 1 (5 reads) is an HBM stream; its weights and offsets are hoisted (8 registers per pixel), and the branch-free
//     body costs 15 instructions per pixel-channel where the first form (fr_tma_body1_legacy) cost 42: 0.087 -> 0.076 ms.
// (Finite features assumed: a non-finite value next to a clamped tap would turn its 0 weight into NaN.)
template <int POINTS>
__device__ __forceinline__ void fr_tma_body(const float* __restrict__ feat, const float* __restrict__ boxes, int C, int H, int W,
                                            float spatial_scale, int rows_per_band, int ch_per_cta, int stages,
                                            float* __restrict__ out, int band, int chunk, int n) {
#ifdef JDET_FR1_LEGACY
  if (POINTS == 1) { fr_tma_body1_legacy(feat, boxes, C, H, W, spatial_scale, rows_per_band, ch_per_cta, stages, out, band, chunk, n); return; }
#endif
  constexpr int PPT = Geo<POINTS>::PPT, kConsumers = Geo<POINTS>::CW * 32, kHalo = Geo<POINTS>::HALO;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = H * W;
  const int c0 = chunk * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  const int r0 = band * rows_per_band, r1 = min(H, r0 + rows_per_band);
  FrRing<POINTS> R(smem_raw, H, W, r0, rows_per_band, stages);
  const int tid = threadIdx.x, lane = tid & 31;
  const int nch = c1 - c0;
  const int band_px = (r1 - r0) * W;

  // ---- every sample of the band: its cell and fractions; the rows the band touches
  uint32_t ta[PPT][POINTS];                                        // top-left tap: map offset, then byte offset in the stage
  float tly[PPT][POINTS], tlx[PPT][POINTS];
  int tmin = 0x7fffffff, tmax = -1;                                // first / last row this thread's samples read
  bool global_only = false;                                        // a sample that cannot be staged at all
  if (tid < kConsumers) {
#pragma unroll
    for (int j = 0; j < PPT; j++) {
      const int i = tid + j * kConsumers;
#pragma unroll
      for (int k = 0; k < POINTS; k++) { ta[j][k] = 0xffffffffu; tly[j][k] = 0.f; tlx[j][k] = 0.f; }
      if (i < band_px) {
        float py[POINTS], px[POINTS];
        fr_points<POINTS>(boxes + ((size_t)n * HW + r0 * W + i) * 5, spatial_scale, py, px);
#pragma unroll
        for (int k = 0; k < POINTS; k++) {
          TapGeo g = fr_tap_geo(py[k], px[k], H, W);
          if (!g.valid) continue;
          // where the reference clamps the lower / right neighbour onto the tap itself (last row / column, fr.py:40-52: fraction
          // exactly 0) the 2 x 2 cell is moved one row / column back with fraction exactly 1: the same two products in the same
          // order, and every read stays inside the map.  (H == 1 cannot: that sample reads global memory.)
          if (g.yh == g.yl) { if (H < 2) { global_only = true; continue; } g.yl -= 1; g.ly = 1.f; }
          if (g.xh == g.xl) { g.xl -= 1; g.lx = 1.f; }             // W % 4 == 0: W >= 4
          ta[j][k] = (uint32_t)(g.yl * W + g.xl);
          tly[j][k] = g.ly; tlx[j][k] = g.lx;
          tmin = min(tmin, g.yl); tmax = max(tmax, g.yl + 1);
        }
      }
    }
    int wmin = tmin, wmax = tmax;
#pragma unroll
    for (int o = 16; o; o >>= 1) { wmin = min(wmin, __shfl_xor_sync(0xffffffffu, wmin, o)); wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o)); }
    if (lane == 0) { atomicMin(&R.span[0], max(wmin, 0)); atomicMax(&R.span[1], wmax); }
  }
  __syncthreads();
  {
    // the staged window: every row the band's samples read when that fits the stage (rows_per_band + 2 * kHalo rows), else the
    // band and kHalo rows either side (the samples beyond read global memory)
    const int need_lo = min(r0, R.span[0]), need_hi = max(r1, R.span[1] + 1);
    if (need_hi - need_lo <= rows_per_band + 2 * kHalo) { R.lo = need_lo; R.hi = need_hi; }
  }
  const int lo = R.lo, hi = R.hi;

  if (tid >= kConsumers) {
    if (lane == 0) R.produce(feat + ((size_t)n * C + c0) * HW, HW, W, nch, stages);
    return;
  }

  const bool straight = !global_only && (tmax < 0 || (tmin >= lo && tmax < hi));   // every sample of this thread reads shared memory
  Tap4 gt[PPT][POINTS];                                            // global taps, only touched when !straight (local memory)
  const uint32_t zb = (uint32_t)R.zoff * 4u, wb = (uint32_t)W * 4u;
#pragma unroll
  for (int j = 0; j < PPT; j++)
#pragma unroll
    for (int k = 0; k < POINTS; k++) ta[j][k] = ta[j][k] == 0xffffffffu ? zb : (ta[j][k] - (uint32_t)(lo * W)) * 4u;
  if (!straight) {
#pragma unroll 1
    for (int j = 0; j < PPT; j++) {
      const int i = tid + j * kConsumers;
      if (i < band_px) fr_decode<POINTS>(boxes + ((size_t)n * HW + r0 * W + i) * 5, spatial_scale, H, W, gt[j]);
    }
  }
  const size_t plane0 = ((size_t)n * C + c0) * HW;
  float* dst = out + plane0 + (size_t)r0 * W + tid;                // pixel j of this thread: dst[j * kConsumers]
  const float* gplane = feat + plane0;
  const uint32_t cb = (uint32_t)((r0 - lo) * W + tid) * 4u;        // centre of pixel 0, byte offset in the stage
  const unsigned char* ring_b = reinterpret_cast<const unsigned char*>(R.ring);
  const uint32_t stage_bytes = (uint32_t)R.stage_elems * 4u;
  const unsigned char* sb = ring_b;

  uint32_t stage = 0, phase = 0;
  for (int c = 0; c < nch; c++) {
    mbar_wait(&R.full[stage], phase);
    auto at = [&](uint32_t off) { return *reinterpret_cast<const float*>(sb + off); };
    if (straight) {
      float v[PPT];
#pragma unroll
      for (int j = 0; j < PPT; j++) {
        const bool exists = tid + j * kConsumers < band_px;
        v[j] = at(exists ? cb + (uint32_t)(j * kConsumers) * 4u : zb);
#pragma unroll
        for (int k = 0; k < POINTS; k++) {
          uint32_t a = ta[j][k];
          float ly = tly[j][k], lx = tlx[j][k];
          // opaque to the optimiser: without this the weights and the four offsets are hoisted out of the channel loop and
          // the 15 registers per pixel are 40 again
          if (POINTS > 1) asm volatile("" : "+r"(a), "+f"(ly), "+f"(lx));   // (points = 1: 8 registers per pixel are affordable, let them be hoisted)
          const float hy = 1.f - ly, hx = 1.f - lx;
          const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
          v[j] += w1 * at(a) + w2 * at(a + 4u) + w3 * at(a + wb) + w4 * at(a + wb + 4u);
        }
      }
#pragma unroll
      for (int j = 0; j < PPT; j++)
        if (tid + j * kConsumers < band_px) st_stream(dst + j * kConsumers, v[j]);
    } else {
#pragma unroll 1
      for (int j = 0; j < PPT; j++) {
        if (tid + j * kConsumers >= band_px) break;
        float v = at(cb + (uint32_t)(j * kConsumers) * 4u);
#pragma unroll
        for (int k = 0; k < POINTS; k++) {
          const Tap4 t = gt[j][k];
          if (t.o00 >= 0)
            v += t.w1 * __ldg(gplane + t.o00) + t.w2 * __ldg(gplane + t.o01) + t.w3 * __ldg(gplane + t.o10) + t.w4 * __ldg(gplane + t.o11);
        }
        st_stream(dst + j * kConsumers, v);
      }
    }
    __syncwarp();                                                  // every lane issued its stores, so its smem reads returned
    if (lane == 0) mbar_arrive(&R.empty[stage]);
    dst += HW;
    gplane += HW;
    sb += stage_bytes;
    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; sb = ring_b; }
  }
}

template <int POINTS>
__global__ void __launch_bounds__(Geo<POINTS>::CW * 32 + 32, Geo<POINTS>::MINB) feature_refine_tma_kernel(const float* __restrict__ feat,
                                                                                       const float* __restrict__ boxes, int C, int H,
                                                                                       int W, float spatial_scale, int rows_per_band,
                                                                                       int ch_per_cta, int stages, float* __restrict__ out) {
  fr_tma_body<POINTS>(feat, boxes, C, H, W, spatial_scale, rows_per_band, ch_per_cta, stages, out, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Every FPN level of a head in ONE launch (FeatureRefineModule applies FR to each level, fr.py:339-346): a 1-D grid of
// (level, image, channel chunk, row band) items, level 0 first.  At cfg4 the five per-level launches took 130 us for 361 MB
// while level 0 alone streams at 4.3 TB/s; the coarse levels (8 x 8 ... 32 x 32 maps) are launch-bound on their own.
constexpr int kMaxLevels = 8;
struct FrLevel { const float* feat; const float* boxes; float* out; int H, W; float scale; int rows, cpc, stages, bands, chunks, item_begin; };
struct FrLevels { FrLevel lv[kMaxLevels]; int n, C; };
template <int POINTS>
__global__ void __launch_bounds__(Geo<POINTS>::CW * 32 + 32, Geo<POINTS>::MINB) feature_refine_tma_multi_kernel(const __grid_constant__ FrLevels L) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < kMaxLevels; i++)
    if (i < L.n && (int)blockIdx.x >= L.lv[i].item_begin) l = i;
  // a static index per case: the level's scalars stay constant-bank operands (a dynamically indexed parameter struct costs
  // 24 more registers and the third resident CTA per SM)
#define JDET_FR_CASE(I)                                                                                                \
  case I: {                                                                                                            \
    int item = blockIdx.x - L.lv[I].item_begin;                                                                        \
    const int band = item % L.lv[I].bands; item /= L.lv[I].bands;                                                      \
    const int chunk = item % L.lv[I].chunks;                                                                           \
    fr_tma_body<POINTS>(L.lv[I].feat, L.lv[I].boxes, L.C, L.lv[I].H, L.lv[I].W, L.lv[I].scale, L.lv[I].rows, L.lv[I].cpc,       \
                        L.lv[I].stages, L.lv[I].out, band, chunk, item / L.lv[I].chunks);                              \
    break;                                                                                                             \
  }
  switch (l) {
    JDET_FR_CASE(0) JDET_FR_CASE(1) JDET_FR_CASE(2) JDET_FR_CASE(3) JDET_FR_CASE(4) JDET_FR_CASE(5) JDET_FR_CASE(6) JDET_FR_CASE(7)
  }
#undef JDET_FR_CASE
}

}  // namespace fr_tma

}  // namespace jdet

JDET_API int jdet_feature_refine(const float* features, const float* best_rbboxes, int N, int C, int H, int W,
                                 int points, float spatial_scale, float* output, void* stream);

namespace jdet {
// band / stage / channel-chunk geometry of the TMA-staged path for one map; false: the map does not qualify.
template <int POINTS>
static bool fr_tma_config(const float* features, int N, int C, int H, int W, fr_tma::FrLevel* out) {
  using namespace fr_tma;
  using G = Geo<POINTS>;
  const int band_pixels = G::CW * 32 * G::PPT;
  if (W % 4 != 0 || W > band_pixels || ((uintptr_t)features & 15) != 0) return false;
  const int rows = max(1, min(H, band_pixels / W));
  const size_t stage_bytes = (size_t)fr_stage_elems<POINTS>(rows, W) * 4;
  // points = 1: ~64 KB of copies in flight per CTA, 3 CTAs per SM; points = 5: one 17-warp CTA per SM, ~100 KB
  int stages = (int)((POINTS == 1 ? JDET_FR1_SMEM_KB * 1024 : JDET_FR5_SMEM_KB * 1024) / stage_bytes);
  stages = stages > kMaxStages ? kMaxStages : stages;
  if (stages < 2) return false;
  const int bands = jdet_ceil_div(H, rows);
  int cpc = C;
  const long long want = 148 * 6;
  while (cpc > 4 * stages && (long long)bands * jdet_ceil_div(C, cpc) * N < want) cpc = (cpc + 1) / 2;
  out->H = H; out->W = W; out->rows = rows; out->stages = stages; out->cpc = cpc; out->bands = bands; out->chunks = jdet_ceil_div(C, cpc);
  return true;
}
template <int POINTS>
static size_t fr_tma_smem(const fr_tma::FrLevel& v) {
  using namespace fr_tma;
  return (size_t)v.stages * fr_stage_elems<POINTS>(v.rows, v.W) * 4 + kRingTail;
}

template <int POINTS>
static int fr_tma_launch(const float* features, const float* best_rbboxes, int N, int C, int H, int W, float spatial_scale,
                         float* output, cudaStream_t st, bool* taken) {
  using namespace fr_tma;
  FrLevel cfg;
  *taken = fr_tma_config<POINTS>(features, N, C, H, W, &cfg);
  if (!*taken) return 0;
  const size_t smem = fr_tma_smem<POINTS>(cfg);
  dim3 g(cfg.bands, cfg.chunks, N);
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_kernel<POINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // without this the driver picks the smallest carve-out that fits ONE block (ncu: occupancy_limit_shared_mem = 1)
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_kernel<POINTS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  feature_refine_tma_kernel<POINTS><<<g, Geo<POINTS>::CW * 32 + 32, smem, st>>>(features, best_rbboxes, C, H, W, spatial_scale, cfg.rows, cfg.cpc,
                                                                                cfg.stages, output);
  return (int)cudaGetLastError();
}

template <int POINTS>
static int fr_multi(const float* const* features, const float* const* best_rbboxes, int nlevels, int N, int C, const int* Hs, const int* Ws,
                    const float* scales, float* const* outputs, void* stream) {
  using namespace fr_tma;
  FrLevels L{};
  L.C = C;
  long long items = 0;
  size_t smem = 0;
  for (int l = 0; l < nlevels; l++) {
    if (Hs[l] < 0 || Ws[l] < 0) return JDET_ERR_BAD_ARG;
    if ((size_t)N * C * Hs[l] * Ws[l] == 0) continue;
    if (!features[l] || !best_rbboxes[l] || !outputs[l]) return JDET_ERR_BAD_ARG;
    FrLevel v;
    if (fr_tma_config<POINTS>(features[l], N, C, Hs[l], Ws[l], &v)) {
      v.feat = features[l]; v.boxes = best_rbboxes[l]; v.out = outputs[l]; v.scale = scales[l]; v.item_begin = (int)items;
      items += (long long)v.bands * v.chunks * N;
      smem = std::max(smem, fr_tma_smem<POINTS>(v));
      L.lv[L.n++] = v;
    } else {
      const int e = jdet_feature_refine(features[l], best_rbboxes[l], N, C, Hs[l], Ws[l], POINTS, scales[l], outputs[l], stream);
      if (e) return e;
    }
  }
  if (L.n == 0) return 0;
  if (items > 0x7fffffffLL) return JDET_ERR_UNSUPPORTED;
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_multi_kernel<POINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_multi_kernel<POINTS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  feature_refine_tma_multi_kernel<POINTS><<<(int)items, Geo<POINTS>::CW * 32 + 32, smem, (cudaStream_t)stream>>>(L);
  return (int)cudaGetLastError();
}
}  // namespace jdet

// jdet.ops.fr.feature_refine(features, best_rbboxes, spatial_scale, points) (ops/fr.py:255-273)
JDET_API int jdet_feature_refine(const float* features, const float* best_rbboxes, int N, int C, int H, int W,
                                 int points, float spatial_scale, float* output, void* stream) {
  using namespace jdet;
  if (N < 0 || C < 0 || H < 0 || W < 0 || (points != 1 && points != 5)) return JDET_ERR_BAD_ARG;
  if ((size_t)N * C * H * W == 0) return 0;
  if (!features || !best_rbboxes || !output) return JDET_ERR_BAD_ARG;
  if (N > 65535) return JDET_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  // TMA-staged path: full-width row bands (contiguous in NCHW), PPT pixels of the band per thread
  // (measured on B200, cfg4: points = 1: 0.130 ms vs 0.203 ms for the 16-B-vector register gather)
  static const bool no_staged_p5 = getenv("JDET_FR_P5_GATHER") != nullptr;      // A/B: points = 5 on the L1 gather kernel
  bool taken = false;
  const int e = points == 1 ? fr_tma_launch<1>(features, best_rbboxes, N, C, H, W, spatial_scale, output, st, &taken)
                : no_staged_p5 ? 0 : fr_tma_launch<5>(features, best_rbboxes, N, C, H, W, spatial_scale, output, st, &taken);
  if (e || taken) return e;
  const int ptiles = jdet_ceil_div(HW, 256);
  // enough CTAs to fill 148 SMs several times over, but keep slabs long enough to amortise the box decode
  int ch_per_cta = C;
  while (ch_per_cta > 16 && (long long)ptiles * jdet_ceil_div(C, ch_per_cta) * N < 148 * 8) ch_per_cta = (ch_per_cta + 1) / 2;
  dim3 grid(ptiles, jdet_ceil_div(C, ch_per_cta), N);
  // (A/B on B200, level 0 of cfg4: this 16-B-vector kernel 121 us; scalar kernel with 8 channels in flight 178 us)
  if (points == 1 && W % 4 == 0 && (((uintptr_t)features | (uintptr_t)output) & 15) == 0) {
    const int vt = jdet_ceil_div(HW / 4, 256);
    int cpc = C;
    while (cpc > 16 && (long long)vt * jdet_ceil_div(C, cpc) * N < 148 * 8) cpc = (cpc + 1) / 2;
    dim3 vgrid(vt, jdet_ceil_div(C, cpc), N);
    feature_refine_p1_vec4_kernel<<<vgrid, 256, 0, st>>>(features, best_rbboxes, C, H, W, spatial_scale, cpc, output);
  } else if (points == 1) feature_refine_kernel<1><<<grid, 256, 0, st>>>(features, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, output);
  else             feature_refine_kernel<5><<<grid, 256, 0, st>>>(features, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, output);
  return (int)cudaGetLastError();
}

// feature_refine on every FPN level of a head in one call (FeatureRefineModule.execute, ops/fr.py:339-346, calls FR per level):
// the levels that qualify for the TMA-staged path (W % 4 == 0, a band fits) share ONE launch; the others take the per-level
// kernels.  features / best_rbboxes / outputs / Hs / Ws / scales: HOST arrays of nlevels (<= 8) entries.
JDET_API int jdet_feature_refine_multi(const float* const* features, const float* const* best_rbboxes, int nlevels, int N, int C,
                                       const int* Hs, const int* Ws, const float* scales, int points, float* const* outputs,
                                       void* stream) {
  using namespace jdet;
  if (nlevels <= 0 || nlevels > fr_tma::kMaxLevels || N < 0 || C < 0 || !features || !best_rbboxes || !Hs || !Ws || !scales || !outputs ||
      (points != 1 && points != 5))
    return JDET_ERR_BAD_ARG;
  if (N > 65535) return JDET_ERR_UNSUPPORTED;
  static const bool no_staged_p5 = getenv("JDET_FR_P5_GATHER") != nullptr;
  if (points == 1) return fr_multi<1>(features, best_rbboxes, nlevels, N, C, Hs, Ws, scales, outputs, stream);
  if (!no_staged_p5) return fr_multi<5>(features, best_rbboxes, nlevels, N, C, Hs, Ws, scales, outputs, stream);
  for (int l = 0; l < nlevels; l++) {
    if (Hs[l] < 0 || Ws[l] < 0) return JDET_ERR_BAD_ARG;
    const int e = jdet_feature_refine(features[l], best_rbboxes[l], N, C, Hs[l], Ws[l], points, scales[l], outputs[l], stream);
    if (e) return e;
  }
  return 0;
}

// backward of jdet_feature_refine w.r.t. features: FeatureRefineFunction.grad (ops/fr.py:266-271, 242-252)
JDET_API int jdet_feature_refine_backward(const float* grad_output, const float* best_rbboxes, int N, int C, int H, int W,
                                          int points, float spatial_scale, float* grad_input, void* stream) {
  using namespace jdet;
  if (N < 0 || C < 0 || H < 0 || W < 0 || (points != 1 && points != 5)) return JDET_ERR_BAD_ARG;
  const size_t total = (size_t)N * C * H * W;
  if (total == 0) return 0;
  if (!grad_output || !best_rbboxes || !grad_input || N > 65535) return JDET_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  JDET_RETURN_IF_CUDA(cudaMemcpyAsync(grad_input, grad_output, total * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const int HW = H * W;
  const int ptiles = jdet_ceil_div(HW, 256);
  int ch_per_cta = C;
  while (ch_per_cta > 16 && (long long)ptiles * jdet_ceil_div(C, ch_per_cta) * N < 148 * 8) ch_per_cta = (ch_per_cta + 1) / 2;
  dim3 grid(ptiles, jdet_ceil_div(C, ch_per_cta), N);
  if (points == 1) feature_refine_bwd_kernel<1><<<grid, 256, 0, st>>>(grad_output, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, grad_input);
  else             feature_refine_bwd_kernel<5><<<grid, 256, 0, st>>>(grad_output, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, grad_input);
  return (int)cudaGetLastError();
}
