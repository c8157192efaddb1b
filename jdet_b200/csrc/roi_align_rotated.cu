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
template <int VERSION>
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
  t.o00 = yl * W + xl; t.o01 = yl * W + xh; t.o10 = yh * W + xl; t.o11 = yh * W + xh;
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
template <int VERSION>
__device__ __forceinline__ void build_tap_table(const RoiGeom& g, int nbins, int PW, int H, int W, int2* raw, int2* fin,
                                                int* cnt) {
  const int spb = g.gh * g.gw, tpb = 4 * spb;
  for (int s = threadIdx.x; s < nbins * spb; s += blockDim.x) {
    const int bin = s / spb, k = s - bin * spb;
    const SampleTap t = make_tap<VERSION>(g, bin / PW, bin % PW, k / g.gw, k % g.gw, H, W);
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
      if (in) {
        fin[bin * tpb + pos] = leader ? make_int2(me.x, __float_as_int(wsum)) : make_int2(-1, 0);
        if (sl == 0) cnt[bin] = nlead;
      }
    }
  } else {
    for (int bin = threadIdx.x; bin < nbins; bin += blockDim.x) {
      int n = 0;
      for (int j = 0; j < tpb; j++) {
        const int2 e = raw[bin * tpb + j];
        if (e.x >= 0) fin[bin * tpb + n++] = e;
      }
      cnt[bin] = n;
      for (; n < tpb; n++) fin[bin * tpb + n] = make_int2(-1, 0);
    }
  }
  __syncthreads();
}

// ---- staged gather kernel ----------------------------------------------------------------------
// grid = (R, C / SLAB), SLAB = 4*QL*NQ channels; 256 threads.  Requires sampling_ratio > 0 and
// PH*PW*gh*gw <= kMaxSamples.  QL lanes span a bin's channels (QL = 32: one bin per warp, every tap is
// 512 contiguous bytes of one channel-last pixel); each lane owns NQ channel quads 128 channels apart,
// so one table lookup + one address feed NQ 16-B loads.  With C = 256 a CTA is a whole RoI: the table is
// built once per RoI.  The output slab (SLAB x PH*PW, contiguous in (R,C,PH,PW)) is assembled in smem
// and leaves as 16-B streaming stores.
template <int VERSION, int QL, int NQ>
__global__ void __launch_bounds__(256) roi_align_nhwc_kernel(const float* __restrict__ feat_nhwc,
                                                              const float* __restrict__ rois, int C, int H, int W,
                                                              int PH, int PW, float spatial_scale, int sample_num,
                                                              float* __restrict__ out) {
  constexpr int SLAB = 4 * QL * NQ;
  extern __shared__ __align__(16) unsigned char smem[];
  const int nbins = PH * PW;
  const int r = blockIdx.x, c0 = blockIdx.y * SLAB;
  __shared__ RoiGeom g;
  if (threadIdx.x == 0) g = roi_geom<VERSION>(rois + (size_t)r * 6, spatial_scale, sample_num, PH, PW);
  __syncthreads();
  const int spb = g.gh * g.gw;                       // samples per bin
  const int tpb = 4 * spb;                           // tap slots per bin
  const int nslots = nbins * tpb;
  const int S = nbins | 1;                           // odd row stride of s_out: see the store below
  int2* fin = reinterpret_cast<int2*>(smem);         // [nslots] compacted (pixel, weight) lists
  int* cnt = reinterpret_cast<int*>(fin + nslots);   // [nbins rounded up to 4]
  int2* raw = reinterpret_cast<int2*>(cnt + ((nbins + 3) & ~3));   // [nslots], dead after the build: aliased by s_out
  float* s_out = reinterpret_cast<float*>(raw);      // [SLAB][S]
  build_tap_table<VERSION>(g, nbins, PW, H, W, raw, fin, cnt);

  const int q = threadIdx.x % QL;                    // channel quad within a 128-channel group
  const float* base = feat_nhwc + (size_t)g.batch * H * W * C + c0 + 4 * q;
  const int icnt = (int)g.inv_count;
  const bool pow2 = (icnt & (icnt - 1)) == 0;        // x / 2^k == x * 2^-k exactly
  const float rcnt = 1.f / g.inv_count;
  const int rot = (q >> 3) & 3;
  for (int bin = threadIdx.x / QL; bin < nbins; bin += blockDim.x / QL) {
    float4 acc[NQ];
#pragma unroll
    for (int u = 0; u < NQ; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int2* e = fin + bin * tpb;
    const int n = cnt[bin];
    int k = 0;
    for (; k + 4 <= n; k += 4) {                     // 4*NQ independent 16-B loads in flight
      const int4 e01 = *reinterpret_cast<const int4*>(e + k), e23 = *reinterpret_cast<const int4*>(e + k + 2);
      const float* p0 = base + (size_t)e01.x * C;
      const float* p1 = base + (size_t)e01.z * C;
      const float* p2 = base + (size_t)e23.x * C;
      const float* p3 = base + (size_t)e23.z * C;
      float4 v0[NQ], v1[NQ], v2[NQ], v3[NQ];
#pragma unroll
      for (int u = 0; u < NQ; u++) {
        v0[u] = __ldg(reinterpret_cast<const float4*>(p0 + 4 * QL * u));
        v1[u] = __ldg(reinterpret_cast<const float4*>(p1 + 4 * QL * u));
        v2[u] = __ldg(reinterpret_cast<const float4*>(p2 + 4 * QL * u));
        v3[u] = __ldg(reinterpret_cast<const float4*>(p3 + 4 * QL * u));
      }
      const float w0 = __int_as_float(e01.y), w1 = __int_as_float(e01.w);
      const float w2 = __int_as_float(e23.y), w3 = __int_as_float(e23.w);
#pragma unroll
      for (int u = 0; u < NQ; u++) {
        acc[u].x += w0 * v0[u].x + w1 * v1[u].x + w2 * v2[u].x + w3 * v3[u].x;
        acc[u].y += w0 * v0[u].y + w1 * v1[u].y + w2 * v2[u].y + w3 * v3[u].y;
        acc[u].z += w0 * v0[u].z + w1 * v1[u].z + w2 * v2[u].z + w3 * v3[u].z;
        acc[u].w += w0 * v0[u].w + w1 * v1[u].w + w2 * v2[u].w + w3 * v3[u].w;
      }
    }
    for (; k < n; k++) {
      const int2 t = e[k];
      const float* p = base + (size_t)t.x * C;
      const float w = __int_as_float(t.y);
#pragma unroll
      for (int u = 0; u < NQ; u++) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + 4 * QL * u));
        acc[u].x += w * v.x; acc[u].y += w * v.y; acc[u].z += w * v.z; acc[u].w += w * v.w;
      }
    }
    // s_out[channel][bin], row stride S odd.  A lane's 4 channels are 4 rows apart from its neighbour's, so
    // a plain store would hit 8 banks 4 lanes deep; lanes 8 apart instead store their components in a
    // rotated order (rot = lane/8), which spreads every store instruction over all 32 banks.
#pragma unroll
    for (int u = 0; u < NQ; u++) {
      float4 a = acc[u];
      if (pow2) { a.x *= rcnt; a.y *= rcnt; a.z *= rcnt; a.w *= rcnt; }
      else { a.x /= g.inv_count; a.y /= g.inv_count; a.z /= g.inv_count; a.w /= g.inv_count; }
      const float b0 = rot & 1 ? a.y : a.x, b1 = rot & 1 ? a.z : a.y, b2 = rot & 1 ? a.w : a.z, b3 = rot & 1 ? a.x : a.w;
      const float d0 = rot & 2 ? b2 : b0, d1 = rot & 2 ? b3 : b1, d2 = rot & 2 ? b0 : b2, d3 = rot & 2 ? b1 : b3;
      float* row = s_out + (size_t)(4 * QL * u + 4 * q) * S + bin;     // d_j is component (j + rot) & 3
      row[((0 + rot) & 3) * S] = d0;
      row[((1 + rot) & 3) * S] = d1;
      row[((2 + rot) & 3) * S] = d2;
      row[((3 + rot) & 3) * S] = d3;
    }
  }
  __syncthreads();
  // out[r][c0 .. c0+SLAB-1][bins] is one contiguous run of SLAB*nbins floats
  float* dst = out + ((size_t)r * C + c0) * nbins;
  const int total = SLAB * nbins;
  if (S == nbins && (total & 3) == 0 && ((((size_t)r * C + c0) * nbins) & 3) == 0) {
    for (int i = threadIdx.x * 4; i < total; i += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(s_out + i);
      st_stream_v4(dst + i, v.x, v.y, v.z, v.w);
    }
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) st_stream(dst + i, s_out[(i / nbins) * S + i % nbins]);
  }
}

// ---- direct NCHW kernel --------------------------------------------------------------------------
// grid = (R, channel chunks); thread task = (channel, bin); taps from the per-RoI table when it fits,
// otherwise recomputed per element (adaptive sampling grids of huge RoIs).
template <int VERSION>
__global__ void __launch_bounds__(256) roi_align_nchw_kernel(const float* __restrict__ feat, const float* __restrict__ rois,
                                                              int C, int H, int W, int PH, int PW, float spatial_scale,
                                                              int sample_num, int ch_per_cta, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
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
  extern __shared__ __align__(16) unsigned char smem[];
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
  build_tap_table<VERSION>(g, nbins, PW, H, W, raw, fin, cnt);  // one atomic per distinct pixel of a bin
  float* base = grad_nhwc + (size_t)g.batch * H * W * C + c0;
  const int q = threadIdx.x % QL;
  const float cntf = (float)spb;                     // the backward divides by gh*gw in both versions
  for (int bin = threadIdx.x / QL; bin < nbins; bin += blockDim.x / QL) {
    const float4 go = make_float4(s_go[(4 * q + 0) * nbins + bin], s_go[(4 * q + 1) * nbins + bin],
                                  s_go[(4 * q + 2) * nbins + bin], s_go[(4 * q + 3) * nbins + bin]);
    const int2* e = fin + bin * tpb;
    const int n = cnt[bin];
    for (int k = 0; k < n; k++) {
      const int2 t = e[k];
      const float w = __int_as_float(t.y);
      const float4 v = make_float4(go.x * w / cntf, go.y * w / cntf, go.z * w / cntf, go.w * w / cntf);
      atomicAdd(reinterpret_cast<float4*>(base + (size_t)t.x * C) + q, v);
    }
  }
}

template <int VERSION>
__global__ void __launch_bounds__(256) roi_align_bwd_nchw_kernel(const float* __restrict__ grad_out,
                                                                  const float* __restrict__ rois, int C, int H, int W,
                                                                  int PH, int PW, float spatial_scale, int sample_num,
                                                                  int ch_per_cta, float* __restrict__ grad_in) {
  extern __shared__ __align__(16) unsigned char smem[];
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

static bool use_staged(int B, int C, int H, int W, int R, int PH, int PW, int sample_num) {
  if (sample_num <= 0 || C % 64 != 0) return false;
  if ((long long)PH * PW * sample_num * sample_num > kMaxSamples) return false;
  // re-laying the map moves 8*B*C*H*W bytes; it pays once the RoI set samples the map densely
  return (long long)R * PH * PW * sample_num * sample_num * 8 >= (long long)B * H * W;
}

}  // namespace jdet

JDET_API size_t jdet_roi_align_rotated_workspace_bytes(int B, int C, int H, int W, int R, int PH, int PW,
                                                       int sampling_ratio) {
  if (!jdet::use_staged(B, C, H, W, R, PH, PW, sampling_ratio)) return 256;
  return jdet_align_up((size_t)B * C * H * W * sizeof(float), 256);
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
  const int nbins = PH * PW;
  if (use_staged(B, C, H, W, R, PH, PW, sampling_ratio)) {
    const size_t need = jdet_align_up((size_t)B * C * H * W * sizeof(float), 256);
    if (!workspace || workspace_bytes < need) return JDET_ERR_WORKSPACE;
    float* nhwc = (float*)workspace;
    launch_nchw_to_nhwc(input, nhwc, B, C, H * W, st);
    const int slab = (C % 256 == 0) ? 256 : (C % 128 == 0) ? 128 : 64;
    const size_t slot_bytes = (size_t)nbins * sampling_ratio * sampling_ratio * 4 * sizeof(int2);
    const size_t smem = slot_bytes + (size_t)((nbins + 3) & ~3) * 4 + std::max(slot_bytes, (size_t)slab * (nbins | 1) * 4);
    dim3 grid(R, C / slab);
    // (measured on B200, cfg2, same relayout: (RoI, 64-ch slab) CTAs with 16 undeduplicated taps 146 us; 128-ch
    //  slabs 136 us; tap merging by a serial per-bin pass 185-250 us, by a parallel pass without compaction 147 us)
#define JDET_LAUNCH_ROI(V, QL_, NQ_)                                                                                   \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(roi_align_nhwc_kernel<V, QL_, NQ_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    roi_align_nhwc_kernel<V, QL_, NQ_><<<grid, 256, smem, st>>>(nhwc, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, output);        \
  } while (0)
    if (version == 1) { if (slab == 256) JDET_LAUNCH_ROI(1, 32, 2); else if (slab == 128) JDET_LAUNCH_ROI(1, 32, 1); else JDET_LAUNCH_ROI(1, 16, 1); }
    else              { if (slab == 256) JDET_LAUNCH_ROI(0, 32, 2); else if (slab == 128) JDET_LAUNCH_ROI(0, 32, 1); else JDET_LAUNCH_ROI(0, 16, 1); }
#undef JDET_LAUNCH_ROI
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

// backward of jdet_roi_align_rotated w.r.t. input: _RotatedROIAlign[_v1].grad (roi_align_rotated_v1.py:328-351,
// roi_align_rotated.py:285-308).  grad_output (R,C,PH,PW) -> grad_input (B,C,H,W), written in full.
JDET_API size_t jdet_roi_align_rotated_backward_workspace_bytes(int B, int C, int H, int W, int R, int PH, int PW,
                                                                int sampling_ratio) {
  return jdet_roi_align_rotated_workspace_bytes(B, C, H, W, R, PH, PW, sampling_ratio);
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
