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

// ---- staged gather kernel ----------------------------------------------------------------------
// grid = (R, C/SLAB slabs); 256 threads.  Requires C % 64 == 0 and PH*PW*gh*gw <= kMaxSamples
// (sampling_ratio > 0).  Thread task = (bin, channel quad): SLAB/4 lanes span the slab's channels.
template <int VERSION, int SLAB>
__global__ void __launch_bounds__(256) roi_align_nhwc_kernel(const float* __restrict__ feat_nhwc,
                                                              const float* __restrict__ rois, int C, int H, int W,
                                                              int PH, int PW, float spatial_scale, int sample_num,
                                                              float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int nbins = PH * PW;
  const int r = blockIdx.x, c0 = blockIdx.y * SLAB;
  constexpr int QL = SLAB / 4;                      // lanes (channel quads) per bin
  __shared__ RoiGeom g;
  if (threadIdx.x == 0) g = roi_geom<VERSION>(rois + (size_t)r * 6, spatial_scale, sample_num, PH, PW);
  __syncthreads();
  const int spb = g.gh * g.gw;                       // samples per bin
  SampleTap* taps = reinterpret_cast<SampleTap*>(smem);
  float* s_out = reinterpret_cast<float*>(taps + nbins * spb);   // [SLAB][nbins]
  for (int s = threadIdx.x; s < nbins * spb; s += blockDim.x) {
    const int bin = s / spb, k = s - bin * spb;
    taps[s] = make_tap<VERSION>(g, bin / PW, bin % PW, k / g.gw, k % g.gw, H, W);
  }
  __syncthreads();
  const float* base = feat_nhwc + (size_t)g.batch * H * W * C + c0;
  const int q = threadIdx.x % QL;                   // channel quad within the slab
  for (int bin = threadIdx.x / QL; bin < nbins; bin += blockDim.x / QL) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const SampleTap* tp = taps + bin * spb;
    for (int k = 0; k < spb; k++) {
      const SampleTap t = tp[k];
      if (t.o00 < 0) continue;
      const float4 a = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o00 * C) + q);
      const float4 b = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o01 * C) + q);
      const float4 c = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o10 * C) + q);
      const float4 d = __ldg(reinterpret_cast<const float4*>(base + (size_t)t.o11 * C) + q);
      acc.x += t.w1 * a.x + t.w2 * b.x + t.w3 * c.x + t.w4 * d.x;
      acc.y += t.w1 * a.y + t.w2 * b.y + t.w3 * c.y + t.w4 * d.y;
      acc.z += t.w1 * a.z + t.w2 * b.z + t.w3 * c.z + t.w4 * d.z;
      acc.w += t.w1 * a.w + t.w2 * b.w + t.w3 * c.w + t.w4 * d.w;
    }
    const float cnt = g.inv_count;
    s_out[(4 * q + 0) * nbins + bin] = acc.x / cnt;
    s_out[(4 * q + 1) * nbins + bin] = acc.y / cnt;
    s_out[(4 * q + 2) * nbins + bin] = acc.z / cnt;
    s_out[(4 * q + 3) * nbins + bin] = acc.w / cnt;
  }
  __syncthreads();
  // out[r][c0 .. c0+SLAB-1][bins] is one contiguous run of SLAB*nbins floats
  float* dst = out + ((size_t)r * C + c0) * nbins;
  const int total = SLAB * nbins;
  if ((total & 3) == 0 && ((((size_t)r * C + c0) * nbins) & 3) == 0) {
    for (int i = threadIdx.x * 4; i < total; i += blockDim.x * 4)
      st_stream_v4(dst + i, s_out[i], s_out[i + 1], s_out[i + 2], s_out[i + 3]);
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) st_stream(dst + i, s_out[i]);
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
  SampleTap* taps = reinterpret_cast<SampleTap*>(smem);
  float* s_go = reinterpret_cast<float*>(taps + nbins * spb);   // [SLAB][nbins], as laid out in grad_out
  for (int s = threadIdx.x; s < nbins * spb; s += blockDim.x) {
    const int bin = s / spb, k = s - bin * spb;
    taps[s] = make_tap<VERSION>(g, bin / PW, bin % PW, k / g.gw, k % g.gw, H, W);
  }
  const float* src = grad_out + ((size_t)r * C + c0) * nbins;
  for (int i = threadIdx.x; i < SLAB * nbins; i += blockDim.x) s_go[i] = __ldg(src + i);
  __syncthreads();
  float* base = grad_nhwc + (size_t)g.batch * H * W * C + c0;
  const int q = threadIdx.x % QL;
  const float cnt = (float)spb;                      // the backward divides by gh*gw in both versions
  for (int bin = threadIdx.x / QL; bin < nbins; bin += blockDim.x / QL) {
    const float4 go = make_float4(s_go[(4 * q + 0) * nbins + bin], s_go[(4 * q + 1) * nbins + bin],
                                  s_go[(4 * q + 2) * nbins + bin], s_go[(4 * q + 3) * nbins + bin]);
    const SampleTap* tp = taps + bin * spb;
    for (int k = 0; k < spb; k++) {
      const SampleTap t = tp[k];
      if (t.o00 < 0) continue;
      const int o[4] = {t.o00, t.o01, t.o10, t.o11};
      const float w[4] = {t.w1, t.w2, t.w3, t.w4};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float4 v = make_float4(go.x * w[j] / cnt, go.y * w[j] / cnt, go.z * w[j] / cnt, go.w * w[j] / cnt);
        atomicAdd(reinterpret_cast<float4*>(base + (size_t)o[j] * C) + q, v);
      }
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
    const int slab = (C % 128 == 0) ? 128 : 64;      // 128: one bin per warp, 512 contiguous bytes per tap
    const size_t smem = (size_t)nbins * sampling_ratio * sampling_ratio * sizeof(SampleTap) + (size_t)slab * nbins * 4;
    dim3 grid(R, C / slab);
    // (measured alternatives on B200, cfg2: one CTA per RoI over all channels 163 us; per-bin tap merging
    //  185-250 us; 64-channel slabs 146 us, with 128/192/64-thread CTAs 144/150/178 us; 128-channel slabs 136 us)
#define JDET_LAUNCH_ROI(V, S)                                                                                          \
  do {                                                                                                                 \
    if (smem > 48 * 1024)                                                                                              \
      JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(roi_align_nhwc_kernel<V, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    roi_align_nhwc_kernel<V, S><<<grid, 256, smem, st>>>(nhwc, rois, C, H, W, PH, PW, spatial_scale, sampling_ratio, output);          \
  } while (0)
    if (version == 1) { if (slab == 128) JDET_LAUNCH_ROI(1, 128); else JDET_LAUNCH_ROI(1, 64); }
    else              { if (slab == 128) JDET_LAUNCH_ROI(0, 128); else JDET_LAUNCH_ROI(0, 64); }
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
    const size_t smem = (size_t)nbins * sampling_ratio * sampling_ratio * sizeof(SampleTap) + (size_t)slab * nbins * 4;
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
