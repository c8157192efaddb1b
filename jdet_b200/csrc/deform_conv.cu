// deform_conv.cu — DeformConv v1 forward (generic shapes) + AlignConv offset field for sm_100a.
//
// Replaces, for the forward path:
//   deformable_im2col_gpu_kernel + jt.matmul  (/root/reference/python/jdet/ops/dcn_v1.py:25-56,
//       131-184, 309-339, 412-454) — the reference materialises columns (C*kh*kw, B*Ho*Wo) in HBM
//       (1.2 GB at S2ANet level 0) and multiplies with cuBLAS SGEMM;
//   AlignConv.get_offset (/root/reference/python/jdet/models/roi_heads/s2anet_head.py:677-713) —
//       ~25 elementwise Jittor kernels per image per level.
//
// This file is the GENERIC path: any kernel size / stride / padding / dilation / groups /
// deformable_groups, fp32 FMA implicit GEMM with the bilinear sampler as the A-operand producer
// (no columns tensor).  The S2ANet AlignConv shape (3x3, stride 1, C % 64 == 0) has its own
// tensor-core kernel in align_conv_tc.cu.
//
// Layout in HBM: x (B,C,H,W); offset (B, dg*2*kh*kw, Ho, Wo) with channel 2t = dy, 2t+1 = dx;
// weight (Co, C/groups, kh, kw); out (B,Co,Ho,Wo); all fp32 contiguous.
#include "common.cuh"

namespace jdet {

// s2anet_head.py:677-713 for all images at once.  anchors (N,H,W,5) -> offset (N, 2*k*k, H, W)
__global__ void __launch_bounds__(256) align_conv_offset_kernel(const float* __restrict__ anchors, int N, int H, int W,
                                                                 float stride, int k, float* __restrict__ offset) {
  const int HW = H * W;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * HW) return;
  const int n = idx / HW, p = idx - n * HW;
  const float* a = anchors + (size_t)idx * 5;
  const float xc = (float)(p % W), yc = (float)(p / W);
  const float x_ctr = __fdiv_rn(a[0], stride), y_ctr = __fdiv_rn(a[1], stride);
  const float w = __fdiv_rn(a[2], stride), h = __fdiv_rn(a[3], stride);
  const float c = cosf(a[4]), s = sinf(a[4]);
  const float dw = __fdiv_rn(w, (float)k), dh = __fdiv_rn(h, (float)k);
  const int pad = (k - 1) / 2;
  float* o = offset + (size_t)n * 2 * k * k * HW + p;
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++) {
      const float xx = (float)(j - pad), yy = (float)(i - pad);
      const float x = __fmul_rn(dw, xx), y = __fmul_rn(dh, yy);
      const float xr = __fsub_rn(__fmul_rn(c, x), __fmul_rn(s, y));
      const float yr = __fadd_rn(__fmul_rn(s, x), __fmul_rn(c, y));
      const float xa = __fadd_rn(xr, x_ctr), ya = __fadd_rn(yr, y_ctr);
      const int t = i * k + j;
      o[(size_t)(2 * t) * HW] = __fsub_rn(ya, __fadd_rn(yc, yy));
      o[(size_t)(2 * t + 1) * HW] = __fsub_rn(xa, __fadd_rn(xc, xx));
    }
}

// dcn_v1.py:25-56 + the validity test at :170
__device__ __forceinline__ float dcn_sample(const float* __restrict__ plane, int H, int W, float h, float w) {
  if (!(h > -1.f && w > -1.f && h < (float)H && w < (float)W)) return 0.f;
  const int hl = (int)floorf(h), wl = (int)floorf(w);
  const int hh = hl + 1, wh = wl + 1;
  const float lh = h - (float)hl, lw = w - (float)wl;
  const float uh = 1.f - lh, uw = 1.f - lw;
  float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
  if (hl >= 0 && wl >= 0) v1 = __ldg(plane + hl * W + wl);
  if (hl >= 0 && wh <= W - 1) v2 = __ldg(plane + hl * W + wh);
  if (hh <= H - 1 && wl >= 0) v3 = __ldg(plane + hh * W + wl);
  if (hh <= H - 1 && wh <= W - 1) v4 = __ldg(plane + hh * W + wh);
  return uh * uw * v1 + uh * lw * v2 + lh * uw * v3 + lh * lw * v4;
}

struct DcnShape {
  int B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, groups, dg, Ho, Wo;
};

constexpr int BM = 64, BN = 64, BK = 16;

// grid = (ceil(B*Ho*Wo / BM), ceil(Cog / BN), groups); 256 threads, 4x4 outputs each.
__global__ void __launch_bounds__(256) deform_conv_simt_kernel(const float* __restrict__ x, const float* __restrict__ offset,
                                                                const float* __restrict__ weight, DcnShape s, int relu,
                                                                float* __restrict__ out) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int HoWo = s.Ho * s.Wo;
  const long long M = (long long)s.B * HoWo;
  const int Cg = s.C / s.groups, Cog = s.Co / s.groups;
  const int K = Cg * s.kh * s.kw;
  const int g = blockIdx.z;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int cpdg = s.C / s.dg;
  const int khw = s.kh * s.kw;

  // A producer: thread -> (pixel = tid % 64, k-lane = tid / 64 .. +4 .. 12)
  const int am = tid & 63, ak0 = tid >> 6;
  const long long m = m0 + am;
  const bool mvalid = m < M;
  int b = 0, ho = 0, wo = 0;
  if (mvalid) { b = (int)(m / HoWo); const int p = (int)(m - (long long)b * HoWo); ho = p / s.Wo; wo = p - ho * s.Wo; }
  const int h_in = ho * s.sh - s.ph, w_in = wo * s.sw - s.pw;

  const int tx = tid & 15, ty = tid >> 4;       // output micro-tile: rows ty*4.., cols tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int kk = ak0 + 4 * r, k = k0 + kk;
      float v = 0.f;
      if (mvalid && k < K) {
        const int cl = k / khw, t = k - cl * khw;
        const int c = g * Cg + cl;
        const int i = t / s.kw, j = t - i * s.kw;
        const float* off = offset + ((size_t)b * s.dg + c / cpdg) * 2 * khw * HoWo + (size_t)ho * s.Wo + wo;
        const float oh = __ldg(off + (size_t)(2 * t) * HoWo), ow = __ldg(off + (size_t)(2 * t + 1) * HoWo);
        const float h_im = (float)(h_in + i * s.dh) + oh, w_im = (float)(w_in + j * s.dw) + ow;
        v = dcn_sample(x + ((size_t)b * s.C + c) * s.H * s.W, s.H, s.W, h_im, w_im);
      }
      As[kk][am] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int idx = tid + 256 * r;             // 1024 = 64 n x 16 k
      const int nn = idx >> 4, kk = idx & 15;
      const int k = k0 + kk, co = n0 + nn;
      Bs[kk][nn] = (k < K && co < Cog) ? __ldg(weight + ((size_t)(g * Cog + co)) * K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bq = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const long long mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
    const int bb = (int)(mm / HoWo);
    const int p = (int)(mm - (long long)bb * HoWo);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int co = n0 + tx * 4 + j;
      if (co >= Cog) continue;
      float v = acc[i][j];
      if (relu) v = fmaxf(v, 0.f);
      out[((size_t)bb * s.Co + g * Cog + co) * HoWo + p] = v;
    }
  }
}

// ---- backward building blocks (ops/dcn_v1.py:131-306) ---------------------------------------------
// columns layout everywhere: [(c*kh*kw + t)][b][ho][wo]  (the reference's), i.e. (K, B*P).

// sampled columns (the forward's A operand, materialised): needed for grad_weight = grad_out x columns^T
__global__ void __launch_bounds__(256) deform_im2col_kernel(const float* __restrict__ x, const float* __restrict__ offset,
                                                             DcnShape s, float* __restrict__ col) {
  const int P = s.Ho * s.Wo, khw = s.kh * s.kw;
  const long long total = (long long)s.C * s.B * P;
  const int cpdg = s.C / s.dg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int b = (int)((idx / P) % s.B);
    const int c = (int)(idx / ((long long)P * s.B));
    const int ho = p / s.Wo, wo = p - ho * s.Wo;
    const float* plane = x + ((size_t)b * s.C + c) * s.H * s.W;
    const float* off = offset + ((size_t)b * s.dg + c / cpdg) * 2 * khw * P + p;
    for (int t = 0; t < khw; t++) {
      const int i = t / s.kw, j = t - i * s.kw;
      const float h_im = (float)(ho * s.sh - s.ph + i * s.dh) + __ldg(off + (size_t)(2 * t) * P);
      const float w_im = (float)(wo * s.sw - s.pw + j * s.dw) + __ldg(off + (size_t)(2 * t + 1) * P);
      col[(((size_t)c * khw + t) * s.B + b) * P + p] = dcn_sample(plane, s.H, s.W, h_im, w_im);
    }
  }
}

// grad_x: scatter of column gradients with the bilinear weights (deformable_col2im_gpu_kernel :185-241)
__global__ void __launch_bounds__(256) deform_col2im_kernel(const float* __restrict__ colg, const float* __restrict__ offset,
                                                             DcnShape s, float* __restrict__ gx) {
  const int P = s.Ho * s.Wo, khw = s.kh * s.kw;
  const long long total = (long long)s.C * khw * s.B * P;
  const int cpdg = s.C / s.dg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int b = (int)((idx / P) % s.B);
    const int k = (int)(idx / ((long long)P * s.B));
    const int c = k / khw, t = k - c * khw;
    const int i = t / s.kw, j = t - i * s.kw;
    const int ho = p / s.Wo, wo = p - ho * s.Wo;
    const float* off = offset + ((size_t)b * s.dg + c / cpdg) * 2 * khw * P + p;
    const float h_im = (float)(ho * s.sh - s.ph + i * s.dh) + __ldg(off + (size_t)(2 * t) * P);
    const float w_im = (float)(wo * s.sw - s.pw + j * s.dw) + __ldg(off + (size_t)(2 * t + 1) * P);
    if (!(h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W)) continue;
    const float g = __ldg(colg + idx);
    const int hl = (int)floorf(h_im), wl = (int)floorf(w_im), hh = hl + 1, wh = wl + 1;
    const float lh = h_im - (float)hl, lw = w_im - (float)wl, uh = 1.f - lh, uw = 1.f - lw;
    float* plane = gx + ((size_t)b * s.C + c) * s.H * s.W;
    if (hl >= 0 && wl >= 0) atomicAdd(plane + hl * s.W + wl, uh * uw * g);
    if (hl >= 0 && wh <= s.W - 1) atomicAdd(plane + hl * s.W + wh, uh * lw * g);
    if (hh <= s.H - 1 && wl >= 0) atomicAdd(plane + hh * s.W + wl, lh * uw * g);
    if (hh <= s.H - 1 && wh <= s.W - 1) atomicAdd(plane + hh * s.W + wh, lh * lw * g);
  }
}

// grad_offset: one thread per (b, deformable group, tap, direction pair, position); sums over the
// group's channels (deformable_col2im_coord_gpu_kernel :243-306, get_coordinate_weight :84-129)
__global__ void __launch_bounds__(256) deform_col2im_coord_kernel(const float* __restrict__ colg, const float* __restrict__ x,
                                                                   const float* __restrict__ offset, DcnShape s,
                                                                   float* __restrict__ goff) {
  const int P = s.Ho * s.Wo, khw = s.kh * s.kw;
  const long long total = (long long)s.B * s.dg * khw * P;
  const int cpdg = s.C / s.dg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int t = (int)((idx / P) % khw);
    const int g = (int)((idx / ((long long)P * khw)) % s.dg);
    const int b = (int)(idx / ((long long)P * khw * s.dg));
    const int i = t / s.kw, j = t - i * s.kw;
    const int ho = p / s.Wo, wo = p - ho * s.Wo;
    const size_t obase = ((size_t)b * s.dg + g) * 2 * khw * P;
    const float h_im = (float)(ho * s.sh - s.ph + i * s.dh) + __ldg(offset + obase + (size_t)(2 * t) * P + p);
    const float w_im = (float)(wo * s.sw - s.pw + j * s.dw) + __ldg(offset + obase + (size_t)(2 * t + 1) * P + p);
    float gh = 0.f, gw = 0.f;
    if (h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W) {
      const int hl = (int)floorf(h_im), wl = (int)floorf(w_im), hh = hl + 1, wh = wl + 1;
      const float lh = h_im - (float)hl, lw = w_im - (float)wl, uh = 1.f - lh, uw = 1.f - lw;
      const bool v1 = hl >= 0 && wl >= 0, v2 = hl >= 0 && wh <= s.W - 1, v3 = hh <= s.H - 1 && wl >= 0, v4 = hh <= s.H - 1 && wh <= s.W - 1;
      for (int cc = 0; cc < cpdg; cc++) {
        const int c = g * cpdg + cc;
        const float* plane = x + ((size_t)b * s.C + c) * s.H * s.W;
        const float p1 = v1 ? __ldg(plane + hl * s.W + wl) : 0.f, p2 = v2 ? __ldg(plane + hl * s.W + wh) : 0.f;
        const float p3 = v3 ? __ldg(plane + hh * s.W + wl) : 0.f, p4 = v4 ? __ldg(plane + hh * s.W + wh) : 0.f;
        const float cg = __ldg(colg + (((size_t)c * khw + t) * s.B + b) * P + p);
        gh += (-uw * p1 - lw * p2 + uw * p3 + lw * p4) * cg;
        gw += (-uh * p1 + uh * p2 - lh * p3 + lh * p4) * cg;
      }
    }
    goff[obase + (size_t)(2 * t) * P + p] = gh;
    goff[obase + (size_t)(2 * t + 1) * P + p] = gw;
  }
}

}  // namespace jdet

// AlignConv.get_offset for a batch (s2anet_head.py:677-721): anchors (N,H,W,5) image space ->
// offset (N, 2*k*k, H, W)
JDET_API int jdet_align_conv_offset(const float* anchors, int N, int H, int W, float stride, int kernel_size,
                                    float* offset, void* stream) {
  if (N < 0 || H < 0 || W < 0 || kernel_size <= 0 || !(kernel_size & 1)) return JDET_ERR_BAD_ARG;
  const long long total = (long long)N * H * W;
  if (total == 0) return 0;
  if (!anchors || !offset || total > 0x7fffffffLL) return JDET_ERR_BAD_ARG;
  jdet::align_conv_offset_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      anchors, N, H, W, stride, kernel_size, offset);
  return (int)cudaGetLastError();
}

// deform_conv(x, offset, weight, stride, padding, dilation, groups, deformable_groups) forward
// (ops/dcn_v1.py:561-600); relu != 0 fuses AlignConv's ReLU (s2anet_head.py:722).
JDET_API int jdet_deform_conv_forward(const float* x, const float* offset, const float* weight, int B, int C, int H,
                                      int W, int Co, int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w,
                                      int dil_h, int dil_w, int groups, int deformable_groups, int relu, float* out,
                                      void* stream) {
  using namespace jdet;
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || Co <= 0 || kh <= 0 || kw <= 0 || stride_h <= 0 || stride_w <= 0 ||
      dil_h <= 0 || dil_w <= 0 || groups <= 0 || deformable_groups <= 0 || C % groups || Co % groups ||
      C % deformable_groups)
    return JDET_ERR_BAD_ARG;
  DcnShape s{B, C, H, W, Co, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, groups, deformable_groups, 0, 0};
  s.Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  s.Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  if (s.Ho <= 0 || s.Wo <= 0) return JDET_ERR_BAD_ARG;
  if (B == 0) return 0;
  if (!x || !offset || !weight || !out) return JDET_ERR_BAD_ARG;
  const long long M = (long long)B * s.Ho * s.Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)jdet_ceil_div(Co / groups, BN), (unsigned)groups);
  deform_conv_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, offset, weight, s, relu, out);
  return (int)cudaGetLastError();
}

// ---- backward building blocks, C ABI (the two GEMMs between them are plain library GEMMs on the host side)
static int jdet_dcn_shape(jdet::DcnShape* s, int B, int C, int H, int W, int kh, int kw, int stride_h, int stride_w, int pad_h,
                          int pad_w, int dil_h, int dil_w, int dg) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0 || stride_h <= 0 || stride_w <= 0 || dil_h <= 0 ||
      dil_w <= 0 || dg <= 0 || C % dg)
    return JDET_ERR_BAD_ARG;
  *s = jdet::DcnShape{B, C, H, W, 0, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, 1, dg, 0, 0};
  s->Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  s->Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  return (s->Ho <= 0 || s->Wo <= 0) ? JDET_ERR_BAD_ARG : 0;
}

// replaces deformable_im2col (ops/dcn_v1.py:309-339): columns (C*kh*kw, B, Ho, Wo)
JDET_API int jdet_deform_im2col(const float* x, const float* offset, int B, int C, int H, int W, int kh, int kw, int stride_h,
                                int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int deformable_groups,
                                float* columns, void* stream) {
  jdet::DcnShape s;
  const int e = jdet_dcn_shape(&s, B, C, H, W, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, deformable_groups);
  if (e) return e;
  if (B == 0) return 0;
  if (!x || !offset || !columns) return JDET_ERR_BAD_ARG;
  jdet::deform_im2col_kernel<<<jdet::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(x, offset, s, columns);
  return (int)cudaGetLastError();
}

// replaces deformable_col2im (ops/dcn_v1.py:376-410): columns gradient (C*kh*kw, B, Ho, Wo) -> grad_x (B,C,H,W), fully written
JDET_API int jdet_deform_col2im(const float* col_grad, const float* offset, int B, int C, int H, int W, int kh, int kw,
                                int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                int deformable_groups, float* grad_x, void* stream) {
  jdet::DcnShape s;
  const int e = jdet_dcn_shape(&s, B, C, H, W, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, deformable_groups);
  if (e) return e;
  if (B == 0) return 0;
  if (!col_grad || !offset || !grad_x) return JDET_ERR_BAD_ARG;
  JDET_RETURN_IF_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)B * C * H * W * 4, (cudaStream_t)stream));
  jdet::deform_col2im_kernel<<<jdet::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(col_grad, offset, s, grad_x);
  return (int)cudaGetLastError();
}

// replaces deformable_col2im_coord (ops/dcn_v1.py:341-374): -> grad_offset (B, dg*2*kh*kw, Ho, Wo), fully written
JDET_API int jdet_deform_col2im_coord(const float* col_grad, const float* x, const float* offset, int B, int C, int H, int W,
                                      int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                      int deformable_groups, float* grad_offset, void* stream) {
  jdet::DcnShape s;
  const int e = jdet_dcn_shape(&s, B, C, H, W, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, deformable_groups);
  if (e) return e;
  if (B == 0) return 0;
  if (!col_grad || !x || !offset || !grad_offset) return JDET_ERR_BAD_ARG;
  jdet::deform_col2im_coord_kernel<<<jdet::num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(col_grad, x, offset, s, grad_offset);
  return (int)cudaGetLastError();
}
