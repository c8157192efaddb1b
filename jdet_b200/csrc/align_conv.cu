// align_conv.cu — AlignConv.execute as one C-ABI call (s2anet_head.py:715-723).
//
// Shapes that qualify (C % 32 == 0, Co % 32 == 0, Co <= 256 — every S2ANet level) run the fused
// tcgen05 kernel in align_conv_tc.cu.  Anything else composes the AlignConv offset field with the
// generic deformable implicit GEMM of deform_conv.cu (fused ReLU).
#include "common.cuh"

extern "C" int jdet_align_conv_offset(const float*, int, int, int, float, int, float*, void*);
extern "C" int jdet_deform_conv_forward(const float*, const float*, const float*, int, int, int, int, int, int, int,
                                        int, int, int, int, int, int, int, int, int, float*, void*);

namespace jdet {
bool align_conv_tc_supported(int C, int Co);
size_t align_conv_tc_workspace_bytes(int N, int C, int H, int W, int Co);
int align_conv_tc_launch(const float* x, const float* anchors, const float* weight, int N, int C, int H, int W, int Co,
                         float stride, float* out, void* workspace, cudaStream_t st);
size_t align_conv_tc_workspace_bytes_multi(int nlevels, int N, int C, const int* Hs, const int* Ws, int Co);
int align_conv_tc_launch_multi(const float* const* xs, const float* const* anchors, const float* weight, int nlevels, int N, int C,
                               const int* Hs, const int* Ws, int Co, const float* strides, float* const* outs, void* workspace,
                               cudaStream_t st, bool x_channels_last);
}  // namespace jdet

JDET_API const char* jdet_version(void) { return "jdet_b200 0.1.0 sm_100a"; }

JDET_API size_t jdet_align_conv_forward_workspace_bytes(int N, int C, int H, int W, int Co) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || Co <= 0) return 256;
  if (jdet::align_conv_tc_supported(C, Co)) return jdet::align_conv_tc_workspace_bytes(N, C, H, W, Co);
  return jdet_align_up((size_t)N * 18 * H * W * sizeof(float), 256);
}

JDET_API int jdet_align_conv_forward(const float* x, const float* anchors, const float* weight, int N, int C, int H,
                                     int W, int Co, float stride, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (N < 0 || C <= 0 || H <= 0 || W <= 0 || Co <= 0) return JDET_ERR_BAD_ARG;
  if (N == 0) return 0;
  if (!x || !anchors || !weight || !out) return JDET_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < jdet_align_conv_forward_workspace_bytes(N, C, H, W, Co)) return JDET_ERR_WORKSPACE;
  if (jdet::align_conv_tc_supported(C, Co))
    return jdet::align_conv_tc_launch(x, anchors, weight, N, C, H, W, Co, stride, out, workspace, (cudaStream_t)stream);
  float* offset = (float*)workspace;
  int e = jdet_align_conv_offset(anchors, N, H, W, stride, 3, offset, stream);
  if (e) return e;
  return jdet_deform_conv_forward(x, offset, weight, N, C, H, W, Co, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, out, stream);
}

// Every FPN level of one head in ONE call (S2ANetHead runs the same AlignConv on its 5 levels, s2anet_head.py:230-237 inside
// forward_single): one persistent tcgen05 launch over all levels' tiles, one weight split.  xs / anchors / outs / Hs / Ws /
// strides are HOST arrays of nlevels entries (device pointers inside).  x_channels_last != 0: every xs[l] is already (N,H,W,C) in
// memory (a torch.channels_last tensor out of the FPN convolutions) and is sampled in place, no re-layout pass.  Needs the tcgen05 shape class (C % 16 == 0,
// Co % 32 == 0, Co <= 256) and nlevels <= 8, else JDET_ERR_UNSUPPORTED (callers loop over jdet_align_conv_forward).
JDET_API size_t jdet_align_conv_forward_multi_workspace_bytes(int nlevels, int N, int C, const int* Hs, const int* Ws, int Co) {
  if (nlevels <= 0 || !Hs || !Ws || N <= 0 || C <= 0 || Co <= 0) return 256;
  return jdet::align_conv_tc_workspace_bytes_multi(nlevels, N, C, Hs, Ws, Co);
}

JDET_API int jdet_align_conv_forward_multi(const float* const* xs, const float* const* anchors, const float* weight, int nlevels,
                                           int N, int C, const int* Hs, const int* Ws, int Co, const float* strides,
                                           float* const* outs, int x_channels_last, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  if (nlevels <= 0 || N < 0 || C <= 0 || Co <= 0 || !xs || !anchors || !weight || !Hs || !Ws || !strides || !outs) return JDET_ERR_BAD_ARG;
  if (N == 0) return 0;
  if (nlevels > 8 || !jdet::align_conv_tc_supported(C, Co)) return JDET_ERR_UNSUPPORTED;
  for (int l = 0; l < nlevels; l++)
    if (!xs[l] || !anchors[l] || !outs[l] || Hs[l] <= 0 || Ws[l] <= 0) return JDET_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < jdet_align_conv_forward_multi_workspace_bytes(nlevels, N, C, Hs, Ws, Co)) return JDET_ERR_WORKSPACE;
  return jdet::align_conv_tc_launch_multi(xs, anchors, weight, nlevels, N, C, Hs, Ws, Co, strides, outs, workspace, (cudaStream_t)stream,
                                          x_channels_last != 0);
}
