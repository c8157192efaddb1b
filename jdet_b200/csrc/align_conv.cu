// align_conv.cu — AlignConv.execute as one C-ABI call (s2anet_head.py:715-723).
//
// v1 composition: AlignConv offset field (deform_conv.cu) -> generic deformable implicit GEMM
// (deform_conv.cu) with fused ReLU.  The tcgen05 fused kernel replaces this body when the shape
// qualifies (align_conv_tc.cu).
#include "common.cuh"

extern "C" int jdet_align_conv_offset(const float*, int, int, int, float, int, float*, void*);
extern "C" int jdet_deform_conv_forward(const float*, const float*, const float*, int, int, int, int, int, int, int,
                                        int, int, int, int, int, int, int, int, int, float*, void*);

JDET_API const char* jdet_version(void) { return "jdet_b200 0.1.0 sm_100a"; }

JDET_API size_t jdet_align_conv_forward_workspace_bytes(int N, int C, int H, int W, int Co) {
  (void)C; (void)Co;
  if (N <= 0 || H <= 0 || W <= 0) return 256;
  return jdet_align_up((size_t)N * 18 * H * W * sizeof(float), 256);
}

JDET_API int jdet_align_conv_forward(const float* x, const float* anchors, const float* weight, int N, int C, int H,
                                     int W, int Co, float stride, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (N < 0 || C <= 0 || H <= 0 || W <= 0 || Co <= 0) return JDET_ERR_BAD_ARG;
  if (N == 0) return 0;
  if (!x || !anchors || !weight || !out) return JDET_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < jdet_align_conv_forward_workspace_bytes(N, C, H, W, Co)) return JDET_ERR_WORKSPACE;
  float* offset = (float*)workspace;
  int e = jdet_align_conv_offset(anchors, N, H, W, stride, 3, offset, stream);
  if (e) return e;
  return jdet_deform_conv_forward(x, offset, weight, N, C, H, W, Co, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, out, stream);
}
