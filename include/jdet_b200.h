/* jdet_b200.h — C ABI of libjdet_b200.so: JDet's oriented-box geometry hot path on B200 (sm_100a).
 *
 * This is the drop-in boundary.  Each entry point replaces one `jt.code(...)` inline-op call site
 * of the reference (Jittor/JDet @ 01974379, paths relative to python/jdet/); the binding a
 * maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - every pointer is DEVICE memory owned by the caller, fp32/int32 contiguous, unless it says "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing allocates, nothing synchronises, nothing touches the host: scratch comes from the
 *     caller via (`workspace`, `workspace_bytes`), sized by the matching *_workspace_bytes();
 *     all calls are CUDA-graph capturable;
 *   - return value: 0 = ok; > 0 = a cudaError_t; JDET_ERR_* (< 0) = rejected arguments;
 *   - angles are radians (ops/box_iou_rotated.py:56-59), boxes are [x_ctr, y_ctr, w, h, theta].
 *   - there is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef JDET_B200_H_
#define JDET_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JDET_ERR_BAD_ARG (-1)
#define JDET_ERR_WORKSPACE (-2)
#define JDET_ERR_UNSUPPORTED (-3)

/* library / build identification: "jdet_b200 <version> sm_100a" */
const char* jdet_version(void);

/* ---- box_iou_rotated -------------------------------------------------------------------------
 * replaces: box_iou_rotated()    ops/box_iou_rotated.py:502-509 (kernel :412-461, launch :464-485)
 *           box_iou_rotated_v1() ops/box_iou_rotated_v1.py:507-525 (incl. the small-box zeroing :516-523)
 * boxes1 (n1,5), boxes2 (n2,5) -> ious (n1,n2) row-major.  version: 0 | 1.                        */
size_t jdet_box_iou_rotated_workspace_bytes(int n1, int n2);
int jdet_box_iou_rotated(const float* boxes1, int n1, const float* boxes2, int n2, float* ious, int version,
                         void* workspace, size_t workspace_bytes, void* stream);

/* same op with the arithmetic of the reference's CPU build selectable: arithmetic 1 = CUDA build (== the call above),
 * 0 = CPU build, i.e. the cpu_src path of ops/box_iou_rotated.py:487-500 (std::sort hull :316-325), bit for bit.   */
int jdet_box_iou_rotated_ex(const float* boxes1, int n1, const float* boxes2, int n2, float* ious, int version,
                            int arithmetic, void* workspace, size_t workspace_bytes, void* stream);

/* ---- nms_rotated -----------------------------------------------------------------------------
 * replaces: nms_rotated_cuda(dets, order_t, iou_threshold, box_length) ops/nms_rotated.py:506-513
 *           (kernel :352-411, launch + host reduce :450-493).
 * dets (n, box_length), box_length 5 = [x,y,w,h,theta] | 6 = [..., label]; order (n,) int32 indices by
 * descending score; keep (n,) bytes, fully written, 1 = kept, indexed like dets.
 * Strict `IoU > iou_threshold` (the reference CUDA path), IoU arguments (higher, lower).            */
size_t jdet_nms_rotated_workspace_bytes(int n, int box_length);
int jdet_nms_rotated(const float* dets, int n, int box_length, const int* order, float iou_threshold,
                     unsigned char* keep, void* workspace, size_t workspace_bytes, void* stream);

/* same op with the reference's CPU-path convention selectable: convention 1 = nms_rotated_cuda (== the call above);
 * 0 = nms_rotated_cpu ops/nms_rotated.py:495-504 (loop :414-449): suppress on IoU >= thr, CPU-build IoU arithmetic.
 * This is what the tile -> image merge (data/devkits/result_merge.py:132-145 py_cpu_nms_obb) runs.                 */
int jdet_nms_rotated_ex(const float* dets, int n, int box_length, const int* order, float iou_threshold, int convention,
                        unsigned char* keep, void* workspace, size_t workspace_bytes, void* stream);

/* stable descending argsort (ties: lower index first)
 * replaces: scores.argsort(0, descending=True) ops/nms_rotated.py:519,532                          */
size_t jdet_argsort_desc_workspace_bytes(int n);
int jdet_argsort_desc(const float* scores, int n, int* order, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---- tile -> image merge NMS over quadrilaterals -------------------------------------------------
 * replaces: py_cpu_nms_poly_fast / py_cpu_nms_poly data/devkits/result_merge.py:69-131, :33-66 (Python loops around
 *           iou_poly, ops/nms_poly.py:247-252 = shapely polygon intersection; restated for convex quadrilaterals in binary64).
 * dets (n, 9) = 4 corner points + score; order (n,) from jdet_argsort_desc(scores); keep (n,) bytes at original indices.
 * fast != 0: the bounding-box pre-filter of the _fast variant.  Suppresses on iou > iou_threshold (a double, >= 0).       */
size_t jdet_nms_poly_workspace_bytes(int n);
int jdet_nms_poly(const float* dets, int n, const int* order, double iou_threshold, int fast, unsigned char* keep,
                  void* workspace, size_t workspace_bytes, void* stream);

/* fixed-size detection record for the end-of-step all-gather: rows [x,y,w,h,theta,score,label] of the kept
 * boxes in descending score order, zero padded to max_per_img rows, + a last row holding the count — what the tail of
 * multiclass_nms_rotated (ops/nms_rotated.py:584-596: re-sort by score, [:max_num]) leaves per image.  One launch, written
 * straight into the caller's send buffer.  dets (n, box_length) [label in column 5 when box_length == 6]; order (n,) from
 * jdet_argsort_desc(scores); keep (n,) bytes from jdet_nms_rotated; record (max_per_img + 1, 7) fp32, fully written.       */
int jdet_pack_detections(const float* dets, int n, int box_length, const float* scores, const int* order,
                         const unsigned char* keep, int max_per_img, float* record, void* stream);
/* one image of a batch whose boxes went through ONE jdet_nms_rotated call with labels offset per image (label' = image *
 * num_classes + label, dets (n,6)): only boxes with label_lo <= label' < label_hi, written with label' - label_lo.       */
int jdet_pack_detections_range(const float* dets, int n, const float* scores, const int* order, const unsigned char* keep,
                               float label_lo, float label_hi, int max_per_img, float* record, void* stream);

/* ---- roi_align_rotated -----------------------------------------------------------------------
 * replaces: version 1: _RotatedROIAlign_v1.execute ops/roi_align_rotated_v1.py:300-326 (kernel :70-147)
 *           version 0: _RotatedROIAlign.execute    ops/roi_align_rotated.py:257-283   (kernel :60-127)
 * input (B,C,H,W); rois (R,6) = [batch, x_ctr, y_ctr, w, h, theta]; output (R,C,PH,PW).
 * sampling_ratio: the int the reference kernel receives (<= 0: adaptive ceil(roi/pooled) grid).
 * Dense RoI sets take the staged path (workspace = channel-last copy of the map + per-RoI tap tables + the gather's
 * work queue: a 256-B counter block, zeroed by the call itself with one memset node, and 8 x R RoI indices); it
 * needs H*W*C < 2^30 per image (tap tables hold 32-bit byte offsets), else cudaErrorInvalidConfiguration.        */
size_t jdet_roi_align_rotated_workspace_bytes(int B, int C, int H, int W, int R, int PH, int PW,
                                              int sampling_ratio);
int jdet_roi_align_rotated(int version, const float* input, int B, int C, int H, int W, const float* rois,
                           int R, int PH, int PW, float spatial_scale, int sampling_ratio, float* output,
                           void* workspace, size_t workspace_bytes, void* stream);

/* same op for a feature map that is ALREADY channel-last in memory, (B,H,W,C) — e.g. a torch channels_last
 * tensor out of the FPN convolutions: no re-layout pass; the workspace only holds the per-RoI tap tables and the work queue.  Needs C % 64 == 0, sampling_ratio > 0 and
 * PH*PW*sampling_ratio^2 <= 1024; otherwise JDET_ERR_UNSUPPORTED (callers fall back to the NCHW entry point). */
size_t jdet_roi_align_rotated_nhwc_workspace_bytes(int R, int PH, int PW, int sampling_ratio);
int jdet_roi_align_rotated_nhwc(int version, const float* input_nhwc, int B, int C, int H, int W, const float* rois,
                                int R, int PH, int PW, float spatial_scale, int sampling_ratio, float* output,
                                void* workspace, size_t workspace_bytes, void* stream);

/* the whole single-level-per-RoI FPN extractor in one call
 * replaces: OrientedSingleRoIExtractor.execute models/roi_extractors/oriented_single_level.py:91-114 (map_roi_levels
 *           :67-70, roi_rescale :72-88) and RboxSingleRoIExtractor.execute (rbox_single_level.py), i.e. per level a
 *           boolean mask, a gather, the RoIAlign op and a masked "+=" into a zero-filled output.
 * feats: HOST array of nlevels device pointers, each (B,C,Hs[l],Ws[l]) fp32 NCHW; Hs/Ws/scales: HOST arrays.
 * A RoI (R,6) is stretched by (ext_w, ext_h), mapped to level clamp(floor(log2(sqrt(w*h)/finest_scale + 1e-6)), 0,
 * nlevels-1), stretched by (rs_w, rs_h) and pooled at scales[level]; output (R,C,PH,PW) rows in RoI order.
 * Needs C % 64 == 0, sampling_ratio > 0, PH*PW*sampling_ratio^2 <= 1024, nlevels <= 8 (else JDET_ERR_UNSUPPORTED). */
size_t jdet_roi_align_rotated_fpn_workspace_bytes(int nlevels, int B, int C, const int* Hs, const int* Ws, int R, int PH,
                                                  int PW, int sampling_ratio);
int jdet_roi_align_rotated_fpn(int version, const float* const* feats, int nlevels, int B, int C, const int* Hs,
                               const int* Ws, const float* scales, const float* rois, int R, int PH, int PW,
                               int sampling_ratio, float ext_w, float ext_h, float rs_w, float rs_h, float finest_scale,
                               float* output, void* workspace, size_t workspace_bytes, void* stream);

/* backward w.r.t. input — replaces: _RotatedROIAlign[_v1].grad ops/roi_align_rotated_v1.py:328-351 (kernel :192-298),
 * ops/roi_align_rotated.py:285-308 (kernel :164-255).  grad_output (R,C,PH,PW) -> grad_input (B,C,H,W), fully written. */
size_t jdet_roi_align_rotated_backward_workspace_bytes(int B, int C, int H, int W, int R, int PH, int PW,
                                                       int sampling_ratio);
int jdet_roi_align_rotated_backward(int version, const float* grad_output, const float* rois, int R, int B, int C,
                                    int H, int W, int PH, int PW, float spatial_scale, int sampling_ratio,
                                    float* grad_input, void* workspace, size_t workspace_bytes, void* stream);

/* ---- feature_refine (rotated_feature_align) --------------------------------------------------
 * replaces: feature_refine_forward ops/fr.py:234-240 (kernel :114-165), FeatureRefineFunction.execute :257-264
 * features (N,C,H,W); best_rbboxes (N,H,W,5); output (N,C,H,W); points: 1 | 5.                     */
int jdet_feature_refine(const float* features, const float* best_rbboxes, int N, int C, int H, int W,
                        int points, float spatial_scale, float* output, void* stream);

/* feature_refine on every FPN level of one head in ONE call — FeatureRefineModule.execute (ops/fr.py:331-347) applies FR level by
 * level; here the levels that qualify for the staged path (W % 4 == 0, a row band fits a CTA: W <= 1024) share one launch.  features / best_rbboxes /
 * outputs / Hs / Ws / scales: HOST arrays of nlevels (<= 8) entries (device pointers inside); N and C are common.          */
int jdet_feature_refine_multi(const float* const* features, const float* const* best_rbboxes, int nlevels, int N, int C,
                              const int* Hs, const int* Ws, const float* scales, int points, float* const* outputs, void* stream);

/* backward w.r.t. features — replaces: FeatureRefineFunction.grad / feature_refine_backward ops/fr.py:242-252,266-271
 * (kernel :167-232).  grad_output (N,C,H,W) -> grad_input (N,C,H,W), fully written. */
int jdet_feature_refine_backward(const float* grad_output, const float* best_rbboxes, int N, int C, int H, int W,
                                 int points, float spatial_scale, float* grad_input, void* stream);

/* ---- AlignConv / DeformConv v1 forward -------------------------------------------------------
 * replaces: AlignConv.get_offset models/roi_heads/s2anet_head.py:677-713 (batched over images :716-721)
 * anchors (N,H,W,5) image space -> offset (N, 2*k*k, H, W), channel 2t = dy, 2t+1 = dx, t = i*k + j */
int jdet_align_conv_offset(const float* anchors, int N, int H, int W, float stride, int kernel_size,
                           float* offset, void* stream);

/* replaces: DeformConvFunction.execute / deform_conv_forward_cuda ops/dcn_v1.py:412-454, 561-600
 *           (deformable_im2col_gpu_kernel :131-184 + jt.matmul :446-447; no columns tensor here).
 * x (B,C,H,W); offset (B, dg*2*kh*kw, Ho, Wo); weight (Co, C/groups, kh, kw); out (B,Co,Ho,Wo).
 * relu != 0 fuses AlignConv's ReLU (s2anet_head.py:722).                                           */
int jdet_deform_conv_forward(const float* x, const float* offset, const float* weight, int B, int C, int H,
                             int W, int Co, int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w,
                             int dil_h, int dil_w, int groups, int deformable_groups, int relu, float* out,
                             void* stream);

/* DeformConv v1 backward building blocks — replace deformable_im2col / deformable_col2im / deformable_col2im_coord
 * (ops/dcn_v1.py:309-410, kernels :131-306).  columns / col_grad: (C*kh*kw, B, Ho, Wo).  The two GEMMs of the backward
 * (W^T x grad_out, grad_out x columns^T — ops/dcn_v1.py:488-489, 545-546) are plain library GEMMs on the caller side. */
int jdet_deform_im2col(const float* x, const float* offset, int B, int C, int H, int W, int kh, int kw, int stride_h,
                       int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int deformable_groups, float* columns,
                       void* stream);
int jdet_deform_col2im(const float* col_grad, const float* offset, int B, int C, int H, int W, int kh, int kw,
                       int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int deformable_groups,
                       float* grad_x, void* stream);
int jdet_deform_col2im_coord(const float* col_grad, const float* x, const float* offset, int B, int C, int H, int W, int kh,
                             int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                             int deformable_groups, float* grad_offset, void* stream);

/* replaces: AlignConv.execute models/roi_heads/s2anet_head.py:715-723 as ONE fused call:
 * offsets from anchors -> deformable 3x3 sampling -> tcgen05 GEMM (3xTF32 split, fp32-class
 * accuracy) -> ReLU.  x (N,C,H,W); anchors (N,H,W,5); weight (Co,C,3,3); out (N,Co,H,W).
 * tcgen05 path: C % 16 == 0, Co % 32 == 0, Co <= 256; other shapes compose offset + generic deform conv. */
size_t jdet_align_conv_forward_workspace_bytes(int N, int C, int H, int W, int Co);
int jdet_align_conv_forward(const float* x, const float* anchors, const float* weight, int N, int C, int H,
                            int W, int Co, float stride, float* out, void* workspace, size_t workspace_bytes,
                            void* stream);

/* the same for every FPN level of one head in ONE call: S2ANetHead applies its AlignConv to 5 levels (s2anet_head.py:230-237);
 * one persistent tcgen05 launch covers all levels' tiles (the 4 / 16 / 64 tiles of the coarse levels fill the last wave
 * instead of costing a launch each) and the weight is split once.  xs / anchors / outs / Hs / Ws / strides: HOST arrays of
 * nlevels (<= 8) entries.  x_channels_last != 0: the maps are already (N,H,W,C) in memory (torch.channels_last) and are sampled in
 * place.  tcgen05 shape class only (else JDET_ERR_UNSUPPORTED: loop over jdet_align_conv_forward).                              */
size_t jdet_align_conv_forward_multi_workspace_bytes(int nlevels, int N, int C, const int* Hs, const int* Ws, int Co);
int jdet_align_conv_forward_multi(const float* const* xs, const float* const* anchors, const float* weight, int nlevels,
                                  int N, int C, const int* Hs, const int* Ws, int Co, const float* strides,
                                  float* const* outs, int x_channels_last, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JDET_B200_H_ */
