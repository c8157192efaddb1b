#!/usr/bin/env python
"""Build ``oracle/_ref/`` — the reference's OWN source compiled as test infrastructure.

TEST INFRASTRUCTURE ONLY.  Nothing under ``jdet_b200/`` may import or link this.

JDet ships its native code as C++/CUDA *source strings* inside ``python/jdet/ops/*.py``
(handed to Jittor's ``jt.code`` JIT at run time).  Jittor is not installable here, but the
strings are plain C++/CUDA.  This script

  1. parses the reference ``.py`` files with ``ast`` (no import, no jittor needed) and
     evaluates the module-level string constants (``IOU_ROTATED_CPU_HEADER`` ...),
  2. drops the single ``#include <executor.h>`` line (a Jittor header the kernels never use),
  3. wraps every header in its own C++ namespace plus a few lines of ``extern "C"`` glue that
     mimics what ``jt.code`` would have generated (``in0_p`` / ``out0_p`` aliases, launch
     geometry copied from the reference launch snippets, cited per function),
  4. compiles  ``oracle/_ref/libref_cpu.so``  (g++, IoU v0/v1 + NMS: the reference CPU path)
     and       ``oracle/_ref/libref_cuda.so`` (nvcc sm_100a: all five reference CUDA kernels,
     plus host-callable copies of the ``__host__ __device__`` CUDA-variant IoU).

The generated translation units live in a temp dir and are deleted; only the ``.so`` files are
kept, under ``oracle/_ref/`` which is git-ignored.  No reference source is copied into the repo.

Runs only where ``/root/reference`` exists (the build container).  The GPU box uses the prebuilt
``.so`` files that travel with the snapshot.
"""
import ast
import os
import shutil
import subprocess
import sys
import tempfile

REF_ROOT = os.environ.get("JDET_REFERENCE_ROOT", "/root/reference")
REF_OPS = os.path.join(REF_ROOT, "python", "jdet", "ops")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

CPU_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off"]
# The reference is JIT-compiled by nvcc with its defaults (fmad on).  Keep that for the device
# kernels: this library is "the reference kernel on the same GPU".
NVCC_FLAGS = ["-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
              "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-ffp-contract=off"]


def module_strings(path):
    """Evaluate module-level ``NAME = <str expr>`` assignments of a reference .py file."""
    with open(path, "r") as f:
        tree = ast.parse(f.read(), filename=path)
    env = {}

    def ev(node):
        if isinstance(node, ast.Constant) and isinstance(node.value, str):
            return node.value
        if isinstance(node, ast.Name) and node.id in env:
            return env[node.id]
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Add):
            return ev(node.left) + ev(node.right)
        raise ValueError("not a constant string expression")

    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 \
                and isinstance(node.targets[0], ast.Name):
            try:
                env[node.targets[0].id] = ev(node.value)
            except ValueError:
                pass
    return env


def strip_jittor(src):
    keep = []
    for line in src.splitlines():
        s = line.strip().replace(" ", "")
        if s.startswith("#include<executor.h>"):
            continue
        keep.append(line)
    return "\n".join(keep)


def strip_alias(src):
    return "\n".join(l for l in src.splitlines() if not l.strip().startswith("@alias"))


PRELUDE = r"""
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <climits>
#include <cfloat>
#include <cstdint>
#include <algorithm>
#include <vector>
#include <math.h>
#include <stdio.h>
#include <float.h>
using std::min; using std::max;
"""


# --------------------------------------------------------------------------------------------
# CPU library: reference cpu_header / cpu_src of box_iou_rotated(.py:312-326,487-500),
# box_iou_rotated_v1 and nms_rotated (.py:314-328,414-449), BOX_LENGTH 5 and 6.
# --------------------------------------------------------------------------------------------
def gen_cpu(iou0, iou1, nms, orn=None):
    parts = [PRELUDE]

    def iou_ns(ns, env):
        hdr = strip_jittor(env["IOU_ROTATED_CPU_HEADER"])
        loop = strip_alias(env["IOU_CPU_SRC"])
        return f"""
namespace {ns} {{
{hdr}
// glue: what jt.code would have bound (in0=boxes1, in1=boxes2, out0=ious)
static void run(const float* boxes1_p, int boxes1_shape0, int boxes1_shape1,
                const float* boxes2_p, int boxes2_shape0, float* ious_p) {{
{loop}
}}
static float one(const float* a, const float* b) {{ return single_box_iou_rotated<float>(a, b); }}
}}  // namespace {ns}
"""
    parts.append(iou_ns("ref_iou_v0_cpu", iou0))
    parts.append(iou_ns("ref_iou_v1_cpu", iou1))

    for bl in (5, 6):
        hdr = strip_jittor(nms["ML_NMS_ROTATED_CPU_HEADER"])
        src = strip_alias(nms["ML_NMS_ROTATED_CPU_SRC"])
        parts.append(f"""
#undef BOX_LENGTH
#define BOX_LENGTH {bl}
namespace ref_nms_cpu_{bl} {{
{hdr}
struct VarStub {{ size_t size; }};
// glue for nms_rotated_cpu (ops/nms_rotated.py:495-504): in0=dets in1=order in2=suppressed out0=keep
static void run(const float* dets_p, int dets_shape0, const int* order_t_p,
                unsigned char* suppressed_t_p, bool* keep_t_p, const float iou_threshold) {{
  VarStub keep_obj{{(size_t)dets_shape0 * sizeof(bool)}}; VarStub* keep_t = &keep_obj;
{src}
}}
static float one(const float* a, const float* b) {{ return single_box_iou_rotated<float>(a, b); }}
}}  // namespace
""")
    if orn is not None:
        # ORN (ops/orn.py): ARF_forward_cpu_kernel :136-170, RIE_forward_cpu_kernel :291-330 — the reference's CPU sources
        parts.append(f"""
#include <cfloat>
namespace ref_orn_cpu {{
namespace arf {{
{strip_jittor(orn["ARF_CPU_HEADER"])}
}}
namespace rie {{
{strip_jittor(orn["RIE_CPU_HEADER"])}
}}
}}  // namespace ref_orn_cpu
extern "C" {{
// weight (nOut, nIn, nOri, kH, kW), indices (nOri, kH, kW, nRot) uint8 1-based -> out (nOut*nRot, nIn*nOri, kH, kW), pre-zeroed by the caller
void ref_arf_forward_cpu(const float* w, const unsigned char* ind, int nOut, int nIn, int nOri, int kH, int kW, int nRot, float* out) {{
  ref_orn_cpu::arf::ARF_forward_cpu_kernel<float>(w, ind, nOut, nIn, nOri, kH, kW, nRot, out);
}}
// feature (nBatch, nFeature*nOri) -> mainDirection (nBatch, nFeature) uint8, aligned (nBatch, nFeature*nOri)
void ref_rie_forward_cpu(const float* f, unsigned char* dir, float* aligned, int nOri, int nBatch, int nFeature) {{
  ref_orn_cpu::rie::RIE_forward_cpu_kernel<float>(f, dir, aligned, nOri, nBatch, nFeature);
}}
}}
""")
    parts.append(r"""
extern "C" {
void ref_box_iou_rotated_cpu(const float* b1, int n1, const float* b2, int n2, float* out) {
  ref_iou_v0_cpu::run(b1, n1, 5, b2, n2, out);
}
void ref_box_iou_rotated_v1_cpu(const float* b1, int n1, const float* b2, int n2, float* out) {
  ref_iou_v1_cpu::run(b1, n1, 5, b2, n2, out);
}
float ref_single_iou_v0_cpu(const float* a, const float* b) { return ref_iou_v0_cpu::one(a, b); }
float ref_single_iou_v1_cpu(const float* a, const float* b) { return ref_iou_v1_cpu::one(a, b); }
float ref_single_iou_nms5_cpu(const float* a, const float* b) { return ref_nms_cpu_5::one(a, b); }
float ref_single_iou_nms6_cpu(const float* a, const float* b) { return ref_nms_cpu_6::one(a, b); }
// keep: n bools (written), suppressed: n bytes scratch (zeroed here, as jt.zeros does)
void ref_nms_rotated_cpu(const float* dets, int n, int box_length, const int* order,
                         float thr, unsigned char* suppressed, bool* keep) {
  memset(suppressed, 0, (size_t)n);
  if (box_length == 5) ref_nms_cpu_5::run(dets, n, order, suppressed, keep, thr);
  else                 ref_nms_cpu_6::run(dets, n, order, suppressed, keep, thr);
}
}
""")
    return "\n".join(parts)


# --------------------------------------------------------------------------------------------
# CUDA library: the five reference kernels, launched exactly as the reference launch snippets do.
# --------------------------------------------------------------------------------------------
def gen_cuda(iou0, iou1, nms, ra0, ra1, fr, dcn):
    parts = [PRELUDE, "#include <cuda_runtime.h>\n"]

    def iou_ns(ns, env):
        hdr = strip_jittor(env["IOU_ROTATED_CUDA_HEADER"])
        return f"""
namespace {ns} {{
{hdr}
// launch geometry: ops/box_iou_rotated.py:464-485
static void launch(const float* boxes1_p, int num_boxes1, const float* boxes2_p, int num_boxes2,
                   float* ious_p, cudaStream_t st) {{
  if (num_boxes1 > 0 && num_boxes2 > 0) {{
    dim3 blocks(CeilDIV(num_boxes1, BLOCK_DIM_X), CeilDIV(num_boxes2, BLOCK_DIM_Y));
    dim3 threads(BLOCK_DIM_X, BLOCK_DIM_Y);
    box_iou_rotated_cuda_kernel<float><<<blocks, threads, 0, st>>>(num_boxes1, num_boxes2,
        boxes1_p, boxes2_p, ious_p);
  }}
}}
static float one_host(const float* a, const float* b) {{ return single_box_iou_rotated<float>(a, b); }}
}}  // namespace {ns}
"""
    parts.append(iou_ns("ref_iou_v0_cuda", iou0))
    parts.append(iou_ns("ref_iou_v1_cuda", iou1))

    for bl in (5, 6):
        hdr = strip_jittor(nms["ML_NMS_ROTATED_CUDA_HEADER"])
        parts.append(f"""
#undef BOX_LENGTH
#define BOX_LENGTH {bl}
namespace ref_nms_cuda_{bl} {{
{hdr}
// launch + host reduce: ops/nms_rotated.py:450-493.  The reference reads the device mask from the
// host through managed memory; here it is copied back explicitly, the reduce loop is the same.
static int run(const float* dets_sorted_p, int dets_num, const int* order_host, float iou_threshold,
               bool* keep_host, float* mask_ms) {{
  memset(keep_host, 0, (size_t)dets_num);
  if (dets_num == 0) return 0;
  const int col_blocks = CeilDIV(dets_num, threadsPerBlock);
  size_t matrices_size = (size_t)dets_num * col_blocks * sizeof(unsigned long long);
  unsigned long long* mask_p = nullptr;
  cudaError_t e = cudaMalloc(&mask_p, matrices_size);
  if (e != cudaSuccess) return (int)e;
  dim3 blocks(col_blocks, col_blocks);
  dim3 threads(threadsPerBlock);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  nms_rotated_cuda_kernel<float><<<blocks, threads, 0>>>(dets_num, iou_threshold, dets_sorted_p, mask_p);
  cudaEventRecord(e1, 0);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {{ cudaFree(mask_p); return (int)e; }}
  if (mask_ms) cudaEventElapsedTime(mask_ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  std::vector<unsigned long long> mask_h((size_t)dets_num * col_blocks);
  cudaMemcpy(mask_h.data(), mask_p, matrices_size, cudaMemcpyDeviceToHost);
  cudaFree(mask_p);
  std::vector<unsigned long long> remv(col_blocks);
  memset(&remv[0], 0, sizeof(unsigned long long) * col_blocks);
  for (int i = 0; i < dets_num; i++) {{
    int nblock = i / threadsPerBlock;
    int inblock = i % threadsPerBlock;
    if (!(remv[nblock] & (1ULL << inblock))) {{
      keep_host[order_host[i]] = true;
      unsigned long long* p = mask_h.data() + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; j++) remv[j] |= p[j];
    }}
  }}
  return 0;
}}
static float one_host(const float* a, const float* b) {{ return single_box_iou_rotated<float>(a, b); }}
}}  // namespace
""")

    def roi_ns(ns, env):
        hdr = env["CUDA_HEADER"]
        return f"""
namespace {ns} {{
{hdr}
// launch: ops/roi_align_rotated*.py  (GET_BLOCKS(output_size), THREADS_PER_BLOCK); sampling_ratio is a
// float constant passed to an int parameter in the reference (truncation) -> int here.
static void launch(const float* input_p, const float* rois_p, int num_rois, int channels, int height,
                   int width, int pooled_height, int pooled_width, float spatial_scale,
                   float sampling_ratio, float* output_p, cudaStream_t st) {{
  int output_size = num_rois * pooled_height * pooled_width * channels;
  if (output_size == 0) return;
  ROIAlignRotatedForward<float><<<GET_BLOCKS(output_size), THREADS_PER_BLOCK, 0, st>>>(
      output_size, input_p, rois_p, spatial_scale, sampling_ratio, channels, height, width,
      pooled_height, pooled_width, output_p);
}}
// backward launch: ops/roi_align_rotated*.py grad() (memset of grad_input, then ROIAlignBackward)
static void launch_bwd(const float* grad_p, const float* rois_p, int num_rois, int batch, int channels, int height,
                       int width, int pooled_height, int pooled_width, float spatial_scale,
                       float sampling_ratio, float* grad_input_p, cudaStream_t st) {{
  cudaMemsetAsync(grad_input_p, 0, (size_t)batch * channels * height * width * 4, st);
  int output_size = num_rois * pooled_height * pooled_width * channels;
  if (output_size == 0) return;
  ROIAlignBackward<float><<<GET_BLOCKS(output_size), THREADS_PER_BLOCK, 0, st>>>(
      output_size, grad_p, rois_p, spatial_scale, sampling_ratio, channels, height, width,
      pooled_height, pooled_width, grad_input_p);
}}
}}  // namespace {ns}
#undef CUDA_1D_KERNEL_LOOP
#undef THREADS_PER_BLOCK
"""
    parts.append(roi_ns("ref_roi_v0", ra0))
    parts.append(roi_ns("ref_roi_v1", ra1))

    parts.append(f"""
namespace ref_fr {{
{fr["HEADER"]}
// launch: ops/fr.py:234-240
static void launch(const float* in0_p, const float* in1_p, int n, int c, int h, int w, int points,
                   float spatial_scale, float* out0_p, cudaStream_t st) {{
  const int output_size = n * c * h * w;
  if (output_size == 0) return;
  feature_refine_forward_kernel<float><<<GET_BLOCKS(output_size), THREADS_PER_BLOCK, 0, st>>>(
      output_size, points, in0_p, in1_p, spatial_scale, c, h, w, out0_p);
}}
// backward: ops/fr.py:242-252 (zeros_like, then feature_refine_backward_kernel)
static void launch_bwd(const float* in0_p, const float* in1_p, int n, int c, int h, int w, int points,
                       float spatial_scale, float* out0_p, cudaStream_t st) {{
  const int output_size = n * c * h * w;
  if (output_size == 0) return;
  cudaMemsetAsync(out0_p, 0, (size_t)output_size * 4, st);
  feature_refine_backward_kernel<float><<<GET_BLOCKS(output_size), THREADS_PER_BLOCK, 0, st>>>(
      output_size, points, in0_p, in1_p, spatial_scale, c, h, w, out0_p);
}}
}}  // namespace ref_fr
#undef CUDA_1D_KERNEL_LOOP
#undef THREADS_PER_BLOCK
""")

    parts.append(f"""
namespace ref_dcn {{
{strip_jittor(dcn["HEADER"])}
// launch: ops/dcn_v1.py:309-339 (memset of columns, then the im2col kernel)
static void launch_im2col(const float* in0_p, const float* in1_p, int channels, int height, int width,
                          int ksize_h, int ksize_w, int pad_h, int pad_w, int stride_h, int stride_w,
                          int dilation_h, int dilation_w, int parallel_imgs, int deformable_group,
                          float* out0_p, cudaStream_t st) {{
  int height_col = (height + 2 * pad_h - (dilation_h * (ksize_h - 1) + 1)) / stride_h + 1;
  int width_col = (width + 2 * pad_w - (dilation_w * (ksize_w - 1) + 1)) / stride_w + 1;
  int num_kernels = channels * height_col * width_col * parallel_imgs;
  int channel_per_deformable_group = channels / deformable_group;
  size_t bytes = (size_t)channels * ksize_h * ksize_w * parallel_imgs * height_col * width_col * 4;
  cudaMemsetAsync(out0_p, 0, bytes, st);
  deformable_im2col_gpu_kernel<float><<<GET_BLOCKS(num_kernels), CUDA_NUM_THREADS, 0, st>>>(
      num_kernels, in0_p, in1_p, height, width, ksize_h, ksize_w, pad_h, pad_w, stride_h, stride_w,
      dilation_h, dilation_w, channel_per_deformable_group, parallel_imgs, channels,
      deformable_group, height_col, width_col, out0_p);
}}
// launches of deformable_col2im / deformable_col2im_coord: ops/dcn_v1.py:341-410 (memset + kernel)
static void launch_col2im(const float* in0_p, const float* in1_p, int channels, int height, int width, int ksize_h,
                          int ksize_w, int pad_h, int pad_w, int stride_h, int stride_w, int dilation_h, int dilation_w,
                          int parallel_imgs, int deformable_group, float* out0_p, cudaStream_t st) {{
  int height_col = (height + 2 * pad_h - (dilation_h * (ksize_h - 1) + 1)) / stride_h + 1;
  int width_col = (width + 2 * pad_w - (dilation_w * (ksize_w - 1) + 1)) / stride_w + 1;
  int num_kernels = channels * ksize_h * ksize_w * height_col * width_col * parallel_imgs;
  int channel_per_deformable_group = channels / deformable_group;
  cudaMemsetAsync(out0_p, 0, (size_t)parallel_imgs * channels * height * width * 4, st);
  deformable_col2im_gpu_kernel<float><<<GET_BLOCKS(num_kernels), CUDA_NUM_THREADS, 0, st>>>(
      num_kernels, in0_p, in1_p, channels, height, width, ksize_h, ksize_w, pad_h, pad_w, stride_h, stride_w,
      dilation_h, dilation_w, channel_per_deformable_group, parallel_imgs, deformable_group, height_col, width_col,
      out0_p, 0);
}}
static void launch_col2im_coord(const float* in0_p, const float* in1_p, const float* in2_p, int channels, int height,
                                int width, int ksize_h, int ksize_w, int pad_h, int pad_w, int stride_h, int stride_w,
                                int dilation_h, int dilation_w, int parallel_imgs, int deformable_group, float* out0_p,
                                cudaStream_t st) {{
  int height_col = (height + 2 * pad_h - (dilation_h * (ksize_h - 1) + 1)) / stride_h + 1;
  int width_col = (width + 2 * pad_w - (dilation_w * (ksize_w - 1) + 1)) / stride_w + 1;
  int num_kernels = height_col * width_col * 2 * ksize_h * ksize_w * deformable_group * parallel_imgs;
  int channel_per_deformable_group = channels * ksize_h * ksize_w / deformable_group;
  cudaMemsetAsync(out0_p, 0, (size_t)num_kernels * 4, st);
  deformable_col2im_coord_gpu_kernel<float><<<GET_BLOCKS(num_kernels), CUDA_NUM_THREADS, 0, st>>>(
      num_kernels, in0_p, in1_p, in2_p, channels, height, width, ksize_h, ksize_w, pad_h, pad_w, stride_h, stride_w,
      dilation_h, dilation_w, channel_per_deformable_group, parallel_imgs, 2 * ksize_h * ksize_w * deformable_group,
      deformable_group, height_col, width_col, out0_p);
}}
}}  // namespace ref_dcn
""")

    parts.append(r"""
extern "C" {
// all pointers are DEVICE pointers unless named *_host
int ref_box_iou_rotated_cuda(const float* b1, int n1, const float* b2, int n2, float* out, void* st) {
  ref_iou_v0_cuda::launch(b1, n1, b2, n2, out, (cudaStream_t)st); return (int)cudaGetLastError();
}
int ref_box_iou_rotated_v1_cuda(const float* b1, int n1, const float* b2, int n2, float* out, void* st) {
  ref_iou_v1_cuda::launch(b1, n1, b2, n2, out, (cudaStream_t)st); return (int)cudaGetLastError();
}
// CUDA-variant (exchange-sort hull) executed on the HOST: pins the oracle's "cuda" variant without a GPU
float ref_single_iou_v0_cudavariant_host(const float* a, const float* b) { return ref_iou_v0_cuda::one_host(a, b); }
float ref_single_iou_v1_cudavariant_host(const float* a, const float* b) { return ref_iou_v1_cuda::one_host(a, b); }
float ref_single_iou_nms5_cudavariant_host(const float* a, const float* b) { return ref_nms_cuda_5::one_host(a, b); }
float ref_single_iou_nms6_cudavariant_host(const float* a, const float* b) { return ref_nms_cuda_6::one_host(a, b); }
// dets_sorted: device (n, box_length) already gathered by order; order_host/keep_host: host
int ref_nms_rotated_cuda(const float* dets_sorted, int n, int box_length, const int* order_host,
                         float thr, bool* keep_host, float* mask_ms) {
  if (box_length == 5) return ref_nms_cuda_5::run(dets_sorted, n, order_host, thr, keep_host, mask_ms);
  return ref_nms_cuda_6::run(dets_sorted, n, order_host, thr, keep_host, mask_ms);
}
int ref_roi_align_rotated_cuda(int version, const float* input, const float* rois, int num_rois,
                               int channels, int height, int width, int ph, int pw,
                               float spatial_scale, float sampling_ratio, float* out, void* st) {
  if (version == 0) ref_roi_v0::launch(input, rois, num_rois, channels, height, width, ph, pw,
                                       spatial_scale, sampling_ratio, out, (cudaStream_t)st);
  else              ref_roi_v1::launch(input, rois, num_rois, channels, height, width, ph, pw,
                                       spatial_scale, sampling_ratio, out, (cudaStream_t)st);
  return (int)cudaGetLastError();
}
int ref_roi_align_rotated_backward_cuda(int version, const float* grad, const float* rois, int num_rois, int batch,
                                        int channels, int height, int width, int ph, int pw,
                                        float spatial_scale, float sampling_ratio, float* grad_input, void* st) {
  if (version == 0) ref_roi_v0::launch_bwd(grad, rois, num_rois, batch, channels, height, width, ph, pw,
                                           spatial_scale, sampling_ratio, grad_input, (cudaStream_t)st);
  else              ref_roi_v1::launch_bwd(grad, rois, num_rois, batch, channels, height, width, ph, pw,
                                           spatial_scale, sampling_ratio, grad_input, (cudaStream_t)st);
  return (int)cudaGetLastError();
}
int ref_feature_refine_backward_cuda(const float* grad, const float* boxes, int n, int c, int h, int w,
                                     int points, float spatial_scale, float* grad_in, void* st) {
  ref_fr::launch_bwd(grad, boxes, n, c, h, w, points, spatial_scale, grad_in, (cudaStream_t)st);
  return (int)cudaGetLastError();
}
int ref_feature_refine_cuda(const float* feat, const float* boxes, int n, int c, int h, int w,
                            int points, float spatial_scale, float* out, void* st) {
  ref_fr::launch(feat, boxes, n, c, h, w, points, spatial_scale, out, (cudaStream_t)st);
  return (int)cudaGetLastError();
}
int ref_deformable_col2im_cuda(const float* col, const float* offset, int channels, int height, int width, int kh, int kw,
                               int pad_h, int pad_w, int stride_h, int stride_w, int dil_h, int dil_w, int parallel_imgs,
                               int deformable_group, float* grad_im, void* st) {
  ref_dcn::launch_col2im(col, offset, channels, height, width, kh, kw, pad_h, pad_w, stride_h, stride_w, dil_h, dil_w,
                         parallel_imgs, deformable_group, grad_im, (cudaStream_t)st);
  return (int)cudaGetLastError();
}
int ref_deformable_col2im_coord_cuda(const float* col, const float* im, const float* offset, int channels, int height,
                                     int width, int kh, int kw, int pad_h, int pad_w, int stride_h, int stride_w,
                                     int dil_h, int dil_w, int parallel_imgs, int deformable_group, float* grad_offset,
                                     void* st) {
  ref_dcn::launch_col2im_coord(col, im, offset, channels, height, width, kh, kw, pad_h, pad_w, stride_h, stride_w,
                               dil_h, dil_w, parallel_imgs, deformable_group, grad_offset, (cudaStream_t)st);
  return (int)cudaGetLastError();
}
int ref_deformable_im2col_cuda(const float* im, const float* offset, int channels, int height,
                               int width, int kh, int kw, int pad_h, int pad_w, int stride_h,
                               int stride_w, int dil_h, int dil_w, int parallel_imgs,
                               int deformable_group, float* columns, void* st) {
  ref_dcn::launch_im2col(im, offset, channels, height, width, kh, kw, pad_h, pad_w, stride_h,
                         stride_w, dil_h, dil_w, parallel_imgs, deformable_group, columns,
                         (cudaStream_t)st);
  return (int)cudaGetLastError();
}
}
""")
    return "\n".join(parts)


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False, verbose=True):
    """Returns True if oracle/_ref/*.so exist (built now or earlier)."""
    cpu_so = os.path.join(OUT, "libref_cpu.so")
    cuda_so = os.path.join(OUT, "libref_cuda.so")
    if not os.path.isdir(REF_OPS):
        ok = os.path.exists(cpu_so) and os.path.exists(cuda_so)
        if verbose:
            print(f"[build_ref] {REF_OPS} not present; using prebuilt oracle/_ref: {ok}")
        return ok
    if not force and os.path.exists(cpu_so) and os.path.exists(cuda_so) \
            and min(os.path.getmtime(cpu_so), os.path.getmtime(cuda_so)) > os.path.getmtime(__file__):
        return True
    os.makedirs(OUT, exist_ok=True)
    iou0 = module_strings(os.path.join(REF_OPS, "box_iou_rotated.py"))
    iou1 = module_strings(os.path.join(REF_OPS, "box_iou_rotated_v1.py"))
    nms = module_strings(os.path.join(REF_OPS, "nms_rotated.py"))
    ra0 = module_strings(os.path.join(REF_OPS, "roi_align_rotated.py"))
    ra1 = module_strings(os.path.join(REF_OPS, "roi_align_rotated_v1.py"))
    fr = module_strings(os.path.join(REF_OPS, "fr.py"))
    dcn = module_strings(os.path.join(REF_OPS, "dcn_v1.py"))
    orn = module_strings(os.path.join(REF_OPS, "orn.py"))
    tmp = tempfile.mkdtemp(prefix="jdet_ref_build_")
    try:
        cpu_cc = os.path.join(tmp, "ref_cpu.cc")
        with open(cpu_cc, "w") as f:
            f.write(gen_cpu(iou0, iou1, nms, orn))
        run(["g++"] + CPU_FLAGS + [cpu_cc, "-o", cpu_so])
        cuda_cu = os.path.join(tmp, "ref_cuda.cu")
        with open(cuda_cu, "w") as f:
            f.write(gen_cuda(iou0, iou1, nms, ra0, ra1, fr, dcn))
        run(["nvcc"] + NVCC_FLAGS + [cuda_cu, "-o", cuda_so])
    finally:
        if os.environ.get("JDET_KEEP_REF_TMP"):
            print("[build_ref] kept", tmp)
        else:
            shutil.rmtree(tmp, ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
