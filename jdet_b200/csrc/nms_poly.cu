// nms_poly.cu — tile -> image merge NMS over general (convex) quadrilaterals, on the GPU.
//
// Replaces py_cpu_nms_poly_fast / py_cpu_nms_poly (data/devkits/result_merge.py:69-131, :33-66): a Python loop over the
// detections of one class of one image that (fast variant) pre-filters by bounding-box overlap (`hbb_ovr > 0`, :100-106),
// evaluates iou_poly (ops/nms_poly.py:247-252: shapely Polygon intersection, iou = inter / max(a1 + a2 - inter, 0.01)) on
// the survivors and drops everything with iou > thresh (:124: `inds = where(hbb_ovr <= thresh)`).
//
// shapely / GEOS is a third-party dependency that is not under /root/reference (version unpinned, requirements.txt); its
// arithmetic is restated here as what it computes for valid convex quadrilaterals — the area of the intersection polygon,
// by clipping one quadrilateral against the four half-planes of the other in binary64 — see oracle/glue.py iou_poly for
// the CPU restatement the tests compare against.  PARITY UNPINNED against shapely itself (absent); pinned against
// box_iou_rotated on rectangles and against analytic cases.
//
//   poly_rec_kernel   per box: orientation-normalised corners (counter-clockwise), bounding box, area, in score order
//   poly_mask_kernel  64 x 64 tiles of the upper triangle: bit (i, j) = iou_poly(i, j) > thr  (i before j in score order)
//   poly_scan_kernel  one CTA: the greedy pass over the mask (the reference's host loop), keep flags at original indices
#include <algorithm>
#include "common.cuh"

namespace jdet {
namespace poly {

struct Rec { double x[4], y[4]; double x1, y1, x2, y2, area; };   // 13 doubles

__device__ __forceinline__ double signed_area4(const double* x, const double* y) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; i++) { const int j = (i + 1) & 3; s += x[i] * y[j] - x[j] * y[i]; }
  return 0.5 * s;
}

__global__ void __launch_bounds__(256) poly_rec_kernel(const float* __restrict__ dets, const int* __restrict__ order, int n, Rec* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* d = dets + (size_t)order[i] * 9;
  Rec r;
#pragma unroll
  for (int k = 0; k < 4; k++) { r.x[k] = (double)d[2 * k]; r.y[k] = (double)d[2 * k + 1]; }
  double a = signed_area4(r.x, r.y);
  if (a < 0.0) {                                   // clockwise -> counter-clockwise (swap corners 1 and 3)
    double t = r.x[1]; r.x[1] = r.x[3]; r.x[3] = t;
    t = r.y[1]; r.y[1] = r.y[3]; r.y[3] = t;
    a = -a;
  }
  r.area = a;
  r.x1 = fmin(fmin(r.x[0], r.x[1]), fmin(r.x[2], r.x[3])); r.x2 = fmax(fmax(r.x[0], r.x[1]), fmax(r.x[2], r.x[3]));
  r.y1 = fmin(fmin(r.y[0], r.y[1]), fmin(r.y[2], r.y[3])); r.y2 = fmax(fmax(r.y[0], r.y[1]), fmax(r.y[2], r.y[3]));
  rec[i] = r;
}

// area of (convex, counter-clockwise) A clipped by the four half-planes of (convex, counter-clockwise) B
__device__ inline double inter_area(const Rec& A, const Rec& B) {
  double px[10], py[10], qx[10], qy[10];
  int n = 4;
#pragma unroll
  for (int i = 0; i < 4; i++) { px[i] = A.x[i]; py[i] = A.y[i]; }
  for (int e = 0; e < 4 && n > 0; e++) {
    const double ax = B.x[e], ay = B.y[e], bx = B.x[(e + 1) & 3], by = B.y[(e + 1) & 3];
    const double ex = bx - ax, ey = by - ay;
    int m = 0;
    for (int i = 0; i < n; i++) {
      const int j = i + 1 == n ? 0 : i + 1;
      const double si = ex * (py[i] - ay) - ey * (px[i] - ax);      // >= 0: inside (left of the edge)
      const double sj = ex * (py[j] - ay) - ey * (px[j] - ax);
      if (si >= 0.0) { qx[m] = px[i]; qy[m] = py[i]; m++; }
      if ((si > 0.0 && sj < 0.0) || (si < 0.0 && sj > 0.0)) {
        const double t = si / (si - sj);
        qx[m] = px[i] + t * (px[j] - px[i]); qy[m] = py[i] + t * (py[j] - py[i]); m++;
      }
    }
    n = m;
    for (int i = 0; i < n; i++) { px[i] = qx[i]; py[i] = qy[i]; }
  }
  if (n < 3) return 0.0;
  double s = 0.0;
  for (int i = 0; i < n; i++) { const int j = i + 1 == n ? 0 : i + 1; s += px[i] * py[j] - px[j] * py[i]; }
  return fabs(0.5 * s);
}

__device__ __forceinline__ double iou_poly(const Rec& A, const Rec& B) {
  const double inter = inter_area(A, B);
  return inter / fmax(A.area + B.area - inter, 0.01);               // ops/nms_poly.py:251
}

// grid = (col blocks, row blocks), upper triangle only; 64 threads: thread t = row t of the tile
__global__ void __launch_bounds__(64) poly_mask_kernel(const Rec* __restrict__ rec, int n, double thr, int fast,
                                                        unsigned long long* __restrict__ mask) {
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb) return;
  const int col_blocks = (n + 63) / 64;
  __shared__ Rec s_col[64];
  const int cj = cb * 64 + threadIdx.x;
  if (cj < n) s_col[threadIdx.x] = rec[cj];
  __syncthreads();
  const int ri = rb * 64 + threadIdx.x;
  if (ri >= n) return;
  const Rec A = rec[ri];
  const int ncol = min(64, n - cb * 64);
  unsigned long long bits = 0ull;
  for (int j = (rb == cb ? threadIdx.x + 1 : 0); j < ncol; j++) {
    const Rec& B = s_col[j];
    if (fast) {   // result_merge.py:93-106: only pairs whose bounding boxes overlap with positive area are evaluated
      const double w = fmax(0.0, fmin(A.x2, B.x2) - fmax(A.x1, B.x1)), h = fmax(0.0, fmin(A.y2, B.y2) - fmax(A.y1, B.y1));
      if (!(w * h > 0.0)) continue;
    }
    if (iou_poly(A, B) > thr) bits |= 1ull << j;
  }
  mask[(size_t)ri * col_blocks + cb] = bits;
}

// one CTA, the reference's greedy pass: box i (score order) is kept unless an earlier kept box set its bit
__global__ void __launch_bounds__(1024) poly_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n,
                                                         unsigned char* __restrict__ keep) {
  extern __shared__ unsigned long long remv[];       // col_blocks words
  const int col_blocks = (n + 63) / 64;
  for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
  for (int i = threadIdx.x; i < n; i += blockDim.x) keep[i] = 0;
  __syncthreads();
  for (int i = 0; i < n; i++) {
    const int nb = i >> 6;
    const bool alive = !((remv[nb] >> (i & 63)) & 1ull);   // every thread reads the same word: uniform branch
    __syncthreads();
    if (alive) {
      if (threadIdx.x == 0) keep[order[i]] = 1;
      const unsigned long long* row = mask + (size_t)i * col_blocks;
      for (int j = nb + threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] |= row[j];
      __syncthreads();
    }
  }
}

}  // namespace poly
}  // namespace jdet

JDET_API size_t jdet_nms_poly_workspace_bytes(int n) {
  if (n <= 0) return 256;
  const size_t cb = ((size_t)n + 63) / 64;
  return jdet_align_up((size_t)n * sizeof(jdet::poly::Rec), 256) + jdet_align_up((size_t)n * cb * 8, 256);
}

// dets (n, 9) = 4 corner points + score; order (n,) = jdet_argsort_desc(scores); keep (n,) bytes at ORIGINAL indices.
// fast != 0: py_cpu_nms_poly_fast (bounding-box pre-filter); 0: py_cpu_nms_poly (every pair).  Suppress on iou > thr, compared in
// binary64 like the reference's Python floats (iou_threshold >= 0).
JDET_API int jdet_nms_poly(const float* dets, int n, const int* order, double iou_threshold, int fast, unsigned char* keep,
                           void* workspace, size_t workspace_bytes, void* stream) {
  using namespace jdet::poly;
  if (n < 0) return JDET_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!dets || !order || !keep || !workspace) return JDET_ERR_BAD_ARG;
  if (workspace_bytes < jdet_nms_poly_workspace_bytes(n)) return JDET_ERR_WORKSPACE;
  const int cb = (n + 63) / 64;
  if ((size_t)cb * 8 > 200 * 1024) return JDET_ERR_UNSUPPORTED;          // the scan keeps one bit per box in shared memory (n <= 1.6 M)
  cudaStream_t st = (cudaStream_t)stream;
  Rec* rec = (Rec*)workspace;
  unsigned long long* mask = (unsigned long long*)((char*)workspace + jdet_align_up((size_t)n * sizeof(Rec), 256));
  poly_rec_kernel<<<jdet_ceil_div(n, 256), 256, 0, st>>>(dets, order, n, rec);
  poly_mask_kernel<<<dim3(cb, cb), 64, 0, st>>>(rec, n, iou_threshold, fast, mask);
  const size_t smem = (size_t)cb * 8;
  if (smem > 48 * 1024) JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(poly_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  poly_scan_kernel<<<1, 1024, smem, st>>>(mask, order, n, keep);
  return (int)cudaGetLastError();
}
