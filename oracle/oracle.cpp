// oracle.cpp — CPU restatement of JDet's oriented-box geometry hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product (jdet_b200/) never does.
//
// Parity status: PINNED.  tests/test_oracle_pin.py checks every function below against
//   (a) oracle/_ref/libref_cpu.so  — the reference's own cpu_src compiled by g++ (IoU v0/v1, NMS),
//   (b) oracle/_ref/libref_cuda.so — the reference's __host__ __device__ CUDA-variant IoU run on
//       the host, and (on the GPU box) the five reference CUDA kernels themselves,
//   (c) tests/golden/*.npz          — vectors generated from (a)/(b) by tests/golden/make_golden*.py.
//
// Arithmetic contract: IEEE-754 binary32 with the reference's binary64 sub-steps, evaluated in
// source order with NO fused multiply-add (build: -ffp-contract=off).  The same contract is
// implemented on the device with __fmul_rn/__fadd_rn, which makes IoU and NMS bit-comparable.
//
// Every function cites the reference lines it restates (paths relative to
// /root/reference/python/jdet/).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct P2 { float x, y; };
inline P2 sub(P2 a, P2 b) { return {a.x - b.x, a.y - b.y}; }
inline float cross2(P2 a, P2 b) { return a.x * b.y - b.x * a.y; }   // ops/box_iou_rotated.py:47-50
inline float dot2(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }     // :42-45

enum { VARIANT_CPU = 0, VARIANT_CUDA = 1 };

struct RBox { float cx, cy, w, h, a; };

// ops/box_iou_rotated.py:52-72 (version 0) and ops/box_iou_rotated_v1.py:53-77 (version 1).
// Angle is promoted to double for cos/sin, narrowed to float, then halved in float.
void corners(const RBox& b, int version, P2 (&p)[4]) {
  const double th = b.a;
  const float c2 = (float)std::cos(th) * 0.5f;
  const float s2 = (float)std::sin(th) * 0.5f;
  if (version == 0) {
    p[0].x = b.cx - s2 * b.h - c2 * b.w;
    p[1].x = b.cx + s2 * b.h - c2 * b.w;
  } else {
    p[0].x = b.cx + s2 * b.h + c2 * b.w;
    p[1].x = b.cx - s2 * b.h + c2 * b.w;
  }
  p[0].y = b.cy + c2 * b.h - s2 * b.w;
  p[1].y = b.cy - c2 * b.h - s2 * b.w;
  p[2].x = 2 * b.cx - p[0].x;
  p[2].y = 2 * b.cy - p[0].y;
  p[3].x = 2 * b.cx - p[1].x;
  p[3].y = 2 * b.cy - p[1].y;
}

// ops/box_iou_rotated.py:74-153 — edge x edge crossings, then corners of 1 in 2, then 2 in 1.
int clip_points(const P2 (&a)[4], const P2 (&b)[4], P2 (&out)[24]) {
  P2 ea[4], eb[4];
  for (int i = 0; i < 4; i++) {
    ea[i] = sub(a[(i + 1) & 3], a[i]);
    eb[i] = sub(b[(i + 1) & 3], b[i]);
  }
  int n = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      const float det = cross2(eb[j], ea[i]);
      if (std::fabs((double)det) <= 1e-14) continue;          // :97
      const P2 d = sub(b[j], a[i]);
      const float t1 = cross2(eb[j], d) / det;
      const float t2 = cross2(ea[i], d) / det;
      if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) {
        out[n].x = a[i].x + ea[i].x * t1;
        out[n].y = a[i].y + ea[i].y * t1;
        n++;
      }
    }
  {  // :111-131
    const P2 AB = eb[0], DA = eb[3];
    const float ABAB = dot2(AB, AB), ADAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      const P2 AP = sub(a[i], b[0]);
      const float pAB = dot2(AP, AB), pAD = -dot2(AP, DA);
      if (pAB >= 0 && pAD >= 0 && pAB <= ABAB && pAD <= ADAD) out[n++] = a[i];
    }
  }
  {  // :133-150
    const P2 AB = ea[0], DA = ea[3];
    const float ABAB = dot2(AB, AB), ADAD = dot2(DA, DA);
    for (int i = 0; i < 4; i++) {
      const P2 AP = sub(b[i], a[0]);
      const float pAB = dot2(AP, AB), pAD = -dot2(AP, DA);
      if (pAB >= 0 && pAD >= 0 && pAB <= ABAB && pAD <= ADAD) out[n++] = b[i];
    }
  }
  return n;
}

// ops/box_iou_rotated.py:155-238.  The angular sort differs between the reference's two builds:
//   VARIANT_CPU  — std::sort with the comparator at :316-325,
//   VARIANT_CUDA — the in-place exchange sort at :335-351 (the one the GPU path runs).
// Returns the hull size; hull points are left relative to the pivot (shift_to_zero = true, :275).
int hull(const P2 (&p)[24], int n, P2 (&q)[24], int variant) {
  int t = 0;
  for (int i = 1; i < n; i++)
    if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
  const P2 start = p[t];
  for (int i = 0; i < n; i++) q[i] = sub(p[i], start);
  std::swap(q[0], q[t]);
  float dist[24];
  for (int i = 0; i < n; i++) dist[i] = dot2(q[i], q[i]);

  if (variant == VARIANT_CPU) {
    std::sort(q + 1, q + n, [](const P2& A, const P2& B) -> bool {
      const float c = cross2(A, B);
      if (std::fabs((double)c) < 1e-6) return dot2(A, A) < dot2(B, B);
      return c > 0;
    });
    // NB (:219-224): the reference does NOT recompute dist[] after std::sort, so Step 4 below
    // reads the pre-sort distances.  Restated faithfully.
  } else {
    for (int i = 1; i < n - 1; i++)
      for (int j = i + 1; j < n; j++) {
        const float c = cross2(q[i], q[j]);
        if (c < -1e-6 || (std::fabs((double)c) < 1e-6 && dist[i] > dist[j])) {
          std::swap(q[i], q[j]);
          std::swap(dist[i], dist[j]);
        }
      }
  }
  int k = 1;
  for (; k < n; k++)
    if (dist[k] > 1e-8) break;
  if (k == n) { q[0] = p[t]; return 1; }
  q[1] = q[k];
  int m = 2;
  for (int i = k + 1; i < n; i++) {
    while (m > 1 && cross2(sub(q[i], q[m - 2]), sub(q[m - 1], q[m - 2])) >= 0) m--;
    q[m++] = q[i];
  }
  return m;
}

// ops/box_iou_rotated.py:240-252
float fan_area(const P2 (&q)[24], int m) {
  if (m <= 2) return 0;
  float area = 0;
  for (int i = 1; i < m - 1; i++)
    area += (float)std::fabs((double)cross2(sub(q[i], q[0]), sub(q[i + 1], q[0])));
  return (float)(area / 2.0);
}

// ops/box_iou_rotated.py:281-310 / ops/nms_rotated.py:281-313 (label column when box_length == 6)
float iou_one(const float* r1, const float* r2, int version, int variant, int box_length) {
  if (box_length == 6 && r1[5] != r2[5]) return 0.0f;
  const double sx = (r1[0] + r2[0]) / 2.0;
  const double sy = (r1[1] + r2[1]) / 2.0;
  RBox b1{(float)(r1[0] - sx), (float)(r1[1] - sy), r1[2], r1[3], r1[4]};
  RBox b2{(float)(r2[0] - sx), (float)(r2[1] - sy), r2[2], r2[3], r2[4]};
  const float area1 = b1.w * b1.h, area2 = b2.w * b2.h;
  if (area1 < 1e-14 || area2 < 1e-14) return 0.f;
  P2 p1[4], p2[4], pts[24], ord[24];
  corners(b1, version, p1);
  corners(b2, version, p2);
  const int n = clip_points(p1, p2, pts);
  float inter = 0.0f;
  if (n > 2) {
    const int m = hull(pts, n, ord, variant);
    inter = fan_area(ord, m);
  }
  return inter / (area1 + area2 - inter);
}

// ---- sampling helpers -------------------------------------------------------------------

// ops/roi_align_rotated_v1.py:23-68 (version 1: clamp on "< 0") and ops/roi_align_rotated.py:21-56,
// ops/fr.py:18-67 (version 0: clamp on "<= 0" — numerically the same thing).
inline float bilinear_clamped(const float* plane, int H, int W, float y, float x) {
  if (y < -1.0 || y > H || x < -1.0 || x > W) return 0;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = y - yl, lx = x - xl;
  const float hy = (float)(1. - ly), hx = (float)(1. - lx);
  const float lt = plane[yl * W + xl], rt = plane[yl * W + xh];
  const float lb = plane[yh * W + xl], rb = plane[yh * W + xh];
  const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
  return w1 * lt + w2 * rt + w3 * lb + w4 * rb;
}

// ops/dcn_v1.py:25-56: floor-based taps, each corner zero-padded individually.
inline float bilinear_zeropad(const float* plane, int H, int W, float h, float w) {
  const int hl = (int)std::floor(h), wl = (int)std::floor(w);
  const int hh = hl + 1, wh = wl + 1;
  const float lh = h - hl, lw = w - wl;
  const float uh = 1 - lh, uw = 1 - lw;
  float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
  if (hl >= 0 && wl >= 0) v1 = plane[hl * W + wl];
  if (hl >= 0 && wh <= W - 1) v2 = plane[hl * W + wh];
  if (hh <= H - 1 && wl >= 0) v3 = plane[hh * W + wl];
  if (hh <= H - 1 && wh <= W - 1) v4 = plane[hh * W + wh];
  const float w1 = uh * uw, w2 = uh * lw, w3 = lh * uw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

// taps + weights of one sample, shared by forward/backward (bilinear_interpolate_gradient,
// ops/roi_align_rotated_v1.py:148-190, ops/fr.py:69-112): returns false when the sample is out of range.
inline bool tap_weights(int H, int W, float y, float x, int (&off)[4], float (&w)[4]) {
  if (y < -1.0 || y > H || x < -1.0 || x > W) return false;
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = y - yl, lx = x - xl;
  const float hy = (float)(1. - ly), hx = (float)(1. - lx);
  off[0] = yl * W + xl; off[1] = yl * W + xh; off[2] = yh * W + xl; off[3] = yh * W + xh;
  w[0] = hy * hx; w[1] = hy * lx; w[2] = ly * hx; w[3] = ly * lx;
  return true;
}

}  // namespace

extern "C" {

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

float orc_single_iou(const float* a, const float* b, int version, int variant, int box_length) {
  return iou_one(a, b, version, variant, box_length);
}

// ops/box_iou_rotated.py:487-500 (rows of 5 floats).  threads <= 0 -> all cores.
void orc_box_iou_rotated(const float* b1, int n1, const float* b2, int n2, float* out, int version,
                         int variant, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
#endif
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < n2; j++)
      out[(size_t)i * n2 + j] = iou_one(b1 + (size_t)i * 5, b2 + (size_t)j * 5, version, variant, 5);
}

// ops/box_iou_rotated_v1.py:516-523 post-pass: rows/cols of boxes with min(w,h) < 1e-3 are zeroed.
void orc_box_iou_rotated_v1_postzero(const float* b1, int n1, const float* b2, int n2, float* out) {
  for (int i = 0; i < n1; i++)
    if (std::min(b1[i * 5 + 2], b1[i * 5 + 3]) < 0.001f)
      for (int j = 0; j < n2; j++) out[(size_t)i * n2 + j] = 0.f;
  for (int j = 0; j < n2; j++)
    if (std::min(b2[j * 5 + 2], b2[j * 5 + 3]) < 0.001f)
      for (int i = 0; i < n1; i++) out[(size_t)i * n2 + j] = 0.f;
}

// Greedy NMS over `order` (descending score).
//   VARIANT_CPU : ops/nms_rotated.py:414-449 — suppress when iou >= thr, pairs skipped once suppressed.
//   VARIANT_CUDA: ops/nms_rotated.py:352-411 + 475-491 — bit (i,j) set when iou > thr for every
//                 i < j in sorted order; a kept box ORs its row.  Same greedy result, strict compare.
// The IoU argument order is (higher-ranked box, lower-ranked box) in both.
void orc_nms_rotated(const float* dets, int n, int box_length, const int* order, float thr,
                     int variant, unsigned char* keep) {
  std::vector<unsigned char> dead(n, 0);
  std::memset(keep, 0, (size_t)n);
  for (int a = 0; a < n; a++) {
    const int i = order[a];
    if (dead[i]) continue;
    keep[i] = 1;
    for (int b = a + 1; b < n; b++) {
      const int j = order[b];
      if (dead[j]) continue;
      const float v = iou_one(dets + (size_t)i * box_length, dets + (size_t)j * box_length, 0, variant,
                              box_length);
      if (variant == VARIANT_CPU ? (v >= thr) : (v > thr)) dead[j] = 1;
    }
  }
}

// ops/roi_align_rotated_v1.py:70-147 (version 1) / ops/roi_align_rotated.py:60-127 (version 0).
// input (B,C,H,W), rois (R,6)=[batch,cx,cy,w,h,theta] -> out (R,C,PH,PW).
// sample_num: the reference bakes `const float sampling_ratio` and passes it to an int parameter.
void orc_roi_align_rotated(int version, const float* input, const float* rois, int R, int C, int H, int W,
                           int PH, int PW, float spatial_scale, int sample_num, float* out,
                           int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
  for (int n = 0; n < R; n++) {
    const float* roi = rois + (size_t)n * 6;
    const int batch = (int)roi[0];
    float cw, ch;
    if (version == 1) {
      cw = roi[1] * spatial_scale - 0.5f;
      ch = roi[2] * spatial_scale - 0.5f;
    } else {
      cw = roi[1] * spatial_scale;
      ch = roi[2] * spatial_scale;
    }
    float rw = roi[3] * spatial_scale, rh = roi[4] * spatial_scale;
    const float theta = roi[5];
    rw = std::max(rw, 1.f);
    rh = std::max(rh, 1.f);
    const float bin_h = rh / (float)PH, bin_w = rw / (float)PW;
    const int gh = sample_num > 0 ? sample_num : (int)std::ceil(rh / PH);
    const int gw = sample_num > 0 ? sample_num : (int)std::ceil(rw / PW);
    const float start_h = (float)(-rh / 2.0), start_w = (float)(-rw / 2.0);
    const float ct = std::cos(theta), st = std::sin(theta);   // float overloads == cosf/sinf
    const float count = version == 1 ? (float)std::max(gh * gw, 1) : (float)(gh * gw);
    for (int c = 0; c < C; c++) {
      const float* plane = input + ((size_t)batch * C + c) * H * W;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; iy++) {
            const float yy = start_h + ph * bin_h + (float)(iy + .5f) * bin_h / (float)gh;
            for (int ix = 0; ix < gw; ix++) {
              const float xx = start_w + pw * bin_w + (float)(ix + .5f) * bin_w / (float)gw;
              float x, y;
              if (version == 1) {
                x = xx * ct + yy * st + cw;
                y = yy * ct - xx * st + ch;
              } else {
                x = xx * ct - yy * st + cw;
                y = xx * st + yy * ct + ch;
              }
              acc += bilinear_clamped(plane, H, W, y, x);
            }
          }
          acc /= count;
          out[(((size_t)n * C + c) * PH + ph) * PW + pw] = acc;
        }
    }
  }
}

// ops/fr.py:114-165.  features (N,C,H,W), boxes (N,H,W,5) -> out (N,C,H,W); points in {1,5}.
// NB bbox[0] scales to the ROW coordinate and bbox[1] to the COLUMN (:133-134).
void orc_feature_refine(const float* feat, const float* boxes, int N, int C, int H, int W, int points,
                        float spatial_scale, float* out, int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for collapse(2) schedule(static) num_threads(threads)
#endif
  for (int n = 0; n < N; n++)
    for (int h = 0; h < H; h++)
      for (int w = 0; w < W; w++) {
        const float* bb = boxes + (((size_t)n * H + h) * W + w) * 5;
        const float roi_y = bb[0] * spatial_scale, roi_x = bb[1] * spatial_scale;
        float px[5] = {roi_x, 0, 0, 0, 0}, py[5] = {roi_y, 0, 0, 0, 0};
        if (points > 1) {
          const float rw = bb[2] * spatial_scale, rh = bb[3] * spatial_scale, ra = bb[4];
          const float w2 = rw / 2, h2 = rh / 2;
          const float ca = cosf(ra), sa = sinf(ra);
          const float wx = ca * w2, wy = sa * w2, hx = -sa * h2, hy = ca * h2;
          px[1] = roi_x + wx + hx; py[1] = roi_y + wy + hy;
          px[2] = roi_x - wx + hx; py[2] = roi_y - wy + hy;
          px[3] = roi_x - wx - hx; py[3] = roi_y - wy - hy;
          px[4] = roi_x + wx - hx; py[4] = roi_y + wy - hy;
        }
        for (int c = 0; c < C; c++) {
          const float* plane = feat + ((size_t)n * C + c) * H * W;
          float v = plane[h * W + w];
          for (int i = 0; i < points; i++) v += bilinear_clamped(plane, H, W, py[i], px[i]);
          out[(((size_t)n * C + c) * H + h) * W + w] = v;
        }
      }
}

// ROIAlignBackward (ops/roi_align_rotated_v1.py:192-298, ops/roi_align_rotated.py:164-255):
// grad_input[b,c,tap] += grad_out[n,c,ph,pw] * w / (gh*gw).  The reference scatters with float
// atomicAdd (order-dependent rounding); this restatement accumulates in double.
void orc_roi_align_rotated_backward(int version, const float* grad_out, const float* rois, int R, int B, int C, int H,
                                    int W, int PH, int PW, float spatial_scale, int sample_num, float* grad_in) {
  std::vector<double> acc((size_t)B * C * H * W, 0.0);
  for (int n = 0; n < R; n++) {
    const float* roi = rois + (size_t)n * 6;
    const int batch = (int)roi[0];
    float cw, ch;
    if (version == 1) { cw = roi[1] * spatial_scale - 0.5f; ch = roi[2] * spatial_scale - 0.5f; }
    else { cw = roi[1] * spatial_scale; ch = roi[2] * spatial_scale; }
    float rw = roi[3] * spatial_scale, rh = roi[4] * spatial_scale;
    const float theta = roi[5];
    rw = std::max(rw, 1.f);
    rh = std::max(rh, 1.f);
    const float bin_h = rh / (float)PH, bin_w = rw / (float)PW;
    const int gh = sample_num > 0 ? sample_num : (int)std::ceil(rh / PH);
    const int gw = sample_num > 0 ? sample_num : (int)std::ceil(rw / PW);
    const float start_h = (float)(-rh / 2.0), start_w = (float)(-rw / 2.0);
    const float ct = std::cos(theta), st = std::sin(theta);
    const float count = (float)(gh * gw);
    for (int ph = 0; ph < PH; ph++)
      for (int pw = 0; pw < PW; pw++)
        for (int iy = 0; iy < gh; iy++) {
          const float yy = start_h + ph * bin_h + (float)(iy + .5f) * bin_h / (float)gh;
          for (int ix = 0; ix < gw; ix++) {
            const float xx = start_w + pw * bin_w + (float)(ix + .5f) * bin_w / (float)gw;
            float x, y;
            if (version == 1) { x = xx * ct + yy * st + cw; y = yy * ct - xx * st + ch; }
            else { x = xx * ct - yy * st + cw; y = xx * st + yy * ct + ch; }
            int off[4]; float w[4];
            if (!tap_weights(H, W, y, x, off, w)) continue;
            for (int c = 0; c < C; c++) {
              const float g = grad_out[(((size_t)n * C + c) * PH + ph) * PW + pw];
              double* plane = acc.data() + ((size_t)batch * C + c) * H * W;
              for (int k = 0; k < 4; k++) plane[off[k]] += (double)(g * w[k] / count);
            }
          }
        }
  }
  for (size_t i = 0; i < acc.size(); i++) grad_in[i] = (float)acc[i];
}

// feature_refine_backward_kernel (ops/fr.py:167-232): grad_in = grad_out + scatter of grad_out * w.
void orc_feature_refine_backward(const float* grad_out, const float* boxes, int N, int C, int H, int W, int points,
                                 float spatial_scale, float* grad_in) {
  std::vector<double> acc((size_t)N * C * H * W);
  for (size_t i = 0; i < acc.size(); i++) acc[i] = grad_out[i];
  for (int n = 0; n < N; n++)
    for (int h = 0; h < H; h++)
      for (int w = 0; w < W; w++) {
        const float* bb = boxes + (((size_t)n * H + h) * W + w) * 5;
        const float roi_y = bb[0] * spatial_scale, roi_x = bb[1] * spatial_scale;
        float px[5] = {roi_x, 0, 0, 0, 0}, py[5] = {roi_y, 0, 0, 0, 0};
        if (points > 1) {
          const float rw = bb[2] * spatial_scale, rh = bb[3] * spatial_scale, ra = bb[4];
          const float w2 = rw / 2, h2 = rh / 2;
          const float ca = cosf(ra), sa = sinf(ra);
          const float wx = ca * w2, wy = sa * w2, hx = -sa * h2, hy = ca * h2;
          px[1] = roi_x + wx + hx; py[1] = roi_y + wy + hy;
          px[2] = roi_x - wx + hx; py[2] = roi_y - wy + hy;
          px[3] = roi_x - wx - hx; py[3] = roi_y - wy - hy;
          px[4] = roi_x + wx - hx; py[4] = roi_y + wy - hy;
        }
        for (int i = 0; i < points; i++) {
          int off[4]; float wt[4];
          if (!tap_weights(H, W, py[i], px[i], off, wt)) continue;
          for (int c = 0; c < C; c++) {
            const float g = grad_out[(((size_t)n * C + c) * H + h) * W + w];
            double* plane = acc.data() + ((size_t)n * C + c) * H * W;
            for (int k = 0; k < 4; k++) plane[off[k]] += (double)(g * wt[k]);
          }
        }
      }
  for (size_t i = 0; i < acc.size(); i++) grad_in[i] = (float)acc[i];
}

// models/roi_heads/s2anet_head.py:677-713 — AlignConv.get_offset for ONE image, kernel_size k (odd).
// anchors (H*W,5) image-space -> offset (2*k*k, H, W), channel 2*t = dy, 2*t+1 = dx, t = i*k + j.
// All steps are fp32 elementwise ops in the reference; same order here.
void orc_align_conv_offset(const float* anchors, int H, int W, float stride, int k, float* offset) {
  const int pad = (k - 1) / 2;
  for (int p = 0; p < H * W; p++) {
    const float* a = anchors + (size_t)p * 5;
    const float xc = (float)(p % W), yc = (float)(p / W);
    const float x_ctr = a[0] / stride, y_ctr = a[1] / stride, w = a[2] / stride, h = a[3] / stride;
    const float c = cosf(a[4]), s = sinf(a[4]);
    const float dw = w / (float)k, dh = h / (float)k;
    for (int i = 0; i < k; i++)
      for (int j = 0; j < k; j++) {
        const float xx = (float)(j - pad), yy = (float)(i - pad);
        const float x = dw * xx, y = dh * yy;
        const float xr = c * x - s * y, yr = s * x + c * y;
        const float xa = xr + x_ctr, ya = yr + y_ctr;
        const float ox = xa - (xc + xx), oy = ya - (yc + yy);
        const int t = i * k + j;
        offset[((size_t)(2 * t) * H * W) + p] = oy;
        offset[((size_t)(2 * t + 1) * H * W) + p] = ox;
      }
  }
}

// ops/dcn_v1.py:131-184 + 412-454 — DeformConv v1 forward (deformable im2col then W x columns),
// stride/pad/dilation general, groups = 1; accumulates in double (the reference GEMM is cuBLAS
// SGEMM whose summation order is unspecified; fp64 is the order-free yardstick, tolerance 1e-4).
// x (B,C,H,W), offset (B, dg*2*kh*kw, Ho, Wo), weight (Co, C, kh, kw) -> out (B,Co,Ho,Wo)
void orc_deform_conv(const float* x, const float* offset, const float* weight, int B, int C, int H, int W,
                     int Co, int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w,
                     int dil_h, int dil_w, int dg, int relu, float* out, int threads) {
  const int Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  const int cpg = C / dg;
  const int K = C * kh * kw;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
  {
    std::vector<float> col(K);
#ifdef _OPENMP
#pragma omp for collapse(2) schedule(static)
#endif
    for (int b = 0; b < B; b++)
      for (int p = 0; p < Ho * Wo; p++) {
        const int ho = p / Wo, wo = p % Wo;
        const int h_in = ho * stride_h - pad_h, w_in = wo * stride_w - pad_w;
        for (int c = 0; c < C; c++) {
          const float* plane = x + ((size_t)b * C + c) * H * W;
          const float* off = offset + ((size_t)b * dg + c / cpg) * 2 * kh * kw * Ho * Wo;
          for (int i = 0; i < kh; i++)
            for (int j = 0; j < kw; j++) {
              const int t = i * kw + j;
              const float oh = off[((size_t)(2 * t) * Ho + ho) * Wo + wo];
              const float ow = off[((size_t)(2 * t + 1) * Ho + ho) * Wo + wo];
              const float h_im = h_in + i * dil_h + oh;
              const float w_im = w_in + j * dil_w + ow;
              float v = 0.f;
              if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)
                v = bilinear_zeropad(plane, H, W, h_im, w_im);
              col[(size_t)c * kh * kw + t] = v;
            }
        }
        for (int co = 0; co < Co; co++) {
          const float* wr = weight + (size_t)co * K;
          double acc = 0.0;
          for (int q = 0; q < K; q++) acc += (double)wr[q] * (double)col[q];
          float r = (float)acc;
          if (relu && r < 0.f) r = 0.f;
          out[(((size_t)b * Co + co) * Ho + ho) * Wo + wo] = r;
        }
      }
  }
}

// DeformConv v1 backward (ops/dcn_v1.py:185-306 kernels, 457-556 host flow), groups = 1:
//   grad_x      : deformable_col2im        — bilinear weights of get_gradient_weight (:58-82) on in-image corners
//   grad_offset : deformable_col2im_coord  — get_coordinate_weight (:84-129), 0 outside (-1,H)x(-1,W)
//   grad_weight : grad_out x columns^T     — columns from deformable_im2col
// Everything accumulates in double (the reference: float atomics + SGEMM).
void orc_deform_conv_backward(const float* x, const float* offset, const float* weight, const float* grad_out, int B, int C,
                              int H, int W, int Co, int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w,
                              int dil_h, int dil_w, int dg, float* gx, float* goff, float* gw) {
  const int Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  const int cpg = C / dg, khw = kh * kw, K = C * khw, P = Ho * Wo;
  std::vector<double> ax((size_t)B * C * H * W, 0.0), ao((size_t)B * dg * 2 * khw * P, 0.0), aw((size_t)Co * K, 0.0);
  std::vector<double> colg(K);
  for (int b = 0; b < B; b++)
    for (int p = 0; p < P; p++) {
      const int ho = p / Wo, wo = p % Wo;
      // column gradient of this output position: W^T x grad_out
      for (int k = 0; k < K; k++) colg[k] = 0.0;
      for (int co = 0; co < Co; co++) {
        const double g = grad_out[(((size_t)b * Co + co) * Ho + ho) * Wo + wo];
        if (g == 0.0) continue;
        const float* wr = weight + (size_t)co * K;
        for (int k = 0; k < K; k++) colg[k] += (double)wr[k] * g;
      }
      for (int c = 0; c < C; c++) {
        const float* plane = x + ((size_t)b * C + c) * H * W;
        const int g = c / cpg;
        const float* off = offset + ((size_t)b * dg + g) * 2 * khw * P;
        for (int t = 0; t < khw; t++) {
          const int i = t / kw, j = t % kw;
          const float oh = off[(size_t)(2 * t) * P + p], ow = off[(size_t)(2 * t + 1) * P + p];
          const float h_im = (ho * stride_h - pad_h) + i * dil_h + oh;
          const float w_im = (wo * stride_w - pad_w) + j * dil_w + ow;
          const double cg = colg[(size_t)c * khw + t];
          // forward sample for grad_weight
          float v = 0.f;
          const bool inside = h_im > -1 && w_im > -1 && h_im < H && w_im < W;
          if (inside) v = bilinear_zeropad(plane, H, W, h_im, w_im);
          for (int co = 0; co < Co; co++)
            aw[(size_t)co * K + (size_t)c * khw + t] += (double)grad_out[(((size_t)b * Co + co) * Ho + ho) * Wo + wo] * v;
          if (!inside) continue;
          const int hl = (int)std::floor(h_im), wl = (int)std::floor(w_im), hh = hl + 1, wh = wl + 1;
          const double lh = (double)h_im - hl, lw = (double)w_im - wl;
          const double uh = 1.0 - lh, uw = 1.0 - lw;
          double* gplane = ax.data() + ((size_t)b * C + c) * H * W;
          const bool v1 = hl >= 0 && wl >= 0, v2 = hl >= 0 && wh <= W - 1, v3 = hh <= H - 1 && wl >= 0, v4 = hh <= H - 1 && wh <= W - 1;
          if (v1) gplane[hl * W + wl] += uh * uw * cg;
          if (v2) gplane[hl * W + wh] += uh * lw * cg;
          if (v3) gplane[hh * W + wl] += lh * uw * cg;
          if (v4) gplane[hh * W + wh] += lh * lw * cg;
          const double p1 = v1 ? plane[hl * W + wl] : 0.0, p2 = v2 ? plane[hl * W + wh] : 0.0;
          const double p3 = v3 ? plane[hh * W + wl] : 0.0, p4 = v4 ? plane[hh * W + wh] : 0.0;
          const double dh_w = -uw * p1 - lw * p2 + uw * p3 + lw * p4;    // d sample / d h   (bp_dir 0)
          const double dw_w = -uh * p1 + uh * p2 - lh * p3 + lh * p4;    // d sample / d w   (bp_dir 1)
          double* go = ao.data() + ((size_t)b * dg + g) * 2 * khw * P;
          go[(size_t)(2 * t) * P + p] += dh_w * cg;
          go[(size_t)(2 * t + 1) * P + p] += dw_w * cg;
        }
      }
    }
  for (size_t i = 0; i < ax.size(); i++) gx[i] = (float)ax[i];
  for (size_t i = 0; i < ao.size(); i++) goff[i] = (float)ao[i];
  for (size_t i = 0; i < aw.size(); i++) gw[i] = (float)aw[i];
}

// ops/dcn_v1.py:131-184 alone: columns (C*kh*kw, B, Ho, Wo) as the reference lays them out.
void orc_deform_im2col(const float* x, const float* offset, int B, int C, int H, int W, int kh, int kw,
                       int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int dg,
                       float* columns) {
  const int Ho = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int Wo = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  const int cpg = C / dg;
  for (int c = 0; c < C; c++)
    for (int b = 0; b < B; b++) {
      const float* plane = x + ((size_t)b * C + c) * H * W;
      const float* off = offset + ((size_t)b * dg + c / cpg) * 2 * kh * kw * Ho * Wo;
      for (int ho = 0; ho < Ho; ho++)
        for (int wo = 0; wo < Wo; wo++)
          for (int i = 0; i < kh; i++)
            for (int j = 0; j < kw; j++) {
              const int t = i * kw + j;
              const float oh = off[((size_t)(2 * t) * Ho + ho) * Wo + wo];
              const float ow = off[((size_t)(2 * t + 1) * Ho + ho) * Wo + wo];
              const float h_im = (ho * stride_h - pad_h) + i * dil_h + oh;
              const float w_im = (wo * stride_w - pad_w) + j * dil_w + ow;
              float v = 0.f;
              if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)
                v = bilinear_zeropad(plane, H, W, h_im, w_im);
              columns[((((size_t)c * kh * kw + t) * B + b) * Ho + ho) * Wo + wo] = v;
            }
    }
}

}  // extern "C"
