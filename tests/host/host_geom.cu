// host_geom.cu — TEST INFRASTRUCTURE: jdet_b200/csrc/rbox_geom.cuh (the PRODUCT's pair geometry) compiled for the
// host (JDET_HOST_CHECK turns its __device__ functions into __host__ __device__), so the CPU test suite can run the
// product's own reject stages and exact-IoU routine against the oracle bit for bit without a GPU.  Built on demand by
// tests/test_host_geom.py with -ffp-contract=off (the header's device build uses explicit _rn intrinsics instead).
// Nothing under jdet_b200/ links or loads this.
#define JDET_HOST_CHECK
#include "../../jdet_b200/csrc/rbox_geom.cuh"

using namespace jdet;

template <int VERSION, int VARIANT>
static float pair(const BoxRec& A, const BoxRec& B, int mode, int* counts) {
  // the kernels' stage order (box_iou_rotated.cu / nms_rotated.cu): circle -> SAT -> forced-zero tag -> exact
  if (mode >= 2) {
    if (circle_disjoint(A.x, A.y, A.qr, B.x, B.y, B.qr)) { counts[0]++; return 0.f; }
    if (sat_disjoint<VERSION>(A, B)) { counts[1]++; return 0.f; }
  }
  if (!(A.tag == 0.f && B.tag == 0.f)) return 0.f;
  counts[2]++;
  if (mode == 3) {   // strided point store with few slots (the exact kernels' shared-memory form): overflow must fall back
    float buf[3 * 4 * 6];
    return iou_exact_core<VERSION, 1, 4, 6>(A, B, buf, buf + 4 * 6, buf + 2 * 4 * 6);
  }
  if (mode == 4) {   // the kernels' own capacity
    float buf[3 * 2 * kExactCap];
    return iou_exact_core<VERSION, 1, 2, kExactCap>(A, B, buf, buf + 2 * kExactCap, buf + 4 * kExactCap);
  }
  return mode == 0 ? iou_exact_general<VERSION, VARIANT>(A, B) : iou_exact<VERSION, VARIANT>(A, B);
}

// mode 0: iou_exact_general for every pair; 1: iou_exact (fast path + fallback) for every pair; 2: full stage pipeline;
// 3 / 4: the strided, capacity-limited point store of the exact kernels (CUDA-build arithmetic only)
extern "C" void host_geom_iou(const float* b1, int n1, const float* b2, int n2, float* out, int version, int variant,
                              int mode, int* counts) {
  BoxRec* r1 = new BoxRec[n1 > 0 ? n1 : 1];
  BoxRec* r2 = new BoxRec[n2 > 0 ? n2 : 1];
  for (int i = 0; i < n1; i++) r1[i] = make_rec(b1[5 * i], b1[5 * i + 1], b1[5 * i + 2], b1[5 * i + 3], b1[5 * i + 4], 0.f, version == 1, false);
  for (int i = 0; i < n2; i++) r2[i] = make_rec(b2[5 * i], b2[5 * i + 1], b2[5 * i + 2], b2[5 * i + 3], b2[5 * i + 4], 0.f, version == 1, false);
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < n2; j++) {
      float v;
      if (version == 0) v = variant ? pair<0, 1>(r1[i], r2[j], mode, counts) : pair<0, 0>(r1[i], r2[j], mode, counts);
      else v = variant ? pair<1, 1>(r1[i], r2[j], mode, counts) : pair<1, 0>(r1[i], r2[j], mode, counts);
      out[(size_t)i * n2 + j] = v;
    }
  delete[] r1;
  delete[] r2;
}

// NMS-side pruning bound: returns iou_upper_bound<0>(A, B) (must never be below the exact IoU)
extern "C" void host_geom_upper_bound(const float* b1, int n1, const float* b2, int n2, float* out) {
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < n2; j++) {
      const BoxRec A = make_rec(b1[5 * i], b1[5 * i + 1], b1[5 * i + 2], b1[5 * i + 3], b1[5 * i + 4], 0.f, false, false);
      const BoxRec B = make_rec(b2[5 * j], b2[5 * j + 1], b2[5 * j + 2], b2[5 * j + 3], b2[5 * j + 4], 0.f, false, false);
      out[(size_t)i * n2 + j] = iou_upper_bound<0>(A, B);
    }
}
