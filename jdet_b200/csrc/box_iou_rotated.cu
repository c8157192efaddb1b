// box_iou_rotated.cu — N x M rotated IoU matrix for sm_100a.
//
// Replaces box_iou_rotated_cuda_kernel + its launch snippet
// (/root/reference/python/jdet/ops/box_iou_rotated.py:412-485, _v1.py:417-490 + the Python
// post-pass _v1.py:516-523).  The reference runs the full clip+hull for every pair from a
// 528-byte local-memory stack; at detection densities >95 % of pairs do not overlap, so the
// matrix is an HBM *write* stream (4*N*M bytes) with a sparse set of expensive entries.
//
// Layout in HBM: boxes (n,5) fp32 row-major in; per-box BoxRec[] (32 B) scratch in the caller's
// workspace; ious (N,M) fp32 row-major out.
//
// Kernel 1  rec_kernel        one thread per box: double-precision cos/sin hoisted out of the
//                             N*M loop (the reference recomputes them per pair), circumradius.
// Kernel 2  iou_tile_kernel   CTA = 64 x 128 output tile, 256 threads, 8 CTAs per SM.
//     phase 1  every pair: circle test whose verdict is a sign bit funnel-shifted into a per-thread
//              32-bit mask (7 issue slots per pair), 16-B streaming stores of +0.0; survivors are
//              compacted into a 4096-entry shared-memory queue (one warp scan + one atomic per warp),
//              in rounds when a tile has more.
//     phase 2  queue -> SAT test -> second queue (warp-aggregated push).
//              The surviving (row, col) candidates are appended to a device-wide queue.
// Kernel 3  iou_exact_kernel  one thread per queued candidate (grid-stride over the device-side count):
//                             reference-exact clip/hull IoU (straight-line edge stage, clip points in a
//                             [slot][thread] shared-memory array), every lane busy, no barriers, 4-B stores
//                             over the zeros.  (Queue full => the tile CTA evaluates its own.)
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "rbox_geom.cuh"

namespace jdet {

// Tile = TR x 128 outputs per 256-thread CTA.  TR = 64 for most sizes; 128 for the largest (A/B on one B200: 16k x 16k
// 425 -> 404 us — the per-CTA fixed work, record loads, five barriers and the compaction scans, is shared by twice the
// pairs — but 4k x 4k clustered 120 -> 130 us and 1k x 1k 21.5 -> 23.6 us: fewer, longer CTAs fill the machine worse).
constexpr int kTC = 128, kThreads = 256;
constexpr long long kBigTilePairs = 1ll << 27;
#ifndef JDET_IOU_TILE_MINB
#define JDET_IOU_TILE_MINB 8          // resident CTAs per SM asked of ptxas (32 registers; A/B on one B200, 16k x 16k: 4 -> 537 us, 6 -> 525, 8 -> 517)
#endif
#ifndef JDET_IOU_QCAP
#define JDET_IOU_QCAP 4096          // (clustered 4k x 4k on one B200: 1024 -> 134 us, 2048 -> 128, 4096 -> 118; 16k x 16k DOTA-shaped: no difference)
#endif
constexpr int kQCap = JDET_IOU_QCAP;        // survivor queue entries per round (a 64 x 128 tile holds 8192 pairs)

// tag handling: IoU has no labels; tag = 1.0f marks a forced-zero box (v1 small-box post pass).
// both box sets in one launch (small problems are launch-bound); also resets the candidate counter
__global__ void __launch_bounds__(256) rec_kernel(const float* __restrict__ boxes1, int n1, const float* __restrict__ boxes2,
                                                  int n2, int zero_small, BoxRec* __restrict__ rec1,
                                                  BoxRec* __restrict__ rec2, unsigned long long* __restrict__ gcount) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *gcount = 0ull;
  const float* b;
  BoxRec* dst;
  if (i < n1) { b = boxes1 + (size_t)i * 5; dst = rec1 + i; }
  else {
    i -= n1;
    if (i >= n2) return;
    b = boxes2 + (size_t)i * 5; dst = rec2 + i;
  }
  *dst = make_rec(b[0], b[1], b[2], b[3], b[4], 0.f, zero_small != 0, false);
}

template <int VERSION, bool VEC4, int kTR>
__global__ void __launch_bounds__(kThreads, JDET_IOU_TILE_MINB) iou_tile_kernel(const BoxRec* __restrict__ rec1, int n1,
                                                             const BoxRec* __restrict__ rec2, int n2,
                                                             float* __restrict__ out, unsigned long long* __restrict__ gcount,
                                                             uint2* __restrict__ gqueue, int gcap, int variant) {
  constexpr int kRPW = kTR / 8, kMasks = kRPW / 8;      // rows per warp; 32-pair masks per thread (8 rows x 4 columns each)
  static_assert(kTR == 64 || kTR == 128, "tile rows");
  __shared__ BoxRec s_row[kTR];
  __shared__ BoxRec s_col[kTC];
  __shared__ __align__(16) float s_cx[kTC], s_cy[kTC], s_cr[kTC];
  __shared__ float4 s_rq[kTR];                          // rows: (x, y, qr, -) for the broadcast loads of phase 1
  __shared__ unsigned short s_q1[kQCap];
  __shared__ unsigned short s_q2[kQCap];
  __shared__ int s_cnt1, s_cnt2;
  __shared__ unsigned long long s_base;               // 64-bit: the reservations of a dense 46k x 46k problem pass 2^31

  const int tid = threadIdx.x;
  const int row0 = blockIdx.y * kTR, col0 = blockIdx.x * kTC;

  if (tid == 0) { s_cnt1 = 0; s_cnt2 = 0; }
  if (tid < kTR + kTC) {
    BoxRec r;
    const bool is_row = tid < kTR;
    const int g = is_row ? row0 + tid : col0 + (tid - kTR);
    const bool ok = is_row ? (g < n1) : (g < n2);
    if (ok) {
      r = is_row ? rec1[g] : rec2[g];
    } else {
      r.x = r.y = r.w = r.h = r.c2 = r.s2 = 0.f; r.qr = -INFINITY; r.tag = 1.f;
    }
    if (is_row) { s_row[tid] = r; s_rq[tid] = make_float4(r.x, r.y, r.qr, 0.f); }
    else {
      const int c = tid - kTR;
      s_col[c] = r; s_cx[c] = r.x; s_cy[c] = r.y; s_cr[c] = r.qr;
    }
  }
  __syncthreads();

  // ---- phase 1: circle test over the whole tile, zero stores ---------------------------------
  // 8 issue slots per pair is what this kernel runs out of (ncu: 79 % SM busy at 4.5 TB/s of stores), so the test
  // is 6 float ops whose SIGN is the verdict, shifted into the per-thread mask by one funnel shift — no compare,
  // select or OR per pair — and a row costs one broadcast LDS.128, one pointer add and one 16-B store.
  const int warp = tid >> 5, lane = tid & 31;
  const float4 cx = reinterpret_cast<const float4*>(s_cx)[lane];
  const float4 cy = reinterpret_cast<const float4*>(s_cy)[lane];
  const float4 cr = reinterpret_cast<const float4*>(s_cr)[lane];
  unsigned rej[kMasks];                                 // mask m: pair p = 4 * (k - 8 * m) + q ends up in bit 31 - p
#pragma unroll
  for (int m = 0; m < kMasks; m++) rej[m] = 0u;
  {
    const int gc = col0 + 4 * lane;
    float* o = out + (size_t)(row0 + warp * kRPW) * n2 + gc;
    const int rows_left = n1 - (row0 + warp * kRPW);
#pragma unroll
    for (int k = 0; k < kRPW; k++) {
      const float4 rq = s_rq[warp * kRPW + k];
      {   // four circle tests as two packed passes (columns x|y, z|w against the broadcast row)
        const unsigned long long X1 = f2_pack(rq.x, rq.x), Y1 = f2_pack(rq.y, rq.y), Q1 = f2_pack(rq.z, rq.z);
        float t0, t1, t2, t3;
        circle_t2(X1, Y1, Q1, f2_pack(cx.x, cx.y), f2_pack(cy.x, cy.y), f2_pack(cr.x, cr.y), t0, t1);
        circle_t2(X1, Y1, Q1, f2_pack(cx.z, cx.w), f2_pack(cy.z, cy.w), f2_pack(cr.z, cr.w), t2, t3);
        unsigned& rj = rej[k / 8];
        rj = __funnelshift_l(__float_as_uint(t0), rj, 1);
        rj = __funnelshift_l(__float_as_uint(t1), rj, 1);
        rj = __funnelshift_l(__float_as_uint(t2), rj, 1);
        rj = __funnelshift_l(__float_as_uint(t3), rj, 1);
      }
      if (k < rows_left) {
        if (VEC4) {
          if (gc < n2) st_stream_v4(o, 0.f, 0.f, 0.f, 0.f);   // n2 % 4 == 0 => all four in range
        } else {
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (gc + q < n2) st_stream(o + q, 0.f);
        }
      }
      o += n2;
    }
  }
  unsigned surv[kMasks];                                // mask m, bit p: pair (row warp * kRPW + 8 * m + p / 4, column 4 * lane + p % 4) survives
#pragma unroll
  for (int m = 0; m < kMasks; m++) surv[m] = __brev(~rej[m]);
  // Survivors -> SAT -> device-wide queue, in rounds of at most kQCap queued survivors (a tile normally has a few
  // hundred; the small queues keep 8 CTAs resident per SM, which is what hides the load -> store -> atomic latency
  // chain of these short CTAs).  Whatever does not fit stays in the per-thread masks for the next round.
  for (;;) {
    {  // compact: warp exclusive scan of popcounts, one atomic per warp
      int cnt = 0;
#pragma unroll
      for (int m = 0; m < kMasks; m++) cnt += __popc(surv[m]);
      int incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int base = 0;
      if (lane == 31 && total > 0) base = atomicAdd(&s_cnt1, total);
      base = __shfl_sync(0xffffffffu, base, 31);
      int pos = base + incl - cnt;
#pragma unroll
      for (int m = 0; m < kMasks; m++) {
        while (surv[m] && pos < kQCap) {
          const int b = __ffs(surv[m]) - 1;
          surv[m] &= surv[m] - 1;
          const int r = warp * kRPW + 8 * m + (b >> 2), c = 4 * lane + (b & 3);
          s_q1[pos++] = (unsigned short)((r << 7) | c);
        }
      }
    }
    unsigned left = 0u;
#pragma unroll
    for (int m = 0; m < kMasks; m++) left |= surv[m];
    const int pending = __syncthreads_or(left != 0u);

    // ---- phase 2: SAT on circle survivors ----------------------------------------------------
    const int cnt1 = min(s_cnt1, kQCap);
    for (int base = 0; base < cnt1; base += kThreads) {
      const int k = base + tid;
      bool keep = false;
      unsigned short e = 0;
      if (k < cnt1) {
        e = s_q1[k];
        keep = !sat_disjoint<VERSION>(s_row[e >> 7], s_col[e & 127]);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (bal) {
        int wbase = 0;
        if (lane == 0) wbase = atomicAdd(&s_cnt2, __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (keep) s_q2[wbase + __popc(bal & ((1u << lane) - 1u))] = e;
      }
    }
    __syncthreads();

    // ---- hand-off: candidates go to the device-wide queue drained by iou_exact_kernel (every lane of
    // every warp busy there, no barriers); if the queue is full this CTA evaluates its own candidates.
    const int cnt2 = s_cnt2;
    if (tid == 0) s_base = cnt2 > 0 ? atomicAdd(gcount, (unsigned long long)cnt2) : 0ull;
    __syncthreads();
    const unsigned long long gbase = s_base;
    if (gbase + (unsigned long long)cnt2 <= (unsigned long long)gcap) {
      for (int k = tid; k < cnt2; k += kThreads) {
        const unsigned short e = s_q2[k];
        gqueue[gbase + k] = make_uint2((unsigned)(row0 + (e >> 7)), (unsigned)(col0 + (e & 127)));
      }
    } else {
      for (int k = tid; k < cnt2; k += kThreads) {
        if (gbase + (unsigned long long)k < (unsigned long long)gcap) gqueue[gbase + k] = make_uint2(0xffffffffu, 0u);   // reserved but unused slot
        const unsigned short e = s_q2[k];
        const int r = e >> 7, c = e & 127;
        const BoxRec& A = s_row[r];
        const BoxRec& B = s_col[c];
        out[(size_t)(row0 + r) * n2 + (col0 + c)] =   // (rare path: the compact routine, for this kernel's register count)
            (A.tag == 0.f && B.tag == 0.f) ? (variant ? iou_exact_general<VERSION, 1>(A, B) : iou_exact_general<VERSION, 0>(A, B)) : 0.f;
      }
    }
    if (!pending) break;
    __syncthreads();
    if (tid == 0) { s_cnt1 = 0; s_cnt2 = 0; }
    __syncthreads();
  }
}

// Exact IoU for the queued candidates: one thread per pair, grid-stride over the device-side count.  The clip points of
// a pair live in a [slot][thread] shared-memory array (48 KB per CTA, 4 CTAs per SM), see iou_exact_shared.
constexpr int kExactThreads = 256;
constexpr size_t kExactSmem = (size_t)3 * kExactCap * kExactThreads * sizeof(float);
template <int VERSION>
__global__ void __launch_bounds__(kExactThreads) iou_exact_kernel(const BoxRec* __restrict__ rec1, const BoxRec* __restrict__ rec2,
                                                         int n2, const unsigned long long* __restrict__ gcount,
                                                         const uint2* __restrict__ gqueue, int gcap,
                                                         float* __restrict__ out, int variant) {
  extern __shared__ float s_pts[];
  float* sq = s_pts + threadIdx.x;
  const unsigned long long reserved = *gcount;
  const int total = reserved < (unsigned long long)gcap ? (int)reserved : gcap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint2 e = gqueue[i];
    if (e.x == 0xffffffffu) continue;
    const BoxRec A = rec1[e.x], B = rec2[e.y];
    out[(size_t)e.x * n2 + e.y] =
        (A.tag == 0.f && B.tag == 0.f) ? (variant ? iou_exact_shared<VERSION, kExactThreads>(A, B, sq) : iou_exact<VERSION, 0>(A, B)) : 0.f;
  }
}

}  // namespace jdet

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
namespace jdet {
static size_t iou_queue_cap(int n1, int n2) {
  const long long pairs = (long long)n1 * n2;
  return (size_t)(pairs < (16ll << 20) ? pairs : (16ll << 20));     // <= 128 MB of (row, col) candidates
}
}  // namespace jdet

JDET_API size_t jdet_box_iou_rotated_workspace_bytes(int n1, int n2) {
  if (n1 <= 0 || n2 <= 0) return 256;
  return jdet_align_up((size_t)n1 * sizeof(jdet::BoxRec), 256) + jdet_align_up((size_t)n2 * sizeof(jdet::BoxRec), 256) +
         256 + jdet_align_up(jdet::iou_queue_cap(n1, n2) * sizeof(uint2), 256);
}

// version 0: jdet.ops.box_iou_rotated      (ops/box_iou_rotated.py:502-509)
// version 1: jdet.ops.box_iou_rotated_v1   (ops/box_iou_rotated_v1.py:507-525, incl. small-box zeroing)
// boxes1 (n1,5), boxes2 (n2,5), ious (n1,n2): device, fp32, contiguous.  Never syncs, never allocates.
JDET_API int jdet_box_iou_rotated_ex(const float* boxes1, int n1, const float* boxes2, int n2, float* ious, int version,
                                     int arithmetic, void* workspace, size_t workspace_bytes, void* stream);

JDET_API int jdet_box_iou_rotated(const float* boxes1, int n1, const float* boxes2, int n2, float* ious,
                                  int version, void* workspace, size_t workspace_bytes, void* stream) {
  return jdet_box_iou_rotated_ex(boxes1, n1, boxes2, n2, ious, version, /*arithmetic=*/1, workspace, workspace_bytes, stream);
}

// arithmetic 1: the reference's CUDA build (exchange-sort hull) — what jdet_box_iou_rotated computes;
// arithmetic 0: its CPU build (std::sort hull with the stale dist[] of box_iou_rotated.py:219-224), bit for bit.
JDET_API int jdet_box_iou_rotated_ex(const float* boxes1, int n1, const float* boxes2, int n2, float* ious, int version,
                                     int arithmetic, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace jdet;
  if (n1 < 0 || n2 < 0 || (version != 0 && version != 1) || (arithmetic != 0 && arithmetic != 1)) return JDET_ERR_BAD_ARG;
  const int variant = arithmetic;
  if (n1 == 0 || n2 == 0) return 0;
  if (!boxes1 || !boxes2 || !ious) return JDET_ERR_BAD_ARG;
  if (workspace_bytes < jdet_box_iou_rotated_workspace_bytes(n1, n2) || !workspace) return JDET_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* wsp = (char*)workspace;
  BoxRec* rec1 = (BoxRec*)wsp;                 wsp += jdet_align_up((size_t)n1 * sizeof(BoxRec), 256);
  BoxRec* rec2 = (BoxRec*)wsp;                 wsp += jdet_align_up((size_t)n2 * sizeof(BoxRec), 256);
  unsigned long long* gcount = (unsigned long long*)wsp;   wsp += 256;
  uint2* gqueue = (uint2*)wsp;
  int gcap = (int)iou_queue_cap(n1, n2);
  if (const char* e = getenv("JDET_TEST_QUEUE_CAP")) gcap = std::max(1, std::min(gcap, atoi(e)));   // tests: force the queue-full path
  rec_kernel<<<jdet_ceil_div(n1 + n2, 256), 256, 0, st>>>(boxes1, n1, boxes2, n2, version == 1, rec1, rec2, gcount);
  const bool vec = (n2 % 4 == 0) && (((uintptr_t)ious & 15) == 0);
  const long long pairs = (long long)n1 * n2;
  const bool big = pairs >= kBigTilePairs;
  dim3 grid(jdet_ceil_div(n2, kTC), jdet_ceil_div(n1, big ? 128 : 64));
  const int xgrid = (int)(pairs < 256 * 1024 ? (pairs + 255) / 256 : num_sms() * 8);
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(iou_exact_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExactSmem));
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(iou_exact_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExactSmem));
#define JDET_IOU_TILE(V, VEC, TR) iou_tile_kernel<V, VEC, TR><<<grid, kThreads, 0, st>>>(rec1, n1, rec2, n2, ious, gcount, gqueue, gcap, variant)
#define JDET_IOU_TILES(V)                                                                                              \
  do {                                                                                                                 \
    if (big) { if (vec) JDET_IOU_TILE(V, true, 128); else JDET_IOU_TILE(V, false, 128); }                              \
    else     { if (vec) JDET_IOU_TILE(V, true, 64);  else JDET_IOU_TILE(V, false, 64); }                               \
  } while (0)
  if (version == 0) {
    JDET_IOU_TILES(0);
    iou_exact_kernel<0><<<xgrid, kExactThreads, kExactSmem, st>>>(rec1, rec2, n2, gcount, gqueue, gcap, ious, variant);
  } else {
    JDET_IOU_TILES(1);
    iou_exact_kernel<1><<<xgrid, kExactThreads, kExactSmem, st>>>(rec1, rec2, n2, gcount, gqueue, gcap, ious, variant);
  }
#undef JDET_IOU_TILES
#undef JDET_IOU_TILE
  return (int)cudaGetLastError();
}
