// nms_rotated.cu — greedy rotated NMS, entirely on the device, for sm_100a.
//
// Replaces nms_rotated_cuda_kernel + its launch/host-reduce snippet
// (/root/reference/python/jdet/ops/nms_rotated.py:352-411, 450-493) behind the signature of
// nms_rotated_cuda(dets, order_t, iou_threshold, box_length) (:506-513).
//
// Reference shape of work: (N/64)^2 tiles over ALL pairs (both triangles, cross-class pairs exit
// inside the IoU routine), a dense N x N/64 u64 mask (1.25 GB at N = 100k) in managed memory,
// cudaDeviceSynchronize, then a single-threaded host loop.  Here:
//
//   1. stable radix sort of the score order by label  -> per-class segments, score order kept
//      inside each class (cross-class pairs can never suppress when thr >= 0: IoU := 0, test is >)
//   2. mask kernel: per segment only the upper-triangular 64x64 tiles; circle -> SAT -> exact IoU
//      with queue compaction between stages (see box_iou_rotated.cu); persistent CTAs pull
//      (segment, row-block, column-chunk) items from an atomic counter
//   3. scan kernel: one CTA per segment walks its 64-box blocks; the diagonal tile is resolved
//      inside one warp with ballots, kept rows are OR-reduced into the running suppression
//      words; keep flags are scattered straight to the caller's (original-index) keep array
//
// No host synchronisation, no managed memory, no allocation: scratch comes from the caller.
// Semantics are the reference CUDA path's: bit(i,j) = IoU(box_i, box_j) > thr for i ranked
// before j, IoU arguments in (higher, lower) order, label column compared with != first.
#include <cub/cub.cuh>

#include <math.h>
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "rbox_geom.cuh"

namespace jdet {

constexpr int kCH = 16;            // column blocks per work item (64 rows x 1024 columns)
constexpr int kColsPerThread = kCH * 64 / 256;
constexpr int kMaskThreads = 256;
#ifndef JDET_NMS_QCAP
#define JDET_NMS_QCAP 4096
#endif
#ifndef JDET_NMS_MASK_MINB
#define JDET_NMS_MASK_MINB 3
#endif
constexpr int kQCap = JDET_NMS_QCAP;   // survivor queue entries per CTA (overflow => extra rounds)
constexpr int kSplitTile = 512;        // boxes per CTA of the class split (256 threads x 2: 196 CTAs at 100k boxes)

struct NmsWs {
  unsigned* keys_in; unsigned* keys_out; int* vals_in; int* vals_out;
  BoxRec* rec; int* sorted_idx; int* flags; int* flag_scan; int* seg_start;
  int* seg_items; int* item_base; long long* seg_tiles; long long* tile_base;
  int* counters;                   // [0] work counter; bytes 8..15: the 64-bit candidate-queue reservation count; [8]: split tiles done
  unsigned char* keys8; int* tile_hist; int* tile_off; int* key_base;   // class split: key per box, per-tile key counts / offsets, first slot per key
  uint4* xqueue; int xcap;         // device-wide exact-IoU candidate queue
  unsigned long long* mask; size_t mask_tiles;
  void* cub_temp; size_t cub_bytes;
};

static size_t mask_tile_bound(int n) {
  const double N = (double)n;
  return (size_t)(N * N / 8192.0 + 3.0 * N / 128.0 + N + 64.0);
}

static size_t carve(NmsWs* w, void* base, int n) {
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off += jdet_align_up(bytes, 256); return p; };
  const size_t n1 = (size_t)n + 2;
  w->keys_in = (unsigned*)take(n1 * 4);  w->keys_out = (unsigned*)take(n1 * 4);
  w->vals_in = (int*)take(n1 * 4);       w->vals_out = (int*)take(n1 * 4);
  w->rec = (BoxRec*)take(n1 * sizeof(BoxRec));
  w->sorted_idx = (int*)take(n1 * 4);    w->flags = (int*)take(n1 * 4);
  w->flag_scan = (int*)take(n1 * 4);     w->seg_start = (int*)take(n1 * 4);
  w->seg_items = (int*)take(n1 * 4);     w->item_base = (int*)take(n1 * 4);
  w->seg_tiles = (long long*)take(n1 * 8); w->tile_base = (long long*)take(n1 * 8);
  w->counters = (int*)take(256);
  w->keys8 = (unsigned char*)take(n1);
  w->tile_hist = (int*)take((size_t)jdet_ceil_div(n, kSplitTile) * 256 * 4);
  w->tile_off = (int*)take((size_t)jdet_ceil_div(n, kSplitTile) * 256 * 4);
  w->key_base = (int*)take(256 * 4);
  w->cub_bytes = (size_t)n * 16 + (1u << 20);
  w->cub_temp = take(w->cub_bytes);
  w->xcap = (int)(((long long)n * n / 2 < (8ll << 20)) ? ((long long)n * n / 2 + 64) : (8ll << 20));
  w->xqueue = (uint4*)take((size_t)w->xcap * sizeof(uint4));
  if (const char* e = getenv("JDET_TEST_QUEUE_CAP")) w->xcap = std::max(1, std::min(w->xcap, atoi(e)));   // tests: force the queue-full path (the layout keeps its full size)
  w->mask_tiles = mask_tile_bound(n);
  w->mask = (unsigned long long*)take(w->mask_tiles * 64 * 8);
  return off;
}

// 8-bit GROUPING key of a label: segments only have to bring equal labels together (one radix pass
// instead of four); labels that share a key are told apart by the tag comparison in phase 2, exactly
// as the reference's `box1_raw[5] != box2_raw[5]` does.  Small integer class ids map to distinct keys.
__device__ __forceinline__ unsigned label_key(float l) {
  if (l == 0.f) l = 0.f;                       // -0.0 == +0.0 under the reference's != test
  if (l >= 0.f && l < 16777216.f && l == truncf(l)) return ((unsigned)l) & 255u;
  const unsigned u = __float_as_uint(l);
  return (u ^ (u >> 8) ^ (u >> 16) ^ (u >> 24)) & 255u;
}

__global__ void key_kernel(const float* __restrict__ dets, const int* __restrict__ order, int n,
                           unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = label_key(dets[(size_t)order[i] * 6 + 5]);
  vals[i] = i;
}

// rec[j] for sorted position j; flags[j] = 1 at segment heads.  pos == nullptr: identity, one segment.
__global__ void gather_kernel(const float* __restrict__ dets, int L, const int* __restrict__ order,
                              const int* __restrict__ pos, const unsigned* __restrict__ keys, int n,
                              int label_in_pair, BoxRec* __restrict__ rec, int* __restrict__ sorted_idx,
                              int* __restrict__ flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  if (j == n) { flags[n] = 0; return; }
  const int src = order[pos ? pos[j] : j];
  const float* b = dets + (size_t)src * L;
  const float tag = (L == 6) ? b[5] : 0.f;
  rec[j] = make_rec(b[0], b[1], b[2], b[3], b[4], tag, false, L == 6);
  sorted_idx[j] = src;
  flags[j] = (j == 0) ? 1 : ((pos && !label_in_pair) ? (keys[j] != keys[j - 1]) : 0);
}

__global__ void seg_start_kernel(const int* __restrict__ flags, const int* __restrict__ scan, int n,
                                 int* __restrict__ seg_start) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  if (j == n) { seg_start[scan[n - 1]] = n; return; }   // sentinel after the last segment
  if (flags[j]) seg_start[scan[j] - 1] = j;
}

// F(m) = sum_{k=1..m} ceil(k / CH)
__host__ __device__ __forceinline__ long long items_prefix(long long m) {
  const long long q = m / kCH, r = m % kCH;
  return (long long)kCH * q * (q + 1) / 2 + r * (q + 1);
}

__device__ __forceinline__ int items_prefix32(int m) {   // the same in 32 bits (m <= 23438 column blocks: n <= 1.5 M)
  const int q = m / kCH, r = m % kCH;
  return (kCH / 2) * q * (q + 1) + r * (q + 1);
}

__global__ void seg_count_kernel(const int* __restrict__ seg_start, const int* __restrict__ scan, int n,
                                 int* __restrict__ seg_items, long long* __restrict__ seg_tiles) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n) return;
  const int nseg = scan[n - 1];
  int items = 0; long long tiles = 0;
  if (s < nseg) {
    const int cnt = seg_start[s + 1] - seg_start[s];
    const long long W = (cnt + 63) / 64;
    items = (int)items_prefix(W);
    tiles = W * (W + 1) / 2;
  }
  seg_items[s] = items; seg_tiles[s] = tiles;
}

// ---- class split: score order -> per-class segments, score order kept inside a class --------------------------------
// Two launches instead of key + radix sort (3) + gather + 3 device scans (6) + 2 segment kernels, whose ~60 us were launch
// latency on 100k boxes: a stable one-pass multi-split on the 8-bit grouping key.
//   split_count_kernel    per tile of 2048 boxes (in score order): the key of every box, the tile's key histogram; the LAST
//                         tile to finish turns the histograms into per-(tile, key) output bases and writes the segment
//                         tables the mask / scan kernels read (seg_start, item_base, tile_base, segment count)
//   split_scatter_kernel  per tile: the stable rank of every box among its tile's boxes of the same key (warp match +
//                         counters per (round, warp, key)), then rec / sorted_idx at base + rank
__global__ void __launch_bounds__(256) split_count_kernel(const float* __restrict__ dets, const int* __restrict__ order, int n, int ntiles,
                                                          unsigned char* __restrict__ keys8, int* __restrict__ tile_hist,
                                                          int* __restrict__ tile_off, int* __restrict__ key_base, int* __restrict__ counters, int* __restrict__ seg_start,
                                                          int* __restrict__ item_base, long long* __restrict__ tile_base,
                                                          int* __restrict__ nseg_out, unsigned char* __restrict__ keep) {
  __shared__ int s_hist[256];
  __shared__ int s_last;
  const int tid = threadIdx.x, tile = blockIdx.x;
  s_hist[tid] = 0;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSplitTile / 256; r++) {
    const int i = tile * kSplitTile + r * 256 + tid;
    if (i < n) {
      const unsigned k = label_key(dets[(size_t)order[i] * 6 + 5]);
      keys8[i] = (unsigned char)k;
      keep[i] = 0;                                       // (the scan kernel sets the kept ones)
      atomicAdd(&s_hist[k], 1);
    }
  }
  __syncthreads();
  tile_hist[(size_t)tile * 256 + tid] = s_hist[tid];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(counters + 8, 1) == ntiles - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- last tile: thread k owns key k
  // tile_off[t][key] = boxes of this key in the tiles before t.  32 loads are issued before the first store: ptxas keeps
  // these GPU-scope loads behind earlier stores, and one L2 round trip per tile made this tail 49 us at 196 tiles
  int run = 0;
  for (int t0 = 0; t0 < ntiles; t0 += 32) {
    int c[32];
#pragma unroll
    for (int j = 0; j < 32; j++) c[j] = t0 + j < ntiles ? __ldcg(tile_hist + (size_t)(t0 + j) * 256 + tid) : 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
      if (t0 + j < ntiles) tile_off[(size_t)(t0 + j) * 256 + tid] = run;
      run += c[j];
    }
  }
  __shared__ int s_a[256], s_seg[256];
  __shared__ long long s_t[256];
  s_a[tid] = run;                                        // boxes of key tid
  __syncthreads();
  int key_start = 0, seg = 0;
  for (int k = 0; k < tid; k++) { key_start += s_a[k]; seg += s_a[k] > 0; }     // 256 x 256 adds: nothing
  key_base[tid] = key_start;                             // first output slot of key tid
  const long long Wk = (run + 63) / 64;
  __syncthreads();                                       // everyone has read the counts
  s_seg[tid] = run > 0 ? seg : -1;
  s_a[tid] = run > 0 ? (int)items_prefix(Wk) : 0;        // work items / mask tiles of this key's segment
  s_t[tid] = run > 0 ? Wk * (Wk + 1) / 2 : 0;
  if (run > 0) seg_start[seg] = key_start;
  __syncthreads();
  if (run > 0) {
    int ib = 0; long long tb = 0;
    for (int k = 0; k < tid; k++) { ib += s_a[k]; tb += s_t[k]; }
    item_base[seg] = ib; tile_base[seg] = tb;
  }
  if (tid == 255) {
    int ib = 0, ns = 0; long long tb = 0;
    for (int k = 0; k < 256; k++) { ib += s_a[k]; tb += s_t[k]; ns += s_seg[k] >= 0; }
    seg_start[ns] = n; item_base[ns] = ib; tile_base[ns] = tb;
    *nseg_out = ns;
  }
}

__global__ void __launch_bounds__(256) split_scatter_kernel(const float* __restrict__ dets, const int* __restrict__ order, int n,
                                                            const unsigned char* __restrict__ keys8, const int* __restrict__ tile_off,
                                                            const int* __restrict__ key_base,
                                                            BoxRec* __restrict__ rec, int* __restrict__ sorted_idx) {
  constexpr int R = kSplitTile / 256;
  __shared__ unsigned short s_cnt[R][8][256];            // boxes of (round, warp) with this key -> exclusive prefix in (round, warp) order
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, tile = blockIdx.x;
  for (int i = tid; i < R * 8 * 256; i += 256) (&s_cnt[0][0][0])[i] = 0;
  __syncthreads();
  unsigned key[R];
  int lrank[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = tile * kSplitTile + r * 256 + tid;
    key[r] = i < n ? (unsigned)keys8[i] : 256u + (unsigned)lane;      // absent boxes match nobody
    const unsigned m = __match_any_sync(0xffffffffu, key[r]);
    lrank[r] = __popc(m & ((1u << lane) - 1u));
    if (i < n && lrank[r] == 0) s_cnt[r][warp][key[r]] = (unsigned short)__popc(m);
  }
  __syncthreads();
  {
    unsigned run = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
      for (int w = 0; w < 8; w++) { const unsigned c = s_cnt[r][w][tid]; s_cnt[r][w][tid] = (unsigned short)run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int i = tile * kSplitTile + r * 256 + tid;
    if (i >= n) continue;
    const int pos = key_base[key[r]] + tile_off[(size_t)tile * 256 + key[r]] + s_cnt[r][warp][key[r]] + lrank[r];
    const int src = order[i];
    const float* b = dets + (size_t)src * 6;
    rec[pos] = make_rec(b[0], b[1], b[2], b[3], b[4], b[5], false, true);
    sorted_idx[pos] = src;
  }
}

// one segment (box_length 5, or the thr < 0 corner): the tables the class split writes otherwise
__global__ void single_segment_kernel(int n, int* __restrict__ seg_start, int* __restrict__ item_base, long long* __restrict__ tile_base,
                                      int* __restrict__ nseg_out) {
  const long long W = (n + 63) / 64;
  seg_start[0] = 0; seg_start[1] = n;
  item_base[0] = 0; item_base[1] = (int)items_prefix(W);
  tile_base[0] = 0; tile_base[1] = W * (W + 1) / 2;
  *nseg_out = 1;
}

// Mask layout: per segment, upper-triangular 64x64 tiles in COLUMN-block-major order, so the scan
// warp that owns column block cb streams tiles (0..cb, cb) from one contiguous run.
__device__ __forceinline__ long long tri_index(long long rb, long long cb) { return cb * (cb + 1) / 2 + rb; }

// ---- mask kernel ---------------------------------------------------------------------------------
// One work item = (segment, row block rb, chunk of up to kCH column blocks).  256 threads.
//   phase 1  thread <-> 2 columns (registers), loop over the 64 rows (broadcast LDS.128): circle test
//   phase 2  SAT on compacted survivors          phase 3  exact IoU, strict "> thr", bits via smem atomics
// label_in_pair != 0 (thr < 0 corner): single segment, no quick rejects, label mismatch => IoU 0.
__global__ void __launch_bounds__(kMaskThreads, JDET_NMS_MASK_MINB) nms_mask_kernel(
    const BoxRec* __restrict__ rec, const int* __restrict__ seg_start, const int* __restrict__ item_base,
    const long long* __restrict__ tile_base, const int* __restrict__ scan, int n, float thr,
    int flags, int* __restrict__ counter, unsigned long long* __restrict__ mask,
    uint4* __restrict__ xqueue, int xcap) {
  const int label_in_pair = flags & 1;       // thr < 0 corner: single segment, label mismatch => IoU 0
  const bool cpu_arith = (flags & 2) != 0;   // the reference CPU build's hull sort; no IoU-upper-bound pruning (its IoU is not bounded by the true one)
  const bool no_prune = (flags & 4) != 0;               // JDET_NMS_NO_PRUNE=1 (tests): every SAT survivor takes the exact routine
  extern __shared__ __align__(16) unsigned char s_dyn[];
  BoxRec* s_row = reinterpret_cast<BoxRec*>(s_dyn);                       //  2 KB
  BoxRec* s_col = s_row + 64;                                             // 16 KB
  float4* s_rowq = reinterpret_cast<float4*>(s_col + kCH * 64);           //  1 KB  (x, y, qr, -)
  unsigned* s_bits = reinterpret_cast<unsigned*>(s_rowq + 64);            //  4 KB: tile words as (lo, hi)
  unsigned short* s_q1 = reinterpret_cast<unsigned short*>(s_bits + kCH * 64 * 2);   // 2 x kQCap entries
  unsigned short* s_q2 = s_q1 + kQCap;
  __shared__ int s_cnt1, s_cnt2, s_item, s_seg, s_rb, s_cb0;
  __shared__ unsigned long long s_xbase;   // 64-bit: a dense same-class cluster of ~65k boxes reserves more than 2^31 candidates

  const int tid = threadIdx.x, lane = tid & 31;
  const int nseg = scan[n - 1];
  const int total = item_base[nseg];

  int next_item = 0;                                     // (warp 0) fetched one item ahead: the atomic's round trip is off the path
  if (tid == 0) next_item = atomicAdd(counter, 1);
  for (;;) {
    __syncthreads();
    if (tid < 32) {
      // item -> (segment, row block, first column block): two binary searches, done by ONE warp in 32-bit arithmetic
      // (n <= 1.5 M: F(W) < 2^25); every warp decoding its own copy in 64 bits was 12 % of the kernel's instructions
      const int item = __shfl_sync(0xffffffffu, next_item, 0);
      if (lane == 0) next_item = atomicAdd(counter, 1);
      int seg = 0, rb = 0, cb0 = 0;
      if (item < total) {
        int lo = 0, hi = nseg;                           // item_base[lo] <= item < item_base[hi]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (item_base[mid] <= item) lo = mid; else hi = mid; }
        seg = lo;
        const int W = (seg_start[seg + 1] - seg_start[seg] + 63) >> 6;
        const int local = item - item_base[seg];
        const int FW = items_prefix32(W);
        int rlo = 0, rhi = W;                            // cum(rb) = F(W) - F(W - rb) <= local
        while (rhi - rlo > 1) { const int mid = (rlo + rhi) >> 1; if (FW - items_prefix32(W - mid) <= local) rlo = mid; else rhi = mid; }
        rb = rlo;
        cb0 = rb + (local - (FW - items_prefix32(W - rb))) * kCH;
      }
      if (lane == 0) { s_item = item; s_seg = seg; s_rb = rb; s_cb0 = cb0; }
    }
    __syncthreads();
    const int item = s_item;
    if (item >= total) break;
    const int seg = s_seg, rb = s_rb, cb0 = s_cb0;
    const int s0 = seg_start[seg], cnt = seg_start[seg + 1] - s0;
    const int W = (cnt + 63) >> 6;
    const int ncb = min(kCH, W - cb0);
    const int ncols = ncb * 64;

    const BoxRec dead{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, -INFINITY, 0.f};
    if (tid < 64) {
      const int p = rb * 64 + tid;
      const BoxRec r = (p < cnt) ? rec[s0 + p] : dead;
      s_row[tid] = r;
      s_rowq[tid] = make_float4(r.x, r.y, r.qr, 0.f);
    }
    // this thread's columns stay in registers for phase 1
    float cx[kColsPerThread], cy[kColsPerThread], cq[kColsPerThread];
#pragma unroll
    for (int j = 0; j < kColsPerThread; j++) {
      const int c = tid + j * kMaskThreads;
      BoxRec r = dead;
      if (c < ncols) {
        const int p = cb0 * 64 + c;
        if (p < cnt) r = rec[s0 + p];
        s_col[c] = r;
      }
      cx[j] = r.x; cy[j] = r.y; cq[j] = r.qr;
    }
    for (int k = tid; k < ncb * 128; k += kMaskThreads) s_bits[k] = 0u;
    if (tid == 0) { s_cnt1 = 0; s_cnt2 = 0; }
    __syncthreads();

    // ---- phase 1 ---------------------------------------------------------------------------------
    unsigned long long m[kColsPerThread];                 // bit r: (row r, my column j) survives
    if (!label_in_pair) {
      // the verdict of a circle test is the SIGN of R|R| - d^2, shifted into the word by one funnel shift (row r lands
      // in bit 31 - r; one bit reversal per word at the end): 7 issue slots per pair, no compare / select / OR
      unsigned wlo[kColsPerThread], whi[kColsPerThread];
#pragma unroll
      for (int j = 0; j < kColsPerThread; j++) { wlo[j] = 0u; whi[j] = 0u; }
#pragma unroll
      for (int r = 0; r < 32; r++) {
        // rows r and r + 32 against column j in one packed pass (FADD2 / FMUL2 / FFMA2: 4 issue slots per pair, not 7)
        const float4 ra = s_rowq[r], rb2 = s_rowq[r + 32];
        const unsigned long long X1 = f2_pack(ra.x, rb2.x), Y1 = f2_pack(ra.y, rb2.y), Q1 = f2_pack(ra.z, rb2.z);
#pragma unroll
        for (int j = 0; j < kColsPerThread; j++) {
          float ta, tb;
          circle_t2(X1, Y1, Q1, f2_pack(cx[j], cx[j]), f2_pack(cy[j], cy[j]), f2_pack(cq[j], cq[j]), ta, tb);
          wlo[j] = __funnelshift_l(__float_as_uint(ta), wlo[j], 1);
          whi[j] = __funnelshift_l(__float_as_uint(tb), whi[j], 1);
        }
      }
#pragma unroll
      for (int j = 0; j < kColsPerThread; j++) m[j] = ((unsigned long long)__brev(~whi[j]) << 32) | __brev(~wlo[j]);
    } else {
      const int nrow = min(64, cnt - rb * 64);
      const unsigned long long rows = nrow >= 64 ? ~0ull : ((1ull << nrow) - 1ull);
#pragma unroll
      for (int j = 0; j < kColsPerThread; j++) {
        const int c = tid + j * kMaskThreads;
        m[j] = (c < ncols && cb0 * 64 + c < cnt) ? rows : 0ull;
      }
    }
    // strictly upper triangle: in the diagonal tile (cb == rb) column c only meets rows r < c
    if (cb0 == rb) {
#pragma unroll
      for (int j = 0; j < kColsPerThread; j++) {
        const int c = tid + j * kMaskThreads;
        if (c < 64) m[j] &= (c == 0) ? 0ull : ((1ull << c) - 1ull);
      }
    }
    // Survivors -> queue.  Candidates per item are normally a few hundred; if they exceed kQCap the
    // remainder stays in the per-thread masks and goes through another round of phases 2 and 3.
    for (;;) {
      int want = 0;
#pragma unroll
      for (int j = 0; j < kColsPerThread; j++) want += __popcll(m[j]);
      int incl = want;                                   // warp inclusive scan, one atomic per warp
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
      const int tot = __shfl_sync(0xffffffffu, incl, 31);
      int base = 0;
      if (lane == 31 && tot > 0) base = atomicAdd(&s_cnt1, tot);
      base = __shfl_sync(0xffffffffu, base, 31);
      int pos = base + incl - want;
#pragma unroll
      for (int j = 0; j < kColsPerThread; j++) {
        const int c = tid + j * kMaskThreads;
        while (m[j] && pos < kQCap) {
          const int r = __ffsll((long long)m[j]) - 1;
          m[j] &= m[j] - 1;
          s_q1[pos++] = (unsigned short)(((c >> 6) << 12) | (r << 6) | (c & 63));
        }
      }
      unsigned long long left = 0ull;
#pragma unroll
      for (int j = 0; j < kColsPerThread; j++) left |= m[j];
      const int pending = __syncthreads_or(left != 0ull);
      const int c1 = min(s_cnt1, kQCap);
      // ---- phase 2: SAT ----------------------------------------------------------------------------
      for (int b0 = 0; b0 < c1; b0 += kMaskThreads) {
        const int k = b0 + tid;
        bool keep = false; unsigned short e = 0;
        if (k < c1) {
          e = s_q1[k];
          const BoxRec& A = s_row[(e >> 6) & 63];
          const BoxRec& B = s_col[(e >> 12) * 64 + (e & 63)];
          // SAT reject, then the IoU upper bound: a pair that provably cannot exceed thr never reaches phase 3
          keep = label_in_pair ? true : (A.tag == B.tag && !sat_disjoint<0>(A, B) && (cpu_arith || no_prune || !(iou_upper_bound<0>(A, B) < thr)));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (bal) {
          int wb = 0;
          if (lane == 0) wb = atomicAdd(&s_cnt2, __popc(bal));
          wb = __shfl_sync(0xffffffffu, wb, 0);
          if (keep) s_q2[wb + __popc(bal & ((1u << lane) - 1u))] = e;
        }
      }
      __syncthreads();
      // ---- hand-off: exact-IoU candidates go to the device-wide queue (nms_exact_kernel: every lane busy,
      // no barriers).  Queue full => this CTA evaluates its own candidates (phase 3 in place).
      const int c2 = s_cnt2;
      if (tid == 0) s_xbase = c2 > 0 ? atomicAdd(reinterpret_cast<unsigned long long*>(counter + 2), (unsigned long long)c2) : 0ull;
      __syncthreads();
      const unsigned long long xbase = s_xbase;
      const long long wbase = tile_base[seg] * 64;
      if (xbase + (unsigned long long)c2 <= (unsigned long long)xcap) {
        for (int k = tid; k < c2; k += kMaskThreads) {
          const unsigned short e = s_q2[k];
          const int r = (e >> 6) & 63, tt = e >> 12, cc = e & 63;
          const long long word = wbase + tri_index(rb, cb0 + tt) * 64 + r;
          xqueue[xbase + k] = make_uint4((unsigned)(s0 + rb * 64 + r), (unsigned)(s0 + (cb0 + tt) * 64 + cc),
                                         (unsigned)(word & 0xffffffffll), (unsigned)((word >> 32) << 8) | (unsigned)cc);
        }
      } else {
        for (int k = tid; k < c2; k += kMaskThreads) {
          if (xbase + (unsigned long long)k < (unsigned long long)xcap) xqueue[xbase + k] = make_uint4(0xffffffffu, 0u, 0u, 0u);   // reserved, unused
          const unsigned short e = s_q2[k];
          const int r = (e >> 6) & 63, tt = e >> 12, cc = e & 63;
          const BoxRec& A = s_row[r];
          const BoxRec& B = s_col[tt * 64 + cc];
          float v;
          if (label_in_pair && A.tag != B.tag) v = 0.f;   // nms_rotated.py:285-286
          else v = cpu_arith ? iou_exact_general<0, 0>(A, B) : iou_exact_general<0, 1>(A, B);   // (rare path: the compact routine keeps this kernel's register count down)
          if (v > thr) atomicOr(&s_bits[(tt * 64 + r) * 2 + (cc >> 5)], 1u << (cc & 31));
        }
      }
      __syncthreads();
      if (!pending) break;
      if (tid == 0) { s_cnt1 = 0; s_cnt2 = 0; }
      __syncthreads();
    }
    // tiles (rb, cb0 + t): column-block-major triangular layout
    unsigned long long* mbase = mask + tile_base[seg] * 64;
    for (int k = tid; k < ncb * 64; k += kMaskThreads) {
      const int t = k >> 6, r = k & 63;
      mbase[tri_index(rb, cb0 + t) * 64 + r] = ((unsigned long long)s_bits[2 * k + 1] << 32) | s_bits[2 * k];
    }
  }
}

// ---- exact kernel --------------------------------------------------------------------------------
// One thread per queued candidate: reference-exact IoU(higher-ranked, lower-ranked), strict "> thr",
// bit set with a 64-bit atomicOr straight into the triangular mask written (as zeros) by the mask kernel.
__global__ void __launch_bounds__(256) nms_exact_kernel(const BoxRec* __restrict__ rec, const int* __restrict__ counter,
                                                         const uint4* __restrict__ xqueue, int xcap, float thr,
                                                         int flags, unsigned long long* __restrict__ mask) {
  const int label_in_pair = flags & 1;
  const bool cpu_arith = (flags & 2) != 0;
  extern __shared__ float s_pts[];                     // clip points, [slot][thread]: see iou_exact_shared
  float* sq = s_pts + threadIdx.x;
  const unsigned long long reserved = *reinterpret_cast<const unsigned long long*>(counter + 2);
  const int total = reserved < (unsigned long long)xcap ? (int)reserved : xcap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint4 e = xqueue[i];
    if (e.x == 0xffffffffu) continue;
    const BoxRec A = rec[e.x], B = rec[e.y];
    float v;
    if (label_in_pair && A.tag != B.tag) v = 0.f;
    else v = cpu_arith ? iou_exact<0, 0>(A, B) : iou_exact_shared<0, 256>(A, B, sq);
    if (v > thr) {
      const unsigned long long word = ((unsigned long long)(e.w >> 8) << 32) | e.z;
      atomicOr(mask + word, 1ull << (e.w & 63u));
    }
  }
}

// ---- scan kernel ---------------------------------------------------------------------------------
// One CTA per segment (persistent over segments), kScanWarps warps.  Warp w OWNS column blocks
// cb = w, w + kScanWarps, ...: it ORs the rows kept in earlier blocks into its own running word
// (a warp-uniform register — there is no shared remv[] array), resolves its diagonal tile with
// ballots, writes the keep flags, and publishes kept(cb) + a progress counter through shared
// memory.  Loads never depend on kept(), so each warp prefetches a batch of tiles and only the
// masking waits; the serial chain per block is "apply tile (cb-1, cb) -> resolve diag -> publish".
constexpr int kScanWarps = 24;
constexpr int kScanThreads = kScanWarps * 32;
constexpr int kScanBatch = 12;

__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(
    const int* __restrict__ seg_start, const long long* __restrict__ tile_base, const int* __restrict__ scan,
    int n, const int* __restrict__ sorted_idx, const unsigned long long* __restrict__ mask,
    unsigned char* __restrict__ keep) {
  extern __shared__ unsigned long long s_kept[];       // [W]
  // number of column blocks resolved so far.  Flag hand-off between warps: the owner stores kept(cb), fences, then
  // advances the counter; readers poll the (volatile) counter, fence, then read kept(rb).  (ld.acquire.cta /
  // st.release.cta instead of volatile + __threadfence_block was measured: the polling loop became 3x slower, the
  // whole 100k x 15 op 0.66 -> 0.81 ms.  compute-sanitizer racecheck reports this hand-off as hazards either way —
  // it only models barriers — see profiles/r01_sanitizer_reentry.txt.)
  __shared__ volatile int s_progress;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nseg = scan[n - 1];
  for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
    const int s0 = seg_start[seg], cnt = seg_start[seg + 1] - s0;
    const int W = (cnt + 63) >> 6;
    const unsigned long long* base = mask + tile_base[seg] * 64;
    __syncthreads();
    if (tid == 0) s_progress = 0;
    __syncthreads();
    for (int cb = warp; cb < W; cb += kScanWarps) {
      const unsigned long long* col = base + tri_index(0, cb) * 64;     // tiles (0..cb, cb), contiguous
      // the diagonal tile and the scatter indices do not depend on anything: fetch them before the chain
      const unsigned long long d0 = col[(size_t)cb * 64 + lane], d1 = col[(size_t)cb * 64 + lane + 32];
      const int p0 = cb * 64 + lane, p1 = p0 + 32;
      const int idx0 = p0 < cnt ? sorted_idx[s0 + p0] : 0, idx1 = p1 < cnt ? sorted_idx[s0 + p1] : 0;
      unsigned acc_lo = 0u, acc_hi = 0u;
      for (int rb0 = 0; rb0 < cb; rb0 += kScanBatch) {
        const int nb = min(kScanBatch, cb - rb0);
        unsigned long long w0[kScanBatch], w1[kScanBatch];
#pragma unroll
        for (int b = 0; b < kScanBatch; b++) {
          if (b < nb) { w0[b] = col[(size_t)(rb0 + b) * 64 + lane]; w1[b] = col[(size_t)(rb0 + b) * 64 + lane + 32]; }
          else { w0[b] = 0ull; w1[b] = 0ull; }
        }
#pragma unroll
        for (int b = 0; b < kScanBatch; b++) {
          if (b < nb) {
            const int rb = rb0 + b;
            while (s_progress <= rb) { }                                  // kept(rb) not published yet
            __threadfence_block();
            const unsigned long long k = s_kept[rb];
            unsigned long long v = 0ull;
            if ((k >> lane) & 1ull) v |= w0[b];
            if ((k >> (lane + 32)) & 1ull) v |= w1[b];
            acc_lo |= __reduce_or_sync(0xffffffffu, (unsigned)v);
            acc_hi |= __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
          }
        }
      }
      // diagonal tile
      unsigned long long cur = ((unsigned long long)acc_hi << 32) | acc_lo;
      const unsigned nz0 = __ballot_sync(0xffffffffu, d0 != 0ull), nz1 = __ballot_sync(0xffffffffu, d1 != 0ull);
      const unsigned long long nz = ((unsigned long long)nz1 << 32) | nz0;
      unsigned long long done = 0ull;
      for (;;) {
        const unsigned long long cand = nz & ~cur & ~done;
        if (!cand) break;
        const int r = __ffsll((long long)cand) - 1;       // lowest kept row with a non-empty word
        const unsigned long long wsel = (r < 32) ? d0 : d1;
        const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)wsel, r & 31);
        const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(wsel >> 32), r & 31);
        cur |= ((unsigned long long)hi << 32) | lo;
        done |= 1ull << r;
      }
      const int valid = min(64, cnt - cb * 64);
      const unsigned long long vmask = valid >= 64 ? ~0ull : ((1ull << valid) - 1ull);
      const unsigned long long kept = ~cur & vmask;
      if (lane == 0) {
        while (s_progress < cb) { }                       // publish strictly in order
        s_kept[cb] = kept;
        __threadfence_block();
        s_progress = cb + 1;
      }
      if ((kept >> lane) & 1ull) keep[idx0] = 1;
      if ((kept >> (lane + 32)) & 1ull) keep[idx1] = 1;
    }
  }
}

}  // namespace jdet

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
JDET_API size_t jdet_nms_rotated_workspace_bytes(int n, int box_length) {
  (void)box_length;
  if (n <= 0) return 256;
  jdet::NmsWs w;
  return jdet::carve(&w, nullptr, n);
}

// dets (n, box_length) fp32, box_length in {5,6}: [x,y,w,h,theta(,label)]; order (n,) int32 = indices
// by descending score; keep (n,) bytes, written in full (1 = kept), indexed like dets.
// == nms_rotated_cuda(dets, order_t, iou_threshold, box_length), ops/nms_rotated.py:506-513.
JDET_API int jdet_nms_rotated_ex(const float* dets, int n, int box_length, const int* order, float iou_threshold, int convention,
                                 unsigned char* keep, void* workspace, size_t workspace_bytes, void* stream);

JDET_API int jdet_nms_rotated(const float* dets, int n, int box_length, const int* order, float iou_threshold,
                              unsigned char* keep, void* workspace, size_t workspace_bytes, void* stream) {
  return jdet_nms_rotated_ex(dets, n, box_length, order, iou_threshold, /*convention=*/1, keep, workspace, workspace_bytes, stream);
}

// convention 1: nms_rotated_cuda (ops/nms_rotated.py:506-513) — suppress on IoU > thr, CUDA-build IoU arithmetic.
// convention 0: nms_rotated_cpu  (ops/nms_rotated.py:495-504, loop :414-449) — suppress on IoU >= thr with the CPU build's
//               IoU arithmetic (std::sort hull).  In fp32 `x >= t` is `x > pred(t)`; the greedy loop's result equals the
//               mask-and-scan's (a box is suppressed iff a KEPT higher-ranked box overlaps it).
JDET_API int jdet_nms_rotated_ex(const float* dets, int n, int box_length, const int* order, float iou_threshold, int convention,
                                 unsigned char* keep, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace jdet;
  if (n < 0 || (box_length != 5 && box_length != 6) || (convention != 0 && convention != 1)) return JDET_ERR_BAD_ARG;
  if (convention == 0) iou_threshold = nextafterf(iou_threshold, -INFINITY);
  if (n == 0) return 0;
  if (!dets || !order || !keep || !workspace) return JDET_ERR_BAD_ARG;
  if (n > 1500000) return JDET_ERR_UNSUPPORTED;      // scan kernel keeps ceil(n/64) words in smem
  NmsWs w;
  if (carve(&w, workspace, n) > workspace_bytes) return JDET_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int T = 256, G = jdet_ceil_div(n + 1, T);
  const int label_in_pair = (iou_threshold < 0.f) ? 1 : 0;   // cross-class pairs then DO suppress (0 > thr)
  const int flags = label_in_pair | (convention == 0 ? 2 : 0) | (getenv("JDET_NMS_NO_PRUNE") ? 4 : 0);
  const bool segment = (box_length == 6) && !label_in_pair;

  JDET_RETURN_IF_CUDA(cudaMemsetAsync(w.counters, 0, 256, st));
  static const bool legacy_split = getenv("JDET_NMS_LEGACY_SPLIT") != nullptr;   // A/B: the radix-sort + device-scan segment build
  if (segment && !legacy_split) {
    const int ntiles = jdet_ceil_div(n, kSplitTile);
    split_count_kernel<<<ntiles, 256, 0, st>>>(dets, order, n, ntiles, w.keys8, w.tile_hist, w.tile_off, w.key_base, w.counters, w.seg_start, w.item_base,
                                               w.tile_base, w.flag_scan + (n - 1), keep);
    split_scatter_kernel<<<ntiles, 256, 0, st>>>(dets, order, n, w.keys8, w.tile_off, w.key_base, w.rec, w.sorted_idx);
  } else if (!segment) {
    JDET_RETURN_IF_CUDA(cudaMemsetAsync(keep, 0, (size_t)n, st));
    gather_kernel<<<G, T, 0, st>>>(dets, box_length, order, nullptr, w.keys_out, n, label_in_pair, w.rec, w.sorted_idx, w.flags);
    single_segment_kernel<<<1, 1, 0, st>>>(n, w.seg_start, w.item_base, w.tile_base, w.flag_scan + (n - 1));
  } else {
    JDET_RETURN_IF_CUDA(cudaMemsetAsync(keep, 0, (size_t)n, st));
    key_kernel<<<G, T, 0, st>>>(dets, order, n, w.keys_in, w.vals_in);
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, w.keys_in, w.keys_out, w.vals_in, w.vals_out, n, 0, 8, st);
    if (need > w.cub_bytes) return JDET_ERR_WORKSPACE;
    need = w.cub_bytes;
    JDET_RETURN_IF_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, need, w.keys_in, w.keys_out, w.vals_in,
                                                       w.vals_out, n, 0, 8, st));
    gather_kernel<<<G, T, 0, st>>>(dets, box_length, order, w.vals_out, w.keys_out, n, label_in_pair, w.rec,
                                   w.sorted_idx, w.flags);
    need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, w.flags, w.flag_scan, n, st);
    if (need > w.cub_bytes) return JDET_ERR_WORKSPACE;
    need = w.cub_bytes;
    JDET_RETURN_IF_CUDA(cub::DeviceScan::InclusiveSum(w.cub_temp, need, w.flags, w.flag_scan, n, st));
    seg_start_kernel<<<G, T, 0, st>>>(w.flags, w.flag_scan, n, w.seg_start);
    seg_count_kernel<<<G, T, 0, st>>>(w.seg_start, w.flag_scan, n, w.seg_items, w.seg_tiles);
    need = w.cub_bytes;
    JDET_RETURN_IF_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, need, w.seg_items, w.item_base, n + 1, st));
    need = w.cub_bytes;
    JDET_RETURN_IF_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_temp, need, w.seg_tiles, w.tile_base, n + 1, st));
  }
  const size_t mask_smem = (size_t)(64 + kCH * 64) * sizeof(BoxRec) + 64 * 16 + (size_t)kCH * 64 * 2 * 4 + (size_t)kQCap * 2 * 2;
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(nms_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mask_smem));
  nms_mask_kernel<<<num_sms() * JDET_NMS_MASK_MINB, kMaskThreads, mask_smem, st>>>(w.rec, w.seg_start, w.item_base, w.tile_base,
                                                        w.flag_scan, n, iou_threshold, flags,
                                                        w.counters, w.mask, w.xqueue, w.xcap);
  const size_t exact_smem = (size_t)3 * kExactCap * 256 * sizeof(float);
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(nms_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exact_smem));
  nms_exact_kernel<<<num_sms() * 8, 256, exact_smem, st>>>(w.rec, w.counters, w.xqueue, w.xcap, iou_threshold, flags, w.mask);
  const size_t smem = (size_t)jdet_ceil_div(n, 64) * 8;
  if (smem > 48 * 1024)
    JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_scan_kernel<<<num_sms(), kScanThreads, smem, st>>>(w.seg_start, w.tile_base, w.flag_scan, n, w.sorted_idx,
                                                       w.mask, keep);
  return (int)cudaGetLastError();
}

// Stable descending argsort of fp32 scores (ties: lower index first) — the order nms_rotated wants.
// Replaces scores.argsort(0, descending=True) (ops/nms_rotated.py:519,532).
namespace jdet {
__global__ void score_key_kernel(const float* __restrict__ s, int n, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = s[i];
  if (v == 0.f) v = 0.f;                             // -0.0 and +0.0 are one score: ties go to the lower index
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending-orderable
  keys[i] = ~u;                                      // descending
  vals[i] = i;
}
}  // namespace jdet

JDET_API size_t jdet_argsort_desc_workspace_bytes(int n) {
  if (n <= 0) return 256;
  return jdet_align_up((size_t)n * 4, 256) * 3 + (size_t)n * 16 + (1u << 20);
}

JDET_API int jdet_argsort_desc(const float* scores, int n, int* order, void* workspace, size_t workspace_bytes,
                               void* stream) {
  using namespace jdet;
  if (n < 0) return JDET_ERR_BAD_ARG;
  if (n == 0) return 0;
  if (!scores || !order || !workspace) return JDET_ERR_BAD_ARG;
  if (workspace_bytes < jdet_argsort_desc_workspace_bytes(n)) return JDET_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t a = jdet_align_up((size_t)n * 4, 256);
  unsigned* k_in = (unsigned*)workspace;
  unsigned* k_out = (unsigned*)((char*)workspace + a);
  int* v_in = (int*)((char*)workspace + 2 * a);
  void* tmp = (char*)workspace + 3 * a;
  size_t tmp_bytes = (size_t)n * 16 + (1u << 20);
  score_key_kernel<<<jdet_ceil_div(n, 256), 256, 0, st>>>(scores, n, k_in, v_in);
  size_t need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, k_in, k_out, v_in, order, n, 0, 32, st);
  if (need > tmp_bytes) return JDET_ERR_WORKSPACE;
  JDET_RETURN_IF_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, order, n, 0, 32, st));
  return (int)cudaGetLastError();
}

// ---- detection record for the end-of-step all-gather ------------------------------------------------------
// After NMS every rank gathers a fixed-size record per image: rows [x, y, w, h, theta, score, label] of the kept
// detections in descending score order, zero padded to max_per_img rows, and a last row whose first column is the count
// (what multiclass_nms_rotated's re-sort + [:max_num], ops/nms_rotated.py:584-596, leaves per image).  One CTA walks the
// score order — already on the device from the NMS call — in blocks of 1024, compacts the kept boxes with a ballot scan
// and writes the rows straight into the caller's (persistent) send buffer: no zeros(), argsort or index kernels in front
// of the collective.
namespace jdet {
__global__ void __launch_bounds__(1024) pack_detections_kernel(const float* __restrict__ dets, int box_length,
                                                                const float* __restrict__ scores, const int* __restrict__ order,
                                                                const unsigned char* __restrict__ keep, int n, int max_out,
                                                                float label_lo, float label_hi, float* __restrict__ rec) {
  // label_hi > label_lo: only boxes whose label (column 5) lies in [label_lo, label_hi) — one image of a batch whose labels
  // were offset per image for a single NMS call — and the record carries label - label_lo
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int base = 0;
  for (int i0 = 0; i0 < n && base < max_out; i0 += 1024) {
    const int i = i0 + tid;
    const int idx = i < n ? order[i] : 0;
    bool k = i < n && keep[idx] != 0;
    if (k && label_hi > label_lo) { const float l = dets[(size_t)idx * box_length + 5]; k = l >= label_lo && l < label_hi; }
    const unsigned bal = __ballot_sync(0xffffffffu, k);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      const int v = s_warp[lane];
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
      s_warp[lane] = incl - v;
      if (lane == 31) s_total = incl;
    }
    __syncthreads();
    const int pos = base + s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
    if (k && pos < max_out) {
      const float* d = dets + (size_t)idx * box_length;
      float* r = rec + (size_t)pos * 7;
      r[0] = d[0]; r[1] = d[1]; r[2] = d[2]; r[3] = d[3]; r[4] = d[4];
      r[5] = scores[idx];
      r[6] = box_length == 6 ? d[5] - (label_hi > label_lo ? label_lo : 0.f) : 0.f;
    }
    base += s_total;
    __syncthreads();
  }
  const int count = min(base, max_out);
  for (int r = count * 7 + tid; r < max_out * 7; r += 1024) rec[r] = 0.f;
  if (tid < 7) rec[(size_t)max_out * 7 + tid] = tid == 0 ? (float)count : 0.f;
}
}  // namespace jdet

// dets (n, box_length) with the label in column 5 when box_length == 6; scores (n,); order (n,) = jdet_argsort_desc(scores);
// keep (n,) bytes from jdet_nms_rotated; record (max_per_img + 1, 7) fp32, fully written.
JDET_API int jdet_pack_detections(const float* dets, int n, int box_length, const float* scores, const int* order,
                                  const unsigned char* keep, int max_per_img, float* record, void* stream) {
  if (n < 0 || max_per_img < 0 || (box_length != 5 && box_length != 6) || !record) return JDET_ERR_BAD_ARG;
  if (n > 0 && (!dets || !scores || !order || !keep)) return JDET_ERR_BAD_ARG;
  jdet::pack_detections_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(dets, box_length, scores, order, keep, n, max_per_img, 0.f, 0.f, record);
  return (int)cudaGetLastError();
}

// the same for ONE image of a batch that went through a single NMS call with per-image label offsets: only boxes with
// label_lo <= label < label_hi (box_length 6), written with label - label_lo
JDET_API int jdet_pack_detections_range(const float* dets, int n, const float* scores, const int* order, const unsigned char* keep,
                                        float label_lo, float label_hi, int max_per_img, float* record, void* stream) {
  if (n < 0 || max_per_img < 0 || !record || !(label_hi > label_lo)) return JDET_ERR_BAD_ARG;
  if (n > 0 && (!dets || !scores || !order || !keep)) return JDET_ERR_BAD_ARG;
  jdet::pack_detections_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(dets, 6, scores, order, keep, n, max_per_img, label_lo, label_hi, record);
  return (int)cudaGetLastError();
}
