// roi_gather_tma.cuh — the staged RoIAlign forward: per-RoI pixel staging by TMA gather4, gather from shared memory.
//
// Replaces the sampling loop of ROIAlignRotatedForward (ops/roi_align_rotated_v1.py:70-147, ops/roi_align_rotated.py:60-127)
// for dense RoI sets, on a channel-last (B*H*W, C) view of the feature map.
//
// Why this shape (measured: profiles/r02_tma_probe_*.txt, tools/tma_probe.cu).  Under load an L2 hit costs ~1 us on
// this part, so the fill rate of an SM is simply (bytes in flight) / 1 us.  The round-1 gather kept its bytes in flight
// in registers (24 warps x 8 x 512 B = 98 KB per SM at best) and ran at 11 TB/s, latency-bound with the load path,
// the issue slots and HBM all about half busy.  Here the bytes in flight live in shared memory instead:
//
//   * the prologue turns a RoI into CHUNKS — the whole RoI when its distinct pixels fit half the ring (the common
//     case: 60 % of the DOTA-shaped RoIs touch < 100 distinct pixels although their bins name ~250 taps), else one
//     bin row, else one bin — and de-duplicates the pixels of a chunk with a bitmap over its bounding box: slot =
//     rank of the pixel's bit.  A tap-table entry is (slot * slab bytes, weight); a chunk's pixel list is the row
//     index list of its TMA gathers;
//   * producer warps issue cp.async.bulk.tensor.2d tile::gather4 (SASS UTMALDG.2D.GATHER4): four arbitrary pixels x
//     one 128-channel slab per instruction, straight into a 128 KB ring, completion counted on a per-chunk mbarrier.
//     One warp issues ~1 gather per 200 cycles however deep its queue, so 8 warps issue (the probe: 14 TB/s);
//   * consumer warps take bins from a shared counter, wait for the bin's chunk, and do nothing but LDS.128 + FFMA2:
//     no global load, no long-scoreboard stall.  Whole-RoI chunks use two bins per warp (a half-warp per bin, 8
//     channels per lane), row / bin chunks one bin per warp so that the bins resident in the ring keep all 16 busy;
//   * a chunk's ring space is released when all 16 consumer warps have moved past it (they arrive on its `empty`
//     barrier as they take a bin of a later chunk); a scheduler warp claims work items and streams the table records
//     two items ahead; the output slab leaves as one bulk store that drains under the next item.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace jdet {
namespace g4 {

constexpr int kConsumerWarps = 16;
constexpr int kProducerWarps = 8;
constexpr int kThreads = 32 * (kConsumerWarps + kProducerWarps + 1);   // + the scheduler warp
constexpr int kNB = 16;          // chunk barrier slots (chunks in flight)
constexpr int kTB = 4;           // table record buffers (RoIs in flight)
constexpr int kMaxSlabs = 32;    // channel slabs of one RoI (C <= 32 * slab)
constexpr int kBitmapWords = 256;   // bounding box of a de-duplicated chunk: <= 8192 pixels

// ---- table record (16-B granular, offsets fixed by nbins and fstride) --------------------------------------------
//   int4 {batch, float bits of the divisor, fstride, FPN level}
//   int4 {bins per chunk, chunks, pixels in all lists, record bytes}
//   int  cnt[nbins up 4]          merged taps of a bin
//   int2 fin[nbins][fstride]      (slot * slab bytes, weight); padding entries repeat the bin's last slot with weight 0
//   int  cnpx[nbins up 4]         pixels of chunk c (a multiple of 4: a list is padded with its last pixel)
//   int  pix[...]                 gather rows (batch * H * W + y * W + x) of chunk 0, 1, ... back to back
struct RecLayout { int nb4, cnt_off, fin_off, cnpx_off, pix_off, max_bytes; };
__host__ __device__ inline RecLayout rec_layout(int nbins, int fstride, int tpb) {
  RecLayout r;
  r.nb4 = (nbins + 3) & ~3;
  r.cnt_off = 32;
  r.fin_off = r.cnt_off + r.nb4 * 4;
  r.cnpx_off = r.fin_off + nbins * fstride * 8;
  r.pix_off = r.cnpx_off + r.nb4 * 4;
  r.max_bytes = (r.pix_off + nbins * (tpb + 4) * 4 + 15) & ~15;
  return r;
}

// ring allocation rule shared by producers and consumers: a chunk never wraps
__device__ __forceinline__ unsigned ring_place(unsigned& pos, unsigned size, unsigned ring) {
  if ((pos & (ring - 1)) + size > ring) pos = (pos | (ring - 1)) + 1;
  const unsigned start = pos;
  pos += size;
  return start;
}

// ---- prologue side: chunking + slot assignment -----------------------------------------------------------------
// On entry fin[bin][k] = ((y << 16) | x, weight) for k < cnt[bin], padding = the bin's last pixel (0 for an empty bin).
// The group routines run either on the whole CTA (whole-RoI chunk) or on one warp (a bin row per warp, rows in parallel).
struct ChunkScratch { unsigned bitmap[kBitmapWords]; unsigned short wpre[kBitmapWords]; int mn_x, mn_y, mx_x, mx_y, total, pad_[3]; };
struct Coop {
  int tid, n; bool warp;
  __device__ __forceinline__ void sync() const { if (warp) __syncwarp(); else __syncthreads(); }
};
struct EntryIndex {   // entry i of a group -> (bin offset, k) without a division when fstride is a power of two
  int fstride, shift;
  __device__ __forceinline__ int bin(int i) const { return shift >= 0 ? i >> shift : i / fstride; }
};
__device__ __forceinline__ EntryIndex entry_index(int f) {
  EntryIndex e;
  e.fstride = f;
  e.shift = (f & (f - 1)) == 0 ? __ffs(f) - 1 : -1;
  return e;
}

// bounding box of the taps of bins [bin0, bin0 + nb) -> S.mn_*/mx_*; false if it does not fit the bitmap
__device__ inline bool group_bbox(const Coop& G, const int2* fin, const int* cnt, int fstride, int bin0, int nb, ChunkScratch& S) {
  const EntryIndex E = entry_index(fstride);
  const int n = nb * fstride;
  G.sync();
  if (G.tid == 0) { S.mn_x = S.mn_y = 0x7fffffff; S.mx_x = S.mx_y = -1; }
  G.sync();
  int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -1, mxy = -1;
  for (int e = G.tid; e < n; e += G.n) {
    const int bo = E.bin(e), k = e - bo * fstride;
    if (k < cnt[bin0 + bo]) {
      const int p = fin[bin0 * fstride + e].x, y = p >> 16, x = p & 0xffff;
      mnx = min(mnx, x); mxx = max(mxx, x); mny = min(mny, y); mxy = max(mxy, y);
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, d)); mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, d));
    mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, d)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
  }
  if ((threadIdx.x & 31) == 0 && mxx >= 0) {
    if (G.warp) { S.mn_x = mnx; S.mn_y = mny; S.mx_x = mxx; S.mx_y = mxy; }
    else { atomicMin(&S.mn_x, mnx); atomicMax(&S.mx_x, mxx); atomicMin(&S.mn_y, mny); atomicMax(&S.mx_y, mxy); }
  }
  G.sync();
  if (S.mx_x < 0) return true;
  return (long long)(S.mx_x - S.mn_x + 1) * (S.mx_y - S.mn_y + 1) <= kBitmapWords * 32;
}

// bitmap + rank prefix of the group whose bounding box sits in S (as left by group_bbox): the number of distinct pixels
__device__ inline int group_mark(const Coop& G, const int2* fin, const int* cnt, int fstride, int bin0, int nb, ChunkScratch& S) {
  const EntryIndex E = entry_index(fstride);
  const int n = nb * fstride;
  if (S.mx_x < 0) return 0;
  const int x0 = S.mn_x, y0 = S.mn_y, bw = S.mx_x - x0 + 1;
  for (int i = G.tid; i < kBitmapWords; i += G.n) S.bitmap[i] = 0u;
  G.sync();
  for (int e = G.tid; e < n; e += G.n) {
    const int bo = E.bin(e), k = e - bo * fstride;
    if (k < cnt[bin0 + bo]) {
      const int p = fin[bin0 * fstride + e].x, bit = ((p >> 16) - y0) * bw + ((p & 0xffff) - x0);
      atomicOr(&S.bitmap[bit >> 5], 1u << (bit & 31));
    }
  }
  G.sync();
  if (G.tid < 32) {                                   // exclusive prefix of the word popcounts (8 words per lane)
    int c[kBitmapWords / 32], sum = 0;
#pragma unroll
    for (int j = 0; j < kBitmapWords / 32; j++) { c[j] = __popc(S.bitmap[G.tid * (kBitmapWords / 32) + j]); sum += c[j]; }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (G.tid >= d) incl += v; }
    int run = incl - sum;
#pragma unroll
    for (int j = 0; j < kBitmapWords / 32; j++) { S.wpre[G.tid * (kBitmapWords / 32) + j] = (unsigned short)run; run += c[j]; }
    if (G.tid == 31) S.total = incl;
  }
  G.sync();
  return S.total;
}

// entries -> (slot * pxbytes, weight), pixel list (padded to a multiple of 4) -> pix_out; needs group_mark's state
__device__ inline void group_commit(const Coop& G, int2* fin, const int* cnt, int fstride, int bin0, int nb, int nd, int pxbytes, int W,
                                    int row_base, int* pix_out, ChunkScratch& S) {
  const EntryIndex E = entry_index(fstride);
  const int n = nb * fstride;
  if (nd == 0) {                                      // no tap at all (RoI outside the map): an empty chunk
    for (int e = G.tid; e < n; e += G.n) fin[bin0 * fstride + e].x = 0;
    G.sync();
    return;
  }
  const int x0 = S.mn_x, y0 = S.mn_y, bw = S.mx_x - x0 + 1;
  for (int e = G.tid; e < n; e += G.n) {              // padding entries follow their bin's last pixel
    int2& f = fin[bin0 * fstride + e];
    if (cnt[bin0 + E.bin(e)] == 0) { f.x = 0; continue; }
    const int p = f.x, bit = ((p >> 16) - y0) * bw + ((p & 0xffff) - x0);
    f.x = (S.wpre[bit >> 5] + __popc(S.bitmap[bit >> 5] & ((1u << (bit & 31)) - 1u))) * pxbytes;
  }
  for (int w = G.tid; w < kBitmapWords; w += G.n) {   // pixel list in slot order
    unsigned m = S.bitmap[w];
    if (!m) continue;
    int slot = S.wpre[w];
    int yy = (w * 32) / bw, xx = (w * 32) - yy * bw;  // one division per non-empty word, then incremental
    int last = 0;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      xx += b - last; last = b;
      while (xx >= bw) { xx -= bw; yy++; }
      pix_out[slot++] = row_base + (y0 + yy) * W + x0 + xx;
    }
  }
  G.sync();
  if (G.tid < 3 && nd + G.tid < ((nd + 3) & ~3)) pix_out[nd + G.tid] = pix_out[nd - 1];
  G.sync();
}

// Tries to turn the RoI into ONE staged chunk: true (and the record is finished, *bytes = its size) when its distinct
// pixels fit `pcap`; false leaves fin untouched (the caller then emits the LSU-path record).
__device__ inline bool chunk_record(unsigned char* rec, int nbins, int fstride, int tpb, int pcap, int pxbytes, int W, int row_base,
                                    ChunkScratch& S, int* bytes) {
  const RecLayout L = rec_layout(nbins, fstride, tpb);
  int* cnt = reinterpret_cast<int*>(rec + L.cnt_off);
  int2* fin = reinterpret_cast<int2*>(rec + L.fin_off);
  int* cnpx = reinterpret_cast<int*>(rec + L.cnpx_off);
  int* pix = reinterpret_cast<int*>(rec + L.pix_off);
  const Coop B{(int)threadIdx.x, (int)blockDim.x, false};
  if (!group_bbox(B, fin, cnt, fstride, 0, nbins, S)) return false;
  const int nd = group_mark(B, fin, cnt, fstride, 0, nbins, S);
  if (nd > pcap) return false;
  group_commit(B, fin, cnt, fstride, 0, nbins, nd, pxbytes, W, row_base, pix, S);
  const int total = (nd + 3) & ~3;
  *bytes = (L.pix_off + total * 4 + 15) & ~15;
  if (threadIdx.x == 0) { cnpx[0] = total; reinterpret_cast<int4*>(rec)[1] = make_int4(nbins, 1, total, *bytes); }
  __syncthreads();
  return true;
}

// ---- gather kernel ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// packed fp32x2 FMA (SASS FFMA2): acc.{x,y} += w * v.{x,y} — two IEEE fmas per issue slot, same bits as fmaf
__device__ __forceinline__ void fma2(float& ax, float& ay, float w, float vx, float vy) {
  unsigned long long a, ww, vv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(ax), "f"(ay));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(vx), "f"(vy));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(ww), "l"(vv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(ax), "=f"(ay) : "l"(a));
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

struct GatherMaps { CUtensorMap map[kMaxLevels]; };   // (B*H*W, C) fp32 view of each level's channel-last map, box {slab, 1}

struct GatherArgs {
  unsigned long long* stats;                                            // JDET_G4_STATS builds: per-role cycle counters (tools only)
  const unsigned char* tables; const int* rec_bytes; size_t stride;   // table records, their sizes, the record pitch
  int* work_counter; float* out;
  int C, nbins, nslabs, items, fstride, tpb;
  unsigned ring;                                                        // ring bytes (a power of two)
};

#ifdef JDET_G4_STATS
#define G4_T0() const long long t0_ = clock64()
#define G4_ADD(slot) do { if (lane == 0) atomicAdd(A.stats + (slot), (unsigned long long)(clock64() - t0_)); } while (0)
#else
#define G4_T0() do {} while (0)
#define G4_ADD(slot) do {} while (0)
#endif
template <int SLAB>
__global__ void __launch_bounds__(kThreads, 1) roi_gather4_kernel(const __grid_constant__ GatherMaps M, const GatherArgs A) {
  constexpr int PXB = SLAB * 4;                         // bytes of one staged pixel
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nbins = A.nbins, S = nbins | 1;
  const int out_floats = (SLAB * S + 3) & ~3;
  unsigned char* ring = smem;
  float* s_out0 = reinterpret_cast<float*>(smem + A.ring);
  unsigned char* tbuf = reinterpret_cast<unsigned char*>(s_out0 + 2 * out_floats);
  __shared__ uint64_t tfull[kTB], tfree[kTB], cfull[kNB], cempty[kNB];
  __shared__ int s_roi[kTB], s_next[kTB][kMaxSlabs];
  __shared__ unsigned s_cstart[kProducerWarps][kNB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RecLayout RL = rec_layout(nbins, A.fstride, A.tpb);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTB; i++) { mbar_init(&tfull[i], 1); mbar_init(&tfree[i], kConsumerWarps + kProducerWarps); }
    // `full` counts every producer warp (not just the one that posts the byte count): a chunk cannot complete before all
    // producers reached it, so none of them can fall two phases behind on the `empty` barrier of its slot
    for (int i = 0; i < kNB; i++) { mbar_init(&cfull[i], kProducerWarps); mbar_init(&cempty[i], kConsumerWarps); }
    mbar_init_fence();
  }
  __syncthreads();

  if (warp == kConsumerWarps + kProducerWarps) {
    // ================= scheduler: claims items, streams their table records ====================================
    if (lane == 0) {
      const int nstaged = A.work_counter[2];             // RoIs the prologue staged; their indices follow the LSU buckets
      const int* glist = A.work_counter + 64 + (size_t)kRoiBuckets * A.items;
      for (int u = 0;; u++) {
        const int b = u % kTB;
        if (u >= kTB) mbar_wait(&tfree[b], ((u / kTB) - 1) & 1);
        const int item = u == 0 ? (int)blockIdx.x : atomicAdd(A.work_counter + 1, 1) + (int)gridDim.x;
        if (item >= nstaged) {                           // end marker: a record whose chunk count is -1
          reinterpret_cast<int4*>(tbuf + (size_t)b * A.stride)[1] = make_int4(0, -1, 0, 0);
          mbar_arrive(&tfull[b]);
          break;
        }
        const int roi = glist[item];
        s_roi[b] = roi;
        for (int sl = 0; sl < A.nslabs; sl++) s_next[b][sl] = 0;
        const uint32_t bytes = (uint32_t)A.rec_bytes[roi];
        mbar_expect_tx(&tfull[b], bytes);
        bulk_g2s(tbuf + (size_t)b * A.stride, A.tables + (size_t)roi * A.stride, bytes, &tfull[b]);
      }
    }
  } else if (warp >= kConsumerWarps) {
    // ================= producers: gather4 the chunks' pixels into the ring =====================================
    const int p = warp - kConsumerWarps;
    unsigned head = 0, tail_pos = 0;
    int seq = 0, tail_seq = 0;
    for (int t = 0;; t++) {
      const int b = t % kTB;
      { G4_T0(); mbar_wait(&tfull[b], (t / kTB) & 1); G4_ADD(0); }
      const unsigned char* rec = tbuf + (size_t)b * A.stride;
      const int4 h0 = reinterpret_cast<const int4*>(rec)[0], h1 = reinterpret_cast<const int4*>(rec)[1];
      const int nchunks = h1.y;
      if (nchunks < 0) break;
      const CUtensorMap* map = &M.map[h0.w];
      const int* cnpx = reinterpret_cast<const int*>(rec + RL.cnpx_off);
      for (int sl = 0; sl < A.nslabs; sl++) {
      const int col = sl * SLAB;
      const int* pix = reinterpret_cast<const int*>(rec + RL.pix_off);
      for (int c = 0; c < nchunks; c++) {
        const int npx = cnpx[c];
        const unsigned size = (unsigned)npx * PXB;
        unsigned pos = head;
        const unsigned start = ring_place(pos, size, A.ring);
        // the barrier slot's previous chunk and enough ring space must have been released by the consumers
        while (seq - tail_seq >= kNB || pos - tail_pos > A.ring) {
          { G4_T0(); mbar_wait(&cempty[tail_seq % kNB], (tail_seq / kNB) & 1); G4_ADD(1); }
          tail_seq++;
          tail_pos = tail_seq < seq ? s_cstart[p][tail_seq % kNB] : start;
        }
        head = pos;
        G4_T0();
        if (lane == 0) {
          s_cstart[p][seq % kNB] = start;
          uint64_t* bar = &cfull[seq % kNB];
          if (p == 0) mbar_expect_tx(bar, size); else mbar_arrive(bar);
          unsigned char* dst = ring + (start & (A.ring - 1));
          for (int j = (p - seq) & (kProducerWarps - 1); 4 * j < npx; j += kProducerWarps) {   // rotate: small chunks have < 8 gathers
            const int4 r = *reinterpret_cast<const int4*>(pix + 4 * j);
            tma_gather4(dst + (size_t)j * 4 * PXB, map, col, r.x, r.y, r.z, r.w, bar);
          }
        }
        __syncwarp();
        G4_ADD(2);
        pix += npx;
        seq++;
      }
      }
      if (lane == 0) mbar_arrive(&tfree[b]);
    }
  } else {
    // ================= consumers ================================================================================
    unsigned pos = 0;                                    // ring position after the chunks passed so far
    int seq_base = 0, sub = 0;                           // chunks / (RoI, slab) sub-items done so far
    const uint32_t ring_u32 = smem_u32(ring);
    for (int t = 0;; t++) {
      const int b = t % kTB;
      { G4_T0(); mbar_wait(&tfull[b], (t / kTB) & 1); G4_ADD(4); }
      const unsigned char* rec = tbuf + (size_t)b * A.stride;
      const int4 h0 = reinterpret_cast<const int4*>(rec)[0], h1 = reinterpret_cast<const int4*>(rec)[1];
      const int nchunks = h1.y, bpc = h1.x;
      if (nchunks < 0) break;
      const int fstride = h0.z;
      const float count = __int_as_float(h0.y);
      const int icnt = (int)count;
      const bool pow2 = (icnt & (icnt - 1)) == 0;          // x / 2^k == x * 2^-k exactly
      const float rcnt = 1.f / count;
      const int* cnt = reinterpret_cast<const int*>(rec + RL.cnt_off);
      const int2* fin = reinterpret_cast<const int2*>(rec + RL.fin_off);
      const int* cnpx = reinterpret_cast<const int*>(rec + RL.cnpx_off);
      const int r = s_roi[b];
      for (int sl = 0; sl < A.nslabs; sl++, sub++) {
      float* s_out = s_out0 + (sub & 1) * out_floats;
      int my_c = 0;                                        // chunks [0, my_c) are behind this warp
      unsigned cstart = 0;
      bool placed = false;                                 // cstart valid for chunk my_c
      constexpr bool pair = true;                          // staged RoIs are one chunk: two bins per warp
      (void)bpc;
      const int ntasks = pair ? (nbins + 1) >> 1 : nbins;
      int c = 0, cend = bpc, ready = -1;                   // chunk of the current bin, its last bin + 1, last chunk seen complete
      for (int task = warp; task < ntasks; task += kConsumerWarps) {   // static order: bins of a RoI cost about the same
        const int bin0 = pair ? 2 * task : task;
        if (!pair) while (bin0 >= cend) { c++; cend += bpc; }
        while (my_c < c) {                                 // leave the chunks before c behind
          if (!placed) ring_place(pos, (unsigned)cnpx[my_c] * PXB, A.ring);
          placed = false;
          if (lane == 0) mbar_arrive(&cempty[(seq_base + my_c) % kNB]);
          my_c++;
        }
        if (!placed) { cstart = ring_place(pos, (unsigned)cnpx[c] * PXB, A.ring); placed = true; }
        if (c > ready) {
          G4_T0(); mbar_wait(&cfull[(seq_base + c) % kNB], ((seq_base + c) / kNB) & 1); G4_ADD(5);
          ready = c;
        }
        G4_T0();
        const uint32_t cbase = ring_u32 + (cstart & (A.ring - 1));
        {
          // two bins per warp: a half-warp per bin, a lane owns channels 4q..4q+3 and SLAB/2 + 4q..
          constexpr int QL = SLAB / 8;                     // lanes per bin (16 at SLAB 128)
          static_assert(QL == 16 || QL == 8, "slab");
          const int grp = lane / QL, q = lane % QL;
          constexpr int G = 32 / QL;                       // bins per warp task... (pair tasks hand out 2 bins; G > 2 only splits lanes)
          const int bin = bin0 + (grp & 1);
          const bool valid = bin < nbins && grp < 2;
          const int n = valid ? cnt[bin] : 0;
          int nmax = max(n, __shfl_xor_sync(0xffffffffu, n, QL));
          if (G > 2) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 2 * QL));
          const int2* e = fin + (valid ? bin : 0) * fstride;
          const uint32_t lbase = cbase + q * 16;
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
          if (n > 0) {
            for (int k = 0; k < nmax; k += 2) {            // entries past n: the last slot again with weight 0
              const int4 t01 = *reinterpret_cast<const int4*>(e + k);
              const float4 v0 = lds_v4(lbase + (unsigned)t01.x), v1 = lds_v4(lbase + (unsigned)t01.x + PXB / 2);
              const float4 v2 = lds_v4(lbase + (unsigned)t01.z), v3 = lds_v4(lbase + (unsigned)t01.z + PXB / 2);
              const float w0 = __int_as_float(t01.y), w1 = __int_as_float(t01.w);
              fma2(a0.x, a0.y, w0, v0.x, v0.y); fma2(a0.z, a0.w, w0, v0.z, v0.w);
              fma2(a1.x, a1.y, w0, v1.x, v1.y); fma2(a1.z, a1.w, w0, v1.z, v1.w);
              fma2(a0.x, a0.y, w1, v2.x, v2.y); fma2(a0.z, a0.w, w1, v2.z, v2.w);
              fma2(a1.x, a1.y, w1, v3.x, v3.y); fma2(a1.z, a1.w, w1, v3.z, v3.w);
            }
          }
          if (valid) {
            const int rot = (q >> 3) & 3;
#pragma unroll
            for (int u = 0; u < 2; u++) {
              float4 a = u ? a1 : a0;
              if (pow2) { a.x *= rcnt; a.y *= rcnt; a.z *= rcnt; a.w *= rcnt; }
              else { a.x /= count; a.y /= count; a.z /= count; a.w /= count; }
              const float b0 = rot & 1 ? a.y : a.x, b1 = rot & 1 ? a.z : a.y, b2 = rot & 1 ? a.w : a.z, b3 = rot & 1 ? a.x : a.w;
              float* row = s_out + (u * (SLAB / 2) + 4 * q) * S + bin;
              row[((0 + rot) & 3) * S] = b0; row[((1 + rot) & 3) * S] = b1; row[((2 + rot) & 3) * S] = b2; row[((3 + rot) & 3) * S] = b3;
            }
          }
        }
        G4_ADD(6);
      }
      G4_T0();
      while (my_c < nchunks) {                             // item done for this warp: release what is left
        if (!placed) ring_place(pos, (unsigned)cnpx[my_c] * PXB, A.ring);
        placed = false;
        if (lane == 0) mbar_arrive(&cempty[(seq_base + my_c) % kNB]);
        my_c++;
      }
      seq_base += nchunks;
      // out[r][c0 .. c0+SLAB-1][bins] is one contiguous run of SLAB*nbins floats
      const int c0 = sl * SLAB;
      float* dst = A.out + ((size_t)r * A.C + c0) * nbins;
      const int total = SLAB * nbins;
      // Plain vector stores, not a bulk store: the TMA queue of this SM is full of the producers' gathers, and a
      // bulk store waiting behind them held the item barrier for ~4 us (clock64 attribution, tools/roi_g4_stats.cu).
      // The two s_out buffers alternate, so a thread still copying sub-item n cannot be overtaken by writes of n + 2:
      // the barrier of n + 1 lies between.
      named_bar_sync(1, 32 * kConsumerWarps);
      G4_ADD(7);
      if (S == nbins && (total & 3) == 0 && (((uintptr_t)dst) & 15) == 0) {
        const float4* src4 = reinterpret_cast<const float4*>(s_out);
        for (int i = threadIdx.x; i < total / 4; i += 32 * kConsumerWarps) {
          const float4 v = src4[i];
          st_stream_v4(dst + 4 * i, v.x, v.y, v.z, v.w);
        }
      } else {
        for (int i = threadIdx.x; i < total; i += 32 * kConsumerWarps) st_stream(dst + i, s_out[(i / nbins) * S + i % nbins]);
      }
      }
      if (lane == 0) mbar_arrive(&tfree[b]);
    }
  }
}

}  // namespace g4
}  // namespace jdet
