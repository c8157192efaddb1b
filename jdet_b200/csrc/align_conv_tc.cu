// align_conv_tc.cu — S2ANet AlignConv as ONE fused tcgen05 implicit GEMM for sm_100a.
//
// Replaces AlignConv.execute (/root/reference/python/jdet/models/roi_heads/s2anet_head.py:715-723):
//   get_offset (:677-713, ~25 elementwise kernels per image) -> deformable_im2col
//   (ops/dcn_v1.py:131-184, a 1.2 GB columns tensor at level 0) -> jt.matmul / cuBLAS SGEMM
//   (ops/dcn_v1.py:446-447) -> ReLU.
//
// B200 design (no offsets tensor, no columns tensor):
//   D[pixel, co] = sum_k A[pixel, k] * Wt[co, k],  k = tap*C + c  (tap-major K; weights re-packed once)
//   M tile = 128 output pixels, N = Co (<= 256, the whole output-channel extent), K block = 16,
//   4 smem stages of 48 KB (the A producer is L2-latency-bound: depth, not width, hides it).
//   A operand  : produced in-kernel.  8 producer warps derive the 9 sampling points of each pixel
//                straight from its refined anchor (same fp32 op order as get_offset + im2col), gather
//                the 4 bilinear corners from a channel-last copy of x with 16-B loads (4 lanes = 16
//                channels = one 64-B row), interpolate in fp32, split each value into two TF32 terms
//                (hi = top 19 bits, lo = exact remainder) and write both into shared memory directly in
//                the UMMA canonical K-major SWIZZLE_64B layout.
//   B operand  : weights pre-split into hi/lo and pre-swizzled in global memory, one contiguous
//                Co x 64 B block per K block -> a single cp.async.bulk (TMA engine, UBLKCP) per term.
//   MMA        : one elected thread issues tcgen05.mma.kind::tf32 128 x Co x 8; 3 products per K step
//                (hi*hi + hi*lo + lo*hi  == "3xTF32", ~2^-21 relative error, fp32-class like the
//                reference's SGEMM); accumulators live in TMEM, double-buffered (2 x Co columns) so the
//                epilogue of tile i overlaps the main loop of tile i+1.
//   Epilogue   : 4 warps, tcgen05.ld 32x32b -> ReLU -> NCHW stores (a warp writes 128 contiguous bytes
//                per output channel).
//   Persistent : grid = min(#tiles, 148); warp roles: 0-3 epilogue, 4-11 A producers, 12 MMA + TMEM
//                alloc, then the B loader.  Barriers are mbarriers (full/empty per smem stage, full/empty per
//                TMEM accumulator stage).
//
// Layout in HBM: x (N,C,H,W) fp32 -> channel-last scratch (N,H,W,C); anchors (N,H,W,5); weight
// (Co,C,3,3) -> scratch [9*C/16][Co][16] hi and lo (swizzled); out (N,Co,H,W).
#include <cstdlib>
#include "common.cuh"

namespace jdet {

void launch_nchw_to_nhwc(const float* in, float* out, int B, int C, int HW, cudaStream_t st);   // relayout.cu

namespace tc {

constexpr int kBlockK = 16;                   // fp32 elements per K block = one 64-B swizzle row
constexpr int kRowBytes = kBlockK * 4;        // 64
constexpr int kQuads = kBlockK / 4;           // lanes (channel quads) per pixel
constexpr int kPixPerStep = 32 / kQuads;      // pixels a warp covers per step
constexpr int kProdWarps = 16;                // A-producer warps
constexpr int kPix = 128 / (kProdWarps * kPixPerStep);  // pixels per lane per stage
constexpr int kMmaWarp = 4 + kProdWarps, kLoadWarp = 5 + kProdWarps;
constexpr int kThreads = (6 + kProdWarps) * 32;
#ifndef JDET_AC_A_STAGES
#define JDET_AC_A_STAGES 4
#endif
#ifndef JDET_AC_B_STAGES
#define JDET_AC_B_STAGES 4
#endif
constexpr int kBlockM = 128;
constexpr int kABytes = kBlockM * kRowBytes;   // 8 KB per term

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the leader issues M = 256 MMAs over both CTAs' A tiles and B halves
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {     // the barrier at the same offset in CTA `cta`
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope (remote arrivals)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {                           // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, SWIZZLE_64B (64-B rows), 8-row groups 512 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * kRowBytes) >> 4) << 32;   // stride byte offset: one 8-row swizzle atom
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
  return d;
}
// 16-B chunk j of row r lives at chunk (j ^ ((r >> 1) & 3))  (Swizzle<2,4,3> on byte addresses)
__device__ __host__ __forceinline__ int swz_chunk(int r, int j) { return j ^ ((r >> 1) & 3); }

constexpr int kMaxLevels = 8;
// One launch covers every FPN level of a head (S2ANet: 5): tiles are numbered level after level, so the 4 / 16 / 64 tiles
// of the coarse levels fill the tail of the last wave instead of costing a launch (and a mostly empty wave) each.
struct Level {
  const float* x_nhwc;      // (N, H, W, C)
  const float* anchors;     // (N, H, W, 5)
  float* out;               // (N, Co, H, W)
  int H, W;
  float stride;
  int tile_begin;           // first tile of this level
};
struct Params {
  Level lv[kMaxLevels];
  int nlevels;
  const float* b_hi;        // [K/16][Co][16] swizzled (shared by the levels: one weight)
  const float* b_lo;
  int N, C, Co;
  int num_tiles;
};
__device__ __forceinline__ int level_of(const Params& p, int tile) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < kMaxLevels; i++)
    if (i < p.nlevels && tile >= p.lv[i].tile_begin) l = i;
  return l;
}

// weight (Co, C, 3, 3) -> hi/lo [kb][co][16] with kb = tap*(C/16) + c/16, 16-B chunks swizzled like the smem rows
__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ w, int Co, int C, float* __restrict__ hi,
                                                           float* __restrict__ lo) {
  const long long total = (long long)Co * C * 9;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int t = (int)(i % 9);
  const int c = (int)((i / 9) % C);
  const int co = (int)(i / (9LL * C));
  const float v = w[i];
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  const float l = v - h;                                      // exact
  const int kb = t * (C / kBlockK) + c / kBlockK, e = c % kBlockK;
  const int chunk = swz_chunk(co, e >> 2);
  const size_t dst = ((size_t)kb * Co + co) * kBlockK + chunk * 4 + (e & 3);
  hi[dst] = h;
  lo[dst] = l;
}

// NCTA = 1: one CTA per tile (M = 128).  NCTA = 2: a CTA pair per two tiles — each CTA produces its own 128 A rows and loads
// HALF of the weights of a K block; the leader's MMAs (cta_group::2, M = 256) read both CTAs' shared memory, so the weight
// stream into each SM is halved (at level 0 the kernel moves 8.3 GB from L2 into the SMs, 4.8 GB of it weights: l1tex 79 %).
template <int NCTA>
__global__ void __launch_bounds__(kThreads, 1) align_conv_tc_kernel(const __grid_constant__ Params p) {
  // separate rings for the two operands: an A stage is 16 KB, a B stage 2 * Co * 64 B (32 KB at Co = 256, per CTA half of it in a
  // pair) — the gathers behind A are what needs depth, and 4 whole 48 KB stages were all that fitted
  constexpr int kAS = NCTA == 2 ? 6 : JDET_AC_A_STAGES, kBS = NCTA == 2 ? 6 : JDET_AC_B_STAGES;
  const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
  const int pair_id = (int)blockIdx.x / NCTA, npairs = (int)gridDim.x / NCTA;
  const int num_pairs = (p.num_tiles + NCTA - 1) / NCTA;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.Co * kRowBytes / NCTA;         // this CTA's share of a K block's weights (per term)
  unsigned char* b_ring = smem + (size_t)kAS * 2 * kABytes;
  auto a_hi = [&](int s) { return smem + (size_t)s * 2 * kABytes; };
  auto a_lo = [&](int s) { return smem + (size_t)s * 2 * kABytes + kABytes; };
  auto b_hi = [&](int s) { return b_ring + (size_t)s * 2 * b_bytes; };
  auto b_lo = [&](int s) { return b_ring + (size_t)s * 2 * b_bytes + b_bytes; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)kBS * 2 * b_bytes);
  uint64_t* afull = bars;                        // [kAS]  the producers' rows are in place
  uint64_t* aempty = bars + kAS;                 // [kAS]  the MMAs reading the stage have retired
  uint64_t* bfull = bars + 2 * kAS;              // [kBS]  the weight block has landed
  uint64_t* bempty = bars + 2 * kAS + kBS;       // [kBS]
  uint64_t* tfull = bars + 2 * kAS + 2 * kBS;    // [2]
  uint64_t* tempty = tfull + 2;
  uint64_t* pfull = tfull + 4;                   // [kAS] (pair leader only): the peer CTA's K block is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pfull + kAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblocks = p.C / kBlockK;
  const int num_kb = 9 * cblocks;

  if (warp == kLoadWarp && lane == 0) {
    for (int s = 0; s < kAS; s++) { mbar_init(&afull[s], kProdWarps); mbar_init(&aempty[s], 1); mbar_init(&pfull[s], 1); }
    for (int s = 0; s < kBS; s++) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4 * NCTA); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    if (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (NCTA == 2) cluster_sync_all();                 // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 4 + kProdWarps) {
    // =============================== A producers ==============================================
    const int pw = warp - 4, sp = lane / kQuads, q = lane % kQuads;
    uint32_t stage = 0, phase = 0;
    for (int pr = pair_id; pr < num_pairs; pr += npairs) {
      const int tile_raw = pr * NCTA + (int)cta_rank;
      const bool dummy = tile_raw >= p.num_tiles;      // the odd tile out of a pair: zeros in, nothing out
      const int tile = dummy ? p.num_tiles - 1 : tile_raw;
      const Level& L = p.lv[level_of(p, tile)];
      const int HW = L.H * L.W;
      const long long P = dummy ? 0 : (long long)p.N * HW;
      const int ltile = tile - L.tile_begin;
      // per-pixel anchor geometry (s2anet_head.py:689-698), 8 pixels per lane
      float gx[kPix], gy[kPix], gdw[kPix], gdh[kPix], gc[kPix], gs[kPix], fh[kPix], fw[kPix];
      long long pbase[kPix];          // n * HW (pixel index of the image origin), -1 => row beyond the tensor
#pragma unroll
      for (int it = 0; it < kPix; it++) {
        const int r = pw * (kPixPerStep * kPix) + it * kPixPerStep + sp;
        const long long pix = (long long)ltile * kBlockM + r;
        if (pix < P) {
          const int n = (int)(pix / HW), hw = (int)(pix - (long long)n * HW);
          const int h = hw / L.W, w = hw - h * L.W;
          const float* a = L.anchors + pix * 5;
          gx[it] = __fdiv_rn(__ldg(a + 0), L.stride);
          gy[it] = __fdiv_rn(__ldg(a + 1), L.stride);
          gdw[it] = __fdiv_rn(__fdiv_rn(__ldg(a + 2), L.stride), 3.f);
          gdh[it] = __fdiv_rn(__fdiv_rn(__ldg(a + 3), L.stride), 3.f);
          const float ang = __ldg(a + 4);
          gc[it] = cosf(ang);
          gs[it] = sinf(ang);
          fh[it] = (float)h;
          fw[it] = (float)w;
          pbase[it] = (long long)n * HW;
        } else {
          gx[it] = gy[it] = gdw[it] = gdh[it] = gc[it] = gs[it] = fh[it] = fw[it] = 0.f;
          pbase[it] = -1;
        }
      }
      // tap geometry once per tap, reused by the C/16 channel blocks
      int o00[kPix], o01[kPix], o10[kPix], o11[kPix];
      float w1[kPix], w2[kPix], w3[kPix], w4[kPix];
      auto tap_geometry = [&](int t) {
        const float xx = (float)(t % 3 - 1), yy = (float)(t / 3 - 1);
#pragma unroll
        for (int it = 0; it < kPix; it++) {
          const float x = __fmul_rn(gdw[it], xx), y = __fmul_rn(gdh[it], yy);
          const float xr = __fsub_rn(__fmul_rn(gc[it], x), __fmul_rn(gs[it], y));
          const float yr = __fadd_rn(__fmul_rn(gs[it], x), __fmul_rn(gc[it], y));
          const float xa = __fadd_rn(xr, gx[it]), ya = __fadd_rn(yr, gy[it]);
          const float cw = __fadd_rn(fw[it], xx), chh = __fadd_rn(fh[it], yy);     // regular conv location
          const float w_im = __fadd_rn(cw, __fsub_rn(xa, cw));                     // dcn_v1.py:168-169
          const float h_im = __fadd_rn(chh, __fsub_rn(ya, chh));
          o00[it] = o01[it] = o10[it] = o11[it] = -1;
          w1[it] = w2[it] = w3[it] = w4[it] = 0.f;
          if (pbase[it] >= 0 && h_im > -1.f && w_im > -1.f && h_im < (float)L.H && w_im < (float)L.W) {
            const int hl = (int)floorf(h_im), wl = (int)floorf(w_im);
            const int hh = hl + 1, wh = wl + 1;
            const float lh = h_im - (float)hl, lw = w_im - (float)wl;
            const float uh = 1.f - lh, uw = 1.f - lw;
            w1[it] = uh * uw; w2[it] = uh * lw; w3[it] = lh * uw; w4[it] = lh * lw;
            if (hl >= 0 && wl >= 0) o00[it] = hl * L.W + wl;
            if (hl >= 0 && wh <= L.W - 1) o01[it] = hl * L.W + wh;
            if (hh <= L.H - 1 && wl >= 0) o10[it] = hh * L.W + wl;
            if (hh <= L.H - 1 && wh <= L.W - 1) o11[it] = hh * L.W + wh;
          }
        }
      };
        // The four corner loads of channel block cb + 1 are issued BEFORE the wait for its stage: a stage's gathers could
        // otherwise only start once the MMAs released it, which ties the loads in flight to the ring depth (4 stages of
        // 48 KB is all that fits) — one more K block per lane rides in registers instead.
        float4 ca[kPix], cb4[kPix], cc[kPix], cd[kPix];
        auto gather = [&](int cbk, float4 (&a)[kPix], float4 (&b)[kPix], float4 (&c)[kPix], float4 (&d)[kPix]) {
#pragma unroll
          for (int it = 0; it < kPix; it++) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            a[it] = z; b[it] = z; c[it] = z; d[it] = z;
            if (pbase[it] >= 0) {
              const float* base = L.x_nhwc + (size_t)cbk * kBlockK + q * 4;
              const size_t pb = (size_t)pbase[it];
              if (o00[it] >= 0) a[it] = __ldg(reinterpret_cast<const float4*>(base + (pb + o00[it]) * p.C));
              if (o01[it] >= 0) b[it] = __ldg(reinterpret_cast<const float4*>(base + (pb + o01[it]) * p.C));
              if (o10[it] >= 0) c[it] = __ldg(reinterpret_cast<const float4*>(base + (pb + o10[it]) * p.C));
              if (o11[it] >= 0) d[it] = __ldg(reinterpret_cast<const float4*>(base + (pb + o11[it]) * p.C));
            }
          }
        };
      // (K order: tap outer, channel block inner.  The other order — the 9 taps of one 16-channel block back to back, for L1
      //  reuse between the taps of neighbouring anchors — measured slower, 1.25-1.40 vs 1.03 ms: the tap geometry then sits
      //  in front of every K block's gathers.)
      for (int t = 0; t < 9; t++) {
        tap_geometry(t);
        gather(0, ca, cb4, cc, cd);
        for (int cb = 0; cb < cblocks; cb++) {
          float4 v[kPix];
#pragma unroll
          for (int it = 0; it < kPix; it++) {
            const float4 a = ca[it], b = cb4[it], c = cc[it], d = cd[it];
            v[it].x = w1[it] * a.x + w2[it] * b.x + w3[it] * c.x + w4[it] * d.x;
            v[it].y = w1[it] * a.y + w2[it] * b.y + w3[it] * c.y + w4[it] * d.y;
            v[it].z = w1[it] * a.z + w2[it] * b.z + w3[it] * c.z + w4[it] * d.z;
            v[it].w = w1[it] * a.w + w2[it] * b.w + w3[it] * c.w + w4[it] * d.w;
          }
          if (cb + 1 < cblocks) gather(cb + 1, ca, cb4, cc, cd);      // in flight across the wait below
          mbar_wait(&aempty[stage], phase ^ 1);
          unsigned char* ah = a_hi(stage);
          unsigned char* al = a_lo(stage);
#pragma unroll
          for (int it = 0; it < kPix; it++) {
            const int r = pw * (kPixPerStep * kPix) + it * kPixPerStep + sp;
            const uint32_t off = (uint32_t)r * (uint32_t)kRowBytes + (uint32_t)(swz_chunk(r, q) << 4);
            float4 h4, l4;
            h4.x = __uint_as_float(__float_as_uint(v[it].x) & 0xffffe000u); l4.x = v[it].x - h4.x;
            h4.y = __uint_as_float(__float_as_uint(v[it].y) & 0xffffe000u); l4.y = v[it].y - h4.y;
            h4.z = __uint_as_float(__float_as_uint(v[it].z) & 0xffffe000u); l4.z = v[it].z - h4.z;
            h4.w = __uint_as_float(__float_as_uint(v[it].w) & 0xffffe000u); l4.w = v[it].w - h4.w;
            *reinterpret_cast<float4*>(ah + off) = h4;
            *reinterpret_cast<float4*>(al + off) = l4;
          }
          fence_proxy_async();             // generic-proxy stores -> visible to the tensor-core (async) proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(&afull[stage]);
          if (++stage == kAS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kLoadWarp) {
    // =============================== B loader ==================================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const size_t my_half = (size_t)cta_rank * (p.Co / NCTA) * kBlockK;     // this CTA's output channels of a K block
      for (int pr = pair_id; pr < num_pairs; pr += npairs) {
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&bempty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bfull[stage], 2u * (uint32_t)b_bytes);
          bulk_g2s(b_hi(stage), p.b_hi + (size_t)kb * p.Co * kBlockK + my_half, (uint32_t)b_bytes, &bfull[stage]);
          bulk_g2s(b_lo(stage), p.b_lo + (size_t)kb * p.Co * kBlockK + my_half, (uint32_t)b_bytes, &bfull[stage]);
          if (++stage == kBS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer =================================================
    // instruction descriptor: D=F32, A=B=TF32, K-major both, N = Co, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Co >> 3) << 17) | ((uint32_t)((kBlockM * NCTA) >> 4) << 24);
    uint32_t stage = 0, phase = 0, bstage = 0, bphase = 0, acc = 0, acc_phase = 0;
    if (NCTA == 2 && cta_rank != 0) {
      // the peer's MMA warp only relays: "my stage is complete" (A rows produced, weight half landed) -> the leader
      for (int pr = pair_id; pr < num_pairs; pr += npairs) {
        for (int kb = 0; kb < num_kb; kb++) {
          mbar_wait(&afull[stage], phase);
          mbar_wait(&bfull[bstage], bphase);
          if (lane == 0) { fence_proxy_async(); mbar_arrive_remote(&pfull[stage], 0u); }
          __syncwarp();
          if (++stage == kAS) { stage = 0; phase ^= 1; }
          if (++bstage == kBS) { bstage = 0; bphase ^= 1; }
        }
      }
    } else
    for (int pr = pair_id; pr < num_pairs; pr += npairs) {
      if (NCTA == 2) mbar_wait_cluster(&tempty[acc], acc_phase ^ 1); else mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Co;
      for (int kb = 0; kb < num_kb; kb++) {
        mbar_wait(&bfull[bstage], bphase);
        mbar_wait(&afull[stage], phase);
        if (NCTA == 2) mbar_wait_cluster(&pfull[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t ah = make_desc(smem_u32(a_hi(stage))), al = make_desc(smem_u32(a_lo(stage)));
          const uint64_t bh = make_desc(smem_u32(b_hi(bstage))), bl = make_desc(smem_u32(b_lo(bstage)));
#pragma unroll
          for (int kk = 0; kk < kBlockK / 8; kk++) {
            const uint64_t adv = (uint64_t)(kk * 2);     // 8 tf32 = 32 B = 2 x 16 B along K inside the swizzle row
            if (NCTA == 2) {
              tc_mma_tf32_2(d_tmem, al + adv, bh + adv, idesc, (kb | kk) ? 1u : 0u);
              tc_mma_tf32_2(d_tmem, ah + adv, bl + adv, idesc, 1u);
              tc_mma_tf32_2(d_tmem, ah + adv, bh + adv, idesc, 1u);
            } else {
              tc_mma_tf32(d_tmem, al + adv, bh + adv, idesc, (kb | kk) ? 1u : 0u);   // small terms first
              tc_mma_tf32(d_tmem, ah + adv, bl + adv, idesc, 1u);
              tc_mma_tf32(d_tmem, ah + adv, bh + adv, idesc, 1u);
            }
          }
          if (NCTA == 2) {
            tc_commit2(&aempty[stage]);                     // frees the stages in both CTAs
            tc_commit2(&bempty[bstage]);
            if (kb == num_kb - 1) tc_commit2(&tfull[acc]);
          } else {
            tc_commit(&aempty[stage]);                      // frees the smem stages when these MMAs retire
            tc_commit(&bempty[bstage]);
            if (kb == num_kb - 1) tc_commit(&tfull[acc]);   // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == kAS) { stage = 0; phase ^= 1; }
        if (++bstage == kBS) { bstage = 0; bphase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =============================== epilogue (warps 0-3) ======================================
    uint32_t acc = 0, acc_phase = 0;
    for (int pr = pair_id; pr < num_pairs; pr += npairs) {
      const int tile_raw = pr * NCTA + (int)cta_rank;
      const bool dummy = tile_raw >= p.num_tiles;
      const int tile = dummy ? p.num_tiles - 1 : tile_raw;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const Level& L = p.lv[level_of(p, tile)];
      const int HW = L.H * L.W;
      const long long P = dummy ? 0 : (long long)p.N * HW;
      const int r = warp * 32 + lane;
      const long long pix = (long long)(tile - L.tile_begin) * kBlockM + r;
      const bool ok = pix < P;
      const int n = ok ? (int)(pix / HW) : 0;
      const int hw = ok ? (int)(pix - (long long)n * HW) : 0;
      float* obase = L.out + (size_t)n * p.Co * HW + hw;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * (uint32_t)p.Co;
      for (int c0 = 0; c0 < p.Co; c0 += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ok) {
#pragma unroll
          for (int j = 0; j < 32; j++) st_stream(obase + (size_t)(c0 + j) * HW, fmaxf(__uint_as_float(v[j]), 0.f));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (NCTA == 2) mbar_arrive_remote(&tempty[acc], 0u); else mbar_arrive(&tempty[acc]); }   // (the leader waits for both CTAs' epilogues)
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (NCTA == 2) cluster_sync_all();                 // nobody leaves while the pair still reads its shared memory / barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace tc

bool align_conv_tc_supported(int C, int Co) { return C % tc::kBlockK == 0 && Co % 32 == 0 && Co >= 32 && Co <= 256; }

size_t align_conv_tc_workspace_bytes_multi(int nlevels, int N, int C, const int* Hs, const int* Ws, int Co) {
  size_t total = 2 * jdet_align_up((size_t)Co * C * 9 * 4, 1024);
  for (int l = 0; l < nlevels; l++) total += jdet_align_up((size_t)N * C * Hs[l] * Ws[l] * 4, 1024);
  return total;
}
size_t align_conv_tc_workspace_bytes(int N, int C, int H, int W, int Co) { return align_conv_tc_workspace_bytes_multi(1, N, C, &H, &W, Co); }

// every level of one head in ONE persistent launch (the levels share the weight): re-layout per level, one weight split
int align_conv_tc_launch_multi(const float* const* xs, const float* const* anchors, const float* weight, int nlevels, int N, int C,
                               const int* Hs, const int* Ws, int Co, const float* strides, float* const* outs, void* workspace,
                               cudaStream_t st, bool x_channels_last) {
  using namespace tc;
  if (nlevels < 1 || nlevels > kMaxLevels) return JDET_ERR_UNSUPPORTED;
  char* wsp = (char*)workspace;
  float* b_hi = (float*)wsp;   wsp += jdet_align_up((size_t)Co * C * 9 * 4, 1024);
  float* b_lo = (float*)wsp;   wsp += jdet_align_up((size_t)Co * C * 9 * 4, 1024);
  Params p{};
  p.nlevels = nlevels; p.b_hi = b_hi; p.b_lo = b_lo; p.N = N; p.C = C; p.Co = Co;
  long long tiles = 0;
  for (int l = 0; l < nlevels; l++) {
    const float* x_nhwc = xs[l];                    // a map that is already (N,H,W,C) in memory is sampled in place
    if (!x_channels_last) {
      float* scratch = (float*)wsp;   wsp += jdet_align_up((size_t)N * C * Hs[l] * Ws[l] * 4, 1024);
      launch_nchw_to_nhwc(xs[l], scratch, N, C, Hs[l] * Ws[l], st);
      x_nhwc = scratch;
    }
    p.lv[l] = Level{x_nhwc, anchors[l], outs[l], Hs[l], Ws[l], strides[l], (int)tiles};
    tiles += ((long long)N * Hs[l] * Ws[l] + kBlockM - 1) / kBlockM;
    if (tiles > 0x7fffffffLL) return JDET_ERR_UNSUPPORTED;
  }
  p.num_tiles = (int)tiles;
  if (p.num_tiles == 0) return 0;
  const long long wt = (long long)Co * C * 9;
  weight_prep_kernel<<<(int)((wt + 255) / 256), 256, 0, st>>>(weight, Co, C, b_hi, b_lo);
  const int sms = num_sms();
  static const bool pair = getenv("JDET_ALIGN_CONV_2CTA") != nullptr;      // opt-in: CTA pairs (cta_group::2)
  if (pair && Co % 64 == 0 && p.num_tiles >= 2) {
    const size_t smem = 1024 + (size_t)6 * (2 * kABytes + (size_t)Co * kRowBytes) + 1024;
    cudaError_t e = cudaFuncSetAttribute(align_conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int pairs = (p.num_tiles + 1) / 2;
    const int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, align_conv_tc_kernel<2>, p);
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
  }
  const size_t smem = 1024 + (size_t)JDET_AC_A_STAGES * 2 * kABytes + (size_t)JDET_AC_B_STAGES * 2 * (size_t)Co * kRowBytes + 512;
  cudaError_t e = cudaFuncSetAttribute(align_conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  align_conv_tc_kernel<1><<<grid, kThreads, smem, st>>>(p);
  return (int)cudaGetLastError();
}

int align_conv_tc_launch(const float* x, const float* anchors, const float* weight, int N, int C, int H, int W, int Co,
                         float stride, float* out, void* workspace, cudaStream_t st) {
  return align_conv_tc_launch_multi(&x, &anchors, weight, 1, N, C, &H, &W, Co, &stride, &out, workspace, st, false);
}

}  // namespace jdet
