// idb_gemm_conv: persistent, warp-specialised tcgen05 implicit-GEMM for sm_100a.
//
//   D[M, N] = epilogue( sum_k im2col(A)[M, k] * W[N, k] )          bf16 x bf16 -> fp32 (TMEM)
//
// One CTA per SM, 576 threads:
//   warps 0-15 : epilogue   (TMEM -> registers -> bias / time-emb / GEGLU / residual -> smem staging -> TMA store)
//   warp  16   : TMA producer (im2col by signed TMA coordinates over the NHWC image; halo = OOB zero fill)
//   warp  17   : tcgen05.mma issuer (one elected lane), owns the TMEM allocation
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) and a ring of 2-4 TMEM accumulator buffers
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the mainloop of tiles i+1...
// A tile is 128 output pixels laid out as a (BW x BH x BB) rectangle of the (x, y, image) space,
// so a 3x3 tap is one 4-D TMA box shifted by (dx-1, dy-1); a stride-2 conv uses a 5-D view
// [B, H/2, 2, W/2, 2*C] of the same tensor.  A Linear is the 1x1 "image" [1, 1, M, K].
// A fused rank-r LoRA rides along as 16 extra UMMA N-columns (T = x A^T in TMEM); the epilogue warps round T to a
// bf16 operand tile and the MMA warp adds T U^T with one more K=16 UMMA -- base weights are never touched.
//
// CG = 2 (`cta_group::2`): two CTAs of a cluster (an SM pair) share one 256 x BLOCK_N tile.  Each CTA
// TMA-loads its own 128 A rows and HALF of the B tile; the leader issues tcgen05.mma.cta_group::2
// which reads both CTAs' smem and writes rows 0-127 / 128-255 of D into the two CTAs' TMEM.  Operand
// bytes per MAC drop by (128 + N/2) / (128 + N): the big convs are L2->SM bandwidth bound otherwise.
#include <atomic>
#include <cstdlib>
#include <string>

#include "../../include/idb.h"
#include "idb_common.cuh"
#include "idb_host.h"

#ifndef IDB_EPI_PROF
#define IDB_EPI_PROF 0
#endif

namespace idb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_THREADS = (NUM_EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 512;
constexpr int EPI_BUF_BYTES = 32 * 32 * 4;                          // per-warp staging buffer: a 32 x 32 fp32 chunk (SWIZZLE_128B rows) or a bf16 chunk
constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * EPI_BUF_BYTES;    // 65,536 B
constexpr int epi_bias_bytes(int tile_n) { return NUM_EPI_WARPS * ((tile_n / 32 + 3) / 4) * 32 * 4; }   // per-warp bias of its chunks of the tile
constexpr int RES_BAR_OFFSET = 512;                                 // residual-load mbarriers [warp] inside the barrier KiB

// Division of small non-negative integers by a launch-time constant: q = (x * ceil(2^40 / d)) >> 40, exact for
// x, d < 2^20 (the host checks).  A runtime `/` costs ~25 dependent instructions; the tile loops of the three
// warp roles decode a tile index with six of them, which dominated small-K tiles.
struct FastDiv {
  unsigned long long mul;
  int d;
  __host__ void set(int dd) {
    d = dd;
    mul = ((1ull << 40) + static_cast<unsigned long long>(dd) - 1) / static_cast<unsigned long long>(dd);
  }
  __device__ __forceinline__ int div(int x) const { return static_cast<int>((static_cast<unsigned long long>(x) * mul) >> 40); }
  __device__ __forceinline__ void divmod(int x, int& q, int& r) const {
    q = div(x);
    r = x - q * d;
  }
};

struct TileCoord {
  int n_blk, ks, m_blk, tx, ty, tb;
};

struct GemmParams {
  FastDiv fd_ntn, fd_ks, fd_tx, fd_ty, fd_seg;
  CUtensorMap tmA0, tmA1, tmW, tmW1, tmL, tmU, tmOutF, tmOutB, tmRes;   // tmU: LoRA up-projection weights [N, 64] bf16
  CUtensorMap _pad_unused;   // tmOut* / tmRes: 4-D [N_out, Wo, Ho, B] maps (32-row boxes); tmW1: W box of the pair's second CTA (LoRA)
  int mode0, cpb0, c0, nkb0, nkb1;
  int tap_off_x, tap_off_y;   // IDB_A_2X2
  int Ho, Wo, B;
  int BW, BH, BB;
  int tiles_x, tiles_y;
  int n_tiles_m, n_tiles_n, k_splits, kb_per_split;
  int N, N_out;
  long long M;
  const float* bias;
  const float* rowvec;
  long long rowvec_ld;
  const float* residual;
  const float* prelu;
  int lora_seg_n;
  int flags;
  int debug;  // profiling only (IDB_GEMM_DEBUG): 1 = no MMA issue, 2 = no TMA loads, 4 = TMEM read only, 5 = no epilogue, 6 = no TMA stores
  float* out_f32;
  __nv_bfloat16* out_bf16;
  float* workspace;
  long long stats_rowblock0;   // first row block of this call inside stats (phased output)
  float2* stats;   // optional [ceil(M/32)][N] (sum, sum of squares) over each 32-row block of the fp32 output
  // optional per-IMAGE sums of the fp32 output over GRANULES of `gran` consecutive channels (gran divides the group size
  // of every GroupNorm that will consume the tensor), in 64-bit FIXED POINT, accumulated with integer atomics straight
  // from the epilogue: integer addition commutes, so the result is bit-identical whatever the arrival order (no
  // counters, no fences, no finalize pass).  [n_img][N_out / gran][2] = (sum * 2^32, sum of squares * 2^24); zero on entry.
  unsigned long long* isums;
  FastDiv fd_rbpi;             // row blocks per image of THIS call's raster
  FastDiv fd_gran;             // channels per granule
  int n_gran;                  // N_out / gran
  long long n_rowblocks;       // row blocks of this call (M / 32)
  // stream-K instantiations only (see WorkIter): groups = images, CTA pairs per group, tiles per group; partial tiles
  // travel through `workspace`, arrival counters live in g_sk_flags
  int sk_groups, sk_units, sk_tiles;
  // IDB_EPI_PHASES4: tile n_blk belongs to output phase n_blk / tiles_per_phase (weights stacked on N); phases 1-3 store
  // through their own maps (tmOutF / tmOutB are phase 0's)
  int phases4, tiles_per_phase;
  CUtensorMap tmOutP[3];
  int ws_tma;   // split-K partials leave through the staged TMA-store path (tmOutF spans the workspace, batch = k_splits x B)
};

constexpr float ISUM_SCALE_S = 4294967296.0f;    // 2^32: |sum of an (image, channel)| < 2^31
constexpr float ISUM_SCALE_SS = 16777216.0f;     // 2^24: sum of squares < 2^39 (rms 11,000 over 4096 pixels)

// ---------------------------------------------------------------------------------------- stream-K schedule
// The one-wave layers (16x16 latents at the bench batch: 64 single-N tiles on 74 CTA pairs, L2-feed bound) and the
// quantised last wave of the others leave SMs idle.  SK instantiations run the dual-N tiles of ONE IMAGE (a "group")
// as a linear space of (tile, k-block) work split evenly over P = pairs / images CTA pairs: a pair's range is
// [tail of a tile some earlier pair started] [whole tiles] [head of a tile].  The pair that holds a tile's k-block 0
// owns its epilogue; the pairs that hold later k-blocks ("contributors") write their raw fp32 accumulators to the
// workspace FIRST THING in their range and bump the tile's arrival counter; the owner finishes its head segment last,
// waits for the counter (all CTAs of the grid are co-resident: grid <= SM count, 1 CTA / SM), adds the partials in
// slot order and runs the normal epilogue.  Every image has the same schedule, so a result does not depend on where the
// image sits in the batch, and the summation order is fixed: deterministic, position independent.
constexpr int SK_MINSEG = 8;          // no segment shorter than this many k-blocks (boundaries snap to the tile edge)
constexpr int SK_MAX_TILES = 4096;
__device__ unsigned int g_sk_flags[SK_MAX_TILES * 4];   // per (tile, CTA rank): [arrived contributor warps, owner warps done]; zero between launches

__device__ __forceinline__ int sk_bound(int s, int W, int P, int nkb) {
  int x = static_cast<int>(static_cast<long long>(s) * W / P);
  const int rem = x % nkb;
  if (rem != 0) {
    if (rem < SK_MINSEG) x -= rem;
    else if (nkb - rem < SK_MINSEG) x += nkb - rem;
  }
  return x;
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* ptr) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// Work items of one CTA (pair): (tile, k-block range).  !SK: tile = unit, unit + num_units, ... with the split-K slice
// encoded in the tile index (decode_tile); SK: the per-image stream-K range described above.
template <bool SK>
struct WorkIter {
  int tile, kb_begin, kb_end;
  int n_contrib;   // SK owner of a split tile: contributors to wait for (they are the next n_contrib units)
  int num_units, total_tiles;
  int pos, hi, tile0, nkb, slot, W, P;
  __device__ __forceinline__ void load(const GemmParams& p) {
    if (SK) {
      const int t = pos / nkb;
      kb_begin = pos - t * nkb;
      kb_end = min(nkb, kb_begin + (hi - pos));
      tile = tile0 + t;
      n_contrib = 0;
      if (kb_begin == 0 && kb_end < nkb) {
        const int tile_end = (t + 1) * nkb;
        for (int s2 = slot + 1; s2 < P && sk_bound(s2, W, P, nkb) < tile_end; ++s2) ++n_contrib;
      }
    }
  }
  __device__ __forceinline__ void init(const GemmParams& p, int unit, int nunits, int ntiles, int nkb_total) {
    num_units = nunits, total_tiles = ntiles, nkb = nkb_total;
    tile = unit, kb_begin = 0, kb_end = nkb_total, n_contrib = 0;
    pos = hi = 0;
    if (SK) {
      P = p.sk_units;
      const int g = unit / P;
      slot = unit - g * P;
      W = p.sk_tiles * nkb;
      tile0 = g * p.sk_tiles;
      if (g < p.sk_groups) pos = sk_bound(slot, W, P, nkb), hi = sk_bound(slot + 1, W, P, nkb);
      if (pos < hi) load(p);
    }
  }
  __device__ __forceinline__ bool valid() const { return SK ? (pos < hi) : (tile < total_tiles); }
  __device__ __forceinline__ bool contrib() const { return SK && kb_begin > 0; }
  __device__ __forceinline__ void next(const GemmParams& p) {
    if (SK) {
      pos += kb_end - kb_begin;
      if (pos < hi) load(p);
    } else {
      tile += num_units;
    }
  }
};

template <int CG>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int rank) {
  TileCoord c;
  int t, mu, m2;
  p.fd_ntn.divmod(tile, t, c.n_blk);
  p.fd_ks.divmod(t, mu, c.ks);
  c.m_blk = mu * CG + rank;
  p.fd_tx.divmod(c.m_blk, m2, c.tx);
  p.fd_ty.divmod(m2, c.tb, c.ty);
  return c;
}

// EPI: compile-time specialisation of the epilogue (the small-K layers are bound by its instruction count):
//   0 = generic (every feature decided at run time; also split-K partials, both outputs, profiling switches)
//   1 = 16-bit output only, no GEGLU        (q/k/v, FF-out; bias / residual / PReLU still optional at run time)
//   2 = 16-bit output of GEGLU(acc + bias)  (FF-in)
//   3 = fp32 output only                    (residual-stream producers; bias / rowvec / residual / stats / PReLU optional)
// NSUB = 2 ("dual-N"): one tile is 2 x BLOCK_N output columns = two accumulators that share every A tile, i.e. two
// UMMAs per K step.  Operand bytes per MAC drop by a further (128 + N) / (128 + N/2) over the CTA pair, which is what
// the big-K convolutions are bound by (L2 -> SM operand feed); the three BLOCK_N-column TMEM buffers are used as a ring
// of which a tile occupies two, so the next tile's mainloop starts once the epilogue has drained the first half.
template <int BLOCK_N, int STAGES, bool LORA, int CG, int EPI, int NSUB, bool SK>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  static_assert(!SK || (NSUB == 2 && EPI == 3 && !LORA), "stream-K: dual-N tiles with the fp32 epilogue only");
  constexpr int UMMA_N = BLOCK_N + (LORA ? 16 : 0);
  constexpr int TILE_N = BLOCK_N * NSUB;
  constexpr int B_ROWS = UMMA_N / CG;                 // B rows this CTA stages per accumulator (half of them in a pair)
  constexpr int B_SUB_BYTES = B_ROWS * BLOCK_K * 2;
  constexpr int B_TILE_BYTES = NSUB * B_SUB_BYTES;
  static_assert(NSUB == 1 || (NSUB == 2 && !LORA && CG == 2 && 3 * BLOCK_N <= TMEM_COLS), "dual-N: CTA pairs, no LoRA, three TMEM buffers");
  constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  constexpr uint32_t IDESC = umma_idesc_bf16(BLOCK_M * CG, UMMA_N, 0, 0);
  static_assert(B_ROWS % 8 == 0, "B tile must be whole 8-row swizzle groups");
  constexpr int STG_OFFSET = STAGES * STAGE_BYTES + 1024;   // barriers live in the first 1 KiB after the ring (keeps 1024-B alignment)
  constexpr int BIAS_STG_OFFSET = STG_OFFSET + EPI_STAGING_BYTES;
  // LoRA only: the x A^T tile (bf16, 128 rows x 128 B, K-major SWIZZLE_128B, first 32 B of a row used) and two buffers
  // for the up-projection weight tile of the current N block (same layout, BLOCK_N / CG rows)
  constexpr int LORA_T_OFFSET = (BIAS_STG_OFFSET + epi_bias_bytes(TILE_N) + 1023) / 1024 * 1024;
  constexpr int LORA_U_BYTES = (BLOCK_N / CG) * BLOCK_K * 2;
  constexpr int LORA_U_OFFSET = LORA_T_OFFSET + A_TILE_BYTES;
  constexpr int LORA_BAR_OFFSET = 704;   // t_ready[4], d_full[4], up_full[2], up_empty[2] inside the barrier KiB
  static_assert(!LORA || LORA_U_BYTES % 1024 == 0, "up-projection tile must keep 1024B alignment");
  static_assert(TILE_N <= 512, "each epilogue warp stages at most four chunks per tile");
  static_assert(UMMA_N % 16 == 0 && UMMA_N <= 256, "invalid UMMA N");
  // accumulator ring in TMEM: as many buffers as fit the 512 columns, so the MMA warp can run further ahead of the
  // (latency-bound) epilogue of small-K tiles
  constexpr int TMEM_BUF_STRIDE = (UMMA_N + 31) / 32 * 32;
  constexpr int NBUF = NSUB == 2 ? 3 : ((TMEM_COLS / TMEM_BUF_STRIDE) > 4 ? 4 : (TMEM_COLS / TMEM_BUF_STRIDE));
  static_assert(NBUF >= 2, "need at least a double-buffered accumulator");
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024B alignment for SWIZZLE_128B");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;   // 0 = leader of the pair
  const int unit = blockIdx.x / CG;                            // CTA (CG=1) or CTA pair (CG=2) index
  const int num_units = gridDim.x / CG;

  if (warp == NUM_EPI_WARPS && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmW);
    if (p.nkb1 > 0) tma_prefetch_desc(&p.tmA1);
    if (LORA) {
      tma_prefetch_desc(&p.tmL);
      tma_prefetch_desc(&p.tmU);
      uint64_t* lb = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + LORA_BAR_OFFSET);
      for (int b = 0; b < 4; ++b) {
        mbar_init(&lb[b], 4 * CG);   // t_ready: the four T-extraction warps of each CTA of the pair
        mbar_init(&lb[4 + b], 1);    // d_full: up-projection MMA retired
      }
      for (int u = 0; u < 2; ++u) {
        mbar_init(&lb[8 + u], CG);   // up_full
        mbar_init(&lb[10 + u], 1);   // up_empty
      }
      for (int q = 0; q < 4; ++q) reinterpret_cast<int*>(&lb[12])[q] = 0;   // T-staging claim counters per lane quarter
    }
    if (p.residual != nullptr) tma_prefetch_desc(&p.tmRes);
    for (int w = 0; w < NUM_EPI_WARPS; ++w)
      mbar_init(reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + RES_BAR_OFFSET) + w, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], CG);     // one producer arrival per CTA of the pair (leader's barrier is the one used)
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], NUM_EPI_WARPS * CG);
    }
    mbar_fence_init();
  }
  if (warp == NUM_EPI_WARPS + 1) {
    if (CG == 2) tmem_alloc2<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // peer barriers must be initialised before remote arrives / multicast commits
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();   // the next kernel's CTAs may take over SMs as ours retire (they block in their own pdl_wait)
  pdl_wait();      // everything above touched only this CTA's smem / TMEM and the kernel parameters

  const int m_units = (p.n_tiles_m + CG - 1) / CG;   // M tiles per unit: a pair covers 2 consecutive M blocks
  const int total_tiles = m_units * p.n_tiles_n * p.k_splits;
  const int nkb_total = p.nkb0 + p.nkb1;

  if (warp == NUM_EPI_WARPS) {
    // ================================================================ TMA producer
    // The whole warp runs the loop convergently on warp-uniform values (so addresses / coordinates
    // live in uniform registers); one elected lane issues the TMA and barrier operations.
    {
      int stage = 0, pit = 0;
      uint32_t phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      WorkIter<SK> wi;
      for (wi.init(p, unit, num_units, total_tiles, nkb_total); wi.valid(); wi.next(p)) {
        const int tile = wi.tile;
        const TileCoord tc_ = decode_tile<CG>(p, tile, rank);
        const int n_blk = tc_.n_blk, ks = tc_.ks;
        const int tx = tc_.tx, ty = tc_.ty, tb = tc_.tb;   // tb >= number of batch tiles for a padding block: TMA zero-fills
        const int x0 = tx * p.BW, y0 = ty * p.BH, b0 = tb * p.BB;
        const int n0 = n_blk * TILE_N + rank * B_ROWS;
        const int lora_row = LORA ? p.fd_seg.div(n_blk * BLOCK_N) * 16 : 0;   // this tile's adapter (16 padded down-projection rows)
        const int ph = (!LORA && p.phases4) ? n_blk / p.tiles_per_phase : 0;    // output parity class of this tile (stacked weights)
        const int tap_off_x = p.phases4 ? (ph & 1) - 1 : p.tap_off_x, tap_off_y = p.phases4 ? (ph >> 1) - 1 : p.tap_off_y;
        const int kb_begin = SK ? wi.kb_begin : ks * p.kb_per_split;
        const int kb_end = SK ? wi.kb_end : min(nkb_total, kb_begin + p.kb_per_split);
        // running (tap, channel-block) counters instead of a division per k-block
        int tap = (kb_begin > 0 && kb_begin < p.nkb0) ? kb_begin / p.cpb0 : 0;   // (only split-K tiles divide)
        int cbi = (kb_begin < p.nkb0) ? kb_begin - tap * p.cpb0 : 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_TILE_BYTES;
          const uint32_t fb = smem_base + STAGES * STAGE_BYTES + stage * 8;   // &full_bar[stage]
          int c0, c1, c2, c3, c4 = 0;
          bool five = false;
          const bool seg1 = kb >= p.nkb0;
          if (!seg1) {
            const int cb = cbi * BLOCK_K;
            if (p.mode0 == IDB_A_1X1) {
              c0 = cb, c1 = x0, c2 = y0, c3 = b0;
            } else {
              const int dy = (tap * 11) >> 5, dx = tap - dy * 3;   // tap / 3 for tap in [0, 9)
              if (p.mode0 == IDB_A_3X3) {
                c0 = cb, c1 = x0 + dx - 1, c2 = y0 + dy - 1, c3 = b0;
              } else if (p.mode0 == IDB_A_2X2) {   // tap = 2 * dy + dx over the low-resolution image
                c0 = cb, c1 = x0 + (tap & 1) + tap_off_x, c2 = y0 + (tap >> 1) + tap_off_y, c3 = b0;
              } else if (p.mode0 == IDB_A_3X3_S2_ASYM) {  // stride 2, padding on the right / bottom only: input (2*yo + dy, 2*xo + dx)
                const int px = (dx == 1) ? 1 : 0, py = (dy == 1) ? 1 : 0;
                const int ox = (dx == 2) ? 1 : 0, oy = (dy == 2) ? 1 : 0;
                five = true;
                c0 = px * p.c0 + cb, c1 = x0 + ox, c2 = py, c3 = y0 + oy, c4 = b0;
              } else {  // stride 2: input (2*yo + dy - 1, 2*xo + dx - 1) in the [B, H/2, 2, W/2, 2C] view
                const int px = (dx == 1) ? 0 : 1, py = (dy == 1) ? 0 : 1;
                const int ox = (dx == 0) ? -1 : 0, oy = (dy == 0) ? -1 : 0;
                five = true;
                c0 = px * p.c0 + cb, c1 = x0 + ox, c2 = py, c3 = y0 + oy, c4 = b0;
              }
            }
            if (++cbi == p.cpb0) cbi = 0, ++tap;
          } else {
            c0 = (kb - p.nkb0) * BLOCK_K, c1 = x0, c2 = y0, c3 = b0;
          }
          const CUtensorMap* tmA = seg1 ? &p.tmA1 : &p.tmA0;
          if (elect_one()) {
            if ((p.debug & 15) == 2 || (p.debug & 15) == 3) {  // feed-rate experiment: signal the stage without moving data
              if (CG == 1 || rank == 0) mbar_expect_tx_a(fb, 0);
              else mbar_arrive_remote_a(fb, 0);
            } else if (CG == 1) {
              mbar_expect_tx_a(fb, STAGE_BYTES);
              if (five) tma_load_5d_a(sa, tmA, fb, c0, c1, c2, c3, c4);
              else tma_load_4d_a(sa, tmA, fb, c0, c1, c2, c3);
              tma_load_2d_a(sb, &p.tmW, fb, kb * BLOCK_K, n0);
              if (LORA) tma_load_2d_a(sb + BLOCK_N * BLOCK_K * 2, &p.tmL, fb, kb * BLOCK_K, lora_row);
            } else {
              if (five) tma2_load_5d_a(sa, tmA, fb, c0, c1, c2, c3, c4);
              else tma2_load_4d_a(sa, tmA, fb, c0, c1, c2, c3);
              if (!LORA) {
                tma2_load_2d_a(sb, &p.tmW, fb, kb * BLOCK_K, n0);
                if (NSUB == 2) tma2_load_2d_a(sb + B_SUB_BYTES, &p.tmW, fb, kb * BLOCK_K, n0 + BLOCK_N);   // second accumulator's rows
              } else if (rank == 0) {   // B rows [0, UMMA_N/2) of the pair's tile: all base-weight rows
                tma2_load_2d_a(sb, &p.tmW, fb, kb * BLOCK_K, n0);
              } else {                  // B rows [UMMA_N/2, UMMA_N): the remaining base rows, then the 16 LoRA-down rows
                constexpr int W1_ROWS = BLOCK_N - B_ROWS;
                tma2_load_2d_a(sb, &p.tmW1, fb, kb * BLOCK_K, n0);
                tma2_load_2d_a(sb + W1_ROWS * BLOCK_K * 2, &p.tmL, fb, kb * BLOCK_K, lora_row);
              }
              if (rank == 0) mbar_expect_tx_a(fb, 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
              else mbar_arrive_remote_a(fb, 0);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (LORA) {
          // up-projection weights of this N block (rows of this CTA's half of the tile), double-buffered; issued AFTER
          // the tile's k-blocks: the MMA warp frees buffer u while it works on those (see the MMA loop)
          const int u = pit & 1;
          uint64_t* lb = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + LORA_BAR_OFFSET);
          mbar_wait(&lb[10 + u], ((pit >> 1) & 1) ^ 1);   // up_empty
          const uint32_t ub = smem_base + STAGES * STAGE_BYTES + LORA_BAR_OFFSET + (8 + u) * 8;   // &up_full[u]
          const uint32_t us = smem_base + LORA_U_OFFSET + u * LORA_U_BYTES;
          const int urow = n_blk * BLOCK_N + rank * (BLOCK_N / CG);
          if (elect_one()) {
            if (CG == 1) {
              mbar_expect_tx_a(ub, LORA_U_BYTES);
              tma_load_2d_a(us, &p.tmU, ub, 0, urow);
            } else {
              tma2_load_2d_a(us, &p.tmU, ub, 0, urow);
              if (rank == 0) mbar_expect_tx_a(ub, 2 * LORA_U_BYTES);
              else mbar_arrive_remote_a(ub, 0);
            }
          }
          __syncwarp();
          ++pit;
        }
      }
    }
  } else if (warp == NUM_EPI_WARPS + 1) {
    // ================================================================ MMA issuer (leader CTA only in a pair)
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0, buf = 0;
      uint32_t bphase = 0;
      uint32_t mma_wait_empty = 0, mma_wait_full = 0;
      const uint32_t mma_t0 = IDB_EPI_PROF ? clock() : 0u;
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t desc_hi = umma_smem_desc_sw128(0);
      // fp16 operands: a_format / b_format fields (bits 7 and 10) are 0 = F16 instead of 1 = BF16
      const uint32_t idesc = (p.flags & IDB_EPI_F16) ? (IDESC & ~((1u << 7) | (1u << 10))) : IDESC;
      bool up_pending = false;
      int up_buf = 0;
      uint32_t up_bphase = 0;
      int dpos = 0;                      // dual-N: next free position of the 3-buffer accumulator ring
      uint32_t duse[3] = {0, 0, 0};      // dual-N: how often each buffer has been handed to a tile
      auto up_ready = [&](int ubuf, uint32_t ubphase, int uit) -> bool {   // warp-uniform probe
        uint64_t* lb = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + LORA_BAR_OFFSET);
        return mbar_try(&lb[ubuf], ubphase) && mbar_try(&lb[8 + (uit & 1)], (uit >> 1) & 1);
      };
      auto issue_up = [&](int ubuf, uint32_t ubphase, int uit) {   // D[ubuf] += (x A^T) . U^T, one K = 16 UMMA over the whole tile
        constexpr uint32_t IDESC_UP = umma_idesc_bf16(BLOCK_M * CG, BLOCK_N, 0, 0);
        uint64_t* lb = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + LORA_BAR_OFFSET);
        const int u = uit & 1;
        mbar_wait(&lb[ubuf], ubphase);             // t_ready: the T tile of that tile is staged (both CTAs)
        mbar_wait(&lb[8 + u], (uit >> 1) & 1);     // up_full: its up-projection weights have landed
        tc_fence_after();
        const uint64_t tdesc = desc_hi | static_cast<uint64_t>(((smem_base + LORA_T_OFFSET) >> 4) & 0x3FFF);
        const uint64_t udesc = desc_hi | static_cast<uint64_t>(((smem_base + LORA_U_OFFSET + u * LORA_U_BYTES) >> 4) & 0x3FFF);
        const uint32_t lb0 = smem_base + STAGES * STAGE_BYTES + LORA_BAR_OFFSET;
        if (elect_one()) {
          if (CG == 1) {
            umma_bf16(tmem_base + ubuf * TMEM_BUF_STRIDE, tdesc, udesc, IDESC_UP, 1u);
            umma_commit_a(lb0 + (4 + ubuf) * 8);   // d_full
            umma_commit_a(lb0 + (10 + u) * 8);     // up_empty
          } else {
            umma2_bf16(tmem_base + ubuf * TMEM_BUF_STRIDE, tdesc, udesc, IDESC_UP, 1u);
            umma2_commit_mc_a(lb0 + (4 + ubuf) * 8);
            umma2_commit_mc_a(lb0 + (10 + u) * 8);
          }
        }
        __syncwarp();
      };
      WorkIter<SK> wi;
      for (wi.init(p, unit, num_units, total_tiles, nkb_total); wi.valid(); wi.next(p), ++it) {
        const int tile = wi.tile;
        const int ks = decode_tile<CG>(p, tile, 0).ks;
        const int kb_begin = SK ? wi.kb_begin : ks * p.kb_per_split;
        const int kb_end = SK ? wi.kb_end : min(nkb_total, kb_begin + p.kb_per_split);
        const uint32_t c0 = IDB_EPI_PROF ? clock() : 0u;
        int buf_b = 0;
        if (NSUB == 2) {   // this tile's two accumulators: ring positions dpos, dpos + 1 (mod 3); each waits for its own drain
          buf = dpos;
          buf_b = dpos == 2 ? 0 : dpos + 1;
          mbar_wait(&tmem_empty[buf], (duse[buf] & 1) ^ 1);
          mbar_wait(&tmem_empty[buf_b], (duse[buf_b] & 1) ^ 1);
          ++duse[buf], ++duse[buf_b];
          dpos = buf_b == 2 ? 0 : buf_b + 1;
        } else if (LORA) {   // while waiting for the accumulator buffer, fire the previous tile's up-projection as soon as it can go
          while (!mbar_try(&tmem_empty[buf], bphase ^ 1)) {
            if (up_pending && up_ready(up_buf, up_bphase, it - 1)) {
              issue_up(up_buf, up_bphase, it - 1);
              up_pending = false;
            }
          }
        } else {
          mbar_wait(&tmem_empty[buf], bphase ^ 1);
        }
        tc_fence_after();
        if (IDB_EPI_PROF) mma_wait_empty += clock() - c0;
        const uint32_t d_tmem = tmem_base + buf * TMEM_BUF_STRIDE;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          // LoRA: D(previous tile) += T U^T once its T tile is staged.  It must be issued before this tile's last
          // k-block so that this tile's accumulator-ready commit also covers it (the T tile is single-buffered).
          if (LORA && up_pending && (kb == kb_end - 1 || up_ready(up_buf, up_bphase, it - 1))) {
            issue_up(up_buf, up_bphase, it - 1);
            up_pending = false;
          }
          const uint32_t c2 = IDB_EPI_PROF ? clock() : 0u;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (IDB_EPI_PROF) mma_wait_full += clock() - c2;
          // descriptors are built convergently from warp-uniform values; one elected lane issues
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = desc_hi | static_cast<uint64_t>((sa >> 4) & 0x3FFF);
          const uint64_t bdesc = desc_hi | static_cast<uint64_t>(((sa + A_TILE_BYTES) >> 4) & 0x3FFF);
          const uint32_t eb = smem_base + STAGES * STAGE_BYTES + (STAGES + stage) * 8;        // &empty_bar[stage]
          const uint32_t tf = smem_base + STAGES * STAGE_BYTES + (2 * STAGES + (NSUB == 2 ? (it & 1) : buf)) * 8;   // &tmem_full[..]
          const uint64_t bdesc2 = bdesc + static_cast<uint64_t>(B_SUB_BYTES >> 4);
          const uint32_t d_tmem2 = tmem_base + buf_b * TMEM_BUF_STRIDE;
          const uint32_t acc0 = (kb > kb_begin) ? 1u : 0u;
          if (elect_one()) {
            if ((p.debug & 15) != 1 && (p.debug & 15) != 3) {
              if (CG == 1) {
                umma_bf16(d_tmem, adesc, bdesc, idesc, acc0);
                umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);   // +32 B per UMMA_K=16 step in the swizzle atom
                umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
              } else {
                umma2_bf16(d_tmem, adesc, bdesc, idesc, acc0);
                umma2_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                umma2_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                umma2_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                if (NSUB == 2) {
                  umma2_bf16(d_tmem2, adesc, bdesc2, idesc, acc0);
                  umma2_bf16(d_tmem2, adesc + 2, bdesc2 + 2, idesc, 1u);
                  umma2_bf16(d_tmem2, adesc + 4, bdesc2 + 4, idesc, 1u);
                  umma2_bf16(d_tmem2, adesc + 6, bdesc2 + 6, idesc, 1u);
                }
              }
            }
            if (CG == 1) {
              umma_commit_a(eb);  // frees the smem slot once these MMAs retire
              if (kb == kb_end - 1) umma_commit_a(tf);
            } else {
              umma2_commit_mc_a(eb);  // ... in both CTAs of the pair
              if (kb == kb_end - 1) umma2_commit_mc_a(tf);
            }
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (LORA) up_pending = true, up_buf = buf, up_bphase = bphase;
        if (NSUB == 1 && ++buf == NBUF) buf = 0, bphase ^= 1;
      }
      if (LORA && up_pending) issue_up(up_buf, up_bphase, it - 1);
      if (IDB_EPI_PROF && (p.debug & 0x400) && lane == 0 && p.workspace != nullptr) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(p.workspace) + static_cast<size_t>(gridDim.x) * NUM_EPI_WARPS * 8 + blockIdx.x * 4;
        dst[0] = mma_wait_empty, dst[1] = mma_wait_full, dst[2] = clock() - mma_t0, dst[3] = it;
      }
    }
  } else {
    // ================================================================ epilogue warps
    // lane = accumulator row.  TMEM -> registers -> (bias, time-embedding row vector, LoRA up-projection,
    // GEGLU, fp32 residual) -> per-warp smem staging buffer (32 rows x 32 columns) -> ONE TMA store per
    // chunk (4-D box over the NHWC output; rows outside the image are clipped by the TMA unit).
    // 16 warps (4 per SM sub-partition): the per-chunk instruction stream is a long dependent chain, so
    // the layers with K <= 1280 are bound by epilogue issue latency unless several warps interleave.
    // Warp w owns TMEM lane quarter w & 3 and every 4th 32-column chunk (rotating with the tile index).
    // The fp32 residual of a chunk is TMA-loaded INTO the staging buffer it is then added to and stored from.
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int slot = warp >> 2;    // which interleaved set of 32-column chunks (rotates per tile)
    const int r = quarter * 32 + lane;
    const int rx = r % p.BW;
    const int ry = (r / p.BW) % p.BH;
    const int rb = r / (p.BW * p.BH);
    const int r0 = quarter * 32;   // first row of this warp's box inside the tile rectangle
    const int bx0 = r0 % p.BW, by0 = (r0 / p.BW) % p.BH, bb0 = r0 / (p.BW * p.BH);
    const bool geglu = EPI ? (EPI == 2) : ((p.flags & IDB_EPI_GEGLU) != 0);
    const bool has_f32 = EPI ? (EPI == 3) : (p.out_f32 != nullptr);
    const bool has_b16 = EPI ? (EPI != 3) : (p.out_bf16 != nullptr);
    const bool ksplit = EPI ? false : (p.k_splits > 1);
    const int dbg = EPI ? 0 : p.debug;
    const bool has_rowvec = (EPI == 1 || EPI == 2) ? false : (p.rowvec != nullptr);
    const bool has_prelu = (EPI == 2) ? false : (p.prelu != nullptr);
    const bool has_stats = (EPI == 1 || EPI == 2) ? false : (p.stats != nullptr || p.isums != nullptr);
    const bool f16 = (p.flags & IDB_EPI_F16) != 0;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sbuf = smem_base + STG_OFFSET + warp * EPI_BUF_BYTES;
    const uint32_t rbar = smem_base + STAGES * STAGE_BYTES + RES_BAR_OFFSET + warp * 8;
    const bool both = has_f32 && has_b16;   // rare: staged and stored one after the other
    const bool res_tma = p.residual != nullptr && !geglu && !ksplit && (dbg & 15) == 0;
    const int sw = lane & 7;       // SWIZZLE_128B phase of this lane's 128-byte fp32 staging row
    constexpr int NCH = TILE_N / 32;
    constexpr int SUBCH = BLOCK_N / 32;   // chunks per accumulator
    constexpr int MAXC = (NCH + 3) / 4;   // chunks per warp per tile
    float* bsm = reinterpret_cast<float*>(smem + BIAS_STG_OFFSET) + warp * (MAXC * 32);
    uint32_t g = 0;                // residual chunks loaded so far (barrier parity = g & 1)
    int buf = 0;                   // accumulator ring position of the current tile
    uint32_t bphase = 0;
    int dpos = 0;                  // dual-N: ring position / per-buffer use counts, mirrored from the MMA warp
    uint32_t duse[3] = {0, 0, 0};
    const bool prof = IDB_EPI_PROF && (dbg & 0x400) != 0;   // per-warp clock() breakdown of the epilogue phases -> p.workspace (build with -DIDB_EPI_PROF=1)
    uint32_t tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t tc = 0;
#define IDB_TICK(k)                       \
  if (prof) {                             \
    const uint32_t now = clock();         \
    tp[k] += now - tc;                    \
    tc = now;                             \
  }
    if (prof) tc = clock();

    float pre_bias[MAXC];
    auto prefetch_tables = [&](int tl, int itn) {   // lane c fetches column col + c of each chunk (coalesced)
      const int nb_ = decode_tile<CG>(p, tl, 0).n_blk;
      const int n0_ = nb_ * TILE_N;
      const int poff_ = (!LORA && p.phases4) ? (nb_ / p.tiles_per_phase) * p.N_out : 0;   // the four phases share the bias
#pragma unroll
      for (int ci = 0; ci < MAXC; ++ci) {
        const int col = n0_ + (((slot + itn) & 3) + 4 * ci) * 32 + lane;
        const bool ok = (((slot + itn) & 3) + 4 * ci) < NCH && col < p.N;
        pre_bias[ci] = (ok && p.bias != nullptr) ? __ldg(p.bias + col - poff_) : 0.f;
      }
    };
    if (!SK && unit < total_tiles) prefetch_tables(unit, 0);

    int it = 0;
    WorkIter<SK> wi;
    for (wi.init(p, unit, num_units, total_tiles, nkb_total); wi.valid(); wi.next(p), ++it) {
      const int tile = wi.tile;
      const bool contrib = wi.contrib();          // stream-K: this item only adds a K segment to a tile another pair owns
      const int n_contrib = SK ? wi.n_contrib : 0;
      const bool res_item = res_tma && !contrib;
      unsigned int* sk_flag = SK ? &g_sk_flags[(tile * CG + static_cast<int>(rank)) * 2] : nullptr;
      if (SK) prefetch_tables(tile, it);          // (no look-ahead: the loads retire while the accumulator is still being computed)
      const TileCoord tc_ = decode_tile<CG>(p, tile, rank);
      const int n_blk = tc_.n_blk, ks = tc_.ks, m_blk = tc_.m_blk;
      const int tx = tc_.tx, ty = tc_.ty, tb = tc_.tb;
      const int x = tx * p.BW + rx, y = ty * p.BH + ry, b = tb * p.BB + rb;
      const bool row_ok = (x < p.Wo) && (y < p.Ho) && (b < p.B);
      const long long orow = (static_cast<long long>(b) * p.Ho + y) * p.Wo + x;
      const int n0 = n_blk * TILE_N;
      const int ph = (!LORA && p.phases4) ? n_blk / p.tiles_per_phase : 0;
      const int ph_col0 = ph * p.N_out;                                   // first stacked column of this tile's phase
      const CUtensorMap* tm_out_f = ph == 0 ? &p.tmOutF : &p.tmOutP[ph - 1];
      const CUtensorMap* tm_out_b = ph == 0 ? &p.tmOutB : &p.tmOutP[ph - 1];
      const int chunk0 = (slot + it) & 3;
      const int cx = tx * p.BW + bx0, cy = ty * p.BH + by0, cb = tb * p.BB + bb0;   // this warp's 32-row box

      // ---- before the accumulator is ready: residual of the first chunk (TMA), bias and LoRA-up weights of
      // all of this warp's chunks of the tile (lane c fetches column col + c, coalesced; kept in per-warp smem)
      if (res_item && chunk0 < NCH && n0 + chunk0 * 32 < p.N && lane == 0) {
        tma_store_wait_read();                      // the previous store has finished reading the staging buffer
        mbar_expect_tx_a(rbar, EPI_BUF_BYTES);
        tma_load_4d_a(sbuf, &p.tmRes, rbar, n0 + chunk0 * 32, cx, cy, cb);
      }
      // bias / LoRA-up weights of this warp's chunks were fetched into registers one tile ago: publish them to
      // the per-warp smem tables, then request the next tile's (their latency hides behind this tile's work)
      __syncwarp();   // the previous tile's reads of bsm / lsm are complete
#pragma unroll
      for (int ci = 0; ci < MAXC; ++ci) {
        if (p.bias != nullptr) bsm[ci * 32 + lane] = pre_bias[ci];
      }
      __syncwarp();
      if (!SK && tile + num_units < total_tiles) prefetch_tables(tile + num_units, it + 1);
      IDB_TICK(0);   // tile prologue (bias / LoRA prefetch, residual request)
      int buf_b = 0;
      uint32_t full_phase = bphase;
      int full_idx = buf;
      if (NSUB == 2) {
        buf = dpos;
        buf_b = dpos == 2 ? 0 : dpos + 1;
        ++duse[buf], ++duse[buf_b];
        dpos = buf_b == 2 ? 0 : buf_b + 1;
        full_idx = it & 1;
        full_phase = (it >> 1) & 1;
      }
      const uint32_t t_lanes = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const uint32_t t_row = t_lanes + buf * TMEM_BUF_STRIDE;
      if (LORA) {
        // fused LoRA: columns [BLOCK_N, BLOCK_N + 16) of the accumulator hold T = x A^T (one adapter per N segment).
        // The slot-0 warps round T to bf16 into a K-major operand tile; the MMA warp then adds T U^T (U = scaled
        // up-projection rows of this N block) to the accumulator with one more UMMA, and everybody waits for that.
        uint64_t* lb = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + LORA_BAR_OFFSET);
        // the first warp of each lane quarter to reach this tile stages T (the others may still be storing the previous tile)
        int* claim = reinterpret_cast<int*>(smem + STAGES * STAGE_BYTES + LORA_BAR_OFFSET + 12 * 8) + quarter;
        int mine = 0;
        if (lane == 0) mine = (atomicCAS(claim, it, it + 1) == it) ? 1 : 0;
        mine = __shfl_sync(0xffffffffu, mine, 0);
        if (mine) {
          mbar_wait(&tmem_full[buf], bphase);
          tc_fence_after();
          uint32_t lv[16];
          IDB_TMEM_LD_X16(t_row + BLOCK_N, lv);
          tmem_ld_wait();
          const uint32_t trow = smem_base + LORA_T_OFFSET + r * 128;   // row r of the T tile; 16-byte chunk c sits at c ^ (r & 7)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(trow + ((0 ^ sw) << 4)),
                       "r"(pack_bf16x2(__uint_as_float(lv[0]), __uint_as_float(lv[1]))),
                       "r"(pack_bf16x2(__uint_as_float(lv[2]), __uint_as_float(lv[3]))),
                       "r"(pack_bf16x2(__uint_as_float(lv[4]), __uint_as_float(lv[5]))),
                       "r"(pack_bf16x2(__uint_as_float(lv[6]), __uint_as_float(lv[7])))
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(trow + ((1 ^ sw) << 4)),
                       "r"(pack_bf16x2(__uint_as_float(lv[8]), __uint_as_float(lv[9]))),
                       "r"(pack_bf16x2(__uint_as_float(lv[10]), __uint_as_float(lv[11]))),
                       "r"(pack_bf16x2(__uint_as_float(lv[12]), __uint_as_float(lv[13]))),
                       "r"(pack_bf16x2(__uint_as_float(lv[14]), __uint_as_float(lv[15])))
                       : "memory");
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2 && rank != 0) mbar_arrive_remote(&lb[buf], 0);
            else mbar_arrive(&lb[buf]);
          }
        }
        mbar_wait(&lb[4 + buf], bphase);   // d_full
      } else {
        mbar_wait(&tmem_full[full_idx], full_phase);
      }
      tc_fence_after();
      IDB_TICK(1);   // waiting for the accumulator
      bool released_a = false;
      if (SK && n_contrib > 0) {   // the other K segments of this tile: every epilogue warp of every contributor has stored and fenced
        if (lane == 0) {
          const unsigned int need = static_cast<unsigned int>(n_contrib) * NUM_EPI_WARPS;
          unsigned int polls = 0;
          while (ld_acquire_gpu(sk_flag) < need) {
            __nanosleep(64);
            // the contributors finished their segments long before this point; seconds of waiting mean the grid is not
            // co-resident (two stream-K kernels of different streams sharing the device, see launch_gemm_e): fail loudly
            // (a launch error the host sees) instead of hanging the GPU
            if (++polls > (1u << 22)) __trap();   // (~0.8 us per poll: a few seconds)
          }
        }
        __syncwarp();
      }

      int ci = 0;
      for (int chunk = chunk0; chunk < NCH; chunk += 4, ++ci) {
        const int col = n0 + chunk * 32;
        uint32_t v[32];
        if ((dbg & 15) == 5) continue;   // profiling: no TMEM read, no stores
        if (res_item && ci > 0 && col < p.N && lane == 0) {   // (the first chunk's residual was requested above)
          tma_store_wait_read();
          mbar_expect_tx_a(rbar, EPI_BUF_BYTES);
          tma_load_4d_a(sbuf, &p.tmRes, rbar, col, cx, cy, cb);
        }
        if (NSUB == 2 && chunk >= SUBCH && !released_a) {   // done with the first accumulator: the next tile's mainloop may reuse it
          released_a = true;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank != 0) mbar_arrive_remote(&tmem_empty[buf], 0);
            else mbar_arrive(&tmem_empty[buf]);
          }
        }
        IDB_TMEM_LD_X32(NSUB == 2 ? (t_lanes + (chunk >= SUBCH ? buf_b : buf) * TMEM_BUF_STRIDE + (chunk >= SUBCH ? chunk - SUBCH : chunk) * 32)
                                  : (t_row + chunk * 32), v);
        tmem_ld_wait();
        IDB_TICK(2);   // TMEM load
        if (col >= p.N || (dbg & 15) == 4) continue;   // warp-uniform (debug 4: TMEM read only)
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);

        if (SK && contrib) {   // raw fp32 segment sums -> this unit's slot of the workspace, [chunk][lane quarter][column][lane]:
          // one coalesced 128-byte line per accumulator column
          float* dst = p.workspace + (((static_cast<long long>(unit) * CG + rank) * NCH + chunk) * 4 + quarter) * 1024 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(dst + j * 32, acc[j]);
          continue;
        }
        if (SK && n_contrib > 0) {   // add the contributors' segments in slot order (fixed summation order)
          for (int jc = 0; jc < n_contrib; ++jc) {
            const float* src = p.workspace + (((static_cast<long long>(unit + 1 + jc) * CG + rank) * NCH + chunk) * 4 + quarter) * 1024 + lane;
            float pv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) pv[j] = __ldcg(src + j * 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += pv[j];
          }
        }
        if (ksplit && p.ws_tma) {   // raw partial sums -> staging buffer -> one TMA store into slice `ks` of the workspace
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          const uint32_t row = sbuf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((j ^ sw) << 4)), "f"(acc[4 * j]),
                         "f"(acc[4 * j + 1]), "f"(acc[4 * j + 2]), "f"(acc[4 * j + 3])
                         : "memory");
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (cb < p.B) tma_store_4d(&p.tmOutF, sbuf, col, cx, cy, ks * p.B + cb);   // (a box past the batch would land in the next slice)
            tma_store_commit();
          }
          continue;
        }
        if (ksplit) {  // raw partial sums (tiny-M layers only); the finalize kernel applies the epilogue
          if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(p.workspace + (static_cast<long long>(ks) * p.M + orow) * p.N + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          }
          continue;
        }
        if (p.bias != nullptr && !(dbg & 0x20)) {
          const float4* bp = reinterpret_cast<const float4*>(bsm + ci * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = bp[j];
            fadd2(acc[4 * j], acc[4 * j + 1], acc[4 * j], acc[4 * j + 1], bv.x, bv.y);
            fadd2(acc[4 * j + 2], acc[4 * j + 3], acc[4 * j + 2], acc[4 * j + 3], bv.z, bv.w);
          }
        }
        if (has_rowvec && row_ok) {
          const float4* rp = reinterpret_cast<const float4*>(p.rowvec + static_cast<long long>(b) * p.rowvec_ld + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 rv = __ldg(rp + j);
            fadd2(acc[4 * j], acc[4 * j + 1], acc[4 * j], acc[4 * j + 1], rv.x, rv.y);
            fadd2(acc[4 * j + 2], acc[4 * j + 3], acc[4 * j + 2], acc[4 * j + 3], rv.z, rv.w);
          }
        }
        if (has_prelu) {   // per-channel PReLU (IResNet)
          const float4* sp = reinterpret_cast<const float4*>(p.prelu + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 sl = __ldg(sp + j);
            acc[4 * j] = acc[4 * j] > 0.f ? acc[4 * j] : acc[4 * j] * sl.x;
            acc[4 * j + 1] = acc[4 * j + 1] > 0.f ? acc[4 * j + 1] : acc[4 * j + 1] * sl.y;
            acc[4 * j + 2] = acc[4 * j + 2] > 0.f ? acc[4 * j + 2] : acc[4 * j + 2] * sl.z;
            acc[4 * j + 3] = acc[4 * j + 3] > 0.f ? acc[4 * j + 3] : acc[4 * j + 3] * sl.w;
          }
        }
        if (!EPI && (p.flags & IDB_EPI_GELU)) {   // plain erf-GELU on every column (CLIP text MLP)
#pragma unroll
          for (int j = 0; j < 32; j += 2) geglu2(acc[j], acc[j + 1], 1.0f, 1.0f, acc[j], acc[j + 1]);
        }
        int nc = 32, ocol = col - ph_col0;
        if (geglu) {  // chunk = [a(16) | g(16)] -> 16 outputs at column col/2
#pragma unroll
          for (int j = 0; j < 16; j += 2) geglu2(acc[j], acc[j + 1], acc[j], acc[j + 1], acc[16 + j], acc[17 + j]);
          nc = 16, ocol = col >> 1;
        }
        IDB_TICK(3);   // bias / rowvec / LoRA / GEGLU math
        if (res_item) {
          // this chunk's residual has landed in the staging buffer: add it in place (each lane touches only its row)
          mbar_wait(reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + RES_BAR_OFFSET) + warp, g & 1);
          ++g;
          const uint32_t row = sbuf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 rv;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(rv.x), "=f"(rv.y), "=f"(rv.z), "=f"(rv.w)
                         : "r"(row + ((j ^ sw) << 4)));
            fadd2(acc[4 * j], acc[4 * j + 1], acc[4 * j], acc[4 * j + 1], rv.x, rv.y);
            fadd2(acc[4 * j + 2], acc[4 * j + 3], acc[4 * j + 2], acc[4 * j + 3], rv.z, rv.w);
          }
        } else {
          if (p.residual != nullptr && row_ok) {
            const float4* rp = reinterpret_cast<const float4*>(p.residual + orow * p.N_out + ocol);
            if (geglu) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 rv = __ldg(rp + j);
                acc[4 * j] += rv.x, acc[4 * j + 1] += rv.y, acc[4 * j + 2] += rv.z, acc[4 * j + 3] += rv.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 rv = __ldg(rp + j);
                acc[4 * j] += rv.x, acc[4 * j + 1] += rv.y, acc[4 * j + 2] += rv.z, acc[4 * j + 3] += rv.w;
              }
            }
          }
          if (lane == 0) tma_store_wait_read();   // the previous store has finished reading the staging buffer
          __syncwarp();
        }
        IDB_TICK(4);   // residual wait + add, or wait for the previous store's smem read
        if (has_stats && !row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = 0.f;
        }
        if (has_f32 && !(dbg & 0x40)) {   // 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row `lane` lands at j ^ (lane & 7)
          const uint32_t row = sbuf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j * 4 < nc)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((j ^ sw) << 4)), "f"(acc[4 * j]),
                           "f"(acc[4 * j + 1]), "f"(acc[4 * j + 2]), "f"(acc[4 * j + 3])
                           : "memory");
          }
          if (has_stats) {
            // GroupNorm statistics of the tensor being written, for free: lane c reduces column c of the
            // staged 32 x 32 fp32 chunk (rows outside the image were staged as zeros)
            __syncwarp();
            float cs = 0.f, cs2 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              float v;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbuf + rr * 128 + ((((lane >> 2) ^ (rr & 7))) << 4) + ((lane & 3) << 2)));
              cs += v;
              cs2 = fmaf(v, v, cs2);
            }
            const long long rb_local = static_cast<long long>(m_blk) * 4 + quarter;   // (padding row blocks of a ragged batch tile write nothing)
            if (p.stats != nullptr && rb_local < p.n_rowblocks && ocol + lane < p.N_out)
              p.stats[(p.stats_rowblock0 + ph * p.n_rowblocks + rb_local) * p.N_out + ocol + lane] = make_float2(cs, cs2);
            if (p.isums != nullptr) {
              // column sums -> granule sums: segmented warp reduction (lanes of one granule are contiguous; fixed tree, so
              // deterministic), then ONE pair of fire-and-forget integer reductions (RED.ADD.64) per granule segment
              const int gid = p.fd_gran.div(ocol + lane);
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_down_sync(0xffffffffu, cs, o), t2 = __shfl_down_sync(0xffffffffu, cs2, o);
                const int g2 = __shfl_down_sync(0xffffffffu, gid, o);
                if (lane + o < 32 && g2 == gid) cs += t, cs2 += t2;
              }
              const int gprev = __shfl_up_sync(0xffffffffu, gid, 1);
              if ((lane == 0 || gprev != gid) && rb_local < p.n_rowblocks && ocol + lane < p.N_out) {
                const int img = p.fd_rbpi.div(static_cast<int>(rb_local));
                unsigned long long* d = p.isums + (static_cast<long long>(img) * p.n_gran + gid) * 2;
                atomicAdd(d, static_cast<unsigned long long>(__float2ll_rn(cs * ISUM_SCALE_S)));
                atomicAdd(d + 1, static_cast<unsigned long long>(__float2ll_rn(cs2 * ISUM_SCALE_SS)));
              }
            }
          }
          IDB_TICK(5);   // staging (+ statistics)
          if (!(dbg & 0x10)) fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          IDB_TICK(6);   // proxy fence
          if (lane == 0) {
            if (!(dbg & 0x80)) tma_store_4d(tm_out_f, sbuf, ocol, cx, cy, cb);
            tma_store_commit();
            if (both) tma_store_wait_read();   // the bf16 copy is staged in the same buffer next
          }
          if (both) __syncwarp();
        }
        if (has_b16 && !(dbg & 0x40)) {  // dense rows of nc * 2 bytes
          const uint32_t row = sbuf + lane * (nc * 2);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j * 8 < nc)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (j << 4)),
                           "r"(pack_16x2(acc[8 * j], acc[8 * j + 1], f16)), "r"(pack_16x2(acc[8 * j + 2], acc[8 * j + 3], f16)),
                           "r"(pack_16x2(acc[8 * j + 4], acc[8 * j + 5], f16)), "r"(pack_16x2(acc[8 * j + 6], acc[8 * j + 7], f16))
                           : "memory");
          }
          IDB_TICK(5);   // staging
          if (!(dbg & 0x10)) fence_proxy_async_smem();
          __syncwarp();
          IDB_TICK(6);   // proxy fence
          if (lane == 0) {
            if (!(dbg & 0x80)) tma_store_4d(tm_out_b, sbuf, ocol, cx, cy, cb);
            tma_store_commit();
          }
        }
        IDB_TICK(7);   // store issue
      }
      if (SK && contrib) {   // publish this warp's share of the segment: stores -> fence -> one arrival per warp
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(sk_flag, 1u);
      } else if (SK && n_contrib > 0) {   // the last owner warp through re-arms the tile's counters for the next launch
        __syncwarp();
        if (lane == 0 && atomicAdd(sk_flag + 1, 1u) == NUM_EPI_WARPS - 1) sk_flag[0] = 0u, sk_flag[1] = 0u;
      }
      // this warp is done reading the accumulator buffer(s)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NSUB == 2) {
          if (!released_a) {
            if (rank != 0) mbar_arrive_remote(&tmem_empty[buf], 0);
            else mbar_arrive(&tmem_empty[buf]);
          }
          if (rank != 0) mbar_arrive_remote(&tmem_empty[buf_b], 0);
          else mbar_arrive(&tmem_empty[buf_b]);
        } else if (CG == 2 && rank != 0) {
          mbar_arrive_remote(&tmem_empty[buf], 0);   // the leader's MMA warp waits on ITS barrier
        } else {
          mbar_arrive(&tmem_empty[buf]);
        }
      }
      if (NSUB == 1 && ++buf == NBUF) buf = 0, bphase ^= 1;
    }
    if (prof && lane == 0 && p.workspace != nullptr) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(p.workspace) + (static_cast<size_t>(blockIdx.x) * NUM_EPI_WARPS + warp) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = tp[k];
    }
#undef IDB_TICK
    if (lane == 0) tma_store_wait_all();   // every output byte is written before the CTA retires
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // the peer's smem / TMEM stay alive until every MMA and epilogue of the pair is done
  else __syncthreads();
  if (warp == NUM_EPI_WARPS + 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------- split-K finalize
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, int k_splits, long long M, int N, int hw,
                                       const float* __restrict__ bias, const float* __restrict__ rowvec,
                                       long long rowvec_ld, const float* __restrict__ prelu,
                                       const float* __restrict__ residual, float* __restrict__ out_f32,
                                       __nv_bfloat16* __restrict__ out_bf16, int f16) {
  pdl_trigger();
  pdl_wait();
  const long long total4 = M * N / 4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 4;
    const long long row = e / N;
    const int col = static_cast<int>(e - row * N);
    // partial sums of the k_splits slices, four independent loads in flight (fixed summation order: deterministic)
    const long long slice = M * N;
    float4 a = *reinterpret_cast<const float4*>(ws + e);
    int s = 1;
    for (; s + 3 < k_splits; s += 4) {
      const float4 v0 = *reinterpret_cast<const float4*>(ws + (s + 0) * slice + e);
      const float4 v1 = *reinterpret_cast<const float4*>(ws + (s + 1) * slice + e);
      const float4 v2 = *reinterpret_cast<const float4*>(ws + (s + 2) * slice + e);
      const float4 v3 = *reinterpret_cast<const float4*>(ws + (s + 3) * slice + e);
      a.x += (v0.x + v1.x) + (v2.x + v3.x), a.y += (v0.y + v1.y) + (v2.y + v3.y);
      a.z += (v0.z + v1.z) + (v2.z + v3.z), a.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; s < k_splits; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(ws + s * slice + e);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    if (bias) {
      const float4 v = *reinterpret_cast<const float4*>(bias + col);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    if (rowvec) {
      const float4 v = *reinterpret_cast<const float4*>(rowvec + (row / hw) * rowvec_ld + col);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    if (prelu) {
      const float4 sl = *reinterpret_cast<const float4*>(prelu + col);
      a.x = a.x > 0.f ? a.x : a.x * sl.x, a.y = a.y > 0.f ? a.y : a.y * sl.y;
      a.z = a.z > 0.f ? a.z : a.z * sl.z, a.w = a.w > 0.f ? a.w : a.w * sl.w;
    }
    if (residual) {
      const float4 v = *reinterpret_cast<const float4*>(residual + e);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + e) = a;
    if (out_bf16) *reinterpret_cast<uint2*>(out_bf16 + e) = make_uint2(pack_16x2(a.x, a.y, f16 != 0), pack_16x2(a.z, a.w, f16 != 0));
  }
}

// Split-K finalize of a tiny-M layer that ALSO needs GroupNorm statistics: one CTA = (32 output columns, image); its 256
// threads are 8 column quads x 32 row lanes that walk the image's rows, add up the K splits (fixed order), apply the
// epilogue, write the output and keep per-column sums, which a fixed-order smem reduction turns into the per-image channel
// sums -- one launch instead of finalize + row-block statistics + GroupNorm finalize.
__global__ void __launch_bounds__(256) splitk_finalize_sums_kernel(const float* __restrict__ ws, int k_splits, long long M, int N, int hw_conv,
                                                                   int hw_img, const float* __restrict__ bias, const float* __restrict__ rowvec,
                                                                   long long rowvec_ld, const float* __restrict__ prelu,
                                                                   const float* __restrict__ residual, float* __restrict__ out_f32,
                                                                   unsigned long long* __restrict__ isums, int gran) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh_s[32][33], sh_ss[32][33];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int col = blockIdx.x * 32 + tx * 4;
  const int img = blockIdx.y;
  const long long slice = M * N;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), sl = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bv = *reinterpret_cast<const float4*>(bias + col);
  if (prelu) sl = *reinterpret_cast<const float4*>(prelu + col);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = ty; r < hw_img; r += 32) {
    const long long row = static_cast<long long>(img) * hw_img + r;
    const long long e = row * N + col;
    float4 a = *reinterpret_cast<const float4*>(ws + e);
    int sp = 1;
    for (; sp + 3 < k_splits; sp += 4) {   // same summation order as splitk_finalize_kernel
      const float4 v0 = *reinterpret_cast<const float4*>(ws + (sp + 0) * slice + e);
      const float4 v1 = *reinterpret_cast<const float4*>(ws + (sp + 1) * slice + e);
      const float4 v2 = *reinterpret_cast<const float4*>(ws + (sp + 2) * slice + e);
      const float4 v3 = *reinterpret_cast<const float4*>(ws + (sp + 3) * slice + e);
      a.x += (v0.x + v1.x) + (v2.x + v3.x), a.y += (v0.y + v1.y) + (v2.y + v3.y);
      a.z += (v0.z + v1.z) + (v2.z + v3.z), a.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; sp < k_splits; ++sp) {
      const float4 v = *reinterpret_cast<const float4*>(ws + sp * slice + e);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    a.x += bv.x, a.y += bv.y, a.z += bv.z, a.w += bv.w;
    if (rowvec) {
      const float4 v = *reinterpret_cast<const float4*>(rowvec + (row / hw_conv) * rowvec_ld + col);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    if (prelu) {
      a.x = a.x > 0.f ? a.x : a.x * sl.x, a.y = a.y > 0.f ? a.y : a.y * sl.y;
      a.z = a.z > 0.f ? a.z : a.z * sl.z, a.w = a.w > 0.f ? a.w : a.w * sl.w;
    }
    if (residual) {
      const float4 v = *reinterpret_cast<const float4*>(residual + e);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    *reinterpret_cast<float4*>(out_f32 + e) = a;
    s[0] += a.x, s[1] += a.y, s[2] += a.z, s[3] += a.w;
    ss[0] = fmaf(a.x, a.x, ss[0]), ss[1] = fmaf(a.y, a.y, ss[1]), ss[2] = fmaf(a.z, a.z, ss[2]), ss[3] = fmaf(a.w, a.w, ss[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sh_s[ty][tx * 4 + j] = s[j], sh_ss[ty][tx * 4 + j] = ss[j];
  __syncthreads();
  if (threadIdx.x < 32) {
    // same fixed-point format and granule layout as the GEMM epilogue (a granule may straddle two CTAs' column blocks, so
    // the totals are ADDED with integer atomics -- exact and order-independent; the accumulators are zero on entry)
    float t = 0.f, tt = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += sh_s[y][threadIdx.x], tt += sh_ss[y][threadIdx.x];
    const int lane = threadIdx.x, colg = blockIdx.x * 32 + lane;
    const int gid = colg / gran;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float u = __shfl_down_sync(0xffffffffu, t, o), u2 = __shfl_down_sync(0xffffffffu, tt, o);
      const int g2 = __shfl_down_sync(0xffffffffu, gid, o);
      if (lane + o < 32 && g2 == gid) t += u, tt += u2;
    }
    const int gprev = __shfl_up_sync(0xffffffffu, gid, 1);
    if (lane == 0 || gprev != gid) {
      unsigned long long* d = isums + (static_cast<long long>(img) * (N / gran) + gid) * 2;
      atomicAdd(d, static_cast<unsigned long long>(__float2ll_rn(t * ISUM_SCALE_S)));
      atomicAdd(d + 1, static_cast<unsigned long long>(__float2ll_rn(tt * ISUM_SCALE_SS)));
    }
  }
}

// statistics of 32-row blocks of a finished fp32 [M, N] tensor (split-K path of tiny-M layers); the
// row-block numbering matches the tile order of the fused epilogue only for raster-ordered tiles,
// which is what the host guarantees before using this path (BW * BH multiple of 32 rows per image).
__global__ void rowblock_stats_kernel(const float* __restrict__ x, long long M, int N, float2* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  const long long rowblock = blockIdx.x;   // grid = (row blocks, ceil(N / 128)), one thread per column
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const long long row0 = rowblock * 32;
  float v[32];
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) v[rr] = (row0 + rr < M) ? x[(row0 + rr) * N + c] : 0.f;   // 32 independent loads in flight
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {   // same summation order as the fused epilogue
    s += v[rr];
    s2 = fmaf(v[rr], v[rr], s2);
  }
  stats[rowblock * N + c] = make_float2(s, s2);
}

// ---------------------------------------------------------------------------------------- host side
static int pow2_divisor(int v, int cap) {
  int d = 1;
  while (d * 2 <= cap && v % (d * 2) == 0) d *= 2;
  return d;
}

// stream-K launches: 0 = cooperative + programmatic dependent launch, 1 = cooperative only, 2 = unavailable on this driver
static std::atomic<int> g_sk_launch_mode{0};
static int env_int(const char* name, int dflt);

template <int BLOCK_N, int STAGES, bool LORA, int CG, int EPI, int NSUB = 1, bool SK = false>
static int launch_gemm_e(const GemmParams& p, int grid, cudaStream_t stream) {
  constexpr int UMMA_N = BLOCK_N + (LORA ? 16 : 0);
  constexpr int smem_bytes = STAGES * (A_TILE_BYTES + NSUB * (UMMA_N / CG) * BLOCK_K * 2) + 1024 + 1024 + EPI_STAGING_BYTES +
                             epi_bias_bytes(BLOCK_N * NSUB) + (LORA ? 1024 + A_TILE_BYTES + 2 * (BLOCK_N / CG) * BLOCK_K * 2 : 0);   // ring + slack + barriers + staging + bias (+ LoRA T / U tiles)
  static_assert(smem_bytes <= 227 * 1024, "shared memory budget");
  auto kern = gemm_tc_kernel<BLOCK_N, STAGES, LORA, CG, EPI, NSUB, SK>;
  static PerDeviceOnce configured;  // per instantiation (and per device inside)
  {
    cudaError_t e = ensure_dynamic_smem(kern, smem_bytes, configured);
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  if (!SK) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    note_launch();
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("gemm_tc launch: ") + cudaGetErrorString(e));
    return IDB_OK;
  }
  // stream-K: CTAs wait for each other inside the launch, so the WHOLE grid must become resident.  The grid is one CTA per
  // SM and nothing the resident CTAs wait for depends on another kernel, so on a stream of its own (this pipeline, its VAE
  // and NCCL's gather kernels included) every CTA gets an SM as soon as the kernels in front retire: a plain launch is
  // deadlock free and is the default -- it is also the only form Nsight Compute can replay (a cooperative cluster launch
  // fails under ncu with LaunchFailed).  Two stream-K kernels of DIFFERENT streams (or MPS clients) could each hold part
  // of the machine and wait forever: such deployments set IDB_GEMM_SK_COOP=1, which makes the launch COOPERATIVE (the
  // driver then guarantees co-residency; combined with programmatic dependent launch when the driver accepts both), or
  // IDB_GEMM_SK=0.
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  attr[2].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[2].val.programmaticStreamSerializationAllowed = 1;
  static const int coop = env_int("IDB_GEMM_SK_COOP", 0);
  if (!coop) {
    attr[1] = attr[2];
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    note_launch();
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("gemm_tc (stream-K, plain) launch: ") + cudaGetErrorString(e));
    return IDB_OK;
  }
  for (;;) {
    const int mode = g_sk_launch_mode.load(std::memory_order_relaxed);
    if (mode >= 2) return IDB_E_UNSUPPORTED;   // (the caller re-plans without stream-K)
    cfg.numAttrs = (mode == 0 && pdl_enabled()) ? 3 : 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e == cudaSuccess) {
      note_launch();
      return IDB_OK;
    }
    cudaGetLastError();   // clear the sticky launch-configuration error and try the next mode
    int expect = mode;
    g_sk_launch_mode.compare_exchange_strong(expect, (mode == 0 && !pdl_enabled()) ? 2 : mode + 1);
  }
}

// epi: 0 generic, 1 / 2 / 3 specialised (see gemm_tc_kernel); only the CTA-pair kernels the UNet spends its time in are
// specialised, everything else runs the generic epilogue
template <int BLOCK_N, int STAGES, bool LORA, int CG>
static int launch_gemm(const GemmParams& p, int grid, cudaStream_t stream, int epi) {
  if (CG == 2) {
    if (epi == 1) return launch_gemm_e<BLOCK_N, STAGES, LORA, CG, (CG == 2 ? 1 : 0)>(p, grid, stream);
    if (epi == 2 && !LORA) return launch_gemm_e<BLOCK_N, STAGES, LORA, CG, (CG == 2 && !LORA ? 2 : 0)>(p, grid, stream);
    if (epi == 3) return launch_gemm_e<BLOCK_N, STAGES, LORA, CG, (CG == 2 ? 3 : 0)>(p, grid, stream);
  }
  return launch_gemm_e<BLOCK_N, STAGES, LORA, CG, 0>(p, grid, stream);
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace idb

using namespace idb;

extern "C" int idb_stream_k_mode(void) { return g_sk_launch_mode.load(std::memory_order_relaxed); }

extern "C" size_t idb_gemm_conv_workspace_bytes(int64_t m, int64_t n, int32_t k_splits) {
  return k_splits > 1 ? static_cast<size_t>(m) * n * k_splits * sizeof(float) : 0;
}

extern "C" int idb_gemm_conv(const idb_gemm_conv_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return fail(IDB_E_BADARG, "idb_gemm_conv: null args");
  if (int rc = require_sm100()) return rc;
  if (!a->a0 || !a->w) return fail(IDB_E_BADARG, "idb_gemm_conv: null operand");
  if (a->c0 <= 0 || a->c0 % 64) return fail(IDB_E_BADARG, "idb_gemm_conv: C0 must be a positive multiple of 64");
  if (a->a1 && (a->c1 <= 0 || a->c1 % 64)) return fail(IDB_E_BADARG, "idb_gemm_conv: C1 must be a multiple of 64");
  if (a->n <= 0 || a->n % 32) return fail(IDB_E_BADARG, "idb_gemm_conv: N must be a positive multiple of 32");
  if (a->batch <= 0 || a->height <= 0 || a->width <= 0) return fail(IDB_E_BADARG, "idb_gemm_conv: bad geometry");
  if (a->a0_mode < IDB_A_1X1 || a->a0_mode > IDB_A_2X2) return fail(IDB_E_BADARG, "idb_gemm_conv: bad a0_mode");
  if (a->a0_mode == IDB_A_2X2 && (a->tap_off_x < -1 || a->tap_off_x > 0 || a->tap_off_y < -1 || a->tap_off_y > 0))
    return fail(IDB_E_BADARG, "idb_gemm_conv: 2x2 tap offsets must be -1 or 0");
  const bool phased = a->out_scale == 2;
  if (a->out_scale != 0 && a->out_scale != 1 && a->out_scale != 2) return fail(IDB_E_BADARG, "idb_gemm_conv: out_scale must be 0, 1 or 2");
  if (phased && ((a->out_phase_x | a->out_phase_y) & ~1)) return fail(IDB_E_BADARG, "idb_gemm_conv: output phase must be 0 or 1");
  if (phased && (a->residual || (a->out_f32 && a->out_bf16) || a->k_splits > 1))
    return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: phased output with residual / both outputs / split-K");
  const bool phases4 = (a->flags & IDB_EPI_PHASES4) != 0;
  if (phases4 && (!phased || a->a0_mode != IDB_A_2X2 || a->n % 4 || (a->n / 4) % 32 || a->a1 || a->lora_down || a->rowvec || a->prelu ||
                  (a->flags & (IDB_EPI_GEGLU | IDB_EPI_GELU))))
    return fail(IDB_E_BADARG, "idb_gemm_conv: IDB_EPI_PHASES4 needs IDB_A_2X2, out_scale = 2, n = 4 * N_out and a plain bias epilogue");
  const bool stride2 = a->a0_mode == IDB_A_3X3_S2 || a->a0_mode == IDB_A_3X3_S2_ASYM;
  if (stride2 && ((a->height | a->width) & 1))
    return fail(IDB_E_BADARG, "idb_gemm_conv: stride-2 conv needs even H and W");
  if (!a->out_f32 && !a->out_bf16) return fail(IDB_E_BADARG, "idb_gemm_conv: no output");
  const bool lora = a->lora_down != nullptr;
  const bool geglu = (a->flags & IDB_EPI_GEGLU) != 0;
  if (lora && (!a->lora_up || a->lora_rank_pad <= 0 || a->lora_rank_pad > 16 || a->lora_rank_pad % 4 ||
               a->lora_seg_n <= 0 || a->lora_seg_n % 160 || a->n % a->lora_seg_n))
    return fail(IDB_E_BADARG, "idb_gemm_conv: bad LoRA arguments (rank_pad in {4,8,12,16}, seg_n % 160 == 0)");
  if (lora && (geglu || a->a1)) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: LoRA with GEGLU / second segment");
  if (geglu && a->out_f32) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: GEGLU writes bf16 only");
  if (a->stats_partials && !a->out_f32) return fail(IDB_E_BADARG, "idb_gemm_conv: stats_partials needs the fp32 output");
  if (a->stats_image_sums && !a->out_f32) return fail(IDB_E_BADARG, "idb_gemm_conv: stats_image_sums needs the fp32 output");
  if (a->prelu && geglu) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: PReLU with GEGLU");
  if ((a->flags & IDB_EPI_GELU) && (geglu || a->k_splits > 1)) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: GELU with GEGLU / forced split-K");
  if ((a->flags & IDB_EPI_F16) && lora) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: fp16 operands with fused LoRA");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int taps0 = (a->a0_mode == IDB_A_1X1) ? 1 : (a->a0_mode == IDB_A_2X2 ? 4 : 9);
  const int H = a->height, W = a->width, B = a->batch;
  p.mode0 = a->a0_mode;
  p.tap_off_x = a->tap_off_x;
  p.tap_off_y = a->tap_off_y;
  p.c0 = a->c0;
  p.cpb0 = a->c0 / 64;
  p.nkb0 = taps0 * p.cpb0;
  p.nkb1 = a->a1 ? a->c1 / 64 : 0;
  p.Ho = stride2 ? H / 2 : H;
  p.Wo = stride2 ? W / 2 : W;
  p.B = B;
  p.M = static_cast<long long>(B) * p.Ho * p.Wo;
  p.N = a->n;
  p.N_out = geglu ? a->n / 2 : (phases4 ? a->n / 4 : a->n);
  p.phases4 = phases4 ? 1 : 0;
  const long long k_total = static_cast<long long>(p.nkb0 + p.nkb1) * 64;

  // tile rectangle: 128 output pixels = BW x BH x BB
  p.BW = pow2_divisor(p.Wo, 128);
  if (p.Ho == 1 && B == 1) p.BW = 128;  // a Linear: rows are just rows; OOB handles the tail
  p.BH = pow2_divisor(p.Ho, 128 / p.BW);
  p.BB = 128 / (p.BW * p.BH);
  p.tiles_x = (p.Wo + p.BW - 1) / p.BW;
  p.tiles_y = (p.Ho + p.BH - 1) / p.BH;
  const int tiles_b = (B + p.BB - 1) / p.BB;
  p.n_tiles_m = p.tiles_x * p.tiles_y * tiles_b;
  if (a->stats_partials) {
    // 32-row blocks of the (BW x BH x BB) tile rectangle must be raster-contiguous runs of one image
    const bool raster = (p.BW == p.Wo) || (p.BH == 1 && p.BB == 1) || (p.Ho == 1 && B == 1);
    if (!raster || (static_cast<long long>(p.Ho) * p.Wo) % 32 != 0)
      return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: stats_partials needs power-of-two Wo (or Wo % 128 == 0) and Ho*Wo % 32 == 0");
  }

  // CTA pairs (cta_group::2) whenever there are at least two M blocks
  static const int force_cg = env_int("IDB_GEMM_CG", 0);
  int cg = (p.n_tiles_m >= 2) ? 2 : 1;
  if (force_cg == 1 || force_cg == 2) cg = force_cg;
  // N tile: minimise  waves x per-tile MMA time / L2-feed efficiency.  The operand feed from L2 caps
  // the tensor pipe at about 8 KB/clk chip-wide: eff = min(1, 8000 / (148 * bytes per clk per SM)).
  static const int max_sms = env_int("IDB_GEMM_MAXSM", 0);   // profiling only
  const int sms = (max_sms > 0 && max_sms < num_sms()) ? max_sms : num_sms();
  const int units = sms / cg;
  const int m_units = (p.n_tiles_m + cg - 1) / cg;
  int block_n = 0;
  bool dual = false;
  if (lora) {
    block_n = 160;
  } else {
    double best = -1.0;
    const int cands[3] = {256, 160, 128};
    for (int ci = 0; ci < 3; ++ci) {
      const int bn = cands[ci];
      if (a->n % bn || (phases4 && p.N_out % bn)) continue;   // (a tile never straddles two phases)
      const long long tiles = static_cast<long long>(m_units) * (a->n / bn);
      const double bytes_per_clk = (128.0 + bn / static_cast<double>(cg)) * 128.0 / (128.0 * bn / 4096.0);
      double eff = 8000.0 / (148.0 * bytes_per_clk);
      if (eff > 1.0) eff = 1.0;
      const double cost = static_cast<double>((tiles + units - 1) / units) * (bn + 24) / eff;
      if (best < 0 || cost < best) best = cost, block_n = bn;
    }
    if (block_n == 0) block_n = (a->n > 160) ? 256 : (a->n > 128 ? 160 : 128);
    if (phases4 && p.N_out % block_n) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: IDB_EPI_PHASES4 needs N_out % 128 == 0 or N_out % 160 == 0");
    static const int force_bn = env_int("IDB_GEMM_BN", 0);   // profiling only
    if ((force_bn == 128 || force_bn == 160 || force_bn == 256) && a->n % force_bn == 0) block_n = force_bn;
    // dual-N (2 x 160 columns per tile, two accumulators sharing every A tile): fewer operand bytes per MAC for the
    // big-K layers, as long as the halved tile count still fills the machine
    static const int force_dual = env_int("IDB_GEMM_DUAL", -1);   // profiling only: 0 = never, 1 = whenever legal
    const bool dual_ok = cg == 2 && !geglu && !(a->flags & IDB_EPI_GELU) && a->n % 320 == 0 && (!phases4 || p.N_out % 320 == 0) && force_bn == 0;
    if (dual_ok && force_dual != 0) {
      const long long tiles = static_cast<long long>(m_units) * (a->n / 320);
      const double cost = static_cast<double>((tiles + units - 1) / units) * (320 + 24) * (128.0 + 160.0) / 320.0;
      const double feed = (128.0 + block_n / 2.0) / block_n;
      const long long tiles1 = static_cast<long long>(m_units) * ((a->n + block_n - 1) / block_n);
      const double cost1 = static_cast<double>((tiles1 + units - 1) / units) * (block_n + 24) * feed;
      // tiny-M layers (8x8 latents) run split-K anyway: with dual-N tiles each split streams the weights once per
      // 256 rows instead of once per 128-column tile
      const bool splitk_regime = !phased && a->k_splits == 0 && a->workspace && tiles1 * 2 <= units && p.nkb0 + p.nkb1 >= 64;   // (split anyway)
      static const int dual_min_kb = env_int("IDB_GEMM_DUAL_MINKB", 32);   // smallest K (in 64-wide blocks) that takes dual-N tiles
      // experiment (off by default): one-wave layers whose dual-N tile grid would leave half the machine idle (16x16 latents,
      // N = 1280) take dual-N tiles AND two K splits -- the same number of work units, 31 % fewer operand bytes per MAC,
      // at the price of the fp32 workspace round trip.  IDB_GEMM_DUAL_SPLIT = smallest K (in 64-wide blocks) that qualifies.
      static const int dual_split_kb = env_int("IDB_GEMM_DUAL_SPLIT", 0);
      const bool dual_split = dual_split_kb > 0 && !phased && a->k_splits == 0 && a->workspace && tiles * 2 <= units &&
                              tiles1 * 2 > units && p.nkb0 + p.nkb1 >= dual_split_kb &&
                              static_cast<size_t>(p.M) * a->n * sizeof(float) * 2 <= a->workspace_bytes;
      if (force_dual == 1 || (p.nkb0 + p.nkb1 >= dual_min_kb && cost < cost1) || splitk_regime || dual_split) dual = true, block_n = 160;
    }
  }
  // stream-K over the dual-N tiles of each image (see WorkIter).  IDB_GEMM_SK: 0 = off, 1 = the layers whose dual-N grid
  // leaves at least half of the machine idle (16x16 latents at the bench batch; default), 2 = every eligible layer.
  const int nkb = p.nkb0 + p.nkb1;
  bool sk = false;
  p.sk_groups = p.sk_units = p.sk_tiles = 0;
  {
    static const int sk_mode = env_int("IDB_GEMM_SK", 1);
    static const int sk_min_kb = env_int("IDB_GEMM_SK_MINKB", 32);   // shortest K range (k-blocks) worth a pair
    static const int dbg0 = env_int("IDB_GEMM_DEBUG", 0);
    static const int no_spec0 = env_int("IDB_GEMM_NOSPEC", 0);
    const int blocks_per_img = p.tiles_x * p.tiles_y;   // 128-row blocks of one image
    const bool shape_ok = !lora && cg == 2 && !geglu && !phased && !(a->flags & (IDB_EPI_GELU | IDB_EPI_F16)) && a->n % 320 == 0 &&
                          a->k_splits == 0 && a->workspace && a->out_f32 && !a->out_bf16 && dbg0 == 0 && !no_spec0 && max_sms == 0 &&
                          p.BB == 1 && blocks_per_img % 2 == 0 && units == num_sms() / 2;
    if (sk_mode > 0 && shape_ok && p.Ho > 1 && g_sk_launch_mode.load(std::memory_order_relaxed) < 2) {   // (rasters only: a group is an image)
      const int groups = B;                                               // one group per image
      const int tiles_per_group = (blocks_per_img / 2) * (a->n / 320);
      const long long tiles_dual = static_cast<long long>(groups) * tiles_per_group;
      int P = groups <= units ? units / groups : 0;
      static const int force_p = env_int("IDB_GEMM_SK_P", 0);   // profiling only
      if (force_p > 0 && force_p < P) P = force_p;
      const long long W = static_cast<long long>(tiles_per_group) * nkb;
      if (P > W / sk_min_kb) P = static_cast<int>(W / sk_min_kb);
      const size_t need = static_cast<size_t>(units) * 2 * 128 * 320 * sizeof(float);
      const bool fills = P >= 1 && static_cast<long long>(groups) * P * 10 >= static_cast<long long>(units) * 9;   // >= 90 % of the pairs busy
      const bool regime = sk_mode >= 2 || tiles_dual * 2 <= units;
      if (fills && regime && P > 1 && tiles_dual <= SK_MAX_TILES && W < (1ll << 30) && need <= a->workspace_bytes && nkb >= 2 * SK_MINSEG) {
        sk = true, dual = true, block_n = 160;
        p.sk_groups = groups, p.sk_units = P, p.sk_tiles = tiles_per_group;
      }
    }
  }
  p.n_tiles_n = (a->n + block_n * (dual ? 2 : 1) - 1) / (block_n * (dual ? 2 : 1));
  p.tiles_per_phase = phases4 ? p.N_out / (block_n * (dual ? 2 : 1)) : p.n_tiles_n;

  int ksp = a->k_splits;
  if (sk) ksp = 1;
  if (ksp == 0) {  // auto: split K when the tile grid leaves most SMs idle
    ksp = 1;
    const long long tiles = static_cast<long long>(m_units) * p.n_tiles_n;
    if (!lora && !geglu && !phased && !(a->flags & IDB_EPI_GELU) && a->workspace && tiles * 2 <= units && nkb >= 16) {
      long long want = units / tiles;
      if (want > nkb / 8) want = nkb / 8;
      const size_t per_split = static_cast<size_t>(p.M) * a->n * sizeof(float);
      if (per_split * want > a->workspace_bytes) want = static_cast<long long>(a->workspace_bytes / per_split);
      if (want > 1) ksp = static_cast<int>(want);
    }
  }
  if (ksp < 1) ksp = 1;
  if (ksp > 1 && (lora || geglu)) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: split-K with LoRA / GEGLU");
  if (ksp > nkb) ksp = nkb;
  p.kb_per_split = (nkb + ksp - 1) / ksp;
  p.k_splits = (nkb + p.kb_per_split - 1) / p.kb_per_split;
  if (p.k_splits > 1 && (!a->workspace || (a->k_splits > 1 && a->workspace_bytes != 0 &&
                                           idb_gemm_conv_workspace_bytes(p.M, a->n, p.k_splits) > a->workspace_bytes)))
    return fail(IDB_E_BADARG, "idb_gemm_conv: split-K needs a (large enough) workspace");

  p.fd_ntn.set(p.n_tiles_n);
  p.fd_ks.set(p.k_splits);
  p.fd_tx.set(p.tiles_x);
  p.fd_ty.set(p.tiles_y);
  p.fd_seg.set(lora ? a->lora_seg_n : 1);
  if (static_cast<long long>(m_units) * p.n_tiles_n * p.k_splits >= (1ll << 20) || p.n_tiles_m >= (1 << 20))
    return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: more than 2^20 tiles");

  p.bias = a->bias;
  p.rowvec = a->rowvec;
  p.rowvec_ld = a->rowvec_ld > 0 ? a->rowvec_ld : a->n;
  p.residual = a->residual;
  p.prelu = a->prelu;
  p.lora_seg_n = lora ? a->lora_seg_n : 1;
  p.flags = a->flags;
  static const int dbg = env_int("IDB_GEMM_DEBUG", 0);
  p.debug = dbg;
  p.out_f32 = a->out_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16);
  p.workspace = a->workspace;
  p.stats = reinterpret_cast<float2*>(a->stats_partials);
  p.stats_rowblock0 = (phased && !phases4) ? static_cast<long long>(2 * a->out_phase_y + a->out_phase_x) * ((p.M + 31) / 32) : 0;
  p.n_rowblocks = (p.M + 31) / 32;
  int stats_hw = 0, n_img = 0;
  if (a->stats_image_sums) {
    // per-image channel sums: the call's rows are n_img images of stats_hw rows each (a Linear over tokens passes the
    // token count per image in stats_hw; a conv uses its output raster)
    stats_hw = a->stats_hw > 0 ? a->stats_hw : p.Ho * p.Wo;
    if (stats_hw % 32 || p.M % stats_hw) return fail(IDB_E_BADARG, "idb_gemm_conv: stats_hw must be a multiple of 32 that divides M");
    const bool raster = (p.BW == p.Wo) || (p.BH == 1 && p.BB == 1) || (p.Ho == 1 && B == 1);
    if (!raster || (static_cast<long long>(p.Ho) * p.Wo) % 32 != 0)
      return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: stats_image_sums needs power-of-two Wo (or Wo % 128 == 0) and Ho*Wo % 32 == 0");
    const int gran = a->stats_gran > 0 ? a->stats_gran : 1;
    if (p.N_out % gran) return fail(IDB_E_BADARG, "idb_gemm_conv: stats_gran must divide N_out");
    n_img = static_cast<int>(p.M / stats_hw);
    p.fd_rbpi.set(stats_hw / 32);
    p.fd_gran.set(gran);
    p.n_gran = p.N_out / gran;
    p.isums = reinterpret_cast<unsigned long long*>(a->stats_image_sums);
  }

  // ---- tensor maps
  if (stride2) {
    const uint64_t C = a->c0;
    uint64_t dims[5] = {2 * C, uint64_t(W / 2), 2, uint64_t(H / 2), uint64_t(B)};
    uint64_t strides[4] = {2 * C * 2, uint64_t(W) * C * 2, 2 * uint64_t(W) * C * 2, uint64_t(H) * W * C * 2};
    uint32_t box[5] = {64, uint32_t(p.BW), 1, uint32_t(p.BH), uint32_t(p.BB)};
    if (int rc = make_tmap_bf16(&p.tmA0, a->a0, 5, dims, strides, box)) return rc;
  } else {
    const uint64_t C = a->c0;
    uint64_t dims[4] = {C, uint64_t(W), uint64_t(H), uint64_t(B)};
    uint64_t strides[3] = {C * 2, uint64_t(W) * C * 2, uint64_t(H) * W * C * 2};
    uint32_t box[4] = {64, uint32_t(p.BW), uint32_t(p.BH), uint32_t(p.BB)};
    if (int rc = make_tmap_bf16(&p.tmA0, a->a0, 4, dims, strides, box)) return rc;
  }
  if (a->a1) {
    const uint64_t C = a->c1;
    uint64_t dims[4] = {C, uint64_t(p.Wo), uint64_t(p.Ho), uint64_t(B)};
    uint64_t strides[3] = {C * 2, uint64_t(p.Wo) * C * 2, uint64_t(p.Ho) * p.Wo * C * 2};
    uint32_t box[4] = {64, uint32_t(p.BW), uint32_t(p.BH), uint32_t(p.BB)};
    if (int rc = make_tmap_bf16(&p.tmA1, a->a1, 4, dims, strides, box)) return rc;
  }
  if (lora) {
    const int nseg = a->n / a->lora_seg_n;
    uint64_t dims[2] = {uint64_t(k_total), uint64_t(nseg * 16)};
    uint64_t strides[1] = {uint64_t(k_total) * 2};
    uint32_t box[2] = {64, 16};
    if (int rc = make_tmap_bf16(&p.tmL, a->lora_down, 2, dims, strides, box)) return rc;
    // up-projection operand [N, 64] bf16: each CTA stages the rows of its share of the N tile
    uint64_t udims[2] = {64, uint64_t(a->n)};
    uint64_t ustrides[1] = {64 * 2};
    uint32_t ubox[2] = {64, uint32_t(block_n / cg)};
    if (int rc = make_tmap_bf16(&p.tmU, a->lora_up, 2, udims, ustrides, ubox)) return rc;
  }

  // W box: in a pair each CTA stages half of the N tile
  {
    const long long k_tot = static_cast<long long>(p.nkb0 + p.nkb1) * 64;
    uint64_t dims[2] = {uint64_t(k_tot), uint64_t(a->n)};
    uint64_t strides[1] = {uint64_t(k_tot) * 2};
    uint32_t box[2] = {64, uint32_t(block_n / cg)};
    if (lora && cg == 2) {
      // the pair's B tile is [160 base rows | 16 LoRA-down rows]: CTA 0 stages base rows 0-87, CTA 1 rows 88-159 + LoRA
      box[1] = (block_n + 16) / 2;
      uint32_t box1[2] = {64, uint32_t(block_n - (block_n + 16) / 2)};
      if (int rc = make_tmap_bf16(&p.tmW1, a->w, 2, dims, strides, box1)) return rc;
    }
    if (int rc = make_tmap_bf16(&p.tmW, a->w, 2, dims, strides, box)) return rc;
  }
  // output store maps: 4-D [N_out, Wo, Ho, B], one box = the 32 tile rows a warp owns x 32 (16 for GEGLU) columns
  {
    const int bw32 = p.BW < 32 ? p.BW : 32;
    const int bh32 = p.BH < 32 / bw32 ? p.BH : 32 / bw32;
    const int bb32 = 32 / (bw32 * bh32);
    const uint32_t nc = geglu ? 16 : 32;
    const uint64_t No = uint64_t(p.N_out);
    uint64_t dims[4] = {No, uint64_t(p.Wo), uint64_t(p.Ho), uint64_t(B)};
    uint32_t box[4] = {nc, uint32_t(bw32), uint32_t(bh32), uint32_t(bb32)};
    // phased output: pixel (y, x) lands at (2y + py, 2x + px) of the [B, 2Ho, 2Wo, N] tensor -> same box, doubled
    // pixel / row strides and a shifted base
    const uint64_t sx = phased ? 2 : 1;
    const uint64_t row_elems = sx * uint64_t(p.Wo) * No;                    // elements per output row of the target tensor
    const uint64_t base_off = (phased && !phases4) ? (uint64_t(a->out_phase_y) * row_elems + uint64_t(a->out_phase_x) * No) : 0;
    if (phases4) {   // phases 1-3 (phase = 2 py + px): the same box and strides from a shifted base
      for (int ph = 1; ph < 4; ++ph) {
        const uint64_t off = uint64_t(ph >> 1) * row_elems + uint64_t(ph & 1) * No;
        if (a->out_f32) {
          uint64_t strides[3] = {sx * No * 4, sx * row_elems * 4, sx * uint64_t(p.Ho) * row_elems * 4};
          if (int rc = make_tmap(&p.tmOutP[ph - 1], a->out_f32 + off, 4, 128, 4, dims, strides, box)) return rc;
        } else {
          uint64_t strides[3] = {sx * No * 2, sx * row_elems * 2, sx * uint64_t(p.Ho) * row_elems * 2};
          if (int rc = make_tmap(&p.tmOutP[ph - 1], static_cast<const char*>(a->out_bf16) + off * 2, 2, 0, 4, dims, strides, box)) return rc;
        }
      }
    }
    static const int ws_tma_on = env_int("IDB_GEMM_WS_TMA", 1);
    p.ws_tma = 0;
    if (p.k_splits > 1 && ws_tma_on && !geglu && p.N_out == p.N && B % bb32 == 0) {
      // split-K partials: slice ks of the workspace is a [B, Ho, Wo, N] fp32 tensor; the slices are stacked on the batch axis
      uint64_t wdims[4] = {No, uint64_t(p.Wo), uint64_t(p.Ho), uint64_t(B) * uint64_t(p.k_splits)};
      uint64_t strides[3] = {No * 4, uint64_t(p.Wo) * No * 4, uint64_t(p.Ho) * p.Wo * No * 4};
      if (int rc = make_tmap(&p.tmOutF, a->workspace, 4, 128, 4, wdims, strides, box)) return rc;
      p.ws_tma = 1;
    } else if (a->out_f32) {
      uint64_t strides[3] = {sx * No * 4, sx * row_elems * 4, sx * uint64_t(p.Ho) * row_elems * 4};
      if (int rc = make_tmap(&p.tmOutF, a->out_f32 + base_off, 4, geglu ? 0 : 128, 4, dims, strides, box)) return rc;
    }
    if (a->out_bf16) {
      uint64_t strides[3] = {sx * No * 2, sx * row_elems * 2, sx * uint64_t(p.Ho) * row_elems * 2};
      if (int rc = make_tmap(&p.tmOutB, static_cast<const char*>(a->out_bf16) + base_off * 2, 2, 0, 4, dims, strides, box)) return rc;
    }
    if (a->residual && !geglu) {   // fp32 [M, N_out] like out_f32: prefetched by TMA into the staging buffers
      uint64_t strides[3] = {No * 4, uint64_t(p.Wo) * No * 4, uint64_t(p.Ho) * p.Wo * No * 4};
      if (int rc = make_tmap(&p.tmRes, a->residual, 4, 128, 4, dims, strides, box)) return rc;
    }
  }
  GemmParams pk = p;
  if (p.k_splits > 1) pk.stats = nullptr, pk.isums = nullptr;   // statistics come from the finalize pass
  const int total_tiles = m_units * p.n_tiles_n * p.k_splits;
  const int grid = sk ? cg * units : cg * (total_tiles < units ? total_tiles : units);
  // epilogue specialisation (generic whenever a profiling switch, split-K or both outputs are in play)
  int epi = 0;
  static const int no_spec = env_int("IDB_GEMM_NOSPEC", 0);
  if (!no_spec && dbg == 0 && p.k_splits == 1 && !(a->out_f32 && a->out_bf16) && !(a->flags & IDB_EPI_GELU)) {
    if (a->out_f32) epi = 3;
    else if (geglu) epi = (a->residual || a->rowvec || a->prelu) ? 0 : 2;
    else epi = (a->rowvec || a->stats_partials) ? 0 : 1;
  }
  int rc;
  if (lora && cg == 2) rc = launch_gemm<160, 4, true, 2>(pk, grid, stream, epi);
  else if (lora) rc = launch_gemm<160, 2, true, 1>(pk, grid, stream, epi);
  else if (cg == 1 && block_n == 256) rc = launch_gemm<256, 3, false, 1>(pk, grid, stream, epi);
  else if (cg == 1 && block_n == 160) rc = launch_gemm<160, 4, false, 1>(pk, grid, stream, epi);
  else if (cg == 1) rc = launch_gemm<128, 4, false, 1>(pk, grid, stream, epi);
  else if (sk && epi == 3) {
    rc = launch_gemm_e<160, 4, false, 2, 3, 2, true>(pk, grid, stream);
    if (rc == IDB_E_UNSUPPORTED) return idb_gemm_conv(a, stream_);   // cooperative launch refused: plan again without stream-K
    if (rc == IDB_OK) note_stream_k_launch();
  }
  else if (sk) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: internal: stream-K schedule chosen for a non-fp32 epilogue");
  else if (dual && epi == 1) rc = launch_gemm_e<160, 4, false, 2, 1, 2>(pk, grid, stream);
  else if (dual && epi == 3) rc = launch_gemm_e<160, 4, false, 2, 3, 2>(pk, grid, stream);
  else if (dual) rc = launch_gemm_e<160, 4, false, 2, 0, 2>(pk, grid, stream);
  else if (block_n == 256) rc = launch_gemm<256, 4, false, 2>(pk, grid, stream, epi);
  else if (block_n == 160) rc = launch_gemm<160, 6, false, 2>(pk, grid, stream, epi);
  else rc = launch_gemm<128, 6, false, 2>(pk, grid, stream, epi);
  if (rc) return rc;

  if (p.k_splits > 1 && p.isums != nullptr && !p.stats && !(p.flags & IDB_EPI_F16) && p.out_f32 && !p.out_bf16) {
    // finalize + per-image channel sums in one pass
    launch_pdl(splitk_finalize_sums_kernel, dim3(p.N / 32, n_img), dim3(256), 0, stream, p.workspace, p.k_splits, p.M, p.N, p.Ho * p.Wo,
               stats_hw, p.bias, p.rowvec, p.rowvec_ld, p.prelu, p.residual, p.out_f32, p.isums, p.fd_gran.d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("splitk_finalize_sums launch: ") + cudaGetErrorString(e));
  } else if (p.k_splits > 1) {
    if (p.isums != nullptr) return fail(IDB_E_UNSUPPORTED, "idb_gemm_conv: split-K with stats_image_sums needs a single fp32 output and no stats_partials");
    const long long total4 = p.M * p.N / 4;
    int blocks = static_cast<int>((total4 + 255) / 256);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    launch_pdl(splitk_finalize_kernel, dim3(blocks), dim3(256), 0, stream, p.workspace, p.k_splits, p.M, p.N, p.Ho * p.Wo, p.bias, p.rowvec,
                                                       p.rowvec_ld, p.prelu, p.residual, p.out_f32, p.out_bf16, (p.flags & IDB_EPI_F16) ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("splitk_finalize launch: ") + cudaGetErrorString(e));
    if (p.stats != nullptr) {
      const unsigned nrb = static_cast<unsigned>((p.M + 31) / 32);
      launch_pdl(rowblock_stats_kernel, dim3(dim3(nrb, (p.N + 127) / 128)), dim3(128), 0, stream, p.out_f32, p.M, p.N, p.stats);
      e = cudaGetLastError();
      if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("rowblock_stats launch: ") + cudaGetErrorString(e));
    }
  }
  return IDB_OK;
}
