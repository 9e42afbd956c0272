// idb_attention: flash-style softmax(Q K^T * scale) V for head_dim 64 on tcgen05 (sm_100a).
//
// One CTA = one (image, head, 128-query tile); 2 CTAs co-reside per SM (96 KiB smem, 256 TMEM
// columns each) so one CTA's softmax (MUFU/FMA pipes) overlaps the other's MMAs (tensor pipe).
//   warps 0-3 : softmax -- thread r owns query row r: S row from TMEM (tcgen05.ld), online max /
//               exp2 / sum in fp32, P (bf16) written to smem in the UMMA K-major SWIZZLE_128B
//               layout, lazy rescale of the O accumulator in TMEM, final O / l -> global
//   warp  4   : TMA producer -- Q once, then K_j, V_j tiles through a 3-slot ring
//   warp  5   : tcgen05.mma issuer: S = Q K_j^T (M128 N128 K64), O += P V_j (M128 N64 K128,
//               V consumed MN-major straight from its [token, d] row-major layout)
#include <cstdlib>
#include <string>

#include "../../include/idb.h"
#include "idb_common.cuh"
#include "idb_host.h"

namespace idb {

constexpr int ATT_BM = 128;      // query rows per CTA
constexpr int ATT_BN = 128;      // kv rows per tile
constexpr int ATT_D = 64;
constexpr int ATT_TILE = 16384;  // 128 x 64 bf16
constexpr int ATT_RING = 3;
constexpr int ATT_THREADS = 192;
constexpr int ATT_TMEM_COLS = 256;
constexpr int ATT_S_COL = 0;
constexpr int ATT_O_COL = 128;
constexpr int ATT_SMEM = ATT_TILE /*Q*/ + ATT_RING * ATT_TILE /*K,V ring*/ + 2 * ATT_TILE /*P*/ + 1024 + 128;

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  int col0_q, col0_k, col0_v;
  __nv_bfloat16* out;
  long long ld_out;
  int B, heads, Tq, Tkv, n_kv_tiles;
  float scale_log2;
  int causal;
  float* lse;   // optional [B, heads, Tq]: log2-domain log-sum-exp of the scaled scores (training forward)
};

__global__ void __launch_bounds__(ATT_THREADS, 2) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sRing = sQ + ATT_TILE;
  uint8_t* sP = sRing + ATT_RING * ATT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + ATT_RING;
  uint64_t* s_full = kv_empty + ATT_RING;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_RING; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);  // one arrive per softmax warp
    mbar_init(o_done, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // q / k / v come from the preceding kernels

  if (warp == 4) {
    // ================================================================ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE);
      tma_load_3d(sQ, &p.tmQ, q_full, p.col0_q + head * ATT_D, q0, b);
      for (int i = 0; i < 2 * n_tiles; ++i) {
        const int slot = i % ATT_RING;
        const uint32_t ph = (i / ATT_RING) & 1;
        mbar_wait(&kv_empty[slot], ph ^ 1);
        mbar_expect_tx(&kv_full[slot], ATT_TILE);
        const int j = i >> 1;
        if ((i & 1) == 0)
          tma_load_3d(sRing + slot * ATT_TILE, &p.tmK, &kv_full[slot], p.col0_k + head * ATT_D, j * ATT_BN, b);
        else
          tma_load_3d(sRing + slot * ATT_TILE, &p.tmV, &kv_full[slot], p.col0_v + head * ATT_D, j * ATT_BN, b);
      }
    }
  } else if (warp == 5) {
    // ================================================================ MMA issuer
    constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
    constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BM, ATT_D, 0, 1);  // B (= V) is MN-major
    const uint32_t tS = tmem_base + ATT_S_COL;
    const uint32_t tO = tmem_base + ATT_O_COL;
    const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ));
    const uint64_t pdesc = umma_smem_desc_sw128(smem_u32(sP));

    auto issue_qk = [&](int j) {
      const int i = 2 * j, slot = i % ATT_RING;
      mbar_wait(&kv_full[slot], (i / ATT_RING) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sRing + slot * ATT_TILE));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_bf16(tS, qdesc + 2 * k, kdesc + 2 * k, IDESC_S, k > 0 ? 1u : 0u);
        umma_commit(&kv_empty[slot]);
        umma_commit(s_full);
      }
      __syncwarp();
    };

    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(p_full, j & 1);  // P_j in smem, S_j consumed, O rescaled
      tc_fence_after();
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int i = 2 * j + 1, slot = i % ATT_RING;
      mbar_wait(&kv_full[slot], (i / ATT_RING) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t vbase = smem_u32(sRing + slot * ATT_TILE);
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          // A = P: K-major, two 64-wide swizzle atoms of 16 KiB; 32 B per K=16 step inside an atom
          const uint64_t ad = pdesc + static_cast<uint64_t>(((kk >> 2) * ATT_TILE + (kk & 3) * 32) >> 4);
          // B = V: MN-major, 16 kv rows (2 KiB) per K=16 step
          const uint64_t bd = umma_smem_desc_sw128(vbase + kk * 2048);
          umma_bf16(tO, ad, bd, IDESC_O, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[slot]);
        umma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    // ================================================================ softmax warps (thread = query row)
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + ATT_S_COL;
    const uint32_t tO = tmem_base + lane_off + ATT_O_COL;
    const float c = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv_valid = min(ATT_BN, p.Tkv - j * ATT_BN);  // columns < kv_valid are real keys
      // ---- pass 1: row max
      float m_tile = -INFINITY;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        IDB_TMEM_LD_X32(tS + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (ch * 32 + i < kv_valid) ? __uint_as_float(v[i]) : -INFINITY;
          m_tile = fmaxf(m_tile, s);
        }
      }
      const float m_new = fmaxf(m_run, m_tile * c);
      const float alpha = ex2(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      m_run = m_new;
      l_run *= alpha;
      // ---- O rescale (needs PV_{j-1} retired; that also frees the P buffer)
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t v[32];
            IDB_TMEM_LD_X32(tO + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            IDB_TMEM_ST_X32(tO + ch * 32, v);
          }
          tmem_st_wait();
        }
      }
      // ---- pass 2: P = exp2(S*c - m), row sum, bf16 pack into the swizzled K-major smem tile
      float l_tile = 0.f;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        IDB_TMEM_LD_X32(tS + ch * 32, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = ex2(__uint_as_float(v[i]) * c - m_new);
          pv[i] = (ch * 32 + i < kv_valid) ? e : 0.f;
          l_tile += pv[i];
        }
        // columns ch*32 .. +31 -> k-atom (ch>>1), 16-byte chunks ((ch&1)*4 + q), q = 0..3
        uint8_t* base = prow + (ch >> 1) * ATT_TILE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (ch & 1) * 4 + q;
          uint4 w = make_uint4(pack_bf16x2(pv[8 * q], pv[8 * q + 1]), pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]),
                               pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]), pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]));
          *reinterpret_cast<uint4*>(base + ((chunk ^ sw) << 4)) = w;
        }
      }
      l_run += l_tile;
      fence_proxy_async_smem();  // P visible to the tensor core's async proxy
      tc_fence_before();         // order our tcgen05.ld/st before the MMA warp's next tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l
    mbar_wait(o_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int row = q0 + r;
    if (p.lse != nullptr && row < p.Tq) p.lse[(static_cast<long long>(b) * p.heads + head) * p.Tq + row] = m_run + log2f(l_run);
    __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Tq + row) * p.ld_out + head * ATT_D;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      IDB_TMEM_LD_X32(tO + ch * 32, v);
      tmem_ld_wait();
      if (row < p.Tq) {
        uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l),
                              pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l),
                              pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l),
                              pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------- shared constants of the 256-query kernels
constexpr int AT2_RING = 4;
constexpr int AT2_TMEM_COLS = 512;
constexpr float AT2_RESCALE_THRESHOLD = 8.0f;   // lazy rescale: the exponent reference moves only when P would exceed 2^8

// ---------------------------------------------------------------------------------------- row-split 256-query variant
// One CTA per SM handles TWO 128-row query tiles ("lanes" a / b) that share every K_j / V_j tile: the TMA load traffic
// per FLOP is half that of two independent 128-row CTAs, and while one lane's warps run the softmax of tile j the
// tensor core works for the other lane.  Each query row's 128 scores are split between TWO threads (64 keys each, warps w and w + 4 of the same TMEM lane
// quarter), i.e. 16 softmax warps = 4 per SM sub-partition.  With one row per thread the SM sub-partitions saw
// ~1 runnable warp (the other lane waits for its MMAs) and the 128 live scores spilled; here every thread keeps
// 64 scores in registers, the pair exchanges its partial row maxima through smem (one 64-thread named barrier
// per tile), row sums stay partial until the end, and each thread rescales / writes 32 of the 64 O columns.
constexpr int AT3_THREADS = 576;   // warps 0-7: softmax lane a, 8-15: lane b, 16: TMA, 17: MMA
constexpr int AT3_XCHG = 2 * 2 * 128 * 2 * 4;   // [parity][lane x][row][half] partial maxima
constexpr int AT3_SMEM = 2 * ATT_TILE /*Q*/ + AT2_RING * ATT_TILE + 4 * ATT_TILE /*P a,b*/ + 1024 + 256 + 2 * AT3_XCHG;

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// two exp2 by the FMA/ALU pipes (Cody-Waite + cubic, packed fp32x2 where the ISA allows)
__device__ __forceinline__ void exp2_poly2(float& y0, float& y1, float x0, float x1) {
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const float magic = 12582912.0f;
  float xf0, xf1, n0, n1, f0, f1, p0, p1;
  fadd2(xf0, xf1, x0, x1, magic, magic);
  fadd2(n0, n1, xf0, xf1, -magic, -magic);
  fadd2(f0, f1, x0, x1, -n0, -n1);
  ffma2(p0, p1, f0, f1, 0.05550410866f, 0.05550410866f, 0.24022650696f, 0.24022650696f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.69314718056f, 0.69314718056f);
  ffma2(p0, p1, p0, p1, f0, f1, 1.0f, 1.0f);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(xf0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(xf1) << 23));
}

// SPLIT: the two threads of a query row are independent online-softmax instances over disjoint key halves of every tile
// (keys 0-63 / 64-127), each with its OWN exponent reference, row sum and 64-column O accumulator (O_x0 / O_x1 in TMEM:
// S_a | S_b | O_a0 | O_a1 | O_b0 | O_b1 = 512 columns); the two partial results are merged once per item
// (O = (O_0 w_0 + O_1 w_1) / (l_0 w_0 + l_1 w_1), w_h = 2^(m_h - max m)).  A tile is then read from TMEM ONCE, 32 scores
// at a time with the lazy reference update done per 32-score chunk, and there is no per-tile exchange / named barrier
// between the two threads of a row.
template <int POLY_MASK, bool SPLIT>   // pairs (i, i+1) with ((i >> 1) & POLY_MASK) == POLY_MASK take the polynomial exp2: 1 -> 1/2, 3 -> 1/4, 7 -> 1/8, 32 -> none
__global__ void __launch_bounds__(AT3_THREADS, 1) attention_rs_kernel(const __grid_constant__ AttnParams p, int n_full, int qblocks, int n_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [2][128 x 64]
  uint8_t* sRing = sQ + 2 * ATT_TILE;
  uint8_t* sP = sRing + AT2_RING * ATT_TILE;            // [2][2 k-atoms][128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * ATT_TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + AT2_RING;
  uint64_t* s_full = kv_empty + AT2_RING;   // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* o_done = p_full + 2;            // [2]
  uint64_t* s_free = o_done + 2;            // [2] every softmax warp of the lane has its last S values in registers
  uint64_t* q_empty = s_free + 2;           // the item's last Q K^T has retired: the Q tiles may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);
  float* xch_m = reinterpret_cast<float*>(sP + 4 * ATT_TILE + 1024 + 256);   // [parity][x][row][half]
  float* xch_l = xch_m + AT3_XCHG / 4;                                         // [x][row][half] (final row sums)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Persistent CTAs walk a list of work items (item = blockIdx.x, + gridDim.x, ...): items [0, n_full) are whole
  // 256-query units; the units that would form a partial last round are split into two single-lane items each
  // (128 queries, lane b idle), which take ~0.6 of a full item's time.  TMEM, barriers and the K/V ring live across
  // items, so the launch / allocation / first-load latency is paid once per CTA instead of once per unit.
  const int n_tiles = p.n_kv_tiles;
  struct Item { bool single; int nl, head, b, q0; };
  auto decode = [&](int item) {
    Item t;
    t.single = item >= n_full;
    const int unit = t.single ? n_full + ((item - n_full) >> 1) : item;
    t.nl = t.single ? 1 : 2;
    const int qblk = unit % qblocks;
    t.head = (unit / qblocks) % p.heads;
    t.b = unit / (qblocks * p.heads);
    t.q0 = qblk * 2 * ATT_BM + (t.single ? ((item - n_full) & 1) * ATT_BM : 0);
    return t;
  };

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT2_RING; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&p_full[x], 8);
      mbar_init(&o_done[x], 1);
      mbar_init(&s_free[x], 8);
    }
    mbar_init(q_empty, 1);
    mbar_fence_init();
  }
  if (warp == 17) tmem_alloc<AT2_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // q / k / v come from the preceding kernels
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 16) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int gi = 0;   // K / V tiles loaded so far (ring position and phase)
      for (int item = blockIdx.x, it = 0; item < n_items; item += gridDim.x, ++it) {
        const Item t = decode(item);
        if (it > 0) mbar_wait(q_empty, (it - 1) & 1);
        mbar_expect_tx(q_full, t.nl * ATT_TILE);
        tma_load_3d(sQ, &p.tmQ, q_full, p.col0_q + t.head * ATT_D, t.q0, t.b);
        if (!t.single) tma_load_3d(sQ + ATT_TILE, &p.tmQ, q_full, p.col0_q + t.head * ATT_D, t.q0 + ATT_BM, t.b);
        for (int i = 0; i < 2 * n_tiles; ++i, ++gi) {
          const int slot = gi % AT2_RING;
          const uint32_t ph = (gi / AT2_RING) & 1;
          mbar_wait(&kv_empty[slot], ph ^ 1);
          mbar_expect_tx(&kv_full[slot], ATT_TILE);
          const int j = i >> 1;
          if ((i & 1) == 0)
            tma_load_3d(sRing + slot * ATT_TILE, &p.tmK, &kv_full[slot], p.col0_k + t.head * ATT_D, j * ATT_BN, t.b);
          else
            tma_load_3d(sRing + slot * ATT_TILE, &p.tmV, &kv_full[slot], p.col0_v + t.head * ATT_D, j * ATT_BN, t.b);
        }
      }
    }
  } else if (warp == 17) {
    // ================================================================ MMA issuer (convergent loop, elected issue)
    constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
    constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BM, ATT_D, 0, 1);  // B (= V) is MN-major
    const uint64_t desc_hi = umma_smem_desc_sw128(0);
    const uint32_t bar0 = smem_base + 2 * ATT_TILE + AT2_RING * ATT_TILE + 4 * ATT_TILE;   // &bars[0]
    auto bar_addr = [&](int idx) { return bar0 + idx * 8; };
    const int I_KVE = 1 + AT2_RING, I_SF = 1 + 2 * AT2_RING, I_OD = I_SF + 4;
    auto mk = [&](uint32_t addr) { return desc_hi | static_cast<uint64_t>((addr >> 4) & 0x3FFF); };

    int gbase = 0;          // K / V tiles consumed by previous items
    int nl = 2;             // active lanes of the current item
    uint32_t gj[2] = {0, 0};   // tiles processed so far per lane (barrier phases)
    auto issue_qk = [&](int x, int j) {   // S_x = Q_x K_j^T
      const int i = gbase + 2 * j, slot = i % AT2_RING;
      if (x == 0) {
        mbar_wait(&kv_full[slot], (i / AT2_RING) & 1);
        tc_fence_after();
      }
      const uint64_t qd = mk(smem_base + x * ATT_TILE);
      const uint64_t kd = mk(smem_base + 2 * ATT_TILE + slot * ATT_TILE);
      const uint32_t tS = tmem_base + x * 128;
      if (elect_one()) {
        umma_bf16(tS, qd, kd, IDESC_S, 0u);
        umma_bf16(tS, qd + 2, kd + 2, IDESC_S, 1u);
        umma_bf16(tS, qd + 4, kd + 4, IDESC_S, 1u);
        umma_bf16(tS, qd + 6, kd + 6, IDESC_S, 1u);
        umma_commit_a(bar_addr(I_SF + x));
        if (x == nl - 1) umma_commit_a(bar_addr(I_KVE + slot));   // covers lane a's MMAs on this K tile too
      }
      __syncwarp();
    };
    auto issue_pv = [&](int x, int j) {   // O_x += P_x V_j
      const int i = gbase + 2 * j + 1, slot = i % AT2_RING;
      if (x == 0) {
        mbar_wait(&kv_full[slot], (i / AT2_RING) & 1);
        tc_fence_after();
      }
      const uint32_t pbase = smem_base + 2 * ATT_TILE + AT2_RING * ATT_TILE + x * 2 * ATT_TILE;
      const uint32_t vbase = smem_base + 2 * ATT_TILE + slot * ATT_TILE;
      const uint32_t tO = tmem_base + 256 + (SPLIT ? x * 128 : x * 64);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
          const uint64_t ad = mk(pbase + (kk >> 2) * ATT_TILE + (kk & 3) * 32);
          const uint64_t bd = mk(vbase + kk * 2048);
          if (SPLIT) umma_bf16(tO + (kk >> 2) * 64, ad, bd, IDESC_O, (j > 0 || (kk & 3) > 0) ? 1u : 0u);   // keys 0-63 -> O_x0, 64-127 -> O_x1
          else umma_bf16(tO, ad, bd, IDESC_O, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit_a(bar_addr(I_OD + x));
        if (x == nl - 1) umma_commit_a(bar_addr(I_KVE + slot));
      }
      __syncwarp();
    };

    for (int item = blockIdx.x, it = 0; item < n_items; item += gridDim.x, ++it) {
      const Item t = decode(item);
      nl = t.nl;
      mbar_wait(q_full, it & 1);
      tc_fence_after();
      for (int x = 0; x < nl; ++x) {
        if (gj[x] > 0) {   // the lane's previous item: its last scores must be out of TMEM before S_x is overwritten
          mbar_wait(&s_free[x], (gj[x] - 1) & 1);
          tc_fence_after();
        }
        issue_qk(x, 0);
        if (n_tiles == 1 && x == nl - 1 && elect_one()) umma_commit_a(bar_addr(I_SF + 8));   // q_empty
      }
      for (int j = 0; j < n_tiles; ++j) {
        for (int x = 0; x < nl; ++x) {
          // S_x(j) is in the softmax warps' registers about half-way through their tile: the next Q K^T overlaps the rest
          if (j + 1 < n_tiles) {
            mbar_wait(&s_free[x], (gj[x] + j) & 1);
            tc_fence_after();
            issue_qk(x, j + 1);
            if (j + 2 == n_tiles && x == nl - 1 && elect_one()) umma_commit_a(bar_addr(I_SF + 8));   // q_empty: last Q K^T of the item
          }
          mbar_wait(&p_full[x], (gj[x] + j) & 1);  // P_x(j) staged, O_x rescaled
          tc_fence_after();
          issue_pv(x, j);
        }
      }
      for (int x = 0; x < nl; ++x) gj[x] += n_tiles;
      gbase += 2 * n_tiles;
    }
  } else {
    // ================================================================ softmax warps (two threads per query row)
    const int x = warp >> 3;            // query tile ("lane") of the CTA
    const int wq = warp & 3;            // TMEM lane quarter (= warp % 4)
    const int hs = (warp >> 2) & 1;     // which 64 keys of the tile / which 32 columns of O
    const int r = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + x * 128 + hs * 64;
    const uint32_t tO = tmem_base + lane_off + 256 + x * 64 + hs * 32;
    const float c = p.scale_log2;
    const int pair_id = 1 + x * 4 + wq;
    uint8_t* prow = sP + x * 2 * ATT_TILE + hs * ATT_TILE + r * 128;   // this thread's 64 keys = k-atom hs of row r
    const int sw = r & 7;
    uint32_t gj = 0;   // tiles this lane has processed so far (barrier phases)

    if constexpr (SPLIT) {
      const uint32_t tOm = tmem_base + lane_off + 256 + (x * 2 + hs) * 64;          // my accumulator: my 64 keys, all 64 columns
      const uint32_t tOo = tmem_base + lane_off + 256 + (x * 2 + (hs ^ 1)) * 64;    // the row's other key half
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (item >= n_full && x == 1) continue;   // lane b idles through single-lane items
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_tiles; ++j, ++gj) {
          mbar_wait(&s_full[x], gj & 1);
          tc_fence_after();
          const int kv_valid = min(ATT_BN, p.Tkv - j * ATT_BN) - hs * 64;   // my keys < kv_valid are real
          float l0 = 0.f, l1 = 0.f;
          bool o_ready = (j == 0);  // PV(j-1) retired: O readable / writable, P buffer free
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t sv[32];
            IDB_TMEM_LD_X32(tS + half * 32, sv);
            tmem_ld_wait();
            if (half == 1) {   // last read of S_x(j) by this warp: the next Q K^T may overwrite it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s_free[x]);
            }
            if (kv_valid < 64) {   // ragged last tile (warp-uniform branch)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (half * 32 + i >= kv_valid) sv[i] = 0xff800000u;
            }
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
              mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
            }
            const float m_t = fmaxf(mx0, mx1) * c;
            // lazy reference update: only when some row of the warp would otherwise produce P > 2^8
            if (__any_sync(0xffffffffu, m_t > m_run + AT2_RESCALE_THRESHOLD)) {
              const float m_upd = fmaxf(m_run, m_t);
              const float alpha = ex2(m_run - m_upd);   // 0 on the first chunk (m_run = -inf), 1 for rows that do not move
              m_run = m_upd;
              l_run *= alpha, l0 *= alpha, l1 *= alpha;
              if (j > 0) {
                if (!o_ready) {
                  mbar_wait(&o_done[x], (gj - 1) & 1);
                  tc_fence_after();
                  o_ready = true;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // my 64 O columns, 16 at a time (register pressure: sv and pk are live)
                  uint32_t v[16];
                  IDB_TMEM_LD_X16(tOm + q * 16, v);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                  IDB_TMEM_ST_X16(tOm + q * 16, v);
                }
                tmem_st_wait();
              }
              if (half == 1) {   // the first 32 probabilities of this tile (already staged) were formed with the old reference
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4* pp = reinterpret_cast<uint4*>(prow + ((q ^ sw) << 4));
                  uint4 w = *pp;
                  w.x = pack_bf16x2(__uint_as_float(w.x << 16) * alpha, __uint_as_float(w.x & 0xffff0000u) * alpha);
                  w.y = pack_bf16x2(__uint_as_float(w.y << 16) * alpha, __uint_as_float(w.y & 0xffff0000u) * alpha);
                  w.z = pack_bf16x2(__uint_as_float(w.z << 16) * alpha, __uint_as_float(w.z & 0xffff0000u) * alpha);
                  w.w = pack_bf16x2(__uint_as_float(w.w << 16) * alpha, __uint_as_float(w.w & 0xffff0000u) * alpha);
                  *pp = w;
                }
              }
            }
            const float neg_m = -m_run;
            uint32_t pk[16];          // bf16x2 probabilities of these 32 keys
#pragma unroll
            for (int i = 0; i < 32; i += 2) {   // masked keys: exp2(-inf) = 0 (MUFU) / 2^-126 (polynomial)
              float x0, x1, e0, e1;
              ffma2(x0, x1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), c, c, neg_m, neg_m);
              if (((i >> 1) & POLY_MASK) == POLY_MASK) {
                exp2_poly2(e0, e1, x0, x1);
              } else {
                e0 = ex2(x0);
                e1 = ex2(x1);
              }
              fadd2(l0, l1, l0, l1, e0, e1);
              pk[i >> 1] = pack_bf16x2(e0, e1);
            }
            if (!o_ready) {   // the P buffer is still being read by P V of the previous tile
              mbar_wait(&o_done[x], (gj - 1) & 1);
              tc_fence_after();
              o_ready = true;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(prow + (((half * 4 + q) ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
          l_run += l0 + l1;
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[x]);
        }
        // ---- epilogue: merge the row's two key halves.  w_h = 2^(m_h - m), O = (O_0 w_0 + O_1 w_1) / (l_0 w_0 + l_1 w_1)
        float* xm = xch_m + (x * 128 + r) * 2;
        float* xl = xch_l + (x * 128 + r) * 2;
        xm[hs] = m_run;
        xl[hs] = l_run;
        pair_barrier(pair_id);
        const float m_o = xm[hs ^ 1], l_o = xl[hs ^ 1];
        const float m_all = fmaxf(m_run, m_o);
        const float w_me = ex2(m_run - m_all), w_o = ex2(m_o - m_all);
        const float inv_l = 1.0f / (l_run * w_me + l_o * w_o);
        const float f_me = w_me * inv_l, f_o = w_o * inv_l;
        mbar_wait(&o_done[x], (gj - 1) & 1);
        tc_fence_after();
        const Item t = decode(item);
        const int row = t.q0 + x * ATT_BM + r;
        if (p.lse != nullptr && hs == 0 && row < p.Tq)
          p.lse[(static_cast<long long>(t.b) * p.heads + t.head) * p.Tq + row] = m_all + log2f(l_run * w_me + l_o * w_o);
        __nv_bfloat16* orow = p.out + (static_cast<long long>(t.b) * p.Tq + row) * p.ld_out + t.head * ATT_D + hs * 32;
#pragma unroll
        for (int q = 0; q < 2; ++q) {   // my 32 output columns, 16 at a time, from both accumulators
          uint32_t va[16], vb[16];
          IDB_TMEM_LD_X16(tOm + hs * 32 + q * 16, va);
          IDB_TMEM_LD_X16(tOo + hs * 32 + q * 16, vb);
          tmem_ld_wait();
          if (row < p.Tq) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              w[i] = pack_bf16x2(fmaf(__uint_as_float(va[2 * i]), f_me, __uint_as_float(vb[2 * i]) * f_o),
                                 fmaf(__uint_as_float(va[2 * i + 1]), f_me, __uint_as_float(vb[2 * i + 1]) * f_o));
            uint4* dst = reinterpret_cast<uint4*>(orow + q * 16);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        tc_fence_before();   // the O reads above precede the next item's first P V (ordered through p_full)
        pair_barrier(pair_id);   // the partner has read my (m, l) before the next item's epilogue overwrites them
      }
    } else {
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    if (item >= n_full && x == 1) continue;   // lane b idles through single-lane items
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j, ++gj) {
      mbar_wait(&s_full[x], gj & 1);
      tc_fence_after();
      const int kv_valid = min(ATT_BN, p.Tkv - j * ATT_BN) - hs * 64;   // my keys < kv_valid are real
      // pass 1: row maximum of my 64 scores (they are re-read from TMEM for pass 2: keeping them would spill)
      float m0 = -INFINITY, m1 = -INFINITY;
      {
        uint32_t s0[32], s1[32];
        IDB_TMEM_LD_X32(tS, s0);
        IDB_TMEM_LD_X32(tS + 32, s1);
        tmem_ld_wait();
        if (kv_valid < 64) {   // ragged last tile (warp-uniform branch)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= kv_valid) s0[i] = 0xff800000u;
            if (32 + i >= kv_valid) s1[i] = 0xff800000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1])));
        }
      }
      // exchange the partial row maxima with the thread that owns the row's other 64 keys
      float* xm = xch_m + (((gj & 1) * 2 + x) * 128 + r) * 2;
      const float m_loc = fmaxf(m0, m1);
      xm[hs] = m_loc;
      pair_barrier(pair_id);
      const float m_tile = fmaxf(m_loc, xm[hs ^ 1]) * c;
      // lazy rescale: the exponent reference m_run only moves when some row of this warp would otherwise
      // produce P > 2^8 (both warps of the pair see the same maxima, so they take the same decision)
      const bool bump = __any_sync(0xffffffffu, m_tile > m_run + AT2_RESCALE_THRESHOLD);
      float alpha = 1.0f;
      if (bump) {
        const float m_upd = fmaxf(m_run, m_tile);
        alpha = ex2(m_run - m_upd);   // 0 on the first tile (m_run = -inf)
        m_run = m_upd;
        l_run *= alpha;
      }
      const float neg_m = -m_run;
      if (j > 0) {
        mbar_wait(&o_done[x], (gj - 1) & 1);   // PV_x(j-1) retired: O readable, P buffer free
        tc_fence_after();
        if (bump) {   // my 32 of the row's 64 O columns
          uint32_t v[32];
          IDB_TMEM_LD_X32(tO, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          IDB_TMEM_ST_X32(tO, v);
          tmem_st_wait();
        }
      }
      float l0 = 0.f, l1 = 0.f;
      auto emit = [&](int half) {   // pass 2 over 32 of my keys: P = exp2(S c - m), row sum, bf16, swizzled K-major smem
        uint32_t sv[32];
        IDB_TMEM_LD_X32(tS + half * 32, sv);
        tmem_ld_wait();
        if (half == 1) {   // last read of S_x(j) by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[x]);
        }
        if (kv_valid < 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (half * 32 + i >= kv_valid) sv[i] = 0xff800000u;
        }
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {   // masked keys: exp2(-inf) = 0 (MUFU) / 2^-126 (polynomial)
          float x0, x1;
          ffma2(x0, x1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), c, c, neg_m, neg_m);
          if (((i >> 1) & POLY_MASK) == POLY_MASK) {
            exp2_poly2(pv[i], pv[i + 1], x0, x1);
          } else {
            pv[i] = ex2(x0);
            pv[i + 1] = ex2(x1);
          }
          fadd2(l0, l1, l0, l1, pv[i], pv[i + 1]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = half * 4 + q;
          uint4 w = make_uint4(pack_bf16x2(pv[8 * q], pv[8 * q + 1]), pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]),
                               pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]), pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]));
          *reinterpret_cast<uint4*>(prow + ((chunk ^ sw) << 4)) = w;
        }
      };
      emit(0);
      emit(1);
      l_run += l0 + l1;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x]);
    }
    // ---- epilogue: combine the two partial row sums, O / l (my 32 columns)
    float* xl = xch_l + (x * 128 + r) * 2;
    xl[hs] = l_run;
    pair_barrier(pair_id);
    const float inv_l = 1.0f / (l_run + xl[hs ^ 1]);
    mbar_wait(&o_done[x], (gj - 1) & 1);
    tc_fence_after();
    const Item t = decode(item);   // decoded here, not before the tile loop: nothing item-specific is live across it
    const int row = t.q0 + x * ATT_BM + r;
    if (p.lse != nullptr && hs == 0 && row < p.Tq)
      p.lse[(static_cast<long long>(t.b) * p.heads + t.head) * p.Tq + row] = m_run + log2f(l_run + xl[hs ^ 1]);
    __nv_bfloat16* orow = p.out + (static_cast<long long>(t.b) * p.Tq + row) * p.ld_out + t.head * ATT_D + hs * 32;
    {
      uint32_t v[32];
      IDB_TMEM_LD_X32(tO, v);
      tmem_ld_wait();
      if (row < p.Tq) {
        uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l),
                              pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l),
                              pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l),
                              pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l));
      }
    }
    tc_fence_before();   // the O reads above precede the next item's first P V (ordered through p_full)
    }   // items
    }   // !SPLIT
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<AT2_TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------- short-context variant
// Cross-attention over the text context (Tkv = 77): ONE K/V tile, so a CTA keeps K and V resident and walks a
// chunk of consecutive 128-row query tiles of one (image, head) instead of paying the launch / TMEM-alloc /
// barrier-init / K,V-load prologue per query tile.  S uses only NKV = ceil16(Tkv) <= 96 accumulator columns and is
// double-buffered in TMEM (S0 | S1 | O = 96 + 96 + 64 columns), Q tiles are double-buffered in smem, so the MMA
// warp computes S(i+1) while the softmax warps work on tile i.  2 CTAs per SM.
constexpr int ATX_THREADS = 192;
constexpr int ATX_SMEM = 2 * ATT_TILE /*K,V*/ + 2 * ATT_TILE /*Q x2*/ + 2 * ATT_TILE /*P*/ + 1024 + 128;
constexpr int ATX_TMEM_COLS = 256;
constexpr int ATX_S_STRIDE = 96;
constexpr int ATX_O_COL = 192;

__global__ void __launch_bounds__(ATX_THREADS, 2) attention_x_kernel(const __grid_constant__ AttnParams p, int qpc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + ATT_TILE;
  uint8_t* sQ = sV + ATT_TILE;                  // [2]
  uint8_t* sP = sQ + 2 * ATT_TILE;              // [2 k-atoms][128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE);
  uint64_t* kv_full = bars;          // 1
  uint64_t* q_full = bars + 1;       // [2]
  uint64_t* q_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;       // [2]
  uint64_t* p_full = bars + 7;       // 1 (count 4)
  uint64_t* o_done = bars + 8;       // 1
  uint64_t* o_free = bars + 9;       // 1 (count 4)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_qtiles = (p.Tq + ATT_BM - 1) / ATT_BM;
  const int qt0 = blockIdx.x * qpc;
  const int nq = min(qpc, n_qtiles - qt0);     // query tiles of this CTA (>= 1 by grid construction)
  const int nkv = (p.Tkv + 15) & ~15;          // accumulator columns / PV depth actually used

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, 4);
    mbar_init(o_done, 1);
    mbar_init(o_free, 4);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<ATX_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();   // q / k / v come from the preceding kernels

  if (warp == 4) {
    // ================================================================ TMA producer
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * ATT_TILE);
      tma_load_3d(sK, &p.tmK, kv_full, p.col0_k + head * ATT_D, 0, b);
      tma_load_3d(sV, &p.tmV, kv_full, p.col0_v + head * ATT_D, 0, b);
      for (int i = 0; i < nq; ++i) {
        const int slot = i & 1;
        mbar_wait(&q_empty[slot], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[slot], ATT_TILE);
        tma_load_3d(sQ + slot * ATT_TILE, &p.tmQ, &q_full[slot], p.col0_q + head * ATT_D, (qt0 + i) * ATT_BM, b);
      }
    }
  } else if (warp == 5) {
    // ================================================================ MMA issuer
    const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, nkv, 0, 0);
    constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BM, ATT_D, 0, 1);  // B (= V) is MN-major
    const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK));
    const uint64_t pdesc = umma_smem_desc_sw128(smem_u32(sP));
    const uint32_t vbase = smem_u32(sV);
    auto issue_qk = [&](int i) {   // S[i & 1] = Q_i K^T
      const int slot = i & 1;
      mbar_wait(&q_full[slot], (i >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t qdesc = umma_smem_desc_sw128(smem_u32(sQ + slot * ATT_TILE));
        const uint32_t tS = tmem_base + slot * ATX_S_STRIDE;
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_bf16(tS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&q_empty[slot]);
        umma_commit(&s_full[slot]);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    tc_fence_after();
    issue_qk(0);
    for (int i = 0; i < nq; ++i) {
      // S buffer (i + 1) & 1 was last read by the softmax of tile i - 1, which finished before p_full(i - 1)
      if (i + 1 < nq) issue_qk(i + 1);
      mbar_wait(p_full, i & 1);            // P_i staged
      if (i > 0) mbar_wait(o_free, (i - 1) & 1);   // O_{i-1} has been read out
      tc_fence_after();
      if (lane == 0) {
        for (int kk = 0; kk < nkv / 16; ++kk) {
          const uint64_t ad = pdesc + static_cast<uint64_t>(((kk >> 2) * ATT_TILE + (kk & 3) * 32) >> 4);
          const uint64_t bd = umma_smem_desc_sw128(vbase + kk * 2048);
          umma_bf16(tmem_base + ATX_O_COL, ad, bd, IDESC_O, kk > 0 ? 1u : 0u);
        }
        umma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    // ================================================================ softmax + epilogue warps (thread = query row)
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float c = p.scale_log2;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    for (int i = 0; i < nq; ++i) {
      const uint32_t tS = tmem_base + lane_off + (i & 1) * ATX_S_STRIDE;
      mbar_wait(&s_full[i & 1], (i >> 1) & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32], s2[32];
      IDB_TMEM_LD_X32(tS, s0);
      IDB_TMEM_LD_X32(tS + 32, s1);
      if (nkv > 64) IDB_TMEM_LD_X32(tS + 64, s2);
      tmem_ld_wait();
      float m = -INFINITY;
      // keys this query row may see: [0, kmax) -- the context length, or the row's own position under a causal mask
      const int kmax = p.causal ? min(p.Tkv, (qt0 + i) * ATT_BM + r + 1) : p.Tkv;
      if (!p.causal && p.Tkv > 64) {   // the UNet's cross-attention (77 keys): only the third chunk has masked columns
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (64 + k >= kmax) s2[k] = 0xff800000u;
          m = fmaxf(m, fmaxf(__uint_as_float(s0[k]), fmaxf(__uint_as_float(s1[k]), __uint_as_float(s2[k]))));
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (k >= kmax) s0[k] = 0xff800000u;
          if (32 + k >= kmax) s1[k] = 0xff800000u;
          if (nkv <= 64 || 64 + k >= kmax) s2[k] = 0xff800000u;
          m = fmaxf(m, fmaxf(__uint_as_float(s0[k]), fmaxf(__uint_as_float(s1[k]), __uint_as_float(s2[k]))));
        }
      }
      const float neg_m = -m * c;
      float l = 0.f;
      if (i > 0) mbar_wait(o_done, (i - 1) & 1);   // PV_{i-1} retired: the P buffer is free
      auto emit = [&](uint32_t (&sv)[32], int ch) {
        float pv[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          pv[k] = ex2(fmaf(__uint_as_float(sv[k]), c, neg_m));   // masked keys: exp2(-inf) = 0
          l += pv[k];
        }
        uint8_t* base = prow + (ch >> 1) * ATT_TILE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (ch & 1) * 4 + q;
          uint4 w = make_uint4(pack_bf16x2(pv[8 * q], pv[8 * q + 1]), pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]),
                               pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]), pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]));
          *reinterpret_cast<uint4*>(base + ((chunk ^ sw) << 4)) = w;
        }
      };
      emit(s0, 0);
      emit(s1, 1);
      if (nkv == 80) {   // 77 keys: the P V product reads 80 columns, so only 16 more exponentials (2 of the 4 smem chunks)
        float pv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          pv[k] = ex2(fmaf(__uint_as_float(s2[k]), c, neg_m));
          l += pv[k];
        }
        uint8_t* base = prow + ATT_TILE;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint4 w = make_uint4(pack_bf16x2(pv[8 * q], pv[8 * q + 1]), pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]),
                               pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]), pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]));
          *reinterpret_cast<uint4*>(base + ((q ^ sw) << 4)) = w;
        }
      } else if (nkv > 64) {
        emit(s2, 2);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- O_i / l -> global
      mbar_wait(o_done, i & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int row = (qt0 + i) * ATT_BM + r;
      if (p.lse != nullptr && row < p.Tq) p.lse[(static_cast<long long>(b) * p.heads + head) * p.Tq + row] = -neg_m + log2f(l);
      __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.Tq + row) * p.ld_out + head * ATT_D;
      uint32_t v0[32], v1[32];
      IDB_TMEM_LD_X32(tmem_base + lane_off + ATX_O_COL, v0);
      IDB_TMEM_LD_X32(tmem_base + lane_off + ATX_O_COL + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (row < p.Tq) {
        uint4* dst = reinterpret_cast<uint4*>(orow);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          dst[q] = make_uint4(pack_bf16x2(__uint_as_float(v0[8 * q]) * inv_l, __uint_as_float(v0[8 * q + 1]) * inv_l),
                              pack_bf16x2(__uint_as_float(v0[8 * q + 2]) * inv_l, __uint_as_float(v0[8 * q + 3]) * inv_l),
                              pack_bf16x2(__uint_as_float(v0[8 * q + 4]) * inv_l, __uint_as_float(v0[8 * q + 5]) * inv_l),
                              pack_bf16x2(__uint_as_float(v0[8 * q + 6]) * inv_l, __uint_as_float(v0[8 * q + 7]) * inv_l));
          dst[4 + q] = make_uint4(pack_bf16x2(__uint_as_float(v1[8 * q]) * inv_l, __uint_as_float(v1[8 * q + 1]) * inv_l),
                                  pack_bf16x2(__uint_as_float(v1[8 * q + 2]) * inv_l, __uint_as_float(v1[8 * q + 3]) * inv_l),
                                  pack_bf16x2(__uint_as_float(v1[8 * q + 4]) * inv_l, __uint_as_float(v1[8 * q + 5]) * inv_l),
                                  pack_bf16x2(__uint_as_float(v1[8 * q + 6]) * inv_l, __uint_as_float(v1[8 * q + 7]) * inv_l));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<ATX_TMEM_COLS>(tmem_base);
  }
}

}  // namespace idb

using namespace idb;

extern "C" int idb_attention(const idb_attention_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return fail(IDB_E_BADARG, "idb_attention: null args");
  if (int rc = require_sm100()) return rc;
  if (!a->q || !a->k || !a->v || !a->out) return fail(IDB_E_BADARG, "idb_attention: null pointer");
  if (a->batch <= 0 || a->heads <= 0 || a->t_q <= 0 || a->t_kv <= 0) return fail(IDB_E_BADARG, "idb_attention: bad sizes");
  if ((a->ld_q | a->ld_k | a->ld_v | a->ld_out) % 8 || (a->col0_q | a->col0_k | a->col0_v) % 8)
    return fail(IDB_E_BADARG, "idb_attention: leading dims / column offsets must be multiples of 8 elements");
  if (a->heads > 65535 || a->batch > 65535) return fail(IDB_E_BADARG, "idb_attention: grid too large");

  AttnParams p;
  memset(&p, 0, sizeof(p));
  const uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {uint64_t(a->ld_q), uint64_t(a->t_q), uint64_t(a->batch)};
    uint64_t strides[2] = {uint64_t(a->ld_q) * 2, uint64_t(a->ld_q) * a->t_q * 2};
    if (int rc = make_tmap_bf16(&p.tmQ, a->q, 3, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[3] = {uint64_t(a->ld_k), uint64_t(a->t_kv), uint64_t(a->batch)};
    uint64_t strides[2] = {uint64_t(a->ld_k) * 2, uint64_t(a->ld_k) * a->t_kv * 2};
    if (int rc = make_tmap_bf16(&p.tmK, a->k, 3, dims, strides, box)) return rc;
  }
  {
    uint64_t dims[3] = {uint64_t(a->ld_v), uint64_t(a->t_kv), uint64_t(a->batch)};
    uint64_t strides[2] = {uint64_t(a->ld_v) * 2, uint64_t(a->ld_v) * a->t_kv * 2};
    if (int rc = make_tmap_bf16(&p.tmV, a->v, 3, dims, strides, box)) return rc;
  }
  p.col0_q = a->col0_q;
  p.col0_k = a->col0_k;
  p.col0_v = a->col0_v;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.ld_out = a->ld_out;
  p.B = a->batch;
  p.heads = a->heads;
  p.Tq = a->t_q;
  p.Tkv = a->t_kv;
  p.n_kv_tiles = (a->t_kv + ATT_BN - 1) / ATT_BN;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.causal = a->causal ? 1 : 0;
  p.lse = a->lse;
  if (p.causal && a->t_kv > 96) return fail(IDB_E_UNSUPPORTED, "idb_attention: causal masking is implemented for t_kv <= 96 (short-context kernel)");

  static PerDeviceOnce configured;
  {
    cudaError_t e = ensure_dynamic_smem(attention_kernel, ATT_SMEM, configured);
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("cudaFuncSetAttribute(attention): ") + cudaGetErrorString(e));
  }
  static const int force_variant = getenv("IDB_ATTN_VARIANT") ? atoi(getenv("IDB_ATTN_VARIANT")) : 0;
  // row-split 256-query kernel from 256 x 256 on (T = 256, B = 8, 20 heads: 10.7 us vs 16.6 us for the 128-query kernel); its
  // first key tile must be full for both key halves, hence t_kv >= 256.  IDB_ATTN_VARIANT: 256 / 128 forces a kernel (profiling)
  const bool use256 = force_variant ? (force_variant == 256 && a->t_kv >= 65) : (a->t_q >= 256 && a->t_kv >= 256);
  if (a->t_kv <= 96 && (force_variant == 0 || p.causal)) {   // short context (cross-attention): K/V resident, chunks of query tiles per CTA
    static PerDeviceOnce configured4;
    {
      cudaError_t e4 = ensure_dynamic_smem(attention_x_kernel, ATX_SMEM, configured4);
      if (e4 != cudaSuccess) return fail(IDB_E_CUDA, std::string("cudaFuncSetAttribute(attention_x): ") + cudaGetErrorString(e4));
    }
    const int n_qtiles = (a->t_q + ATT_BM - 1) / ATT_BM;
    const long long units = static_cast<long long>(n_qtiles) * a->heads * a->batch;
    int qpc = static_cast<int>((units + 2LL * num_sms() - 1) / (2LL * num_sms()));   // one resident wave at 2 CTAs / SM
    if (qpc < 1) qpc = 1;
    if (qpc > n_qtiles) qpc = n_qtiles;
    dim3 grid4((n_qtiles + qpc - 1) / qpc, a->heads, a->batch);
    launch_pdl(attention_x_kernel, dim3(grid4), dim3(ATX_THREADS), ATX_SMEM, stream, p, qpc);
    cudaError_t e4 = cudaGetLastError();
    if (e4 != cudaSuccess) return fail(IDB_E_CUDA, std::string("attention_x launch: ") + cudaGetErrorString(e4));
    return IDB_OK;
  }
  static const int rs_poly = getenv("IDB_ATTN_RSPOLY") ? atoi(getenv("IDB_ATTN_RSPOLY")) : 3;
  if (use256) {   // row-split 256-query kernel (16 softmax warps)
    static const int rs_split = getenv("IDB_ATTN_SPLIT") ? atoi(getenv("IDB_ATTN_SPLIT")) : 1;   // 0: exchange the row maximum every tile (first version)
    void (*all[8])(AttnParams, int, int, int) = {attention_rs_kernel<1, false>, attention_rs_kernel<3, false>, attention_rs_kernel<7, false>,
                                                 attention_rs_kernel<32, false>, attention_rs_kernel<1, true>, attention_rs_kernel<3, true>,
                                                 attention_rs_kernel<7, true>, attention_rs_kernel<32, true>};
    const int pi = rs_poly == 1 ? 0 : (rs_poly == 7 ? 2 : (rs_poly == 32 ? 3 : 1));
    void (*kern)(AttnParams, int, int, int) = all[pi + (rs_split ? 4 : 0)];
    static PerDeviceOnce configured3[8];
    for (int i = 0; i < 8; ++i) {
      cudaError_t e3 = ensure_dynamic_smem(all[i], AT3_SMEM, configured3[i]);
      if (e3 != cudaSuccess) return fail(IDB_E_CUDA, std::string("cudaFuncSetAttribute(attention_rs): ") + cudaGetErrorString(e3));
    }
    // 256-query units; the ones that would form a partial last wave run as two single-lane CTAs each
    const int qblocks = (a->t_q + 2 * ATT_BM - 1) / (2 * ATT_BM);
    const long long units = static_cast<long long>(qblocks) * a->heads * a->batch;
    const int sms = num_sms();
    long long n_full = units;
    static const int split_tail = getenv("IDB_ATTN_TAILSPLIT") ? atoi(getenv("IDB_ATTN_TAILSPLIT")) : 1;
    const long long rem = units % sms;
    if (split_tail && units > sms && rem != 0 && 2 * rem <= sms && a->t_q % (2 * ATT_BM) == 0) n_full = units - rem;
    const long long items = n_full + 2 * (units - n_full);
    static const int persistent = getenv("IDB_ATTN_PERSISTENT") ? atoi(getenv("IDB_ATTN_PERSISTENT")) : 1;
    const long long ctas = (persistent && items > sms) ? sms : items;   // persistent CTAs walk the item list
    launch_pdl(kern, dim3(static_cast<unsigned>(ctas)), dim3(AT3_THREADS), AT3_SMEM, stream, p, static_cast<int>(n_full), qblocks,
               static_cast<int>(items));
    cudaError_t e3 = cudaGetLastError();
    if (e3 != cudaSuccess) return fail(IDB_E_CUDA, std::string("attention_rs launch: ") + cudaGetErrorString(e3));
    return IDB_OK;
  }
  dim3 grid((a->t_q + ATT_BM - 1) / ATT_BM, a->heads, a->batch);
  launch_pdl(attention_kernel, dim3(grid), dim3(ATT_THREADS), ATT_SMEM, stream, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("attention launch: ") + cudaGetErrorString(e));
  return IDB_OK;
}
