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

  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("cudaFuncSetAttribute(attention): ") + cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((a->t_q + ATT_BM - 1) / ATT_BM, a->heads, a->batch);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("attention launch: ") + cudaGetErrorString(e));
  return IDB_OK;
}
