// idb_attention_backward: gradients of O = softmax(Q K^T * scale) V with respect to Q, K and V (head_dim 64) on tcgen05,
// for the LoRA-only training backward (/root/reference/train_ID-Booth.py:1140: the q / k / v / out
// projections of every attention carry the trainable rank-4 adapters, train_ID-Booth.py:672-678).
//
// One CTA = one (image, head, 128-key tile j); it keeps K_j, V_j resident and walks the query tiles i:
//   S  = Q_i K_j^T              dP = dO_i V_j^T                    (two UMMA groups into TMEM)
//   P  = exp2(S c - LSE_i)      dS = P o (dP - D_i) * scale        (softmax warps: TMEM -> registers -> bf16 smem tiles)
//   dV_j += P^T dO_i            dK_j += dS^T Q_i                   (A operands MN-major: the transposition is the tensor core's)
//   dQ_i  = dS K_j              -> fp32 atomics into dQ (every key tile contributes to every query row)
// LSE_i (log2 domain) comes from the forward kernels (idb_attention_args.lse), D_i = rowsum(dO o O) from attn_bwd_prep_kernel.
//   warps 0-3 : thread r = query row r of the tile (P / dS formation, dQ read-out), finally key row r (dK / dV read-out)
//   warp  4   : TMA producer          warp 5 : tcgen05.mma issuer, owns the TMEM allocation
// A first version: Q / dO tiles are double-buffered (the next tile loads, and its S / dP MMAs queue, behind the current
// tile's dV / dK / dQ MMAs); S, dP, P, dS are single-buffered, so the softmax-backward phase and the MMAs of one query
// tile still alternate.
#include <cstdlib>
#include <string>

#include "../../include/idb.h"
#include "idb_common.cuh"
#include "idb_host.h"

namespace idb {

constexpr int AB_TILE = 16384;   // 128 x 64 bf16
constexpr int AB_THREADS = 192;
constexpr int AB_TMEM_COLS = 512;
constexpr int AB_S = 0, AB_DP = 128, AB_DV = 256, AB_DK = 320, AB_DQ = 384;
constexpr int AB_SMEM = 10 * AB_TILE /* K V | Q(2) dO(2) | P(2) dS(2) */ + 1024 + 256;

struct AttnBwdParams {
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int col0_q, col0_k, col0_v, col0_do;
  const float* lse;     // [B, heads, Tq] log2-domain log-sum-exp of the scaled scores
  const float* dsum;    // [B, heads, Tq] rowsum(dO o O)
  float* dq;            // fp32 [B * Tq, ld_dq]; head h at columns col0_dq + 64 h; zero on entry (atomically accumulated)
  long long ld_dq;
  int col0_dq;
  __nv_bfloat16* dk;    // bf16 [B * Tkv, ld_dk], head h at col0_dk + 64 h
  __nv_bfloat16* dv;
  long long ld_dk, ld_dv;
  int col0_dk, col0_dv;
  int B, heads, Tq, Tkv, n_q_tiles;
  float scale_log2, scale;
};

// D[b, h, t] = sum_d dO[b, t, h, d] * O[b, t, h, d]   (one thread per (row, head); 8 x 16-byte loads each)
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, long long ld_o, int col0_o,
                                     const __nv_bfloat16* __restrict__ d_o, long long ld_do, int col0_do, float* __restrict__ dsum,
                                     int B, int heads, int Tq) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(B) * heads * Tq;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % Tq);
    const int h = static_cast<int>((i / Tq) % heads);
    const int b = static_cast<int>(i / (static_cast<long long>(Tq) * heads));
    const uint4* po = reinterpret_cast<const uint4*>(o + (static_cast<long long>(b) * Tq + t) * ld_o + col0_o + h * 64);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + (static_cast<long long>(b) * Tq + t) * ld_do + col0_do + h * 64);
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 a = po[q], g = pd[q];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc = fmaf(__uint_as_float(aw[k] << 16), __uint_as_float(gw[k] << 16), acc);
        acc = fmaf(__uint_as_float(aw[k] & 0xffff0000u), __uint_as_float(gw[k] & 0xffff0000u), acc);
      }
    }
    dsum[(static_cast<long long>(b) * heads + h) * Tq + t] = acc;
  }
}

// smem descriptor of an MN-major SWIZZLE_128B operand that is TWO 64-element blocks wide along M / N (blocks 16 KiB apart):
// canonical layout ((8, n), (8, k)) : ((1, LBO), (8, SBO)) in 16-byte units -> LBO = 16384 B, SBO = 1024 B
__device__ __forceinline__ uint64_t umma_smem_desc_sw128_mn2(uint32_t smem_addr) {
  return umma_smem_desc_sw128(smem_addr) | (static_cast<uint64_t>((AB_TILE >> 4) & 0x3FFF) << 16);
}

__global__ void __launch_bounds__(AB_THREADS, 1) attention_bwd_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + AB_TILE;
  uint8_t* sQ = sV + AB_TILE;         // [2 slots]: the next query tile's Q / dO load while this one is being worked on
  uint8_t* sDO = sQ + 2 * AB_TILE;    // [2 slots]
  uint8_t* sP = sDO + 2 * AB_TILE;    // [2 key atoms][128 query rows x 64 keys]
  uint8_t* sDS = sP + 2 * AB_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * AB_TILE);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;      // [2]
  uint64_t* qdo_empty = bars + 3;     // [2] every MMA that reads the slot has retired
  uint64_t* sdp_full = bars + 5;
  uint64_t* pds_full = bars + 6;      // count 4 (softmax warps)
  uint64_t* mma2_done = bars + 7;     // dV / dK / dQ MMAs of the query tile retired: P, dS free, dQ readable
  uint64_t* dq_free = bars + 8;       // count 4: dQ read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nq = p.n_q_tiles;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmDO);
    mbar_init(kv_full, 1);
    for (int k = 0; k < 2; ++k) {
      mbar_init(&qdo_full[k], 1);
      mbar_init(&qdo_empty[k], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 4);
    mbar_init(mma2_done, 1);
    mbar_init(dq_free, 4);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<AB_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 4) {
    // ================================================================ TMA producer
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * AB_TILE);
      tma_load_3d(sK, &p.tmK, kv_full, p.col0_k + head * 64, j * 128, b);
      tma_load_3d(sV, &p.tmV, kv_full, p.col0_v + head * 64, j * 128, b);
      for (int i = 0; i < nq; ++i) {
        const int slot = i & 1;
        mbar_wait(&qdo_empty[slot], ((i >> 1) & 1) ^ 1);   // every MMA that read the slot's previous tile has retired
        mbar_expect_tx(&qdo_full[slot], 2 * AB_TILE);
        tma_load_3d(sQ + slot * AB_TILE, &p.tmQ, &qdo_full[slot], p.col0_q + head * 64, i * 128, b);
        tma_load_3d(sDO + slot * AB_TILE, &p.tmDO, &qdo_full[slot], p.col0_do + head * 64, i * 128, b);
      }
    }
  } else if (warp == 5) {
    // ================================================================ MMA issuer
    constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, 0, 0);   // S = Q K^T, dP = dO V^T: both operands K-major (d contiguous)
    constexpr uint32_t IDESC_T = umma_idesc_bf16(128, 64, 1, 1);    // dV = P^T dO, dK = dS^T Q: A and B MN-major
    constexpr uint32_t IDESC_Q = umma_idesc_bf16(128, 64, 0, 1);    // dQ = dS K: A K-major, B (= K_j) MN-major
    const uint64_t kdesc = umma_smem_desc_sw128(smem_u32(sK)), vdesc = umma_smem_desc_sw128(smem_u32(sV));
    const uint64_t qdesc0 = umma_smem_desc_sw128(smem_u32(sQ)), dodesc0 = umma_smem_desc_sw128(smem_u32(sDO));
    const uint64_t pdesc_mn = umma_smem_desc_sw128_mn2(smem_u32(sP)), dsdesc_mn = umma_smem_desc_sw128_mn2(smem_u32(sDS));
    const uint64_t dsdesc_k = umma_smem_desc_sw128(smem_u32(sDS));
    mbar_wait(kv_full, 0);
    tc_fence_after();
    for (int i = 0; i < nq; ++i) {
      const int slot = i & 1;
      const uint64_t qdesc = qdesc0 + static_cast<uint64_t>((slot * AB_TILE) >> 4), dodesc = dodesc0 + static_cast<uint64_t>((slot * AB_TILE) >> 4);
      mbar_wait(&qdo_full[slot], (i >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + AB_S, qdesc + 2 * k, kdesc + 2 * k, IDESC_S, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + AB_DP, dodesc + 2 * k, vdesc + 2 * k, IDESC_S, k > 0 ? 1u : 0u);
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_full, i & 1);                    // P_i / dS_i staged; S / dP consumed
      if (i > 0) mbar_wait(dq_free, (i - 1) & 1);    // dQ_{i-1} has been read out of TMEM
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // K = 128 query rows, 16 per step = 2 KiB inside every 64-wide block
          const uint64_t step = static_cast<uint64_t>((kk * 2048) >> 4);
          umma_bf16(tmem_base + AB_DV, pdesc_mn + step, dodesc + step, IDESC_T, (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t step = static_cast<uint64_t>((kk * 2048) >> 4);
          umma_bf16(tmem_base + AB_DK, dsdesc_mn + step, qdesc + step, IDESC_T, (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {   // K = 128 keys: two 64-wide K atoms of dS, 16 key rows of K_j per step
          const uint64_t ad = dsdesc_k + static_cast<uint64_t>(((kk >> 2) * AB_TILE + (kk & 3) * 32) >> 4);
          const uint64_t bd = kdesc + static_cast<uint64_t>((kk * 2048) >> 4);
          umma_bf16(tmem_base + AB_DQ, ad, bd, IDESC_Q, kk > 0 ? 1u : 0u);
        }
        umma_commit(mma2_done);
        umma_commit(&qdo_empty[slot]);
      }
      __syncwarp();
    }
  } else {
    // ================================================================ softmax-backward warps (thread = query row, then key row)
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float c = p.scale_log2;
    uint8_t* prow = sP + r * 128;
    uint8_t* dsrow = sDS + r * 128;
    const int sw = r & 7;
    const long long stat0 = (static_cast<long long>(b) * p.heads + head) * p.Tq;

    auto flush_dq = [&](int i) {   // dQ_i (this key tile's contribution) -> global fp32, atomically
      mbar_wait(mma2_done, i & 1);
      tc_fence_after();
      const int row = i * 128 + r;
      float* dst = p.dq + (static_cast<long long>(b) * p.Tq + row) * p.ld_dq + p.col0_dq + head * 64;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        IDB_TMEM_LD_X32(tmem_base + lane_off + AB_DQ + ch * 32, v);
        tmem_ld_wait();
        if (row < p.Tq) {
#pragma unroll
          for (int k = 0; k < 32; ++k) atomicAdd(dst + ch * 32 + k, __uint_as_float(v[k]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
    };

    for (int i = 0; i < nq; ++i) {
      const int row = i * 128 + r;
      const float lse_r = row < p.Tq ? p.lse[stat0 + row] : 0.f;
      const float d_r = row < p.Tq ? p.dsum[stat0 + row] : 0.f;
      if (i > 0) flush_dq(i - 1);       // also: the MMAs that read P_{i-1} / dS_{i-1} have retired
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t sv[32], dv[32];
        IDB_TMEM_LD_X32(tmem_base + lane_off + AB_S + ch * 32, sv);
        IDB_TMEM_LD_X32(tmem_base + lane_off + AB_DP + ch * 32, dv);
        tmem_ld_wait();
        uint32_t pw[16], dw[16];
        const int key0 = j * 128 + ch * 32;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          float p0 = ex2(fmaf(__uint_as_float(sv[k]), c, -lse_r));
          float p1 = ex2(fmaf(__uint_as_float(sv[k + 1]), c, -lse_r));
          if (key0 + k >= p.Tkv) p0 = 0.f;        // keys past the context: TMA zero-filled K / V rows
          if (key0 + k + 1 >= p.Tkv) p1 = 0.f;
          const float g0 = p0 * (__uint_as_float(dv[k]) - d_r) * p.scale;
          const float g1 = p1 * (__uint_as_float(dv[k + 1]) - d_r) * p.scale;
          pw[k >> 1] = pack_bf16x2(p0, p1);
          dw[k >> 1] = pack_bf16x2(g0, g1);
        }
        // query row r, keys ch*32 .. +31 -> key atom (ch >> 1), 16-byte chunks ((ch & 1) * 4 + q) of the 128-byte row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((ch & 1) * 4 + q) ^ sw;
          *reinterpret_cast<uint4*>(prow + (ch >> 1) * AB_TILE + (chunk << 4)) = make_uint4(pw[4 * q], pw[4 * q + 1], pw[4 * q + 2], pw[4 * q + 3]);
          *reinterpret_cast<uint4*>(dsrow + (ch >> 1) * AB_TILE + (chunk << 4)) = make_uint4(dw[4 * q], dw[4 * q + 1], dw[4 * q + 2], dw[4 * q + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    flush_dq(nq - 1);
    // ---- dK_j, dV_j: TMEM lane = key row
    const int krow = j * 128 + r;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* base = which == 0 ? p.dv : p.dk;
      const long long ld = which == 0 ? p.ld_dv : p.ld_dk;
      const int c0 = which == 0 ? p.col0_dv : p.col0_dk;
      __nv_bfloat16* dst = base + (static_cast<long long>(b) * p.Tkv + krow) * ld + c0 + head * 64;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        IDB_TMEM_LD_X32(tmem_base + lane_off + (which == 0 ? AB_DV : AB_DK) + ch * 32, v);
        tmem_ld_wait();
        if (krow < p.Tkv) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            d4[q] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q]), __uint_as_float(v[8 * q + 1])),
                               pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3])),
                               pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5])),
                               pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<AB_TMEM_COLS>(tmem_base);
  }
}

}  // namespace idb

using namespace idb;

extern "C" int idb_attention_backward(const idb_attention_bwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return fail(IDB_E_BADARG, "idb_attention_backward: null args");
  if (int rc = require_sm100()) return rc;
  if (!a->q || !a->k || !a->v || !a->o || !a->d_o || !a->lse || !a->dsum || !a->dq || !a->dk || !a->dv)
    return fail(IDB_E_BADARG, "idb_attention_backward: null pointer");
  if (a->batch <= 0 || a->heads <= 0 || a->t_q <= 0 || a->t_kv <= 0) return fail(IDB_E_BADARG, "idb_attention_backward: bad sizes");
  if ((a->ld_q | a->ld_k | a->ld_v | a->ld_o | a->ld_do | a->ld_dk | a->ld_dv) % 8 ||
      (a->col0_q | a->col0_k | a->col0_v | a->col0_o | a->col0_do | a->col0_dk | a->col0_dv) % 8 || a->ld_dq % 4 || a->col0_dq % 4)
    return fail(IDB_E_BADARG, "idb_attention_backward: leading dims / column offsets must be multiples of 8 elements");
  if (a->heads > 65535 || a->batch > 65535) return fail(IDB_E_BADARG, "idb_attention_backward: grid too large");

  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  const uint32_t box[3] = {64, 128, 1};
  auto tmap = [&](CUtensorMap* m, const void* ptr, int64_t ld, int t) {
    uint64_t dims[3] = {uint64_t(ld), uint64_t(t), uint64_t(a->batch)};
    uint64_t strides[2] = {uint64_t(ld) * 2, uint64_t(ld) * t * 2};
    return make_tmap_bf16(m, ptr, 3, dims, strides, box);
  };
  if (int rc = tmap(&p.tmQ, a->q, a->ld_q, a->t_q)) return rc;
  if (int rc = tmap(&p.tmK, a->k, a->ld_k, a->t_kv)) return rc;
  if (int rc = tmap(&p.tmV, a->v, a->ld_v, a->t_kv)) return rc;
  if (int rc = tmap(&p.tmDO, a->d_o, a->ld_do, a->t_q)) return rc;
  p.col0_q = a->col0_q, p.col0_k = a->col0_k, p.col0_v = a->col0_v, p.col0_do = a->col0_do;
  p.lse = a->lse, p.dsum = a->dsum;
  p.dq = a->dq, p.ld_dq = a->ld_dq, p.col0_dq = a->col0_dq;
  p.dk = static_cast<__nv_bfloat16*>(a->dk), p.dv = static_cast<__nv_bfloat16*>(a->dv);
  p.ld_dk = a->ld_dk, p.ld_dv = a->ld_dv, p.col0_dk = a->col0_dk, p.col0_dv = a->col0_dv;
  p.B = a->batch, p.heads = a->heads, p.Tq = a->t_q, p.Tkv = a->t_kv;
  p.n_q_tiles = (a->t_q + 127) / 128;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;

  {
    const long long total = static_cast<long long>(a->batch) * a->heads * a->t_q;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    launch_pdl(attn_bwd_prep_kernel, dim3(blocks), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(a->o), static_cast<long long>(a->ld_o),
               a->col0_o, static_cast<const __nv_bfloat16*>(a->d_o), static_cast<long long>(a->ld_do), a->col0_do, a->dsum, a->batch, a->heads, a->t_q);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("attn_bwd_prep launch: ") + cudaGetErrorString(e));
  }
  static PerDeviceOnce configured;
  {
    cudaError_t e = ensure_dynamic_smem(attention_bwd_kernel, AB_SMEM, configured);
    if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("cudaFuncSetAttribute(attention_bwd): ") + cudaGetErrorString(e));
  }
  dim3 grid((a->t_kv + 127) / 128, a->heads, a->batch);
  launch_pdl(attention_bwd_kernel, dim3(grid), dim3(AB_THREADS), AB_SMEM, stream, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(IDB_E_CUDA, std::string("attention_bwd launch: ") + cudaGetErrorString(e));
  return IDB_OK;
}
