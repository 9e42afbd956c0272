// Bandwidth-bound kernels of the LoRA-only training backward (/root/reference/train_ID-Booth.py:1140-1146:
// `accelerator.backward(loss)` with only the rank-4 attention adapters trainable, :672-678): the input gradients of
// LayerNorm, GroupNorm(+SiLU) and GEGLU, the adapter weight gradients (a tall-skinny reduction over the rows), and the
// resampling glue of the strided / upsampling convolutions.  The contractions of the backward (input gradients of every
// Linear / conv, the attention backward) run on the tensor-core kernels (idb_gemm_conv on transposed packed weights,
// idb_attention_backward).  fp32 math throughout; reductions in a fixed order (deterministic).
#include <cstdlib>
#include <string>

#include "../../include/idb.h"
#include "idb_common.cuh"
#include "idb_host.h"

namespace idb {

#define IDB_CHECK_LAUNCH_B(name)                                                                                  \
  do {                                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                                         \
    if (e__ != cudaSuccess) return fail(IDB_E_CUDA, std::string(name " launch: ") + cudaGetErrorString(e__));     \
  } while (0)

// ---------------------------------------------------------------------------------------- LayerNorm backward (warp per row)
// y = xhat * gamma + beta, xhat = (x - mu) * rstd.  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
template <int MAXQ>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                     float* __restrict__ dx, int add, long long rows, int C, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nq = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  const float4* gr = reinterpret_cast<const float4*>(dy + row * C);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float4 v[MAXQ], g[MAXQ];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int q = lane + i * 32;
    if (q < nq) {
      v[i] = xr[q];
      const float4 d = gr[q], ga = __ldg(g4 + q);
      g[i] = make_float4(d.x * ga.x, d.y * ga.y, d.z * ga.z, d.w * ga.w);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int q = lane + i * 32;
    if (q < nq) {
      v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / C + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int q = lane + i * 32;
    if (q < nq) {
      v[i].x *= rstd, v[i].y *= rstd, v[i].z *= rstd, v[i].w *= rstd;   // xhat
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
  }
  const float mg = sg / C, mgx = sgx / C;
  float4* outr = reinterpret_cast<float4*>(dx + row * C);
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int q = lane + i * 32;
    if (q < nq) {
      float4 r = make_float4(rstd * (g[i].x - mg - v[i].x * mgx), rstd * (g[i].y - mg - v[i].y * mgx),
                             rstd * (g[i].z - mg - v[i].z * mgx), rstd * (g[i].w - mg - v[i].w * mgx));
      if (add) {
        const float4 o4 = outr[q];
        r.x += o4.x, r.y += o4.y, r.z += o4.z, r.w += o4.w;
      }
      outr[q] = r;
    }
  }
}

// ---------------------------------------------------------------------------------------- GroupNorm(+SiLU) backward
struct GnBwdParams {
  const float* dy;       // [B, hw, C]: gradient with respect to act(GN(x))
  const float* x0;
  const float* x1;       // the two concatenated sources ([B, hw, c0], [B, hw, c1]); x1 may be null
  int c0, c1, C, hw, groups, cpg, silu;
  const float* stats;    // [B, groups, 2] (mean, rstd) of the forward
  const float* gamma;
  const float* beta;
  float* red;            // [B, groups, nslab, 2] scratch: per pixel-slab partials of (sum g, sum g * xhat)
  int nslab;
  float* dx0;
  float* dx1;            // outputs, same geometry as x0 / x1
  int add0, add1;        // accumulate into the output instead of overwriting it
};

__device__ __forceinline__ float gn_bwd_g(const GnBwdParams& p, float dy, float xhat, int c) {
  const float ga = p.gamma[c];
  if (!p.silu) return dy * ga;
  const float z = xhat * ga + p.beta[c];
  const float s = 1.0f / (1.0f + __expf(-z));
  return dy * (s * (1.0f + z * (1.0f - s))) * ga;
}

// CTAs (group, image, pixel slab): fixed-order partial reductions of (sum g, sum g * xhat) -> red[b][g][slab][2]; the apply
// kernel adds the slabs up in index order
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const GnBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int g = blockIdx.x, b = blockIdx.y, slab = blockIdx.z, nslab = gridDim.z;
  const float mean = p.stats[(static_cast<long long>(b) * p.groups + g) * 2], rstd = p.stats[(static_cast<long long>(b) * p.groups + g) * 2 + 1];
  const int pix_per = (p.hw + nslab - 1) / nslab;
  const int pix0 = slab * pix_per, pix1 = min(p.hw, pix0 + pix_per);
  const long long n = static_cast<long long>(max(0, pix1 - pix0)) * p.cpg;
  float s = 0.f, sx = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int pl = static_cast<int>(i / p.cpg);
    const int pix = pix0 + pl;
    const int c = g * p.cpg + static_cast<int>(i - static_cast<long long>(pl) * p.cpg);
    const float xv = c < p.c0 ? p.x0[(static_cast<long long>(b) * p.hw + pix) * p.c0 + c]
                              : p.x1[(static_cast<long long>(b) * p.hw + pix) * p.c1 + (c - p.c0)];
    const float xhat = (xv - mean) * rstd;
    const float gv = gn_bwd_g(p, p.dy[(static_cast<long long>(b) * p.hw + pix) * p.C + c], xhat, c);
    s += gv;
    sx = fmaf(gv, xhat, sx);
  }
  __shared__ float sh[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
  }
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s, sh[1][threadIdx.x >> 5] = sx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c2 = 0.f;
    for (int w = 0; w < 8; ++w) a += sh[0][w], c2 += sh[1][w];
    float* dst = p.red + ((static_cast<long long>(b) * p.groups + g) * nslab + slab) * 2;
    dst[0] = a;
    dst[1] = c2;
  }
}

__global__ void gn_bwd_apply_kernel(const GnBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(p.hw) * p.C;   // per image (blockIdx.y)
  const int b = blockIdx.y;
  const float inv_n = 1.0f / (static_cast<float>(p.hw) * p.cpg);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int pix = static_cast<int>(i / p.C);
    const int c = static_cast<int>(i - static_cast<long long>(pix) * p.C);
    const int g = c / p.cpg;
    const float mean = p.stats[(static_cast<long long>(b) * p.groups + g) * 2], rstd = p.stats[(static_cast<long long>(b) * p.groups + g) * 2 + 1];
    float sg = 0.f, sgx = 0.f;
    {
      const float2* part = reinterpret_cast<const float2*>(p.red) + (static_cast<long long>(b) * p.groups + g) * p.nslab;
      for (int k = 0; k < p.nslab; ++k) sg += part[k].x, sgx += part[k].y;   // fixed order (the partials are L1 / L2 resident)
    }
    const bool first = c < p.c0;
    const long long off = first ? (static_cast<long long>(b) * p.hw + pix) * p.c0 + c : (static_cast<long long>(b) * p.hw + pix) * p.c1 + (c - p.c0);
    const float xv = first ? p.x0[off] : p.x1[off];
    const float xhat = (xv - mean) * rstd;
    const float gv = gn_bwd_g(p, p.dy[(static_cast<long long>(b) * p.hw + pix) * p.C + c], xhat, c);
    const float r = rstd * (gv - sg * inv_n - xhat * (sgx * inv_n));
    float* dst = first ? p.dx0 : p.dx1;
    if (dst == nullptr) continue;
    dst[off] = ((first ? p.add0 : p.add1) ? dst[off] : 0.f) + r;
  }
}

// ---------------------------------------------------------------------------------------- GEGLU backward
// u: bf16 [M, 2H], rows interleaved in 16-blocks [a(16) | g(16)] (the pre-activation of the fused FF-in GEMM, recomputed);
// dh: bf16 [M, H]; du (same layout as u): da = dh * gelu(g), dg = dh * a * gelu'(g)   (exact erf GELU)
__global__ void geglu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ du,
                                 long long M, int H) {
  pdl_trigger();
  pdl_wait();
  const long long total = M * H;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / H;
    const int j = static_cast<int>(i - row * H);
    const long long ia = row * 2 * H + (j >> 4) * 32 + (j & 15), ig = ia + 16;
    const float a = __bfloat162float(u[ia]), g = __bfloat162float(u[ig]), d = __bfloat162float(dh[i]);
    const float cdf = 0.5f * (1.0f + erff(g * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);
    du[ia] = __float2bfloat16(d * g * cdf);
    du[ig] = __float2bfloat16(d * a * (cdf + g * pdf));
  }
}

// ---------------------------------------------------------------------------------------- adapter weight gradient
// out[w, r] = sum_m wide[m, w] * skinny[m, r]   (dB = dY^T (x A^T), dA = (dY B)^T x): tall-skinny, bandwidth bound on `wide`.
// grid = (ceil(W / 64), n_slabs): a CTA reduces its slab of rows for 64 columns; the slab partials are then added up in
// slab order by lora_wgrad_finish_kernel (deterministic; no float atomics).
constexpr int WG_R = 16;
constexpr int WG_ROWS = 64;   // rows per round
__global__ void __launch_bounds__(256) lora_wgrad_partial_kernel(const __nv_bfloat16* __restrict__ wide, long long ld_w, int col0_w,
                                                                 const __nv_bfloat16* __restrict__ skinny, long long ld_s, int col0_s,
                                                                 float* __restrict__ partial, long long M, int W, int R) {
  pdl_trigger();
  pdl_wait();
  const int cp = threadIdx.x & 31, rl = threadIdx.x >> 5;   // 32 column pairs (64 columns) x 8 row lanes
  const int w = blockIdx.x * 64 + 2 * cp;
  const long long rows_per = ((M + gridDim.y - 1) / gridDim.y + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  const long long m0 = blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float acc0[WG_R], acc1[WG_R];
#pragma unroll
  for (int r = 0; r < WG_R; ++r) acc0[r] = acc1[r] = 0.f;
  __shared__ float sk[WG_ROWS][WG_R];
  const bool pair_ok = (w + 1 < W) && ((ld_w | col0_w) % 2 == 0);
  for (long long mb = m0; mb < m1; mb += WG_ROWS) {
    // sixty-four rows per round: their skinny vectors go through smem (every thread needs all R values of its rows);
    // each thread then handles rows mb + rl, + 8, ... with eight independent 4-byte loads of `wide` in flight
#pragma unroll
    for (int k = 0; k < WG_ROWS * WG_R / 256; ++k) {
      const int idx = threadIdx.x + k * 256, rr = idx >> 4, cc = idx & 15;
      const long long m = mb + rr;
      sk[rr][cc] = (m < m1 && cc < R) ? __bfloat162float(skinny[m * ld_s + col0_s + cc]) : 0.f;
    }
    __syncthreads();
    float x0[8], x1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const long long m = mb + rl + 8 * k;
      x0[k] = x1[k] = 0.f;
      if (m < m1 && w < W) {
        if (pair_ok) {
          const uint32_t v = *reinterpret_cast<const uint32_t*>(wide + m * ld_w + col0_w + w);
          x0[k] = __uint_as_float(v << 16), x1[k] = __uint_as_float(v & 0xffff0000u);
        } else {
          x0[k] = __bfloat162float(wide[m * ld_w + col0_w + w]);
          if (w + 1 < W) x1[k] = __bfloat162float(wide[m * ld_w + col0_w + w + 1]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int r = 0; r < WG_R; ++r) {
        const float sv = sk[rl + 8 * k][r];
        acc0[r] = fmaf(x0[k], sv, acc0[r]);
        acc1[r] = fmaf(x1[k], sv, acc1[r]);
      }
    __syncthreads();
  }
  __shared__ float red[8][64][WG_R + 1];
#pragma unroll
  for (int r = 0; r < WG_R; ++r) red[rl][2 * cp][r] = acc0[r], red[rl][2 * cp + 1][r] = acc1[r];
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * WG_R; i += 256) {   // fixed-order sum over the eight row lanes
    const int wl = i >> 4, r = i & 15;
    if (blockIdx.x * 64 + wl < W) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][wl][r];
      partial[(static_cast<long long>(blockIdx.y) * W + blockIdx.x * 64 + wl) * WG_R + r] = t;
    }
  }
}

__global__ void lora_wgrad_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, int n_slabs, int W, int R, int out_ld,
                                         int transpose, float scale, int add) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (w, r)
  if (i >= W * R) return;
  const int w = i / R, r = i - w * R;
  float s = 0.f;
  for (int k = 0; k < n_slabs; ++k) s += partial[(static_cast<long long>(k) * W + w) * WG_R + r];
  float* dst = transpose ? out + static_cast<long long>(r) * out_ld + w : out + static_cast<long long>(w) * out_ld + r;
  *dst = (add ? *dst : 0.f) + s * scale;
}

// ---------------------------------------------------------------------------------------- resampling glue
// Z[b, 2y, 2x, :] = g[b, y, x, :], zeros elsewhere: the input gradient of a stride-2 3x3 convolution is the (flipped-weight)
// stride-1 3x3 convolution of Z
__global__ void zero_insert2x_kernel(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ z, int B, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int cq_n = C / 8;
  const long long total = static_cast<long long>(B) * 2 * H * 2 * W * cq_n;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(idx % cq_n);
    long long pix = idx / cq_n;
    const int xo = static_cast<int>(pix % (2 * W));
    pix /= 2 * W;
    const int yo = static_cast<int>(pix % (2 * H));
    const int b = static_cast<int>(pix / (2 * H));
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (!(xo & 1) && !(yo & 1)) v = *reinterpret_cast<const uint4*>(g + ((static_cast<long long>(b) * H + (yo >> 1)) * W + (xo >> 1)) * C + cq * 8);
    *reinterpret_cast<uint4*>(z + idx * 8) = v;
  }
}

// out[b, y, x, :] (+)= sum of the 2x2 block of g: the input gradient of nearest-neighbour 2x upsampling
__global__ void sumpool2x_kernel(const float* __restrict__ g, float* __restrict__ out, int add, int B, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int cq_n = C / 4;
  const long long total = static_cast<long long>(B) * H * W * cq_n;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(idx % cq_n);
    long long pix = idx / cq_n;
    const int x = static_cast<int>(pix % W);
    pix /= W;
    const int y = static_cast<int>(pix % H);
    const int b = static_cast<int>(pix / H);
    const float* src = g + ((static_cast<long long>(b) * 2 * H + 2 * y) * 2 * W + 2 * x) * C + cq * 4;
    const float4 a = *reinterpret_cast<const float4*>(src), c = *reinterpret_cast<const float4*>(src + C);
    const float4 d = *reinterpret_cast<const float4*>(src + static_cast<long long>(2 * W) * C), e = *reinterpret_cast<const float4*>(src + static_cast<long long>(2 * W) * C + C);
    float4 r = make_float4((a.x + c.x) + (d.x + e.x), (a.y + c.y) + (d.y + e.y), (a.z + c.z) + (d.z + e.z), (a.w + c.w) + (d.w + e.w));
    float4* dst = reinterpret_cast<float4*>(out + idx * 4);
    if (add) {
      const float4 o = *dst;
      r.x += o.x, r.y += o.y, r.z += o.z, r.w += o.w;
    }
    *dst = r;
  }
}

static int grid_for_b(long long work_items, int block, int max_blocks) {
  long long g = (work_items + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace idb

using namespace idb;

extern "C" int idb_layernorm_backward(const float* dy, const float* x, const float* gamma, float* dx, int32_t add, int64_t rows, int32_t c,
                                      float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!dy || !x || !gamma || !dx || rows <= 0 || c <= 0 || c % 4 || c > 2048) return fail(IDB_E_BADARG, "idb_layernorm_backward: bad arguments");
  const int warps = 8;
  const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
  const int qpl = (c / 4 + 31) / 32;
  if (qpl <= 3) launch_pdl(layernorm_bwd_kernel<3>, dim3(grid), dim3(warps * 32), 0, stream, dy, x, gamma, dx, add, rows, c, eps);
  else if (qpl <= 5) launch_pdl(layernorm_bwd_kernel<5>, dim3(grid), dim3(warps * 32), 0, stream, dy, x, gamma, dx, add, rows, c, eps);
  else if (qpl <= 10) launch_pdl(layernorm_bwd_kernel<10>, dim3(grid), dim3(warps * 32), 0, stream, dy, x, gamma, dx, add, rows, c, eps);
  else launch_pdl(layernorm_bwd_kernel<16>, dim3(grid), dim3(warps * 32), 0, stream, dy, x, gamma, dx, add, rows, c, eps);
  IDB_CHECK_LAUNCH_B("layernorm_bwd");
  return IDB_OK;
}

extern "C" int idb_groupnorm_backward(const idb_groupnorm_bwd_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return fail(IDB_E_BADARG, "idb_groupnorm_backward: null args");
  if (int rc = require_sm100()) return rc;
  const int C = a->c0 + (a->x1 ? a->c1 : 0);
  if (!a->dy || !a->x0 || !a->stats || !a->gamma || !a->beta || !a->scratch || (!a->dx0 && !a->dx1))
    return fail(IDB_E_BADARG, "idb_groupnorm_backward: null pointer");
  if (a->groups <= 0 || C % a->groups || a->batch <= 0 || a->hw <= 0) return fail(IDB_E_BADARG, "idb_groupnorm_backward: bad geometry");
  GnBwdParams p;
  p.dy = a->dy, p.x0 = a->x0, p.x1 = a->x1, p.c0 = a->c0, p.c1 = a->x1 ? a->c1 : 0, p.C = C, p.hw = a->hw, p.groups = a->groups;
  p.cpg = C / a->groups, p.silu = a->silu, p.stats = a->stats, p.gamma = a->gamma, p.beta = a->beta, p.red = a->scratch;
  p.dx0 = a->dx0, p.dx1 = a->dx1, p.add0 = a->add0, p.add1 = a->add1;
  p.nslab = a->hw >= 2048 ? 16 : (a->hw >= 256 ? 4 : 1);   // (scratch holds 16 slabs)
  launch_pdl(gn_bwd_reduce_kernel, dim3(dim3(a->groups, a->batch, p.nslab)), dim3(256), 0, stream, p);
  IDB_CHECK_LAUNCH_B("gn_bwd_reduce");
  const long long per_image = static_cast<long long>(a->hw) * C;
  launch_pdl(gn_bwd_apply_kernel, dim3(dim3(grid_for_b(per_image, 256, num_sms() * 8 / (a->batch > 8 ? 8 : a->batch) + 1), a->batch)), dim3(256), 0, stream, p);
  IDB_CHECK_LAUNCH_B("gn_bwd_apply");
  return IDB_OK;
}

extern "C" int idb_geglu_backward(const void* dh_bf16, const void* u_bf16, void* du_bf16, int64_t m, int32_t h, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!dh_bf16 || !u_bf16 || !du_bf16 || m <= 0 || h <= 0 || h % 16) return fail(IDB_E_BADARG, "idb_geglu_backward: bad arguments");
  launch_pdl(geglu_bwd_kernel, dim3(grid_for_b(m * h, 256, num_sms() * 16)), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(dh_bf16),
             static_cast<const __nv_bfloat16*>(u_bf16), static_cast<__nv_bfloat16*>(du_bf16), static_cast<long long>(m), h);
  IDB_CHECK_LAUNCH_B("geglu_bwd");
  return IDB_OK;
}

extern "C" size_t idb_lora_wgrad_workspace_bytes(int32_t w) { return static_cast<size_t>(64) * w * WG_R * sizeof(float); }

extern "C" int idb_lora_wgrad(const void* wide_bf16, int64_t ld_w, int32_t col0_w, const void* skinny_bf16, int64_t ld_s, int32_t col0_s,
                              float* out, int32_t out_ld, int32_t transpose_out, float scale, int32_t add, int64_t m, int32_t w, int32_t r,
                              float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!wide_bf16 || !skinny_bf16 || !out || !workspace || m <= 0 || w <= 0 || r <= 0 || r > WG_R)
    return fail(IDB_E_BADARG, "idb_lora_wgrad: bad arguments (rank <= 16)");
  int slabs = static_cast<int>((m + 511) / 512);
  if (slabs > 64) slabs = 64;
  launch_pdl(lora_wgrad_partial_kernel, dim3(dim3((w + 63) / 64, slabs)), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(wide_bf16),
             static_cast<long long>(ld_w), col0_w, static_cast<const __nv_bfloat16*>(skinny_bf16), static_cast<long long>(ld_s), col0_s, workspace,
             static_cast<long long>(m), w, r);
  IDB_CHECK_LAUNCH_B("lora_wgrad_partial");
  launch_pdl(lora_wgrad_finish_kernel, dim3((w * r + 255) / 256), dim3(256), 0, stream, static_cast<const float*>(workspace), out, slabs, w, r, out_ld,
             transpose_out, scale, add);
  IDB_CHECK_LAUNCH_B("lora_wgrad_finish");
  return IDB_OK;
}

extern "C" int idb_zero_insert2x(const void* g_bf16, void* z_bf16, int32_t batch, int32_t h, int32_t w, int32_t c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!g_bf16 || !z_bf16 || c % 8) return fail(IDB_E_BADARG, "idb_zero_insert2x: bad arguments");
  const long long total = static_cast<long long>(batch) * 4 * h * w * (c / 8);
  launch_pdl(zero_insert2x_kernel, dim3(grid_for_b(total, 256, num_sms() * 16)), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(g_bf16),
             static_cast<__nv_bfloat16*>(z_bf16), batch, h, w, c);
  IDB_CHECK_LAUNCH_B("zero_insert2x");
  return IDB_OK;
}

extern "C" int idb_sumpool2x(const float* g, float* out, int32_t add, int32_t batch, int32_t h, int32_t w, int32_t c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!g || !out || c % 4) return fail(IDB_E_BADARG, "idb_sumpool2x: bad arguments");
  const long long total = static_cast<long long>(batch) * h * w * (c / 4);
  launch_pdl(sumpool2x_kernel, dim3(grid_for_b(total, 256, num_sms() * 16)), dim3(256), 0, stream, g, out, add, batch, h, w, c);
  IDB_CHECK_LAUNCH_B("sumpool2x");
  return IDB_OK;
}
