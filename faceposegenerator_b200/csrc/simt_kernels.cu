// Bandwidth-bound SIMT kernels of the hot path: GroupNorm(+SiLU), LayerNorm, row softmax,
// time embedding, edge convolutions (tiny Cin / tiny Cout), nearest-2x upsample, casts and the
// fused CFG + DDPMScheduler.step.  All vectorised (16 B per thread per access where the layout
// allows), coalesced along the NHWC channel axis, fp32 math.
#include <cstdlib>
#include <string>

#include "../../include/idb.h"
#include "idb_common.cuh"
#include "idb_host.h"

namespace idb {

#define IDB_CHECK_LAUNCH(name)                                                                \
  do {                                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) return fail(IDB_E_CUDA, std::string(name " launch: ") + cudaGetErrorString(e__)); \
  } while (0)

constexpr int GN_MAX_CHUNKS = 256;
constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_FUSE_MAX_RB = 8;     // fused finalize: at most 8 row blocks (256 pixels) per image ...
constexpr int GN_FUSE_MAX_CQ = 640;   // ... and at most 2560 channels (one thread per channel quad)

// ---------------------------------------------------------------------------------------- GroupNorm
struct GnParams {
  const float* x0;
  const float* x1;
  int c0, c1, C, CQ, PY;
  int hw, groups, cpg, nchunks, pix_per_chunk;
  float eps;
  const float* gamma;
  const float* beta;
  int silu;
  __nv_bfloat16* out_norm;
  __nv_bfloat16* out_raw;
  float* partials;  // [B, nchunks, groups, 2]
  float* stats;     // [B, groups, 2] = (mean, rstd), written by the last stats CTA of each image
  unsigned int* counters;  // [B] arrival tickets (zero on entry, reset to zero by the last CTA)
  // fused finalize (small rasters): the apply CTAs reduce the producers' row-block channel sums themselves
  const float2* rb0;   // non-null selects the fused path; [row block][c0] (sum, sum of squares) of source 0
  const float2* rb1;   // same for source 1
  int rbpi, phases0;   // row blocks per image (<= GN_FUSE_MAX_RB), phased layout of source 0 (see gn_finalize_kernel)
  // per-image channel sums accumulated by the producing GEMMs (idb_gemm_conv stats_image_sums): [batch][c] 64-bit fixed
  // point (sum * 2^32, sum of squares * 2^24)
  const longlong2* sums0;
  const longlong2* sums1;
  double inv_n;   // 1 / (hw * channels per group)
  int gran;       // channels per granule of sums0 / sums1
  int b1_mod;     // > 0: source 1 (and its sums) holds b1_mod images, image b reads image b % b1_mod (a skip tensor shared by the CFG pair)
};

__device__ __forceinline__ float4 gn_load(const GnParams& p, int b, int pix, int cq) {
  const int c = cq * 4;
  if (c < p.c0) return *reinterpret_cast<const float4*>(p.x0 + (static_cast<long long>(b) * p.hw + pix) * p.c0 + c);
  const int b1 = p.b1_mod > 0 ? b % p.b1_mod : b;
  return *reinterpret_cast<const float4*>(p.x1 + (static_cast<long long>(b1) * p.hw + pix) * p.c1 + (c - p.c0));
}

// Deterministic: per-thread partial sums go to smem and are combined in a fixed order (no float
// atomics), so repeated runs / different GPUs give bit-identical statistics.
__global__ void gn_stats_kernel(const GnParams p) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float gn_sm[];  // [PY][C] sums, then [PY][C] sums of squares
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int cq = threadIdx.x % p.CQ, py = threadIdx.x / p.CQ;
  const int pbeg = chunk * p.pix_per_chunk;
  const int pend = min(p.hw, pbeg + p.pix_per_chunk);
  if (py < p.PY) {
    float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    int pix = pbeg + py;
    for (; pix + 3 * p.PY < pend; pix += 4 * p.PY) {  // 4 independent 16-byte loads in flight
      const float4 v0 = gn_load(p, b, pix, cq), v1 = gn_load(p, b, pix + p.PY, cq);
      const float4 v2 = gn_load(p, b, pix + 2 * p.PY, cq), v3 = gn_load(p, b, pix + 3 * p.PY, cq);
      s[0] += (v0.x + v1.x) + (v2.x + v3.x), s[1] += (v0.y + v1.y) + (v2.y + v3.y);
      s[2] += (v0.z + v1.z) + (v2.z + v3.z), s[3] += (v0.w + v1.w) + (v2.w + v3.w);
      ss[0] += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x);
      ss[1] += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
      ss[2] += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z);
      ss[3] += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
    }
    for (; pix < pend; pix += p.PY) {
      const float4 v = gn_load(p, b, pix, cq);
      s[0] += v.x, s[1] += v.y, s[2] += v.z, s[3] += v.w;
      ss[0] += v.x * v.x, ss[1] += v.y * v.y, ss[2] += v.z * v.z, ss[3] += v.w * v.w;
    }
    float* ds = gn_sm + py * p.C + cq * 4;
    float* dss = gn_sm + (p.PY + py) * p.C + cq * 4;
    *reinterpret_cast<float4*>(ds) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(dss) = make_float4(ss[0], ss[1], ss[2], ss[3]);
  }
  __syncthreads();
  // fixed-order combine: warp w handles groups w, w+nw, ...; lanes stride the (py, channel) cells
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int g = warp; g < p.groups; g += nw) {
    float a = 0.f, a2 = 0.f;
    const int cells = p.PY * p.cpg;
    for (int i = lane; i < cells; i += 32) {
      const int yy = i / p.cpg, c = g * p.cpg + (i - yy * p.cpg);
      a += gn_sm[yy * p.C + c];
      a2 += gn_sm[(p.PY + yy) * p.C + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
      float* dst = p.partials + ((static_cast<long long>(b) * p.nchunks + chunk) * p.groups + g) * 2;
      dst[0] = a;
      dst[1] = a2;
    }
  }
  // ---- the last CTA of image b to arrive folds the per-chunk partials (fixed order -> deterministic)
  __shared__ unsigned int s_ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&p.counters[b], 1u);
  __syncthreads();
  if (s_ticket != static_cast<unsigned int>(p.nchunks - 1)) return;
  __threadfence();
  for (int g = warp; g < p.groups; g += nw) {
    double s = 0.0, ss = 0.0;
    const float* src = p.partials + (static_cast<long long>(b) * p.nchunks * p.groups + g) * 2;
    for (int i = lane; i < p.nchunks; i += 32) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(src + static_cast<long long>(i) * p.groups * 2));
      s += v.x;
      ss += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
      const double n = static_cast<double>(p.hw) * p.cpg;
      const double mean = s / n;
      double var = ss / n - mean * mean;
      if (var < 0.0) var = 0.0;
      float* dst = p.stats + (static_cast<long long>(b) * p.groups + g) * 2;
      dst[0] = static_cast<float>(mean);
      dst[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
    }
  }
  if (threadIdx.x == 0) p.counters[b] = 0u;   // ready for the next GroupNorm call on this workspace
}

// Group statistics from the row-block channel sums that the producing GEMM epilogues wrote
// (fixed summation order -> deterministic).  grid = (groups, batch), 256 threads per (image, group).
__global__ void gn_finalize_kernel(const float2* __restrict__ s0, int c0, const float2* __restrict__ s1, int c1,
                                   int rb_per_image, int cpg, int hw, float eps, float* __restrict__ stats, int phases0) {
  pdl_trigger();
  pdl_wait();
  const int g = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int groups = gridDim.x;
  __shared__ double sh_s[8], sh_ss[8];
  double s = 0.0, ss = 0.0;
  // cells (row block, channel of the group) are spread over the 256 threads, four independent loads in flight per
  // thread (a group may straddle the two concatenated sources, so the source is chosen per channel)
  const int cbase = g * cpg;
  const long long row0 = static_cast<long long>(b) * rb_per_image;
  const int cells = rb_per_image * cpg;
  auto cell = [&](int i) -> float2 {
    if (i >= cells) return make_float2(0.f, 0.f);
    const int rb = i / cpg, c = cbase + (i - rb * cpg);
    if (c >= c0) return s1[(row0 + rb) * c1 + (c - c0)];
    if (phases0 <= 1) return s0[(row0 + rb) * c0 + c];
    // source 0 written as `phases0` phased GEMM outputs: [phase][image][row block of the low-resolution raster]
    const int rbl = rb_per_image / phases0;
    const int ph = rb / rbl, r2 = rb - ph * rbl;
    const long long row = (static_cast<long long>(ph) * gridDim.y + b) * rbl + r2;
    return s0[row * c0 + c];
  };
  for (int i = threadIdx.x; i < cells; i += 4 * 256) {
    const float2 v0 = cell(i), v1 = cell(i + 256), v2 = cell(i + 512), v3 = cell(i + 768);
    s += (static_cast<double>(v0.x) + v1.x) + (static_cast<double>(v2.x) + v3.x);
    ss += (static_cast<double>(v0.y) + v1.y) + (static_cast<double>(v2.y) + v3.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) sh_s[warp] = s, sh_ss[warp] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tss = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) ts += sh_s[w], tss += sh_ss[w];
    const double n = static_cast<double>(hw) * cpg;
    const double mean = ts / n;
    double var = tss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    float* dst = stats + (static_cast<long long>(b) * groups + g) * 2;
    dst[0] = static_cast<float>(mean);
    dst[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

__device__ __forceinline__ void gn_emit(const GnParams& p, int b, int pix, int c, const float4 v, const float* sc,
                                        const float* sh) {
  float y[4] = {v.x * sc[0] + sh[0], v.y * sc[1] + sh[1], v.z * sc[2] + sh[2], v.w * sc[3] + sh[3]};
  if (p.silu) {
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = silu_f(y[i]);
  }
  const long long o = (static_cast<long long>(b) * p.hw + pix) * p.C + c;
  *reinterpret_cast<uint2*>(p.out_norm + o) = make_uint2(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]));
  if (p.out_raw) *reinterpret_cast<uint2*>(p.out_raw + o) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

__global__ void __launch_bounds__(1024) gn_apply_kernel(const GnParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float g_mean[GN_MAX_GROUPS], g_rstd[GN_MAX_GROUPS];
  __shared__ double f_s[GN_FUSE_MAX_CQ], f_ss[GN_FUSE_MAX_CQ];
  const int chunk = blockIdx.x, b = blockIdx.y;
  if (p.sums0 != nullptr) {
    // Group statistics straight from the producers' per-image granule sums (a few hundred bytes per image): thread g adds
    // up the cpg / gran granules of group g (which may lie in either of the two concatenated sources) -- integer sums,
    // exact; the only double-precision work is the cancellation-prone E[x^2] - mean^2.  No statistics pass, no finalize.
    const int g = threadIdx.x;
    if (g < p.groups) {
      const int per = p.cpg / p.gran, g0c = g * p.cpg;
      const int n0 = p.c0 / p.gran, n1 = p.c1 / p.gran;
      long long s = 0, ss = 0;
      for (int i = 0; i < per; ++i) {
        const int q = g0c / p.gran + i;     // granule index in the concatenated channel space
        const longlong2 v = (q < n0) ? __ldg(p.sums0 + static_cast<long long>(b) * n0 + q)
                                     : __ldg(p.sums1 + static_cast<long long>(p.b1_mod > 0 ? b % p.b1_mod : b) * n1 + (q - n0));
        s += v.x;
        ss += v.y;
      }
      const double mean = static_cast<double>(s) * (1.0 / 4294967296.0) * p.inv_n;
      double var = fma(-mean, mean, static_cast<double>(ss) * (1.0 / 16777216.0) * p.inv_n);
      if (var < 0.0) var = 0.0;
      g_mean[g] = static_cast<float>(mean);
      g_rstd[g] = rsqrtf(static_cast<float>(var) + p.eps);
    }
  } else if (p.rb0 != nullptr) {
    // Fused finalize (rasters of <= 256 pixels: the row-block sums of an image are a few KB, so every CTA reduces
    // them itself instead of waiting for a separate tiny kernel).  Thread t owns channels 4t .. 4t+3 (one group,
    // cpg % 4 == 0); all its loads are issued before the first use; fixed summation order -> deterministic.
    const int t = threadIdx.x;
    if (t < p.CQ) {
      const int c = 4 * t;
      const long long row0 = static_cast<long long>(b) * p.rbpi;
      const int rbl = p.phases0 > 1 ? p.rbpi / p.phases0 : p.rbpi;
      auto src_of = [&](int rb) -> const float4* {
        const float2* src;
        if (c >= p.c0) {
          src = p.rb1 + (row0 + rb) * p.c1 + (c - p.c0);
        } else if (p.phases0 <= 1) {
          src = p.rb0 + (row0 + rb) * p.c0 + c;
        } else {   // [phase][image][row block of the low-resolution raster]
          const int ph = rb / rbl, r2 = rb - ph * rbl;
          src = p.rb0 + ((static_cast<long long>(ph) * gridDim.y + b) * rbl + r2) * p.c0 + c;
        }
        return reinterpret_cast<const float4*>(src);   // [0] = (s, ss) of channels c, c + 1; [1] = c + 2, c + 3
      };
      double s = 0.0, ss = 0.0;
      for (int rb0 = 0; rb0 < p.rbpi; rb0 += 4) {   // four row blocks = eight 16-byte loads in flight
        float4 va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          va[i] = vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rb0 + i < p.rbpi) {
            const float4* src = src_of(rb0 + i);
            va[i] = src[0];
            vb[i] = src[1];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s += (static_cast<double>(va[i].x) + va[i].z) + (static_cast<double>(vb[i].x) + vb[i].z);
          ss += (static_cast<double>(va[i].y) + va[i].w) + (static_cast<double>(vb[i].y) + vb[i].w);
        }
      }
      f_s[t] = s;
      f_ss[t] = ss;
    }
    __syncthreads();
    if (t < p.groups) {
      const int nq = p.cpg >> 2, q0 = t * nq;
      double ts = 0.0, tss = 0.0;
      for (int q = 0; q < nq; ++q) ts += f_s[q0 + q], tss += f_ss[q0 + q];
      const double n = static_cast<double>(p.hw) * p.cpg;
      const double mean = ts / n;
      double var = tss / n - mean * mean;
      if (var < 0.0) var = 0.0;
      g_mean[t] = static_cast<float>(mean);
      g_rstd[t] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
    }
  } else if (threadIdx.x < p.groups) {
    const float2 st = *reinterpret_cast<const float2*>(p.stats + (static_cast<long long>(b) * p.groups + threadIdx.x) * 2);
    g_mean[threadIdx.x] = st.x;
    g_rstd[threadIdx.x] = st.y;
  }
  __syncthreads();
  const int cq = threadIdx.x % p.CQ, py = threadIdx.x / p.CQ;
  if (py >= p.PY) return;
  const int c = cq * 4;
  float sc[4], sh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int g = (c + i) / p.cpg;
    const float ga = p.gamma[c + i] * g_rstd[g];
    sc[i] = ga;
    sh[i] = p.beta[c + i] - g_mean[g] * ga;
  }
  // a CTA walks chunks blockIdx.x, + gridDim.x, ... of its image: the grid is one resident wave, so the statistics
  // prologue above is paid once per CTA instead of once per 16-pixel chunk
  for (int ch = chunk; ch < p.nchunks; ch += gridDim.x) {
    const int pbeg = ch * p.pix_per_chunk;
    const int pend = min(p.hw, pbeg + p.pix_per_chunk);
    int pix = pbeg + py;
    for (; pix + 3 * p.PY < pend; pix += 4 * p.PY) {
      const float4 v0 = gn_load(p, b, pix, cq), v1 = gn_load(p, b, pix + p.PY, cq);
      const float4 v2 = gn_load(p, b, pix + 2 * p.PY, cq), v3 = gn_load(p, b, pix + 3 * p.PY, cq);
      gn_emit(p, b, pix, c, v0, sc, sh);
      gn_emit(p, b, pix + p.PY, c, v1, sc, sh);
      gn_emit(p, b, pix + 2 * p.PY, c, v2, sc, sh);
      gn_emit(p, b, pix + 3 * p.PY, c, v3, sc, sh);
    }
    for (; pix < pend; pix += p.PY) gn_emit(p, b, pix, c, gn_load(p, b, pix, cq), sc, sh);
  }
}

// ---------------------------------------------------------------------------------------- LayerNorm (warp per row)
template <int MAXQ>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, long long rows, int C,
                                 float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nq = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  float4 v[MAXQ];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int q = lane + i * 32;
    if (q < nq) {
      v[i] = xr[q];
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
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / C + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* orow = reinterpret_cast<uint2*>(out + row * C);
#pragma unroll
  for (int i = 0; i < MAXQ; ++i) {
    const int q = lane + i * 32;
    if (q < nq) {
      const float4 g = __ldg(g4 + q), bb = __ldg(b4 + q);
      const float y0 = (v[i].x - mean) * rstd * g.x + bb.x, y1 = (v[i].y - mean) * rstd * g.y + bb.y;
      const float y2 = (v[i].z - mean) * rstd * g.z + bb.z, y3 = (v[i].w - mean) * rstd * g.w + bb.w;
      orow[q] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    }
  }
}

// ---------------------------------------------------------------------------------------- row softmax (VAE attention)
__global__ void softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int cols, float scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* sr = s + row * cols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int i = tid * 4; i < cols; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(sr + i);
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  const float c = scale * 1.4426950408889634f;
  float sum = 0.f;
  for (int i = tid * 4; i < cols; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(sr + i);
    sum += (exp2f((v.x - m) * c) + exp2f((v.y - m) * c)) + (exp2f((v.z - m) * c) + exp2f((v.w - m) * c));
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < nw; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int i = tid * 4; i < cols; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(sr + i);
    *reinterpret_cast<uint2*>(p + row * cols + i) =
        make_uint2(pack_bf16x2(exp2f((v.x - m) * c) * inv, exp2f((v.y - m) * c) * inv),
                   pack_bf16x2(exp2f((v.z - m) * c) * inv, exp2f((v.w - m) * c) * inv));
  }
}

// ---------------------------------------------------------------------------------------- time embedding
__global__ void sinusoid_kernel(const float* __restrict__ t, float* __restrict__ out, int batch, int dim) {
  pdl_trigger();
  pdl_wait();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * half) return;
  const int b = i / half, k = i % half;
  const float f = expf(-logf(10000.0f) * static_cast<float>(k) / static_cast<float>(half));
  const float a = t[b] * f;
  out[b * dim + k] = cosf(a);  // flip_sin_to_cos: [cos | sin]
  out[b * dim + half + k] = sinf(a);
}

// y[b, n] = act(sum_k x[b,k] W[n,k] + bias[n]); one warp per output feature, fp32 weights read once.
__global__ void skinny_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ y, int batch, int K, int N,
                                     int silu_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float4* wr = reinterpret_cast<const float4*>(w + static_cast<long long>(n) * K);
  for (int b0 = 0; b0 < batch; b0 += 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int nb = min(8, batch - b0);
    for (int q = lane; q < K / 4; q += 32) {
      const float4 wv = __ldg(wr + q);
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) {
        if (bb < nb) {
          const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<long long>(b0 + bb) * K + q * 4);
          acc[bb] += (wv.x * xv.x + wv.y * xv.y) + (wv.z * xv.z + wv.w * xv.w);
        }
      }
    }
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) {
      float a = acc[bb];
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0 && bb < nb) {
        a += bias ? bias[n] : 0.f;
        y[static_cast<long long>(b0 + bb) * N + n] = silu_out ? silu_f(a) : a;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------- edge convolutions
// conv3x3 pad 1, Cin = 4: one CTA = 32 consecutive pixels of one image row; thread = output channel
// with its 36 weights in registers; the 3 x 34 x 4 input patch sits in smem and is read as broadcast
// float4s; stores are coalesced along the NHWC channel axis.
constexpr int CIN_TILE_W = 32;
__global__ void conv_small_cin_kernel(const float* __restrict__ x, int x_nchw, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ out_f32,
                                      __nv_bfloat16* __restrict__ out_bf16, int B, int H, int W, int Cout) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 patch[3][CIN_TILE_W + 2];
  const int tiles_x = (W + CIN_TILE_W - 1) / CIN_TILE_W;
  const int tx = blockIdx.x % tiles_x;
  const int yh = (blockIdx.x / tiles_x) % H;
  const int b = blockIdx.x / (tiles_x * H);
  const int x0 = tx * CIN_TILE_W;
  for (int i = threadIdx.x; i < 3 * (CIN_TILE_W + 2) * 4; i += blockDim.x) {
    const int c = i & 3, col = (i >> 2) % (CIN_TILE_W + 2), r = (i >> 2) / (CIN_TILE_W + 2);
    const int yy = yh + r - 1, xx = x0 + col - 1;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      v = x_nchw ? x[((static_cast<long long>(b) * 4 + c) * H + yy) * W + xx]
                 : x[((static_cast<long long>(b) * H + yy) * W + xx) * 4 + c];
    reinterpret_cast<float*>(&patch[r][col])[c] = v;
  }
  __syncthreads();
  const int co = threadIdx.x;
  if (co >= Cout) return;
  float4 wr[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wr[t] = __ldg(reinterpret_cast<const float4*>(w + (static_cast<long long>(co) * 9 + t) * 4));
  const float bv = bias ? bias[co] : 0.f;
  const int npx = min(CIN_TILE_W, W - x0);
  for (int px = 0; px < npx; ++px) {
    float a = bv;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float4 v = patch[dy][px + dx];
        const float4 k = wr[dy * 3 + dx];
        a += (v.x * k.x + v.y * k.y) + (v.z * k.z + v.w * k.w);
      }
    const long long off = ((static_cast<long long>(b) * H + yh) * W + x0 + px) * Cout + co;
    if (out_f32) out_f32[off] = a;
    if (out_bf16) out_bf16[off] = __float2bfloat16(a);
  }
}

// conv3x3 pad 1, tiny Cout: a warp computes CSC_PPW consecutive output pixels of one image row; lanes stride the channel
// axis (8-byte loads).  Every input pixel of the 3 x (CSC_PPW + 2) window is loaded ONCE and feeds the up to three output
// pixels it belongs to; the 3 x Cout weight vectors of a tap row live in registers across the window (the first version
// walked pixel by pixel: 72 instead of 30 loads and 216 instead of 27 smem weight reads per 8 pixels, 1.1 ms for the VAE's
// 512 x 512 conv_out, 0.73 ms now; persistent CTAs that stage the weights once were slower, 0.92 ms).
constexpr int CSC_PPW = 8;
template <int COUT>
__global__ void __launch_bounds__(256) conv_small_cout_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out, int postprocess,
                                                              int B, int H, int W, int Cin) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sw[];  // [COUT][9][Cin]
  for (int i = threadIdx.x * 4; i < COUT * 9 * Cin; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(sw + i) = *reinterpret_cast<const float4*>(w + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int segs = (W + CSC_PPW - 1) / CSC_PPW;                       // row segments per image row
  const long long seg = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seg >= static_cast<long long>(B) * H * segs) return;
  const int x0 = static_cast<int>(seg % segs) * CSC_PPW;
  const int yh = static_cast<int>((seg / segs) % H);
  const int b = static_cast<int>(seg / (static_cast<long long>(segs) * H));
  float acc[CSC_PPW][COUT];
#pragma unroll
  for (int i = 0; i < CSC_PPW; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[i][o] = 0.f;
  for (int c = lane * 4; c < Cin; c += 128) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = yh + r - 1;
      if (yy < 0 || yy >= H) continue;     // warp-uniform
      float4 wv[COUT][3];
#pragma unroll
      for (int o = 0; o < COUT; ++o)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) wv[o][dx] = *reinterpret_cast<const float4*>(sw + (o * 9 + r * 3 + dx) * Cin + c);
      const __nv_bfloat16* xr = x + ((static_cast<long long>(b) * H + yy) * W) * Cin + c;
#pragma unroll
      for (int j = 0; j < CSC_PPW + 2; ++j) {
        const int xx = x0 - 1 + j;
        if (xx < 0 || xx >= W) continue;   // warp-uniform (zero padding)
        const uint2 raw = *reinterpret_cast<const uint2*>(xr + static_cast<long long>(xx) * Cin);
        const __nv_bfloat162 p0 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
        const __nv_bfloat162 p1 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        const float v0 = __low2float(p0), v1 = __high2float(p0), v2 = __low2float(p1), v3 = __high2float(p1);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int i = j - dx;            // output pixel x0 + i reads input x0 + i + dx - 1 = x0 - 1 + j with tap column dx
          if (i < 0 || i >= CSC_PPW) continue;   // (compile time after unrolling)
#pragma unroll
          for (int o = 0; o < COUT; ++o)
            acc[i][o] += (v0 * wv[o][dx].x + v1 * wv[o][dx].y) + (v2 * wv[o][dx].z + v3 * wv[o][dx].w);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < CSC_PPW; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o)
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) acc[i][o] += __shfl_xor_sync(0xffffffffu, acc[i][o], sft);
  // lane l < CSC_PPW * COUT writes one value: NHWC image (postprocess) -> l = pixel * COUT + channel, contiguous;
  // NCHW -> l = channel * CSC_PPW + pixel, contiguous per channel
  float v = 0.f;
  int pi = 0, po = 0;
#pragma unroll
  for (int i = 0; i < CSC_PPW; ++i)
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      const int l = postprocess ? i * COUT + o : o * CSC_PPW + i;
      if (lane == l) v = acc[i][o], pi = i, po = o;
    }
  if (lane < CSC_PPW * COUT && x0 + pi < W) {
    v += bias ? bias[po] : 0.f;
    if (postprocess) {
      v = fminf(fmaxf(v * 0.5f + 0.5f, 0.f), 1.f);
      out[((static_cast<long long>(b) * H + yh) * W + x0 + pi) * COUT + po] = v;   // NHWC image
    } else {
      out[((static_cast<long long>(b) * COUT + po) * H + yh) * W + x0 + pi] = v;   // NCHW
    }
  }
}

__global__ void upsample2x_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                  int C) {
  pdl_trigger();
  pdl_wait();
  const int cq_n = C / 4;
  const long long total = static_cast<long long>(B) * 2 * H * 2 * W * cq_n;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(idx % cq_n);
    long long pix = idx / cq_n;
    const int xo = static_cast<int>(pix % (2 * W));
    pix /= 2 * W;
    const int yo = static_cast<int>(pix % (2 * H));
    const int b = static_cast<int>(pix / (2 * H));
    const float4 v = *reinterpret_cast<const float4*>(x + ((static_cast<long long>(b) * H + (yo >> 1)) * W + (xo >> 1)) * C + cq * 4);
    *reinterpret_cast<uint2*>(out + idx * 4) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n4) {
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

__global__ void vae_latent_prep_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                       const float* __restrict__ bias, float inv_scaling, float* __restrict__ out, int B,
                                       int hw) {
  pdl_trigger();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * hw) return;
  const int b = static_cast<int>(i / hw), pix = static_cast<int>(i % hw);
  float v[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) v[c] = z[(static_cast<long long>(b) * 4 + c) * hw + pix] * inv_scaling;
  float o[4];
#pragma unroll
  for (int oc = 0; oc < 4; ++oc) {
    float a = bias[oc];
#pragma unroll
    for (int c = 0; c < 4; ++c) a += w[oc * 4 + c] * v[c];
    o[oc] = a;
  }
  *reinterpret_cast<float4*>(out + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// conv_in operand: latents fp32 NCHW [B,4,H,W] -> bf16 NHWC [B,H,W,64]: channels 0-3 = hi(x), 4-7 = lo(x) = bf16(x - hi),
// 8-11 = hi(x), rest zero.  Against weights packed [w_hi | w_hi | w_lo] the tensor-core conv computes
// x_hi w_hi + x_lo w_hi + x_hi w_lo, i.e. the fp32 product to ~2^-17 (the 4-channel conv stays at full precision).
__global__ void latent_operand_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int hw) {
  pdl_trigger();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;   // (pixel, 16-byte chunk of the 128-byte row)
  if (i >= static_cast<long long>(B) * hw * 8) return;
  const int chunk = static_cast<int>(i & 7);
  const long long pixi = i >> 3;
  uint4 w = make_uint4(0u, 0u, 0u, 0u);
  if (chunk < 2) {
    const int b = static_cast<int>(pixi / hw), pix = static_cast<int>(pixi % hw);
    float v[4], hi[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      v[c] = x[(static_cast<long long>(b) * 4 + c) * hw + pix];
      hi[c] = __bfloat162float(__float2bfloat16(v[c]));
    }
    if (chunk == 0) w = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]),
                                   pack_bf16x2(v[0] - hi[0], v[1] - hi[1]), pack_bf16x2(v[2] - hi[2], v[3] - hi[3]));
    else w = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), 0u, 0u);
  }
  *reinterpret_cast<uint4*>(out + pixi * 64 + chunk * 8) = w;
}

// ---------------------------------------------------------------------------------------- CFG + DDPM step
__global__ void cfg_ddpm_step_kernel(const float* __restrict__ eps2, const float* __restrict__ x,
                                     const float* __restrict__ noise, const float* __restrict__ coef, float gs,
                                     int use_cfg, int vpred, float* __restrict__ x_prev, float* __restrict__ x0_out,
                                     long long n) {
  pdl_trigger();
  pdl_wait();
  const float sa = coef[0], sb = coef[1], c0 = coef[2], ct = coef[3], sigma = coef[4];
  const float inv_sa = 1.0f / sa;
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 eu = reinterpret_cast<const float4*>(eps2)[i];
    float e[4] = {eu.x, eu.y, eu.z, eu.w};
    if (use_cfg) {
      const float4 ec = reinterpret_cast<const float4*>(eps2 + n)[i];
      e[0] += gs * (ec.x - eu.x), e[1] += gs * (ec.y - eu.y), e[2] += gs * (ec.z - eu.z), e[3] += gs * (ec.w - eu.w);
    }
    const float4 xv4 = reinterpret_cast<const float4*>(x)[i];
    const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
    float nz[4] = {0, 0, 0, 0};
    if (noise != nullptr) {
      const float4 t = reinterpret_cast<const float4*>(noise)[i];
      nz[0] = t.x, nz[1] = t.y, nz[2] = t.z, nz[3] = t.w;
    }
    float x0[4], xp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      x0[k] = vpred ? (sa * xv[k] - sb * e[k]) : (xv[k] - sb * e[k]) * inv_sa;
      xp[k] = c0 * x0[k] + ct * xv[k] + sigma * nz[k];
    }
    reinterpret_cast<float4*>(x_prev)[i] = make_float4(xp[0], xp[1], xp[2], xp[3]);
    if (x0_out) reinterpret_cast<float4*>(x0_out)[i] = make_float4(x0[0], x0[1], x0[2], x0[3]);
  }
}

// ---------------------------------------------------------------------------------------- IResNet glue
// y[b, yo, xo, c] = bf16(x[b, s*yo, s*xo, c] * scale[c] + shift[c])   (eval-mode BatchNorm2d in front of a conv,
// and/or the stride-s sampling of a 1x1 stride-s shortcut conv; scale / shift may be NULL = plain cast)
__global__ void channel_affine_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                      const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int f16, int B,
                                      int H, int W, int C, int stride) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / stride, Wo = W / stride, CQ = C / 4;
  const long long total = static_cast<long long>(B) * Ho * Wo * CQ;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(i % CQ);
    long long pix = i / CQ;
    const int xo = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int yo = static_cast<int>(pix % Ho);
    const int b = static_cast<int>(pix / Ho);
    const float4 v = *reinterpret_cast<const float4*>(x + ((static_cast<long long>(b) * H + yo * stride) * W + xo * stride) * C + cq * 4);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) sc = *reinterpret_cast<const float4*>(scale + cq * 4);
    if (shift) sh = *reinterpret_cast<const float4*>(shift + cq * 4);
    *reinterpret_cast<uint2*>(out + i * 4) = make_uint2(pack_16x2(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), f16 != 0),
                                                        pack_16x2(fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w), f16 != 0));
  }
}

// ArcFace input from decoded images (train_ID-Booth.py:433-455): crop [y0:y1, x0:x1] of the [0,1] NHWC image,
// bilinear resize (align_corners = False, no antialias = torchvision resize(antialias=None)) to S x S,
// ((v * 255) / 255 - 0.5) / 0.5, written as the bf16 NHWC stem operand with C_pad channels (3 real, rest 0).
__global__ void crop_resize_norm_kernel(const float* __restrict__ img, const int* __restrict__ bbox, __nv_bfloat16* __restrict__ out,
                                        int f16, int n, int H, int W, int S, int c_pad) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(n) * S * S;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xo = static_cast<int>(i % S), yo = static_cast<int>((i / S) % S), b = static_cast<int>(i / (static_cast<long long>(S) * S));
    const int x0 = max(0, bbox[4 * b]), y0 = max(0, bbox[4 * b + 1]);
    const int x1 = min(bbox[4 * b + 2], W), y1 = min(bbox[4 * b + 3], H);
    const int cw = x1 - x0, ch = y1 - y0;
    // PyTorch upsample_bilinear2d, align_corners = False: src = max(0, (dst + 0.5) * scale - 0.5)
    const float sx = fmaxf(0.f, (xo + 0.5f) * (static_cast<float>(cw) / S) - 0.5f);
    const float sy = fmaxf(0.f, (yo + 0.5f) * (static_cast<float>(ch) / S) - 0.5f);
    const int ix = min(static_cast<int>(sx), cw - 1), iy = min(static_cast<int>(sy), ch - 1);
    const int ix1 = min(ix + 1, cw - 1), iy1 = min(iy + 1, ch - 1);
    const float lx = sx - ix, ly = sy - iy;
    const float* base = img + static_cast<long long>(b) * H * W * 3;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p00 = base[((y0 + iy) * static_cast<long long>(W) + x0 + ix) * 3 + c];
      const float p01 = base[((y0 + iy) * static_cast<long long>(W) + x0 + ix1) * 3 + c];
      const float p10 = base[((y0 + iy1) * static_cast<long long>(W) + x0 + ix) * 3 + c];
      const float p11 = base[((y0 + iy1) * static_cast<long long>(W) + x0 + ix1) * 3 + c];
      const float r = (1.f - ly) * ((1.f - lx) * p00 + lx * p01) + ly * ((1.f - lx) * p10 + lx * p11);
      v[c] = (r - 0.5f) / 0.5f;
    }
    unsigned short* o = reinterpret_cast<unsigned short*>(out) + i * c_pad;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<unsigned short>(pack_16x2(v[c], 0.f, f16 != 0) & 0xffffu);
    for (int c = 3; c < c_pad; ++c) o[c] = 0;
  }
}

static int grid_for(long long work_items, int block, int max_blocks) {
  long long g = (work_items + block - 1) / block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace idb

using namespace idb;

// workspace = [partials: B*GN_MAX_CHUNKS*G*2 floats][stats: B*G*2 floats][counters: B uint32]
extern "C" size_t idb_groupnorm_workspace_bytes(int32_t batch, int32_t groups) {
  return (static_cast<size_t>(batch) * GN_MAX_CHUNKS * groups * 2 + static_cast<size_t>(batch) * groups * 2 + batch) *
         sizeof(float);
}

extern "C" int idb_groupnorm(const idb_groupnorm_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return fail(IDB_E_BADARG, "idb_groupnorm: null args");
  if (int rc = require_sm100()) return rc;
  const int C = a->c0 + (a->x1 ? a->c1 : 0);
  if (!a->x0 || !a->gamma || !a->beta || !a->out_norm || !a->partials) return fail(IDB_E_BADARG, "idb_groupnorm: null pointer");
  if (a->c0 % 4 || (a->x1 && a->c1 % 4) || C % a->groups || a->groups > GN_MAX_GROUPS || a->groups <= 0)
    return fail(IDB_E_BADARG, "idb_groupnorm: channels must be multiples of 4 and divisible by groups (<= 64)");
  if (C / 4 > 1024) return fail(IDB_E_BADARG, "idb_groupnorm: C too large");
  GnParams p;
  p.x0 = a->x0, p.x1 = a->x1, p.c0 = a->c0, p.c1 = a->x1 ? a->c1 : 0, p.C = C, p.CQ = C / 4;
  p.PY = p.CQ >= 256 ? 1 : 256 / p.CQ;
  p.hw = a->hw, p.groups = a->groups, p.cpg = C / a->groups, p.eps = a->eps;
  p.b1_mod = 0;
  if (a->x1 && a->x1_batch > 0 && a->x1_batch != a->batch) {
    if (a->batch % a->x1_batch) return fail(IDB_E_BADARG, "idb_groupnorm: x1_batch must divide batch");
    if (a->x1_stats != nullptr) return fail(IDB_E_UNSUPPORTED, "idb_groupnorm: a shared source 1 (x1_batch) takes per-image sums or no statistics, not row-block sums");
    p.b1_mod = a->x1_batch;
  }
  const bool have_stats = a->x0_stats != nullptr && (a->x1 == nullptr || a->x1_stats != nullptr) && a->hw % 32 == 0 &&
                          (a->x0_stats_phases <= 1 || (a->x0_stats_phases == 4 && a->hw % 128 == 0));
  // small rasters: the apply CTAs finalize the statistics themselves (one launch instead of two).  Measured on the UNet
  // step (B = 8, same box): 8x8 rasters only (2 row blocks) 8.851 -> 8.745 ms; extending it to 16x16 (8 row blocks,
  // 164 KB of sums re-read per CTA at 2560 channels) gives the gain back, so the default stops at 2.
  // IDB_GN_FUSE_RB = 0 disables, up to GN_FUSE_MAX_RB widens.
  static const int fuse_rb = getenv("IDB_GN_FUSE_RB") ? atoi(getenv("IDB_GN_FUSE_RB")) : 2;
  const bool fused = have_stats && p.PY == 1 && p.cpg % 4 == 0 && p.CQ <= GN_FUSE_MAX_CQ && p.CQ >= a->groups &&
                     a->hw / 32 <= fuse_rb && a->hw / 32 <= GN_FUSE_MAX_RB;
  // pixels per CTA: enough for >= 4 unrolled rounds of PY rows, while keeping >= ~4 CTAs per SM in flight
  int ppc = 16 * p.PY;
  while (ppc > 4 * p.PY && static_cast<long long>(a->batch) * ((a->hw + ppc - 1) / ppc) < 4LL * num_sms()) ppc /= 2;
  if (fused && ppc < 2 * (a->hw / 32)) ppc = 2 * (a->hw / 32);   // every CTA re-reads the image's row-block sums: keep that below its own pixel traffic
  int nchunks = (a->hw + ppc - 1) / ppc;
  if (nchunks > GN_MAX_CHUNKS) nchunks = GN_MAX_CHUNKS;
  p.pix_per_chunk = (a->hw + nchunks - 1) / nchunks;
  p.nchunks = (a->hw + p.pix_per_chunk - 1) / p.pix_per_chunk;
  p.gamma = a->gamma, p.beta = a->beta, p.silu = a->silu;
  p.out_norm = static_cast<__nv_bfloat16*>(a->out_norm);
  p.out_raw = static_cast<__nv_bfloat16*>(a->out_raw);
  p.partials = a->partials;
  p.stats = a->partials + static_cast<size_t>(a->batch) * GN_MAX_CHUNKS * a->groups * 2;
  p.counters = reinterpret_cast<unsigned int*>(p.stats + static_cast<size_t>(a->batch) * a->groups * 2);
  int threads = p.CQ * p.PY;
  threads = (threads + 31) / 32 * 32;
  if (threads < 64) threads = 64;  // the apply kernel's first `groups` threads publish mean / rstd
  dim3 grid(p.nchunks, a->batch);
  p.rb0 = p.rb1 = nullptr, p.rbpi = a->hw / 32, p.phases0 = a->x0_stats_phases;
  p.sums0 = p.sums1 = nullptr;
  if (a->x0_sums != nullptr && (a->x1 == nullptr || a->x1_sums != nullptr)) {
    p.gran = a->sums_gran > 0 ? a->sums_gran : 1;
    if (p.cpg % p.gran || p.c0 % p.gran || p.c1 % p.gran)
      return fail(IDB_E_BADARG, "idb_groupnorm: sums_gran must divide the channels per group and both source widths");
    p.sums0 = reinterpret_cast<const longlong2*>(a->x0_sums);
    p.sums1 = reinterpret_cast<const longlong2*>(a->x1_sums);
    p.inv_n = 1.0 / (static_cast<double>(a->hw) * p.cpg);
  } else if (fused) {
    p.rb0 = reinterpret_cast<const float2*>(a->x0_stats);
    p.rb1 = reinterpret_cast<const float2*>(a->x1_stats);
  } else if (have_stats) {   // statistics already produced by the GEMM epilogues of the sources
    launch_pdl(gn_finalize_kernel, dim3(dim3(a->groups, a->batch)), dim3(256), 0, stream, 
        reinterpret_cast<const float2*>(a->x0_stats), p.c0, reinterpret_cast<const float2*>(a->x1_stats), p.c1,
        a->hw / 32, p.cpg, a->hw, a->eps, p.stats, a->x0_stats_phases);
    IDB_CHECK_LAUNCH("gn_finalize");
  } else {
    const size_t stats_smem = static_cast<size_t>(2) * p.PY * C * sizeof(float);
    launch_pdl(gn_stats_kernel, dim3(grid), dim3(threads), stats_smem, stream, p);
    IDB_CHECK_LAUNCH("gn_stats");
  }
  // apply: one resident wave (64 registers per thread), CTAs stride over the chunks of their image
  static const int apply_waves = getenv("IDB_GN_APPLY_WAVES") ? atoi(getenv("IDB_GN_APPLY_WAVES")) : 1;   // 0 = one CTA per chunk (profiling)
  int gx = p.nchunks;
  if (apply_waves > 0) {
    const int resident = 65536 / (64 * threads) > 0 ? 65536 / (64 * threads) : 1;
    const long long slots = static_cast<long long>(num_sms()) * resident * apply_waves;
    const int per_image = static_cast<int>((slots + a->batch - 1) / a->batch);
    if (per_image < gx) gx = per_image < 1 ? 1 : per_image;
  }
  launch_pdl(gn_apply_kernel, dim3(dim3(gx, a->batch)), dim3(threads), 0, stream, p);
  IDB_CHECK_LAUNCH("gn_apply");
  return IDB_OK;
}

extern "C" int idb_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16, int64_t rows,
                             int32_t c, float eps, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!x || !gamma || !beta || !out_bf16) return fail(IDB_E_BADARG, "idb_layernorm: null pointer");
  if (c % 4 || c <= 0 || c > 2048) return fail(IDB_E_BADARG, "idb_layernorm: C must be a multiple of 4, <= 2048");
  const int warps = 8;
  const int grid = static_cast<int>((rows + warps - 1) / warps);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_bf16);
  const int quads_per_lane = (c / 4 + 31) / 32;  // registers (and therefore occupancy) scale with this
  if (quads_per_lane <= 3) launch_pdl(layernorm_kernel<3>, dim3(grid), dim3(warps * 32), 0, stream, x, gamma, beta, o, rows, c, eps);
  else if (quads_per_lane <= 5) launch_pdl(layernorm_kernel<5>, dim3(grid), dim3(warps * 32), 0, stream, x, gamma, beta, o, rows, c, eps);
  else if (quads_per_lane <= 10) launch_pdl(layernorm_kernel<10>, dim3(grid), dim3(warps * 32), 0, stream, x, gamma, beta, o, rows, c, eps);
  else launch_pdl(layernorm_kernel<16>, dim3(grid), dim3(warps * 32), 0, stream, x, gamma, beta, o, rows, c, eps);
  IDB_CHECK_LAUNCH("layernorm");
  return IDB_OK;
}

extern "C" int idb_softmax_rows(const float* s, void* p_bf16, int64_t rows, int32_t cols, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!s || !p_bf16 || cols % 4 || rows <= 0) return fail(IDB_E_BADARG, "idb_softmax_rows: bad arguments");
  launch_pdl(softmax_rows_kernel, dim3(static_cast<unsigned>(rows)), dim3(256), 0, stream, s, static_cast<__nv_bfloat16*>(p_bf16), cols, scale);
  IDB_CHECK_LAUNCH("softmax_rows");
  return IDB_OK;
}

extern "C" int idb_time_embed(const idb_time_embed_args* a, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (a == nullptr) return fail(IDB_E_BADARG, "idb_time_embed: null args");
  if (int rc = require_sm100()) return rc;
  if (a->dim_sin % 4 || a->dim_emb % 4 || a->batch <= 0) return fail(IDB_E_BADARG, "idb_time_embed: bad dims");
  float* sin_buf = a->scratch;
  float* h1 = sin_buf + static_cast<long long>(a->batch) * a->dim_sin;
  float* se = h1 + static_cast<long long>(a->batch) * a->dim_emb;
  const int half_total = a->batch * a->dim_sin / 2;
  launch_pdl(sinusoid_kernel, dim3((half_total + 127) / 128), dim3(128), 0, stream, a->timesteps, sin_buf, a->batch, a->dim_sin);
  IDB_CHECK_LAUNCH("sinusoid");
  const int wpb = 8;
  launch_pdl(skinny_linear_kernel, dim3((a->dim_emb + wpb - 1) / wpb), dim3(wpb * 32), 0, stream, sin_buf, a->w1, a->b1, h1, a->batch,
                                                                              a->dim_sin, a->dim_emb, 1);
  IDB_CHECK_LAUNCH("time linear_1");
  launch_pdl(skinny_linear_kernel, dim3((a->dim_emb + wpb - 1) / wpb), dim3(wpb * 32), 0, stream, h1, a->w2, a->b2, se, a->batch, a->dim_emb,
                                                                              a->dim_emb, 1);
  IDB_CHECK_LAUNCH("time linear_2");
  launch_pdl(skinny_linear_kernel, dim3((a->n_all + wpb - 1) / wpb), dim3(wpb * 32), 0, stream, se, a->w_all, a->b_all, a->proj_out, a->batch,
                                                                            a->dim_emb, a->n_all, 0);
  IDB_CHECK_LAUNCH("time_emb_proj");
  return IDB_OK;
}

extern "C" int idb_conv3x3_small_cin(const float* x, int32_t x_nchw, const float* w, const float* bias, float* out_f32,
                                     void* out_bf16, int32_t batch, int32_t h, int32_t wd, int32_t cin, int32_t cout,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!x || !w || (!out_f32 && !out_bf16)) return fail(IDB_E_BADARG, "idb_conv3x3_small_cin: null pointer");
  if (cin != 4 || cout > 1024) return fail(IDB_E_UNSUPPORTED, "idb_conv3x3_small_cin: Cin must be 4, Cout <= 1024");
  const long long blocks = static_cast<long long>(batch) * h * ((wd + CIN_TILE_W - 1) / CIN_TILE_W);
  const int threads = (cout + 31) / 32 * 32;
  launch_pdl(conv_small_cin_kernel, dim3(static_cast<unsigned>(blocks)), dim3(threads), 0, stream, 
      x, x_nchw, w, bias, out_f32, static_cast<__nv_bfloat16*>(out_bf16), batch, h, wd, cout);
  IDB_CHECK_LAUNCH("conv_small_cin");
  return IDB_OK;
}

extern "C" int idb_conv3x3_small_cout(const void* x_bf16, const float* w, const float* bias, float* out,
                                      int32_t postprocess, int32_t batch, int32_t h, int32_t wd, int32_t cin,
                                      int32_t cout, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!x_bf16 || !w || !out) return fail(IDB_E_BADARG, "idb_conv3x3_small_cout: null pointer");
  if (cin % 4 || (cout != 3 && cout != 4)) return fail(IDB_E_UNSUPPORTED, "idb_conv3x3_small_cout: Cout in {3,4}, Cin % 4 == 0");
  const size_t smem = static_cast<size_t>(cout) * 9 * cin * sizeof(float);
  if (smem > 48 * 1024) return fail(IDB_E_UNSUPPORTED, "idb_conv3x3_small_cout: weights exceed 48 KiB of shared memory");
  const long long nseg = static_cast<long long>(batch) * h * ((wd + CSC_PPW - 1) / CSC_PPW);   // one warp per row segment
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((nseg + wpb - 1) / wpb);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x_bf16);
  if (cout == 4)
    launch_pdl(conv_small_cout_kernel<4>, dim3(grid), dim3(wpb * 32), smem, stream, xb, w, bias, out, postprocess, batch, h, wd, cin);
  else
    launch_pdl(conv_small_cout_kernel<3>, dim3(grid), dim3(wpb * 32), smem, stream, xb, w, bias, out, postprocess, batch, h, wd, cin);
  IDB_CHECK_LAUNCH("conv_small_cout");
  return IDB_OK;
}

extern "C" int idb_upsample2x(const float* x, void* out_bf16, int32_t batch, int32_t h, int32_t wd, int32_t c,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!x || !out_bf16 || c % 4) return fail(IDB_E_BADARG, "idb_upsample2x: bad arguments");
  const long long total = static_cast<long long>(batch) * 4 * h * wd * (c / 4);
  launch_pdl(upsample2x_kernel, dim3(grid_for(total, 256, num_sms() * 16)), dim3(256), 0, stream, x, static_cast<__nv_bfloat16*>(out_bf16),
                                                                               batch, h, wd, c);
  IDB_CHECK_LAUNCH("upsample2x");
  return IDB_OK;
}

extern "C" int idb_cast_bf16(const float* x, void* out_bf16, int64_t n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!x || !out_bf16 || n % 4) return fail(IDB_E_BADARG, "idb_cast_bf16: bad arguments");
  launch_pdl(cast_bf16_kernel, dim3(grid_for(n / 4, 256, num_sms() * 16)), dim3(256), 0, stream, x, static_cast<__nv_bfloat16*>(out_bf16), n / 4);
  IDB_CHECK_LAUNCH("cast_bf16");
  return IDB_OK;
}

extern "C" int idb_vae_latent_prep(const float* z_nchw, const float* w, const float* bias, float inv_scaling,
                                   float* out_nhwc, int32_t batch, int32_t hw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!z_nchw || !w || !bias || !out_nhwc) return fail(IDB_E_BADARG, "idb_vae_latent_prep: null pointer");
  const long long total = static_cast<long long>(batch) * hw;
  launch_pdl(vae_latent_prep_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, z_nchw, w, bias, inv_scaling,
                                                                                         out_nhwc, batch, hw);
  IDB_CHECK_LAUNCH("vae_latent_prep");
  return IDB_OK;
}

extern "C" int idb_latent_operand(const float* x_nchw, void* out_bf16_nhwc64, int32_t batch, int32_t hw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!x_nchw || !out_bf16_nhwc64 || batch <= 0 || hw <= 0) return fail(IDB_E_BADARG, "idb_latent_operand: bad arguments");
  const long long total = static_cast<long long>(batch) * hw * 8;
  launch_pdl(latent_operand_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, x_nchw,
             static_cast<__nv_bfloat16*>(out_bf16_nhwc64), batch, hw);
  IDB_CHECK_LAUNCH("latent_operand");
  return IDB_OK;
}

extern "C" int idb_cfg_ddpm_step(const float* eps2, const float* x, const float* noise, const float* coef,
                                 float guidance_scale, int32_t use_cfg, int32_t v_prediction, float* x_prev,
                                 float* x0_out, int64_t n_per_branch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!eps2 || !x || !coef || !x_prev) return fail(IDB_E_BADARG, "idb_cfg_ddpm_step: null pointer");
  if (n_per_branch <= 0 || n_per_branch % 4) return fail(IDB_E_BADARG, "idb_cfg_ddpm_step: n must be a positive multiple of 4");
  launch_pdl(cfg_ddpm_step_kernel, dim3(grid_for(n_per_branch / 4, 256, num_sms() * 8)), dim3(256), 0, stream, 
      eps2, x, noise, coef, guidance_scale, use_cfg, v_prediction, x_prev, x0_out, n_per_branch);
  IDB_CHECK_LAUNCH("cfg_ddpm_step");
  return IDB_OK;
}

extern "C" int idb_channel_affine(const float* x, const float* scale, const float* shift, void* out_bf16, int32_t out_f16,
                                  int32_t batch, int32_t h, int32_t wd, int32_t c, int32_t stride, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!x || !out_bf16 || c % 4 || stride < 1 || h % stride || wd % stride)
    return fail(IDB_E_BADARG, "idb_channel_affine: bad arguments (C % 4 == 0, H and W divisible by stride)");
  const long long total = static_cast<long long>(batch) * (h / stride) * (wd / stride) * (c / 4);
  launch_pdl(channel_affine_kernel, dim3(grid_for(total, 256, num_sms() * 16)), dim3(256), 0, stream, 
      x, scale, shift, static_cast<__nv_bfloat16*>(out_bf16), out_f16, batch, h, wd, c, stride);
  IDB_CHECK_LAUNCH("channel_affine");
  return IDB_OK;
}

extern "C" int idb_crop_resize_norm(const float* img_nhwc, const int32_t* bbox_xyxy, void* out_bf16, int32_t out_f16, int32_t n,
                                    int32_t h, int32_t wd, int32_t size, int32_t c_pad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = require_sm100()) return rc;
  if (!img_nhwc || !bbox_xyxy || !out_bf16 || n <= 0 || size <= 0 || c_pad < 3)
    return fail(IDB_E_BADARG, "idb_crop_resize_norm: bad arguments");
  const long long total = static_cast<long long>(n) * size * size;
  launch_pdl(crop_resize_norm_kernel, dim3(grid_for(total, 128, num_sms() * 8)), dim3(128), 0, stream, 
      img_nhwc, bbox_xyxy, static_cast<__nv_bfloat16*>(out_bf16), out_f16, n, h, wd, size, c_pad);
  IDB_CHECK_LAUNCH("crop_resize_norm");
  return IDB_OK;
}
