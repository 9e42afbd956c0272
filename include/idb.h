/* idb.h -- C ABI of libidb_b200.so: the B200 (sm_100a) kernels behind the ID-Booth
 * Stable Diffusion 2.1 denoising hot path.
 *
 * The reference (rangasaishreyas/FacePoseGenerator) is pure Python and reaches this
 * arithmetic through torch ops inside diffusers==0.32.2 (requirements.txt:4); it has
 * no FFI of its own.  Each entry point below therefore names the reference call site
 * (file:line under /root/reference) and the diffusers module whose torch-op sequence
 * it replaces.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless noted;
 *    the caller owns every buffer (including workspaces) -- the library never
 *    allocates or frees device memory and keeps no reference past return.
 *  - every function enqueues on `stream` (a cudaStream_t passed as void*), never
 *    synchronises, and is CUDA-graph capturable.
 *  - return 0 on success, a negative IDB_E_* code on error; idb_last_error() gives
 *    the message of the calling thread's last failure.  No C++ exception crosses
 *    the ABI.  A device that is not sm_100 is an error, never a fallback.
 *  - activations: NHWC.  "stream" tensors (the residual stream) are fp32, GEMM/conv
 *    operands are bf16.  Weights are pre-packed once at load (see each call).
 */
#ifndef IDB_H_
#define IDB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDB_VERSION 100

enum {
  IDB_OK = 0,
  IDB_E_BADARG = -1,   /* shape / alignment / null pointer */
  IDB_E_ARCH = -2,     /* device is not compute capability 10.x */
  IDB_E_CUDA = -3,     /* CUDA runtime / driver error at launch */
  IDB_E_UNSUPPORTED = -4
};

int idb_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf. */
int idb_last_error(char* buf, size_t n);
/* 0 if the current device is sm_100 (B200), IDB_E_ARCH otherwise. */
int idb_device_check(void);
int idb_num_sms(void);
/* Kernels this library has launched (or captured into a CUDA graph) so far in this process: monotonic, exact. */
uint64_t idb_launch_count(void);
/* idb_gemm_conv calls so far that ran the stream-K schedule (a diagnostic for tests and benchmarks). */
uint64_t idb_stream_k_launch_count(void);
/* How stream-K kernels are launched in this process: 0 = cooperative + programmatic dependent launch, 1 = cooperative
 * only (the driver refused the combination), 2 = cooperative launch unavailable, stream-K disabled. */
int idb_stream_k_mode(void);
/* sizeof of the argument structs below as THIS build sees them (0 = idb_gemm_conv_args, 1 = idb_attention_args,
 * 2 = idb_groupnorm_args, 3 = idb_time_embed_args, 4 = idb_attention_bwd_args, 5 = idb_groupnorm_bwd_args): a binding
 * checks its own layout against it at load time. */
size_t idb_sizeof_args(int32_t which);

/* ------------------------------------------------------------------------------------------
 * idb_gemm_conv: D[M,N] = epilogue( sum_seg im2col(A_seg)[M,K_seg] . W[N, K]^T )
 * tcgen05/TMEM implicit GEMM fed by TMA.  One kernel covers
 *   - nn.Linear            (diffusers Attention.to_q/k/v/to_out, FeedForward, proj_in/out;
 *                           reached from inference_ID-Booth.py:138)
 *   - peft lora.Linear     (unmerged rank-r delta fused in; inference_ID-Booth.py:107,
 *                           train_ID-Booth.py:672-678)
 *   - nn.Conv2d 3x3 s1/s2, 1x1 (ResnetBlock2D.conv1/conv2/conv_shortcut, Down/Upsample2D)
 * A operands are bf16 NHWC images [B,H,W,C] (a Linear is the 1x1 "image" [1,1,M,K]);
 * up to two K segments are concatenated (segment 1 is always a 1x1 tap: the fused
 * conv_shortcut, or nothing).  W is bf16 [N, K_total] row-major with K ordered
 * (tap-major, channel-minor) per segment.  C of every segment must be a multiple of 64,
 * N a multiple of 32.
 * ---------------------------------------------------------------------------------------- */
enum { IDB_A_1X1 = 0, IDB_A_3X3 = 1, IDB_A_3X3_S2 = 2,
       IDB_A_3X3_S2_ASYM = 3, /* stride 2 with the (0,1,0,1) right/bottom padding of the VAE encoder's Downsample2D: in(2y+dy, 2x+dx) */
       IDB_A_2X2 = 4          /* 2x2 taps at in(y + dy + tap_off_y, x + dx + tap_off_x), dy, dx in {0,1}, offsets in {-1,0}: one output
                                 parity class of nearest-2x-upsample + conv3x3 (Upsample2D) evaluated on the LOW-resolution input
                                 with the 3x3 taps that hit the same input pixel pre-summed (4 GEMMs with K = 4C instead of one
                                 with K = 9C on the upsampled tensor) */ };
enum {
  IDB_EPI_GEGLU = 1, /* W rows interleaved in 16-blocks [a(16) | g(16)]; out[:, j] = a_j * gelu_erf(g_j); N_out = N/2 */
  IDB_EPI_GELU = 4,  /* out = gelu_erf(acc + bias) (CLIP text MLP fc1); not combined with GEGLU */
  IDB_EPI_F16 = 2,   /* the 16-bit tensors of this call (a0, a1, w, out_bf16) are IEEE fp16 instead of bf16 (the ArcFace IResNet
                        runs under fp16 autocast in the reference, iresnet.py:149); not combined with LoRA */
  IDB_EPI_PHASES4 = 8 /* IDB_A_2X2 with out_scale = 2 only: ALL FOUR output parity classes of an Upsample2D in one call.  W holds
                        the four phase weight matrices stacked on N ([4 * N_out, 4 * C0], phase = 2 * py + px major), n = 4 * N_out;
                        the tile of phase (py, px) reads its taps at offsets (py - 1, px - 1) and writes output pixels
                        (2y + py, 2x + px); tap_off_* / out_phase_* are ignored; bias [N_out] is shared by the phases;
                        stats_partials is filled for all four phases, stats_image_sums accumulates over them */
};

typedef struct {
  /* segment 0 */
  const void* a0;      /* bf16 [B, H, W, C0] */
  int32_t a0_mode;     /* IDB_A_* */
  int32_t c0;
  /* segment 1 (optional, 1x1, same OUTPUT geometry) */
  const void* a1;      /* bf16 [B, Ho, Wo, C1] or NULL */
  int32_t c1;
  /* geometry of the INPUT image of segment 0 */
  int32_t batch, height, width;
  /* weights */
  const void* w;       /* bf16 [N, K_total], K_total = taps0*C0 + C1 */
  int32_t n;
  /* epilogue inputs (all optional) */
  const float* bias;    /* [N] */
  const float* rowvec;  /* [batch, N]  added per image (ResnetBlock2D time_emb_proj term) */
  int64_t rowvec_ld;    /* elements between consecutive images of rowvec (0 => N) */
  const float* residual;/* fp32 [M, N_out] */
  /* fused LoRA (optional): down is bf16 [n_seg*16, K] (each adapter's A zero-padded to 16
   * rows; segment s = column / lora_seg_n), up is bf16 [N, 64] (row n = B[n, :] * scale of
   * its segment's adapter in columns [0, rank), zeros elsewhere; 64 columns = one 128-byte
   * K-major operand row).  x A^T is accumulated in fp32 next to the base product, rounded to
   * bf16 and multiplied by `up` with one more tensor-core instruction per tile.
   * lora_rank_pad: rank rounded up to a multiple of 4 (<= 16; informational). */
  const void* lora_down;
  const void* lora_up;
  int32_t lora_rank_pad;
  int32_t lora_seg_n;
  int32_t flags;        /* IDB_EPI_* */
  /* outputs: row-major [M, N_out] with M = batch*Ho*Wo in (b, y, x) raster order */
  float* out_f32;       /* or NULL */
  void* out_bf16;       /* or NULL */
  /* split-K: writes raw partial sums to workspace (fp32 [k_splits, M, N]) and a second kernel
   * applies the epilogue.  k_splits: 1 = off, >1 = forced, 0 = auto (library picks a split that
   * fills the SMs, limited by workspace_bytes).  workspace may be NULL when k_splits == 1. */
  int32_t k_splits;
  float* workspace;
  size_t workspace_bytes;
  /* optional: per-32-row-block channel statistics of the fp32 output, [ceil(M/32), N_out, 2] floats =
   * (sum, sum of squares), written by the epilogue for free; idb_groupnorm consumes them instead of
   * re-reading the tensor.  Needs out_f32, Wo a power of two (or a multiple of 128) and Ho*Wo % 32 == 0. */
  float* stats_partials;
  /* optional: per-IMAGE sums of the fp32 output over granules of stats_gran consecutive channels, in 64-bit fixed point:
   * int64 [n_images, N_out / stats_gran, 2] = (sum * 2^32, sum of squares * 2^24) over the image's stats_hw rows,
   * n_images = M / stats_hw.  stats_gran (0 = 1) must divide the group size of every GroupNorm that will consume the
   * tensor and the channel offset at which it is concatenated (SD2.1 UNet: 10, VAE: 4).  The epilogue ADDS its 32-row
   * sums with integer atomics (exact and order-independent: bit-reproducible without any ordering), so the caller ZEROES the
   * tensor before the call -- before the first of the four phase calls that share it when out_scale = 2.
   * idb_groupnorm(x0_sums) then needs neither a statistics pass nor a finalize launch.  Range: |sum| < 2^31 and sum of
   * squares < 2^39 per (image, granule).  Same geometry precondition as stats_partials (which may be NULL).
   * stats_hw: rows per image (0 = Ho*Wo; a Linear over tokens passes the tokens per image); a multiple of 32. */
  int64_t* stats_image_sums;
  int32_t stats_hw;
  int32_t stats_gran;
  /* optional: per-output-channel PReLU slopes [N], applied after bias / rowvec and before the residual
   * (ArcFace IResNet: bn2 folded into conv1, then nn.PReLU(planes)).  Not combined with GEGLU. */
  const float* prelu;
  /* IDB_A_2X2 only */
  int32_t tap_off_x, tap_off_y;
  /* phased output: out_scale = 2 writes output pixel (y, x) of image b at (2y + out_phase_y, 2x + out_phase_x) of a
   * [batch, 2*Ho, 2*Wo, N] tensor (and numbers its statistics row blocks per phase: stats_partials is then
   * [4 phases][ceil(M/32)][N][2], this call filling phase 2*out_phase_y + out_phase_x).  0 / 1 = plain output.
   * Not combined with residual, split-K or both outputs. */
  int32_t out_scale, out_phase_x, out_phase_y;
} idb_gemm_conv_args;

int idb_gemm_conv(const idb_gemm_conv_args* args, void* stream);
/* Bytes of workspace idb_gemm_conv needs for (M, N, k_splits). */
size_t idb_gemm_conv_workspace_bytes(int64_t m, int64_t n, int32_t k_splits);

/* ------------------------------------------------------------------------------------------
 * idb_attention: O = softmax(Q K^T * scale) V per (batch, head), flash-style on tcgen05
 * (S and O accumulators in TMEM, online softmax in fp32).  Replaces AttnProcessor2_0 ->
 * F.scaled_dot_product_attention (diffusers Attention, self- and cross-; head_dim 64) inside the UNet call of
 * inference_ID-Booth.py:138 / train_ID-Booth.py:1040-1046 (the attention projections are the LoRA targets named at
 * train_ID-Booth.py:676); with causal = 1 the CLIPAttention behind encode_prompt (train_ID-Booth.py:476-491).
 * q/k/v: bf16 row-major token matrices; row (b, t) at ptr + ((b*T + t)*ld + col0 + head*64).
 * out: bf16 [B*Tq, heads*64] (ld_out).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* q; int64_t ld_q; int32_t col0_q;
  const void* k; int64_t ld_k; int32_t col0_k;
  const void* v; int64_t ld_v; int32_t col0_v;
  void* out;     int64_t ld_out;
  int32_t batch, heads, t_q, t_kv;
  float scale;
  int32_t causal;   /* 1: key j is visible to query i only if j <= i (CLIP text tower); supported for t_kv <= 96 */
  /* optional: log-sum-exp of the scaled scores of every query row in the LOG2 domain, fp32 [batch, heads, t_q]
   * (= log2 sum_j exp(scale * q.k_j)); saved by a training forward for idb_attention_backward */
  float* lse;
} idb_attention_args;
int idb_attention(const idb_attention_args* args, void* stream);

/* idb_attention_backward: dQ, dK, dV of O = softmax(Q K^T * scale) V given dO (LoRA-only training backward,
 * train_ID-Booth.py:1140; the attention projections carry the trainable adapters, :672-678).  Same operand addressing as
 * idb_attention; o = the forward output, lse = the forward's log2-domain log-sum-exp.
 *   dq   : fp32 [batch * t_q, ld_dq], head h at columns col0_dq + 64 h; MUST BE ZERO on entry (every key tile adds its
 *          contribution atomically; fp32 atomics: the last bits depend on the arrival order)
 *   dk/dv: bf16 [batch * t_kv, ld], head h at col0 + 64 h (written, not accumulated)
 *   dsum : fp32 scratch [batch, heads, t_q] (rowsum(dO o O), filled by the call) */
typedef struct {
  const void* q; int64_t ld_q; int32_t col0_q;
  const void* k; int64_t ld_k; int32_t col0_k;
  const void* v; int64_t ld_v; int32_t col0_v;
  const void* o; int64_t ld_o; int32_t col0_o;
  const void* d_o; int64_t ld_do; int32_t col0_do;
  const float* lse;
  float* dsum;
  float* dq; int64_t ld_dq; int32_t col0_dq;
  void* dk; int64_t ld_dk; int32_t col0_dk;
  void* dv; int64_t ld_dv; int32_t col0_dv;
  int32_t batch, heads, t_q, t_kv;
  float scale;
} idb_attention_bwd_args;
int idb_attention_backward(const idb_attention_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) over NHWC, fp32 statistics.  Replaces F.group_norm + F.silu in
 * ResnetBlock2D.norm1/norm2, Transformer2DModel.norm, conv_norm_out (UNet call of inference_ID-Booth.py:138 /
 * train_ID-Booth.py:1040-1046; VAE decode of train_ID-Booth.py:412,437).  Reads the logical
 * channel-concatenation [x0 | x1] (UpBlock skip `torch.cat([h, skip], 1)`) without
 * materialising it.  Inputs fp32 (stream) ; outputs bf16 [B,H,W,C0+C1]:
 *   out_norm = act(GN(x))        out_raw (optional) = bf16(x)   (operand of conv_shortcut)
 * partials: fp32 workspace of idb_groupnorm_workspace_bytes(); it must be ZERO-INITIALISED once by the caller
 * (it holds arrival counters that every call leaves at zero again); calls sharing a workspace must be
 * stream-ordered.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const float* x0; int32_t c0;
  const float* x1; int32_t c1;      /* NULL / 0 when there is no concat */
  int32_t batch, hw, groups;
  float eps;
  const float* gamma; const float* beta;   /* [C0+C1] */
  int32_t silu;
  void* out_norm;                    /* bf16 */
  void* out_raw;                     /* bf16 or NULL */
  float* partials;
  /* optional: row-block statistics produced by idb_gemm_conv(stats_partials) for x0 / x1
   * ([hw/32 * batch, C, 2] each).  When given for every source the statistics pass over x is skipped. */
  const float* x0_stats;
  const float* x1_stats;
  /* 4 when x0 was written by four phased idb_gemm_conv calls (out_scale = 2): x0_stats is then
   * [4 phases][batch * hw/128][C0][2]; 0 / 1 otherwise */
  int32_t x0_stats_phases;
  /* optional, preferred: per-image granule sums accumulated by idb_gemm_conv(stats_image_sums) for x0 / x1 (int64 fixed
   * point [batch, C / sums_gran, 2] each; sums_gran divides the group size and c0).  When given for every source the call
   * is ONE launch: each CTA derives the group statistics of its image from these few hundred bytes itself. */
  const int64_t* x0_sums;
  const int64_t* x1_sums;
  int32_t sums_gran;
  /* 0 (= batch), or the number of images x1 (and x1_sums) holds when that is fewer than batch: image b then reads image
   * b % x1_batch of x1 -- a skip tensor computed once for both halves of a classifier-free-guidance pair. */
  int32_t x1_batch;
} idb_groupnorm_args;
int idb_groupnorm(const idb_groupnorm_args* args, void* stream);
size_t idb_groupnorm_workspace_bytes(int32_t batch, int32_t groups);

/* LayerNorm over the last dim (eps 1e-5, affine): fp32 [rows, C] -> bf16 [rows, C].
 * Replaces BasicTransformerBlock.norm1/2/3 (UNet call of inference_ID-Booth.py:138 / train_ID-Booth.py:1040-1046) and the
 * LayerNorms of the CLIP text tower (encode_prompt, train_ID-Booth.py:476-491). C % 4 == 0, C <= 2048. */
int idb_layernorm(const float* x, const float* gamma, const float* beta, void* out_bf16,
                  int64_t rows, int32_t c, float eps, void* stream);

/* Row softmax for the VAE mid-block attention (1 head, d = 512; vae.decode at train_ID-Booth.py:412,437 and the pipeline
 * tail of inference_ID-Booth.py:138): fp32 [rows, cols] * scale ->
 * bf16 probabilities [rows, cols]. */
int idb_softmax_rows(const float* s, void* p_bf16, int64_t rows, int32_t cols, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small / bandwidth-bound pieces
 * ---------------------------------------------------------------------------------------- */
/* Timesteps(320, flip_sin_to_cos, shift 0) -> linear_1 -> SiLU -> linear_2 -> SiLU (the SiLU
 * that every ResnetBlock2D applies before time_emb_proj), then ALL time_emb_proj layers at
 * once: proj_out[b, :] = W_all . silu(emb[b]) + b_all  with W_all = row-concat of the 22
 * time_emb_proj weights.  fp32 weights.  scratch: fp32 [batch, 2*dim_emb + dim_sin].
 * Reference: the `timesteps` argument of the UNet call, train_ID-Booth.py:1040-1046 / the pipeline loop of
 * inference_ID-Booth.py:138. */
typedef struct {
  const float* timesteps;            /* [batch] */
  int32_t batch, dim_sin, dim_emb;   /* 320, 1280 */
  const float* w1; const float* b1;  /* [dim_emb, dim_sin] */
  const float* w2; const float* b2;  /* [dim_emb, dim_emb] */
  const float* w_all; const float* b_all; int32_t n_all;   /* [n_all, dim_emb] */
  float* proj_out;                   /* [batch, n_all] */
  float* scratch;
} idb_time_embed_args;
int idb_time_embed(const idb_time_embed_args* args, void* stream);

/* 3x3 pad-1 conv with tiny Cin (<= 8): conv_in of the UNet (4->320) and of the VAE decoder
 * (4->512).  x fp32, NCHW when x_nchw else NHWC; w fp32 [Cout, 3, 3, Cin]; out fp32 NHWC
 * and/or bf16 NHWC. */
int idb_conv3x3_small_cin(const float* x, int32_t x_nchw, const float* w, const float* bias,
                          float* out_f32, void* out_bf16,
                          int32_t batch, int32_t h, int32_t wd, int32_t cin, int32_t cout, void* stream);
/* 3x3 pad-1 conv with tiny Cout (<= 4): conv_out of the UNet (320->4) and VAE (128->3).
 * x bf16 NHWC (already GroupNorm+SiLU'd); w fp32 [Cout, 3, 3, Cin]; out fp32 NCHW, or when
 * postprocess != 0 NHWC with (v*0.5+0.5).clamp(0,1) (VaeImageProcessor.postprocess "np"; in-tree twin
 * train_ID-Booth.py:413-415, output_type="np" at inference_ID-Booth.py:138). */
int idb_conv3x3_small_cout(const void* x_bf16, const float* w, const float* bias, float* out,
                           int32_t postprocess, int32_t batch, int32_t h, int32_t wd,
                           int32_t cin, int32_t cout, void* stream);
/* nearest-neighbour 2x upsample, NHWC: fp32 in -> bf16 out (Upsample2D's F.interpolate). */
int idb_upsample2x(const float* x, void* out_bf16, int32_t batch, int32_t h, int32_t wd, int32_t c, void* stream);
/* fp32 -> bf16 cast (operand staging for Downsample2D). */
int idb_cast_bf16(const float* x, void* out_bf16, int64_t n, void* stream);
/* VAE front: z/scaling_factor -> post_quant_conv 1x1 (4->4): NCHW fp32 -> NHWC fp32
 * (`latents = (1 / 0.18215) * latents; vae.decode(latents)`, train_ID-Booth.py:410-412,435-437). */
int idb_vae_latent_prep(const float* z_nchw, const float* w, const float* bias, float inv_scaling,
                        float* out_nhwc, int32_t batch, int32_t hw, void* stream);

/* ------------------------------------------------------------------------------------------
 * idb_cfg_ddpm_step: classifier-free-guidance combine + DDPMScheduler.step in ONE launch.
 * Replaces `eps_u + s*(eps_c - eps_u)` and diffusers DDPMScheduler.step (pipeline loop behind
 * inference_ID-Booth.py:138; guidance_scale from :49).  All tensors fp32 NCHW [n, ...]:
 *   eps = cfg ? eps2[0:n] + s*(eps2[n:2n] - eps2[0:n]) : eps2[0:n]
 *   x0  = epsilon: (x - sqrt_1m_acp*eps)/sqrt_acp ; v-pred: sqrt_acp*x - sqrt_1m_acp*eps
 *   x_prev = c_x0*x0 + c_xt*x + sigma*noise           (noise may be NULL when sigma == 0)
 * coef: DEVICE fp32[5] = {sqrt_acp, sqrt_1m_acp, c_x0, c_xt, sigma} (a row of the per-step
 * table the scheduler uploads at set_timesteps -- no host sync inside the loop).
 * ---------------------------------------------------------------------------------------- */
/* UNet conv_in operand: fp32 NCHW latents [B,4,H,W] -> bf16 NHWC [B,H,W,64] = [hi(x) | lo(x) | hi(x) | 0...]; with conv_in
 * weights packed [w_hi | w_hi | w_lo] per tap, idb_gemm_conv computes the 4-channel 3x3 conv at fp32-product precision. */
int idb_latent_operand(const float* x_nchw, void* out_bf16_nhwc64, int32_t batch, int32_t hw, void* stream);

int idb_cfg_ddpm_step(const float* eps2, const float* x, const float* noise, const float* coef,
                      float guidance_scale, int32_t use_cfg, int32_t v_prediction,
                      float* x_prev, float* x0_out /* or NULL */, int64_t n_per_branch, void* stream);

/* ------------------------------------------------------------------------------------------
 * ArcFace IResNet-100 glue (reference: ArcFace_files/backbones/iresnet.py:29-162, train_ID-Booth.py:433-455).
 * idb_channel_affine: out[b, yo, xo, c] = bf16(x[b, s*yo, s*xo, c] * scale[c] + shift[c]) -- an eval-mode
 *   BatchNorm2d in front of a conv (IBasicBlock.bn1, IResNet.bn2) and / or the stride-s sampling of the 1x1
 *   stride-s `downsample` conv; scale / shift may be NULL.  x fp32 NHWC [B,H,W,C], out bf16 (fp16 when out_f16)
 *   [B,H/s,W/s,C].
 * idb_crop_resize_norm: crop bbox (x0,y0,x1,y1 per image, int32, clamped to the image) from fp32 NHWC [n,H,W,3]
 *   images in [0,1], bilinear resize (align_corners = False, no antialias) to size x size, (v - 0.5) / 0.5, written
 *   as bf16 NHWC [n, size, size, c_pad] (channels >= 3 zero): the IResNet stem operand.
 * ---------------------------------------------------------------------------------------- */
int idb_channel_affine(const float* x, const float* scale, const float* shift, void* out_16, int32_t out_f16, int32_t batch,
                       int32_t h, int32_t w, int32_t c, int32_t stride, void* stream);
int idb_crop_resize_norm(const float* img_nhwc, const int32_t* bbox_xyxy, void* out_16, int32_t out_f16, int32_t n, int32_t h,
                         int32_t w, int32_t size, int32_t c_pad, void* stream);

/* ------------------------------------------------------------------------------------------
 * LoRA-only training backward (train_ID-Booth.py:1140-1146 with the adapters of :672-678 as the only
 * trainable tensors).  The contractions of the backward run on idb_gemm_conv (transposed packed weights; the fused-LoRA
 * form covers dX = dY W + (dY B) A with the adapter roles swapped) and idb_attention_backward; these are the
 * bandwidth-bound pieces.  All reductions run in a fixed order.
 * ---------------------------------------------------------------------------------------- */
/* dx (+)= LayerNorm input gradient: dy, x fp32 [rows, C]; C % 4 == 0, C <= 2048 */
int idb_layernorm_backward(const float* dy, const float* x, const float* gamma, float* dx, int32_t add, int64_t rows, int32_t c,
                           float eps, void* stream);
/* GroupNorm(+SiLU) input gradient over the logical concatenation [x0 | x1] (as idb_groupnorm reads it).  dy fp32
 * [batch, hw, C0+C1] = gradient with respect to act(GN(x)); stats fp32 [batch, groups, 2] = (mean, rstd) of the forward;
 * scratch fp32 [batch, groups, 16, 2]; dx0 / dx1 fp32 like x0 / x1 (either may be NULL), accumulated into when add0 / add1. */
typedef struct {
  const float* dy;
  const float* x0; int32_t c0;
  const float* x1; int32_t c1;
  int32_t batch, hw, groups, silu;
  const float* stats;
  const float* gamma; const float* beta;
  float* scratch;
  float* dx0; float* dx1;
  int32_t add0, add1;
} idb_groupnorm_bwd_args;
int idb_groupnorm_backward(const idb_groupnorm_bwd_args* args, void* stream);
/* GEGLU backward: dh bf16 [M, H]; u bf16 [M, 2H] = the pre-activation in the interleaved [a(16) | g(16)] layout of
 * IDB_EPI_GEGLU (recomputed by the caller); du bf16 [M, 2H] in the same layout */
int idb_geglu_backward(const void* dh_bf16, const void* u_bf16, void* du_bf16, int64_t m, int32_t h, void* stream);
/* Adapter weight gradient: out[w, r] (+)= scale * sum_m wide[m, col0_w + w] * skinny[m, col0_s + r], r < rank <= 16
 * (dB = dY^T (x A^T); dA = ((dY B)^T x) with transpose_out = 1: out[r, w]).  wide / skinny bf16 with row strides ld_w /
 * ld_s; out fp32 with row stride out_ld; workspace: idb_lora_wgrad_workspace_bytes(w) bytes. */
size_t idb_lora_wgrad_workspace_bytes(int32_t w);
int idb_lora_wgrad(const void* wide_bf16, int64_t ld_w, int32_t col0_w, const void* skinny_bf16, int64_t ld_s, int32_t col0_s,
                   float* out, int32_t out_ld, int32_t transpose_out, float scale, int32_t add, int64_t m, int32_t w, int32_t r,
                   float* workspace, void* stream);
/* z[b, 2y, 2x, :] = g[b, y, x, :], zero elsewhere (bf16 NHWC): operand of the input gradient of a stride-2 conv */
int idb_zero_insert2x(const void* g_bf16, void* z_bf16, int32_t batch, int32_t h, int32_t w, int32_t c, void* stream);
/* out[b, y, x, :] (+)= sum of the 2x2 block of g (fp32 NHWC [batch, 2h, 2w, c]): input gradient of nearest-2x upsampling */
int idb_sumpool2x(const float* g, float* out, int32_t add, int32_t batch, int32_t h, int32_t w, int32_t c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDB_H_ */
