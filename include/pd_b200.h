/*
 * pd_b200.h — C-ABI of the B200-native Prompt-Diffusion denoising hot path.
 *
 * The reference (david3684/Prompt-Diffusion) is 100 % Python/PyTorch and has no
 * FFI of its own; its "operator API" on this path is the set of torch ops that
 * cldm/cldm.py, ldm/modules/diffusionmodules/openaimodel.py, ldm/modules/attention.py
 * and cldm/ddim_hacked.py call.  Each entry point below replaces one such op (or
 * a fused group of them) and cites the reference lines it stands in for.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch / C++ types.
 *  - every pointer is a DEVICE pointer unless the name ends in _host.
 *  - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it
 *    and never synchronise, allocate or free (CUDA-graph capturable).
 *  - activations are pixel-major ("NHWC"): element (b,y,x,c) of a [B,H,W,C]
 *    tensor lives at base + ((b*H + y)*W + x)*ld + c, with row pitch ld >= C in
 *    ELEMENTS (lets a producer write straight into a channel-concat slot).
 *  - dtype codes: PD_F32 (fp32 mode: SIMT FFMA kernels, fp32 storage) and
 *    PD_BF16 (bf16 storage, fp32 accumulation; tcgen05/TMEM/TMA GEMMs).
 *  - concurrency: launches that use library-owned scratch — the stream-K workspace of pd_conv2d (one per device)
 *    and, for callers that do not pass their own, the cooperative GroupNorm scratch — must not overlap on two
 *    streams of one device; serialise such users with an event, or give each its own `partial` scratch
 *    (pd_group_norm) and run pd_conv2d from one stream per device.  Host-side caches are mutex-guarded.
 *  - return value: 0 on success, a negative PD_ERR_* / positive cudaError_t
 *    otherwise; pd_last_error() gives a human-readable message.  There is no
 *    CPU fallback anywhere behind this interface.
 */
#ifndef PD_B200_H_
#define PD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_F32 0
#define PD_BF16 1

#define PD_ACT_NONE 0
#define PD_ACT_SILU 1
/* FeedForward's GEGLU (attention.py:54-56) fused into the linear that feeds it (tcgen05 engine, bf16 only):
 * w has Cout = 2F rows interleaved in blocks of 32 (32 value rows, then their 32 gate rows), bias likewise;
 * out has F = Cout/2 columns: out[:, j] = (x W_v^T + b_v)[:, j] * gelu_erf((x W_g^T + b_g)[:, j]) */
#define PD_ACT_GEGLU 2

#define PD_ERR_BAD_ARG (-1)
#define PD_ERR_UNSUPPORTED (-2)
#define PD_ERR_NO_DEVICE (-3)

/* engine selection for pd_conv2d */
#define PD_ENGINE_AUTO 0   /* tcgen05 when dtype==BF16 and the shape qualifies, else SIMT */
#define PD_ENGINE_SIMT 1   /* force the SIMT implicit-GEMM kernel (any shape, any dtype)    */
#define PD_ENGINE_TC 2     /* force tcgen05; PD_ERR_UNSUPPORTED if the shape does not fit   */

const char* pd_last_error(void);
int pd_abi_version(void);
/* number of kernels this library has launched since load (bench "gpu_launches") */
uint64_t pd_launch_count(void);
/* 1 when the current device is sm_100 (tcgen05/TMEM present) */
int pd_device_is_sm100(void);

/* Measurement hook (bench.py roofline leg): per-launch CUDA-event timing of the tcgen05 engine.
 * pd_prof_enable(1) clears the log and starts recording, pd_prof_read() synchronises and returns the
 * summed kernel milliseconds, the summed algorithmic FLOPs (2*M*Cout*K per launch) and the launch count. */
int pd_prof_enable(int on);
int pd_prof_read(double* total_ms, double* total_flops, uint64_t* launches);
/* one CSV line per recorded launch (shape, tile choice, milliseconds, TFLOP/s) */
int pd_prof_dump(const char* path_host);
#ifdef PD_DEBUG
/* PD_DEBUG builds only (scripts/build_variant.sh; never the shipped library): device buffer of 3*64*2 uint64 that
 * receives CTA 0's per-role (TMA / MMA / epilogue) tile start/end globaltimer stamps; NULL switches it off */
int pd_debug_timeline(void* dev_buf);
/* PD_DEBUG builds only: timing experiments on the conv engine (results are WRONG when non-zero):
 * 1 = issue no MMAs, 2 = issue no TMA loads, 3 = no TMA stores, 8 = no epilogue, 9 = 1 + 2 */
int pd_debug_gemm_mode(int32_t mode);
#endif
/* tile order of the tcgen05 conv engine: 1 (default) = the N tiles of one M tile are adjacent in the schedule,
 * 0 = all M tiles of one N tile first */
int pd_debug_tile_order(int32_t n_fast);
/* Launch variants (single-CTA / CTA-pair tiles, stream-K, tile width) are chosen per layer shape from a committed
 * table (csrc/tune_table.inc), so results never depend on timing noise; PD_B200_AUTOTUNE=1 times the candidates of
 * shapes missing from the table on first use (table regeneration only).  pd_tune_dump writes the table rows
 * currently in effect (committed + tuned) in tune_table.inc syntax. */
int pd_tune_dump(const char* path_host);
/* tile-shape override of the tcgen05 conv engine: 0 auto, 1 single-CTA 128-row tiles, 2 CTA-pair 256-row tiles */
int pd_debug_force_cta_group(int32_t cg);
/* experiments: pin the N extent of the tile (multiple of 32 up to 256; 0 = heuristic) */
int pd_debug_force_bn(int32_t bn);
/* resident-B schedule of the tcgen05 conv engine (short-K layers: a worker loads its weight tile, all of K, into shared
 * memory ONCE, keeps one N tile for life and streams only activations): 1 = wherever it fits, 0 = per the variant table.
 * pd_debug_bres_launches counts the launches that took it. */
int pd_debug_force_bres(int32_t on);
uint64_t pd_debug_bres_launches(void);
/* vertical-halo schedule of the tcgen05 conv engine (3x3, stride 1, one K segment, output W % 8 == 0 and H % 16 == 0): per
 * 64-channel chunk the activation operand is fetched as three column-shifted 8 x 18-pixel boxes whose three row offsets are
 * the dy taps — 3 instead of 9 activation loads per chunk.  1 = wherever it applies, 0 = per the variant table;
 * pd_debug_vh_launches counts the launches that took it. */
int pd_debug_force_vh(int32_t on);
uint64_t pd_debug_vh_launches(void);
/* together with a forced CTA group: 1 = stream-K schedule of the tcgen05 conv engine wherever it applies (the (tile, k-block)
 * space is cut evenly over the CTAs; partial tiles are summed in K order by the last CTA to arrive), 0 = data-parallel */
int pd_debug_force_stream_k(int32_t on);
/* 1 (default): GroupNorm is one cooperative launch; 0: statistics kernel + apply kernel */
int pd_debug_group_norm_fused(int32_t on);
/* A/B switch: 0 = pd_attention's auto selection never picks the four-group kernel (engine 5), 1 = default */
int pd_debug_attention_tc4(int32_t on);
/* same for the persistent kernel (engine 8) */
int pd_debug_attention_persistent(int32_t on);
/* same for the three-group / 128-key-tile kernel (engine 6, head dim <= 40) */
int pd_debug_attention_tc3(int32_t on);
/* same for the tcgen05 attention kernel: 6 phases x 32 tiles of uint64 stamps from block (0,0,0) */
int pd_debug_attention_timeline(void* dev_buf);

/*
 * Convolution / linear as one implicit GEMM.
 *   replaces  nn.Conv2d 3x3 / 1x1 (conv_nd, util.py:221-231) as used by ResBlock
 *             (openaimodel.py:200-240,254-274), Downsample (:150-159), Upsample's
 *             conv (:106,116-117), SpatialTransformer.proj_in/proj_out
 *             (attention.py:298-316,329,339), ControlNet hint stacks / zero convs
 *             (cldm.py:147-181,299-300,320-323), UNet `out` conv (openaimodel.py:726-730);
 *             and nn.Linear (time_embed, emb_layers, to_q/k/v, to_out, GEGLU proj,
 *             ff.net.2: attention.py:52,72,154-161) with H=W=1.
 *   out[b,yo,xo,n] = act( alpha * ( sum_k A[..] * Wt[n,k] + bias[n] ) + rowvec[b,n] + res[b,yo,xo,n] )
 *   K runs over segment 0 (ksize x ksize taps of x, tap-major then channel) followed by
 *   optional segment 1 (a 1x1 tap of x2 sampled at the OUTPUT pixel): this fuses
 *   ResBlock's skip_connection 1x1 conv into out_layers' conv3x3 (openaimodel.py:274).
 *   Weights are [Cout, ksize*ksize*C + C2] row-major (K contiguous), same dtype as x.
 */
typedef struct pd_conv_params {
  const void* x;       /* [B,H,W,C] pitch ldx                               */
  const void* x2;      /* optional [B,Ho,Wo,C2] pitch ldx2 (segment 1)      */
  const void* w;       /* [Cout, Ktot]                                      */
  const float* bias;   /* [Cout] fp32 or NULL                               */
  const float* rowvec; /* optional [B, Cout] fp32, pitch ldrv (timestep emb) */
  const void* res;     /* optional residual [B,Ho,Wo,Cout] pitch ldr, dtype = out dtype */
  void* out;           /* [B,Ho,Wo,Cout] pitch ldo                          */
  int32_t B, H, W, C;  /* input geometry of x (H, W BEFORE the optional upsample) */
  int32_t C2;          /* channels of x2 (0 = no segment 1)                 */
  int32_t Cout;
  int32_t ksize;       /* 1 or 3 (padding = ksize/2); 2 with pad_y / pad_x (tcgen05 engine) */
  int32_t stride;      /* 1 or 2                                            */
  int32_t upsample;    /* 1: x is read through a nearest x2 upsample (Upsample.forward) */
  int32_t ldx, ldx2, ldr, ldo, ldrv;
  int32_t act;         /* PD_ACT_*                                          */
  int32_t dtype;       /* dtype of x, x2, w                                 */
  int32_t out_dtype;   /* dtype of out and res                              */
  int32_t engine;      /* PD_ENGINE_*                                       */
  float alpha;
  int32_t w_blocked;   /* 0: w is [Cout][K] (K contiguous).  1: k-block-major [K/64][Cout][64] (tcgen05 engine only,
                        * K % 64 == 0): the B tile of a k-block is one contiguous run in HBM (measured: no gain) */
  /* LayerNorm folded into a linear layer (tcgen05 engine, bf16, ksize 1, act NONE or GEGLU, no res / rowvec, alpha 1):
   * x is the UN-normalised tensor, w holds W[n,k] * gamma[k], bias holds bias[n] + sum_k W[n,k] * beta[k], and
   *   out[m,n] = rstd[m] * (acc[m,n] - mean[m] * ln_colsum[n]) + bias[n]      (then GEGLU if requested)
   * which equals Linear(LayerNorm(x)) (attention.py:271-275: attn1 / attn2.to_q / ff of BasicTransformerBlock)
   * without ever writing the normalised tensor.  NULL = off. */
  const float* ln_stats;   /* [rows][2] fp32 (mean, rstd) from pd_layer_norm_stats */
  const float* ln_colsum;  /* [Cout] fp32: sum_k of the (rounded) scaled weights of row n */
  /* ---- statistics handed from a GEMM's epilogue to the norm that follows it (tcgen05 engine, bf16 output; all
   * zero = off).  A norm's statistics pass re-reads a tensor the producing GEMM held in registers a moment earlier;
   * the epilogue emits (sum, sum of squares) partials of the values AS STORED (bf16-rounded) instead. ---- */
  const float* ln_parts;   /* consumer side of a folded LayerNorm, alternative to ln_stats: the producer's ln_parts_out */
  float* ln_parts_out;     /* producer side: 16-byte header {int32 parts} followed by float2 (sum, sumsq) [parts][ln_rows];
                            * one partial per (N tile, epilogue warp group) of this launch and output row; needs
                            * space for pd_conv2d_ln_parts_floats(rows) floats */
  float* gn_stats_out;     /* producer side of GroupNorm (util.py:217-219 statistics): float2 (sum, sumsq) per
                            * (64-pixel record, channel): [(image * gn_recs_per_image + gn_rec_off + record)][gn_ld],
                            * this launch's channel 0 at the pointer.  Needs Ho*Wo % 64 == 0 and a pixel-box tile of
                            * at least 64 pixels (pd_conv2d_gn_stats_supported); consumed by pd_group_norm_apply */
  int64_t ln_rows;         /* row pitch of the ln_parts / ln_parts_out arrays (>= output rows) */
  int64_t out_sx, out_sy, out_sb; /* output pixel strides in ELEMENTS (x, y, image); all 0 = dense (ldo, Wo*ldo, Ho*Wo*ldo).
                            * Non-dense: ldo is ignored, res must be NULL.  Lets a launch write one phase of a 2x
                            * upsampled tensor (see pad_y / pad_x) */
  float ln_eps;            /* eps of the LayerNorm behind ln_parts */
  int32_t gn_ld, gn_rec_off, gn_recs_per_image;
  int32_t pad_y, pad_x;    /* ksize == 2 only: zero rows / columns before the first tap (0 or 1); Ho = H, Wo = W.
                            * Upsample.forward (openaimodel.py:108-118: nearest x2 then conv3x3) == four such 2x2
                            * convolutions of the LOW-resolution tensor, one per output phase (py, px), with the 3x3
                            * taps that fall on the same source pixel pre-summed: 4/9 of the MACs, no 4x tensor */
} pd_conv_params;
/* floats needed behind ln_parts_out for `rows` output rows (header + 64 partial slots) */
int64_t pd_conv2d_ln_parts_floats(int64_t rows);
/* 1 when a launch with this output geometry can emit gn_stats_out */
int pd_conv2d_gn_stats_supported(int32_t B, int32_t Ho, int32_t Wo, int32_t ksize, int32_t stride);
int pd_conv2d(const pd_conv_params* p, void* stream);

/* Re-lay a reference OIHW fp32 conv weight (or [out,in] linear weight with kh=kw=1)
 * into the [Cout, kh*kw*Cin] tap-major K-contiguous layout pd_conv2d expects,
 * channel-padding Cin -> cin_pad with zeros, converting to `dtype`.
 * (checkpoint layout: SURVEY.md appendix B; loader cldm/model.py:12-21) */
int pd_repack_conv_weight(const float* w_oihw, void* w_out, int32_t cout, int32_t cin,
                          int32_t kh, int32_t kw, int32_t cin_pad, int32_t ldk, int32_t k_offset,
                          int32_t dtype, void* stream);

/* GroupNorm(32 groups) + optional SiLU, fp32 statistics.
 *   replaces GroupNorm32 (util.py:217-219) + nn.SiLU in ResBlock.in_layers/out_layers
 *   (openaimodel.py:200-231), UNet `out` (:726-728) and Normalize (attention.py:88-89, eps 1e-6).
 *   `partial` is caller-provided scratch of pd_group_norm_scratch_floats(B) floats that must be ZERO-filled
 *   once when allocated (it holds self-resetting arrival counters) and must not be shared by concurrent streams. */
int64_t pd_group_norm_scratch_floats(int32_t B);
int pd_group_norm(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma,
                  const float* beta, float* partial, int32_t B, int32_t HW, int32_t C,
                  int32_t groups, float eps, int32_t act, int32_t dtype, int32_t out_dtype,
                  void* stream);

/* GroupNorm(+SiLU) of a bf16 tensor whose (sum, sum of squares) records were emitted by the epilogue of the GEMM(s)
 * that produced it (pd_conv_params.gn_stats_out: float2 [B * recs_per_image][stats_ld], one record per 64 pixels):
 * the records are folded per (image, group) in a fixed order, then x is streamed once.  Same arithmetic as
 * pd_group_norm (fp32 statistics over the stored bf16 values, GroupNorm32 util.py:217-219).  `scratch` receives the
 * B * 64 (mean, rstd) floats.  split != 0: out gets 2C columns, bf16(y) at column c and bf16(y - bf16(y)) at C + c. */
int pd_group_norm_apply(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma,
                        const float* beta, const float* colstats, int32_t stats_ld,
                        int32_t recs_per_image, float* scratch, int32_t B, int32_t HW, int32_t C,
                        int32_t groups, float eps, int32_t act, int32_t split, void* stream);

/* Row statistics of LayerNorm: stats[row] = (mean, 1/sqrt(var + eps)) in fp32, for the folded form of
 * LayerNorm -> Linear (pd_conv_params.ln_stats). */
int pd_layer_norm_stats(const void* x, int32_t ldx, float* stats, int64_t rows, int32_t C, float eps,
                        int32_t dtype, void* stream);

/* LayerNorm over the last dim (nn.LayerNorm, attention.py:263-265, eps 1e-5). */
int pd_layer_norm(const void* x, int32_t ldx, void* out, int32_t ldo, const float* gamma,
                  const float* beta, int64_t rows, int32_t C, float eps, int32_t dtype,
                  void* stream);

/* GEGLU gate: out[m,f] = x[m,f] * gelu_erf(x[m,F+f])  (attention.py:54-56). */
int pd_geglu(const void* x, int32_t ldx, void* out, int32_t ldo, int64_t rows, int32_t F,
             int32_t dtype, void* stream);

/* Row softmax of a materialised score matrix: out[r,:] = softmax(x[r,:] * scale) (fp32 statistics).
 *   The first-stage decoder's single-head, 512-channel attention (AttnBlock.forward,
 *   ldm/modules/diffusionmodules/model.py:190-193: bmm -> * c^-0.5 -> softmax(dim=2)) runs as
 *   pd_conv2d (q k^T) -> pd_softmax_rows -> pd_conv2d (p v).  cols <= 16384 (bf16) / 8192 (fp32). */
int pd_softmax_rows(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int32_t cols,
                    float scale, int32_t dtype, void* stream);

/* Multi-head softmax attention without materialising the score matrix.
 *   replaces CrossAttention.forward's einsum/softmax/einsum (attention.py:171-193):
 *   q [B,Nq,heads*d] pitch ldq, k/v [B,Nk,heads*d] pitch ldk/ldv, out [B,Nq,heads*d];
 *   sim = q k^T * scale in fp32, softmax over Nk, out = sim v.  No mask (the path never
 *   passes one). */
int pd_attention(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v,
                 int32_t ldv, void* out, int32_t ldo, int32_t B, int32_t heads, int32_t Nq,
                 int32_t Nk, int32_t d, float scale, int32_t dtype, void* stream);
/* same, with an explicit engine: 0 auto, 1 SIMT (fp32 math, any dtype), 2 warp-level mma.sync (bf16),
 * 3 tcgen05/TMEM/TMA (bf16, head dim <= 192; one query group per CTA above 128: the d = 160 levels),
 * 4 single-pass mma.sync short-key kernel (bf16, Nk <= 128, head dim <= 80; auto keeps it for small query counts and the
 * causal CLIP tower), 5 four-query-group / 64-key-tile tcgen05 kernel (bf16, head dim <= 64; measured equal to engine 3 on
 * B200, kept as an explicit engine / PD_B200_ATTN4=1 only), 6 three-query-group / 128-key-tile tcgen05 kernel (bf16, head
 * dim <= 40; what auto picks from 6144 queries up), 7 persistent tcgen05 short-key kernel (bf16, Nk <= 128, head dim <= 128:
 * the 77-token cross-attention; what auto picks for it once there are >= 296 256-query units), 8 persistent form of engine 3
 * (bf16, head dim <= 64: one CTA per SM walks the (batch, head, 256-query) units, the next unit's Q / K loads and first
 * Q K^T run under the current unit's tail; measured equal to engine 3 (the unit boundary is not where that kernel loses time):
 * explicit engine, auto only with PD_B200_ATTN_PERSIST=1) */
int pd_attention_ex(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v,
                    int32_t ldv, void* out, int32_t ldo, int32_t B, int32_t heads, int32_t Nq,
                    int32_t Nk, int32_t d, float scale, int32_t dtype, int32_t engine, void* stream);

/* Causal self-attention (query i attends keys <= i), Nq == Nk == N: the CLIP text tower behind FrozenCLIPEmbedder
 * (ldm/modules/encoders/modules.py:117-128 -> transformers CLIPTextModel, 77 tokens, 12 heads, d 64).  bf16: the
 * single-pass short-key kernel (N <= 128, d <= 80); otherwise / fp32: the SIMT engine. */
int pd_attention_causal(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v,
                        int32_t ldv, void* out, int32_t ldo, int32_t B, int32_t heads, int32_t N,
                        int32_t d, float scale, int32_t dtype, void* stream);

/* CLIP text embeddings: out[r,:] = token_embedding[ids[r],:] + position_embedding[r % L,:] (fp32 tables, ids int64,
 * out in `dtype`, pitch ldo); ids outside [0, vocab) are clamped. */
int pd_embedding_lookup(const int64_t* ids, const float* tok, const float* pos, void* out, int32_t ldo,
                        int64_t rows, int32_t L, int32_t C, int32_t vocab, int32_t dtype, void* stream);

/* quick-GELU of the CLIP MLP: out = x * sigmoid(1.702 x), row-pitched 2-D tensors (in place allowed). */
int pd_quick_gelu(const void* x, int32_t ldx, void* out, int32_t ldo, int64_t rows, int32_t C,
                  int32_t dtype, void* stream);

/* timestep_embedding (util.py:154-174): t [B] int64 -> [B, dim] = [cos(t f), sin(t f)].
 * `freqs` (optional, dim/2 device floats) is the frequency table exp(-ln(max_period) j / half) as the
 * caller's host library rounds it (the reference builds it with torch.exp on the host, util.py:165-167);
 * NULL = computed in-kernel (may differ from a given host exp() by 1 ulp, i.e. ~5e-5 in cos(981 f)). */
int pd_timestep_embedding(const int64_t* t, const float* freqs, void* out, int32_t ldo, int32_t B,
                          int32_t dim, float max_period, int32_t out_dtype, void* stream);

/* elementwise SiLU (nn.SiLU between hint-stack convs is fused via pd_conv_params.act;
 * this one serves emb_layers' leading SiLU, openaimodel.py:217-218). */
int pd_silu(const void* x, void* out, int64_t n, int32_t dtype, void* stream);

/* Layout / dtype bridges at the NCHW-fp32 boundary of the reference API
 * (ControlLDM.apply_model tensors, cldm.py:369-382). */
/* accumulate != 0: out += x (adds an externally supplied NCHW control tensor in place,
 * ControlledUnetModel.forward `h += control.pop()`, cldm.py:35,41) */
int pd_nchw_to_nhwc(const float* x, void* out, int32_t ldo, int32_t B, int32_t C, int32_t H,
                    int32_t W, int32_t out_dtype, int32_t accumulate, void* stream);
int pd_nhwc_to_nchw(const void* x, int32_t ldx, float* out, int32_t B, int32_t C, int32_t H,
                    int32_t W, int32_t dtype, float scale, void* stream);
/* Split-precision entry of the latent (bf16 mode): out[:, 0:C] = hi = bf16(x), out[:, C:2C] = bf16(x - hi),
 * out[:, 2C:3C] = hi; with conv_in's weights packed [w_hi | w_hi | w_lo] per tap (input_blocks.0.0 of both nets,
 * openaimodel.py:560-566 / cldm.py:183-189) the 4-channel latent is convolved at ~16 mantissa bits inside the
 * 64-channel K padding the tcgen05 engine needs anyway. */
int pd_nchw_to_nhwc_split(const float* x, void* out, int32_t ldo, int32_t B, int32_t C, int32_t H,
                          int32_t W, void* stream);
/* out[r, c] = x[r, c] + y[r, c], fp32, row-pitched (sums the hi / lo weight halves of the UNet `out` conv). */
int pd_add2d(const float* x, int32_t ldx, const float* y, int32_t ldy, float* out, int32_t ldo,
             int64_t rows, int32_t cols, void* stream);
/* strided 2-D cast/copy: out[r, c] = (out_dtype) x[r, c], r < rows, c < cols */
int pd_cast2d(const void* x, int32_t ldx, int32_t dtype, void* out, int32_t ldo,
              int32_t out_dtype, int64_t rows, int32_t cols, void* stream);
/* nearest-neighbour x2 upsample of a pixel-major tensor (F.interpolate, openaimodel.py:115). */
int pd_upsample2x(const void* x, int32_t ldx, void* out, int32_t ldo, int32_t B, int32_t H,
                  int32_t W, int32_t C, int32_t dtype, void* stream);

/* Fused classifier-free guidance + DDIM update for one step
 *   replaces cldm/ddim_hacked.py:193 (e = e_u + s (e_c - e_u)) and :211-233
 *   (pred_x0, dir_xt, noise, x_prev).  `coef` points to 6 floats on the DEVICE:
 *   {a_t, a_prev, sigma_t, sqrt_one_minus_at, cfg_scale, temperature}; eps_uncond may be
 *   NULL (no guidance: e = eps_cond); noise may be NULL (treated as 0; valid when sigma_t == 0).
 *   All tensors fp32, n elements, any (matching) layout. */
int pd_cfg_ddim_step(const float* eps_uncond, const float* eps_cond, const float* x,
                     const float* noise, const float* coef, float* x_prev, float* pred_x0,
                     float* e_t, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PD_B200_H_ */
