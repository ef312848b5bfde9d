/* diffpure_b200.h -- C ABI of the B200-native diffusion-purification engine (libdiffpure_b200.so).
 *
 * The reference (NVlabs/DiffPure) has no FFI: its hot path is reached through Python classes
 * (runners/diffpure_sde.py:150-247 RevGuidedDiffusion / RevVPSDE, runners/diffpure_guided.py:17-89
 * GuidedDiffusion, runners/diffpure_ddpm.py:57-142 Diffusion, called by eval_sde_adv.py:68-93
 * SDE_Adv_Model.forward). This header is the boundary those runner classes bind instead of calling
 * torch modules: plain pointers and sizes, no torch types, integer return codes (0 = ok), no C++
 * exceptions across the ABI; the message of the last failure is returned by dp_last_error().
 *
 * Model: an engine owns device buffers and a *program* -- the score-model UNet lowered by the host
 * (diffpure_b200/lowering.py) to a sequence of fused sm_100a kernels -- for one fixed batch size.
 *   dp_op_embed      <- timestep embedding        (score_sde/models/layers.py:515-529, guided_diffusion/nn.py:111-129,
 *                                                   ddpm/unet_ddpm.py:14-32)
 *   dp_op_gemm       <- every conv3x3 / conv1x1 / NIN / Linear / attention matmul as a tcgen05 implicit GEMM
 *                       (score_sde/models/layers.py:100-124,546-555; layerspp.py:75-91,242-274;
 *                        guided_diffusion/unet.py:151-362; ddpm/unet_ddpm.py:63-197)
 *   dp_op_gn_apply   <- GroupNorm + SiLU (+ FiLM, + nearest-up / 2x2-mean-down, + channel concat)
 *                       (layerspp.py:219,231,245-258; unet.py:244-260; nn.py:25-27; unet_ddpm.py:40-41,55-60)
 *   dp_op_stats      <- GroupNorm statistics of a tensor not produced by dp_op_gemm
 *   dp_op_conv_in    <- the 3->C input conv        (ncsnpp.py:268, unet.py:486, unet_ddpm.py:229)
 *   dp_op_update     <- the per-step update behind the C->3|6 output conv (itself a dp_op_gemm padded to 8 columns)
 *                       (ncsnpp.py:371-374 + runners/diffpure_sde.py:86-147 + torchsde Euler step;
 *                        guided_diffusion/gaussian_diffusion.py:240-334,403-447; runners/diffpure_ddpm.py:37-54;
 *                        runners/diffpure_ode.py:90-131; runners/diffpure_ldsde.py:92-148)
 *   dp_op_attn_small <- whole-sequence attention for short sequences (T <= 64)
 *   dp_op_attn_block <- the whole attention block (projections, softmax, output projection, residual) for T = C = 256
 * Threading: one engine per (process, device); calls on an engine are stream-ordered and not re-entrant.
 * Ownership: the caller owns every pointer it passes to dp_unet_forward / dp_purify; the engine owns
 * buffers obtained from dp_buffer_alloc, its tensor maps and CUDA graphs.
 */
#ifndef DIFFPURE_B200_H_
#define DIFFPURE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dp_engine dp_engine;

#define DP_OK 0
#define DP_ERR_INVALID 1
#define DP_ERR_CUDA 2
#define DP_ERR_STATE 3

/* ---- lifecycle ------------------------------------------------------------------------------ */
int dp_create(dp_engine** out, int device);
void dp_destroy(dp_engine* e);
const char* dp_last_error(const dp_engine* e); /* e may be NULL: error of the last failed dp_create */
int dp_version(void);
int dp_device_sm_count(const dp_engine* e);

/* ---- engine-owned device buffers (weights, activations, tables) ------------------------------ */
int dp_buffer_alloc(dp_engine* e, size_t bytes, int* buf_id);
/* Registers caller-owned device memory (256-byte aligned, on the engine's device) as an engine buffer without copying:
 * the packed weight blob is uploaded / NCCL-broadcast once per device and shared by the engines of every batch size
 * (the reference re-replicates the whole module per forward through nn.DataParallel, eval_sde_adv.py:227-229).
 * The engine never frees it; it must outlive the engine. */
int dp_buffer_adopt(dp_engine* e, void* device_ptr, size_t bytes, int* buf_id);
void* dp_buffer_ptr(dp_engine* e, int buf_id);
int dp_buffer_write(dp_engine* e, int buf_id, size_t offset, const void* host_src, size_t bytes);
int dp_buffer_read(dp_engine* e, int buf_id, size_t offset, void* host_dst, size_t bytes);
size_t dp_bytes_allocated(const dp_engine* e);

/* ---- program construction (all pointers are device pointers inside engine buffers) ----------- */

/* Timestep embedding of the per-step / per-sample condition -> bf16 [B, dim]. */
typedef struct {
  void* out_bf16;   /* [B, dim] */
  int B, dim;
  int cos_first;    /* 0: [sin | cos] (DDPM++, ddpm)   1: [cos | sin] (ADM) */
  int half_minus_1; /* 1: freq_i = exp(-ln(1e4) i/(half-1))   0: /half (ADM) */
} dp_embed_desc;

/* One K segment of the implicit-GEMM A operand. */
typedef struct {
  const void* act_bf16; /* NHWC bf16 [B, Hin, Win, c_total]; plain GEMM: [rows, c_total] with Hin = 1 */
  int C;                /* channels read (multiple of 64) */
  int c_total;          /* channel pitch of the tensor */
  int taps;             /* 1 or 9 */
  int stride;           /* 1, or 2 (pad right/bottom, ddpm/unet_ddpm.py:63-82) */
  int pad;              /* 1 for 'same' 3x3, else 0 */
} dp_gemm_aseg;

typedef struct {
  dp_gemm_aseg a[2];
  int nseg;
  const void* w_bf16;   /* [rows, Ktotal] bf16, K contiguous: Ktotal = sum taps*C over segments */
  long long w_rows;     /* rows of the weight/B matrix visible to the map */
  long long w_pitch;    /* elements between rows */
  long long w_cols;     /* K columns visible to the map (0 = Ktotal) */
  int B, H, W;          /* output grid (plain GEMM: B = 1, H = 1, W = rows) */
  int N;                /* output columns (multiple of 8) */
  int batch;            /* batched GEMM count (attention), else 1 */
  int a_batch_rows, b_batch_rows;
  long long out_batch_stride;
  int inner;            /* heads per batch entry (two-level batch: index = entry*inner + head), 0/1 = none */
  int a_inner_k, a_inner_rows, b_inner_k, b_inner_rows; /* per-head offsets: A channel, A row, B K column, B row */
  long long out_inner_stride;
  const float* bias; int bias_along_m;
  const float* rowvec; int rowvec_ld; int rowvec_rows_per_sample; /* per-sample additive vector */
  const float* rowscale; /* out *= 1/rowscale[b*M + row] */
  const float* resid;
  float alpha; int silu;
  float* out_f32; void* out_bf16; long long ldc;
  float* stats;          /* [ceil(M/seg)][N][2] partial (sum, sumsq), seg = min(H*W,128); or NULL */
  int softmax; float softmax_scale; float* rowsum_out;
  /* Fused GroupNorm(+SiLU) of the result: gn_out_bf16 = act(GN(v)) [same layout as out_bf16], v = the epilogue's result.
   *  - without out_f32 (score_sde/models/layerspp.py:259-266: Conv_0 + Dense_0(act(temb)) -> GroupNorm_1 -> act, read only
   *    by Conv_1): v = acc + bias + rowvec is never materialised (out_bf16 / stats / resid must be NULL, alpha 1): the
   *    sample's accumulators stay resident in TMEM between the statistics pass and the normalising pass;
   *  - with out_f32 (+ stats, optionally resid / alpha / a raw bf16 copy in out_bf16; layerspp.py:245-251,78: Conv_1 / NIN_3
   *    whose result is the next block's GroupNorm_0 input): v is written as usual and the normalising pass re-reads it.
   * Shapes whose sample does not fit the tile / TMEM scheme fall back inside the engine to GEMM + gn_finalize + gn_apply. */
  void* gn_out_bf16; const float* gn_gamma; const float* gn_beta; int gn_groups; float gn_eps; int gn_silu;
} dp_gemm_desc;

typedef struct {
  const float* src0; const float* stats0; int C0; int P0; /* fp32 NHWC + [B][P0][C0][2] partials;
                                                             stats0 == NULL: identity (cast / resample only) */
  int src0_is_bf16;                                       /* src0 points to bf16 data (C1 == 0, resample == 0) */
  const float* src1; const float* stats1; int C1; int P1; /* optional channel-concat second source */
  const float* gamma; const float* beta;                  /* [C0+C1] */
  const float* film; int film_ld;                         /* optional [B, film_ld]: scale = [:C], shift = [C:2C] */
  int B, H, W;            /* input grid */
  int groups; float eps;
  int silu;
  int resample;           /* 0 none, 1 nearest x2, 2 2x2 mean */
  void* out_bf16;         /* [B, H', W', C] normalised (+SiLU) */
  void* raw_bf16;         /* optional: resampled raw input as bf16 (1x1 shortcut operand) */
  float* raw_f32;         /* optional: resampled raw input as fp32 (identity residual after resample) */
} dp_gn_desc;

/* Row softmax of fp32 logits [rows, T] -> normalised bf16 probabilities (long-sequence attention, T > 256). */
typedef struct {
  const float* src; void* out_bf16; long long rows; int T;
} dp_softmax_desc;

typedef struct {
  const float* src; int B, HW, C; /* fp32 [B, HW, C] */
  float* stats;                   /* [B][P][C][2], P = ceil(HW/128) */
} dp_stats_desc;

typedef struct {
  const float* w;    /* [27][Cout] fp32: ((ky*3+kx)*3 + ci) major */
  const float* bias; /* [Cout] */
  float* out;        /* fp32 NHWC [B,H,W,Cout] */
  float* stats;      /* optional per-channel partials of `out` ([B][P][Cout][2]) */
  int B, H, W, Cout;
} dp_conv_in_desc;

/* Output stage when the C->3|6 conv runs as a dp_op_gemm (N padded to 8): consumes its fp32 result and either
 * returns it (dp_unet_forward) or applies the fused per-step update of dp_purify. */
typedef struct {
  const float* eps; int ld; /* fp32 [B*H*W, ld], first Cout columns valid */
  int B, H, W, Cout;
} dp_update_desc;

/* The whole AttnBlockpp behind its GroupNorm (score_sde/models/layerspp.py:75-91: NIN_0..2, einsum . C^-1/2, softmax,
 * einsum, NIN_3, (x + h) / sqrt 2) as ONE kernel for T = H*W = 256 tokens of C = 256 channels: q, k, v, the logits and P
 * stay in tensor memory / shared memory of a CTA pair; replaces four dp_op_gemm launches + the output projection. */
typedef struct {
  const void* hn_bf16;  /* GroupNorm_0(x), bf16 [B*T, C] */
  const void* w_bf16;   /* bf16 [4*C, C]: Wq | Wk | Wv | W3, each [out, in] (NIN.W transposed) */
  const float* bias;    /* [4*C]: bq | bk | bv | b3 */
  const float* resid;   /* x, fp32 [B*T, C] */
  float* out_f32;       /* (x + h) * alpha, fp32 [B*T, C] */
  float* stats;         /* optional: per-(128-row tile, channel) (sum, sumsq) of out, [B*T/128][C][2] (as dp_gemm_desc.stats) */
  int B, T, C;          /* T and C must be 256 */
  float scale;          /* applied to q.k before softmax */
  float alpha;
} dp_attn_block_desc;
int dp_op_attn_block(dp_engine* e, const dp_attn_block_desc* d);

typedef struct {
  const void* qkv_bf16; /* [B*T, 3*heads*d]: q | k | v, each [heads][d] */
  void* out_bf16;       /* [B*T, heads*d] */
  int B, T, heads, d;
  float scale;          /* applied to q.k before softmax */
} dp_attn_small_desc;

/* ---- data-gradient ops (input gradient of the score network; the reference differentiates through the loop with
 *      torchsde's adjoint, runners/diffpure_sde.py:233-239, for the white-box attacks of eval_sde_adv.py:126-128) -------- */

/* GroupNorm(+SiLU, +resample, +concat) backward of a dp_op_gn_apply: sources / statistics are the forward tensors, g is
 * dL/d(output) in fp32 NHWC at the forward OUTPUT resolution (C0+C1 channels). Results:
 *   d0 = dL/d(src0) + add0_scale * resample^T(add0)[:, :C0] + add1      (fp32 and / or bf16)
 *   d1 = dL/d(src1) + add0_scale * resample^T(add0)[:, C0:]             (fp32; the skip connection's gradient) */
typedef struct {
  const float* src0; const float* stats0; int C0; int P0; int src0_is_bf16;
  const float* src1; const float* stats1; int C1; int P1;
  const float* gamma; const float* beta;
  int B, H, W; int groups; float eps; int silu; int resample;
  const float* g;
  const float* add0; float add0_scale; /* optional, output resolution, C0+C1 channels (shortcut branch) */
  const float* add1;                   /* optional, [B,H,W,C0] (gradient that reached src0 through a skip connection) */
  float* d0_f32; void* d0_bf16; float* d1_f32;
  const float* film; int film_ld;      /* optional scale-shift rows of the forward op (guided_diffusion unet.py:255-258), first source only */
} dp_gn_bwd_desc;

/* softmax backward per row: P = pnum / rowsum (rowsum NULL: pnum is P), dS = P (dP - sum_j dP_j P_j); writes dS and P as bf16. */
typedef struct {
  const void* pnum_bf16; const float* rowsum; const float* dp; void* ds_bf16; void* pn_bf16; long long rows; int T;
} dp_softmax_bwd_desc;

/* batched bf16 transpose: out[b][c][r] = in[b][r][c] */
typedef struct {
  const void* in_bf16; void* out_bf16; int rows, cols, ld_in, ld_out, batch; long long in_batch_stride, out_batch_stride;
} dp_transpose_desc;

/* backward of dp_op_attn_small: qkv as in the forward op, go = dL/d(out) [B*T, heads*d]; out = dq | dk | dv [B*T, 3*heads*d] */
typedef struct {
  const void* qkv_bf16; const void* go_bf16; void* out_bf16; int B, T, heads, d; float scale;
} dp_attn_small_bwd_desc;

/* dL/d(UNet output), handed to dp_unet_vjp as NCHW fp32 [B,C,H,W] -> bf16 NHWC [B,H,W,Cpad], zero padded */
typedef struct {
  void* out_bf16; int B, H, W, C, Cpad;
} dp_grad_in_desc;

int dp_op_gn_bwd(dp_engine* e, const dp_gn_bwd_desc* d);
int dp_op_softmax_bwd(dp_engine* e, const dp_softmax_bwd_desc* d);
int dp_op_transpose(dp_engine* e, const dp_transpose_desc* d);
int dp_op_attn_small_bwd(dp_engine* e, const dp_attn_small_bwd_desc* d);
int dp_op_grad_in(dp_engine* e, const dp_grad_in_desc* d);

/* The engine's state x (fp32, 3 channels) as a bf16 NHWC tensor zero-padded to Cpad channels: the A operand of the 3->C
 * input conv (ncsnpp.py:268, unet.py:486, unet_ddpm.py:229) when it runs as a dp_op_gemm on the tensor cores. */
typedef struct {
  void* out_bf16; int B, H, W, Cpad;
} dp_pad_in_desc;
int dp_op_pad_in(dp_engine* e, const dp_pad_in_desc* d);

int dp_op_embed(dp_engine* e, const dp_embed_desc* d);
int dp_op_gemm(dp_engine* e, const dp_gemm_desc* d);
int dp_op_gn_apply(dp_engine* e, const dp_gn_desc* d);
int dp_op_stats(dp_engine* e, const dp_stats_desc* d);
int dp_op_conv_in(dp_engine* e, const dp_conv_in_desc* d);
int dp_op_attn_small(dp_engine* e, const dp_attn_small_desc* d);
int dp_op_softmax_rows(dp_engine* e, const dp_softmax_desc* d);
int dp_op_update(dp_engine* e, const dp_update_desc* d);
int dp_program_size(const dp_engine* e);

/* Freeze the program for images of [B, C=3, H, W]; captures the CUDA graphs. */
int dp_finalize(dp_engine* e, int B, int H, int W);

/* ---- execution ------------------------------------------------------------------------------- */

/* Stream contract of dp_unet_forward / dp_purify: `stream` (a cudaStream_t) != NULL -> all work is enqueued on that
 * stream, ordered behind what the caller enqueued before, and the call returns without synchronising; `stream` == NULL ->
 * the call waits for the caller's default-stream work, runs on the engine's private stream and blocks until the result
 * is complete. Calls on one engine must not overlap (shared run state).
 *
 * One UNet evaluation: x [B,3,H,W] fp32 (device), cond [B] fp32 (device; DDPM++: 999*t, ADM/ddpm: timestep),
 * out [B,Cout,H,W] fp32 (device). Mirrors model(x, t) of the reference modules. */
int dp_unet_forward(dp_engine* e, const float* x_nchw, const float* cond, float* out_nchw, void* stream);

/* Vector-Jacobian product of one UNet evaluation for a program that contains the forward ops followed by the data-gradient
 * ops (diffpure_b200/lowering_ncsnpp.py:lower_vjp): gx [B,3,H,W] = J(x, cond)^T g, g = dL/d(UNet output) [B,C,H,W].
 * Same stream contract as dp_unet_forward. */
int dp_unet_vjp(dp_engine* e, const float* x_nchw, const float* cond, const float* g_nchw, float* gx_nchw, void* stream);

#define DP_UPDATE_LINEAR 0 /* x <- c[0]*x + c[1]*eps + c[2]*z                 (VP-SDE Euler-Maruyama; ddpm fixed-var) */
#define DP_UPDATE_LEARNED_RANGE 1 /* guided_diffusion p_sample with learned-range variance and x0 clamp; 8 coefs */
#define DP_UPDATE_LINEAR_ANCHORED 2 /* x <- c[0]*x + c[1]*eps + c[2]*z + c[3]*x_init  (Langevin-dynamics SDE, runners/diffpure_ldsde.py) */

typedef struct {
  int steps;
  int update_kind;
  int ncoef;            /* coefficients per step (3 or 8) */
  const float* cond;    /* host [steps]: UNet condition of step k (same for the whole batch) */
  const float* coef;    /* host [steps][ncoef] */
  float init_scale_x;   /* forward diffusion x = init_scale_x * x0 + init_scale_e * e  */
  float init_scale_e;
  const float* init_noise;  /* device [B,3,H,W] or NULL -> counter-based generator */
  const float* step_noise;  /* device [steps,B,3,H,W] standard normals or NULL -> counter-based generator */
  uint64_t seed;
  uint64_t sample_offset;   /* global index of sample 0 (multi-GPU sharding keeps streams identical) */
  const float* anchor;      /* DP_UPDATE_LINEAR_ANCHORED: device [B,3,H,W] x_init, or NULL -> the (diffused) initial state */
  float* states;            /* optional device [steps+1,B,3,H,W]: the state before every step and the final one (the
                               discretise-then-differentiate backward pass replays the loop from them) */
  /* Fused pre / post steps of the caller (SDE_Adv_Model.forward, eval_sde_adv.py:73-89; classifier wrappers,
   * utils.py:144-153). All zero = x0 and out are [B,3,H,W] in [-1,1] (the runner API). */
  int in_h, in_w;           /* x0 is [B,3,in_h,in_w]: bilinear resize (align_corners = False) to the model grid; 0 = model grid */
  int in_unit_range;        /* 1: x0 is in [0,1], mapped to [-1,1] after the resize */
  int out_h, out_w;         /* out is [B,3,out_h,out_w]: bilinear resize of the purified image; 0 = model grid */
  int out_unit_range;       /* 1: (x + 1) / 2 */
  float out_mean[3], out_std[3]; /* classifier normalisation (x - mean) / std behind the range map; out_std[0] == 0: none */
} dp_purify_params;

/* The whole purification loop on the device: forward-diffuse, then `steps` x (UNet + fused update).
 * x0, out: [B,3,H,W] fp32 device, values in [-1,1] (runners/diffpure_sde.py:197-247). */
int dp_purify(dp_engine* e, const float* x0_nchw, float* out_nchw, const dp_purify_params* p, void* stream);

/* Measurement aid: runs the program once, op by op (mode 0 = forward, 1 = step without advancing the step
 * counter), each launch bracketed by CUDA events on the engine's stream. ms[i] = device time of op i,
 * kinds[i] = 0 embed,1 gemm,2 gn_apply,3 stats,4 stats_reduce,5 conv_in,6 attn_small,7 softmax_rows,8 update,
 * 9 gn_bwd,10 softmax_bwd,11 transpose,12 attn_small_bwd,13 grad_in,14 gn_finalize,15 pad_in,16 attn_block,
 * flops[i] = 2*M*N*K*batch executed by GEMM op i (0 otherwise). */
int dp_profile_ops(dp_engine* e, int mode, float* ms, int* kinds, double* flops, int cap);

/* Host evaluation of the counter-based normal generator (tests). */
float dp_normal_host(uint64_t seed, uint64_t sample, uint32_t stream, uint32_t pixel, int c);

/* Number of kernels one UNet evaluation launches (dp_unet_forward; also the length of dp_profile_ops' arrays). */
int dp_launches_per_eval(const dp_engine* e);
/* Number of kernels one loop step launches (UNet evaluation + update). When the output conv's epilogue applies the update
 * (the C -> 3|6 conv on the narrow tcgen05 tile) this is dp_launches_per_eval - 1: no update kernel, no step-counter kernel. */
int dp_launches_per_step(const dp_engine* e);

/* How many of the program's GEMM ops run on CTA-pair (tcgen05 cta_group::2) tiles (tests / reporting). */
int dp_gemm_pair_count(const dp_engine* e);
/* How many GEMM ops carry the fused GroupNorm epilogue (requested and resident in TMEM, i.e. not fallen back). */
int dp_gemm_fused_gn_count(const dp_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* DIFFPURE_B200_H_ */
