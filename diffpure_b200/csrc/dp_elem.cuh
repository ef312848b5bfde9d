// dp_elem.cuh -- the bandwidth-bound kernels around the GEMMs (all NHWC).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace dp {

// Per-step scalars the captured graph reads from device memory (updated by the host loop with
// cudaMemcpyAsync-free pointer indexing: kernels read table[*step]).
struct StepTables {
  const int* step;      // device: current step index
  const float* cond;    // device [steps]
  const float* coef;    // device [steps][ncoef]
  int ncoef;
};

struct EmbedParams {
  __nv_bfloat16* out;  // [B, dim]
  int B, dim, cos_first, half_minus_1;
  const float* cond_per_sample;  // [B] or null -> tables.cond[*tables.step]
  StepTables tables;
};
int launch_embed(const EmbedParams& p, cudaStream_t s);

struct GnParams {
  const float* src0; const float* stats0; int C0, P0;
  const __nv_bfloat16* src0h;  // bf16 source instead of src0 (single source, no resample)
  const float* src1; const float* stats1; int C1, P1;
  const float* gamma; const float* beta;
  const float* film; int film_ld;
  int B, H, W, groups; float eps; int silu; int resample;
  __nv_bfloat16* out; __nv_bfloat16* raw; float* raw_f32;
};
// finalize: partial statistics -> ss [B][2][C] (scale, shift); apply: streaming normalise (+SiLU, resample) -> bf16.
// `ss == nullptr` in apply = identity (cast / resample only).
// (Folding the finalize into every apply CTA was measured in round 2: 163 fewer launches per DDPM++ evaluation, but the
//  redundant per-CTA fold -- 16 CTAs per sample, two barriers and a double-precision group reduction in front of the
//  streaming part -- cost 11.0 ms instead of 7.0 + 1.7 ms per evaluation at B=512; the two-kernel form stays.)
int launch_gn_finalize(const GnParams& p, float* ss, cudaStream_t s);
int launch_gn_apply(const GnParams& p, const float* ss, int num_sms, cudaStream_t s);

// fp32 [B, HW, C] -> per-channel partial (sum, sumsq) [B][P][C][2], P = ceil(HW / 128)
int launch_stats(const float* src, float* stats, int B, int HW, int C, cudaStream_t s);
// [B][P][C][2] -> [B][1][C][2] (in a separate buffer)
int launch_stats_reduce(const float* in, float* out, int B, int P, int C, cudaStream_t s);

struct ConvInParams {
  const float* x;  // fp32 NHWC [B,H,W,3]
  const float* w; const float* bias; float* out; int B, H, W, Cout;
};
int launch_conv_in(const ConvInParams& p, cudaStream_t s);

// Per-call values the captured step graph reads from device memory.
struct CallParams {
  const float* step_noise;  // [steps,B,3,H,W] standard normals, or null -> counter-based generator
  unsigned long long seed, sample_offset;
  int update_kind;  // 0: x <- k0 x + k1 eps + k2 z ; 1: learned-range DDPM step ; 2: kind 0 + k3 x_init
  float* states;    // optional [steps+1,B,3,H,W] NCHW: the update of step k also stores the new state at index k+1
};

// Per-step update from an fp32 eps buffer [B*H*W, ld] (written by the output conv run as a tensor-core GEMM):
// mode 0 copies the first Cout columns to out_nchw, mode 1 applies the fused SDE / DDPM update to x (NHWC [.,3]).
struct UpdateParams {
  const float* eps; int ld; int B, H, W, Cout;
  int mode; float* out_nchw; float* x; const float* x_init;
  StepTables tables; const CallParams* call;
};
int launch_update(const UpdateParams& p, cudaStream_t s);

struct AttnSmallParams {
  const __nv_bfloat16* qkv; __nv_bfloat16* out; int B, T, heads, d; float scale;
};
int launch_attn_small(const AttnSmallParams& p, cudaStream_t s);
// fp32 [rows, T] -> softmax over T -> bf16 [rows, T]
int launch_softmax_rows(const float* src, __nv_bfloat16* out, long long rows, int T, cudaStream_t s);

// x_state[b,h,w,c] = sx * x0[b,c,h,w] + se * noise  (noise: tensor or counter-based normal stream 0)
int launch_init_state(const float* x0_nchw, const float* noise_nchw, float* x_nhwc, int B, int C, int HW,
                      float sx, float se, unsigned long long seed, unsigned long long sample_offset,
                      cudaStream_t s);
int launch_nhwc_to_nchw(const float* x_nhwc, float* out_nchw, int B, int C, int HW, cudaStream_t s);
// engine state (fp32 NHWC, 3 channels) -> bf16 NHWC zero-padded to Cpad channels (A operand of the tensor-core input conv)
int launch_pad_in(const float* x_nhwc3, __nv_bfloat16* out, long long npix, int Cpad, cudaStream_t s);
// Fused pre / post steps of the caller (eval_sde_adv.py:73-89, utils.py:144-153): bilinear resize (align_corners = False)
// + [0,1] -> [-1,1] in front of the forward diffusion; resize + [-1,1] -> [0,1] + classifier normalisation behind the loop.
struct PostParams {
  int unit_range;
  float mean[3], std[3];  // std[0] == 0: no normalisation
};
int launch_init_state_pre(const float* x0_nchw, const float* noise_nchw, float* x_nhwc, int B, int C, int H, int W,
                          int Hin, int Win, int unit_range, float sx, float se, unsigned long long seed,
                          unsigned long long sample_offset, cudaStream_t s);
int launch_final_post(const float* x_nhwc, float* out_nchw, int B, int C, int H, int W, int Ho, int Wo,
                      const PostParams& pp, cudaStream_t s);
int launch_step_advance(int* step, cudaStream_t s);

// ---- data-gradient kernels (dp_bwd.cu) ---------------------------------------------------------------------------------
// GroupNorm(+SiLU, +resample, +concat) backward. Sources / statistics as in GnParams (the forward tensors);
// g: dL/d(output), fp32 NHWC at the forward OUTPUT resolution with C0+C1 channels.
struct GnBwdParams {
  const float* src0; const __nv_bfloat16* src0h; const float* stats0; int C0, P0;
  const float* src1; const float* stats1; int C1, P1;
  const float* gamma; const float* beta;
  const float* film; int film_ld;       // optional per-sample [scale | shift] rows (scale-shift norm), as GnParams.film
  int B, H, W, groups; float eps; int silu; int resample;
  const float* g;
  const float* add0; float add0_scale;  // optional fp32 at the output resolution, C0+C1 channels: d += scale * resample^T(add0)
  const float* add1;                    // optional fp32 [B,H,W,C0]: added to the first source's gradient (skip connection)
  float* part;                          // scratch [B][ceil(HW/128)][C0+C1][2]
  float* d0_f32; __nv_bfloat16* d0_bf16;  // gradient wrt source 0 [B,H,W,C0]
  float* d1_f32;                          // gradient wrt source 1 [B,H,W,C1]
};
int launch_gn_bwd(const GnBwdParams& p, cudaStream_t s);
int launch_softmax_bwd(const __nv_bfloat16* pnum, const float* rowsum, const float* dp_, __nv_bfloat16* ds,
                       __nv_bfloat16* pn, long long rows, int T, cudaStream_t s);
struct TransposeParams {
  const __nv_bfloat16* in; __nv_bfloat16* out; int rows, cols, ld_in, ld_out, batch;
  long long in_batch_stride, out_batch_stride;
};
int launch_transpose(const TransposeParams& p, cudaStream_t s);
struct AttnSmallBwdParams {
  const __nv_bfloat16* qkv; const __nv_bfloat16* go; __nv_bfloat16* out; int B, T, heads, d; float scale;
};
int launch_attn_small_bwd(const AttnSmallBwdParams& p, cudaStream_t s);
int launch_grad_in(const float* g_nchw, __nv_bfloat16* out, int B, int C, int HW, int Cpad, cudaStream_t s);

// Counter-based standard normal: Philox4x32-10 keyed by seed, counter (sample, step+1 | 0 = init, pixel),
// Box-Muller on the four outputs; component c in [0,3).
__host__ __device__ float dp_normal(unsigned long long seed, unsigned long long sample, unsigned int stream,
                                    unsigned int pixel, int c);

}  // namespace dp
