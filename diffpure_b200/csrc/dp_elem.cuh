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
// Streaming normalise (+SiLU, FiLM, resample, concat) -> bf16; every CTA folds its sample's partial statistics into the
// per-channel scale / shift itself. `stats0 == nullptr` = identity (cast / resample only).
int launch_gn_apply(const GnParams& p, int num_sms, cudaStream_t s);

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
int launch_step_advance(int* step, cudaStream_t s);

// Counter-based standard normal: Philox4x32-10 keyed by seed, counter (sample, step+1 | 0 = init, pixel),
// Box-Muller on the four outputs; component c in [0,3).
__host__ __device__ float dp_normal(unsigned long long seed, unsigned long long sample, unsigned int stream,
                                    unsigned int pixel, int c);

}  // namespace dp
