// dp_attn.cuh -- the whole self-attention block of the DDPM++ score network as ONE tcgen05 kernel (sm_100a).
//
// Reference: score_sde/models/layerspp.py:75-91 (AttnBlockpp.forward, single head of width C, skip_rescale):
//   h = GroupNorm_0(x); q = NIN_0(h); k = NIN_1(h); v = NIN_2(h)
//   w = softmax(einsum('bchw,bcij->bhwij', q, k) * C^-1/2) over the T = H W keys
//   h = NIN_3(einsum('bhwij,bcij->bchw', w, v));  return (x + h) / sqrt(2)
// Everything after GroupNorm_0 runs here for T = 256 tokens of C = 256 channels (the 16 x 16 level of the CIFAR-10 model):
// six 256 x 256 x 256 GEMMs per sample on a CTA pair (cta_group::2), q, k, v^T, the logits, P and the attention output
// never leave the SM pair (TMEM accumulators -> bf16 operand tiles in shared memory), the residual / scale / fp32 store /
// GroupNorm partial statistics of the block's output are the last GEMM's epilogue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace dp {

struct AttnBlockParams {
  CUtensorMap tmap_h;  // GroupNorm_0(x), bf16 [B * 256, 256] (token-major, channels contiguous); box (64, 128)
  CUtensorMap tmap_w;  // bf16 [4 * 256, 256]: Wq | Wk | Wv | W3, each [out, in] (the transposed NIN matrices); box (64, 128)
  const float* bias;   // [4 * 256]: bq | bk | bv | b3
  const float* resid;  // x, fp32 [B * 256, 256]
  float* out_f32;      // (x + h) * alpha, fp32 [B * 256, 256]
  float* stats;        // per-(128-row tile, channel) (sum, sum of squares) of the output: [B * 2][256][2], or nullptr
  int B;
  float scale;         // C^-1/2, folded into q
  float alpha;         // 1/sqrt(2) (skip_rescale) or 1
  int x_prefetch;      // when x is pulled into L2 ahead of the output epilogue: 0 never, 1 at the start of the sample, 2 after G4
};

constexpr int kAttnBlockT = 256;
constexpr int kAttnBlockC = 256;

int attn_block_init();  // opt in to the large dynamic shared-memory carve-out (once per device)
int launch_attn_block(const AttnBlockParams& p, int num_sms, cudaStream_t stream);

}  // namespace dp
