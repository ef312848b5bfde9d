// dp_gemm.cuh -- parameter block of the tcgen05 implicit-GEMM convolution / GEMM kernel.
//
// One kernel family covers every dense contraction of the three score-model UNets
// (reference call sites: 3x3 conv  score_sde/models/layers.py:118-124, guided_diffusion/unet.py:193,219,
//  ddpm/unet_ddpm.py:95-108; 1x1 conv / NIN layers.py:100-105,546-555, unet.py:230, unet_ddpm.py:117-121;
//  stride-2 conv unet_ddpm.py:63-82; attention matmuls layerspp.py:75-91, unet.py:345-362,
//  unet_ddpm.py:172-197; time-embedding Linears ncsnpp.py:246-255, unet.py:478-483):
//
//   D[M, N] = sum_seg sum_tap sum_c  A_seg[pixel(m) + tap, c] * Wt[n, k(seg, tap, c)]
//
// A operand: bf16 NHWC activations read through 4-D TMA tensor maps (C, W, H, B); the 3x3 halo and
// the zero padding come from TMA out-of-bounds zero fill, a stride-2 conv from the map's element
// strides. Up to two K segments (main 3x3 input + fused 1x1 shortcut input).
// B operand: bf16 weights [N, Ktotal] (K contiguous) through a 2-D tensor map.
// Accumulation: fp32 in TMEM (tcgen05.mma kind::f16), 128 x BN tile per CTA, double-buffered.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace dp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // bf16 elements = one 128-byte swizzle row

struct GemmASeg {
  CUtensorMap tmap;  // (C, W, H, B) bf16, box (64, bw, bh, bn), SWIZZLE_128B; patch mode: box (64, bw, bh + 2, 1)
  int patch;         // 1: 3x3 taps read from shared row patches (see GemmParams::stage_bytes)
  int taps;          // 1 or 9
  int kchunks;       // C / 64
  int stride;        // coordinate multiplier (2 for the stride-2 downsample conv)
  int pad;           // subtracted from the tap offset (1 for 'same' 3x3, else 0)
};

struct GemmParams {
  GemmASeg a[2];
  CUtensorMap tmap_b;  // (Ktotal, Nrows) bf16, box (64, BN), SWIZZLE_128B
  int nseg;
  // output pixel grid of one batch entry (plain GEMM: H = 1, W = M)
  int H, W, bw, bh;
  int tiles_w, tiles_per_img;  // tiles_per_img = tiles_w * tiles_h, 0 if several images share a tile
  int imgs_per_tile;           // 128 / (H*W) when H*W < 128, else 1
  int M;                       // valid rows per batch entry
  int N;                       // valid columns
  int m_tiles, n_tiles, batch;
  int a_batch_rows;  // added to the W coordinate of A per (outer) batch entry (batched GEMM)
  int b_batch_rows;  // added to the row coordinate of B per (outer) batch entry
  // two-level batch (attention heads): batch index = outer * inner + head
  int inner;             // heads per outer entry (>= 1)
  int a_inner_k;         // A channel-coordinate offset per head
  int a_inner_rows;      // A row (W coordinate) offset per head
  int b_inner_k;         // B K-coordinate offset per head
  int b_inner_rows;      // B row offset per head
  long long out_inner_stride;  // output element offset per head
  int num_stages;
  // Patch mode of a 3x3 segment (stride 1, one image per tile, W in {16, 32, 64}): the nine taps of a 64-channel chunk are
  // NOT nine 16 KiB TMA tiles (each A byte fetched from L2 nine times -- the N = 128 convolutions ran at the L2 throughput cap,
  // 47 B/clk/SM of the chip's ~43) but three row patches, one per horizontal shift kx: box (64, bw, bh + 2, 1) = the tile's rows
  // plus one above and one below; the tap (ky, kx) operand is the patch from row ky on -- a plain K-major tile at a
  // 1024-byte-aligned offset ky * bw * 128 B. A stage then holds one patch + the three weight tiles of its taps: A bytes
  // from L2 drop by 9 bh / (3 (bh + 2)) = 2x at 32x32, 2.4x at 16x16. stage_bytes == 0: every stage is one 16 KiB A tile +
  // one weight tile.
  int stage_bytes;    // bytes per pipeline stage in patch mode (a_patch_bytes + 3 weight tiles), multiple of 1024
  int a_patch_bytes;  // (bh + 2) * bw * 128
  // ---- epilogue ----
  const float* bias;  // [N] (or [M] if bias_along_m)
  int bias_along_m;
  const float* rowvec;  // per-sample additive vector: rowvec[(row >> rowvec_shift) * rowvec_ld + col]
  int rowvec_ld, rowvec_shift;
  const float* rowscale;  // out *= 1 / rowscale[b*M + row]
  const float* resid;     // fp32 residual, same addressing as out_f32
  float alpha;            // final scale
  int silu;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  long long ldc;               // row stride (elements) of out / resid
  long long out_batch_stride;  // elements
  float* stats;                // [segments][N][2] per-channel (sum, sumsq) partials, or null
  int stat_nseg;               // row segments per tile: 1 (HW>=128), 2 (HW=64), 8 (HW=16)
  // ---- softmax epilogue (kSoftmax kernels): P = exp(scale*(S - rowmax)) (bf16), rowsum out ----
  float softmax_scale;
  float* rowsum_out;  // [batch*M]
  // ---- accumulator staging ----
  int tpg;         // tiles per work unit and CTA pair: consecutive M units of ONE sample (1 unless the fused GroupNorm
                   // epilogue keeps a 32x32 sample resident in TMEM)
  int upc;         // CTA pairs per work unit (1; 2 = a "super-pair": two adjacent CTA pairs share one 32x32 sample, two
                   // 256-row tiles each, so that half of every SM's TMEM stays free for the next sample's MMAs while the
                   // statistics are exchanged -- through global memory, pairs of different clusters share no DSMEM)
  int acc_stages;  // TMEM accumulator stages (2; 4 = all 512 columns with BN = 128)
  // ---- fused GroupNorm(+SiLU) output (E_GN kernels): gn_out = act(GN(acc + bias + rowvec)) as bf16, nothing else ----
  // The statistics of a sample need every row of the sample: its tiles stay in TMEM (tpg stages), pass 1 reduces the
  // per-channel sums (CTA pairs exchange theirs through DSMEM), pass 2 re-reads TMEM and writes the normalised operand.
  const float* gn_gamma;
  const float* gn_beta;
  __nv_bfloat16* gn_out;  // same addressing as out_bf16 (ldc, batch strides)
  int gn_cpg;             // channels per group (divides BN)
  int gn_hw;              // rows per sample
  float gn_eps;
  int gn_silu;
  int gn_xchg;            // 1: a sample spans both CTAs of the pair (upc == 1)
  int gn_p1;              // 1: resident variant, H*W >= 128: pass 1 keeps group sums per lane (no smem transpose); 0: per-channel pass
  // super-pair exchange (upc == 2), engine-owned, shared by all launches of a stream:
  float* xg_data;         // [super-pairs][2 parities][4 CTAs][BN / cpg groups][2] partial (sum, sum of squares)
  unsigned long long* xg_flag;   // [super-pairs][4 CTAs]: token of the last unit each CTA published (64-bit: never wraps)
  unsigned long long* xg_epoch;  // [0] launch epoch (tokens of earlier launches are always smaller), [1] CTA arrival counter
  // ---- fused per-step update (E_UPDATE kernel: the C -> 3|6 output conv, BN = 32): the epilogue applies the SDE / DDPM
  // update to the state instead of writing eps, draws the noise, and the last CTA advances the step counter ----
  float* upd_x;             // state, NHWC fp32 [B*H*W, 3]; null = plain epilogue
  const float* upd_x_init;  // anchor of the Langevin-dynamics update
  const int* upd_step;      // device step counter
  const float* upd_coef;    // [steps][8]
  const void* upd_call;     // dp::CallParams*: noise source, seed, sample offset, update kind, state recording
  int* upd_arrive;          // CTA arrival counter next to the step counter (null: do not advance, profiling)
  int* upd_step_rw;
  int upd_cout, upd_hw, upd_B;
};

// Launches the persistent kernel (grid = min(tiles, num_sms)). BN in {128, 256}.
// Returns cudaError_t as int.
// cg = 2: CTA pairs (clusters of 2, tcgen05 cta_group::2) on 256 x BN tiles; needs gemm_pair_supported(), a B tensor map
// with BN/2-row boxes and geometry filled with the same cg.
int launch_gemm(const GemmParams& p, int bn, bool softmax, int num_sms, cudaStream_t stream, int cg = 1);
bool gemm_pair_supported(const GemmParams& p, int bn, bool softmax);
// Dynamic shared memory needed for (bn, stages); and the max stage count that fits. gn: the fused-GroupNorm kernels
// (p.gn_out != nullptr) carry a scale/shift table and the pair exchange buffers.
size_t gemm_smem_bytes(int bn, int stages, int cg = 1, bool gn = false);
int gemm_max_stages(int bn, int cg = 1, bool gn = false);
// Switches segment 0 (a stride-1 3x3 conv on one-image tiles with W in {16, 32, 64}) to patch mode when two or more of its
// stages fit in shared memory: fills a[0].patch, a_patch_bytes, stage_bytes, num_stages (geometry and gn_out must be set).
// The caller then builds a[0].tmap with a (64, bw, bh + 2, 1) box.
bool gemm_enable_patch(GemmParams& p, int bn, int cg);
// One-time cudaFuncSetAttribute for all instantiations.
int gemm_init();

// Tile box of the output pixel grid: full rows when W < 128, else 128-pixel row segments.
struct TileBox {
  int bw, bh, bn;
};
TileBox gemm_tile_box(int H, int W);
// Fills H/W/bw/bh/tiles_*/imgs_per_tile/M/N/m_tiles/n_tiles/stat_nseg/num_stages for an output grid of
// B images of H x W pixels (plain GEMM: B = 1, H = 1, W = rows) and N output columns.
void gemm_fill_geometry(GemmParams& p, int B, int H, int W, int N, int bn, int cg = 1);

}  // namespace dp
