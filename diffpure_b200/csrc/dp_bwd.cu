// dp_bwd.cu -- data-gradient kernels of the score network (white-box attacks differentiate through the purification
// loop: eval_sde_adv.py:126-128 through runners/diffpure_sde.py:233-239 sdeint_adjoint). The weights are frozen, so every
// layer contributes only its input gradient:
//   conv / NIN / Linear : the same tcgen05 implicit GEMM with the weights flipped over the taps and transposed
//                         (dp_gemm.cu; packing on the host, lowering_ncsnpp.py)
//   GroupNorm (+SiLU, + nearest-up / 2x2-mean-down, + channel concat) : gn_bwd_stats + gn_bwd_apply below
//   attention            : GEMMs + softmax_bwd (row-wise) + batched bf16 transposes; short sequences (T <= 64) in one
//                          shared-memory kernel
// Everything is deterministic (fixed summation order, no atomics). Restated on the CPU by oracle/ncsnpp_vjp.py.
#include "dp_elem.cuh"
#include "dp_launch.cuh"

#include <cstdio>

namespace dp {

namespace {

__device__ __forceinline__ float bf16_to_f(const __nv_bfloat16 v) { return __bfloat162float(v); }

// d silu(u) / du = s (1 + u (1 - s)), s = sigmoid(u)
__device__ __forceinline__ float silu_grad_f(float u) {
  const float s = 1.0f / (1.0f + __expf(-u));
  return s * (1.0f + u * (1.0f - s));
}

// Group statistics of the forward pass from the producer's per-channel partial sums: sc/sh = per-channel sum / sum of
// squares, gs[2g] = mean, gs[2g+1] = rstd (same arithmetic as the forward gn_apply kernel, dp_elem.cu).
__device__ void fold_group_stats(const GnBwdParams& p, int b, float* sc, float* sh, float* gs) {
  const int C = p.C0 + p.C1;
  const int G = p.groups;
  const int tid = threadIdx.x;
  for (int c = tid; c < C; c += blockDim.x) {
    const float* st;
    int P, Cx, cl;
    if (c < p.C0) { st = p.stats0; P = p.P0; Cx = p.C0; cl = c; }
    else          { st = p.stats1; P = p.P1; Cx = p.C1; cl = c - p.C0; }
    const float2* s2 = reinterpret_cast<const float2*>(st) + (static_cast<size_t>(b) * P * Cx + cl);
    float s = 0.f, q = 0.f;
    for (int pp = 0; pp < P; ++pp) {
      const float2 v = __ldg(s2 + static_cast<size_t>(pp) * Cx);
      s += v.x;
      q += v.y;
    }
    sc[c] = s;
    sh[c] = q;
  }
  __syncthreads();
  const int cpg = C / G;
  const int HW = p.H * p.W;
  for (int g = tid; g < G; g += blockDim.x) {
    double S = 0.0, Q = 0.0;
    for (int j = 0; j < cpg; ++j) {
      S += sc[g * cpg + j];
      Q += sh[g * cpg + j];
    }
    const double n = static_cast<double>(cpg) * HW;
    const double mean = S / n;
    double var = Q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    gs[2 * g] = static_cast<float>(mean);
    gs[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
  }
  __syncthreads();
}

// Gradient arriving at the GroupNorm(+act) output of input pixel (h, w), channel c: the transpose of the forward resample
// applied to `t` (fp32 NHWC at the forward OUTPUT resolution, `ct` channels): nearest x2 -> sum of the 4 children,
// 2x2 mean -> a quarter of the parent.
__device__ __forceinline__ float fetch_resampled_T(const float* __restrict__ t, int resample, int b, int h, int w, int H,
                                                   int W, int ct, int c) {
  if (resample == 1) {
    const int Wo = 2 * W;
    const float* r0 = t + ((static_cast<size_t>(b) * 2 * H + 2 * h) * Wo + 2 * w) * ct + c;
    const float* r1 = r0 + static_cast<size_t>(Wo) * ct;
    return (__ldg(r0) + __ldg(r0 + ct)) + (__ldg(r1) + __ldg(r1 + ct));
  }
  if (resample == 2) {
    const int Ho = H / 2, Wo = W / 2;
    return 0.25f * __ldg(t + ((static_cast<size_t>(b) * Ho + (h >> 1)) * Wo + (w >> 1)) * ct + c);
  }
  return __ldg(t + ((static_cast<size_t>(b) * H + h) * W + w) * ct + c);
}

__device__ __forceinline__ float load_x(const GnBwdParams& p, int b, int pix, int c) {
  const int HW = p.H * p.W;
  if (c < p.C0) {
    const size_t i = (static_cast<size_t>(b) * HW + pix) * p.C0 + c;
    return p.src0h ? bf16_to_f(p.src0h[i]) : __ldg(p.src0 + i);
  }
  return __ldg(p.src1 + (static_cast<size_t>(b) * HW + pix) * p.C1 + (c - p.C0));
}

// FiLM / scale-shift norm (guided_diffusion unet.py:255-258: h = norm(h) * (1 + scale) + shift, then SiLU): for the input
// gradient the per-sample scale and shift just replace gamma, beta by gamma (1 + scale), beta (1 + scale) + shift.
__device__ __forceinline__ void film_fold(const GnBwdParams& p, int b, int c, int C, float& gam, float& bet) {
  if (p.film == nullptr) return;
  const float sc = 1.0f + __ldg(p.film + static_cast<size_t>(b) * p.film_ld + c);
  const float sh = __ldg(p.film + static_cast<size_t>(b) * p.film_ld + C + c);
  bet = bet * sc + sh;
  gam = gam * sc;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// GroupNorm backward. y = act(xhat * gamma + beta), xhat = (x - mean_g) * rstd_g, a = resample(y); given g = dL/da:
//   g_y = resample^T(g) * act'(u);  gx = g_y * gamma;  m1 = <gx>_group, m2 = <gx xhat>_group
//   dL/dx = rstd (gx - m1 - xhat m2)                                   (oracle/ncsnpp_vjp.py:gn_silu_vjp)
// Pass 1 (stats): per 128-pixel chunk and channel the partial sums (sum gx, sum gx xhat) -> part [B][P][C][2].
// Pass 2 (apply): folds the partials per group and writes dL/dx (+ the shortcut / skip gradients) as fp32 and bf16.
// One thread = one channel, walking the chunk's pixels (adjacent threads = adjacent channels: coalesced NHWC rows).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(GnBwdParams p) {
  pdl_entry();
  extern __shared__ float gsm[];
  const int C = p.C0 + p.C1;
  float* gs = gsm + 2 * C;
  const int b = blockIdx.y, chunk = blockIdx.x;
  fold_group_stats(p, b, gsm, gsm + C, gs);
  const int HW = p.H * p.W;
  const int P = (HW + 127) / 128;
  const int px0 = chunk * 128, px1 = min(HW, px0 + 128);
  const int cpg = C / p.groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = gs[2 * g], rstd = gs[2 * g + 1];
    float gam = __ldg(p.gamma + c);
    float bet = p.silu ? __ldg(p.beta + c) : 0.f;
    film_fold(p, b, c, C, gam, bet);
    float s1 = 0.f, s2 = 0.f;
    for (int pix = px0; pix < px1; ++pix) {
      const int h = pix / p.W, w = pix - h * p.W;
      const float xhat = (load_x(p, b, pix, c) - mean) * rstd;
      float gy = fetch_resampled_T(p.g, p.resample, b, h, w, p.H, p.W, C, c);
      if (p.silu) gy *= silu_grad_f(xhat * gam + bet);
      const float gx = gy * gam;
      s1 += gx;
      s2 += gx * xhat;
    }
    float* o = p.part + ((static_cast<size_t>(b) * P + chunk) * C + c) * 2;
    o[0] = s1;
    o[1] = s2;
  }
}

__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(GnBwdParams p) {
  pdl_entry();
  extern __shared__ float gsm[];
  const int C = p.C0 + p.C1;
  float* gs = gsm + 2 * C;          // forward group mean / rstd
  float* gm = gs + 2 * p.groups;    // backward group means m1, m2
  const int b = blockIdx.y, chunk = blockIdx.x;
  fold_group_stats(p, b, gsm, gsm + C, gs);
  const int HW = p.H * p.W;
  const int P = (HW + 127) / 128;
  const int cpg = C / p.groups;
  // per-channel totals of the backward partials (fixed order), then per-group means
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int pp = 0; pp < P; ++pp) {
      const float2 v = *reinterpret_cast<const float2*>(p.part + ((static_cast<size_t>(b) * P + pp) * C + c) * 2);
      s1 += v.x;
      s2 += v.y;
    }
    gsm[c] = s1;
    gsm[C + c] = s2;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
    double S1 = 0.0, S2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
      S1 += gsm[g * cpg + j];
      S2 += gsm[C + g * cpg + j];
    }
    const double n = static_cast<double>(cpg) * HW;
    gm[2 * g] = static_cast<float>(S1 / n);
    gm[2 * g + 1] = static_cast<float>(S2 / n);
  }
  __syncthreads();
  const int px0 = chunk * 128, px1 = min(HW, px0 + 128);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = gs[2 * g], rstd = gs[2 * g + 1], m1 = gm[2 * g], m2 = gm[2 * g + 1];
    float gam = __ldg(p.gamma + c);
    float bet = p.silu ? __ldg(p.beta + c) : 0.f;
    film_fold(p, b, c, C, gam, bet);
    for (int pix = px0; pix < px1; ++pix) {
      const int h = pix / p.W, w = pix - h * p.W;
      const float xhat = (load_x(p, b, pix, c) - mean) * rstd;
      float gy = fetch_resampled_T(p.g, p.resample, b, h, w, p.H, p.W, C, c);
      if (p.silu) gy *= silu_grad_f(xhat * gam + bet);
      float d = rstd * (gy * gam - m1 - xhat * m2);
      if (p.add0) d += p.add0_scale * fetch_resampled_T(p.add0, p.resample, b, h, w, p.H, p.W, C, c);
      if (c < p.C0) {
        const size_t o = (static_cast<size_t>(b) * HW + pix) * p.C0 + c;
        if (p.add1) d += __ldg(p.add1 + o);
        if (p.d0_f32) p.d0_f32[o] = d;
        if (p.d0_bf16) p.d0_bf16[o] = __float2bfloat16_rn(d);
      } else {
        p.d1_f32[(static_cast<size_t>(b) * HW + pix) * p.C1 + (c - p.C0)] = d;
      }
    }
  }
}

int launch_gn_bwd(const GnBwdParams& p, cudaStream_t s) {
  const int C = p.C0 + p.C1;
  const int HW = p.H * p.W;
  const int P = (HW + 127) / 128;
  const size_t smem = static_cast<size_t>(2 * C + 4 * p.groups) * sizeof(float);
  const int threads = C < 256 ? ((C + 31) / 32) * 32 : 256;
  (void)launch_k(gn_bwd_stats_kernel, dim3(P, p.B), dim3(threads), smem, s, 1, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  (void)launch_k(gn_bwd_apply_kernel, dim3(P, p.B), dim3(threads), smem, s, 1, p);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// softmax backward, one warp per row: P = Pnum / rowsum (forward: Pnum = exp(scale (S - rowmax)) in bf16),
// dS = P (dP - sum_j dP_j P_j). Writes dS and the normalised P, both bf16.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __nv_bfloat16* __restrict__ pnum,
                                                          const float* __restrict__ rowsum,
                                                          const float* __restrict__ dp_, __nv_bfloat16* __restrict__ ds,
                                                          __nv_bfloat16* __restrict__ pn, long long rows, int T) {
  pdl_entry();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float inv = rowsum != nullptr ? 1.0f / rowsum[row] : 1.0f;   // no row sums: pnum is already normalised
  const __nv_bfloat16* pr = pnum + row * T;
  const float* dr = dp_ + row * T;
  float dot = 0.f;
  for (int j = lane; j < T; j += 32) dot += dr[j] * (bf16_to_f(pr[j]) * inv);
  for (int m = 16; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
  for (int j = lane; j < T; j += 32) {
    const float pj = bf16_to_f(pr[j]) * inv;
    ds[row * T + j] = __float2bfloat16_rn(pj * (dr[j] - dot));
    pn[row * T + j] = __float2bfloat16_rn(pj);
  }
}

int launch_softmax_bwd(const __nv_bfloat16* pnum, const float* rowsum, const float* dp_, __nv_bfloat16* ds,
                       __nv_bfloat16* pn, long long rows, int T, cudaStream_t s) {
  (void)launch_k(softmax_bwd_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, 1, pnum, rowsum, dp_,
                 ds, pn, rows, T);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// batched bf16 transpose: out[b][c][r] = in[b][r][c]   (32 x 32 tiles through shared memory)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(TransposeParams p) {
  pdl_entry();
  __shared__ __nv_bfloat16 tile[32][34];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const __nv_bfloat16* in = p.in + static_cast<size_t>(b) * p.in_batch_stride;
  __nv_bfloat16* out = p.out + static_cast<size_t>(b) * p.out_batch_stride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < p.rows && c < p.cols) tile[i][tx] = in[static_cast<size_t>(r) * p.ld_in + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < p.rows && c < p.cols) out[static_cast<size_t>(c) * p.ld_out + r] = tile[tx][i];
  }
}

int launch_transpose(const TransposeParams& p, cudaStream_t s) {
  const dim3 grid((p.cols + 31) / 32, (p.rows + 31) / 32, p.batch);
  (void)launch_k(transpose_kernel, grid, dim3(256), 0, s, 1, p);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// short-sequence attention backward (T <= 64): one CTA per (head, sample); q, k, v, dO and the T x T matrices in smem.
//   P = softmax(scale q k^T); dV = P^T dO; dP = dO V^T; dS = P (dP - rowsum(dP P)); dQ = scale dS K; dK = scale dS^T Q
// qkv: [B*T, 3*heads*d] (q | k | v, as the forward attn_small kernel reads it); go: [B*T, heads*d];
// out: [B*T, 3*heads*d] (dq | dk | dv).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_small_bwd_kernel(AttnSmallBwdParams p) {
  pdl_entry();
  extern __shared__ __align__(16) unsigned char smraw[];
  const int T = p.T, d = p.d;
  const int pitch = d + 2;
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* sk = sq + T * pitch;
  __nv_bfloat16* sv = sk + T * pitch;
  __nv_bfloat16* sg = sv + T * pitch;
  float* sp = reinterpret_cast<float*>(sg + T * pitch);   // P   [T][T+1]   (T * pitch * 4 bf16 = multiple of 4 bytes)
  float* sd = sp + T * (T + 1);                            // dS  [T][T+1]
  const int head = blockIdx.x, b = blockIdx.y;
  const int ld = 3 * p.heads * d;
  const __nv_bfloat16* g = p.qkv + static_cast<size_t>(b) * T * ld + head * d;
  const __nv_bfloat16* go = p.go + static_cast<size_t>(b) * T * (p.heads * d) + head * d;
  for (int i = threadIdx.x; i < T * d; i += blockDim.x) {
    const int r = i / d, c = i - r * d;
    sq[r * pitch + c] = g[static_cast<size_t>(r) * ld + c];
    sk[r * pitch + c] = g[static_cast<size_t>(r) * ld + p.heads * d + c];
    sv[r * pitch + c] = g[static_cast<size_t>(r) * ld + 2 * p.heads * d + c];
    sg[r * pitch + c] = go[static_cast<size_t>(r) * (p.heads * d) + c];
  }
  __syncthreads();
  for (int ij = threadIdx.x; ij < T * T; ij += blockDim.x) {
    const int i = ij / T, j = ij - i * T;
    float acc = 0.f, accp = 0.f;
    for (int c = 0; c < d; ++c) {
      acc += bf16_to_f(sq[i * pitch + c]) * bf16_to_f(sk[j * pitch + c]);
      accp += bf16_to_f(sg[i * pitch + c]) * bf16_to_f(sv[j * pitch + c]);
    }
    sp[i * (T + 1) + j] = acc * p.scale;
    sd[i * (T + 1) + j] = accp;   // dP
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < T; i += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, sp[i * (T + 1) + j]);
    for (int m = 16; m > 0; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = __expf(sp[i * (T + 1) + j] - mx);
      sp[i * (T + 1) + j] = e;
      sum += e;
    }
    for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    const float inv = 1.0f / sum;
    float dot = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float pj = sp[i * (T + 1) + j] * inv;
      sp[i * (T + 1) + j] = pj;
      dot += pj * sd[i * (T + 1) + j];
    }
    for (int m = 16; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
    for (int j = lane; j < T; j += 32) sd[i * (T + 1) + j] = sp[i * (T + 1) + j] * (sd[i * (T + 1) + j] - dot);
  }
  __syncthreads();
  __nv_bfloat16* o = p.out + static_cast<size_t>(b) * T * ld + head * d;
  for (int ic = threadIdx.x; ic < T * d; ic += blockDim.x) {
    const int i = ic / d, c = ic - i * d;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < T; ++j) {
      dq += sd[i * (T + 1) + j] * bf16_to_f(sk[j * pitch + c]);
      dk += sd[j * (T + 1) + i] * bf16_to_f(sq[j * pitch + c]);
      dv += sp[j * (T + 1) + i] * bf16_to_f(sg[j * pitch + c]);
    }
    o[static_cast<size_t>(i) * ld + c] = __float2bfloat16_rn(dq * p.scale);
    o[static_cast<size_t>(i) * ld + p.heads * d + c] = __float2bfloat16_rn(dk * p.scale);
    o[static_cast<size_t>(i) * ld + 2 * p.heads * d + c] = __float2bfloat16_rn(dv);
  }
}

int launch_attn_small_bwd(const AttnSmallBwdParams& p, cudaStream_t s) {
  const int pitch = p.d + 2;
  const size_t smem = static_cast<size_t>(4) * p.T * pitch * 2 + static_cast<size_t>(2) * p.T * (p.T + 1) * 4;
  if (smem > 227 * 1024) return static_cast<int>(cudaErrorInvalidValue);
  cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem > 48 * 1024 ? smem : 48 * 1024));
  if (e != cudaSuccess) return static_cast<int>(e);
  (void)launch_k(attn_small_bwd_kernel, dim3(p.heads, p.B), dim3(256), smem, s, 1, p);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// gradient wrt the UNet output, NCHW fp32 [B, C, H, W] -> NHWC bf16 [B, H, W, Cpad] (zero padded: the dgrad GEMM of the
// output conv reads 64-channel K chunks)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_in_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ out, int B,
                                                      int C, int HW, int Cpad) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * HW * Cpad;
  if (i >= n) return;
  const int c = static_cast<int>(i % Cpad);
  const long long bp = i / Cpad;
  const int pix = static_cast<int>(bp % HW);
  const int b = static_cast<int>(bp / HW);
  out[i] = __float2bfloat16_rn(c < C ? g[(static_cast<size_t>(b) * C + c) * HW + pix] : 0.f);
}

int launch_grad_in(const float* g_nchw, __nv_bfloat16* out, int B, int C, int HW, int Cpad, cudaStream_t s) {
  const long long n = static_cast<long long>(B) * HW * Cpad;
  (void)launch_k(grad_in_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, s, 1, g_nchw, out, B, C, HW,
                 Cpad);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace dp
