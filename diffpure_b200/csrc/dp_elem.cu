// dp_elem.cu -- bandwidth-bound kernels around the tcgen05 GEMMs (sm_100a): GroupNorm finalize / apply (+SiLU, FiLM,
// resample, concat) producing the bf16 GEMM operands where the GroupNorm is not fused into the producing GEMM's epilogue,
// GroupNorm statistics, timestep embedding, the state cast for the tensor-core input conv (and the SIMT input conv it
// replaced), the stand-alone per-step update (forward-mode output; the step graph applies the update in the output conv's
// epilogue, dp_gemm.cu), short-sequence attention, row softmax, the fused pre / post steps and layout conversion.
// All activations are NHWC; vector width is 8 channels (32 B fp32 in, 16 B bf16 out).
#include <cstdlib>

#include "dp_elem.cuh"
#include "dp_launch.cuh"
#include "dp_rng.cuh"

#include <cstdio>

namespace dp {

namespace {

// x * sigmoid(x) with ex2.approx + rcp.approx (2 MUFU ops; relative error ~1e-6, far below the bf16 rounding that follows)
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& a, float& b) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  a = __bfloat162float(t.x);
  b = __bfloat162float(t.y);
}

}  // namespace

__host__ __device__ float dp_normal(unsigned long long seed, unsigned long long sample, unsigned int stream,
                                    unsigned int pixel, int c) {
  return dp_normal_impl(seed, sample, stream, pixel, c);
}

// ------------------------------------------------------------------------------------------------
// timestep embedding
// ------------------------------------------------------------------------------------------------
__global__ void embed_kernel(EmbedParams p) {
  pdl_entry();
  const int b = blockIdx.x;
  const int half = p.dim / 2;
  const float cond = p.cond_per_sample ? p.cond_per_sample[b] : p.tables.cond[*p.tables.step];
  const float coef = -logf(10000.0f) / static_cast<float>(p.half_minus_1 ? half - 1 : half);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float freq = expf(static_cast<float>(i) * coef);
    const float arg = cond * freq;
    const float s = sinf(arg), c = cosf(arg);
    p.out[static_cast<size_t>(b) * p.dim + i] = __float2bfloat16_rn(p.cos_first ? c : s);
    p.out[static_cast<size_t>(b) * p.dim + half + i] = __float2bfloat16_rn(p.cos_first ? s : c);
  }
}

int launch_embed(const EmbedParams& p, cudaStream_t s) {
  (void)launch_k(embed_kernel, dim3(p.B), dim3(128), 0, s, 1, p);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// GroupNorm: finalize (per sample: partial sums -> per-channel scale / shift) + streaming apply
// ------------------------------------------------------------------------------------------------
// ss layout: [B][2][C] fp32 (scale[C] then shift[C]); y = x * scale + shift folds mean, rstd, gamma, beta and FiLM.
__global__ void __launch_bounds__(256) gn_finalize_kernel(GnParams p, float* __restrict__ ss) {
  pdl_entry();
  extern __shared__ float sm[];
  const int C = p.C0 + p.C1;
  const int G = p.groups;
  float* sc = sm;
  float* sh = sm + C;
  float* gs = sm + 2 * C;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int HW = p.H * p.W;
  // affine / FiLM parameters do not depend on the statistics: fetch them first so their latency overlaps the partial sums
  // (C <= 2 * blockDim.x for every UNet here; further channels are fetched in the last loop)
  float pg[2], pb[2], pfs[2], pfb[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = tid + i * 256;
    pg[i] = pb[i] = pfs[i] = pfb[i] = 0.f;
    if (c < C) {
      pg[i] = __ldg(p.gamma + c);
      pb[i] = __ldg(p.beta + c);
      if (p.film) {
        pfs[i] = __ldg(p.film + static_cast<size_t>(b) * p.film_ld + c);
        pfb[i] = __ldg(p.film + static_cast<size_t>(b) * p.film_ld + C + c);
      }
    }
  }
  for (int c = tid; c < C; c += blockDim.x) {
    const float* st;
    int P, Cx, cl;
    if (c < p.C0) { st = p.stats0; P = p.P0; Cx = p.C0; cl = c; }
    else          { st = p.stats1; P = p.P1; Cx = p.C1; cl = c - p.C0; }
    const float2* s2 = reinterpret_cast<const float2*>(st) + (static_cast<size_t>(b) * P * Cx + cl);
    float s = 0.f, q = 0.f;
    int pp = 0;
    for (; pp + 8 <= P; pp += 8) {  // eight independent loads in flight, summed in a fixed order
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(s2 + static_cast<size_t>(pp + u) * Cx);
#pragma unroll
      for (int u = 0; u < 8; ++u) { s += v[u].x; q += v[u].y; }
    }
    for (; pp + 4 <= P; pp += 4) {
      const float2 v0 = __ldg(s2 + static_cast<size_t>(pp) * Cx), v1 = __ldg(s2 + static_cast<size_t>(pp + 1) * Cx);
      const float2 v2 = __ldg(s2 + static_cast<size_t>(pp + 2) * Cx), v3 = __ldg(s2 + static_cast<size_t>(pp + 3) * Cx);
      s += v0.x; q += v0.y; s += v1.x; q += v1.y; s += v2.x; q += v2.y; s += v3.x; q += v3.y;
    }
    for (; pp < P; ++pp) {
      const float2 v = __ldg(s2 + static_cast<size_t>(pp) * Cx);
      s += v.x;
      q += v.y;
    }
    sc[c] = s;
    sh[c] = q;
  }
  __syncthreads();
  const int cpg = C / G;
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
  for (int c = tid, i = 0; c < C; c += blockDim.x, ++i) {
    const int g = c / cpg;
    float gam, bet, fsv = 0.f, fbv = 0.f;
    if (i < 2) { gam = pg[i]; bet = pb[i]; fsv = pfs[i]; fbv = pfb[i]; }
    else {
      gam = p.gamma[c]; bet = p.beta[c];
      if (p.film) { fsv = p.film[static_cast<size_t>(b) * p.film_ld + c]; fbv = p.film[static_cast<size_t>(b) * p.film_ld + C + c]; }
    }
    float a = gam * gs[2 * g + 1];
    float bb = bet - gs[2 * g] * a;
    if (p.film) {
      const float fs = 1.0f + fsv;
      a *= fs;
      bb = bb * fs + fbv;
    }
    ss[(static_cast<size_t>(b) * 2) * C + c] = a;
    ss[(static_cast<size_t>(b) * 2 + 1) * C + c] = bb;
  }
}

int launch_gn_finalize(const GnParams& p, float* ss, cudaStream_t s) {
  const int C = p.C0 + p.C1;
  const size_t smem = static_cast<size_t>(2 * C + 2 * p.groups) * sizeof(float);
  (void)launch_k(gn_finalize_kernel, dim3(p.B), dim3(256), smem, s, 1, p, ss);
  return static_cast<int>(cudaGetLastError());
}

// Streaming apply: plain large grid (measured 5.9-6.1 TB/s for this access shape vs 3.7 TB/s for a persistent loop,
// tools/bench_stream.cu). One CTA = U*rpi output pixels of one sample; scale/shift of the sample staged in smem;
// every thread owns one fixed 8-channel vector and issues all its loads before any compute / store.
template <int RES, bool SRC16, int U>
__global__ void __launch_bounds__(256) gn_apply_kernel(GnParams p, const float* __restrict__ ss) {
  pdl_entry();
  const int C = p.C0 + p.C1;
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int HW = p.H * p.W;
  const int Wo = RES == 1 ? p.W * 2 : (RES == 2 ? p.W / 2 : p.W);
  const int Ho = RES == 1 ? p.H * 2 : (RES == 2 ? p.H / 2 : p.H);
  const int HWo = Ho * Wo;
  const int vpp = C / 8;
  const int rpi = blockDim.x / vpp;
  const int c = (tid % vpp) * 8;
  const int pr = tid / vpp;
  static_assert(RES != 2 || U == 1, "the 2x2-mean variant handles one output pixel per thread");
  constexpr int LD = RES == 2 ? 8 : 2;  // float4 loads per output pixel
  const float* src;
  int Cx;
  if (c < p.C0) { src = p.src0 + c; Cx = p.C0; }
  else          { src = p.src1 + (c - p.C0); Cx = p.C1; }
  src += static_cast<size_t>(b) * HW * Cx;
  const int px0 = blockIdx.x * (U * rpi) + pr;

  float4 v[U][LD];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int px = px0 + u * rpi;
    if (px < HWo) {
      if constexpr (RES == 2) {
        const int ho = px / Wo, wo = px - ho * Wo;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4* s4 = reinterpret_cast<const float4*>(
              src + (static_cast<size_t>(2 * ho + (d >> 1)) * p.W + (2 * wo + (d & 1))) * Cx);
          v[u][2 * d] = __ldg(s4);
          v[u][2 * d + 1] = __ldg(s4 + 1);
        }
      } else if constexpr (SRC16) {
        // bf16 source (a conv output stored in bf16): one 16-byte load per 8 channels
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(p.src0h + (static_cast<size_t>(b) * HW + px) * Cx + c));
        unpack_bf16x2(h.x, v[u][0].x, v[u][0].y); unpack_bf16x2(h.y, v[u][0].z, v[u][0].w);
        unpack_bf16x2(h.z, v[u][1].x, v[u][1].y); unpack_bf16x2(h.w, v[u][1].z, v[u][1].w);
      } else {
        int pin = px;
        if constexpr (RES == 1) {
          const int ho = px / Wo, wo = px - ho * Wo;
          pin = (ho >> 1) * p.W + (wo >> 1);
        }
        const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<size_t>(pin) * Cx);
        v[u][0] = __ldg(s4);
        v[u][1] = __ldg(s4 + 1);
      }
    }
  }
  // per-channel scale / shift of this sample: 4 x 16-byte loads per thread straight from L1/L2 (all CTAs resident
  // on an SM work on the same few samples); issued right behind the data loads, no barrier in the kernel
  float a8[8], b8[8];
  if (ss != nullptr) {
    const float4* qa = reinterpret_cast<const float4*>(ss + static_cast<size_t>(b) * 2 * C + c);
    const float4* qb = reinterpret_cast<const float4*>(ss + (static_cast<size_t>(b) * 2 + 1) * C + c);
    const float4 a0 = __ldg(qa), a1 = __ldg(qa + 1), b0 = __ldg(qb), b1 = __ldg(qb + 1);
    a8[0] = a0.x; a8[1] = a0.y; a8[2] = a0.z; a8[3] = a0.w; a8[4] = a1.x; a8[5] = a1.y; a8[6] = a1.z; a8[7] = a1.w;
    b8[0] = b0.x; b8[1] = b0.y; b8[2] = b0.z; b8[3] = b0.w; b8[4] = b1.x; b8[5] = b1.y; b8[6] = b1.z; b8[7] = b1.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { a8[j] = 1.f; b8[j] = 0.f; }
  }
  const bool act = p.silu != 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int px = px0 + u * rpi;
    if (px < HWo) {
      float y[8], r[8];
      if constexpr (RES == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { y[j] = 0.f; r[j] = 0.f; }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float q[8] = {v[u][2 * d].x, v[u][2 * d].y, v[u][2 * d].z, v[u][2 * d].w,
                              v[u][2 * d + 1].x, v[u][2 * d + 1].y, v[u][2 * d + 1].z, v[u][2 * d + 1].w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = q[j] * a8[j] + b8[j];
            y[j] += act ? silu_f(t) : t;
            r[j] += q[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { y[j] *= 0.25f; r[j] *= 0.25f; }
      } else {
        const float q[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          r[j] = q[j];
          const float t = q[j] * a8[j] + b8[j];
          y[j] = act ? silu_f(t) : t;
        }
      }
      const size_t o = (static_cast<size_t>(b) * HWo + px) * C + c;
      uint4 pk;
      pk.x = pack_bf16x2(y[0], y[1]); pk.y = pack_bf16x2(y[2], y[3]);
      pk.z = pack_bf16x2(y[4], y[5]); pk.w = pack_bf16x2(y[6], y[7]);
      *reinterpret_cast<uint4*>(p.out + o) = pk;
      if (p.raw) {
        uint4 pq;
        pq.x = pack_bf16x2(r[0], r[1]); pq.y = pack_bf16x2(r[2], r[3]);
        pq.z = pack_bf16x2(r[4], r[5]); pq.w = pack_bf16x2(r[6], r[7]);
        *reinterpret_cast<uint4*>(p.raw + o) = pq;
      }
      if (p.raw_f32) {
        float4* d = reinterpret_cast<float4*>(p.raw_f32 + o);
        d[0] = make_float4(r[0], r[1], r[2], r[3]);
        d[1] = make_float4(r[4], r[5], r[6], r[7]);
      }
    }
  }
}

int launch_gn_apply(const GnParams& p, const float* ss, int num_sms, cudaStream_t s) {
  (void)num_sms;
  const int C = p.C0 + p.C1;
  const int vpp = C / 8;
  if (vpp > 256) return static_cast<int>(cudaErrorInvalidValue);
  const int threads = vpp * (256 / vpp);  // every thread keeps one fixed 8-channel vector
  const int rpi = threads / vpp;
  const int Ho = p.resample == 1 ? p.H * 2 : (p.resample == 2 ? p.H / 2 : p.H);
  const int Wo = p.resample == 1 ? p.W * 2 : (p.resample == 2 ? p.W / 2 : p.W);
  const int HWo = Ho * Wo;
  // pixels per thread: 4 on the large un-resampled tensors (more 16-byte loads in flight per thread; measured per DDPM++
  // evaluation at B=512 on one box: 7.59 ms with 2, 6.70-6.81 ms with 4, 7.62 ms with 8; DP_GN_U=2 restores 2),
  // 2 otherwise, 1 for the 2x2-mean variant (8 loads per output pixel already)
  static const int u_big = [] { const char* v = std::getenv("DP_GN_U"); return v ? std::atoi(v) : 4; }();
  const int U = p.resample == 2 ? 1 : ((p.resample == 0 && HWo >= 256 && u_big >= 4) ? 4 : 2);
  const dim3 grid((HWo + U * rpi - 1) / (U * rpi), p.B);
  const size_t smem = 0;
  if (p.src0h != nullptr) {
    if (p.resample != 0 || p.C1 != 0) return static_cast<int>(cudaErrorInvalidValue);
    if (U == 4) (void)launch_k(gn_apply_kernel<0, true, 4>, dim3(grid), dim3(threads), smem, s, 1, p, ss);
    else (void)launch_k(gn_apply_kernel<0, true, 2>, dim3(grid), dim3(threads), smem, s, 1, p, ss);
  } else if (p.resample == 0) {
    if (U == 4) (void)launch_k(gn_apply_kernel<0, false, 4>, dim3(grid), dim3(threads), smem, s, 1, p, ss);
    else (void)launch_k(gn_apply_kernel<0, false, 2>, dim3(grid), dim3(threads), smem, s, 1, p, ss);
  } else if (p.resample == 1) (void)launch_k(gn_apply_kernel<1, false, 2>, dim3(grid), dim3(threads), smem, s, 1, p, ss);
  else (void)launch_k(gn_apply_kernel<2, false, 1>, dim3(grid), dim3(threads), smem, s, 1, p, ss);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------------
__global__ void stats_kernel(const float* __restrict__ src, float* __restrict__ stats, int HW, int C, int P) {
  pdl_entry();
  const int pp = blockIdx.x, b = blockIdx.y;
  const int r0 = pp * 128;
  const int r1 = min(HW, r0 + 128);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    const float* x = src + (static_cast<size_t>(b) * HW) * C + c;
    for (int r = r0; r < r1; ++r) {
      const float v = x[static_cast<size_t>(r) * C];
      s += v;
      q += v * v;
    }
    float* o = stats + ((static_cast<size_t>(b) * P + pp) * C + c) * 2;
    o[0] = s;
    o[1] = q;
  }
}

int launch_stats(const float* src, float* stats, int B, int HW, int C, cudaStream_t s) {
  const int P = (HW + 127) / 128;
  (void)launch_k(stats_kernel, dim3(P, B), dim3(256), 0, s, 1, src, stats, HW, C, P);
  return static_cast<int>(cudaGetLastError());
}

__global__ void stats_reduce_kernel(const float* __restrict__ in, float* __restrict__ out, int P, int C2) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= C2) return;
  float s = 0.f;
  for (int pp = 0; pp < P; ++pp) s += in[(static_cast<size_t>(b) * P + pp) * C2 + i];
  out[static_cast<size_t>(b) * C2 + i] = s;
}

int launch_stats_reduce(const float* in, float* out, int B, int P, int C, cudaStream_t s) {
  const int C2 = 2 * C;
  (void)launch_k(stats_reduce_kernel, dim3((C2 + 255) / 256, B), dim3(256), 0, s, 1, in, out, P, C2);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// input conv 3 -> Cout (fp32 SIMT: K = 27 is too thin for the tensor pipe)
// ------------------------------------------------------------------------------------------------
// Each thread computes 4 consecutive pixels (along W) x 4 output channels: the 3 x 6 x 3 input patch lives in
// registers and every 16-byte weight load from shared memory feeds 16 FMAs.
__global__ void __launch_bounds__(256) conv_in_kernel(ConvInParams p) {
  pdl_entry();
  extern __shared__ float sw[];  // [27][Cout] + bias[Cout]
  const int Cout = p.Cout;
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[i] = p.w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[27 * Cout + i] = p.bias[i];
  __syncthreads();
  const int vpp = Cout / 4;
  const int gpb = blockDim.x / vpp;  // 4-pixel groups per block
  const int HW = p.H * p.W;
  const int wq = p.W / 4;
  const long long ngroups = static_cast<long long>(p.B) * p.H * wq;
  const long long gg = static_cast<long long>(blockIdx.x) * gpb + threadIdx.x / vpp;
  if (gg >= ngroups) return;
  const int co = (threadIdx.x % vpp) * 4;
  const int b = static_cast<int>(gg / (p.H * wq));
  const int rem = static_cast<int>(gg - static_cast<long long>(b) * p.H * wq);
  const int h = rem / wq, w0 = (rem - h * wq) * 4;
  float xin[3][6][3];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int hh = h + ky - 1;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int ww = w0 + j - 1;
      const bool ok = hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
      const float* src = p.x + (static_cast<size_t>(b) * HW + static_cast<size_t>(ok ? hh : 0) * p.W + (ok ? ww : 0)) * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) xin[ky][j][ci] = ok ? __ldg(src + ci) : 0.f;
    }
  }
  const float4 bias4 = *reinterpret_cast<const float4*>(sw + 27 * Cout + co);
  float4 acc[4] = {bias4, bias4, bias4, bias4};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float4 wv = *reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * 3 + ci) * Cout + co);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float xv = xin[ky][px + kx][ci];
          acc[px].x += xv * wv.x; acc[px].y += xv * wv.y; acc[px].z += xv * wv.z; acc[px].w += xv * wv.w;
        }
      }
  float* out = p.out + (static_cast<size_t>(b) * HW + static_cast<size_t>(h) * p.W + w0) * Cout + co;
#pragma unroll
  for (int px = 0; px < 4; ++px) *reinterpret_cast<float4*>(out + static_cast<size_t>(px) * Cout) = acc[px];
}

int launch_conv_in(const ConvInParams& p, cudaStream_t s) {
  if (p.W % 4) return static_cast<int>(cudaErrorInvalidValue);
  const int vpp = p.Cout / 4;
  const int gpb = 256 / vpp;
  const long long ngroups = static_cast<long long>(p.B) * p.H * (p.W / 4);
  const size_t smem = static_cast<size_t>(28 * p.Cout) * sizeof(float);
  (void)launch_k(conv_in_kernel, dim3(static_cast<unsigned>((ngroups + gpb - 1) / gpb)), dim3(256), smem, s, 1, p);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// per-step update from the output conv's result (fp32 [B*HW, ld], first Cout columns valid)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) update_kernel(UpdateParams p) {
  pdl_entry();
  const long long gp = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int HW = p.H * p.W;
  const long long npix = static_cast<long long>(p.B) * HW;
  if (gp >= npix) return;
  const int b = static_cast<int>(gp / HW);
  const int pix = static_cast<int>(gp - static_cast<long long>(b) * HW);
  const float* e = p.eps + gp * p.ld;
  float o[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) o[j] = j < p.Cout ? e[j] : 0.f;
  if (p.mode == 0) {
    for (int j = 0; j < p.Cout; ++j) p.out_nchw[(static_cast<size_t>(b) * p.Cout + j) * HW + pix] = o[j];
    return;
  }
  const int step = *p.tables.step;
  const CallParams cp = *p.call;
  float k[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) k[i] = p.tables.coef[step * 8 + i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float* xp = p.x + gp * 3 + c;
    const float xv = *xp;
    const float z = cp.step_noise ? cp.step_noise[((static_cast<size_t>(step) * p.B + b) * 3 + c) * HW + pix]
                                  : dp_normal(cp.seed, cp.sample_offset + b, static_cast<unsigned>(step) + 1u,
                                              static_cast<unsigned>(pix), c);
    float xn;
    if (cp.update_kind == 0) {
      xn = k[0] * xv + k[1] * o[c] + k[2] * z;
    } else if (cp.update_kind == 2) {
      // Langevin-dynamics SDE (runners/diffpure_ldsde.py:92-131): extra pull towards the initial image
      xn = k[0] * xv + k[1] * o[c] + k[2] * z + k[3] * p.x_init[gp * 3 + c];
    } else {
      // guided_diffusion/gaussian_diffusion.py:277-284,305,317-322,438-446
      float x0 = k[0] * xv - k[1] * o[c];
      x0 = fminf(1.f, fmaxf(-1.f, x0));
      const float mean = k[2] * x0 + k[3] * xv;
      const float frac = (o[3 + c] + 1.f) * 0.5f;
      const float logvar = frac * k[4] + (1.f - frac) * k[5];
      xn = mean + k[6] * expf(0.5f * logvar) * z;
    }
    *xp = xn;
    if (cp.states) cp.states[((static_cast<size_t>(step + 1) * p.B + b) * 3 + c) * HW + pix] = xn;
  }
}

int launch_update(const UpdateParams& p, cudaStream_t s) {
  const long long npix = static_cast<long long>(p.B) * p.H * p.W;
  (void)launch_k(update_kernel, dim3(static_cast<unsigned>((npix + 255) / 256)), dim3(256), 0, s, 1, p);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// short-sequence attention (T <= 64): one CTA per (head, sample), everything resident in smem
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_small_kernel(AttnSmallParams p) {
  pdl_entry();
  extern __shared__ __align__(16) unsigned char smraw[];
  const int T = p.T, d = p.d;
  const int pitch = d + 2;  // bf16 elements; odd word pitch -> conflict-free row-strided reads
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smraw);
  __nv_bfloat16* sk = sq + T * pitch;
  __nv_bfloat16* sv = sk + T * pitch;
  float* ss = reinterpret_cast<float*>(sv + T * pitch + ((T * pitch * 3) & 1));
  const int head = blockIdx.x, b = blockIdx.y;
  const int ld = 3 * p.heads * d;
  const __nv_bfloat16* g = p.qkv + static_cast<size_t>(b) * T * ld + head * d;
  for (int i = threadIdx.x; i < T * d; i += blockDim.x) {
    const int r = i / d, c = i - r * d;
    sq[r * pitch + c] = g[static_cast<size_t>(r) * ld + c];
    sk[r * pitch + c] = g[static_cast<size_t>(r) * ld + p.heads * d + c];
    sv[r * pitch + c] = g[static_cast<size_t>(r) * ld + 2 * p.heads * d + c];
  }
  __syncthreads();
  for (int ij = threadIdx.x; ij < T * T; ij += blockDim.x) {
    const int i = ij / T, j = ij - i * T;
    float acc = 0.f;
    for (int c = 0; c < d; c += 2) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(sq + i * pitch + c);
      const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(sk + j * pitch + c);
      acc += __bfloat162float(a.x) * __bfloat162float(bb.x) + __bfloat162float(a.y) * __bfloat162float(bb.y);
    }
    ss[i * (T + 1) + j] = acc * p.scale;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < T; i += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, ss[i * (T + 1) + j]);
    for (int m = 16; m > 0; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = __expf(ss[i * (T + 1) + j] - mx);
      ss[i * (T + 1) + j] = e;
      sum += e;
    }
    for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    const float inv = 1.0f / sum;
    for (int j = lane; j < T; j += 32) ss[i * (T + 1) + j] *= inv;
  }
  __syncthreads();
  __nv_bfloat16* o = p.out + static_cast<size_t>(b) * T * (p.heads * d) + head * d;
  for (int ic = threadIdx.x; ic < T * d; ic += blockDim.x) {
    const int i = ic / d, c = ic - i * d;
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc += ss[i * (T + 1) + j] * __bfloat162float(sv[j * pitch + c]);
    o[static_cast<size_t>(i) * (p.heads * d) + c] = __float2bfloat16_rn(acc);
  }
}

int launch_attn_small(const AttnSmallParams& p, cudaStream_t s) {
  const int pitch = p.d + 2;
  const size_t smem = static_cast<size_t>(3) * p.T * pitch * 2 + 4 + static_cast<size_t>(p.T) * (p.T + 1) * 4;
  if (smem > 227 * 1024) return static_cast<int>(cudaErrorInvalidValue);
  cudaError_t e = cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem > 48 * 1024 ? smem : 48 * 1024));
  if (e != cudaSuccess) return static_cast<int>(e);
  (void)launch_k(attn_small_kernel, dim3(p.heads, p.B), dim3(256), smem, s, 1, p);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// row softmax for long sequences (T > 256): one warp per row, fp32 logits -> normalised bf16
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ src,
                                                           __nv_bfloat16* __restrict__ out, long long rows, int T) {
  pdl_entry();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* s4 = reinterpret_cast<const float4*>(src + row * T);
  const int n4 = T / 4;
  float mx = -INFINITY;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = __ldg(s4 + i);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int m = 16; m > 0; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
  float sum = 0.f;
  for (int i = lane; i < n4; i += 32) {
    const float4 v = __ldg(s4 + i);
    sum += __expf(v.x - mx) + __expf(v.y - mx) + __expf(v.z - mx) + __expf(v.w - mx);
  }
  for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
  const float inv = 1.0f / sum;
  uint2* o2 = reinterpret_cast<uint2*>(out + row * T);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = __ldg(s4 + i);
    uint2 pk;
    pk.x = pack_bf16x2(__expf(v.x - mx) * inv, __expf(v.y - mx) * inv);
    pk.y = pack_bf16x2(__expf(v.z - mx) * inv, __expf(v.w - mx) * inv);
    o2[i] = pk;
  }
}

int launch_softmax_rows(const float* src, __nv_bfloat16* out, long long rows, int T, cudaStream_t s) {
  if (T % 4) return static_cast<int>(cudaErrorInvalidValue);
  (void)launch_k(softmax_rows_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, s, 1, src, out, rows, T);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// state init / layout conversion / step counter
// ------------------------------------------------------------------------------------------------
__global__ void init_state_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                  float* __restrict__ x, int B, int C, int HW, float sx, float se,
                                  unsigned long long seed, unsigned long long sample_offset) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * C * HW;
  if (i >= n) return;
  const int pix = static_cast<int>(i % HW);
  const int c = static_cast<int>((i / HW) % C);
  const int b = static_cast<int>(i / (static_cast<long long>(HW) * C));
  const float e = noise ? noise[i] : dp_normal(seed, sample_offset + b, 0u, static_cast<unsigned>(pix), c);
  x[(static_cast<size_t>(b) * HW + pix) * C + c] = sx * x0[i] + se * e;
}

int launch_init_state(const float* x0_nchw, const float* noise_nchw, float* x_nhwc, int B, int C, int HW,
                      float sx, float se, unsigned long long seed, unsigned long long sample_offset,
                      cudaStream_t s) {
  const long long n = static_cast<long long>(B) * C * HW;
  init_state_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x0_nchw, noise_nchw, x_nhwc, B, C, HW,
                                                                          sx, se, seed, sample_offset);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// fused pre / post steps either side of the loop (eval_sde_adv.py:73-89: [0,1] -> [-1,1], ImageNet 224 <-> 256 bilinear
// resize; utils.py:144-153: classifier normalisation). Bilinear = F.interpolate(mode='bilinear', align_corners=False):
// src = scale (dst + 0.5) - 0.5 clamped at 0, scale = in / out, in fp32 and in ATen's association order.
// ------------------------------------------------------------------------------------------------
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_of(int dst, int in_size, int out_size) {
  Lerp r;
  if (in_size == out_size) { r.i0 = r.i1 = dst; r.l0 = 1.f; r.l1 = 0.f; return r; }
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  r.i0 = static_cast<int>(src);
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - static_cast<float>(r.i0);
  r.l0 = 1.f - r.l1;
  return r;
}

// x_state[b,h,w,c] = sx * map(resize(x0))[b,c,h,w] + se * noise     (x0: NCHW [B,C,Hin,Win])
__global__ void init_state_pre_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                      float* __restrict__ x, int B, int C, int H, int W, int Hin, int Win, int unit_range,
                                      float sx, float se, unsigned long long seed, unsigned long long sample_offset) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int HW = H * W;
  const long long n = static_cast<long long>(B) * C * HW;
  if (i >= n) return;
  const int pix = static_cast<int>(i % HW);
  const int c = static_cast<int>((i / HW) % C);
  const int b = static_cast<int>(i / (static_cast<long long>(HW) * C));
  const int h = pix / W, w = pix - h * W;
  const Lerp ly = lerp_of(h, Hin, H), lx = lerp_of(w, Win, W);
  const float* src = x0 + (static_cast<size_t>(b) * C + c) * Hin * Win;
  float v = ly.l0 * (lx.l0 * src[ly.i0 * Win + lx.i0] + lx.l1 * src[ly.i0 * Win + lx.i1]) +
            ly.l1 * (lx.l0 * src[ly.i1 * Win + lx.i0] + lx.l1 * src[ly.i1 * Win + lx.i1]);
  if (unit_range) v = (v - 0.5f) * 2.f;
  const float e = noise ? noise[i] : dp_normal(seed, sample_offset + b, 0u, static_cast<unsigned>(pix), c);
  x[(static_cast<size_t>(b) * HW + pix) * C + c] = sx * v + se * e;
}

int launch_init_state_pre(const float* x0_nchw, const float* noise_nchw, float* x_nhwc, int B, int C, int H, int W,
                          int Hin, int Win, int unit_range, float sx, float se, unsigned long long seed,
                          unsigned long long sample_offset, cudaStream_t s) {
  const long long n = static_cast<long long>(B) * C * H * W;
  init_state_pre_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x0_nchw, noise_nchw, x_nhwc, B, C, H, W,
                                                                              Hin, Win, unit_range, sx, se, seed,
                                                                              sample_offset);
  return static_cast<int>(cudaGetLastError());
}

// out[b,c,ho,wo] = ((resize(x)[b,c,ho,wo] + 1) / 2 - mean[c]) / std[c]      (x: NHWC [B,H,W,C], out: NCHW)
__global__ void final_post_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int H, int W,
                                  int Ho, int Wo, PostParams pp) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * C * Ho * Wo;
  if (i >= n) return;
  const int wo = static_cast<int>(i % Wo);
  const int ho = static_cast<int>((i / Wo) % Ho);
  const int c = static_cast<int>((i / (static_cast<long long>(Wo) * Ho)) % C);
  const int b = static_cast<int>(i / (static_cast<long long>(Wo) * Ho * C));
  const Lerp ly = lerp_of(ho, H, Ho), lx = lerp_of(wo, W, Wo);
  const float* src = x + static_cast<size_t>(b) * H * W * C + c;
  float v = ly.l0 * (lx.l0 * src[(ly.i0 * W + lx.i0) * C] + lx.l1 * src[(ly.i0 * W + lx.i1) * C]) +
            ly.l1 * (lx.l0 * src[(ly.i1 * W + lx.i0) * C] + lx.l1 * src[(ly.i1 * W + lx.i1) * C]);
  if (pp.unit_range) v = (v + 1.f) * 0.5f;
  if (pp.std[0] != 0.f) v = (v - pp.mean[c]) / pp.std[c];
  out[i] = v;
}

int launch_final_post(const float* x_nhwc, float* out_nchw, int B, int C, int H, int W, int Ho, int Wo,
                      const PostParams& pp, cudaStream_t s) {
  const long long n = static_cast<long long>(B) * C * Ho * Wo;
  final_post_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x_nhwc, out_nchw, B, C, H, W, Ho, Wo, pp);
  return static_cast<int>(cudaGetLastError());
}

// state x (fp32 NHWC, 3 channels) -> bf16 NHWC zero-padded to Cpad channels: the A operand of the input conv on the tensor
// cores (K = 9 * 64 of which 27 columns are non-zero: 21x redundant MMAs that still beat the SIMT conv 6x at 256x256)
__global__ void __launch_bounds__(256) pad_in_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                     long long npix, int Cpad) {
  pdl_entry();
  const int vpp = Cpad / 8;                                   // 16-byte vectors per pixel
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix * vpp) return;
  const long long pix = i / vpp;
  const int v = static_cast<int>(i - pix * vpp);
  uint4 pk = make_uint4(0u, 0u, 0u, 0u);
  if (v == 0) {
    const float a = x[pix * 3], b = x[pix * 3 + 1], c = x[pix * 3 + 2];
    pk.x = pack_bf16x2(a, b);
    pk.y = pack_bf16x2(c, 0.f);
  }
  *reinterpret_cast<uint4*>(out + pix * Cpad + v * 8) = pk;
}

int launch_pad_in(const float* x_nhwc3, __nv_bfloat16* out, long long npix, int Cpad, cudaStream_t s) {
  if (Cpad % 8) return static_cast<int>(cudaErrorInvalidValue);
  const long long n = npix * (Cpad / 8);
  (void)launch_k(pad_in_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, s, 1, x_nhwc3, out, npix, Cpad);
  return static_cast<int>(cudaGetLastError());
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int HW) {
  pdl_entry();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(B) * C * HW;
  if (i >= n) return;
  const int pix = static_cast<int>(i % HW);
  const int c = static_cast<int>((i / HW) % C);
  const int b = static_cast<int>(i / (static_cast<long long>(HW) * C));
  out[i] = x[(static_cast<size_t>(b) * HW + pix) * C + c];
}

int launch_nhwc_to_nchw(const float* x_nhwc, float* out_nchw, int B, int C, int HW, cudaStream_t s) {
  const long long n = static_cast<long long>(B) * C * HW;
  nhwc_to_nchw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(x_nhwc, out_nchw, B, C, HW);
  return static_cast<int>(cudaGetLastError());
}

__global__ void step_advance_kernel(int* step) {
  pdl_entry(); *step += 1; }

int launch_step_advance(int* step, cudaStream_t s) {
  (void)launch_k(step_advance_kernel, dim3(1), dim3(1), 0, s, 1, step);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace dp
