// selftest_attn.cu -- standalone on-GPU check of the fused attention-block kernel (diffpure_b200/csrc/dp_attn.cu) against a
// plain host loop that rounds to bf16 at the same points (q, k, v^T, P, o). Built by diffpure_b200/csrc/Makefile, run by
// tests/test_gpu_parity.py::test_attn_block_kernel_selftest (-m gpu). Exit code 0 = within tolerance.
//   selftest_attn            correctness at B = 1, 3, 150 (more samples than CTA pairs: the persistent loop wraps)
//   selftest_attn perf [B]   device time per launch at B (default 512)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../diffpure_b200/csrc/dp_attn.cuh"
#include "../diffpure_b200/csrc/dp_tmap.h"

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                            \
    }                                                                                     \
  } while (0)

static std::mt19937 rng(4321);
static float frand(float s = 1.f) {
  std::uniform_real_distribution<float> d(-s, s);
  return d(rng);
}
static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int T = 256, C = 256;

struct Data {
  int B;
  std::vector<float> h, w, bias, x;  // h, w already bf16-representable
  __nv_bfloat16 *d_h = nullptr, *d_w = nullptr;
  float *d_bias = nullptr, *d_x = nullptr, *d_out = nullptr, *d_stats = nullptr;
  dp::AttnBlockParams p;
};

static void make(Data& d, int B, float wscale) {
  d.B = B;
  d.h.resize(static_cast<size_t>(B) * T * C);
  d.w.resize(4 * C * C);
  d.bias.resize(4 * C);
  d.x.resize(d.h.size());
  for (auto& v : d.h) v = bf(frand(1.7f));
  for (auto& v : d.w) v = bf(frand(wscale));
  for (auto& v : d.bias) v = frand(0.2f);
  for (auto& v : d.x) v = frand(2.f);
  std::vector<__nv_bfloat16> hb(d.h.size()), wb(d.w.size());
  for (size_t i = 0; i < hb.size(); ++i) hb[i] = __float2bfloat16_rn(d.h[i]);
  for (size_t i = 0; i < wb.size(); ++i) wb[i] = __float2bfloat16_rn(d.w[i]);
  CK(cudaMalloc(&d.d_h, hb.size() * 2));
  CK(cudaMalloc(&d.d_w, wb.size() * 2));
  CK(cudaMalloc(&d.d_bias, d.bias.size() * 4));
  CK(cudaMalloc(&d.d_x, d.x.size() * 4));
  CK(cudaMalloc(&d.d_out, d.x.size() * 4));
  CK(cudaMalloc(&d.d_stats, static_cast<size_t>(B) * 2 * C * 2 * 4));
  CK(cudaMemcpy(d.d_h, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_w, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_bias, d.bias.data(), d.bias.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.d_x, d.x.data(), d.x.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d.d_out, 0xff, d.x.size() * 4));
  CK(cudaMemset(d.d_stats, 0xff, static_cast<size_t>(B) * 2 * C * 2 * 4));
  std::string err;
  if (dp::make_mat_tmap(&d.p.tmap_h, d.d_h, C, static_cast<long long>(B) * T, C, 128, &err) ||
      dp::make_mat_tmap(&d.p.tmap_w, d.d_w, C, 4 * C, C, 128, &err)) {
    printf("tensor map: %s\n", err.c_str());
    exit(2);
  }
  d.p.bias = d.d_bias;
  d.p.resid = d.d_x;
  d.p.out_f32 = d.d_out;
  d.p.stats = d.d_stats;
  d.p.B = B;
  d.p.scale = 1.0f / 16.0f;
  d.p.alpha = 0.70710678f;
  d.p.x_prefetch = getenv("DP_ATTN_XPF") ? atoi(getenv("DP_ATTN_XPF")) : 2;
}

static void release(Data& d) {
  cudaFree(d.d_h); cudaFree(d.d_w); cudaFree(d.d_bias); cudaFree(d.d_x); cudaFree(d.d_out); cudaFree(d.d_stats);
}

// y[t][o] = sum_c a[t][c] w[o][c]
static void matmul_nt(const float* a, const float* w, float* y, int M, int N, int K) {
  for (int i = 0; i < M; ++i)
    for (int o = 0; o < N; ++o) {
      float acc = 0.f;
      const float* ar = a + static_cast<size_t>(i) * K;
      const float* wr = w + static_cast<size_t>(o) * K;
      for (int c = 0; c < K; ++c) acc += ar[c] * wr[c];
      y[static_cast<size_t>(i) * N + o] = acc;
    }
}

static void reference(const Data& d, int s, std::vector<float>& out) {
  const float* h = d.h.data() + static_cast<size_t>(s) * T * C;
  const float* x = d.x.data() + static_cast<size_t>(s) * T * C;
  std::vector<float> q(T * C), k(T * C), v(T * C), sc(T * T), pm(T * T), o(T * C), vt(C * T);
  matmul_nt(h, d.w.data(), q.data(), T, C, C);
  matmul_nt(h, d.w.data() + C * C, k.data(), T, C, C);
  matmul_nt(h, d.w.data() + 2 * C * C, v.data(), T, C, C);
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < C; ++c) {
      q[t * C + c] = bf((q[t * C + c] + d.bias[c]) * d.p.scale);
      k[t * C + c] = bf(k[t * C + c] + d.bias[C + c]);
      vt[c * T + t] = bf(v[t * C + c] + d.bias[2 * C + c]);
    }
  matmul_nt(q.data(), k.data(), sc.data(), T, T, C);
  std::vector<float> rsum(T);
  for (int t = 0; t < T; ++t) {
    float mx = -INFINITY;
    for (int j = 0; j < T; ++j) mx = fmaxf(mx, sc[t * T + j]);
    float sum = 0.f;
    for (int j = 0; j < T; ++j) {
      pm[t * T + j] = bf(expf(sc[t * T + j] - mx));
      sum += pm[t * T + j];
    }
    rsum[t] = sum;
  }
  matmul_nt(pm.data(), vt.data(), o.data(), T, C, T);
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < C; ++c) o[t * C + c] = bf(o[t * C + c] / rsum[t]);
  out.resize(T * C);
  matmul_nt(o.data(), d.w.data() + 3 * C * C, out.data(), T, C, C);
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < C; ++c) out[t * C + c] = (out[t * C + c] + d.bias[3 * C + c] + x[t * C + c]) * d.p.alpha;
}

static int check(int B, int num_sms) {
  Data d;
  make(d, B, 0.11f);
  int rc = dp::launch_attn_block(d.p, num_sms, nullptr);
  if (rc) {
    printf("launch failed: %s\n", cudaGetErrorString(static_cast<cudaError_t>(rc)));
    return 1;
  }
  CK(cudaDeviceSynchronize());
  std::vector<float> out(d.x.size()), stats(static_cast<size_t>(B) * 2 * C * 2);
  CK(cudaMemcpy(out.data(), d.d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(stats.data(), d.d_stats, stats.size() * 4, cudaMemcpyDeviceToHost));
  int fails = 0;
  // samples checked on the host: the first, the last, and one from the wrap-around region
  std::vector<int> samples = {0};
  if (B > 1) samples.push_back(B - 1);
  if (B > 80) samples.push_back(77);
  for (int s : samples) {
    std::vector<float> ref;
    reference(d, s, ref);
    double num = 0, den = 0, mx = 0;
    for (int i = 0; i < T * C; ++i) {
      const double e = out[static_cast<size_t>(s) * T * C + i] - ref[i];
      num += e * e;
      den += static_cast<double>(ref[i]) * ref[i];
      mx = fmax(mx, fabs(e));
    }
    const double rel = sqrt(num / den);
    // the x + h sum is dominated by x (exact): judge the attention part on its own scale as well
    double hn = 0, he = 0;
    for (int i = 0; i < T * C; ++i) {
      const double hr = ref[i] / d.p.alpha - d.x[static_cast<size_t>(s) * T * C + i];
      const double hg = out[static_cast<size_t>(s) * T * C + i] / d.p.alpha - d.x[static_cast<size_t>(s) * T * C + i];
      hn += hr * hr;
      he += (hg - hr) * (hg - hr);
    }
    const double relh = sqrt(he / hn);
    const bool ok = rel < 2e-3 && relh < 1e-2 && std::isfinite(rel);
    printf("attn_block B=%d sample %d: rel-L2 %.3e (attention branch alone %.3e) max abs %.3e  %s\n", B, s, rel, relh, mx,
           ok ? "OK" : "FAIL");
    fails += !ok;
    // partial statistics of the two 128-row tiles of this sample, against the kernel's own output
    double smax = 0;
    for (int half = 0; half < 2; ++half)
      for (int c = 0; c < C; ++c) {
        double sm = 0, sq = 0;
        for (int t = 0; t < 128; ++t) {
          const double v = out[(static_cast<size_t>(s) * T + half * 128 + t) * C + c];
          sm += v;
          sq += v * v;
        }
        const float* g = stats.data() + ((static_cast<size_t>(s) * 2 + half) * C + c) * 2;
        smax = fmax(smax, fabs(g[0] - sm) / (1.0 + fabs(sm)));
        smax = fmax(smax, fabs(g[1] - sq) / (1.0 + fabs(sq)));
      }
    const bool sok = smax < 1e-4;
    printf("attn_block B=%d sample %d: statistics max rel err %.3e  %s\n", B, s, smax, sok ? "OK" : "FAIL");
    fails += !sok;
  }
  // determinism + sample independence: the same sample data at another position gives the same bits
  if (B >= 3) {
    CK(cudaMemcpy(d.d_h + static_cast<size_t>(1) * T * C, d.d_h, static_cast<size_t>(T) * C * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(d.d_x + static_cast<size_t>(1) * T * C, d.d_x, static_cast<size_t>(T) * C * 4, cudaMemcpyDeviceToDevice));
    rc = dp::launch_attn_block(d.p, num_sms, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<float> o2(2 * T * C);
    CK(cudaMemcpy(o2.data(), d.d_out, o2.size() * 4, cudaMemcpyDeviceToHost));
    const bool same = rc == 0 && memcmp(o2.data(), o2.data() + T * C, static_cast<size_t>(T) * C * 4) == 0 &&
                      memcmp(o2.data(), out.data(), static_cast<size_t>(T) * C * 4) == 0;
    printf("attn_block B=%d: sample 0 repeated at position 1 and across launches is bit-identical  %s\n", B, same ? "OK" : "FAIL");
    fails += !same;
  }
  release(d);
  return fails;
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (dp::attn_block_init()) {
    printf("attn_block_init failed\n");
    return 2;
  }
  if (argc > 1 && !strcmp(argv[1], "perf")) {
    const int B = argc > 2 ? atoi(argv[2]) : 512;
    Data d;
    make(d, B, 0.11f);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) dp::launch_attn_block(d.p, prop.multiProcessorCount, nullptr);
    CK(cudaDeviceSynchronize());
    const int iters = 20;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) dp::launch_attn_block(d.p, prop.multiProcessorCount, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / iters;
    const double flops = 6.0 * 2.0 * T * C * C * B;
    const double bytes = static_cast<double>(B) * T * C * (2 + 4 + 4);
    printf("perf attn_block B=%d: %.1f us  %.1f TF/s  %.0f GB/s (h bf16 + x fp32 in, fp32 out)\n", B, us, flops / us * 1e-6,
           bytes / us * 1e-3);
    release(d);
    return 0;
  }
  int fails = 0;
  fails += check(1, prop.multiProcessorCount);
  fails += check(3, prop.multiProcessorCount);
  fails += check(150, prop.multiProcessorCount);
  printf("selftest_attn: %d failure(s)\n", fails);
  return fails ? 1 : 0;
}
