// selftest_gemm.cu -- standalone on-GPU check of the tcgen05 implicit-GEMM kernel against a plain
// host loop (fp32 accumulation of the same bf16-rounded operands). Built by diffpure_b200/csrc/Makefile,
// run by tests/test_gpu_parity.py::test_gemm_kernel_selftest (-m gpu). Exit code 0 = all cases within tolerance.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../diffpure_b200/csrc/dp_gemm.cuh"
#include "../diffpure_b200/csrc/dp_tmap.h"

using dp::GemmParams;

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                            \
    }                                                                                     \
  } while (0)

static std::mt19937 rng(1234);
static float frand(float s = 1.f) {
  std::uniform_real_distribution<float> d(-s, s);
  return d(rng);
}
static float bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <class T>
T* dev(const std::vector<T>& h) {
  T* d;
  CK(cudaMalloc(&d, h.size() * sizeof(T) + 16));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
static std::vector<__nv_bfloat16> tobf(const std::vector<float>& v) {
  std::vector<__nv_bfloat16> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __float2bfloat16_rn(v[i]);
  return o;
}

struct ConvCase {
  const char* name;
  int B, H, W;        // output grid
  int C0, taps0;      // segment 0 (input spatial = output spatial * stride)
  int C1;             // segment 1 (1x1), 0 = none
  int N, bn;
  int stride;         // 1 or 2 (stride 2: pad right/bottom only, taps 9)
  bool bias, rowvec, resid, silu, stats, out_bf16;
  float alpha;
  int cg;             // 0/1 = one CTA per tile, 2 = CTA pair (cta_group::2)
  int gn;             // 1 = fused GroupNorm + SiLU epilogue (bf16 output only; the sample's accumulators stay in TMEM)
                      // 2 = "late" variant: fp32 result (+ residual, alpha, statistics, raw bf16 copy) AND its GroupNorm + SiLU
};

static int num_sms = 148;
static int failures = 0;

static void run_conv(const ConvCase& cs) {
  const int B = cs.B, H = cs.H, W = cs.W, N = cs.N;
  const int Hin = H * cs.stride, Win = W * cs.stride;
  const int K0 = cs.taps0 * cs.C0, Kt = K0 + cs.C1;
  const int M = B * H * W;
  std::vector<float> a0((size_t)B * Hin * Win * cs.C0), a1((size_t)M * (cs.C1 ? cs.C1 : 1)), w((size_t)N * Kt);
  for (auto& v : a0) v = bf(frand());
  for (auto& v : a1) v = bf(frand());
  const float ws = 1.0f / sqrtf((float)Kt);
  for (auto& v : w) v = bf(frand(ws * 1.7f));
  std::vector<float> bias(N), rowvec((size_t)B * N), resid((size_t)M * N);
  for (auto& v : bias) v = frand();
  for (auto& v : rowvec) v = frand();
  for (auto& v : resid) v = frand();

  auto a0b = tobf(a0);
  auto a1b = tobf(a1);
  auto wb = tobf(w);
  __nv_bfloat16* d_a0 = dev(a0b);
  __nv_bfloat16* d_a1 = dev(a1b);
  __nv_bfloat16* d_w = dev(wb);
  float* d_bias = dev(bias);
  float* d_rowvec = dev(rowvec);
  float* d_resid = dev(resid);
  float *d_out, *d_stats;
  __nv_bfloat16* d_outb;
  CK(cudaMalloc(&d_out, (size_t)M * N * 4));
  CK(cudaMalloc(&d_outb, (size_t)M * N * 2));
  CK(cudaMemset(d_out, 0xff, (size_t)M * N * 4));

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.batch = 1;
  const int cg = cs.cg == 2 ? 2 : 1;
  const int groups = N / 4 < 32 ? N / 4 : 32;
  std::vector<float> gamma(N), beta(N);
  for (auto& v : gamma) v = 1.0f + 0.3f * frand();
  for (auto& v : beta) v = 0.3f * frand();
  float* d_gamma = dev(gamma);
  float* d_beta = dev(beta);
  __nv_bfloat16* d_gn = nullptr;
  CK(cudaMalloc(&d_gn, (size_t)M * N * 2));
  if (cs.gn) {
    p.gn_out = cs.gn == 2 ? d_gn : d_outb;
    p.gn_gamma = d_gamma; p.gn_beta = d_beta;
    p.gn_cpg = N / groups; p.gn_hw = H * W; p.gn_eps = 1e-6f; p.gn_silu = 1;
    p.gn_p1 = getenv("DP_GN_P1") ? atoi(getenv("DP_GN_P1")) : 1;   // group-sum pass 1 (default) / per-channel pass 1
    const bool super_pair = strstr(cs.name, "super-pair") != nullptr;
    if (H * W == 1024 && !super_pair) { p.tpg = 4; p.acc_stages = 4; }  // the sample's four tiles resident in one pair
    if (cs.bn == 128) p.acc_stages = 4;
    if (H * W == 1024 && super_pair) {  // two CTA pairs per sample, exchange through global memory
      p.tpg = 2; p.upc = 2; p.acc_stages = 4;
      static void* xg = nullptr;
      const size_t kData = 128 * 2 * 4 * 64 * 2 * sizeof(float), kFlags = 128 * 4 * 8;
      if (!xg) { CK(cudaMalloc(&xg, kData + kFlags + 64)); CK(cudaMemset(xg, 0, kData + kFlags + 64)); }
      p.xg_data = static_cast<float*>(xg);
      p.xg_flag = reinterpret_cast<unsigned long long*>(static_cast<char*>(xg) + kData);
      p.xg_epoch = reinterpret_cast<unsigned long long*>(static_cast<char*>(xg) + kData + kFlags);
    }
    p.gn_xchg = (cg == 2 && H * W >= 256 && p.upc != 2) ? 1 : 0;
  }
  dp::gemm_fill_geometry(p, B, H, W, N, cs.bn, cg);
  const int nsegs_total = p.m_tiles * p.stat_nseg;
  CK(cudaMalloc(&d_stats, (size_t)nsegs_total * N * 2 * 4));
  CK(cudaMemset(d_stats, 0, (size_t)nsegs_total * N * 2 * 4));
  const dp::TileBox tb = dp::gemm_tile_box(H, W);
  std::string err;
  if (dp::make_act_tmap(&p.a[0].tmap, d_a0, cs.C0, cs.C0, Win, Hin, B, tb.bw, tb.bh, tb.bn, cs.stride, &err)) {
    printf("[%s] tmap a0: %s\n", cs.name, err.c_str());
    failures++;
    return;
  }
  p.a[0].taps = cs.taps0;
  p.a[0].kchunks = cs.C0 / 64;
  p.a[0].stride = cs.stride;
  p.a[0].pad = (cs.taps0 == 9 && cs.stride == 1) ? 1 : 0;
  p.nseg = 1;
  if (cs.C1) {
    if (dp::make_act_tmap(&p.a[1].tmap, d_a1, cs.C1, cs.C1, W, H, B, tb.bw, tb.bh, tb.bn, 1, &err)) {
      printf("[%s] tmap a1: %s\n", cs.name, err.c_str());
      failures++;
      return;
    }
    p.a[1].taps = 1;
    p.a[1].kchunks = cs.C1 / 64;
    p.a[1].stride = 1;
    p.a[1].pad = 0;
    p.nseg = 2;
  }
  if (dp::make_mat_tmap(&p.tmap_b, d_w, Kt, N, Kt, cs.bn / cg, &err)) {
    printf("[%s] tmap b: %s\n", cs.name, err.c_str());
    failures++;
    return;
  }
  // shared row patches for the 3x3 taps wherever the shape allows (DP_SELFTEST_PATCH=0: the tile-per-tap mainloop)
  static const bool patch_on = !(getenv("DP_SELFTEST_PATCH") && atoi(getenv("DP_SELFTEST_PATCH")) == 0);
  bool patched = false;
  if (patch_on && dp::gemm_enable_patch(p, cs.bn, cg)) {
    patched = true;
    if (dp::make_act_tmap(&p.a[0].tmap, d_a0, cs.C0, cs.C0, Win, Hin, B, tb.bw, tb.bh + 2, 1, 1, &err)) {
      printf("[%s] tmap a0 patch: %s\n", cs.name, err.c_str());
      failures++;
      return;
    }
  }
  p.bias = cs.bias ? d_bias : nullptr;
  p.rowvec = cs.rowvec ? d_rowvec : nullptr;
  p.rowvec_ld = N;
  int sh = 0;
  while ((1 << sh) < H * W) ++sh;
  p.rowvec_shift = sh;
  p.resid = cs.resid ? d_resid : nullptr;
  p.alpha = cs.alpha;
  p.silu = cs.silu;
  // pair kernels exist for the lowerings' epilogues only: a bf16 case writes bf16 alone there
  const bool f32_out = cs.gn == 2 || (!cs.gn && !((cg == 2 || !strncmp(cs.name, "bf16only", 8)) && cs.out_bf16));
  p.out_f32 = f32_out ? d_out : nullptr;
  p.out_bf16 = (cs.out_bf16 && cs.gn != 1) ? d_outb : nullptr;
  p.ldc = N;
  p.out_batch_stride = 0;
  p.stats = (cs.stats && cs.gn != 1) ? d_stats : nullptr;

  int e = dp::launch_gemm(p, cs.bn, false, num_sms, 0, cg);
  cudaError_t se = cudaDeviceSynchronize();
  if (e || se != cudaSuccess) {
    printf("[%s] launch/sync error %d / %s\n", cs.name, e, cudaGetErrorString(se));
    failures++;
    exit(3);
  }
  std::vector<float> out((size_t)M * N), stats((size_t)nsegs_total * N * 2);
  std::vector<__nv_bfloat16> outb((size_t)M * N), gnb((size_t)M * N);
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(outb.data(), d_outb, outb.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gnb.data(), d_gn, gnb.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(stats.data(), d_stats, stats.size() * 4, cudaMemcpyDeviceToHost));

  // host reference
  double maxerr = 0, maxref = 0, maxerr_b = 0;
  std::vector<double> rs((size_t)nsegs_total * N * 2, 0.0);
  std::vector<float> vall(cs.gn ? (size_t)M * N : 0);
  const int hw = H * W;
  const int seg_rows = hw >= 128 ? 128 : hw;
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t row = ((size_t)b * H + y) * W + x;
        for (int n = 0; n < N; ++n) {
          float acc = 0.f;
          const float* wr = &w[(size_t)n * Kt];
          for (int t = 0; t < cs.taps0; ++t) {
            const int ky = cs.taps0 == 9 ? t / 3 : 0, kx = cs.taps0 == 9 ? t % 3 : 0;
            const int pad = (cs.taps0 == 9 && cs.stride == 1) ? 1 : 0;
            const int yy = y * cs.stride + ky - pad, xx = x * cs.stride + kx - pad;
            if (yy < 0 || yy >= Hin || xx < 0 || xx >= Win) continue;
            const float* ar = &a0[(((size_t)b * Hin + yy) * Win + xx) * cs.C0];
            const float* wt = wr + (size_t)t * cs.C0;
            for (int c = 0; c < cs.C0; ++c) acc += ar[c] * wt[c];
          }
          if (cs.C1) {
            const float* ar = &a1[row * cs.C1];
            const float* wt = wr + K0;
            for (int c = 0; c < cs.C1; ++c) acc += ar[c] * wt[c];
          }
          float v = acc;
          if (cs.bias) v += bias[n];
          if (cs.rowvec) v += rowvec[(size_t)b * N + n];
          if (cs.silu) v = v / (1.f + expf(-v));
          if (cs.resid) v += resid[row * N + n];
          v *= cs.alpha;
          if (cs.gn) vall[row * N + n] = v;
          if (cs.gn == 1) continue;
          const double d = f32_out ? fabs((double)v - out[row * N + n]) : 0.0;
          if (d > maxerr) maxerr = d;
          if (fabs(v) > maxref) maxref = fabs(v);
          if (cs.out_bf16 && cs.gn != 2) {
            const double db = fabs((double)v - __bfloat162float(outb[row * N + n]));
            if (db > maxerr_b) maxerr_b = db;
          }
          const size_t sg = row / seg_rows;
          rs[(sg * N + n) * 2 + 0] += v;
          rs[(sg * N + n) * 2 + 1] += (double)v * v;
        }
      }
  if (cs.gn) {  // per (sample, group) statistics in double, then normalise + affine + SiLU; compare with the bf16 output
    const int cpg = N / groups;
    for (int b = 0; b < B; ++b)
      for (int g = 0; g < groups; ++g) {
        double S = 0, Q = 0;
        for (int px = 0; px < hw; ++px)
          for (int j = 0; j < cpg; ++j) {
            const double v = vall[((size_t)b * hw + px) * N + g * cpg + j];
            S += v;
            Q += v * v;
          }
        const double n = (double)cpg * hw, mean = S / n, rstd = 1.0 / sqrt(fmax(Q / n - mean * mean, 0.0) + 1e-6);
        for (int px = 0; px < hw; ++px)
          for (int j = 0; j < cpg; ++j) {
            const int n_ = g * cpg + j;
            const size_t i = ((size_t)b * hw + px) * N + n_;
            double y = (vall[i] - mean) * rstd * gamma[n_] + beta[n_];
            y = y / (1.0 + exp(-y));
            maxerr_b = fmax(maxerr_b, fabs(y - __bfloat162float(cs.gn == 2 ? gnb[i] : outb[i])));
            if (cs.gn == 2 && cs.out_bf16) maxerr_b = fmax(maxerr_b, fabs((double)vall[i] - __bfloat162float(outb[i])));
            maxref = fmax(maxref, fabs(y));
          }
      }
  }
  double maxs = 0, maxsref = 0;
  if (cs.stats && cs.gn != 1) {
    const size_t nvalid = (size_t)((M + seg_rows - 1) / seg_rows) * N * 2;
    for (size_t i = 0; i < nvalid; ++i) {
      maxs = fmax(maxs, fabs(rs[i] - stats[i]));
      maxsref = fmax(maxsref, fabs(rs[i]));
    }
  }
  const bool ok = maxerr <= 2e-3 * fmax(1.0, maxref) && (!(cs.out_bf16 || cs.gn) || maxerr_b <= 1e-2 * fmax(1.0, maxref)) &&
                  (!cs.stats || cs.gn == 1 || maxs <= 1e-3 * fmax(1.0, maxsref));
  printf("[%s]%s %s  max|err|=%.3e (max|ref|=%.3f) bf16out err=%.3e stats err=%.3e (ref %.2f)\n", cs.name,
         patched ? " (row patches)" : "", ok ? "OK  " : "FAIL", maxerr, maxref, maxerr_b, maxs, maxsref);
  if (!ok) failures++;
  cudaFree(d_a0); cudaFree(d_a1); cudaFree(d_w); cudaFree(d_bias); cudaFree(d_rowvec); cudaFree(d_resid);
  cudaFree(d_out); cudaFree(d_outb); cudaFree(d_stats); cudaFree(d_gamma); cudaFree(d_beta); cudaFree(d_gn);
}

// Batched attention-style GEMMs: S = softmax-numerator(Q K^T) then O = P V^T-layout.
static void run_attention(int Bt, int T, int C, int bn_s) {
  // qk: [Bt*T, 2C] bf16 (q | k); vt: [Bt, C, T] bf16
  std::vector<float> qk((size_t)Bt * T * 2 * C), vt((size_t)Bt * C * T);
  for (auto& v : qk) v = bf(frand(1.5f));
  for (auto& v : vt) v = bf(frand());
  auto qkb = tobf(qk);
  auto vtb = tobf(vt);
  __nv_bfloat16* d_qk = dev(qkb);
  __nv_bfloat16* d_vt = dev(vtb);
  __nv_bfloat16* d_p;
  float *d_rs, *d_o;
  CK(cudaMalloc(&d_p, (size_t)Bt * T * T * 2));
  CK(cudaMalloc(&d_rs, (size_t)Bt * T * 4));
  CK(cudaMalloc(&d_o, (size_t)Bt * T * C * 4));
  const float scale = 1.0f / sqrtf((float)C);
  std::string err;
  {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.batch = Bt;
    dp::gemm_fill_geometry(p, 1, 1, T, T, bn_s);
    // A = q rows: matrix [Bt*T, C] with pitch 2C ; as 4-D (C, rows, 1, 1)
    if (dp::make_act_tmap(&p.a[0].tmap, d_qk, C, 2 * C, Bt * T, 1, 1, 128, 1, 1, 1, &err)) { printf("tmap q: %s\n", err.c_str()); failures++; return; }
    p.a[0].taps = 1; p.a[0].kchunks = C / 64; p.a[0].stride = 1; p.a[0].pad = 0; p.nseg = 1;
    p.a_batch_rows = T;
    if (dp::make_mat_tmap(&p.tmap_b, d_qk + C, C, (long long)Bt * T, 2 * C, bn_s, &err)) { printf("tmap k: %s\n", err.c_str()); failures++; return; }
    p.b_batch_rows = T;
    p.out_bf16 = d_p; p.ldc = T; p.out_batch_stride = (long long)T * T;
    p.softmax_scale = scale; p.rowsum_out = d_rs; p.alpha = 1.f;
    int e = dp::launch_gemm(p, bn_s, true, num_sms, 0);
    cudaError_t se = cudaDeviceSynchronize();
    if (e || se != cudaSuccess) { printf("[attn S] launch/sync error %d / %s\n", e, cudaGetErrorString(se)); exit(3); }
  }
  {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.batch = Bt;
    const int bn_o = (C % 256 == 0) ? 256 : 128;
    dp::gemm_fill_geometry(p, 1, 1, T, C, bn_o);
    if (dp::make_act_tmap(&p.a[0].tmap, d_p, T, T, Bt * T, 1, 1, 128, 1, 1, 1, &err)) { printf("tmap p: %s\n", err.c_str()); failures++; return; }
    p.a[0].taps = 1; p.a[0].kchunks = T / 64; p.a[0].stride = 1; p.a[0].pad = 0; p.nseg = 1;
    p.a_batch_rows = T;
    if (dp::make_mat_tmap(&p.tmap_b, d_vt, T, (long long)Bt * C, T, bn_o, &err)) { printf("tmap vt: %s\n", err.c_str()); failures++; return; }
    p.b_batch_rows = C;
    p.out_f32 = d_o; p.ldc = C; p.out_batch_stride = (long long)T * C;
    p.rowscale = d_rs; p.alpha = 1.f;
    int e = dp::launch_gemm(p, bn_o, false, num_sms, 0);
    cudaError_t se = cudaDeviceSynchronize();
    if (e || se != cudaSuccess) { printf("[attn O] launch/sync error %d / %s\n", e, cudaGetErrorString(se)); exit(3); }
  }
  std::vector<float> o((size_t)Bt * T * C);
  CK(cudaMemcpy(o.data(), d_o, o.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  std::vector<float> s(T), pr(T);
  for (int b = 0; b < Bt; ++b)
    for (int i = 0; i < T; ++i) {
      float mx = -1e30f;
      for (int j = 0; j < T; ++j) {
        float acc = 0;
        for (int c = 0; c < C; ++c) acc += qk[((size_t)b * T + i) * 2 * C + c] * qk[((size_t)b * T + j) * 2 * C + C + c];
        s[j] = acc;
        mx = fmaxf(mx, acc);
      }
      float sum = 0;
      for (int j = 0; j < T; ++j) { pr[j] = bf(expf((s[j] - mx) * scale)); sum += pr[j]; }
      for (int c = 0; c < C; ++c) {
        float acc = 0;
        for (int j = 0; j < T; ++j) acc += pr[j] * vt[((size_t)b * C + c) * T + j];
        const float v = acc / sum;
        maxerr = fmax(maxerr, fabs((double)v - o[((size_t)b * T + i) * C + c]));
        maxref = fmax(maxref, fabs(v));
      }
    }
  const bool ok = maxerr <= 5e-3 * fmax(1.0, maxref);
  printf("[attention Bt=%d T=%d C=%d] %s max|err|=%.3e (max|ref|=%.3f)\n", Bt, T, C, ok ? "OK  " : "FAIL", maxerr, maxref);
  if (!ok) failures++;
  cudaFree(d_qk); cudaFree(d_vt); cudaFree(d_p); cudaFree(d_rs); cudaFree(d_o);
}

// perf mode: time one conv shape (no host reference): selftest_gemm perf B H W C0 taps N resid
static void run_perf(int B, int H, int W, int C0, int taps, int N, int resid, int bnforce, int cg) {
  const size_t M = (size_t)B * H * W;
  const int Kt = taps * C0;
  __nv_bfloat16 *d_a, *d_w;
  float *d_out, *d_res, *d_bias, *d_rowvec, *d_stats;
  CK(cudaMalloc(&d_a, M * C0 * 2)); CK(cudaMemset(d_a, 0, M * C0 * 2));
  CK(cudaMalloc(&d_w, (size_t)N * Kt * 2)); CK(cudaMemset(d_w, 0, (size_t)N * Kt * 2));
  CK(cudaMalloc(&d_out, M * N * 4)); CK(cudaMalloc(&d_res, M * N * 4)); CK(cudaMemset(d_res, 0, M * N * 4));
  CK(cudaMalloc(&d_bias, N * 4)); CK(cudaMemset(d_bias, 0, N * 4));
  CK(cudaMalloc(&d_rowvec, (size_t)B * N * 4)); CK(cudaMemset(d_rowvec, 0, (size_t)B * N * 4));
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.batch = 1;
  int bn = bnforce ? bnforce : ((N % 256 == 0) ? 256 : 128);
  dp::gemm_fill_geometry(p, B, H, W, N, bn, cg);
  CK(cudaMalloc(&d_stats, (size_t)p.m_tiles * p.stat_nseg * N * 8));
  const dp::TileBox tb = dp::gemm_tile_box(H, W);
  std::string err;
  if (dp::make_act_tmap(&p.a[0].tmap, d_a, C0, C0, W, H, B, tb.bw, tb.bh, tb.bn, 1, &err)) { printf("%s\n", err.c_str()); exit(1); }
  p.a[0].taps = taps; p.a[0].kchunks = C0 / 64; p.a[0].stride = 1; p.a[0].pad = taps == 9 ? 1 : 0; p.nseg = 1;
  if (dp::make_mat_tmap(&p.tmap_b, d_w, Kt, N, Kt, bn / cg, &err)) { printf("%s\n", err.c_str()); exit(1); }
  p.bias = d_bias; p.alpha = 1.f; p.out_f32 = d_out; p.ldc = N; p.stats = d_stats;
  const bool want_patch = !(getenv("DP_SELFTEST_PATCH") && atoi(getenv("DP_SELFTEST_PATCH")) == 0);
  if (getenv("DP_PERF_BF16")) { p.out_f32 = nullptr; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d_out); }  // bf16 output epilogue
  if (getenv("DP_PERF_GN")) {  // fused GroupNorm + SiLU epilogue (bf16 output only)
    p.out_f32 = nullptr; p.out_bf16 = nullptr; p.stats = nullptr;
    p.gn_out = reinterpret_cast<__nv_bfloat16*>(d_out);
    p.gn_gamma = d_bias; p.gn_beta = d_bias; p.gn_cpg = N / 32; p.gn_hw = H * W; p.gn_eps = 1e-6f; p.gn_silu = 1;
    p.gn_p1 = getenv("DP_GN_P1") ? atoi(getenv("DP_GN_P1")) : 1;
    const bool super_pair = getenv("DP_GN_UPC") && atoi(getenv("DP_GN_UPC")) == 2;
    if (H * W == 1024 && !super_pair) { p.tpg = 4; p.acc_stages = 4; }
    if (H * W == 1024 && super_pair) {
      p.tpg = 2; p.upc = 2; p.acc_stages = 4;
      void* xg = nullptr;
      const size_t kData = 128 * 2 * 4 * 64 * 2 * sizeof(float), kFlags = 128 * 4 * 8;
      CK(cudaMalloc(&xg, kData + kFlags + 64)); CK(cudaMemset(xg, 0, kData + kFlags + 64));
      p.xg_data = static_cast<float*>(xg);
      p.xg_flag = reinterpret_cast<unsigned long long*>(static_cast<char*>(xg) + kData);
      p.xg_epoch = reinterpret_cast<unsigned long long*>(static_cast<char*>(xg) + kData + kFlags);
    }
    p.gn_xchg = (cg == 2 && H * W >= 256 && p.upc != 2) ? 1 : 0;
    dp::gemm_fill_geometry(p, B, H, W, N, bn, cg);
  }
  int sh = 0; while ((1 << sh) < H * W) ++sh;
  if (resid) { p.resid = d_res; p.alpha = 0.70710678f; } else { p.rowvec = d_rowvec; p.rowvec_ld = N; p.rowvec_shift = sh; }
  bool patched = false;
  if (want_patch && dp::gemm_enable_patch(p, bn, cg)) {   // after the epilogue kind (gn_out) is known: it sizes the stages
    patched = true;
    if (dp::make_act_tmap(&p.a[0].tmap, d_a, C0, C0, W, H, B, tb.bw, tb.bh + 2, 1, 1, &err)) { printf("%s\n", err.c_str()); exit(1); }
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) {
    int e = dp::launch_gemm(p, bn, false, num_sms, 0, cg);
    if (e) { printf("launch failed: %d\n", e); exit(1); }
  }
  CK(cudaDeviceSynchronize());
  const int iters = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) dp::launch_gemm(p, bn, false, num_sms, 0, cg);
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fl = 2.0 * M * N * Kt;
  printf("perf B%d %dx%d C%d taps%d N%d bn%d cg%d resid%d%s stages=%d: %.1f us  %.1f TF/s\n", B, H, W, C0, taps, N, bn, cg, resid,
         patched ? " patch" : "", p.num_stages, ms / iters * 1e3, fl / (ms / iters * 1e-3) / 1e12);
}

int main(int argc, char** argv) {
  int devid = 0;
  CK(cudaSetDevice(devid));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, devid));
  num_sms = prop.multiProcessorCount;
  if (getenv("DP_SELFTEST_SMS")) num_sms = atoi(getenv("DP_SELFTEST_SMS"));   // a smaller grid: several work units per CTA (pair)
  printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, num_sms);
  int e = dp::gemm_init();
  if (e) { printf("gemm_init failed %d\n", e); return 2; }
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  const bool pair_only = argc > 1 && !strcmp(argv[1], "pair");
  const bool gn_only = argc > 1 && !strcmp(argv[1], "gn");      // the fused-GroupNorm cases only (A/B of DP_GN_P1)
  if (argc > 8 && !strcmp(argv[1], "perf")) {
    run_perf(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]), argc > 9 ? atoi(argv[9]) : 0,
             argc > 10 ? atoi(argv[10]) : 1);
    return 0;
  }

  const ConvCase cases[] = {
      // name                      B  H   W   C0  taps C1   N   bn  st  bias  rowv  resid silu  stats bf16  alpha
      {"gemm 512x128x256",         1, 1, 512, 256, 1,  0, 128, 128, 1, false,false,false,false,false,false, 1.f},
      {"gemm 300x256x512 bn256",   1, 1, 300, 512, 1,  0, 256, 256, 1, true, false,false,false,false,true,  1.f},
      {"gemm M=16 (temb-like)",    1, 1, 16,  128, 1,  0, 512, 256, 1, true, false,false,true, false,true,  1.f},
      {"gemm M=3 N=192",           1, 1, 3,   128, 1,  0, 192, 128, 1, true, false,false,false,false,true,  1.f},
      {"conv3x3 32x32 128->8 bn32", 2, 32, 32, 128, 9,  0, 8,   32,  1, true, false,false,false,false,false, 1.f},
      {"conv3x3 32x32 128->128",   2, 32, 32, 128, 9,  0, 128, 128, 1, true, true, true, false,true, false, 0.70710678f},
      {"conv3x3 16x16 256->256",   4, 16, 16, 256, 9,  0, 256, 256, 1, true, true, false,false,true, true,  1.f},
      {"conv3x3 8x8 256->256",     5, 8,  8,  256, 9,  0, 256, 128, 1, true, true, true, false,true, false, 0.70710678f},
      {"conv3x3 4x4 256->256",     11, 4, 4,  256, 9,  0, 256, 256, 1, true, true, true, false,true, false, 0.70710678f},
      {"conv3x3+1x1 16x16",        3, 16, 16, 256, 9, 512, 256, 256, 1, true, false,false,false,true, false, 0.70710678f},
      {"conv1x1 32x32 256->128",   2, 32, 32, 256, 1,  0, 128, 128, 1, true, false,true, false,false,false, 1.f},
      {"conv3x3 s2 32->16 128",    2, 16, 16, 128, 9,  0, 128, 128, 2, true, false,false,false,true, false, 1.f},
      {"conv3x3 256x256 64->128",  1, 256,256, 64, 9,  0, 128, 128, 1, true, true, false,false,true, false, 1.f},
      {"conv3x3 s2 256->128 64",   1, 128,128, 64, 9,  0, 128, 128, 2, true, false,false,false,false,false, 1.f},
      // bf16-only outputs: the row-per-thread epilogue without the shared-memory transpose
      {"bf16only gemm 300x256x512",    1, 1, 300, 512, 1,  0, 256, 128, 1, true, false,false,false,false,true,  1.f},
      {"bf16only gemm 300x264 silu",   1, 1, 300, 128, 1,  0, 264, 128, 1, true, false,false,true, false,true,  1.f},
      {"bf16only conv3x3 bn256",       3, 16, 16, 64, 9,  0, 256, 256, 1, false,false,false,false,false,true,  1.f},
      // CTA pairs (cta_group::2): 256 x BN tiles across two SMs
      {"pair conv3x3 32x32 128->128",  2, 32, 32, 128, 9,  0, 128, 128, 1, true, true, false,false,true, false, 1.f, 2},
      {"pair conv3x3 32x32 resid",     4, 32, 32, 128, 9,  0, 128, 128, 1, true, false,true, false,true, false, 0.70710678f, 2},
      {"pair conv3x3+1x1 32x32 bf16",  2, 32, 32, 128, 9, 256, 128, 128, 1, true, false,false,false,true, true, 1.f, 2},
      {"pair conv3x3 16x16 256->256",  4, 16, 16, 256, 9,  0, 256, 256, 1, true, true, false,false,true, true,  1.f, 2},
      {"pair conv3x3 8x8 256 bn128",   8, 8,  8,  256, 9,  0, 256, 128, 1, true, false,true, false,true, false, 0.70710678f, 2},
      {"pair conv3x3 4x4 256->256",    32, 4, 4,  256, 9,  0, 256, 256, 1, true, false,true, false,true, false, 0.70710678f, 2},
      {"pair conv3x3 s2 32->16 128",   2, 16, 16, 128, 9,  0, 128, 128, 2, true, false,false,false,true, false, 1.f, 2},
      {"pair conv 64x64 many tiles",   20, 64, 64, 64, 9,  0, 128, 128, 1, true, true, false,false,true, false, 1.f, 2},
      // fused GroupNorm + SiLU epilogue: the sample's accumulators resident in TMEM, two passes
      {"gn super-pair 32x32 128->128 (global exchange)", 3, 32, 32, 128, 9, 0, 128, 128, 1, true, true, false,false,false,true, 1.f, 2, 1},
      {"gn pair 32x32 128->128 (4 resident tiles, DSMEM exchange)", 3, 32, 32, 128, 9, 0, 128, 128, 1, true, true, false,false,false,true, 1.f, 2, 1},
      {"gn pair 32x32 256->256 (two N tiles)",                    2, 32, 32, 256, 9, 0, 256, 128, 1, true, true, false,false,false,true, 1.f, 2, 1},
      {"gn pair 32x32 many samples",                              40, 32, 32, 128, 9, 0, 128, 128, 1, true, true, false,false,false,true, 1.f, 2, 1},
      {"gn pair 16x16 256->256 bn256 (pair exchange)",            5, 16, 16, 256, 9, 0, 256, 256, 1, true, true, false,false,false,true, 1.f, 2, 1},
      {"gn pair 16x16 512->256 bn128 many",                       90, 16, 16, 512, 9, 0, 256, 128, 1, true, true, false,false,false,true, 1.f, 2, 1},
      {"gn single 8x8 256 bn128 (ragged)",                        5, 8,  8,  256, 9, 0, 256, 128, 1, true, true, false,false,false,true, 1.f, 1, 1},
      {"gn pair 8x8 256 bn128",                                   8, 8,  8,  256, 9, 0, 256, 128, 1, true, true, false,false,false,true, 1.f, 2, 1},
      {"gn single 4x4 256 bn256 (ragged)",                        11, 4, 4,  256, 9, 0, 256, 256, 1, true, true, false,false,false,true, 1.f, 1, 1},
      {"gn pair 4x4 256 bn128 no rowvec",                         32, 4, 4,  256, 9, 0, 256, 128, 1, true, false,false,false,false,true, 1.f, 2, 1},
      // "late" variant: fp32 result with residual / alpha / statistics AND the next block's GroupNorm + SiLU operand
      {"gn-late pair 32x32 128->128 resid",                       3, 32, 32, 128, 9, 0, 128, 128, 1, true, false,true, false,true, false, 0.70710678f, 2, 2},
      {"gn-late pair 32x32 +1x1 raw16",                           2, 32, 32, 128, 9, 256, 128, 128, 1, true, false,false,false,true, true, 0.70710678f, 2, 2},
      {"gn-late pair 16x16 256->256 bn256 resid",                 5, 16, 16, 256, 9, 0, 256, 256, 1, true, false,true, false,true, false, 0.70710678f, 2, 2},
      {"gn-late pair 16x16 1x1 (NIN_3) resid",                    6, 16, 16, 256, 1, 0, 256, 128, 1, true, false,true, false,true, false, 0.70710678f, 2, 2},
      {"gn-late single 8x8 256 bn128 resid (ragged)",             5, 8,  8,  256, 9, 0, 256, 128, 1, true, false,true, false,true, false, 0.70710678f, 1, 2},
      {"gn-late pair 4x4 256 bn128 resid",                        32, 4, 4,  256, 9, 0, 256, 128, 1, true, false,true, false,true, false, 0.70710678f, 2, 2},
      {"gn-late single 4x4 256 bn256 resid (ragged)",             11, 4, 4,  256, 9, 0, 256, 256, 1, true, false,true, false,true, false, 0.70710678f, 1, 2},
  };
  const int ncases = sizeof(cases) / sizeof(cases[0]);
  for (int i = 0; i < ncases; ++i) {
    if (quick && i >= 4) break;
    if (pair_only && cases[i].cg != 2) continue;
    if (gn_only && !cases[i].gn) continue;
    run_conv(cases[i]);
  }
  if (!pair_only && !gn_only) run_attention(3, 256, 256, 256);
  if (!quick && !pair_only && !gn_only) run_attention(2, 128, 512, 128);
  printf("selftest_gemm: %d failure(s)\n", failures);
  return failures ? 1 : 0;
}
