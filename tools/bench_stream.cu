// Microbenchmark: fp32 -> (scale/shift/SiLU) -> bf16 streaming variants, to calibrate gn_apply against HBM peak.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
__device__ __forceinline__ uint32_t pk(float a, float b) { __nv_bfloat162 t = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&t); }

// variant A: one 8-channel vector per thread, plain grid (no persistence)
__global__ void k_simple(const float4* __restrict__ in, uint4* __restrict__ out, size_t nvec, float a, float b) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  float4 v0 = __ldg(in + 2 * i), v1 = __ldg(in + 2 * i + 1);
  uint4 o;
  o.x = pk(silu_f(v0.x * a + b), silu_f(v0.y * a + b)); o.y = pk(silu_f(v0.z * a + b), silu_f(v0.w * a + b));
  o.z = pk(silu_f(v1.x * a + b), silu_f(v1.y * a + b)); o.w = pk(silu_f(v1.z * a + b), silu_f(v1.w * a + b));
  out[i] = o;
}
// variant B: U vectors per thread, strided by blockDim (coalesced), loads first
template <int U>
__global__ void k_unroll(const float4* __restrict__ in, uint4* __restrict__ out, size_t nvec, float a, float b) {
  size_t base = ((size_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
  float4 v[U][2];
#pragma unroll
  for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < nvec) { v[u][0] = __ldg(in + 2 * i); v[u][1] = __ldg(in + 2 * i + 1); } }
#pragma unroll
  for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < nvec) {
    uint4 o;
    o.x = pk(silu_f(v[u][0].x * a + b), silu_f(v[u][0].y * a + b)); o.y = pk(silu_f(v[u][0].z * a + b), silu_f(v[u][0].w * a + b));
    o.z = pk(silu_f(v[u][1].x * a + b), silu_f(v[u][1].y * a + b)); o.w = pk(silu_f(v[u][1].z * a + b), silu_f(v[u][1].w * a + b));
    out[i] = o; } }
}
// variant C: each lane loads 16B contiguous per instruction (fully coalesced 512B/warp), two planes
template <int U>
__global__ void k_coal(const float4* __restrict__ in, uint2* __restrict__ out, size_t nvec4, float a, float b) {
  size_t base = ((size_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
  float4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < nvec4) v[u] = __ldg(in + i); }
#pragma unroll
  for (int u = 0; u < U; ++u) { size_t i = base + (size_t)u * blockDim.x; if (i < nvec4) {
    uint2 o; o.x = pk(silu_f(v[u].x * a + b), silu_f(v[u].y * a + b)); o.y = pk(silu_f(v[u].z * a + b), silu_f(v[u].w * a + b)); out[i] = o; } }
}
__global__ void k_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ldg(in + i);
}

int main() {
  const size_t nelem = (size_t)512 * 1024 * 128;  // 32x32x128 x B=512
  float* in; __nv_bfloat16* out; float* out32;
  CK(cudaMalloc(&in, nelem * 4)); CK(cudaMalloc(&out, nelem * 2)); CK(cudaMalloc(&out32, nelem * 4));
  CK(cudaMemset(in, 0, nelem * 4));
  float* flush; CK(cudaMalloc(&flush, 256 << 20));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, double bytes, auto launch) {
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
      CK(cudaMemsetAsync(flush, 1, 256 << 20));
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    printf("%-28s %8.1f us  %7.1f GB/s\n", name, best * 1e3, bytes / best / 1e6);
  };
  const size_t nvec = nelem / 8, nvec4 = nelem / 4;
  const double bytes = nelem * 6.0;
  run("copy f32->f32 (8B/elt)", nelem * 8.0, [&] { k_copy<<<(unsigned)((nvec4 + 255) / 256), 256>>>((const float4*)in, (float4*)out32, nvec4); });
  run("simple 32B/thr", bytes, [&] { k_simple<<<(unsigned)((nvec + 255) / 256), 256>>>((const float4*)in, (uint4*)out, nvec, 1.1f, 0.1f); });
  run("unroll2 32B/thr", bytes, [&] { k_unroll<2><<<(unsigned)((nvec + 511) / 512), 256>>>((const float4*)in, (uint4*)out, nvec, 1.1f, 0.1f); });
  run("unroll4 32B/thr", bytes, [&] { k_unroll<4><<<(unsigned)((nvec + 1023) / 1024), 256>>>((const float4*)in, (uint4*)out, nvec, 1.1f, 0.1f); });
  run("coal16 x1", bytes, [&] { k_coal<1><<<(unsigned)((nvec4 + 255) / 256), 256>>>((const float4*)in, (uint2*)out, nvec4, 1.1f, 0.1f); });
  run("coal16 x4", bytes, [&] { k_coal<4><<<(unsigned)((nvec4 + 1023) / 1024), 256>>>((const float4*)in, (uint2*)out, nvec4, 1.1f, 0.1f); });
  run("coal16 x8", bytes, [&] { k_coal<8><<<(unsigned)((nvec4 + 2047) / 2048), 256>>>((const float4*)in, (uint2*)out, nvec4, 1.1f, 0.1f); });
  return 0;
}
