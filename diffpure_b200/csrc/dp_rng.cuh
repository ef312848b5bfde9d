// dp_rng.cuh -- counter-based standard normals shared by the update paths (dp_elem.cu update_kernel / init_state_kernel,
// the fused update epilogue of the output-conv GEMM in dp_gemm.cu) and the host check dp_normal_host().
// Philox4x32-10 keyed by the seed, counter = (global sample index, stream = step + 1 | 0 for the forward diffusion, pixel);
// Box-Muller on the four outputs gives the three channel values. Replaces torchsde's BrownianInterval stream
// (runners/diffpure_sde.py:234-238) and the torch.randn_like draws of the DDPM chains.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dp {

__host__ __device__ inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
  const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
  const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
  const uint32_t n1 = static_cast<uint32_t>(p1);
  const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
  const uint32_t n3 = static_cast<uint32_t>(p0);
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ inline float dp_normal_impl(unsigned long long seed, unsigned long long sample, unsigned int stream,
                                                unsigned int pixel, int c) {
  uint32_t c0 = static_cast<uint32_t>(sample), c1 = static_cast<uint32_t>(sample >> 32), c2 = stream, c3 = pixel;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const uint32_t a = (c < 2) ? c0 : c2, b = (c < 2) ? c1 : c3;
  const float u1 = (static_cast<float>(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (static_cast<float>(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * logf(u1));
  const float ang = 6.283185307179586f * u2;
  return (c == 1) ? rad * sinf(ang) : rad * cosf(ang);
}

}  // namespace dp
