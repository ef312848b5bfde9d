// dp_gemm.cu -- persistent, warp-specialised tcgen05 implicit-GEMM convolution / GEMM kernel (sm_100a).
//
// Warp roles (256 threads, 1 CTA per SM):
//   warp 0 (one elected lane) : TMA producer   -- A halo tiles (4-D map, OOB zero fill = conv padding)
//                                                 and weight tiles into a num_stages smem ring
//   warp 1 (one elected lane) : tcgen05.mma issuer, 128 x BN x 16 per instruction, fp32 accum in TMEM,
//                                                 two accumulator stages so the epilogue overlaps the next tile
//   warp 2                    : TMEM allocator / deallocator
//   warps 4..7                : epilogue: tcgen05.ld -> smem transpose -> fused bias / time-embedding add /
//                               residual / scale / SiLU -> coalesced fp32|bf16 stores, plus deterministic
//                               per-channel GroupNorm partial statistics of the tile; or row softmax.
#include "dp_gemm.cuh"
#include "dp_ptx.cuh"

#include <cstdio>

namespace dp {

namespace {

constexpr int kNumThreads = 256;
constexpr int kStageABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kStgPitch = 36;  // floats; 144-byte rows keep float4 accesses aligned and conflict-free
constexpr int kStagingFloats = 4 * 32 * kStgPitch;
constexpr int kMaxStages = 8;

template <int BN>
struct Smem {
  static constexpr int kStageBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kStageABytes + kStageBBytes;
  static constexpr int kStatsFloats = 8 * BN * 2;
  static constexpr int kBarBytes = 256;
  static constexpr size_t total(int stages) {
    return static_cast<size_t>(stages) * kStageBytes + kStagingFloats * 4 + kStatsFloats * 4 + kBarBytes;
  }
};

struct Tile {
  int b, mt, nt, n0, h0, w0;
};

__device__ __forceinline__ Tile decode_tile(const GemmParams& p, int t) {
  Tile c;
  c.nt = t % p.n_tiles;
  const int r = t / p.n_tiles;
  c.mt = r % p.m_tiles;
  c.b = r / p.m_tiles;
  if (p.imgs_per_tile > 1) {
    c.n0 = c.mt * p.imgs_per_tile;
    c.h0 = 0;
    c.w0 = 0;
  } else {
    c.n0 = c.mt / p.tiles_per_img;
    const int rem = c.mt - c.n0 * p.tiles_per_img;
    c.h0 = (rem / p.tiles_w) * p.bh;
    c.w0 = (rem % p.tiles_w) * p.bw;
  }
  return c;
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

template <int BN, bool kSoftmax>
__global__ void __launch_bounds__(kNumThreads, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
  using L = Smem<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* sm = smem_raw;
  if ((base & 1023u) != 0u) __trap();

  const int stages = p.num_stages;
  float* staging = reinterpret_cast<float*>(sm + static_cast<size_t>(stages) * L::kStageBytes);
  float* sstats = staging + kStagingFloats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sstats + L::kStatsFloats);
  const uint32_t bar0 = smem_u32(bars);
  // barrier map (8 bytes each): full[0..8) empty[8..16) tfull[16..18) tempty[18..20) ; holder at 20
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (16 + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (18 + a); };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.a[0].tmap);
    if (p.nseg > 1) tma_prefetch_desc(&p.a[1].tmap);
    tma_prefetch_desc(&p.tmap_b);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), 2 * BN);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  const int total_tiles = p.m_tiles * p.n_tiles * p.batch;
  int num_kb = 0;
  for (int s = 0; s < p.nseg; ++s) num_kb += p.a[s].taps * p.a[s].kchunks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const Tile c = decode_tile(p, t);
        int kglobal = 0;
        const int brow = c.nt * BN + c.b * p.b_batch_rows;
        for (int s = 0; s < p.nseg; ++s) {
          const GemmASeg& seg = p.a[s];
          for (int tap = 0; tap < seg.taps; ++tap) {
            const int ky = (seg.taps == 9) ? tap / 3 : 0;
            const int kx = (seg.taps == 9) ? tap - 3 * ky : 0;
            const int c1 = c.w0 * seg.stride + kx - seg.pad + c.b * p.a_batch_rows;
            const int c2 = c.h0 * seg.stride + ky - seg.pad;
            for (int kc = 0; kc < seg.kchunks; ++kc) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              mbar_arrive_expect_tx(full_bar(stage), L::kStageBytes);
              const uint32_t sa = base + stage * L::kStageBytes;
              tma_load_4d(sa, &seg.tmap, full_bar(stage), kc * kBlockK, c1, c2, c.n0);
              tma_load_2d(sa + kStageABytes, &p.tmap_b, full_bar(stage), kglobal, brow);
              kglobal += kBlockK;
              if (++stage == stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = base + stage * L::kStageBytes;
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + kStageABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    // TMEM (lane = row) -> registers -> smem [32 rows][36] -> registers (lane = 4 columns x 4 row groups):
    // every shared/global access below is a conflict-free / fully coalesced 128-bit access.
    const int q = warp - 4;  // TMEM lane quadrant == warp_id % 4
    float* stg = staging + q * (32 * kStgPitch);
    const int c4 = (lane & 7) * 4;
    const int rsub = lane >> 3;
    const bool has_bias_n = p.bias != nullptr && !p.bias_along_m;
    const bool has_bias_m = p.bias != nullptr && p.bias_along_m;
    const bool has_rowvec = p.rowvec != nullptr;
    const bool has_rowscale = p.rowscale != nullptr;
    const bool has_resid = p.resid != nullptr;
    const bool has_f32 = p.out_f32 != nullptr;
    const bool has_bf16 = p.out_bf16 != nullptr;
    const bool do_silu = p.silu != 0;
    const bool do_stats = p.stats != nullptr;
    const float alpha = p.alpha;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const Tile c = decode_tile(p, t);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      const int row0 = c.mt * kBlockM + q * 32;  // row within the batch entry
      const long long obase = static_cast<long long>(c.b) * p.out_batch_stride;
      uint32_t r[32];

      if constexpr (kSoftmax) {
        float mx = -INFINITY;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          tmem_ld_32x32b_x32(taddr + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        const float sc = p.softmax_scale * 1.4426950408889634f;
        float sum = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          tmem_ld_32x32b_x32(taddr + ch * 32, r);
          tmem_ld_wait();
          if (ch == BN / 32 - 1) {
            tc_fence_before_sync();
            mbar_arrive(tempty_bar(as));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 e4;
            float* e = reinterpret_cast<float*>(&e4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float ev = exp2f((__uint_as_float(r[4 * j + u]) - mx) * sc);
              e[u] = __bfloat162float(__float2bfloat16_rn(ev));
              sum += e[u];
            }
            *reinterpret_cast<float4*>(stg + lane * kStgPitch + 4 * j) = e4;
          }
          __syncwarp();
          const int col = c.nt * BN + ch * 32 + c4;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int rr = i8 * 4 + rsub;
            const int row = row0 + rr;
            const float4 v = *reinterpret_cast<const float4*>(stg + rr * kStgPitch + c4);
            if (row < p.M && col < p.N) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(p.out_bf16 + obase + static_cast<long long>(row) * p.ldc + col) = pk;
            }
          }
          __syncwarp();
        }
        if (row0 + lane < p.M) p.rowsum_out[static_cast<long long>(c.b) * p.M + row0 + lane] = sum;
      } else {
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          tmem_ld_32x32b_x32(taddr + ch * 32, r);
          tmem_ld_wait();
          if (ch == BN / 32 - 1) {
            // accumulator fully read: hand the TMEM stage back to the MMA warp
            tc_fence_before_sync();
            mbar_arrive(tempty_bar(as));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * kStgPitch + 4 * j) =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
          const int col = c.nt * BN + ch * 32 + c4;
          const bool colok = col < p.N;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_bias_n && colok) bias4 = *reinterpret_cast<const float4*>(p.bias + col);
          float st[16];  // [half][sum|sumsq][4 cols]
#pragma unroll
          for (int i = 0; i < 16; ++i) st[i] = 0.f;
          // Issue all residual loads of this 32x32 block before any store: `resid` and `out` may alias
          // from the compiler's point of view, which would otherwise serialise one HBM round trip per row.
          float4 rs8[8];
          if (has_resid) {
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8) {
              const int row = row0 + i8 * 4 + rsub;
              rs8[i8] = (row < p.M && colok)
                            ? __ldg(reinterpret_cast<const float4*>(p.resid + obase + static_cast<long long>(row) * p.ldc + col))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int rr = i8 * 4 + rsub;
            const int row = row0 + rr;
            float4 v = *reinterpret_cast<const float4*>(stg + rr * kStgPitch + c4);
            if (row < p.M && colok) {
              if (has_rowscale) {
                const float rinv = 1.0f / p.rowscale[static_cast<long long>(c.b) * p.M + row];
                v.x *= rinv; v.y *= rinv; v.z *= rinv; v.w *= rinv;
              }
              v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
              if (has_bias_m) {
                const float bm = p.bias[row];
                v.x += bm; v.y += bm; v.z += bm; v.w += bm;
              }
              if (has_rowvec) {
                const float4 rv = *reinterpret_cast<const float4*>(
                    p.rowvec + static_cast<long long>(row >> p.rowvec_shift) * p.rowvec_ld + col);
                v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
              }
              if (do_silu) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
              const long long o = obase + static_cast<long long>(row) * p.ldc + col;
              if (has_resid) {
                const float4 rs = rs8[i8];
                v.x += rs.x; v.y += rs.y; v.z += rs.z; v.w += rs.w;
              }
              v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
              if (has_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = v;
              if (has_bf16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(p.out_bf16 + o) = pk;
              }
              float* h = st + (i8 >= 4 ? 8 : 0);
              h[0] += v.x; h[1] += v.y; h[2] += v.z; h[3] += v.w;
              h[4] += v.x * v.x; h[5] += v.y * v.y; h[6] += v.z * v.z; h[7] += v.w * v.w;
            }
          }
          if (do_stats) {
            // fold the 4 row groups (lanes l, l+8, l+16, l+24) in a fixed order
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              st[i] += __shfl_xor_sync(0xffffffffu, st[i], 8);
              st[i] += __shfl_xor_sync(0xffffffffu, st[i], 16);
            }
            if (rsub == 0) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float* d = sstats + ((2 * q + hh) * BN + ch * 32 + c4 + u) * 2;
                  d[0] = st[hh * 8 + u];
                  d[1] = st[hh * 8 + 4 + u];
                }
            }
          }
          __syncwarp();
        }
        if (do_stats) {
          // combine the 8 half-warp slots in a fixed order -> deterministic partial sums
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int te = q * 32 + lane;
          const int nseg = p.stat_nseg;
          const int per = 8 / nseg;
          for (int chn = te; chn < BN; chn += 128) {
            const int col = c.nt * BN + chn;
            if (col < p.N) {
              for (int sg = 0; sg < nseg; ++sg) {
                float s = 0.f, qq = 0.f;
                for (int k = 0; k < per; ++k) {
                  const float* d = sstats + ((sg * per + k) * BN + chn) * 2;
                  s += d[0];
                  qq += d[1];
                }
                const long long segid =
                    (static_cast<long long>(c.b) * p.m_tiles + c.mt) * nseg + sg;
                float* o = p.stats + (segid * p.N + col) * 2;
                o[0] = s;
                o[1] = qq;
              }
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, bool kSoftmax>
int launch_t(const GemmParams& p, int num_sms, cudaStream_t stream) {
  const int total = p.m_tiles * p.n_tiles * p.batch;
  if (total <= 0) return 0;
  const int grid = total < num_sms ? total : num_sms;
  const size_t smem = Smem<BN>::total(p.num_stages);
  gemm_kernel<BN, kSoftmax><<<grid, kNumThreads, smem, stream>>>(p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

size_t gemm_smem_bytes(int bn, int stages) {
  return bn == 256 ? Smem<256>::total(stages) : Smem<128>::total(stages);
}

int gemm_max_stages(int bn) {
  const size_t cap = 232448;  // 227 KiB opt-in maximum per CTA on sm_100
  int s = kMaxStages;
  while (s > 2 && gemm_smem_bytes(bn, s) > cap) --s;
  return s;
}

int gemm_init() {
  cudaError_t e;
#define DP_SET(BN, SM)                                                                         \
  e = cudaFuncSetAttribute(gemm_kernel<BN, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           static_cast<int>(Smem<BN>::total(gemm_max_stages(BN))));            \
  if (e != cudaSuccess) return static_cast<int>(e);
  DP_SET(128, false)
  DP_SET(256, false)
  DP_SET(128, true)
  DP_SET(256, true)
#undef DP_SET
  return 0;
}

TileBox gemm_tile_box(int H, int W) {
  TileBox t;
  if (W >= kBlockM || H == 1) {  // plain GEMM rows (H == 1): always 128-row boxes, OOB rows read as zero
    t.bw = kBlockM;
    t.bh = 1;
    t.bn = 1;
  } else {
    t.bw = W;
    t.bh = (kBlockM / W) < H ? (kBlockM / W) : H;
    t.bn = kBlockM / (t.bw * t.bh);
  }
  return t;
}

void gemm_fill_geometry(GemmParams& p, int B, int H, int W, int N, int bn) {
  const TileBox t = gemm_tile_box(H, W);
  p.H = H;
  p.W = W;
  p.bw = t.bw;
  p.bh = t.bh;
  const int hw = H * W;
  if (hw >= kBlockM || H == 1) {
    p.imgs_per_tile = 1;
    p.tiles_w = (W + t.bw - 1) / t.bw;
    const int tiles_h = (H + t.bh - 1) / t.bh;
    p.tiles_per_img = p.tiles_w * tiles_h;
    p.m_tiles = B * p.tiles_per_img;
    p.stat_nseg = 1;
  } else {
    p.imgs_per_tile = kBlockM / hw;
    p.tiles_w = 1;
    p.tiles_per_img = 1;
    p.m_tiles = (B + p.imgs_per_tile - 1) / p.imgs_per_tile;
    p.stat_nseg = p.imgs_per_tile;
  }
  p.M = B * hw;
  p.N = N;
  p.n_tiles = (N + bn - 1) / bn;
  if (p.batch <= 0) p.batch = 1;
  p.num_stages = gemm_max_stages(bn);
}

int launch_gemm(const GemmParams& p, int bn, bool softmax, int num_sms, cudaStream_t stream) {
  if (bn == 256) {
    return softmax ? launch_t<256, true>(p, num_sms, stream) : launch_t<256, false>(p, num_sms, stream);
  }
  return softmax ? launch_t<128, true>(p, num_sms, stream) : launch_t<128, false>(p, num_sms, stream);
}

}  // namespace dp
