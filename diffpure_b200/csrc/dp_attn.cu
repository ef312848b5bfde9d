// dp_attn.cu -- AttnBlockpp (layerspp.py:75-91) behind its GroupNorm as one persistent tcgen05 kernel: see dp_attn.cuh.
//
// One CTA pair (cluster of 2, tcgen05 cta_group::2) per sample, persistent over samples. CTA `rank` owns tokens
// [128 rank, 128 rank + 128) of the sample (the M rows of q, the logits, the attention output and the block output), channels
// [128 rank, +128) of v^T, and the matching half of every N-split B operand. Per CTA: three 64 KiB operand buffers X, Y, Z
// (4 K-chunks of [128 rows][64 bf16], 128-byte swizzle -- what TMA writes and the UMMA descriptors read) and the whole TMEM
// (T0 = columns 0-255, T1 = 256-511). Per sample:
//
//   TMA      h -> X, Wv -> Y, Wq -> Z
//   G1  T0 = Wv . h^T         (A = Y, B = X)   v^T: lanes = channels, columns = the sample's 256 tokens
//   G2  T1 = h . Wq^T         (A = X, B = Z)
//   TMA      Wk -> Y                            (after G1)
//   D1  Z  = bf16(T0 + bv)                      (after G2: Z is free)  -- the B operand of G5, K-major along the keys
//   G3  T0 = h . Wk^T         (A = X, B = Y)
//   D2  X  = bf16((T1 + bq) scale)   D3  Y = bf16(T0 + bk)            (after G3)
//   G4  T1 = q . k^T          (A = X, B = Y)    logits, 128 queries x 256 keys per CTA
//   TMA      W3 -> Y                            (after G4)
//   D4  X  = bf16(exp2((T1 - rowmax) log2 e)), row sums kept           (after G4)
//   G5  T0 = P . v            (A = X, B = Z)
//   D5  X  = bf16(T0 / rowsum)                                         (after G5)
//   G6  T1 = o . W3^T         (A = X, B = Y)
//   TMA      Wq of the NEXT sample -> Z          (after G5)
//   E   out = (T1 + b3 + x) alpha, fp32, + per-channel partial statistics; the next sample's h / Wv loads and G1 run
//       underneath it. x was pulled into L2 at the start of the sample (cp.async.bulk.prefetch.L2): all CTA pairs reach E at
//       about the same time, and with x coming from HBM there E took 38 % of the kernel (ncu) while HBM idled elsewhere.
//
// Warp roles as in dp_gemm.cu: warp 0 = TMA producer (one lane, both CTAs), warp 1 = MMA issuer (leader CTA only),
// warp 2 = TMEM allocator, warps 4-11 = the eight drain / epilogue warps (two per TMEM lane quadrant, 128 columns each).
#include "dp_attn.cuh"
#include "dp_launch.cuh"
#include "dp_ptx.cuh"

#include <cuda_bf16.h>

namespace dp {

namespace {

constexpr int kBufBytes = 128 * 256 * 2;    // one operand buffer: 4 K-chunks
constexpr int kChunkBytes = 128 * 64 * 2;   // [128 rows][64 bf16]
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kStgPitch = 36;               // floats per staged row: 32 columns + 4 (conflict-free 128-bit accesses)
constexpr int kBiasFloats = 4 * 256;
constexpr int kStatFloats = 4 * 256 * 2;    // [4 lane quadrants][256 channels][sum | sumsq]
constexpr int kRedFloats = 2 * 2 * 128;     // [max | sum][column half][row]
constexpr int kStgFloats = kEpiWarps * 16 * kStgPitch;   // transposing staging of the output epilogue: 16 rows x 32 columns per warp
constexpr int kBarBytes = 256;
constexpr size_t kSmemBytes = 3 * kBufBytes + (kBiasFloats + kStatFloats + kRedFloats + kStgFloats) * 4 + kBarBytes;
static_assert(kSmemBytes <= 232448, "shared memory budget");

// barrier map (8 bytes each)
constexpr int BAR_LD = 0;      // 5: h, Wv, Wq, Wk, W3 landed (leader CTA's barriers, both CTAs' bytes)
constexpr int BAR_G = 5;       // 6: GEMM i complete (tcgen05.commit multicast to both CTAs)
constexpr int BAR_D = 11;      // 4: drains D1, D2+D3, D4, D5 complete (leader's barriers, 16 warp arrivals)
constexpr int BAR_E = 15;      // epilogue done with T1 (leader's barrier, 16 warp arrivals)
constexpr int BAR_HOLDER = 20; // TMEM base address

__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pulls [p, p + bytes) into L2 (no destination: the later loads of the epilogue hit there)
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// the same for one TMA box of a tensor map
__device__ __forceinline__ void prefetch_l2_box(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// 32 consecutive K elements [col0, col0 + 32) of operand row `row` into a K-major SWIZZLE_128B buffer: four 16-byte chunks,
// chunk j of a 128-byte row stored at position j ^ (row & 7) (the layout TMA produces and make_kmajor_sw128_desc describes)
__device__ __forceinline__ void put32(uint8_t* buf, int row, int col0, const uint32_t (&pk)[16]) {
  uint8_t* tile = buf + (col0 >> 6) * kChunkBytes + row * 128;
  const int j0 = (col0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(tile + (((j0 + j) ^ (row & 7)) << 4)) =
        make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
}

__global__ void __launch_bounds__(kThreads, 1) attn_block_kernel(const __grid_constant__ AttnBlockParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  uint8_t* const Xp = smem_raw;
  uint8_t* const Yp = smem_raw + kBufBytes;
  uint8_t* const Zp = smem_raw + 2 * kBufBytes;
  const uint32_t X = base, Y = base + kBufBytes, Z = base + 2 * kBufBytes;
  float* const sbias = reinterpret_cast<float*>(smem_raw + 3 * kBufBytes);
  float* const sstats = sbias + kBiasFloats;
  float* const sred = sstats + kStatFloats;
  float* const staging = sred + kRedFloats;
  uint64_t* const bars = reinterpret_cast<uint64_t*>(staging + kStgFloats);
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int i) { return bar0 + 8u * i; };
  volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(bars + BAR_HOLDER);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmap_h);
    tma_prefetch_desc(&p.tmap_w);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 5; ++i) mbar_init(bar(BAR_LD + i), 1);
    for (int i = 0; i < 6; ++i) mbar_init(bar(BAR_G + i), 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar(BAR_D + i), 2 * kEpiWarps);
    mbar_init(bar(BAR_E), 2 * kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_holder)), 512);
  tc_fence_before_sync();
  cluster_sync_all();  // the peer's barriers must exist before any remote signal
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;
  pdl_entry();

  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = static_cast<int>(blockIdx.x >> 1);
  const int npairs = static_cast<int>(gridDim.x >> 1);
  const uint32_t T0 = tmem_base, T1 = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (each CTA loads its own halves)
    if (elect_one()) {
      auto load = [&](uint32_t dst, const CUtensorMap* map, int ld, int row) {
        if (rank == 0) mbar_arrive_expect_tx(bar(BAR_LD + ld), 2 * kBufBytes);  // both CTAs' bytes land on the leader's barrier
        const uint32_t sig = mapa_u32(bar(BAR_LD + ld), 0);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) tma_load_2d_pair(dst + kc * kChunkBytes, map, sig, kc * 64, row);
      };
      // this CTA's 128 rows of x (one contiguous 128 KiB block) on their way into L2 ahead of the output epilogue
      auto prefetch_x = [&](int s) {
        const char* xr = reinterpret_cast<const char*>(p.resid) +
                         (static_cast<size_t>(s) * kAttnBlockT + rank * 128) * kAttnBlockC * sizeof(float);
#pragma unroll 1
        for (int i = 0; i < 8; ++i) prefetch_l2_bulk(xr + i * 16384, 16384);
      };
      int n = 0;
      if (pair < p.B) load(Z, &p.tmap_w, 2, 0 * 256 + rank * 128);   // Wq of the first sample
      for (int s = pair; s < p.B; s += npairs, ++n) {
        const uint32_t ph = n & 1, pph = ph ^ 1u;
        if (n > 0) mbar_wait(bar(BAR_G + 5), pph);            // G6 of the previous sample: X and Y are free
        load(X, &p.tmap_h, 0, s * kAttnBlockT + rank * 128);
        load(Y, &p.tmap_w, 1, 2 * 256 + rank * 128);          // Wv
        if (p.x_prefetch == 1) prefetch_x(s);
        mbar_wait(bar(BAR_G + 0), ph);                        // G1 read Wv
        load(Y, &p.tmap_w, 3, 1 * 256 + rank * 128);          // Wk
        mbar_wait(bar(BAR_G + 3), ph);                        // G4 read k
        load(Y, &p.tmap_w, 4, 3 * 256 + rank * 128);          // W3
        if (p.x_prefetch >= 2) prefetch_x(s);
        if (s + npairs < p.B) {
          if (p.x_prefetch == 3) {   // the next sample's h rows as well: their TMA load is issued the moment G6 retires
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) prefetch_l2_box(&p.tmap_h, kc * 64, (s + npairs) * kAttnBlockT + rank * 128);
          }
          mbar_wait(bar(BAR_G + 4), ph);                      // G5 read v^T: Z is free for the next sample's Wq
          load(Z, &p.tmap_w, 2, 0 * 256 + rank * 128);
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, for both SMs)
    constexpr uint32_t idesc = make_idesc_bf16(256, 256);
    auto gemm = [&](uint32_t a, uint32_t b, uint32_t d, int done) {
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const uint64_t adesc = make_kmajor_sw128_desc(a + kc * kChunkBytes);
          const uint64_t bdesc = make_kmajor_sw128_desc(b + kc * kChunkBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_pair(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
        }
        umma_commit_pair(bar(BAR_G + done), 3);
      }
      __syncwarp();
    };
    int n = 0;
    for (int s = pair; s < p.B; s += npairs, ++n) {
      const uint32_t ph = n & 1, pph = ph ^ 1u;
      mbar_wait(bar(BAR_LD + 0), ph);
      mbar_wait(bar(BAR_LD + 1), ph);
      gemm(Y, X, T0, 0);                                      // G1: v^T (T0 was drained by the previous D5)
      mbar_wait(bar(BAR_LD + 2), ph);
      if (n > 0) mbar_wait(bar(BAR_E), pph);                  // the previous epilogue has read T1
      gemm(X, Z, T1, 1);                                      // G2: q
      mbar_wait(bar(BAR_LD + 3), ph);
      mbar_wait(bar(BAR_D + 0), ph);                          // D1: T0 drained
      gemm(X, Y, T0, 2);                                      // G3: k
      mbar_wait(bar(BAR_D + 1), ph);                          // D2, D3: q in X, k in Y, T0 and T1 drained
      gemm(X, Y, T1, 3);                                      // G4: logits
      mbar_wait(bar(BAR_D + 2), ph);                          // D4: P in X, T1 drained
      gemm(X, Z, T0, 4);                                      // G5: P v
      mbar_wait(bar(BAR_LD + 4), ph);
      mbar_wait(bar(BAR_D + 3), ph);                          // D5: o in X, T0 drained
      gemm(X, Y, T1, 5);                                      // G6: output projection
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ drains and the epilogue
    const int q = warp & 3;              // TMEM lane quadrant == warp_id % 4
    const int hlf = (warp - 4) >> 2;     // which 128 columns of the 256
    const int row = q * 32 + lane;       // this thread's accumulator row (TMEM lane) inside the CTA's 128
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const int te = (warp - 4) * 32 + lane;
    for (int i = te; i < kBiasFloats; i += 32 * kEpiWarps) sbias[i] = __ldg(p.bias + i);
    named_bar(1, 32 * kEpiWarps);
    float* const stg = staging + (warp - 4) * (16 * kStgPitch);
    const int c4 = (lane & 7) * 4;       // epilogue: 32 columns at a time, this thread's 4 of them ...
    const int rsub = lane >> 3;          // ... in rows rsub, rsub + 4, ..., rsub + 28 of the warp's 32
    const float alpha = p.alpha;
    const uint32_t arrive_leader[5] = {mapa_u32(bar(BAR_D + 0), 0), mapa_u32(bar(BAR_D + 1), 0), mapa_u32(bar(BAR_D + 2), 0),
                                       mapa_u32(bar(BAR_D + 3), 0), mapa_u32(bar(BAR_E), 0)};
    // the operand tile this warp wrote must be visible to the tensor core (async proxy) of its SM before the leader's
    // MMA warp is told; the TMEM reads are ordered by the tcgen05 fence
    auto drained = [&](int which) {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_release(arrive_leader[which]);
    };
    // TMEM [row][c0, c0 + 32) -> (acc + add) * mul -> bf16 -> operand buffer; `per_col` selects a per-column or per-row add
    auto drain = [&](uint32_t tsrc, uint8_t* dst, const float* col_add, float row_add, float mul) {
      uint32_t r[32], r2[32];
#pragma unroll 1
      for (int blk = 0; blk < 4; blk += 2) {
        const int c0 = hlf * 128 + blk * 32;
        tmem_ld_32x32b_x32(tsrc + lane_bits + c0, r);
        tmem_ld_32x32b_x32(tsrc + lane_bits + c0 + 32, r2);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a0 = col_add ? col_add[c0 + 2 * j] : row_add, a1 = col_add ? col_add[c0 + 2 * j + 1] : row_add;
          pk[j] = pack_bf16((__uint_as_float(r[2 * j]) + a0) * mul, (__uint_as_float(r[2 * j + 1]) + a1) * mul);
        }
        put32(dst, row, c0, pk);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a0 = col_add ? col_add[c0 + 32 + 2 * j] : row_add, a1 = col_add ? col_add[c0 + 33 + 2 * j] : row_add;
          pk[j] = pack_bf16((__uint_as_float(r2[2 * j]) + a0) * mul, (__uint_as_float(r2[2 * j + 1]) + a1) * mul);
        }
        put32(dst, row, c0 + 32, pk);
      }
    };
    float* const smax = sred;                 // [column half][row]
    float* const ssum = sred + 2 * 128;
    int n = 0;
    for (int s = pair; s < p.B; s += npairs, ++n) {
      const uint32_t ph = n & 1;
      // ---- D1: v^T (+ bv, a per-row = per-channel constant here) -> Z
      mbar_wait(bar(BAR_G + 0), ph);
      mbar_wait(bar(BAR_G + 1), ph);          // G2 has read Wq out of Z
      tc_fence_after_sync();
      drain(T0, Zp, nullptr, sbias[2 * 256 + rank * 128 + row], 1.0f);
      drained(0);
      // ---- D2: q (+ bq, scaled) -> X; D3: k (+ bk) -> Y
      mbar_wait(bar(BAR_G + 2), ph);
      tc_fence_after_sync();
      drain(T1, Xp, sbias, 0.f, p.scale);
      drain(T0, Yp, sbias + 256, 0.f, 1.0f);
      drained(1);
      // ---- D4: softmax numerator of this thread's row -> X, row sum kept
      mbar_wait(bar(BAR_G + 3), ph);
      tc_fence_after_sync();
      float rinv;
      {
        uint32_t r[32], r2[32];
        float mx = -INFINITY;
#pragma unroll 1
        for (int blk = 0; blk < 4; blk += 2) {
          const int c0 = hlf * 128 + blk * 32;
          tmem_ld_32x32b_x32(T1 + lane_bits + c0, r);
          tmem_ld_32x32b_x32(T1 + lane_bits + c0 + 32, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r[j]), __uint_as_float(r2[j])));
        }
        smax[hlf * 128 + row] = mx;
        named_bar(2 + q, 64);                 // the two warps of this lane quadrant
        mx = fmaxf(mx, smax[(hlf ^ 1) * 128 + row]);
        constexpr float kLog2e = 1.4426950408889634f;
        const float moff = mx * kLog2e;
        float sum = 0.f;
#pragma unroll 1
        for (int blk = 0; blk < 4; blk += 2) {
          const int c0 = hlf * 128 + blk * 32;
          tmem_ld_32x32b_x32(T1 + lane_bits + c0, r);
          tmem_ld_32x32b_x32(T1 + lane_bits + c0 + 32, r2);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(exp2f(fmaf(__uint_as_float(r[2 * j]), kLog2e, -moff)),
                                                            exp2f(fmaf(__uint_as_float(r[2 * j + 1]), kLog2e, -moff)));
            sum += __low2float(b2) + __high2float(b2);   // over the rounded values the P v product reads
            pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
          }
          put32(Xp, row, c0, pk);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(exp2f(fmaf(__uint_as_float(r2[2 * j]), kLog2e, -moff)),
                                                            exp2f(fmaf(__uint_as_float(r2[2 * j + 1]), kLog2e, -moff)));
            sum += __low2float(b2) + __high2float(b2);
            pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
          }
          put32(Xp, row, c0 + 32, pk);
        }
        ssum[hlf * 128 + row] = sum;
        named_bar(2 + q, 64);
        // fixed order (half 0 + half 1) so both warps of the quadrant divide by the same value
        rinv = 1.0f / (ssum[row] + ssum[128 + row]);
      }
      drained(2);
      // ---- D5: attention output / row sum -> X
      mbar_wait(bar(BAR_G + 4), ph);
      tc_fence_after_sync();
      drain(T0, Xp, nullptr, 0.f, rinv);
      drained(3);
      // ---- E: (T1 + b3 + x) alpha -> fp32 output + per-channel partial statistics (as dp_gemm.cu's plain epilogue)
      mbar_wait(bar(BAR_G + 5), ph);
      tc_fence_after_sync();
      {
        const long long tbase = (static_cast<long long>(s) * kAttnBlockT + rank * 128 + q * 32) * kAttnBlockC;
        const float* const resp = p.resid + tbase;
        float* const outf = p.out_f32 + tbase;
        struct Pre {
          float4 rs[8];
        };
        auto prefetch = [&](Pre& f, int blk) {   // the residual of one 32-column block: rows rsub + 4 i8
          const int cc = hlf * 128 + blk * 32 + c4;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8)
            f.rs[i8] = __ldg(reinterpret_cast<const float4*>(resp + (i8 * 4 + rsub) * kAttnBlockC + cc));
        };
        Pre cur, nxt;
        prefetch(cur, 0);
        uint32_t r[32];
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
          const int cc = hlf * 128 + blk * 32 + c4;
          if (blk + 1 < 4) prefetch(nxt, blk + 1);
          tmem_ld_32x32b_x32(T1 + lane_bits + hlf * 128 + blk * 32, r);
          tmem_ld_wait();
          if (blk == 3) {  // this warp has read its share of T1: the next sample's q may overwrite it
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(arrive_leader[4]);
          }
          const float4 b3 = *reinterpret_cast<const float4*>(sbias + 3 * 256 + cc);
          float st[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) st[i] = 0.f;
          // the transposing staging holds 16 rows (the shared-memory budget): rows 0-15 of the warp, then rows 16-31
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if ((lane >> 4) == hf) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stg + (lane & 15) * kStgPitch + 4 * j) =
                    make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3]));
            }
            __syncwarp();
            float4 vv[4];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) vv[i4] = *reinterpret_cast<const float4*>(stg + (i4 * 4 + rsub) * kStgPitch + c4);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              float4 v = vv[i4];
              const float4 rs = cur.rs[hf * 4 + i4];
              v.x = (v.x + b3.x + rs.x) * alpha;
              v.y = (v.y + b3.y + rs.y) * alpha;
              v.z = (v.z + b3.z + rs.z) * alpha;
              v.w = (v.w + b3.w + rs.w) * alpha;
              *reinterpret_cast<float4*>(outf + (hf * 16 + i4 * 4 + rsub) * kAttnBlockC + cc) = v;
              st[0] += v.x; st[1] += v.y; st[2] += v.z; st[3] += v.w;
              st[4] += v.x * v.x; st[5] += v.y * v.y; st[6] += v.z * v.z; st[7] += v.w * v.w;
            }
            __syncwarp();
          }
          // fold the 4 row groups (lanes l, l+8, l+16, l+24) in a fixed order
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            st[i] += __shfl_xor_sync(0xffffffffu, st[i], 8);
            st[i] += __shfl_xor_sync(0xffffffffu, st[i], 16);
          }
          if (rsub == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float* d = sstats + (q * 256 + cc + u) * 2;
              d[0] = st[u];
              d[1] = st[4 + u];
            }
          }
          if (blk + 1 < 4) cur = nxt;
        }
        named_bar(1, 32 * kEpiWarps);
        if (p.stats != nullptr) {                  // the four lane quadrants in a fixed order: deterministic partial sums
          float sm = 0.f, sq = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            sm += sstats[(k * 256 + te) * 2];
            sq += sstats[(k * 256 + te) * 2 + 1];
          }
          float* o = p.stats + ((static_cast<long long>(s) * 2 + rank) * kAttnBlockC + te) * 2;
          o[0] = sm;
          o[1] = sq;
        }
        named_bar(1, 32 * kEpiWarps);
      }
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();  // no CTA of the pair may exit while its peer can still signal it
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace

int attn_block_init() {
  return static_cast<int>(cudaFuncSetAttribute(attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(kSmemBytes)));
}

int launch_attn_block(const AttnBlockParams& p, int num_sms, cudaStream_t stream) {
  if (p.B <= 0) return 0;
  const int slots = num_sms / 2;
  const int pairs = p.B < slots ? p.B : slots;
  return static_cast<int>(launch_k(attn_block_kernel, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, 2, p));
}

}  // namespace dp
