// dp_engine.cpp -- the C-ABI engine: device buffers, the lowered UNet program, CUDA-graph capture and the
// device-resident purification loop (see include/diffpure_b200.h for the reference call sites it replaces).
#include "diffpure_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dp_attn.cuh"
#include "dp_elem.cuh"
#include "dp_gemm.cuh"
#include "dp_tmap.h"

namespace {

std::string g_create_error;

enum OpKind { OP_EMBED, OP_GEMM, OP_GN, OP_STATS, OP_STATS_REDUCE, OP_CONV_IN, OP_ATTN_SMALL, OP_SOFTMAX, OP_UPDATE,
              OP_GN_BWD, OP_SOFTMAX_BWD, OP_TRANSPOSE, OP_ATTN_SMALL_BWD, OP_GRAD_IN, OP_GN_FINALIZE, OP_PAD_IN, OP_ATTN_BLOCK };

struct StatsReduce {
  const float* in;
  float* out;
  int B, P, C;
};

// CTA-pair GEMM tiles (see dp_op_gemm): 0 = off, 1 = BN 128 only, 2 = BN 128 and 256; the DP_GEMM_PAIR environment
// variable overrides it for A/B runs
constexpr int kDefaultPairMode = 2;
// fewer pair tiles than this leave most SMs idle either way; 32 admits the 4x4 level of the B=512 CIFAR-10 model
// (64 pair tiles, one per CTA pair: measured 0.91 vs 1.01 ms per evaluation for its 40 convolutions)
constexpr long long kMinPairTiles = 32;

struct Op {
  OpKind kind;
  dp::EmbedParams embed;
  dp::GemmParams gemm;
  int bn = 128;
  int cg = 1;  // CTAs per tile
  bool softmax = false;
  int fused_update_op = -1;  // GEMM: index of the OP_UPDATE its epilogue absorbs in step mode; OP_UPDATE: index of that GEMM
  dp::GnParams gn;
  dp_stats_desc stats;
  StatsReduce sred;
  dp::ConvInParams cin;
  dp::AttnSmallParams attn;
  dp_softmax_desc smax;
  dp::UpdateParams upd;
  dp::GnBwdParams gnb;
  dp_softmax_bwd_desc smb;
  dp::TransposeParams tr;
  dp::AttnSmallBwdParams attnb;
  dp_grad_in_desc gin;
  dp_pad_in_desc pin;
  dp::AttnBlockParams ablk;
};

}  // namespace

struct dp_engine {
  int device = 0;
  int num_sms = 148;
  mutable std::string err;
  std::vector<void*> buffers;
  std::vector<size_t> sizes;
  std::vector<char> owned;  // 0: borrowed from the caller (dp_buffer_adopt), never freed here
  size_t total_bytes = 0;
  std::vector<Op> ops;
  bool finalized = false;
  bool update_fused = false;  // the per-step update (and the step-counter advance) run inside the output conv's epilogue
  int B = 0, H = 0, W = 0, Cout = 0;
  // engine-owned run state
  float* x_state = nullptr;          // NHWC fp32 [B,H,W,3]
  float* x_init = nullptr;           // copy of the initial state (anchor of the Langevin-dynamics update)
  float* eps_out = nullptr;          // NCHW fp32 [B,Cout,H,W]
  float* cond_per_sample = nullptr;  // [B]
  float* g_in = nullptr;             // NCHW fp32 [B,Cg,H,W]: dL/d(UNet output) of dp_unet_vjp
  int g_in_channels = 0;
  float* gn_ss = nullptr;            // [B][2][C] scale/shift scratch shared by all GroupNorm ops (stream-ordered)
  size_t gn_ss_floats = 0;
  void* xg = nullptr;                // super-pair exchange scratch of the fused-GroupNorm GEMMs: data | flags | epoch
  float* bwd_part = nullptr;         // scratch of the GroupNorm backward partial sums (stream-ordered, shared by all ops)
  size_t bwd_part_floats = 0;
  int* d_step = nullptr;
  float* d_cond = nullptr;
  float* d_coef = nullptr;
  int table_cap = 0;
  dp::CallParams* d_call = nullptr;
  cudaStream_t stream = nullptr;       // private stream: capture, profiling, calls made with stream == NULL
  cudaStream_t last_stream = nullptr;  // stream of the latest dp_unet_forward / dp_purify (dp_buffer_read waits for it)
  float* h_stage = nullptr;            // pinned staging of the per-call tables: coef [cap][8], cond [cap], CallParams
  cudaEvent_t staging_done = nullptr;
  cudaGraphExec_t g_forward = nullptr, g_step = nullptr;
};

namespace {

int fail(const dp_engine* e, int code, const std::string& msg) {
  if (e) e->err = msg;
  return code;
}
int cuda_fail(const dp_engine* e, cudaError_t ce, const char* what) {
  return fail(e, DP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
}
#define DP_CUDA(e, call)                                  \
  do {                                                    \
    cudaError_t ce_ = (call);                             \
    if (ce_ != cudaSuccess) return cuda_fail(e, ce_, #call); \
  } while (0)

bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2(long long v) {
  int s = 0;
  while ((1LL << s) < v) ++s;
  return s;
}

dp::StepTables tables_of(dp_engine* e, int ncoef) {
  dp::StepTables t;
  t.step = e->d_step;
  t.cond = e->d_cond;
  t.coef = e->d_coef;
  t.ncoef = ncoef;
  return t;
}

int ensure_run_state(dp_engine* e) {
  if (e->d_step) return DP_OK;
  DP_CUDA(e, cudaSetDevice(e->device));
  DP_CUDA(e, cudaMalloc(&e->d_step, 64));
  DP_CUDA(e, cudaMemset(e->d_step, 0, 64));
  e->table_cap = 4096;
  DP_CUDA(e, cudaMalloc(&e->d_cond, sizeof(float) * e->table_cap));
  DP_CUDA(e, cudaMalloc(&e->d_coef, sizeof(float) * e->table_cap * 8));
  DP_CUDA(e, cudaMemset(e->d_cond, 0, sizeof(float) * e->table_cap));
  DP_CUDA(e, cudaMemset(e->d_coef, 0, sizeof(float) * e->table_cap * 8));
  DP_CUDA(e, cudaMalloc(&e->d_call, sizeof(dp::CallParams)));
  DP_CUDA(e, cudaMemset(e->d_call, 0, sizeof(dp::CallParams)));
  DP_CUDA(e, cudaMallocHost(&e->h_stage, sizeof(float) * e->table_cap * 9 + sizeof(dp::CallParams) + 64));
  DP_CUDA(e, cudaEventCreateWithFlags(&e->staging_done, cudaEventDisableTiming));
  return DP_OK;
}

// mode: 0 = forward (per-sample cond, eps to eps_out), 1 = step (per-step tables, fused state update)
int run_op(dp_engine* e, size_t i, int mode, cudaStream_t s) {
  Op& op = e->ops[i];
  int rc = 0;
  switch (op.kind) {
    case OP_EMBED: {
      dp::EmbedParams p = op.embed;
      p.cond_per_sample = mode == 0 ? e->cond_per_sample : nullptr;
      rc = dp::launch_embed(p, s);
      break;
    }
    case OP_GEMM:
      if (mode != 0 && op.fused_update_op >= 0) {
        // step mode: the output conv applies the update in its epilogue (no eps tensor, no update / step-advance launch)
        const dp::UpdateParams& u = e->ops[op.fused_update_op].upd;
        dp::GemmParams g = op.gemm;
        g.out_f32 = nullptr;
        g.upd_x = e->x_state; g.upd_x_init = e->x_init;
        g.upd_step = e->d_step; g.upd_step_rw = e->d_step; g.upd_coef = e->d_coef; g.upd_call = e->d_call;
        g.upd_arrive = mode == 1 ? e->d_step + 1 : nullptr;   // mode 2 (profiling): the step counter stays put
        g.upd_cout = u.Cout; g.upd_hw = u.H * u.W; g.upd_B = u.B;
        rc = dp::launch_gemm(g, op.bn, op.softmax, e->num_sms, s, op.cg);
        break;
      }
      rc = dp::launch_gemm(op.gemm, op.bn, op.softmax, e->num_sms, s, op.cg);
      break;
    case OP_GN:
      rc = dp::launch_gn_apply(op.gn, op.gn.stats0 ? e->gn_ss : nullptr, e->num_sms, s);
      break;
    case OP_GN_FINALIZE:
      rc = dp::launch_gn_finalize(op.gn, e->gn_ss, s);
      break;
    case OP_STATS:
      rc = dp::launch_stats(op.stats.src, op.stats.stats, op.stats.B, op.stats.HW, op.stats.C, s);
      break;
    case OP_STATS_REDUCE:
      rc = dp::launch_stats_reduce(op.sred.in, op.sred.out, op.sred.B, op.sred.P, op.sred.C, s);
      break;
    case OP_CONV_IN: {
      dp::ConvInParams p = op.cin;
      p.x = e->x_state;
      rc = dp::launch_conv_in(p, s);
      break;
    }
    case OP_ATTN_SMALL:
      rc = dp::launch_attn_small(op.attn, s);
      break;
    case OP_UPDATE: {
      if (mode != 0 && op.fused_update_op >= 0) break;  // absorbed by the output conv's epilogue
      dp::UpdateParams u = op.upd;
      u.mode = mode;
      u.out_nchw = e->eps_out;
      u.x = e->x_state;
      u.x_init = e->x_init;
      u.call = e->d_call;
      rc = dp::launch_update(u, s);
      break;
    }
    case OP_GN_BWD: {
      dp::GnBwdParams g = op.gnb;
      g.part = e->bwd_part;
      rc = dp::launch_gn_bwd(g, s);
      break;
    }
    case OP_SOFTMAX_BWD:
      rc = dp::launch_softmax_bwd(static_cast<const __nv_bfloat16*>(op.smb.pnum_bf16), op.smb.rowsum, op.smb.dp,
                                  static_cast<__nv_bfloat16*>(op.smb.ds_bf16), static_cast<__nv_bfloat16*>(op.smb.pn_bf16),
                                  op.smb.rows, op.smb.T, s);
      break;
    case OP_TRANSPOSE:
      rc = dp::launch_transpose(op.tr, s);
      break;
    case OP_ATTN_SMALL_BWD:
      rc = dp::launch_attn_small_bwd(op.attnb, s);
      break;
    case OP_PAD_IN:
      rc = dp::launch_pad_in(e->x_state, static_cast<__nv_bfloat16*>(op.pin.out_bf16),
                             static_cast<long long>(op.pin.B) * op.pin.H * op.pin.W, op.pin.Cpad, s);
      break;
    case OP_GRAD_IN:
      rc = dp::launch_grad_in(e->g_in, static_cast<__nv_bfloat16*>(op.gin.out_bf16), op.gin.B, op.gin.C,
                              op.gin.H * op.gin.W, op.gin.Cpad, s);
      break;
    case OP_ATTN_BLOCK:
      rc = dp::launch_attn_block(op.ablk, e->num_sms, s);
      break;
    case OP_SOFTMAX:
      rc = dp::launch_softmax_rows(op.smax.src, static_cast<__nv_bfloat16*>(op.smax.out_bf16), op.smax.rows,
                                   op.smax.T, s);
      break;
  }
  if (rc != 0)
    return fail(e, DP_ERR_CUDA,
                "launch of op " + std::to_string(i) + " (kind " + std::to_string(op.kind) +
                    ") failed: " + cudaGetErrorString(static_cast<cudaError_t>(rc)));
  return DP_OK;
}

int run_ops(dp_engine* e, int mode, cudaStream_t s) {
  for (size_t i = 0; i < e->ops.size(); ++i)
    if (int rc = run_op(e, i, mode, s)) return rc;
  if (mode == 1 && !e->update_fused) {
    int rc = dp::launch_step_advance(e->d_step, s);
    if (rc) return fail(e, DP_ERR_CUDA, "launch_step_advance failed");
  }
  return DP_OK;
}

int capture(dp_engine* e, int mode, cudaGraphExec_t* out) {
  cudaGraph_t graph = nullptr;
  DP_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed));
  int rc = run_ops(e, mode, e->stream);
  cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
  if (rc != DP_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (ce != cudaSuccess) return cuda_fail(e, ce, "cudaStreamEndCapture");
  ce = cudaGraphInstantiate(out, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return cuda_fail(e, ce, "cudaGraphInstantiate");
  return DP_OK;
}

}  // namespace

extern "C" {

int dp_version(void) { return 100; }

int dp_create(dp_engine** out, int device) {
  if (!out) return DP_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(ce);
    return DP_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return DP_ERR_INVALID;
  }
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, device);
  if (ce != cudaSuccess) {
    g_create_error = cudaGetErrorString(ce);
    return DP_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "diffpure_b200 requires an sm_100a (B200) device, found sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor);
    return DP_ERR_CUDA;
  }
  ce = cudaSetDevice(device);
  if (ce != cudaSuccess) {
    g_create_error = cudaGetErrorString(ce);
    return DP_ERR_CUDA;
  }
  int rc = dp::gemm_init();
  if (rc) {
    g_create_error = std::string("gemm_init: ") + cudaGetErrorString(static_cast<cudaError_t>(rc));
    return DP_ERR_CUDA;
  }
  rc = dp::attn_block_init();
  if (rc) {
    g_create_error = std::string("attn_block_init: ") + cudaGetErrorString(static_cast<cudaError_t>(rc));
    return DP_ERR_CUDA;
  }
  dp_engine* e = new dp_engine();
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) {
    g_create_error = cudaGetErrorString(ce);
    delete e;
    return DP_ERR_CUDA;
  }
  *out = e;
  return DP_OK;
}

void dp_destroy(dp_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  if (e->g_forward) cudaGraphExecDestroy(e->g_forward);
  if (e->g_step) cudaGraphExecDestroy(e->g_step);
  for (size_t i = 0; i < e->buffers.size(); ++i)
    if (e->owned[i]) cudaFree(e->buffers[i]);
  cudaFree(e->x_state);
  cudaFree(e->x_init);
  cudaFree(e->eps_out);
  cudaFree(e->cond_per_sample);
  cudaFree(e->g_in);
  cudaFree(e->xg);
  cudaFree(e->gn_ss);
  cudaFree(e->bwd_part);
  cudaFree(e->d_step);
  cudaFree(e->d_cond);
  cudaFree(e->d_coef);
  cudaFree(e->d_call);
  cudaFreeHost(e->h_stage);
  if (e->staging_done) cudaEventDestroy(e->staging_done);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char* dp_last_error(const dp_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }
int dp_device_sm_count(const dp_engine* e) { return e ? e->num_sms : 0; }
size_t dp_bytes_allocated(const dp_engine* e) { return e ? e->total_bytes : 0; }
int dp_program_size(const dp_engine* e) { return e ? static_cast<int>(e->ops.size()) : 0; }
int dp_launches_per_eval(const dp_engine* e) { return e ? static_cast<int>(e->ops.size()) : 0; }
int dp_launches_per_step(const dp_engine* e) {
  // one replay of the step graph: with the update in the output conv's epilogue there is neither an update nor a
  // step-advance launch; otherwise both
  return e ? static_cast<int>(e->ops.size()) + (e->update_fused ? -1 : 1) : 0;
}

int dp_gemm_fused_gn_count(const dp_engine* e) {
  int n = 0;
  if (e)
    for (const Op& op : e->ops) n += (op.kind == OP_GEMM && op.gemm.gn_out != nullptr) ? 1 : 0;
  return n;
}

int dp_gemm_pair_count(const dp_engine* e) {
  int n = 0;
  if (e)
    for (const Op& op : e->ops) n += (op.kind == OP_GEMM && op.cg == 2) ? 1 : 0;
  return n;
}

int dp_buffer_alloc(dp_engine* e, size_t bytes, int* buf_id) {
  if (!e || !buf_id) return DP_ERR_INVALID;
  DP_CUDA(e, cudaSetDevice(e->device));
  void* p = nullptr;
  const size_t padded = (bytes + 255) & ~static_cast<size_t>(255);
  DP_CUDA(e, cudaMalloc(&p, padded ? padded : 256));
  DP_CUDA(e, cudaMemset(p, 0, padded ? padded : 256));
  e->buffers.push_back(p);
  e->sizes.push_back(bytes);
  e->owned.push_back(1);
  e->total_bytes += padded;
  *buf_id = static_cast<int>(e->buffers.size()) - 1;
  return DP_OK;
}

int dp_buffer_adopt(dp_engine* e, void* device_ptr, size_t bytes, int* buf_id) {
  if (!e || !buf_id || !device_ptr) return DP_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(device_ptr) & 255u) return fail(e, DP_ERR_INVALID, "dp_buffer_adopt: pointer must be 256-byte aligned");
  cudaPointerAttributes at;
  DP_CUDA(e, cudaPointerGetAttributes(&at, device_ptr));
  if (at.type != cudaMemoryTypeDevice || at.device != e->device)
    return fail(e, DP_ERR_INVALID, "dp_buffer_adopt: not device memory of this engine's device");
  e->buffers.push_back(device_ptr);
  e->sizes.push_back(bytes);
  e->owned.push_back(0);
  *buf_id = static_cast<int>(e->buffers.size()) - 1;
  return DP_OK;
}

void* dp_buffer_ptr(dp_engine* e, int buf_id) {
  if (!e || buf_id < 0 || buf_id >= static_cast<int>(e->buffers.size())) return nullptr;
  return e->buffers[buf_id];
}

int dp_buffer_write(dp_engine* e, int buf_id, size_t offset, const void* host_src, size_t bytes) {
  if (!e || buf_id < 0 || buf_id >= static_cast<int>(e->buffers.size())) return DP_ERR_INVALID;
  if (offset + bytes > e->sizes[buf_id]) return fail(e, DP_ERR_INVALID, "dp_buffer_write out of range");
  DP_CUDA(e, cudaMemcpy(static_cast<char*>(e->buffers[buf_id]) + offset, host_src, bytes, cudaMemcpyHostToDevice));
  return DP_OK;
}

int dp_buffer_read(dp_engine* e, int buf_id, size_t offset, void* host_dst, size_t bytes) {
  if (!e || buf_id < 0 || buf_id >= static_cast<int>(e->buffers.size())) return DP_ERR_INVALID;
  if (offset + bytes > e->sizes[buf_id]) return fail(e, DP_ERR_INVALID, "dp_buffer_read out of range");
  DP_CUDA(e, cudaStreamSynchronize(e->stream));
  if (e->last_stream && e->last_stream != e->stream) DP_CUDA(e, cudaStreamSynchronize(e->last_stream));
  DP_CUDA(e, cudaMemcpy(host_dst, static_cast<char*>(e->buffers[buf_id]) + offset, bytes, cudaMemcpyDeviceToHost));
  return DP_OK;
}

int dp_op_embed(dp_engine* e, const dp_embed_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (int rc = ensure_run_state(e)) return rc;
  if (d->dim % 2) return fail(e, DP_ERR_INVALID, "embed dim must be even");
  Op op;
  op.kind = OP_EMBED;
  op.embed.out = static_cast<__nv_bfloat16*>(d->out_bf16);
  op.embed.B = d->B;
  op.embed.dim = d->dim;
  op.embed.cos_first = d->cos_first;
  op.embed.half_minus_1 = d->half_minus_1;
  op.embed.cond_per_sample = nullptr;
  op.embed.tables = tables_of(e, 3);
  e->ops.push_back(op);
  return DP_OK;
}

// DP_GEMM_BN256: 1 (default) = 256-wide tiles wherever the heuristics below allow; 0 = never; 2 = not for 3x3 convolutions at
// 16x16 / 32x32 / 64x64, where BN = 128 tiles can use the row-patch mainloop (A/B switch)
static bool bn256_mode(const dp_gemm_desc* d, int hw) {
  static const int mode = [] { const char* v = std::getenv("DP_GEMM_BN256"); return v ? std::atoi(v) : 1; }();
  if (mode == 0) return false;
  if (mode == 2 && d->a[0].taps == 9 && d->a[0].stride == 1 && (d->W == 16 || d->W == 32 || d->W == 64) && hw >= 128) return false;
  return true;
}

int dp_op_gemm(dp_engine* e, const dp_gemm_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->nseg < 1 || d->nseg > 2) return fail(e, DP_ERR_INVALID, "gemm: nseg must be 1 or 2");
  if (d->N % 8 || d->N <= 0) return fail(e, DP_ERR_INVALID, "gemm: N must be a positive multiple of 8");
  const int hw = d->H * d->W;
  if (d->H > 1 && !(is_pow2(d->H) && is_pow2(d->W))) return fail(e, DP_ERR_INVALID, "gemm: H, W must be powers of two");
  if (d->H > 1 && hw < 16) return fail(e, DP_ERR_INVALID, "gemm: H*W must be >= 16");
  // ---- fused GroupNorm output: decide whether the sample's accumulators can stay resident in TMEM ----------------
  // HW <= 128: every tile holds whole samples (any tile kind). HW == 256: one CTA-pair tile per sample. HW == 1024: four
  // CTA-pair tiles of BN = 128 = all 512 TMEM columns of both SMs. Anything else (or DP_GEMM_GN=0): the unfused sequence.
  bool gn_fused = false, gn_super = false;
  int gn_tpg = 1, gn_stages = 2;
  if (d->gn_out_bf16) {
    const bool late = d->out_f32 != nullptr;   // the result itself is an output too (fp32 residual stream)
    if (d->rowscale || d->silu || d->softmax || d->bias_along_m || !d->gn_gamma || !d->gn_beta || d->gn_groups <= 0 ||
        d->N % d->gn_groups)
      return fail(e, DP_ERR_INVALID, "gemm: fused GroupNorm output: unsupported epilogue combination");
    if (!late && (d->out_bf16 || d->stats || d->resid || d->alpha != 1.0f))
      return fail(e, DP_ERR_INVALID, "gemm: a GroupNorm output without out_f32 replaces every other output (bias / rowvec only)");
    if (late && !d->stats)
      return fail(e, DP_ERR_INVALID, "gemm: a GroupNorm output next to out_f32 needs the partial-statistics tensor too");
    static const int gn_on = [] { const char* v = std::getenv("DP_GEMM_GN"); return v ? std::atoi(v) : 1; }();
    const int cpg = d->N / d->gn_groups;
    const bool shape_ok = (hw == 16 || hw == 64 || hw == 128 || hw == 256 || hw == 1024) && d->H > 1 && d->N % 128 == 0 &&
                          128 % cpg == 0 && (d->batch <= 1);
    // DP_GEMM_GN: 0 none, 1 default, 2 every late candidate, 3 resident kind only. The late kind pays where the MMAs of a
    // tile outlast the two-pass epilogue (measured at B=512, profiles/r02_gn_late_per_shape.txt): 3x3 convs at <= 16x16 win
    // 17-32 us per launch; the 32x32 convs (+88 us vs 84 us of GroupNorm kernels saved) and the short-K NIN_3 projection
    // (+54 vs 50 us) are epilogue / HBM bound already and keep the separate kernels.
    long long kall = 0;
    for (int sgi = 0; sgi < d->nseg; ++sgi) kall += static_cast<long long>(d->a[sgi].taps) * d->a[sgi].C;
    const bool late_pays = hw <= 256 && kall >= 2048;
    gn_fused = shape_ok && (late ? (gn_on == 2 || (gn_on == 1 && late_pays)) : gn_on >= 1);
    if (!gn_fused) {
      // unfused sequence: the GEMM without the normalised output, then gn_finalize + gn_apply
      dp_gemm_desc g = *d;
      g.gn_out_bf16 = nullptr;
      dp_gn_desc n;
      std::memset(&n, 0, sizeof(n));
      if (late) {
        n.src0 = d->out_f32;
        n.stats0 = d->stats;
      } else {   // engine-owned scratch: raw result in bf16 + partial statistics
        int braw, bst;
        const size_t rows = static_cast<size_t>(d->B) * hw;
        if (int rc = dp_buffer_alloc(e, rows * d->N * 2, &braw)) return rc;
        const size_t srows = hw >= 128 ? rows / 128 : static_cast<size_t>((d->B + 128 / hw - 1) / (128 / hw)) * (128 / hw);
        if (int rc = dp_buffer_alloc(e, srows * d->N * 2 * sizeof(float), &bst)) return rc;
        g.out_bf16 = e->buffers[braw];
        g.stats = static_cast<float*>(e->buffers[bst]);
        n.src0 = static_cast<const float*>(e->buffers[braw]);
        n.src0_is_bf16 = 1;
        n.stats0 = static_cast<const float*>(e->buffers[bst]);
      }
      if (int rc = dp_op_gemm(e, &g)) return rc;
      n.C0 = d->N;
      n.P0 = hw >= 128 ? hw / 128 : 1;
      n.gamma = d->gn_gamma; n.beta = d->gn_beta;
      n.B = d->B; n.H = d->H; n.W = d->W; n.groups = d->gn_groups; n.eps = d->gn_eps; n.silu = d->gn_silu;
      n.out_bf16 = d->gn_out_bf16;
      return dp_op_gn_apply(e, &n);
    }
    // 32x32: the sample's four 256-row tiles stay in the four accumulator stages of ONE CTA pair (exchange over DSMEM).
    // DP_GN_UPC=2 selects the measured alternative, two pairs per sample with half of TMEM free for the next sample: the
    // exchange then crosses clusters through global memory, whose round trip (possibly across the two dies) cost more than
    // the overlap won (221 vs 176 us on the 128->128 convolution at B=512).
    static const int gn_upc = [] { const char* v = std::getenv("DP_GN_UPC"); return v && std::atoi(v) == 2 ? 2 : 1; }();
    if (hw == 1024) { gn_tpg = gn_upc == 2 ? 2 : 4; gn_stages = 4; gn_super = gn_upc == 2; }
  }
  Op op;
  op.kind = OP_GEMM;
  dp::GemmParams& p = op.gemm;
  std::memset(&p, 0, sizeof(p));
  p.batch = d->batch > 0 ? d->batch : 1;
  op.softmax = d->softmax != 0;
  if (gn_fused) {
    p.gn_out = static_cast<__nv_bfloat16*>(d->gn_out_bf16);
    p.gn_gamma = d->gn_gamma; p.gn_beta = d->gn_beta;
    p.gn_cpg = d->N / d->gn_groups; p.gn_hw = hw; p.gn_eps = d->gn_eps; p.gn_silu = d->gn_silu;
    p.tpg = gn_tpg; p.acc_stages = gn_stages;
    static const int gn_p1 = [] { const char* v = std::getenv("DP_GN_P1"); return v ? std::atoi(v) : 1; }();
    p.gn_p1 = gn_p1;
    if (gn_super) {
      constexpr size_t kSlots = 128, kData = kSlots * 2 * 4 * 64 * 2 * sizeof(float), kFlags = kSlots * 4 * 8;
      if (!e->xg) {
        DP_CUDA(e, cudaSetDevice(e->device));
        DP_CUDA(e, cudaMalloc(&e->xg, kData + kFlags + 64));
        DP_CUDA(e, cudaMemset(e->xg, 0, kData + kFlags + 64));
      }
      p.upc = 2;
      p.xg_data = static_cast<float*>(e->xg);
      p.xg_flag = reinterpret_cast<unsigned long long*>(static_cast<char*>(e->xg) + kData);
      p.xg_epoch = reinterpret_cast<unsigned long long*>(static_cast<char*>(e->xg) + kData + kFlags);
    }
  }
  long long kprobe = 0;
  for (int sgi = 0; sgi < d->nseg; ++sgi) kprobe += static_cast<long long>(d->a[sgi].taps) * d->a[sgi].C;
  // tile width
  int bn = 128;
  if (op.softmax) {
    if (d->N != 128 && d->N != 256) return fail(e, DP_ERR_INVALID, "gemm softmax: N must be 128 or 256");
    bn = d->N;
  } else if (d->N <= 32 && !d->stats) {
    bn = 32;  // narrow output (the C->3|6 conv padded to 8 columns): 128x32 tiles waste 4x instead of 16x of the MMA
  } else if (gn_fused && hw == 1024 && !d->out_f32) {
    bn = 128;  // four resident accumulator stages need BN = 128
  } else if (d->N % 256 == 0 && kprobe > 512 && bn256_mode(d, hw) ) {
    // (K <= 512: four to eight k-blocks per tile, the epilogue dominates and the 8-warp BN = 128 epilogue wins: measured
    //  80 vs 90 us and 107 vs 159 us on the 16x16 attention projections, tests/selftest_gemm perf)
    dp::GemmParams probe;
    std::memset(&probe, 0, sizeof(probe));
    probe.batch = p.batch;
    dp::gemm_fill_geometry(probe, d->B, d->H, d->W, d->N, 256);
    if (static_cast<long long>(probe.m_tiles) * probe.n_tiles * probe.batch >= e->num_sms) bn = 256;
  }
  op.bn = bn;
  if (gn_fused && bn == 128) p.acc_stages = 4;  // all 512 TMEM columns: the MMAs may run further ahead of the two-pass epilogue
  dp::gemm_fill_geometry(p, d->B, d->H, d->W, d->N, bn);
  const dp::TileBox tb = dp::gemm_tile_box(d->H, d->W);
  std::string err;
  long long ktotal = 0;
  for (int s = 0; s < d->nseg; ++s) {
    const dp_gemm_aseg& a = d->a[s];
    if (a.C % 64 || a.C <= 0) return fail(e, DP_ERR_INVALID, "gemm: segment channels must be a positive multiple of 64");
    if (a.taps != 1 && a.taps != 9) return fail(e, DP_ERR_INVALID, "gemm: taps must be 1 or 9");
    if (a.stride != 1 && a.stride != 2) return fail(e, DP_ERR_INVALID, "gemm: stride must be 1 or 2");
    const int win = d->W * a.stride, hin = d->H * a.stride;
    // plain / batched GEMM: the map spans all batch entries' rows
    const int inner_n = d->inner > 0 ? d->inner : 1;
    const int outer_n = p.batch / inner_n;
    const int wdim = (d->H == 1) ? ((outer_n - 1) * d->a_batch_rows + (inner_n - 1) * d->a_inner_rows + d->W) : win;
    // the map exposes the whole channel pitch so per-head channel offsets stay in bounds
    if (dp::make_act_tmap(&p.a[s].tmap, a.act_bf16, a.c_total, a.c_total, wdim, hin, d->B, tb.bw, tb.bh, tb.bn,
                          a.stride, &err))
      return fail(e, DP_ERR_CUDA, "gemm: A tensor map: " + err);
    p.a[s].taps = a.taps;
    p.a[s].kchunks = a.C / 64;
    p.a[s].stride = a.stride;
    p.a[s].pad = a.pad;
    ktotal += static_cast<long long>(a.taps) * a.C;
  }
  p.nseg = d->nseg;
  p.a_batch_rows = d->a_batch_rows;
  p.b_batch_rows = d->b_batch_rows;
  p.out_batch_stride = d->out_batch_stride;
  p.inner = d->inner > 0 ? d->inner : 1;
  p.a_inner_k = d->a_inner_k;
  p.a_inner_rows = d->a_inner_rows;
  p.b_inner_k = d->b_inner_k;
  p.b_inner_rows = d->b_inner_rows;
  p.out_inner_stride = d->out_inner_stride;
  if (p.batch % p.inner) return fail(e, DP_ERR_INVALID, "gemm: batch must be a multiple of inner");
  p.bias = d->bias;
  p.bias_along_m = d->bias_along_m;
  p.rowvec = d->rowvec;
  p.rowvec_ld = d->rowvec_ld;
  if (d->rowvec) {
    if (!is_pow2(d->rowvec_rows_per_sample)) return fail(e, DP_ERR_INVALID, "gemm: rowvec rows/sample must be 2^k");
    p.rowvec_shift = ilog2(d->rowvec_rows_per_sample);
  }
  p.rowscale = d->rowscale;
  p.resid = d->resid;
  p.alpha = d->alpha;
  p.silu = d->silu;
  p.out_f32 = d->out_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16);
  p.ldc = d->ldc;
  p.stats = d->stats;
  p.softmax_scale = d->softmax_scale;
  p.rowsum_out = d->rowsum_out;
  if (op.softmax && (!p.out_bf16 || !p.rowsum_out)) return fail(e, DP_ERR_INVALID, "gemm softmax needs out_bf16 + rowsum_out");
  if (!op.softmax && !p.out_f32 && !p.out_bf16 && !p.gn_out) return fail(e, DP_ERR_INVALID, "gemm: no output");
  // CTA pairs (cta_group::2, 256 x BN tiles over two SMs) for the convolutions: less shared-memory operand traffic per
  // SM. DP_GEMM_PAIR=0 disables, 1 = BN 128 only, 2 = BN 128 and 256 (measurement switch).
  op.cg = 1;
  {
    const char* pm = std::getenv("DP_GEMM_PAIR");
    const int mode = pm ? std::atoi(pm) : kDefaultPairMode;
    const long long units = static_cast<long long>(p.m_tiles / 2) * p.n_tiles * p.batch;
    // measured (tests/selftest_gemm perf, B=512): 32x32 128->128 1102 vs 952 TF/s, 256->128 1380 vs 1084, with fp32 residual
    // 991 vs 911, 16x16 512->256 1843 vs 1652
    // (small grids too: a lone tile per CTA is L2->SM bandwidth bound and a pair CTA loads 25% fewer operand bytes)
    if (mode > 0 && (bn == 128 || mode > 1) && ktotal >= 1024 && units >= kMinPairTiles &&
        dp::gemm_pair_supported(p, bn, op.softmax))
      op.cg = 2;
    if (gn_fused && hw >= 256) {  // the sample spans the pair's two CTAs: pairs are part of the algorithm, not a heuristic
      if (!dp::gemm_pair_supported(p, bn, op.softmax)) return fail(e, DP_ERR_STATE, "gemm: fused GroupNorm needs CTA pairs");
      op.cg = 2;
      p.gn_xchg = p.upc == 2 ? 0 : 1;  // DSMEM exchange inside the pair (a super-pair's four CTAs go through global memory)
    }
  }
  p.num_stages = dp::gemm_max_stages(bn, op.cg, p.gn_out != nullptr);
  {
    // shared row patches for the 3x3 taps (halves the A bytes fetched from L2 at 32x32; DP_GEMM_PATCH=0 disables)
    static const int patch_on = [] { const char* v = std::getenv("DP_GEMM_PATCH"); return v ? std::atoi(v) : 1; }();
    if (patch_on && dp::gemm_enable_patch(p, bn, op.cg)) {
      const dp_gemm_aseg& a0 = d->a[0];
      if (dp::make_act_tmap(&p.a[0].tmap, a0.act_bf16, a0.c_total, a0.c_total, d->W, d->H, d->B, tb.bw, tb.bh + 2, 1, 1, &err))
        return fail(e, DP_ERR_CUDA, "gemm: A patch tensor map: " + err);
    }
  }
  if (dp::make_mat_tmap(&p.tmap_b, d->w_bf16, d->w_cols > 0 ? d->w_cols : ktotal, d->w_rows, d->w_pitch, bn / op.cg,
                        &err))
    return fail(e, DP_ERR_CUDA, "gemm: B tensor map: " + err);
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_gn_apply(dp_engine* e, const dp_gn_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  const int C = d->C0 + d->C1;
  if (d->C0 % 8 || d->C1 % 8 || (d->stats0 && C % d->groups)) return fail(e, DP_ERR_INVALID, "gn: channel counts");
  dp::GnParams p;
  std::memset(&p, 0, sizeof(p));
  p.src0 = d->src0; p.stats0 = d->stats0; p.C0 = d->C0; p.P0 = d->P0;
  if (d->src0_is_bf16) {
    if (d->C1 || d->resample) return fail(e, DP_ERR_INVALID, "gn: bf16 source needs a single un-resampled source");
    p.src0h = reinterpret_cast<const __nv_bfloat16*>(d->src0);
    p.src0 = nullptr;
  }
  p.src1 = d->src1; p.stats1 = d->stats1; p.C1 = d->C1; p.P1 = d->P1;
  p.gamma = d->gamma; p.beta = d->beta; p.film = d->film; p.film_ld = d->film_ld;
  p.B = d->B; p.H = d->H; p.W = d->W; p.groups = d->groups; p.eps = d->eps; p.silu = d->silu;
  p.resample = d->resample;
  p.out = static_cast<__nv_bfloat16*>(d->out_bf16);
  p.raw = static_cast<__nv_bfloat16*>(d->raw_bf16);
  p.raw_f32 = d->raw_f32;
  // Collapse long partial lists once (256x256 images: 512 partials / sample) so each CTA reads one row.
  for (int which = 0; which < 2; ++which) {
    const int P = which ? p.P1 : p.P0;
    const int Cx = which ? p.C1 : p.C0;
    const float* st = which ? p.stats1 : p.stats0;
    if (Cx == 0 || P <= 16) continue;
    int buf;
    if (int rc = dp_buffer_alloc(e, static_cast<size_t>(d->B) * Cx * 2 * sizeof(float), &buf)) return rc;
    Op r;
    r.kind = OP_STATS_REDUCE;
    r.sred.in = st;
    r.sred.out = static_cast<float*>(e->buffers[buf]);
    r.sred.B = d->B;
    r.sred.P = P;
    r.sred.C = Cx;
    e->ops.push_back(r);
    if (which) { p.stats1 = r.sred.out; p.P1 = 1; } else { p.stats0 = r.sred.out; p.P0 = 1; }
  }
  if (p.stats0) {
    const size_t need = static_cast<size_t>(d->B) * 2 * C;
    if (need > e->gn_ss_floats) {  // ops read the pointer from the engine at launch time: growing it here is safe
      cudaFree(e->gn_ss);
      e->gn_ss = nullptr;
      DP_CUDA(e, cudaMalloc(&e->gn_ss, need * sizeof(float)));
      e->gn_ss_floats = need;
    }
    Op f;
    f.kind = OP_GN_FINALIZE;
    f.gn = p;
    e->ops.push_back(f);
  }
  Op op;
  op.kind = OP_GN;
  op.gn = p;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_stats(dp_engine* e, const dp_stats_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  Op op;
  op.kind = OP_STATS;
  op.stats = *d;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_conv_in(dp_engine* e, const dp_conv_in_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->Cout % 4 || 256 % (d->Cout / 4)) return fail(e, DP_ERR_INVALID, "conv_in: Cout must divide 1024 and be a multiple of 4");
  Op op;
  op.kind = OP_CONV_IN;
  op.cin.x = nullptr;
  op.cin.w = d->w; op.cin.bias = d->bias; op.cin.out = d->out;
  op.cin.B = d->B; op.cin.H = d->H; op.cin.W = d->W; op.cin.Cout = d->Cout;
  e->ops.push_back(op);
  if (d->stats) {
    dp_stats_desc sd;
    sd.src = d->out; sd.B = d->B; sd.HW = d->H * d->W; sd.C = d->Cout; sd.stats = d->stats;
    return dp_op_stats(e, &sd);
  }
  return DP_OK;
}

int dp_op_attn_small(dp_engine* e, const dp_attn_small_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->T > 64 || d->d % 2) return fail(e, DP_ERR_INVALID, "attn_small: T <= 64 and even head dim required");
  Op op;
  op.kind = OP_ATTN_SMALL;
  op.attn.qkv = static_cast<const __nv_bfloat16*>(d->qkv_bf16);
  op.attn.out = static_cast<__nv_bfloat16*>(d->out_bf16);
  op.attn.B = d->B; op.attn.T = d->T; op.attn.heads = d->heads; op.attn.d = d->d; op.attn.scale = d->scale;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_attn_block(dp_engine* e, const dp_attn_block_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->T != dp::kAttnBlockT || d->C != dp::kAttnBlockC || d->B <= 0)
    return fail(e, DP_ERR_INVALID, "attn_block: the fused attention block is built for T = 256 tokens of C = 256 channels");
  if (!d->hn_bf16 || !d->w_bf16 || !d->bias || !d->resid || !d->out_f32)
    return fail(e, DP_ERR_INVALID, "attn_block: hn, w, bias, resid and out are required");
  Op op;
  op.kind = OP_ATTN_BLOCK;
  dp::AttnBlockParams& a = op.ablk;
  std::string err;
  if (dp::make_mat_tmap(&a.tmap_h, d->hn_bf16, d->C, static_cast<long long>(d->B) * d->T, d->C, 128, &err) ||
      dp::make_mat_tmap(&a.tmap_w, d->w_bf16, d->C, 4LL * d->C, d->C, 128, &err))
    return fail(e, DP_ERR_CUDA, "attn_block: " + err);
  a.bias = d->bias;
  a.resid = d->resid;
  a.out_f32 = d->out_f32;
  a.stats = d->stats;
  a.B = d->B;
  a.scale = d->scale;
  a.alpha = d->alpha;
  a.x_prefetch = [] { const char* v = std::getenv("DP_ATTN_XPF"); return v ? std::atoi(v) : 2; }();
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_softmax_rows(dp_engine* e, const dp_softmax_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->T % 4) return fail(e, DP_ERR_INVALID, "softmax_rows: T must be a multiple of 4");
  Op op;
  op.kind = OP_SOFTMAX;
  op.smax = *d;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_gn_bwd(dp_engine* e, const dp_gn_bwd_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  const int C = d->C0 + d->C1;
  if (C <= 0 || C % d->groups || !d->stats0 || !d->g) return fail(e, DP_ERR_INVALID, "gn_bwd: channels / statistics / gradient");
  if (d->C1 && (!d->src1 || !d->stats1 || !d->d1_f32)) return fail(e, DP_ERR_INVALID, "gn_bwd: second source incomplete");
  if (d->resample && (d->H % 2 || d->W % 2) && d->resample == 2) return fail(e, DP_ERR_INVALID, "gn_bwd: odd grid");
  Op op;
  op.kind = OP_GN_BWD;
  dp::GnBwdParams& p = op.gnb;
  std::memset(&p, 0, sizeof(p));
  if (d->src0_is_bf16) p.src0h = reinterpret_cast<const __nv_bfloat16*>(d->src0);
  else p.src0 = d->src0;
  p.stats0 = d->stats0; p.C0 = d->C0; p.P0 = d->P0;
  p.src1 = d->src1; p.stats1 = d->stats1; p.C1 = d->C1; p.P1 = d->P1;
  p.gamma = d->gamma; p.beta = d->beta;
  p.film = d->film; p.film_ld = d->film_ld;
  if (d->film && (d->C1 || !d->silu)) return fail(e, DP_ERR_INVALID, "gn_bwd: scale-shift rows go with a single source and SiLU");
  p.B = d->B; p.H = d->H; p.W = d->W; p.groups = d->groups; p.eps = d->eps; p.silu = d->silu; p.resample = d->resample;
  p.g = d->g; p.add0 = d->add0; p.add0_scale = d->add0_scale; p.add1 = d->add1;
  p.d0_f32 = d->d0_f32; p.d0_bf16 = static_cast<__nv_bfloat16*>(d->d0_bf16); p.d1_f32 = d->d1_f32;
  const size_t need = static_cast<size_t>(d->B) * ((d->H * d->W + 127) / 128) * C * 2;
  if (need > e->bwd_part_floats) {   // ops read the pointer from the engine at launch time: growing it here is safe
    DP_CUDA(e, cudaSetDevice(e->device));
    cudaFree(e->bwd_part);
    e->bwd_part = nullptr;
    DP_CUDA(e, cudaMalloc(&e->bwd_part, need * sizeof(float)));
    e->bwd_part_floats = need;
  }
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_softmax_bwd(dp_engine* e, const dp_softmax_bwd_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  Op op;
  op.kind = OP_SOFTMAX_BWD;
  op.smb = *d;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_transpose(dp_engine* e, const dp_transpose_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->batch <= 0 || d->batch > 65535) return fail(e, DP_ERR_INVALID, "transpose: batch out of range");
  Op op;
  op.kind = OP_TRANSPOSE;
  op.tr.in = static_cast<const __nv_bfloat16*>(d->in_bf16);
  op.tr.out = static_cast<__nv_bfloat16*>(d->out_bf16);
  op.tr.rows = d->rows; op.tr.cols = d->cols; op.tr.ld_in = d->ld_in; op.tr.ld_out = d->ld_out; op.tr.batch = d->batch;
  op.tr.in_batch_stride = d->in_batch_stride; op.tr.out_batch_stride = d->out_batch_stride;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_attn_small_bwd(dp_engine* e, const dp_attn_small_bwd_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->T > 64) return fail(e, DP_ERR_INVALID, "attn_small_bwd: T <= 64 required");
  Op op;
  op.kind = OP_ATTN_SMALL_BWD;
  op.attnb.qkv = static_cast<const __nv_bfloat16*>(d->qkv_bf16);
  op.attnb.go = static_cast<const __nv_bfloat16*>(d->go_bf16);
  op.attnb.out = static_cast<__nv_bfloat16*>(d->out_bf16);
  op.attnb.B = d->B; op.attnb.T = d->T; op.attnb.heads = d->heads; op.attnb.d = d->d; op.attnb.scale = d->scale;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_pad_in(dp_engine* e, const dp_pad_in_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->Cpad < 8 || d->Cpad % 8 || !d->out_bf16) return fail(e, DP_ERR_INVALID, "pad_in: Cpad must be a multiple of 8");
  Op op;
  op.kind = OP_PAD_IN;
  op.pin = *d;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_grad_in(dp_engine* e, const dp_grad_in_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (d->C <= 0 || d->Cpad < d->C) return fail(e, DP_ERR_INVALID, "grad_in: channels");
  Op op;
  op.kind = OP_GRAD_IN;
  op.gin = *d;
  e->g_in_channels = d->C;
  e->ops.push_back(op);
  return DP_OK;
}

int dp_op_update(dp_engine* e, const dp_update_desc* d) {
  if (!e || !d) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (int rc = ensure_run_state(e)) return rc;
  if ((d->Cout != 3 && d->Cout != 6) || d->ld < d->Cout) return fail(e, DP_ERR_INVALID, "update: Cout in {3,6}, ld >= Cout");
  Op op;
  op.kind = OP_UPDATE;
  std::memset(&op.upd, 0, sizeof(op.upd));
  op.upd.eps = d->eps; op.upd.ld = d->ld; op.upd.B = d->B; op.upd.H = d->H; op.upd.W = d->W; op.upd.Cout = d->Cout;
  op.upd.tables = tables_of(e, 8);
  e->Cout = d->Cout;
  // The producer of `eps` is the C -> 3|6 output conv on the narrow tile: in step mode its epilogue applies the update
  // (DP_FUSE_UPDATE=0 keeps the separate kernel, for A/B runs).
  static const int fuse = [] { const char* v = std::getenv("DP_FUSE_UPDATE"); return v ? std::atoi(v) : 1; }();
  if (fuse && !e->ops.empty()) {
    Op& g = e->ops.back();
    if (g.kind == OP_GEMM && g.bn == 32 && g.cg == 1 && !g.softmax && g.gemm.out_f32 == d->eps && !g.gemm.out_bf16 &&
        !g.gemm.resid && !g.gemm.rowvec && !g.gemm.rowscale && !g.gemm.stats && !g.gemm.silu && g.gemm.alpha == 1.0f &&
        !g.gemm.bias_along_m && g.gemm.batch == 1 && g.gemm.ldc == d->ld && g.gemm.M == d->B * d->H * d->W) {
      g.fused_update_op = static_cast<int>(e->ops.size());
      op.fused_update_op = static_cast<int>(e->ops.size()) - 1;
      e->update_fused = true;
    }
  }
  e->ops.push_back(op);
  return DP_OK;
}

int dp_finalize(dp_engine* e, int B, int H, int W) {
  if (!e) return DP_ERR_INVALID;
  if (e->finalized) return fail(e, DP_ERR_STATE, "program already finalized");
  if (e->ops.empty() || e->Cout == 0) return fail(e, DP_ERR_STATE, "program has no output conv");
  if (int rc = ensure_run_state(e)) return rc;
  e->B = B; e->H = H; e->W = W;
  const size_t hw = static_cast<size_t>(H) * W;
  DP_CUDA(e, cudaMalloc(&e->x_state, B * hw * 3 * sizeof(float)));
  DP_CUDA(e, cudaMemset(e->x_state, 0, B * hw * 3 * sizeof(float)));
  DP_CUDA(e, cudaMalloc(&e->x_init, B * hw * 3 * sizeof(float)));
  DP_CUDA(e, cudaMemset(e->x_init, 0, B * hw * 3 * sizeof(float)));
  DP_CUDA(e, cudaMalloc(&e->eps_out, B * hw * e->Cout * sizeof(float)));
  if (e->g_in_channels > 0) {
    DP_CUDA(e, cudaMalloc(&e->g_in, B * hw * e->g_in_channels * sizeof(float)));
    DP_CUDA(e, cudaMemset(e->g_in, 0, B * hw * e->g_in_channels * sizeof(float)));
  }
  DP_CUDA(e, cudaMalloc(&e->cond_per_sample, sizeof(float) * B));
  DP_CUDA(e, cudaMemset(e->cond_per_sample, 0, sizeof(float) * B));
  // eager warm-up run of both modes (sets function attributes, surfaces launch errors), then capture
  for (int mode = 0; mode < 2; ++mode) {
    if (int rc = run_ops(e, mode, e->stream)) return rc;
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) return cuda_fail(e, ce, "warm-up run of the program");
  }
  DP_CUDA(e, cudaMemset(e->d_step, 0, 64));
  if (int rc = capture(e, 0, &e->g_forward)) return rc;
  if (int rc = capture(e, 1, &e->g_step)) return rc;
  e->finalized = true;
  return DP_OK;
}

// Stream contract of dp_unet_forward / dp_purify: `stream` != NULL -> all work is enqueued on that stream (ordered behind
// whatever the caller enqueued before, e.g. the producer of x) and the call returns without synchronising; the caller
// synchronises that stream before reading the output on the host. `stream` == NULL -> the engine's private stream is
// used and the call blocks: it first waits for the caller's default-stream work, then for its own result. The engine's buffers are shared run state: calls on one engine
// must not overlap (not re-entrant), whichever streams they use.
static int pick_stream(dp_engine* e, void* stream, cudaStream_t* out) {
  if (stream) {
    *out = static_cast<cudaStream_t>(stream);
    return DP_OK;
  }
  DP_CUDA(e, cudaStreamSynchronize(nullptr));  // the caller's default-stream work (e.g. the producer of x) is complete
  *out = e->stream;
  return DP_OK;
}

int dp_unet_forward(dp_engine* e, const float* x_nchw, const float* cond, float* out_nchw, void* stream) {
  if (!e || !x_nchw || !cond || !out_nchw) return DP_ERR_INVALID;
  if (!e->finalized) return fail(e, DP_ERR_STATE, "dp_finalize has not been called");
  DP_CUDA(e, cudaSetDevice(e->device));
  cudaStream_t s = nullptr;
  if (int rc0 = pick_stream(e, stream, &s)) return rc0;
  e->last_stream = s;
  const int HW = e->H * e->W;
  int rc = dp::launch_init_state(x_nchw, x_nchw, e->x_state, e->B, 3, HW, 1.0f, 0.0f, 0, 0, s);
  if (rc) return fail(e, DP_ERR_CUDA, "init_state launch failed");
  DP_CUDA(e, cudaMemcpyAsync(e->cond_per_sample, cond, sizeof(float) * e->B, cudaMemcpyDeviceToDevice, s));
  DP_CUDA(e, cudaGraphLaunch(e->g_forward, s));
  DP_CUDA(e, cudaMemcpyAsync(out_nchw, e->eps_out, sizeof(float) * e->B * e->Cout * HW, cudaMemcpyDeviceToDevice, s));
  if (!stream) DP_CUDA(e, cudaStreamSynchronize(s));
  return DP_OK;
}

int dp_unet_vjp(dp_engine* e, const float* x_nchw, const float* cond, const float* g_nchw, float* gx_nchw, void* stream) {
  if (!e || !x_nchw || !cond || !g_nchw || !gx_nchw) return DP_ERR_INVALID;
  if (!e->finalized) return fail(e, DP_ERR_STATE, "dp_finalize has not been called");
  if (!e->g_in) return fail(e, DP_ERR_STATE, "the program has no data-gradient ops (no dp_op_grad_in)");
  DP_CUDA(e, cudaSetDevice(e->device));
  cudaStream_t s = nullptr;
  if (int rc0 = pick_stream(e, stream, &s)) return rc0;
  e->last_stream = s;
  const int HW = e->H * e->W;
  int rc = dp::launch_init_state(x_nchw, x_nchw, e->x_state, e->B, 3, HW, 1.0f, 0.0f, 0, 0, s);
  if (rc) return fail(e, DP_ERR_CUDA, "init_state launch failed");
  DP_CUDA(e, cudaMemcpyAsync(e->cond_per_sample, cond, sizeof(float) * e->B, cudaMemcpyDeviceToDevice, s));
  DP_CUDA(e, cudaMemcpyAsync(e->g_in, g_nchw, sizeof(float) * e->B * e->g_in_channels * HW, cudaMemcpyDeviceToDevice, s));
  DP_CUDA(e, cudaGraphLaunch(e->g_forward, s));
  DP_CUDA(e, cudaMemcpyAsync(gx_nchw, e->eps_out, sizeof(float) * e->B * e->Cout * HW, cudaMemcpyDeviceToDevice, s));
  if (!stream) DP_CUDA(e, cudaStreamSynchronize(s));
  return DP_OK;
}

int dp_purify(dp_engine* e, const float* x0_nchw, float* out_nchw, const dp_purify_params* p, void* stream) {
  if (!e || !x0_nchw || !out_nchw || !p) return DP_ERR_INVALID;
  if (!e->finalized) return fail(e, DP_ERR_STATE, "dp_finalize has not been called");
  if (p->steps <= 0 || p->steps > e->table_cap) return fail(e, DP_ERR_INVALID, "steps out of range");
  if (p->ncoef < 3 || p->ncoef > 8) return fail(e, DP_ERR_INVALID, "ncoef out of range");
  if (p->update_kind < 0 || p->update_kind > 2) return fail(e, DP_ERR_INVALID, "unknown update_kind");
  const int want = p->update_kind == DP_UPDATE_LEARNED_RANGE ? 6 : 3;
  if (e->Cout != want && !(p->update_kind != DP_UPDATE_LEARNED_RANGE && e->Cout == 6))
    return fail(e, DP_ERR_INVALID, "update_kind does not match the model's output channels");
  DP_CUDA(e, cudaSetDevice(e->device));
  cudaStream_t s = nullptr;
  if (int rc0 = pick_stream(e, stream, &s)) return rc0;
  e->last_stream = s;
  // per-step tables: the captured graph reads cond[*step] / coef[*step][:], laid out with a fixed pitch of 8 so the
  // graph's kernel parameters never change. Staged through engine-owned pinned memory: the previous call's copies have
  // completed before it is rewritten (cudaEventSynchronize on the staging event).
  DP_CUDA(e, cudaEventSynchronize(e->staging_done));
  float* coef8 = e->h_stage;
  float* hcond = e->h_stage + static_cast<size_t>(e->table_cap) * 8;
  dp::CallParams* cp = reinterpret_cast<dp::CallParams*>(hcond + e->table_cap);
  std::memset(coef8, 0, sizeof(float) * static_cast<size_t>(p->steps) * 8);
  for (int i = 0; i < p->steps; ++i)
    for (int j = 0; j < p->ncoef; ++j) coef8[static_cast<size_t>(i) * 8 + j] = p->coef[static_cast<size_t>(i) * p->ncoef + j];
  std::memcpy(hcond, p->cond, sizeof(float) * p->steps);
  cp->step_noise = p->step_noise;
  cp->seed = p->seed;
  cp->sample_offset = p->sample_offset;
  cp->update_kind = p->update_kind;
  cp->states = p->states;
  DP_CUDA(e, cudaMemcpyAsync(e->d_cond, hcond, sizeof(float) * p->steps, cudaMemcpyHostToDevice, s));
  DP_CUDA(e, cudaMemcpyAsync(e->d_coef, coef8, sizeof(float) * static_cast<size_t>(p->steps) * 8, cudaMemcpyHostToDevice, s));
  DP_CUDA(e, cudaMemcpyAsync(e->d_call, cp, sizeof(*cp), cudaMemcpyHostToDevice, s));
  DP_CUDA(e, cudaEventRecord(e->staging_done, s));
  DP_CUDA(e, cudaMemsetAsync(e->d_step, 0, 2 * sizeof(int), s));  // step counter + the CTA arrival counter beside it
  const int HW = e->H * e->W;
  int rc;
  if (p->in_h < 0 || p->in_w < 0 || p->out_h < 0 || p->out_w < 0 || (p->in_h == 0) != (p->in_w == 0) ||
      (p->out_h == 0) != (p->out_w == 0))
    return fail(e, DP_ERR_INVALID, "pre / post grids: give both extents or neither");
  if (p->in_h || p->in_unit_range)
    rc = dp::launch_init_state_pre(x0_nchw, p->init_noise, e->x_state, e->B, 3, e->H, e->W, p->in_h ? p->in_h : e->H,
                                   p->in_w ? p->in_w : e->W, p->in_unit_range, p->init_scale_x, p->init_scale_e, p->seed,
                                   p->sample_offset, s);
  else
    rc = dp::launch_init_state(x0_nchw, p->init_noise, e->x_state, e->B, 3, HW, p->init_scale_x, p->init_scale_e,
                               p->seed, p->sample_offset, s);
  if (rc) return fail(e, DP_ERR_CUDA, "init_state launch failed");
  if (p->states) {
    rc = dp::launch_nhwc_to_nchw(e->x_state, p->states, e->B, 3, HW, s);
    if (rc) return fail(e, DP_ERR_CUDA, "state record launch failed");
  }
  if (p->update_kind == DP_UPDATE_LINEAR_ANCHORED) {
    if (p->anchor) {
      rc = dp::launch_init_state(p->anchor, p->anchor, e->x_init, e->B, 3, HW, 1.f, 0.f, p->seed, p->sample_offset, s);
      if (rc) return fail(e, DP_ERR_CUDA, "anchor layout launch failed");
    } else {
      DP_CUDA(e, cudaMemcpyAsync(e->x_init, e->x_state, sizeof(float) * e->B * HW * 3, cudaMemcpyDeviceToDevice, s));
    }
  }
  for (int i = 0; i < p->steps; ++i) DP_CUDA(e, cudaGraphLaunch(e->g_step, s));
  if (p->out_h || p->out_unit_range || p->out_std[0] != 0.f) {
    dp::PostParams pp;
    pp.unit_range = p->out_unit_range;
    for (int i = 0; i < 3; ++i) { pp.mean[i] = p->out_mean[i]; pp.std[i] = p->out_std[i]; }
    rc = dp::launch_final_post(e->x_state, out_nchw, e->B, 3, e->H, e->W, p->out_h ? p->out_h : e->H,
                               p->out_w ? p->out_w : e->W, pp, s);
  } else {
    rc = dp::launch_nhwc_to_nchw(e->x_state, out_nchw, e->B, 3, HW, s);
  }
  if (rc) return fail(e, DP_ERR_CUDA, "final layout launch failed");
  if (!stream) DP_CUDA(e, cudaStreamSynchronize(s));
  return DP_OK;
}

int dp_profile_ops(dp_engine* e, int mode, float* ms, int* kinds, double* flops, int cap) {
  if (!e || !ms || !kinds || !flops) return DP_ERR_INVALID;
  if (!e->finalized) return fail(e, DP_ERR_STATE, "dp_finalize has not been called");
  const int n = static_cast<int>(e->ops.size());
  if (cap < n) return fail(e, DP_ERR_INVALID, "dp_profile_ops: capacity too small");
  DP_CUDA(e, cudaSetDevice(e->device));
  std::vector<cudaEvent_t> ev(2 * n);
  for (auto& x : ev) DP_CUDA(e, cudaEventCreate(&x));
  int rc = DP_OK;
  for (int i = 0; i < n && rc == DP_OK; ++i) {
    cudaEventRecord(ev[2 * i], e->stream);
    rc = run_op(e, i, mode == 1 ? 2 : mode, e->stream);  // 2 = step mode without advancing the step counter
    cudaEventRecord(ev[2 * i + 1], e->stream);
  }
  if (rc == DP_OK) {
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) rc = cuda_fail(e, ce, "dp_profile_ops");
  }
  for (int i = 0; i < n && rc == DP_OK; ++i) {
    cudaEventElapsedTime(&ms[i], ev[2 * i], ev[2 * i + 1]);
    kinds[i] = static_cast<int>(e->ops[i].kind);
    flops[i] = 0.0;
    if (e->ops[i].kind == OP_GEMM) {
      const dp::GemmParams& g = e->ops[i].gemm;
      double k = 0;
      for (int s2 = 0; s2 < g.nseg; ++s2) k += 64.0 * g.a[s2].taps * g.a[s2].kchunks;
      flops[i] = 2.0 * g.M * g.N * k * g.batch;
    } else if (e->ops[i].kind == OP_ATTN_BLOCK) {  // six T x C x C GEMMs per sample (T = C)
      flops[i] = 6.0 * 2.0 * dp::kAttnBlockT * dp::kAttnBlockC * dp::kAttnBlockC * e->ops[i].ablk.B;
    }
  }
  for (auto& x : ev) cudaEventDestroy(x);
  return rc;
}

float dp_normal_host(uint64_t seed, uint64_t sample, uint32_t stream, uint32_t pixel, int c) {
  return dp::dp_normal(seed, sample, stream, pixel, c);
}

}  // extern "C"
