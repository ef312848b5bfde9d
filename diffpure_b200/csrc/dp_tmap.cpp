#include "dp_tmap.h"

#include <mutex>

namespace dp {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  static std::string init_err;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
      init_err = std::string("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: ") +
                 cudaGetErrorString(e);
      return;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (fn == nullptr && err) *err = init_err;
  return fn;
}

int encode(CUtensorMap* out, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box, const cuuint32_t* estr, std::string* err) {
  EncodeTiledFn fn = get_encode(err);
  if (!fn) return -1;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
             " (rank " + std::to_string(rank) + ", dims";
      for (int i = 0; i < rank; ++i) *err += " " + std::to_string(dims[i]);
      *err += ", box";
      for (int i = 0; i < rank; ++i) *err += " " + std::to_string(box[i]);
      *err += ")";
    }
    return -2;
  }
  return 0;
}

}  // namespace

int make_act_tmap(CUtensorMap* out, const void* base, int c, int c_total, int w, int h, int b, int bw,
                  int bh, int bn, int stride, std::string* err) {
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w),
                              static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(b)};
  const cuuint64_t pitch = static_cast<cuuint64_t>(c_total) * 2;
  const cuuint64_t strides[3] = {pitch, pitch * w, pitch * w * h};
  // With a traversal stride s the box spans bw*s input pixels and picks every s-th one.
  const cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(bw * stride), static_cast<cuuint32_t>(bh * stride),
                             static_cast<cuuint32_t>(bn)};
  const cuuint32_t estr[4] = {1u, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1u};
  return encode(out, const_cast<void*>(base), 4, dims, strides, box, estr, err);
}

int make_mat_tmap(CUtensorMap* out, const void* base, long long k, long long rows, long long pitch_elems,
                  int box_rows, std::string* err) {
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_elems) * 2};
  const cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1u, 1u};
  return encode(out, const_cast<void*>(base), 2, dims, strides, box, estr, err);
}

}  // namespace dp
