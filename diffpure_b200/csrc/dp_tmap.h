// dp_tmap.h -- host-side construction of the TMA tensor maps the GEMM kernel consumes.
// The driver entry point cuTensorMapEncodeTiled is resolved at run time through the CUDA runtime,
// so the library does not link against libcuda directly.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

namespace dp {

// NHWC bf16 activation tensor seen as (C, W, H, B); box (64, bw, bh, bn) output pixels with an
// optional spatial traversal stride (stride-2 convolution). 128-byte swizzle, OOB -> zeros.
// `c_total` is the channel count of the underlying tensor (row pitch), `c` the channels exposed.
int make_act_tmap(CUtensorMap* out, const void* base, int c, int c_total, int w, int h, int b, int bw,
                  int bh, int bn, int stride, std::string* err);

// bf16 matrix [rows, k] with K contiguous; box (64, box_rows).
int make_mat_tmap(CUtensorMap* out, const void* base, long long k, long long rows, long long pitch_elems,
                  int box_rows, std::string* err);

}  // namespace dp
