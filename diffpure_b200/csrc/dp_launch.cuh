// dp_launch.cuh -- kernel launch helper with programmatic dependent launch (PDL).
//
// One UNet evaluation is ~530 dependent kernels, many of them a few microseconds long; with plain stream order every
// boundary pays grid drain + launch latency. With the programmatic-stream-serialization attribute the next kernel's CTAs
// are scheduled while the current kernel drains and park in griddepcontrol.wait, which returns once the predecessor
// grid has completed and its memory is visible. Contract: every kernel launched through launch_k() calls pdl_entry()
// before it touches global memory (so the ordering is exactly that of plain stream order); captured into CUDA graphs the
// attribute becomes a programmatic dependency edge.
//
// Measured on the CIFAR-10 headline loop (bench.py, B200, power-capped at ~1.66 GHz): 185.1 img/s with PDL vs 187.9
// without -- the loop is limited by the power cap, not by launch gaps, so PDL is OFF by default (DP_PDL=1 enables it).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace dp {

inline bool pdl_enabled() {
  static const int on = [] {
    const char* v = std::getenv("DP_PDL");
    return v ? std::atoi(v) : 0;
  }();
  return on != 0;
}

// cluster_x > 1 adds a cluster dimension (CTA pairs)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#ifdef __CUDACC__
// Wait for the stream predecessor (no-op without the launch attribute), then let the successor be scheduled.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

}  // namespace dp
