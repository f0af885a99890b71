// fsweep_pdl.cuh — programmatic dependent launch (Hopper / Blackwell) for the kernel chain of a captured step.
//
// The headline step is a chain of eight short kernels (exp map, two FFT passes, sweep, finalize, criteria total, exp
// map adjoint, Adam), each 3 - 14 us: the gap between the last block of one kernel retiring and the first block of the
// next one running (~1.5 - 2 us per edge, even inside a CUDA graph) is a sizeable share of the step.  A kernel launched
// through launch_pdl carries cudaLaunchAttributeProgrammaticStreamSerialization: its blocks may be SCHEDULED while
// the preceding kernel of the stream is still running, and pdl_sync() at the top of the kernel
// (griddepcontrol.wait) holds them until that kernel has completed and its memory is visible — only then does the
// kernel release ITS dependents (griddepcontrol.launch_dependents), so at most two kernels of the chain are ever
// co-resident and nothing is read or written early.  In a kernel launched the ordinary way both instructions are
// no-ops.  Stream capture turns the attribute into a programmatic dependency edge of the graph.  FSWEEP_PDL=0 launches
// everything with full serialisation (A/B measurements).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace fsweep {

__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("FSWEEP_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace fsweep
