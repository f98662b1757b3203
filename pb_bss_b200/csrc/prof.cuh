// Launch accounting + optional CUDA-event timing of the library's own kernels
// (pbb_launch_count / pbb_profile_* in include/pbb.h).  bench.py uses it to time
// the dominant kernel with events on the launching stream.
#pragma once
#include <cuda_runtime.h>

namespace pbb {

void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);

struct LaunchScope {
  cudaStream_t st;
  LaunchScope(const char* name, cudaStream_t s) : st(s) { prof_begin(name, s); }
  ~LaunchScope() { prof_end(st); }
};

}  // namespace pbb
