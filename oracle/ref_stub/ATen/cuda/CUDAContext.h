// ORACLE BUILD STUB (test infrastructure), see ../ATen.h: the "current stream" is whatever fps_ref() was called with.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
using std::max;
using std::min;
namespace at {
namespace cuda {
cudaStream_t& oracle_ref_stream();
inline cudaStream_t getCurrentCUDAStream() { return oracle_ref_stream(); }
}  // namespace cuda
}  // namespace at
