// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry point around the REFERENCE's own furthest-point-sampling kernel
// (mask3d/third_party/pointnet2/_ext_src/src/sampling_gpu.cu:72-176), which oracle/make_ref.py compiles from the source where
// it lies under /root/reference into oracle/_ref/libfps_ref.so.  Nothing of the product links or loads this library.
#include <cuda_runtime.h>

namespace at {
namespace cuda {
cudaStream_t& oracle_ref_stream() {
  static cudaStream_t s = nullptr;
  return s;
}
}  // namespace cuda
}  // namespace at

// defined in the reference's sampling_gpu.cu
void furthest_point_sampling_kernel_wrapper(int b, int n, int m, const float* dataset, float* temp, int* idxs);

extern "C" int fps_ref(int b, int n, int m, const float* xyz, float* temp, int* idxs, void* stream) {
  at::cuda::oracle_ref_stream() = (cudaStream_t)stream;
  furthest_point_sampling_kernel_wrapper(b, n, m, xyz, temp, idxs);
  return (int)cudaGetLastError();
}
