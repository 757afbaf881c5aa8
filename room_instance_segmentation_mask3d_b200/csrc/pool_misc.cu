// Local average / sum pooling over the 2^3 stride-2 kernel map, casts and channel-slice copies.
//
// Replaces MinkowskiAvgPooling's cuSPARSE csrmm + per-row nnz divide with one
// gather pass over the output-stationary neighbour table: each thread owns 4 channels of one output row,
// reads <= K input rows, writes once.  HBM-bound, no atomics.
#include "common.cuh"

namespace mk {

template <typename TIn>
__global__ void __launch_bounds__(256) k_pool_fwd(const TIn* __restrict__ in, int C, const int32_t* __restrict__ nbr, int K,
                                                  int64_t n_out, int average, float* __restrict__ out) {
  const int CV = (C + 3) >> 2;
  const int64_t total = n_out * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / CV;
    const int c = (int)(i - o * CV) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    for (int k = 0; k < K; ++k) {
      const int r = nbr[(int64_t)k * n_out + o];
      if (r < 0) continue;
      ++cnt;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < C) acc[j] += to_f32<TIn>(in[(int64_t)r * C + c + j]);
    }
    const float inv = (average && cnt > 0) ? 1.f / (float)cnt : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c + j < C) out[o * C + c + j] = average ? acc[j] / (float)(cnt > 0 ? cnt : 1) : acc[j];
    (void)inv;
  }
}

// din[child] = dout[parent] / cnt(parent): every child has exactly one parent in a stride-2 2^3 map, so this is a
// scatter with unique targets (no atomics).  Rows of `din` without a parent entry are zeroed by the caller.
__global__ void __launch_bounds__(256) k_pool_bwd(const float* __restrict__ dout, int C, const int32_t* __restrict__ nbr, int K,
                                                  int64_t n_out, int average, float* __restrict__ din) {
  const int64_t total = n_out * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = i / C;
    const int c = (int)(i - o * C);
    int cnt = 0;
    for (int k = 0; k < K; ++k) cnt += nbr[(int64_t)k * n_out + o] >= 0;
    const float g = dout[i] / (float)((average && cnt > 0) ? cnt : 1);
    for (int k = 0; k < K; ++k) {
      const int r = nbr[(int64_t)k * n_out + o];
      if (r >= 0) din[(int64_t)r * C + c] = g;
    }
  }
}

__global__ void __launch_bounds__(256) k_threshold(const float* __restrict__ x, int64_t n, uint8_t* __restrict__ blocked) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // reference: sigmoid(x) < 0.5 in fp32 (mask3d.py:550).  sigmoid is monotone with sigmoid(0) = 0.5 exactly, and
    // fp32 sigmoid(x) rounds to exactly 0.5 only for |x| below ~6e-8 where 1/(1+exp(-x)) is within half an ulp of 0.5;
    // evaluate the same expression so the tie region matches bit for bit.
    const float s = 1.f / (1.f + expf(-x[i]));
    blocked[i] = s < 0.5f ? 1 : 0;
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) k_cast(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = from_f32<TD>(to_f32<TS>(s[i]));
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) k_copy_channels(const TS* __restrict__ s, int C, TD* __restrict__ d, int C_dst, int c_off,
                                                       int64_t n) {
  const int64_t total = n * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    d[r * C_dst + c_off + c] = from_f32<TD>(to_f32<TS>(s[i]));
  }
}

inline unsigned grid_for(int64_t total, int threads = 256) {
  int64_t b = ceil_div(total, threads);
  int64_t cap = 8 * (int64_t)num_sms();
  return (unsigned)(b > cap ? cap : (b > 0 ? b : 1));
}

}  // namespace mk

using namespace mk;

extern "C" {

int mk_pool_fwd(const float* in, int C, const int32_t* nbr, int K, int64_t n_out, int average, float* out, void* stream_) {
  if (n_out == 0) return 0;
  k_pool_fwd<float><<<grid_for(n_out * ((C + 3) / 4)), 256, 0, (cudaStream_t)stream_>>>(in, C, nbr, K, n_out, average, out);
  return check_launch(__func__);
}

int mk_pool_fwd_dt(const void* in, int in_dtype, int C, const int32_t* nbr, int K, int64_t n_out, int average, float* out, void* stream_) {
  if (n_out == 0) return 0;
  if (in_dtype == MK_BF16)
    k_pool_fwd<__nv_bfloat16><<<grid_for(n_out * ((C + 3) / 4)), 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)in, C, nbr, K, n_out,
                                                                                                    average, out);
  else
    k_pool_fwd<float><<<grid_for(n_out * ((C + 3) / 4)), 256, 0, (cudaStream_t)stream_>>>((const float*)in, C, nbr, K, n_out, average, out);
  return check_launch(__func__);
}

int mk_pool_bwd(const float* dout, int C, const int32_t* nbr, int K, int64_t n_out, int average, int64_t n_in, float* din,
                void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  cudaMemsetAsync(din, 0, (size_t)n_in * C * sizeof(float), st);
  if (n_out == 0) return check_launch(__func__);
  k_pool_bwd<<<grid_for(n_out * C), 256, 0, st>>>(dout, C, nbr, K, n_out, average, din);
  return check_launch(__func__);
}

int mk_pool_threshold(const float* pooled, int64_t n, int Q, uint8_t* blocked, void* stream_) {
  if (n == 0) return 0;
  k_threshold<<<grid_for(n * Q), 256, 0, (cudaStream_t)stream_>>>(pooled, n * Q, blocked);
  return check_launch(__func__);
}

int mk_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n == 0) return 0;
  unsigned g = grid_for(n);
  if (src_dtype == MK_F32 && dst_dtype == MK_BF16) k_cast<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == MK_BF16 && dst_dtype == MK_F32) k_cast<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == MK_F32 && dst_dtype == MK_F32) k_cast<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else k_cast<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  return check_launch(__func__);
}

int mk_copy_channels(const void* src, int src_dtype, int C, void* dst, int dst_dtype, int C_dst, int c_off, int64_t n,
                     void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (n == 0) return 0;
  unsigned g = grid_for(n * C);
#define MK_CC(TS, TD) k_copy_channels<TS, TD><<<g, 256, 0, st>>>((const TS*)src, C, (TD*)dst, C_dst, c_off, n)
  if (src_dtype == MK_F32 && dst_dtype == MK_F32) MK_CC(float, float);
  else if (src_dtype == MK_F32 && dst_dtype == MK_BF16) MK_CC(float, __nv_bfloat16);
  else if (src_dtype == MK_BF16 && dst_dtype == MK_F32) MK_CC(__nv_bfloat16, float);
  else MK_CC(__nv_bfloat16, __nv_bfloat16);
#undef MK_CC
  return check_launch(__func__);
}

}  // extern "C"
