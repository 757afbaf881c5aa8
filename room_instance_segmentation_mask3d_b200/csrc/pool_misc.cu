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

// Eight channels per thread: one 16-byte (bf16) / two 16-byte (fp32) loads per child row instead of four scalar ones, 32-bit index
// arithmetic, 32 bytes of output per thread.  Same summation order per channel (k ascending) => bit-identical to k_pool_fwd.
template <typename TIn>
__global__ void __launch_bounds__(256) k_pool_fwd8(const TIn* __restrict__ in, unsigned CV, const int32_t* __restrict__ nbr, int K,
                                                   unsigned n_out, int average, float* __restrict__ out) {
  const unsigned C = CV * 8u;
  const unsigned total = n_out * CV;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned o = i / CV;
    const unsigned c = (i - o * CV) * 8u;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int cnt = 0;
    for (int k = 0; k < K; ++k) {
      const int r = nbr[(size_t)k * n_out + o];
      if (r < 0) continue;
      ++cnt;
      float v[8];
      const TIn* src = in + (size_t)r * C + c;
      if constexpr (sizeof(TIn) == 2) {
        const uint4 q = *reinterpret_cast<const uint4*>(src);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
          v[2 * j] = f.x;
          v[2 * j + 1] = f.y;
        }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    if (average) {
      const float d = (float)(cnt > 0 ? cnt : 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = acc[j] / d;
    }
    float* dst = out + (size_t)o * C + c;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
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
  // four elements per thread (16-byte load, 4-byte store); same expression per element as the scalar tail below
  const int64_t nvec = ((((uintptr_t)x & 15) | ((uintptr_t)blocked & 3)) == 0) ? n >> 2 : 0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    const float4 f = reinterpret_cast<const float4*>(x)[v];
    const float e[4] = {f.x, f.y, f.z, f.w};
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = 1.f / (1.f + expf(-e[j]));
      w |= (s < 0.5f ? 1u : 0u) << (8 * j);
    }
    reinterpret_cast<uint32_t*>(blocked)[v] = w;
  }
  for (int64_t i = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    // reference: sigmoid(x) < 0.5 in fp32 (mask3d.py:550).  sigmoid is monotone with sigmoid(0) = 0.5 exactly, and
    // fp32 sigmoid(x) rounds to exactly 0.5 only for |x| below ~6e-8 where 1/(1+exp(-x)) is within half an ulp of 0.5;
    // evaluate the same expression so the tie region matches bit for bit.
    const float s = 1.f / (1.f + expf(-x[i]));
    blocked[i] = s < 0.5f ? 1 : 0;
  }
}

// Compact inference result of one decoder output (trainer/trainer.py:373-402 needs exactly this of the mask logits): per voxel the
// binary masks of all Q <= 32 queries as one bit word (bit q = logit > 0 <=> sigmoid > 0.5), per (scene, query) the mask score
// sum_n sigmoid(x) [x > 0] / (sum_n [x > 0] + 1e-6).  One pass over the [N, Q] fp32 logits (4 Q bytes read, 4 bytes written
// per voxel instead of 4 Q bytes copied to the host).  Block = 256 rows of one scene segment; lane = row; the per-query sums are
// reduced by shuffles inside the warp, through shared memory across the 8 warps, and merged per block in a fixed order by
// k_mask_compact_final (deterministic: no floating-point atomics).
__global__ void __launch_bounds__(256) k_mask_compact(const float* __restrict__ logits, int64_t n, int Q,
                                                      const int32_t* __restrict__ scene_offsets, int B, uint32_t* __restrict__ bits,
                                                      float* __restrict__ part /* [blocks][2][32] */, int32_t* __restrict__ block_scene) {
  __shared__ float s_sum[8][2][32];
  // blocks are laid out per scene: block -> (scene, chunk) by scanning the (tiny) offsets table
  int blk = blockIdx.x, b = 0;
  while (b < B) {
    const int nb = (scene_offsets[b + 1] - scene_offsets[b] + 255) / 256;
    if (blk < nb) break;
    blk -= nb;
    ++b;
  }
  if (b >= B) {
    if (threadIdx.x == 0) block_scene[blockIdx.x] = -1;
    return;
  }
  const int64_t row = (int64_t)scene_offsets[b] + (int64_t)blk * 256 + threadIdx.x;
  const bool in = row < scene_offsets[b + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t w = 0;
  float hs = 0.f, cs = 0.f;  // this lane's query column sums after the transpose-reduce below
  // every lane walks the Q columns of its own row; column q's partial sums are reduced across the warp with one xor-butterfly
  for (int q = 0; q < Q; ++q) {
    float h = 0.f, c = 0.f;
    if (in) {
      const float x = logits[row * Q + q];
      if (x > 0.f) {
        w |= 1u << q;
        h = 1.f / (1.f + __expf(-x));
        c = 1.f;
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      h += __shfl_xor_sync(0xffffffffu, h, d);
      c += __shfl_xor_sync(0xffffffffu, c, d);
    }
    if (lane == q) {
      hs = h;
      cs = c;
    }
  }
  if (in) bits[row] = w;
  s_sum[warp][0][lane] = hs;
  s_sum[warp][1][lane] = cs;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int k = threadIdx.x >> 5, q = threadIdx.x & 31;
    float a = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) a += s_sum[ww][k][q];
    part[((int64_t)blockIdx.x * 2 + k) * 32 + q] = a;
  }
  if (threadIdx.x == 0) block_scene[blockIdx.x] = b;
}

// score[b, q] = sum of the scene's block partials: the scene's blocks are the contiguous range k_mask_compact laid out for it; 32
// warps sum interleaved sub-sequences (fixed assignment), warp 0 adds the 32 partial sums in order -- deterministic, and one
// scene no longer walks every block of the batch serially (the first version took 0.49 ms for ONE 889 k-voxel scene)
__global__ void __launch_bounds__(1024) k_mask_compact_final(const float* __restrict__ part, const int32_t* __restrict__ scene_offsets, int B,
                                                            int Q, float* __restrict__ score /* [B][Q] */) {
  __shared__ float s_h[32][33], s_c[32][33];
  const int b = blockIdx.x, q = threadIdx.x & 31, w = threadIdx.x >> 5;
  int first = 0;
  for (int i = 0; i < b; ++i) first += (scene_offsets[i + 1] - scene_offsets[i] + 255) / 256;
  const int nb = (scene_offsets[b + 1] - scene_offsets[b] + 255) / 256;
  float h = 0.f, c = 0.f;
  for (int i = w; i < nb; i += 32) {
    h += part[((int64_t)(first + i) * 2 + 0) * 32 + q];
    c += part[((int64_t)(first + i) * 2 + 1) * 32 + q];
  }
  s_h[w][q] = h;
  s_c[w][q] = c;
  __syncthreads();
  if (w == 0) {
    h = 0.f;
    c = 0.f;
    for (int i = 0; i < 32; ++i) {
      h += s_h[i][q];
      c += s_c[i][q];
    }
    if (q < Q) score[b * Q + q] = h / (c + 1e-6f);
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
  if (C % 8 == 0 && n_out * (int64_t)(C / 8) < ((int64_t)1 << 31) && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
    const unsigned g = grid_for(n_out * (C / 8));
    if (in_dtype == MK_BF16)
      k_pool_fwd8<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)in, (unsigned)(C / 8), nbr, K, (unsigned)n_out, average, out);
    else
      k_pool_fwd8<float><<<g, 256, 0, (cudaStream_t)stream_>>>((const float*)in, (unsigned)(C / 8), nbr, K, (unsigned)n_out, average, out);
    return check_launch(__func__);
  }
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
  k_threshold<<<grid_for((n * Q + 3) / 4), 256, 0, (cudaStream_t)stream_>>>(pooled, n * Q, blocked);
  return check_launch(__func__);
}

int64_t mk_mask_compact_workspace_bytes(int64_t n, int B) {
  const int64_t nblocks = (n + 255) / 256 + B;
  return nblocks * (2 * 32 * (int64_t)sizeof(float) + (int64_t)sizeof(int32_t)) + 256;
}

int mk_mask_compact(const float* logits, int64_t n, int Q, const int32_t* scene_offsets, int B, uint32_t* bits, float* score, void* workspace,
                    int64_t workspace_bytes, void* stream_) {
  MK_REQUIRE(Q >= 1 && Q <= 32, "compact results pack one bit per query into a 32-bit word: Q <= 32");
  MK_REQUIRE(workspace_bytes >= mk_mask_compact_workspace_bytes(n, B), "workspace too small");
  if (B <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  const int nblocks = (int)((n + 255) / 256 + B);  // upper bound of sum_b ceil(n_b / 256)
  float* part = (float*)workspace;
  int32_t* block_scene = (int32_t*)(part + (int64_t)nblocks * 64);
  k_mask_compact<<<nblocks, 256, 0, st>>>(logits, n, Q, scene_offsets, B, bits, part, block_scene);
  k_mask_compact_final<<<B, 1024, 0, st>>>(part, scene_offsets, B, Q, score);
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
