// fp32 CUDA-core sparse convolution: the exact-fp32 parity path, the Cin=1 stem, and the fp32 backward.
//
// Replaces ME ConvolutionForward/BackwardKernelGPU (convolution_kernel.cu) for the
// "fp32-accumulate parity mode" of BASELINE.json; the throughput path is conv_umma.cu.  Output-stationary:
// one warp owns one output row and walks the K offsets of the neighbour table, so there are no atomics
// and the summation order over k is fixed (k ascending, then ci ascending) => bit-reproducible.
#include "common.cuh"

namespace mk {

constexpr int kMaxCoutPerLane = 12;  // cout <= 384 (dgrad of the 384-channel concat inputs)

// out[o, c] = sum_k sum_ci in[nbr[k,o], ci] * W[widx(k), ci, c]
// `TRANSPOSE_W`: read W as [cout_w, cin_w] transposed, i.e. use W[k][c][ci] (dgrad: reduce over cout).
template <bool TRANSPOSE_W>
__global__ void __launch_bounds__(256) k_conv_f32(const float* __restrict__ in, int cin, const float* __restrict__ W,
                                                  int cout, const int32_t* __restrict__ nbr, int K, int64_t n_out,
                                                  const int32_t* __restrict__ kidx, const float* __restrict__ bias,
                                                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nper = (cout + 31) >> 5;
  for (int64_t o = warp; o < n_out; o += nwarps) {
    float acc[kMaxCoutPerLane];
#pragma unroll
    for (int j = 0; j < kMaxCoutPerLane; ++j) acc[j] = 0.f;
    for (int k = 0; k < K; ++k) {
      const int i = nbr[(int64_t)k * n_out + o];
      if (i < 0) continue;
      const int kw = kidx ? kidx[k] : k;
      const float* xrow = in + (int64_t)i * cin;
      const float* Wk = W + (int64_t)kw * cin * cout;
      for (int ci0 = 0; ci0 < cin; ci0 += 32) {
        const float xv = (ci0 + lane < cin) ? xrow[ci0 + lane] : 0.f;
        const int lim = min(32, cin - ci0);
        for (int t = 0; t < lim; ++t) {
          const float x = __shfl_sync(0xffffffffu, xv, t);
          const int ci = ci0 + t;
#pragma unroll
          for (int j = 0; j < kMaxCoutPerLane; ++j) {
            if (j < nper) {
              const int c = lane + 32 * j;
              if (c < cout) {
                const float w = TRANSPOSE_W ? Wk[(int64_t)c * cin + ci] : Wk[(int64_t)ci * cout + c];
                acc[j] = fmaf(x, w, acc[j]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxCoutPerLane; ++j) {
      if (j < nper) {
        const int c = lane + 32 * j;
        if (c < cout) out[o * cout + c] = acc[j] + (bias ? bias[c] : 0.f);
      }
    }
  }
}

// Stem: cin == 1.  out[o, c] = sum_k in[nbr[k,o]] * W[k, c]; one warp per row, lanes over channels.
template <typename TOut>
__global__ void __launch_bounds__(256) k_conv_stem(const float* __restrict__ in, const float* __restrict__ W, int cout,
                                                   const int32_t* __restrict__ nbr, int K, int64_t n_out,
                                                   TOut* __restrict__ out) {
  extern __shared__ float sW[];  // [K, cout]
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) sW[i] = W[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = warp; o < n_out; o += nwarps) {
    float acc[kMaxCoutPerLane];
#pragma unroll
    for (int j = 0; j < kMaxCoutPerLane; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int kk = k0 + lane;
      const int idx = kk < K ? nbr[(int64_t)kk * n_out + o] : -1;
      const float xv = idx >= 0 ? in[idx] : 0.f;
      unsigned present = __ballot_sync(0xffffffffu, idx >= 0);
      while (present) {
        const int t = __ffs(present) - 1;
        present &= present - 1;
        const float x = __shfl_sync(0xffffffffu, xv, t);
        const float* w = sW + (k0 + t) * cout;
#pragma unroll
        for (int j = 0; j < kMaxCoutPerLane; ++j) {
          const int c = lane + 32 * j;
          if (c < cout) acc[j] = fmaf(x, w[c], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxCoutPerLane; ++j) {
      const int c = lane + 32 * j;
      if (c < cout) out[o * cout + c] = from_f32<TOut>(acc[j]);
    }
  }
}

// Stem, row-per-thread form (cout <= CO): thread = output row, so the K table reads of a warp are 128-byte coalesced
// (nbr is [K, n_out]) and the weights are shared-memory broadcasts; the one-warp-per-row form above reads 32 different
// table lines per load.  Same summation order (k ascending) => same bits.
template <typename TOut, int CO>
__global__ void __launch_bounds__(256) k_conv_stem_rows(const float* __restrict__ in, const float* __restrict__ W, int cout,
                                                        const int32_t* __restrict__ nbr, int K, int64_t n_out,
                                                        TOut* __restrict__ out) {
  extern __shared__ __align__(16) float sW[];  // [K, CO], zero padded
  for (int i = threadIdx.x; i < K * CO; i += blockDim.x) {
    const int k = i / CO, c = i - k * CO;
    sW[i] = c < cout ? W[k * cout + c] : 0.f;
  }
  __syncthreads();
  for (int64_t o0 = (int64_t)blockIdx.x * blockDim.x; o0 < n_out; o0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = o0 + threadIdx.x;
    const bool valid = o < n_out;
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    int idx_next = valid ? nbr[o] : -1;
    for (int k = 0; k < K; ++k) {
      const int idx = idx_next;
      idx_next = (valid && k + 1 < K) ? nbr[(int64_t)(k + 1) * n_out + o] : -1;
      if (idx >= 0) {
        const float xv = in[idx];
        const float4* w = reinterpret_cast<const float4*>(sW + k * CO);
#pragma unroll
        for (int c4 = 0; c4 < CO / 4; ++c4) {
          const float4 wv = w[c4];
          acc[4 * c4 + 0] = fmaf(xv, wv.x, acc[4 * c4 + 0]);
          acc[4 * c4 + 1] = fmaf(xv, wv.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(xv, wv.z, acc[4 * c4 + 2]);
          acc[4 * c4 + 3] = fmaf(xv, wv.w, acc[4 * c4 + 3]);
        }
      }
    }
    if (valid) {
      TOut* dst = out + o * cout;
      const bool sectors = ((((uintptr_t)out) | (uintptr_t)(cout * sizeof(TOut))) & 31) == 0;
      if ((cout & 15) == 0) {
#pragma unroll
        for (int c0 = 0; c0 < CO; c0 += 16) {
          if (c0 < cout) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = acc[c0 + j];
            if (sectors) store16_sectors(dst + c0, v);  // thread = row: a whole 32-byte sector per store instead of two halves
            else store16(dst + c0, v);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < CO; ++c)
          if (c < cout) dst[c] = from_f32<TOut>(acc[c]);
      }
    }
  }
}

// Stem weight gradient (cin == 1, cout <= 32, K <= 128): dW[k, c] = sum_o in[nbr[k,o]] * dout[o, c].
// Block = 8 warps = 2 row tiles (32 rows each) x 4 offset groups (32 offsets each).  A warp keeps 32 accumulators
// (its offsets x lane = channel) in registers over all its row tiles; the dout tile sits in shared memory; table reads
// are coalesced (lane = row).  Partials per (block, row-tile slot) go to a workspace and are merged in fixed order by
// k_conv_stem_wgrad_merge => deterministic.
__global__ void __launch_bounds__(256) k_conv_stem_wgrad(const float* __restrict__ in, const float* __restrict__ dout, int cout,
                                                         const int32_t* __restrict__ nbr, int K, int64_t n_out,
                                                         float* __restrict__ partial) {
  __shared__ float dys[2][32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kg = warp & 3, rt = warp >> 2;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  const int64_t ntiles = (n_out + 31) >> 5;
  for (int64_t t0 = (int64_t)blockIdx.x * 2; t0 < ntiles; t0 += (int64_t)gridDim.x * 2) {
    const int64_t tile = t0 + rt;
    const int64_t o0 = tile * 32;
    __syncthreads();
    // the 4 warps of a row tile load 8 rows each
    for (int r = kg * 8; r < kg * 8 + 8; ++r) {
      const int64_t o = o0 + r;
      dys[rt][r][lane] = (tile < ntiles && o < n_out && lane < cout) ? dout[o * cout + lane] : 0.f;
    }
    __syncthreads();
    if (tile < ntiles) {
      const int64_t o = o0 + lane;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = kg * 32 + j;
        if (k < K) {  // warp-uniform
          const int idx = o < n_out ? nbr[(int64_t)k * n_out + o] : -1;
          const float xv = idx >= 0 ? in[idx] : 0.f;
          unsigned present = __ballot_sync(0xffffffffu, idx >= 0);
          while (present) {
            const int t = __ffs(present) - 1;
            present &= present - 1;
            acc[j] = fmaf(__shfl_sync(0xffffffffu, xv, t), dys[rt][t][lane], acc[j]);
          }
        }
      }
    }
  }
  // partial[(block * 2 + rt), k, c]
  float* dst = partial + ((int64_t)blockIdx.x * 2 + rt) * K * cout;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int k = kg * 32 + j;
    if (k < K && lane < cout) dst[k * cout + lane] = acc[j];
  }
}

__global__ void __launch_bounds__(256) k_conv_stem_wgrad_merge(const float* __restrict__ partial, int nparts, int n, float* __restrict__ dW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * n + i];
  dW[i] = s;
}

// wgrad: dW[k, ci, c] = sum_o in[nbr[k,o], ci] * dout[o, c].  One block per (k, ci-tile of 8); threads over c;
// sequential over o => fixed order, deterministic.  Parity/backward-correctness path, not a throughput path.
__global__ void __launch_bounds__(256) k_conv_wgrad_f32(const float* __restrict__ in, int cin, const float* __restrict__ dout,
                                                        int cout, const int32_t* __restrict__ nbr, int64_t n_out,
                                                        float* __restrict__ dW) {
  const int k = blockIdx.y;
  const int ci0 = blockIdx.x * 8;
  const int c = threadIdx.x;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  __shared__ float sx[64][8];
  __shared__ int sidx[64];
  for (int64_t o0 = 0; o0 < n_out; o0 += 64) {
    __syncthreads();
    if (threadIdx.x < 64) {
      int64_t o = o0 + threadIdx.x;
      sidx[threadIdx.x] = o < n_out ? nbr[(int64_t)k * n_out + o] : -1;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 64 * 8; t += blockDim.x) {
      int r = t >> 3, j = t & 7;
      int i = sidx[r];
      sx[r][j] = (i >= 0 && ci0 + j < cin) ? in[(int64_t)i * cin + ci0 + j] : 0.f;
    }
    __syncthreads();
    if (c < cout) {
      const int lim = (int)min((int64_t)64, n_out - o0);
      for (int r = 0; r < lim; ++r) {
        if (sidx[r] < 0) continue;
        const float g = dout[(o0 + r) * cout + c];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(sx[r][j], g, acc[j]);
      }
    }
  }
  if (c < cout) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (ci0 + j < cin) dW[((int64_t)k * cin + ci0 + j) * cout + c] = acc[j];
  }
}

}  // namespace mk

using namespace mk;

extern "C" {

int mk_conv_fwd_f32(const float* in, int cin, const float* W, int cout, const int32_t* nbr, int K, int64_t n_out,
                    const float* bias, float* out, void* stream_) {
  MK_REQUIRE(cout >= 1 && cout <= 32 * kMaxCoutPerLane, "cout must be <= 384");
  if (n_out == 0) return 0;
  int64_t blocks = ceil_div(n_out, 8);
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  k_conv_f32<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(in, cin, W, cout, nbr, K, n_out, nullptr, bias, out);
  return check_launch(__func__);
}

int mk_conv_dgrad_f32(const float* dout, int cout, const float* W, int cin, const int32_t* nbr_t, int K, int64_t n_in,
                      const int32_t* kidx, float* din, void* stream_) {
  // din[i, ci] = sum_k sum_c dout[nbr_t[k,i], c] * W[kidx[k], ci, c]: the forward kernel with W read transposed.
  MK_REQUIRE(cin >= 1 && cin <= 32 * kMaxCoutPerLane, "cin must be <= 384");
  if (n_in == 0) return 0;
  int64_t blocks = ceil_div(n_in, 8);
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  k_conv_f32<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(dout, cout, W, cin, nbr_t, K, n_in, kidx, nullptr, din);
  return check_launch(__func__);
}

int mk_conv_wgrad_f32(const float* in, int cin, const float* dout, int cout, const int32_t* nbr, int K, int64_t n_out,
                      float* dW, void* stream_) {
  MK_REQUIRE(cout <= 256, "cout must be <= 256");
  dim3 grid((unsigned)ceil_div(cin, 8), (unsigned)K);
  k_conv_wgrad_f32<<<grid, 256, 0, (cudaStream_t)stream_>>>(in, cin, dout, cout, nbr, n_out, dW);
  return check_launch(__func__);
}

static int stem_wgrad_grid(int64_t n_out) {
  int64_t g = ceil_div(ceil_div(n_out, 32), 2);
  if (g > 4 * num_sms()) g = 4 * num_sms();
  return (int)(g < 1 ? 1 : g);
}

int64_t mk_conv_stem_wgrad_workspace_bytes(int64_t n_out, int K, int cout) {
  return (int64_t)stem_wgrad_grid(n_out) * 2 * K * cout * (int64_t)sizeof(float);
}

int mk_conv_stem_wgrad_f32(const float* in, const float* dout, int cout, const int32_t* nbr, int K, int64_t n_out, float* dW,
                           void* workspace, int64_t workspace_bytes, void* stream_) {
  MK_REQUIRE(cout >= 1 && cout <= 32 && K >= 1 && K <= 128, "stem wgrad: cout <= 32 and K <= 128 required");
  MK_REQUIRE(workspace_bytes >= mk_conv_stem_wgrad_workspace_bytes(n_out, K, cout), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  if (n_out == 0) {
    cudaMemsetAsync(dW, 0, (size_t)K * cout * sizeof(float), st);
    return check_launch(__func__);
  }
  const int grid = stem_wgrad_grid(n_out);
  k_conv_stem_wgrad<<<grid, 256, 0, st>>>(in, dout, cout, nbr, K, n_out, (float*)workspace);
  k_conv_stem_wgrad_merge<<<(unsigned)ceil_div((int64_t)K * cout, 256), 256, 0, st>>>((const float*)workspace, grid * 2, K * cout, dW);
  return check_launch(__func__);
}

int mk_conv_stem_f32(const float* in, const float* W, int cout, const int32_t* nbr, int K, int64_t n_out, void* out,
                     int out_dtype, void* stream_) {
  MK_REQUIRE(cout <= 256, "cout must be <= 256");
  MK_REQUIRE((size_t)K * cout * sizeof(float) <= 48 * 1024, "stem weights must fit 48 KB of shared memory");
  if (n_out == 0) return 0;
  if (cout <= 32) {
    // row-per-thread form: coalesced table reads (see k_conv_stem_rows)
    int64_t blocks = ceil_div(n_out, 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    const size_t smem32 = (size_t)K * 32 * sizeof(float);
    if (out_dtype == MK_F32)
      k_conv_stem_rows<float, 32><<<(unsigned)blocks, 256, smem32, (cudaStream_t)stream_>>>(in, W, cout, nbr, K, n_out, (float*)out);
    else
      k_conv_stem_rows<__nv_bfloat16, 32><<<(unsigned)blocks, 256, smem32, (cudaStream_t)stream_>>>(in, W, cout, nbr, K, n_out,
                                                                                                (__nv_bfloat16*)out);
    return check_launch(__func__);
  }
  int64_t blocks = ceil_div(n_out, 8);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  size_t smem = (size_t)K * cout * sizeof(float);
  if (out_dtype == MK_F32)
    k_conv_stem<float><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream_>>>(in, W, cout, nbr, K, n_out, (float*)out);
  else
    k_conv_stem<__nv_bfloat16><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream_>>>(in, W, cout, nbr, K, n_out,
                                                                                     (__nv_bfloat16*)out);
  return check_launch(__func__);
}

}  // extern "C"
