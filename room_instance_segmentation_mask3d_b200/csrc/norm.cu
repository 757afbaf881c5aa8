// Instance norm (per scene, per channel) + affine + residual + ReLU, and its backward.
//
// Replaces MinkowskiInstanceNorm's 2 global-avg-pool + 2 broadcast launches (4 full passes over [N,C],
// SURVEY.md section 8a-a9) and the separate MinkowskiReLU / `out += residual` passes (a11) with
//   stats:  one read of x   (chunked, shifted sums, deterministic fp64 merge in fixed order)
//   apply:  one read of x (+ residual), one write, fused affine + residual + ReLU + dtype cast.
// HBM-bound: vectorised 4-channel loads, coalesced along channels, grid sized in multiples of the SM count.
#include "common.cuh"

namespace mk {

constexpr int kNormThreads = 256;

// 1-D grid over the NON-EMPTY (scene, chunk) pairs of a batch (host: ceil(n / rows_per_chunk) + B blocks is an upper bound):
// block g -> scene b and chunk g - (chunks of the scenes before b); false past the end.  A (chunks, B) grid sized for the
// largest possible scene launches B - 1 empty blocks per useful one.
__device__ __forceinline__ bool resolve_chunk(const int32_t* __restrict__ offs, int B, int rpc, int g, int& b, int& chunk) {
  for (b = 0; b < B; ++b) {
    const int nb = (offs[b + 1] - offs[b] + rpc - 1) / rpc;
    if (g < nb) {
      chunk = g;
      return true;
    }
    g -= nb;
  }
  return false;
}

// partial layout: [chunk][3][C] = {count (as float), shifted sum, shifted sum of squares}, shift[chunk][C]
template <typename T>
__global__ void __launch_bounds__(kNormThreads) k_in_stats_partial(const T* __restrict__ x, int C,
                                                                   const int32_t* __restrict__ scene_offsets, int rows_per_chunk,
                                                                   int chunks_per_scene, float* __restrict__ partial) {
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t start = (int64_t)scene_offsets[b] + (int64_t)chunk * rows_per_chunk;
  const int64_t end = min((int64_t)scene_offsets[b + 1], start + rows_per_chunk);
  const int CV = C >> 2;                 // 4-channel vectors per row
  const int RL = kNormThreads / CV;      // row lanes
  const int cv = threadIdx.x % CV, rl = threadIdx.x / CV;
  float* out = partial + ((int64_t)(b * chunks_per_scene + chunk)) * 4 * C;
  extern __shared__ float sm[];  // [RL][2][C]
  float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0), shift = make_float4(0, 0, 0, 0);
  if (start < end) shift = load4(x + start * C + 4 * cv);
  if (rl < RL) {
    for (int64_t r = start + rl; r < end; r += RL) {
      float4 v = load4(x + r * C + 4 * cv);
      v.x -= shift.x; v.y -= shift.y; v.z -= shift.z; v.w -= shift.w;
      s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
      s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
    }
    float* p = sm + (int64_t)rl * 2 * C;
    store4(p + 4 * cv, s1);
    store4(p + C + 4 * cv, s2);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < RL; ++l) {  // fixed order
      a += sm[l * 2 * C + c];
      q += sm[l * 2 * C + C + c];
    }
    out[c] = (float)(end > start ? end - start : 0);
    out[C + c] = a;
    out[2 * C + c] = q;
  }
  if (rl == 0) store4(out + 3 * C + 4 * cv, shift);
}

// grid (ceil(C / 8), B), block 256 = 8 warps, one channel per warp.  Chunk statistics are merged with Chan's parallel formula in
// fp64: lane l folds chunks l, l+32, ... in ascending order, then the 32 lanes are folded by a fixed shuffle tree (lane i
// absorbs lane i + d, d = 16 .. 1) => a fixed merge order, bit-reproducible run to run.
__device__ __forceinline__ void chan_merge(double& n, double& mu, double& m2, double nb, double mub, double m2b) {
  if (nb <= 0) return;
  const double nt = n + nb, d = mub - mu, f = nb / nt;
  mu += d * f;
  m2 += m2b + d * d * n * f;
  n = nt;
}

__global__ void __launch_bounds__(256) k_in_stats_final(const float* __restrict__ partial, int C, int chunks_per_scene_cap, float eps,
                                                        float* __restrict__ mean, float* __restrict__ rstd,
                                                        const int32_t* __restrict__ scene_offsets, int rows_per_chunk) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  // chunks are laid out for the largest possible scene (stride chunks_per_scene_cap); only the ones this scene fills are read
  const int scene_rows = scene_offsets[b + 1] - scene_offsets[b];
  if (scene_rows <= 0) {  // no chunk of an empty scene was written
    if (c < C && lane == 0) {
      mean[b * C + c] = 0.f;
      rstd[b * C + c] = (float)(1.0 / sqrt((double)eps));
    }
    return;
  }
  const int chunks_per_scene = min(chunks_per_scene_cap, (scene_rows + rows_per_chunk - 1) / rows_per_chunk);
  if (c >= C) return;  // whole warp
  // Sums about ONE reference value (the first chunk's shift, a sample of the channel): per chunk
  //   sum(x - ref) = s1 + n d,  sum((x - ref)^2) = s2 + 2 d s1 + n d^2,  d = shift_chunk - ref
  // in fp64 with multiply-adds only (the pairwise Chan merge needed an fp64 division per chunk on the critical path of a
  // kernel that is pure latency: ~10 us per call, 62 calls per forward).  Fixed order: lane l folds chunks l, l+32, ...,
  // then a fixed shuffle tree => bit-reproducible run to run.
  const float* p0 = partial + ((int64_t)(b * chunks_per_scene_cap)) * 4 * C;
  const double ref = p0[3 * C + c];
  double n = 0, S = 0, SS = 0;
  // four chunks per lane per round, all 16 loads issued before the first use (the kernel is a chain of L2 latencies)
  constexpr int U = 8;  // chunks per lane and round: all 4 * U loads are issued before the first use
  for (int ch0 = lane; ch0 < chunks_per_scene; ch0 += 32 * U) {
    float v[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int ch = ch0 + 32 * u;
      const float* p = partial + ((int64_t)(b * chunks_per_scene_cap + (ch < chunks_per_scene ? ch : 0))) * 4 * C;
      const bool ok = ch < chunks_per_scene;
      v[u][0] = ok ? p[c] : 0.f;
      v[u][1] = ok ? p[C + c] : 0.f;
      v[u][2] = ok ? p[2 * C + c] : 0.f;
      v[u][3] = ok ? p[3 * C + c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {  // ascending chunk order per lane, as before
      const double nb = v[u][0];
      if (nb <= 0) continue;
      const double s1 = v[u][1], s2 = v[u][2], d = (double)v[u][3] - ref;
      n += nb;
      S += fma(nb, d, s1);
      SS += fma(d, fma(nb, d, 2.0 * s1), s2);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    n += __shfl_down_sync(0xffffffffu, n, d);
    S += __shfl_down_sync(0xffffffffu, S, d);
    SS += __shfl_down_sync(0xffffffffu, SS, d);
  }
  if (lane == 0) {
    double mu = ref, var = 0.0;
    if (n > 0) {
      const double m = S / n;
      mu = ref + m;
      var = SS / n - m * m;
      if (var < 0) var = 0;
    }
    mean[b * C + c] = (float)mu;
    rstd[b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <typename T>
__global__ void __launch_bounds__(kNormThreads) k_in_stats(const T* __restrict__ x, int C, const int32_t* __restrict__ scene_offsets, int B,
                                                           int rows_per_chunk, int chunks_per_scene, float eps,
                                                           float* __restrict__ partial, unsigned* __restrict__ counters,
                                                           float* __restrict__ mean, float* __restrict__ rstd) {
  constexpr int V = 16 / (int)sizeof(T);  // channels per 16-byte load
  int b, chunk;
  if (!resolve_chunk(scene_offsets, B, rows_per_chunk, (int)blockIdx.x, b, chunk)) return;
  const int64_t start = (int64_t)scene_offsets[b] + (int64_t)chunk * rows_per_chunk;
  const int64_t end = min((int64_t)scene_offsets[b + 1], start + rows_per_chunk);
  const int CV = C / V;
  const int RL = kNormThreads / CV;
  const int cv = threadIdx.x % CV, rl = threadIdx.x / CV;
  float* out = partial + ((int64_t)(b * chunks_per_scene + chunk)) * 4 * C;
  extern __shared__ float sm[];  // [RL][2][C]
  __shared__ int s_last;
  float s1[V], s2[V], shift[V];
  auto ld = [&](int64_t r, float (&v)[V]) {
    if constexpr (sizeof(T) == 2) {
      const uint4 q = *reinterpret_cast<const uint4*>(x + r * C + V * cv);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
      }
    } else {
      const float4 q = *reinterpret_cast<const float4*>(x + r * C + V * cv);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
  };
#pragma unroll
  for (int j = 0; j < V; ++j) s1[j] = s2[j] = shift[j] = 0.f;
  if (rl < RL) {
    if (start < end) ld(start, shift);
    int64_t r = start + rl;
    for (; r + 3 * (int64_t)RL < end; r += 4 * (int64_t)RL) {
      float v[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) ld(r + u * (int64_t)RL, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float d = v[u][j] - shift[j];
          s1[j] += d;
          s2[j] = fmaf(d, d, s2[j]);
        }
      }
    }
    for (; r < end; r += RL) {
      float v[V];
      ld(r, v);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float d = v[j] - shift[j];
        s1[j] += d;
        s2[j] = fmaf(d, d, s2[j]);
      }
    }
    float* p = sm + (int64_t)rl * 2 * C + V * cv;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      p[j] = s1[j];
      p[C + j] = s2[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    float a = 0.f, q = 0.f;
    for (int l = 0; l < RL; ++l) {  // fixed order
      a += sm[l * 2 * C + c];
      q += sm[l * 2 * C + C + c];
    }
    out[c] = (float)(end > start ? end - start : 0);
    out[C + c] = a;
    out[2 * C + c] = q;
  }
  if (rl == 0) {
#pragma unroll
    for (int j = 0; j < V; ++j) out[3 * C + V * cv + j] = shift[j];
  }
  if (counters == nullptr) return;  // statistics merged by k_in_stats_final (separate launch)
  // ---- last block of the scene merges the chunk statistics (Chan, fp64, fixed order)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&counters[b], 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31;
  for (int c = threadIdx.x >> 5; c < C; c += kNormThreads / 32) {
    double n = 0, mu = 0, m2 = 0;
    for (int ch = lane; ch < chunks_per_scene; ch += 32) {
      const float* p = partial + ((int64_t)(b * chunks_per_scene + ch)) * 4 * C;
      const double nb = __ldcg(p + c);
      if (nb <= 0) continue;
      const double a1 = __ldcg(p + C + c), a2 = __ldcg(p + 2 * C + c), sh = __ldcg(p + 3 * C + c);
      double m2b = a2 - a1 * a1 / nb;
      if (m2b < 0) m2b = 0;
      chan_merge(n, mu, m2, nb, sh + a1 / nb, m2b);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double nb = __shfl_down_sync(0xffffffffu, n, d), mub = __shfl_down_sync(0xffffffffu, mu, d),
                   m2b = __shfl_down_sync(0xffffffffu, m2, d);
      chan_merge(n, mu, m2, nb, mub, m2b);
    }
    if (lane == 0) {
      const double var = n > 0 ? m2 / n : 0.0;
      mean[b * C + c] = (float)mu;
      rstd[b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

template <typename T, int V>
__device__ __forceinline__ void ld_vec(const T* p, float (&v)[V]) {
  if constexpr (sizeof(T) == 2) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  } else {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
}
template <typename T, int V>
__device__ __forceinline__ void st_vec(T* p, const float (&v)[V]) {
  if constexpr (sizeof(T) == 2) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// bf16 -> bf16 apply, 8 channels (16 bytes) per thread.  The grid-stride step is a multiple of the vectors per row (host: grid
// rounded to a multiple of CV / gcd(CV, 256)), so a thread keeps ONE channel group for its whole life: mean / rstd / gamma / beta
// of that group live in registers (reloaded when the thread's rows cross into the next scene) and an element costs one 16-byte
// load, the arithmetic and one 16-byte store.  The first version re-derived (row, channel) with a division and re-read 128 bytes
// of parameters per 16 bytes of data: ncu had it at 0.38 of the HBM rate (profiles/r02_families_infer_summary.txt).
__global__ void __launch_bounds__(256) k_norm_apply_bf16x8(const __nv_bfloat16* __restrict__ x, unsigned n_rows, unsigned CV,
                                                          const int32_t* __restrict__ scene_offsets, int B,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const __nv_bfloat16* __restrict__ residual, int relu,
                                                          __nv_bfloat16* __restrict__ out) {
  const unsigned C = CV * 8u;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned r = i0 / CV;
  const unsigned c = (i0 - r * CV) * 8u;
  const unsigned r_step = gridDim.x * blockDim.x / CV;  // exact by construction of the grid
  float gg[8], ee[8], mm[8], ss[8];
  {
    const float4 g0 = load4(gamma + c), g1 = load4(gamma + c + 4), e0 = load4(beta + c), e1 = load4(beta + c + 4);
    gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
    ee[0] = e0.x; ee[1] = e0.y; ee[2] = e0.z; ee[3] = e0.w; ee[4] = e1.x; ee[5] = e1.y; ee[6] = e1.z; ee[7] = e1.w;
  }
  int b = -1;
  unsigned next = 0;  // first row of scene b + 1
  for (; r < n_rows; r += r_step) {
    if (r >= next) {  // rows are grouped by scene: a thread crosses at most B - 1 boundaries
      do {
        ++b;
        next = b + 1 < B ? (unsigned)scene_offsets[b + 1] : 0xFFFFFFFFu;
      } while (r >= next);
      const float4 m0 = load4(mean + b * C + c), m1 = load4(mean + b * C + c + 4);
      const float4 s0 = load4(rstd + b * C + c), s1 = load4(rstd + b * C + c + 4);
      mm[0] = m0.x; mm[1] = m0.y; mm[2] = m0.z; mm[3] = m0.w; mm[4] = m1.x; mm[5] = m1.y; mm[6] = m1.z; mm[7] = m1.w;
      ss[0] = s0.x; ss[1] = s0.y; ss[2] = s0.z; ss[3] = s0.w; ss[4] = s1.x; ss[5] = s1.y; ss[6] = s1.z; ss[7] = s1.w;
    }
    const size_t off = (size_t)r * C + c;
    float v[8];
    ld_vec<__nv_bfloat16, 8>(x + off, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf((v[j] - mm[j]) * ss[j], gg[j], ee[j]);
    if (residual) {
      float q[8];
      ld_vec<__nv_bfloat16, 8>(residual + off, q);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += q[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    st_vec<__nv_bfloat16, 8>(out + off, v);
  }
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_norm_apply(const TIn* __restrict__ x, int64_t n, int C,
                                                    const int32_t* __restrict__ scene_offsets, int B,
                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    const TOut* __restrict__ residual, int relu, TOut* __restrict__ out) {
  const int CV = C >> 2;
  const int64_t total = n * CV;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / CV;
    const int c = (int)(i - r * CV) * 4;
    float4 v = load4(x + r * C + c);
    if (mean) {
      const int b = scene_of_row(scene_offsets, B, r);
      const float4 m = load4(mean + b * C + c), s = load4(rstd + b * C + c);
      v.x = (v.x - m.x) * s.x; v.y = (v.y - m.y) * s.y; v.z = (v.z - m.z) * s.z; v.w = (v.w - m.w) * s.w;
    }
    if (gamma) {
      const float4 g = load4(gamma + c), be = load4(beta + c);
      v.x = fmaf(v.x, g.x, be.x); v.y = fmaf(v.y, g.y, be.y); v.z = fmaf(v.z, g.z, be.z); v.w = fmaf(v.w, g.w, be.w);
    }
    if (residual) {
      const float4 q = load4(residual + r * C + c);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    store4(out + r * C + c, v);
  }
}

// ---- backward: per (scene, channel) sums of dy*gamma and dy*gamma*yhat, then dx -------------------
// partial [chunk][2][C]
__global__ void __launch_bounds__(kNormThreads) k_in_bwd_partial(const float* __restrict__ x, const float* __restrict__ dy,
                                                                 int C, const int32_t* __restrict__ scene_offsets,
                                                                 int rows_per_chunk, int chunks_per_scene,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 float* __restrict__ partial) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int64_t start = (int64_t)scene_offsets[b] + (int64_t)chunk * rows_per_chunk;
  const int64_t end = min((int64_t)scene_offsets[b + 1], start + rows_per_chunk);
  float* out = partial + ((int64_t)(b * chunks_per_scene + chunk)) * 2 * C;
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    const float m = mean[b * C + c], s = rstd[b * C + c];
    float a = 0.f, q = 0.f;
    for (int64_t r = start; r < end; ++r) {
      const float g = dy[r * C + c];
      const float yh = (x[r * C + c] - m) * s;
      a += g;
      q = fmaf(g, yh, q);
    }
    out[c] = a;
    out[C + c] = q;
  }
}

__global__ void k_in_bwd_final(const float* __restrict__ partial, int C, int chunks_per_scene, int B,
                               const int32_t* __restrict__ scene_offsets, float* __restrict__ sum_dy,
                               float* __restrict__ sum_dyy, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  // one block; thread per channel; deterministic order
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double tg = 0, tb = 0;
    for (int b = 0; b < B; ++b) {
      double a = 0, q = 0;
      for (int ch = 0; ch < chunks_per_scene; ++ch) {
        const float* p = partial + ((int64_t)(b * chunks_per_scene + ch)) * 2 * C;
        a += p[c];
        q += p[C + c];
      }
      sum_dy[b * C + c] = (float)a;
      sum_dyy[b * C + c] = (float)q;
      tb += a;
      tg += q;
    }
    if (dgamma) dgamma[c] = (float)tg;
    if (dbeta) dbeta[c] = (float)tb;
  }
}

__global__ void __launch_bounds__(256) k_in_bwd_dx(const float* __restrict__ x, const float* __restrict__ dy, int64_t n, int C,
                                                   const int32_t* __restrict__ scene_offsets, int B,
                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                   const float* __restrict__ gamma, const float* __restrict__ sum_dy,
                                                   const float* __restrict__ sum_dyy, float* __restrict__ dx) {
  const int64_t total = n * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const int b = scene_of_row(scene_offsets, B, r);
    const float nb = (float)(scene_offsets[b + 1] - scene_offsets[b]);
    const float s = rstd[b * C + c];
    const float yh = (x[i] - mean[b * C + c]) * s;
    const float g = gamma ? gamma[c] : 1.f;
    // dy here is the gradient w.r.t. the affine output; dgamma/dbeta sums were taken on dy directly
    dx[i] = g * s * (dy[i] - sum_dy[b * C + c] / nb - yh * sum_dyy[b * C + c] / nb);
  }
}

// ---- fused backward of (instance norm -> [+ residual] -> [ReLU]): one statistics pass + one apply pass, any dtype ----------
// g = dy where the unit's output was positive (or dy when there is no ReLU); x_hat = (x - mean) * rstd with x the conv output.
//   sums per (scene, channel):  A = sum g,  Q = sum g * x_hat            (also dbeta = sum_b A, dgamma = sum_b Q)
//   dx = gamma * rstd * (g - A / n_b - x_hat * Q / n_b),   dres = g
// partial [chunk][2][C] = {A, Q}; grid (chunks, B)
template <typename T>
__global__ void __launch_bounds__(kNormThreads) k_inr_bwd_partial(const T* __restrict__ dy, const T* __restrict__ out,
                                                                  const T* __restrict__ x, int C,
                                                                  const int32_t* __restrict__ scene_offsets, int B, int rows_per_chunk,
                                                                  int chunks_per_scene, const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, float* __restrict__ partial) {
  constexpr int V = 16 / (int)sizeof(T);
  int b, chunk;
  if (!resolve_chunk(scene_offsets, B, rows_per_chunk, (int)blockIdx.x, b, chunk)) return;
  const int64_t start = (int64_t)scene_offsets[b] + (int64_t)chunk * rows_per_chunk;
  const int64_t end = min((int64_t)scene_offsets[b + 1], start + rows_per_chunk);
  const int CV = C / V;
  const int RL = kNormThreads / CV;
  const int cv = threadIdx.x % CV, rl = threadIdx.x / CV;
  float* po = partial + ((int64_t)(b * chunks_per_scene + chunk)) * 2 * C;
  extern __shared__ float sm[];  // [RL][2][C]
  float a[V], q[V], mu[V], rs[V];
#pragma unroll
  for (int j = 0; j < V; ++j) a[j] = q[j] = 0.f;
  if (rl < RL) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      mu[j] = mean[b * C + V * cv + j];
      rs[j] = rstd[b * C + V * cv + j];
    }
    auto acc = [&](int64_t r) {
      float g[V], xv[V];
      ld_vec<T, V>(dy + r * C + V * cv, g);
      ld_vec<T, V>(x + r * C + V * cv, xv);
      if (out) {
        float o[V];
        ld_vec<T, V>(out + r * C + V * cv, o);
#pragma unroll
        for (int j = 0; j < V; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        a[j] += g[j];
        q[j] = fmaf(g[j], (xv[j] - mu[j]) * rs[j], q[j]);
      }
    };
    int64_t r = start + rl;
    for (; r + (int64_t)RL < end; r += 2 * (int64_t)RL) {  // two rows (six loads) in flight
      acc(r);
      acc(r + RL);
    }
    if (r < end) acc(r);
    float* p = sm + (int64_t)rl * 2 * C + V * cv;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      p[j] = a[j];
      p[C + j] = q[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kNormThreads) {
    float sa = 0.f, sq = 0.f;
    for (int l = 0; l < RL; ++l) {  // fixed order
      sa += sm[l * 2 * C + c];
      sq += sm[l * 2 * C + C + c];
    }
    po[c] = sa;
    po[C + c] = sq;
  }
}

// grid ceil(C / 8), block 256: one warp per channel, scenes in sequence (B is small), lanes over chunks, fixed merge order.
__global__ void __launch_bounds__(256) k_inr_bwd_final(const float* __restrict__ partial, int C, int chunks_per_scene_cap, int B,
                                                       float* __restrict__ sum_a, float* __restrict__ sum_q,
                                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                       const int32_t* __restrict__ scene_offsets, int rows_per_chunk) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double ta = 0, tq = 0;
  for (int b = 0; b < B; ++b) {
    double sa = 0, sq = 0;
    const int chunks_per_scene = min(chunks_per_scene_cap, (scene_offsets[b + 1] - scene_offsets[b] + rows_per_chunk - 1) / rows_per_chunk);
    constexpr int U = 8;
    for (int ch0 = lane; ch0 < chunks_per_scene; ch0 += 32 * U) {
      float va[U], vq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int ch = ch0 + 32 * u;
        const bool ok = ch < chunks_per_scene;
        const float* p = partial + ((int64_t)(b * chunks_per_scene_cap + (ok ? ch : 0))) * 2 * C;
        va[u] = ok ? p[c] : 0.f;
        vq[u] = ok ? p[C + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        sa += va[u];
        sq += vq[u];
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sa += __shfl_down_sync(0xffffffffu, sa, d);
      sq += __shfl_down_sync(0xffffffffu, sq, d);
    }
    if (lane == 0) {
      sum_a[b * C + c] = (float)sa;
      sum_q[b * C + c] = (float)sq;
    }
    ta += sa;
    tq += sq;
  }
  if (lane == 0) {
    if (dbeta) dbeta[c] = (float)ta;
    if (dgamma) dgamma[c] = (float)tq;
  }
}

template <typename T, typename TD>
__global__ void __launch_bounds__(256) k_inr_bwd_apply(const T* __restrict__ dy, const T* __restrict__ out, const T* __restrict__ x,
                                                       int64_t n, int C, const int32_t* __restrict__ scene_offsets, int B,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ sum_a,
                                                       const float* __restrict__ sum_q, TD* __restrict__ dx, T* __restrict__ dres) {
  // same thread-keeps-its-channel-group structure as k_norm_apply_bf16x8 (grid from row_invariant_grid): per-scene constants
  //   d = k1 * g - k0 - k2 * x   with  k1 = gamma rstd,  k2 = gamma rstd^2 sum_q / n,  k0 = gamma rstd sum_a / n - k2 mean
  // would change the rounding of the first version, so the original expression is kept and only its operands are cached
  constexpr int V = 16 / (int)sizeof(T);
  const int CV = C / V;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t r = i0 / CV;
  const int c = (int)(i0 - r * CV) * V;
  const int64_t r_step = (int64_t)gridDim.x * blockDim.x / CV;
  float gm[V], rs[V], mu[V], sa[V], sq[V];
#pragma unroll
  for (int j = 0; j < V; ++j) gm[j] = gamma ? gamma[c + j] : 1.f;
  int b = -1;
  int64_t next = 0;
  float inv_n = 0.f;
  for (; r < n; r += r_step) {
    if (r >= next) {
      do {
        ++b;
        next = b + 1 < B ? (int64_t)scene_offsets[b + 1] : (int64_t)1 << 62;
      } while (r >= next);
      inv_n = 1.f / (float)(scene_offsets[b + 1] - scene_offsets[b]);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        rs[j] = rstd[b * C + c + j];
        mu[j] = mean[b * C + c + j];
        sa[j] = sum_a[b * C + c + j];
        sq[j] = sum_q[b * C + c + j];
      }
    }
    const int64_t off = r * C + c;
    float g[V], xv[V], d[V];
    ld_vec<T, V>(dy + off, g);
    ld_vec<T, V>(x + off, xv);
    if (out) {
      float o[V];
      ld_vec<T, V>(out + off, o);
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xh = (xv[j] - mu[j]) * rs[j];
      d[j] = gm[j] * rs[j] * (g[j] - sa[j] * inv_n - xh * sq[j] * inv_n);
    }
    if constexpr (sizeof(TD) == sizeof(T)) {
      st_vec<TD, V>(dx + off, d);
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) dx[off + j] = from_f32<TD>(d[j]);
    }
    if (dres) st_vec<T, V>(dres + off, g);
  }
}

// grid of 256-thread blocks whose total thread count is a multiple of `cv` (vectors per row): a grid-stride thread then keeps its
// channel group, see k_norm_apply_bf16x8
inline int64_t row_invariant_grid(int64_t total_vecs, int cv) {
  int g = cv, t = 256;
  while (t) {  // gcd(cv, 256)
    const int q = g % t;
    g = t;
    t = q;
  }
  const int64_t m = cv / g;
  int64_t blocks = ceil_div(total_vecs, 256);
  const int64_t cap = 16 * (int64_t)num_sms();
  if (blocks > cap) blocks = cap;
  return ceil_div(blocks, m) * m;
}

inline void norm_chunking(int64_t n, int B, int* rows_per_chunk, int* chunks_per_scene) {
  // Chunks are per scene, sized for the largest possible scene (all n rows): in a batch of B similar scenes only 1 / B of the
  // (scene, chunk) blocks have rows, so the target is 8 per SM -- with 2 per SM the statistics passes ran 2 blocks of 256
  // threads per SM, 0.39 of the HBM rate (profiles/r02_families_infer_summary.txt)
  int64_t target = 8 * (int64_t)num_sms();
  int64_t rpc = ceil_div(n > 0 ? n : 1, target);
  if (rpc < 64) rpc = 64;
  *rows_per_chunk = (int)rpc;
  *chunks_per_scene = (int)ceil_div(n > 0 ? n : 1, rpc);
  (void)B;
}

}  // namespace mk

using namespace mk;

extern "C" {

int64_t mk_instnorm_workspace_bytes(int64_t n, int C, int B) {
  int rpc, cps;
  norm_chunking(n, B, &rpc, &cps);
  return ((int64_t)B * cps * 4 * C + 2 * (int64_t)B * C + 64 + (int64_t)B) * (int64_t)sizeof(float);
}

int mk_instnorm_stats(const void* x, int x_dtype, int64_t n, int C, const int32_t* scene_offsets, int B, float eps,
                      float* mean, float* rstd, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(C % 4 == 0 && C >= 4 && C <= 4 * kNormThreads, "C must be a multiple of 4, <= 1024");
  MK_REQUIRE(workspace_bytes >= mk_instnorm_workspace_bytes(n, C, B), "workspace too small");
  MK_REQUIRE(B >= 1, "B must be >= 1");
  int rpc, cps;
  norm_chunking(n, B, &rpc, &cps);
  float* partial = (float*)workspace;
  // The in-kernel merge by the last block (pass `counters`) saves a launch but serialises C channels x chunks of fp64 merge
  // on ONE SM (measured 0.115 ms per call on B200 vs ~0.01 ms for the 12-block k_in_stats_final), so it is off.
  unsigned* counters = nullptr;
  dim3 grid((unsigned)cps, (unsigned)B);
  const unsigned grid1 = (unsigned)(ceil_div(n > 0 ? n : 1, rpc) + B);  // >= the non-empty (scene, chunk) pairs, see resolve_chunk
  const int V = x_dtype == MK_F32 ? 4 : 8;
  if (C % V == 0 && C / V <= kNormThreads && ((uintptr_t)x & 15) == 0) {
    const int RL = kNormThreads / (C / V);
    const size_t smem = (size_t)RL * 2 * C * sizeof(float);
    if (x_dtype == MK_F32)
      k_in_stats<float><<<grid1, kNormThreads, smem, st>>>((const float*)x, C, scene_offsets, B, rpc, cps, eps, partial, counters, mean, rstd);
    else
      k_in_stats<__nv_bfloat16><<<grid1, kNormThreads, smem, st>>>((const __nv_bfloat16*)x, C, scene_offsets, B, rpc, cps, eps, partial,
                                                                   counters, mean, rstd);
    k_in_stats_final<<<dim3((unsigned)ceil_div(C, 8), (unsigned)B), 256, 0, st>>>(partial, C, cps, eps, mean, rstd, scene_offsets, rpc);
    return check_launch(__func__);
  }
  const int CV = C / 4, RL = kNormThreads / CV;
  size_t smem = (size_t)RL * 2 * C * sizeof(float);
  if (x_dtype == MK_F32)
    k_in_stats_partial<float><<<grid, kNormThreads, smem, st>>>((const float*)x, C, scene_offsets, rpc, cps, partial);
  else
    k_in_stats_partial<__nv_bfloat16><<<grid, kNormThreads, smem, st>>>((const __nv_bfloat16*)x, C, scene_offsets, rpc, cps, partial);
  k_in_stats_final<<<dim3((unsigned)ceil_div(C, 8), (unsigned)B), 256, 0, st>>>(partial, C, cps, eps, mean, rstd, scene_offsets, rpc);
  return check_launch(__func__);
}

int mk_norm_apply(const void* x, int x_dtype, int64_t n, int C, const int32_t* scene_offsets, int B, const float* mean,
                  const float* rstd, const float* gamma, const float* beta, const void* residual, int relu, void* out,
                  int out_dtype, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(C % 4 == 0, "C must be a multiple of 4");
  if (n == 0) return 0;
  int64_t total = n * (C / 4);
  int64_t blocks = ceil_div(total, 256);
  int64_t cap = 8 * (int64_t)num_sms();
  if (blocks > cap) blocks = cap;
  if (x_dtype == MK_BF16 && out_dtype == MK_BF16 && C % 8 == 0 && mean && gamma && n * (int64_t)(C / 8) < ((int64_t)1 << 31) &&
      (((uintptr_t)x | (uintptr_t)out | (uintptr_t)residual) & 15) == 0) {
    const int64_t tot8 = n * (C / 8);
    const int64_t b8 = row_invariant_grid(tot8, C / 8);
    k_norm_apply_bf16x8<<<(unsigned)b8, 256, 0, st>>>((const __nv_bfloat16*)x, (unsigned)n, (unsigned)(C / 8), scene_offsets, B, mean, rstd,
                                                      gamma, beta, (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)out);
    return check_launch(__func__);
  }
#define MK_APPLY(TI, TO)                                                                                              \
  k_norm_apply<TI, TO><<<(unsigned)blocks, 256, 0, st>>>((const TI*)x, n, C, scene_offsets, B, mean, rstd, gamma, beta, \
                                                         (const TO*)residual, relu, (TO*)out)
  if (x_dtype == MK_F32 && out_dtype == MK_F32) MK_APPLY(float, float);
  else if (x_dtype == MK_F32 && out_dtype == MK_BF16) MK_APPLY(float, __nv_bfloat16);
  else if (x_dtype == MK_BF16 && out_dtype == MK_F32) MK_APPLY(__nv_bfloat16, float);
  else MK_APPLY(__nv_bfloat16, __nv_bfloat16);
#undef MK_APPLY
  return check_launch(__func__);
}

int mk_instnorm_relu_bwd(const void* dy, const void* out, const void* x, int dtype, int64_t n, int C, const int32_t* scene_offsets,
                         int B, const float* mean, const float* rstd, const float* gamma, void* dx, int dx_dtype, void* dres,
                         float* dgamma, float* dbeta, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  const int V = dtype == MK_F32 ? 4 : 8;
  MK_REQUIRE(C % V == 0 && C / V <= kNormThreads, "C must be a multiple of 4 (fp32) / 8 (bf16)");
  MK_REQUIRE(dx_dtype == dtype || (dtype == MK_BF16 && dx_dtype == MK_F32), "dx must have the input dtype (or fp32 for bf16 inputs)");
  MK_REQUIRE(workspace_bytes >= mk_instnorm_workspace_bytes(n, C, B), "workspace too small");
  MK_REQUIRE((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)out | (uintptr_t)dx | (uintptr_t)dres) & 15) == 0, "16-byte alignment");
  int rpc, cps;
  norm_chunking(n, B, &rpc, &cps);
  float* partial = (float*)workspace;
  float* sum_a = partial + (int64_t)B * cps * 4 * C;
  float* sum_q = sum_a + (int64_t)B * C;
  const unsigned grid1 = (unsigned)(ceil_div(n > 0 ? n : 1, rpc) + B);
  const int RL = kNormThreads / (C / V);
  const size_t smem = (size_t)RL * 2 * C * sizeof(float);
  if (dtype == MK_F32)
    k_inr_bwd_partial<float><<<grid1, kNormThreads, smem, st>>>((const float*)dy, (const float*)out, (const float*)x, C, scene_offsets, B, rpc,
                                                               cps, mean, rstd, partial);
  else
    k_inr_bwd_partial<__nv_bfloat16><<<grid1, kNormThreads, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)out,
                                                                       (const __nv_bfloat16*)x, C, scene_offsets, B, rpc, cps, mean, rstd,
                                                                       partial);
  k_inr_bwd_final<<<(unsigned)ceil_div(C, 8), 256, 0, st>>>(partial, C, cps, B, sum_a, sum_q, dgamma, dbeta, scene_offsets, rpc);
  if (n > 0) {
    const int64_t blocks = row_invariant_grid(n * (C / V), C / V);
    if (dtype == MK_F32)
      k_inr_bwd_apply<float, float><<<(unsigned)blocks, 256, 0, st>>>((const float*)dy, (const float*)out, (const float*)x, n, C,
                                                                      scene_offsets, B, mean, rstd, gamma, sum_a, sum_q, (float*)dx,
                                                                      (float*)dres);
    else if (dx_dtype == MK_BF16)
      k_inr_bwd_apply<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)out, (const __nv_bfloat16*)x, n, C, scene_offsets, B, mean, rstd, gamma, sum_a,
          sum_q, (__nv_bfloat16*)dx, (__nv_bfloat16*)dres);
    else
      k_inr_bwd_apply<__nv_bfloat16, float><<<(unsigned)blocks, 256, 0, st>>>(
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)out, (const __nv_bfloat16*)x, n, C, scene_offsets, B, mean, rstd, gamma, sum_a,
          sum_q, (float*)dx, (__nv_bfloat16*)dres);
  }
  return check_launch(__func__);
}

int mk_instnorm_bwd(const float* x, const float* dy, int64_t n, int C, const int32_t* scene_offsets, int B,
                    const float* mean, const float* rstd, const float* gamma, float* dx, float* dgamma, float* dbeta,
                    void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(workspace_bytes >= mk_instnorm_workspace_bytes(n, C, B), "workspace too small");
  int rpc, cps;
  norm_chunking(n, B, &rpc, &cps);
  float* partial = (float*)workspace;
  float* sum_dy = partial + (int64_t)B * cps * 4 * C;
  float* sum_dyy = sum_dy + (int64_t)B * C;
  dim3 grid((unsigned)cps, (unsigned)B);
  k_in_bwd_partial<<<grid, kNormThreads, 0, st>>>(x, dy, C, scene_offsets, rpc, cps, mean, rstd, partial);
  k_in_bwd_final<<<1, 256, 0, st>>>(partial, C, cps, B, scene_offsets, sum_dy, sum_dyy, dgamma, dbeta);
  if (n > 0) {
    int64_t blocks = ceil_div(n * C, 256);
    int64_t cap = 8 * (int64_t)num_sms();
    if (blocks > cap) blocks = cap;
    k_in_bwd_dx<<<(unsigned)blocks, 256, 0, st>>>(x, dy, n, C, scene_offsets, B, mean, rstd, gamma, sum_dy, sum_dyy, dx);
  }
  return check_launch(__func__);
}

}  // extern "C"
