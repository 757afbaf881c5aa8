// Mask3D decoder kernels: mask-logit product, fourier positional encoding, masked cross-attention core.
//
// Replaces, for one decoder layer (reference: mask3d/models/mask3d.py:397-478, 520-561):
//   * the per-scene Python loop `decomposed_features[i] @ mask_embed[i].T` (:534-535)      -> mk_mask_logits
//   * get_pos_encs' per-scene fourier embedding (:261-278, position_embedding.py:127-156)  -> mk_scene_minmax + mk_fourier_pe
//   * pad-to-max-K + dense [B*8,Q,K] bool mask + stock MHA with materialised weights (:400-478)
//       -> mk_masked_xattn_fwd: varlen over scenes (no pads exist), split over keys (flash-decoding style),
//          online softmax per (query, head) in registers, "query blocks every key => unblock all" rule (:456)
//          resolved by a per-(scene, query) any-open flag computed in the same launch sequence.
// All three are HBM/L2-bound (Q = 25 queries => <= 25 FLOP per K/V byte).
#include <cfloat>

#include "common.cuh"

namespace mk {

// ---------------------------------------------------------------------------------------------------
// mask logits: logits[n, q] = feat[n, :] . embed[b, q, :]
// ---------------------------------------------------------------------------------------------------
constexpr int kMlRows = 64;
constexpr int kMlThreads = 256;
constexpr int kMlMaxQPerThread = 32;  // Q <= 128

template <typename T>
__global__ void __launch_bounds__(kMlThreads) k_mask_logits(const T* __restrict__ feat, int C, const int32_t* __restrict__ scene_offsets,
                                                           const float* __restrict__ embed, int Q, float* __restrict__ logits) {
  extern __shared__ float sm[];
  float* sE = sm;                 // [Q][C]
  float* sF = sm + (size_t)Q * C; // [kMlRows][C + 1]
  const int b = blockIdx.y;
  const int64_t start = (int64_t)scene_offsets[b] + (int64_t)blockIdx.x * kMlRows;
  const int64_t end = min((int64_t)scene_offsets[b + 1], start + kMlRows);
  if (start >= end) return;
  const int nrows = (int)(end - start);
  for (int i = threadIdx.x; i < Q * C; i += kMlThreads) sE[i] = embed[(int64_t)b * Q * C + i];
  const int CV = C >> 2;
  for (int i = threadIdx.x; i < nrows * CV; i += kMlThreads) {
    const int r = i / CV, c = (i - r * CV) * 4;
    float4 v = load4(feat + (start + r) * C + c);
    float* d = sF + r * (C + 1) + c;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int r = threadIdx.x % kMlRows, qg = threadIdx.x / kMlRows;  // 4 query groups
  if (r >= nrows) return;
  float acc[kMlMaxQPerThread];
#pragma unroll
  for (int j = 0; j < kMlMaxQPerThread; ++j) acc[j] = 0.f;
  const float* f = sF + r * (C + 1);
  const int nq = (Q - qg + 3) / 4;  // queries qg, qg+4, ...
  for (int c = 0; c < C; ++c) {
    const float x = f[c];
#pragma unroll
    for (int j = 0; j < kMlMaxQPerThread; ++j)
      if (j < nq) acc[j] = fmaf(x, sE[(qg + 4 * j) * C + c], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < kMlMaxQPerThread; ++j)
    if (j < nq) logits[(start + r) * Q + qg + 4 * j] = acc[j];
}

// ---------------------------------------------------------------------------------------------------
// per-scene min / max of fp32 [n,3] coordinates
// ---------------------------------------------------------------------------------------------------
// order-preserving float <-> uint map so that min / max can use integer atomics
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__global__ void k_minmax_init(unsigned* __restrict__ omin, unsigned* __restrict__ omax, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    omin[i] = 0xffffffffu;
    omax[i] = 0u;
  }
}

// grid (chunks, B)
__global__ void __launch_bounds__(256) k_scene_minmax(const float* __restrict__ xyz, const int32_t* __restrict__ scene_offsets,
                                                      int rows_per_chunk, unsigned* __restrict__ omin, unsigned* __restrict__ omax) {
  const int b = blockIdx.y;
  const int64_t s = (int64_t)scene_offsets[b] + (int64_t)blockIdx.x * rows_per_chunk;
  const int64_t e = min((int64_t)scene_offsets[b + 1], s + rows_per_chunk);
  if (s >= e) return;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t r = s + threadIdx.x; r < e; r += blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = xyz[r * 3 + a];
      mn[a] = fminf(mn[a], v);
      mx[a] = fmaxf(mx[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn[a] = fminf(mn[a], __shfl_down_sync(0xffffffffu, mn[a], d));
      mx[a] = fmaxf(mx[a], __shfl_down_sync(0xffffffffu, mx[a], d));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(&omin[b * 3 + a], f2ord(mn[a]));
      atomicMax(&omax[b * 3 + a], f2ord(mx[a]));
    }
  }
}

__global__ void k_minmax_finish(unsigned* __restrict__ omin, unsigned* __restrict__ omax, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    reinterpret_cast<float*>(omin)[i] = ord2f(omin[i]);
    reinterpret_cast<float*>(omax)[i] = ord2f(omax[i]);
  }
}

// ---------------------------------------------------------------------------------------------------
// fourier PE: out[n, j] = sin(p_j), out[n, D/2 + j] = cos(p_j), p = ((xyz - min) / (max - min) * 2pi) @ gauss_B
// optionally out = base + pe (the attention key input `src + pos`)
// ---------------------------------------------------------------------------------------------------
template <typename TB, typename TO>
__global__ void __launch_bounds__(256) k_fourier_pe(const float* __restrict__ xyz, int64_t n, const int32_t* __restrict__ scene_offsets,
                                                    int B, const float* __restrict__ cmin, const float* __restrict__ cmax,
                                                    const float* __restrict__ gauss_B, int D, const TB* __restrict__ base,
                                                    TO* __restrict__ out) {
  const int half = D >> 1;
  const int64_t total = n * half;
  // half is a power of two in practice (64): shift / mask instead of a 64-bit division per element
  const int sh = (half & (half - 1)) == 0 ? 31 - __clz(half) : -1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = sh >= 0 ? (i >> sh) : i / half;
    const int j = (int)(i - r * half);
    const int b = scene_of_row(scene_offsets, B, r);
    float p = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float lo = cmin[b * 3 + a], hi = cmax[b * 3 + a];
      // same association as position_embedding.py:33-43,146: ((x - lo) * 1) / (hi - lo) + 0, then * 2pi
      float x = ((xyz[r * 3 + a] - lo) * 1.0f) / (hi - lo) + 0.0f;
      x *= 6.283185307179586f;
      p = fmaf(x, gauss_B[a * half + j], p);
    }
    // Cody-Waite reduction to [-pi, pi] with a two-term 2*pi, then the SFU sine / cosine (absolute error ~5e-7 there): the
    // accurate sincosf is ~40 instructions with a slow path, and this kernel evaluates 57 M of them per 0.9 M-voxel scene
    float s, c;
    {
      const float kq = rintf(p * 0.15915494309189535f);
      float r = fmaf(kq, -6.2831854820251465f, p);      // 2*pi rounded to fp32
      r = fmaf(kq, 1.7484555314695172e-07f, r);          // 2*pi - fp32(2*pi)
      s = __sinf(r);
      c = __cosf(r);
    }
    if (base) {
      s += to_f32<TB>(base[r * D + j]);
      c += to_f32<TB>(base[r * D + half + j]);
    }
    out[r * D + j] = from_f32<TO>(s);
    out[r * D + half + j] = from_f32<TO>(c);
  }
}

// Eight frequencies per thread: the three normalised coordinates (three fp32 divisions, the scene lookup, the coordinate loads) are
// computed once per 8 outputs instead of once per output, and sine / cosine halves go out as 16-byte (bf16) or 32-byte (fp32)
// vectors instead of 2- / 4-byte scalars.  Same expressions in the same order per element => bit-identical to k_fourier_pe.
template <typename T>
__device__ __forceinline__ void pe_load8(const T* p, float (&v)[8]) {
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
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
template <typename T>
__device__ __forceinline__ void pe_store8(T* p, const float (&v)[8]) {
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
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <typename TB, typename TO>
__global__ void __launch_bounds__(256) k_fourier_pe8(const float* __restrict__ xyz, int64_t n, const int32_t* __restrict__ scene_offsets,
                                                     int B, const float* __restrict__ cmin, const float* __restrict__ cmax,
                                                     const float* __restrict__ gauss_B, int D, const TB* __restrict__ base,
                                                     TO* __restrict__ out) {
  const int half = D >> 1, groups = half >> 3;  // groups of 8 frequencies per row
  const int64_t total = n * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / groups;
    const int j0 = (int)(i - r * groups) * 8;
    const int b = scene_of_row(scene_offsets, B, r);
    float x[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float lo = cmin[b * 3 + a], hi = cmax[b * 3 + a];
      x[a] = ((xyz[r * 3 + a] - lo) * 1.0f) / (hi - lo) + 0.0f;
      x[a] *= 6.283185307179586f;
    }
    float s[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) p = fmaf(x[a], gauss_B[a * half + j0 + j], p);
      const float kq = rintf(p * 0.15915494309189535f);
      float rr = fmaf(kq, -6.2831854820251465f, p);
      rr = fmaf(kq, 1.7484555314695172e-07f, rr);
      s[j] = __sinf(rr);
      c[j] = __cosf(rr);
    }
    if (base) {
      float bs[8], bc[8];
      pe_load8<TB>(base + r * D + j0, bs);
      pe_load8<TB>(base + r * D + half + j0, bc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += bs[j];
        c[j] += bc[j];
      }
    }
    pe_store8<TO>(out + r * D + j0, s);
    pe_store8<TO>(out + r * D + half + j0, c);
  }
}

// ---------------------------------------------------------------------------------------------------
// masked cross-attention core
// ---------------------------------------------------------------------------------------------------
// any_open[b, q] = 1 iff some key of scene b is not blocked for query q
__global__ void __launch_bounds__(256) k_any_open(const uint8_t* __restrict__ blocked, int64_t n, int Q,
                                                  const int32_t* __restrict__ scene_offsets, int B, int32_t* __restrict__ any_open) {
  // 16 mask bytes per thread: most are "blocked" (non-zero), so the common case is one 16-byte load and four compares; only open
  // entries pay the division / scene lookup, and only while their flag is still clear
  const int64_t total = n * Q;
  const int64_t nvec = (((uintptr_t)blocked & 15) == 0) ? total >> 4 : 0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    const uint4 w4 = reinterpret_cast<const uint4*>(blocked)[v];
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      // a zero byte somewhere in the word?  (classic has-zero-byte test)
      if (((w[u] - 0x01010101u) & ~w[u] & 0x80808080u) == 0u) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (((w[u] >> (8 * j)) & 0xffu) == 0u) {
          const int64_t i = (v << 4) + 4 * u + j;
          const int64_t r = i / Q;
          const int q = (int)(i - r * Q);
          int32_t* f = any_open + scene_of_row(scene_offsets, B, r) * Q + q;
          if (*f == 0) *f = 1;  // benign race: every writer stores 1
        }
      }
    }
  }
  for (int64_t i = (nvec << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (!blocked[i]) {
      const int64_t r = i / Q;
      const int q = (int)(i - r * Q);
      any_open[scene_of_row(scene_offsets, B, r) * Q + q] = 1;
    }
  }
}

constexpr int kXaKeysPerTile = 32;
constexpr int kXaMaxHd = 32;  // head dim <= 32 (partial record stride)

// grid (splits, B); block = Q*H threads rounded up to a warp multiple; thread = (query, head).
// partial[(b*splits + split)][q*H + h] = {m, l, acc[hd]}
template <typename T, int HD>
__global__ void __launch_bounds__(1024) k_xattn_partial(const float* __restrict__ qp, const T* __restrict__ kp, const T* __restrict__ vp,
                                                        const uint8_t* __restrict__ blocked, const int32_t* __restrict__ any_open,
                                                        const int32_t* __restrict__ scene_offsets, int Q, int D, int H, float scale,
                                                        int keys_per_split, float* __restrict__ partial) {
  extern __shared__ __align__(16) float sm[];
  const int hd = D / H;
  // [tile][H][hd + 4]: the 4-float pad per head spreads the 8 heads a quarter-warp reads (16 bytes each, stride hd floats = 64 B
  // => heads 0,2,4,6 on the same banks: 4-way conflicts) over distinct bank groups (stride 80 B)
  const int HP = hd + 4, RS = H * HP;
  float* sK = sm;
  float* sV = sm + (size_t)kXaKeysPerTile * RS;
  uint8_t* sM = (uint8_t*)(sV + (size_t)kXaKeysPerTile * RS);  // [tile][Q]
  const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
  const int64_t s0 = (int64_t)scene_offsets[b] + (int64_t)split * keys_per_split;
  const int64_t s1 = min((int64_t)scene_offsets[b + 1], s0 + keys_per_split);
  const int t = threadIdx.x;
  const bool active = t < Q * H;
  const int q = active ? t / H : 0, h = active ? t % H : 0;
  float qv[HD], acc[HD];
#pragma unroll
  for (int j = 0; j < HD; ++j) {
    // scores are kept in log2 units (scale * log2(e) folded into q): exp2f is one MUFU.EX2, expf a ~10-instruction sequence
    qv[j] = (active && j < hd) ? qp[((int64_t)b * Q + q) * D + h * hd + j] * (scale * 1.4426950408889634f) : 0.f;
    acc[j] = 0.f;
  }
  const bool use_mask = active && any_open[b * Q + q] != 0;
  float m = -INFINITY, l = 0.f;
  for (int64_t k0 = s0; k0 < s1; k0 += kXaKeysPerTile) {
    const int nk = (int)min((int64_t)kXaKeysPerTile, s1 - k0);
    __syncthreads();
    for (int i = t; i < nk * (D >> 2); i += blockDim.x) {
      const int r = i / (D >> 2), c = (i - r * (D >> 2)) * 4;
      const int hh = c / hd, jj = c - hh * hd;
      store4(sK + r * RS + hh * HP + jj, load4(kp + (k0 + r) * D + c));
      store4(sV + r * RS + hh * HP + jj, load4(vp + (k0 + r) * D + c));
    }
    for (int i = t; i < nk * Q; i += blockDim.x) sM[i] = blocked[k0 * Q + i];
    __syncthreads();
    if (active) {
      for (int r = 0; r < nk; ++r) {
        if (use_mask && sM[r * Q + q]) continue;
        const float4* kr = reinterpret_cast<const float4*>(sK + r * RS + h * HP);  // hd % 4 == 0 (checked by the entry point)
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < HD; j += 4) {
          if (j < hd) {
            const float4 kk = kr[j >> 2];
            s = fmaf(qv[j], kk.x, s);
            s = fmaf(qv[j + 1], kk.y, s);
            s = fmaf(qv[j + 2], kk.z, s);
            s = fmaf(qv[j + 3], kk.w, s);
          }
        }
        const float mn = fmaxf(m, s);
        const float corr = exp2f(m - mn);  // m = -inf on the first key: exp2(-inf) = 0
        const float pw = exp2f(s - mn);
        l = l * corr + pw;
        const float4* vr = reinterpret_cast<const float4*>(sV + r * RS + h * HP);
#pragma unroll
        for (int j = 0; j < HD; j += 4) {
          if (j < hd) {
            const float4 vv = vr[j >> 2];
            acc[j] = fmaf(acc[j], corr, pw * vv.x);
            acc[j + 1] = fmaf(acc[j + 1], corr, pw * vv.y);
            acc[j + 2] = fmaf(acc[j + 2], corr, pw * vv.z);
            acc[j + 3] = fmaf(acc[j + 3], corr, pw * vv.w);
          }
        }
        m = mn;
      }
    }
  }
  if (active) {
    float* o = partial + (((int64_t)b * nsplit + split) * Q * H + t) * (2 + kXaMaxHd);
    o[0] = m;
    o[1] = l;
#pragma unroll
    for (int j = 0; j < HD; ++j)
      if (j < hd) o[2 + j] = acc[j];
  }
}

// ---------------------------------------------------------------------------------------------------
// Tensor-core form of k_xattn_partial for the shapes the model runs (bf16 keys / values, D = 128, H = 8 heads of 16 channels,
// Q <= 32 queries): a flash-attention style kernel with warp-level softmax.  The SIMT kernel above spends ~60 instructions per
// (key, query, head) triple on CUDA cores (1.0-1.8 ms per launch on the 865 k keys of level 1 of a 4-scene batch, 15-25x above
// the time HBM needs for the 443 MB of keys and values); here a warp owns one head for all (padded to 32) queries:
//   S[32 q x 64 keys] = Q_h K_h^T   16 x mma.sync.m16n8k16 (bf16, fp32 accumulate), scores in log2 units (scale * log2e folded into Q)
//   mask (blocked bits, "all blocked => unblock" rule through any_open), online softmax in the accumulator fragments (row
//   max / sum by quad shuffles), P re-used in registers as the A operand of
//   O[32 q x 16] += P V_h       16 x mma.sync, V fragments by ldmatrix.trans
// K / V tiles (64 keys x 128 channels, rows of a scene are contiguous) arrive by 16-byte cp.async, double buffered.  The op is
// HBM-bound (Q = 25: <= 25 FLOP per key / value byte), so mma.sync is enough to get the arithmetic out
// of the way of the memory system; partial records and k_xattn_combine are those of the SIMT kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int kXmKeys = 64;            // keys per tile
constexpr int kXmPitch = 136;          // bf16 elements per staged row (128 + 8: consecutive rows shift by 4 banks -> conflict-free fragments)
constexpr int kXmThreads = 256;        // 8 warps = 8 heads

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(kXmThreads) k_xattn_partial_mma(const float* __restrict__ qp, const __nv_bfloat16* __restrict__ kp,
                                                                  const __nv_bfloat16* __restrict__ vp, const uint8_t* __restrict__ blocked,
                                                                  const int32_t* __restrict__ any_open,
                                                                  const int32_t* __restrict__ scene_offsets, int Q, float scale,
                                                                  int keys_per_split, float* __restrict__ partial) {
  constexpr int D = 128, H = 8, HD = 16;
  extern __shared__ __align__(16) uint8_t xm_smem[];
  __nv_bfloat16* sK = (__nv_bfloat16*)xm_smem;                          // [2][64][136]
  __nv_bfloat16* sV = sK + 2 * kXmKeys * kXmPitch;                      // [2][64][136]
  uint32_t* sM = (uint32_t*)(sV + 2 * kXmKeys * kXmPitch);              // [2][64] blocked bits per key (bit q)
  const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
  const int64_t s0 = (int64_t)scene_offsets[b] + (int64_t)split * keys_per_split;
  const int64_t s1 = min((int64_t)scene_offsets[b + 1], s0 + keys_per_split);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int h = warp;                 // head of this warp
  const int g = lane >> 2, tq = lane & 3;

  // Q_h as A fragments (2 m-tiles of 16 queries), scaled into log2 units, bf16
  uint32_t qa[2][4];
  const float qs = scale * 1.4426950408889634f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int q = mt * 16 + g + (r & 1) * 8;      // a0,a2: row g; a1,a3: row g + 8
      const int d = tq * 2 + (r >> 1) * 8;          // a0,a1: cols 2t..; a2,a3: cols 2t + 8..
      float x0 = 0.f, x1 = 0.f;
      if (q < Q) {
        const float* src = qp + ((int64_t)b * Q + q) * D + h * HD + d;
        x0 = src[0] * qs;
        x1 = src[1] * qs;
      }
      qa[mt][r] = pack_bf16x2(x0, x1);
    }
  }
  // per row (query) flags: rows g, g+8 of each m-tile
  uint32_t row_use_mask = 0;  // bit (mt*2 + hi): the query's mask is applied (some key of the scene is open for it)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int q = mt * 16 + g + hi * 8;
      if (q < Q && any_open[b * Q + q] != 0) row_use_mask |= 1u << (mt * 2 + hi);
    }

  float m_run[2][2], l_run[2][2], o[2][2][4];  // [m-tile][row half]; o[m-tile][d-block][4]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    l_run[mt][0] = l_run[mt][1] = 0.f;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[mt][nb][j] = 0.f;
  }

  const int64_t ntiles = s1 > s0 ? (s1 - s0 + kXmKeys - 1) / kXmKeys : 0;
  auto load_tile = [&](int64_t tile, int buf) {
    const int64_t k0 = s0 + tile * kXmKeys;
    // 64 rows x 16 pieces of 16 B per matrix = 1024 pieces: 4 per thread and matrix
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pidx = t + i * kXmThreads;
      const int r = pidx >> 4, c = (pidx & 15) * 8;
      const bool in = k0 + r < s1;
      const int64_t src = (in ? (k0 + r) : s0) * D + c;
      const uint32_t dk = (uint32_t)__cvta_generic_to_shared(sK + ((size_t)buf * kXmKeys + r) * kXmPitch + c);
      const uint32_t dv = (uint32_t)__cvta_generic_to_shared(sV + ((size_t)buf * kXmKeys + r) * kXmPitch + c);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dk), "l"(kp + src), "r"(in ? 16u : 0u) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dv), "l"(vp + src), "r"(in ? 16u : 0u) : "memory");
    }
    // blocked bits of the tile's keys: 4 threads per key, 8 queries each, OR-ed by quad shuffles
    {
      const int r = t >> 2, part = t & 3;
      uint32_t w = 0;
      if (k0 + r < s1) {
        const uint8_t* src = blocked + (k0 + r) * Q;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int q = part * 8 + j;
          if (q < Q && src[q]) w |= 1u << q;
        }
      }
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      if (part == 0) sM[buf * kXmKeys + r] = w;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (ntiles > 0) load_tile(0, 0);
  for (int64_t tile = 0; tile < ntiles; ++tile) {
    const int buf = (int)(tile & 1);
    if (tile + 1 < ntiles) {
      load_tile(tile + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __nv_bfloat16* tK = sK + (size_t)buf * kXmKeys * kXmPitch;
    const __nv_bfloat16* tV = sV + (size_t)buf * kXmKeys * kXmPitch;
    const uint32_t* tM = sM + buf * kXmKeys;
    const int nvalid = (int)min((int64_t)kXmKeys, s1 - (s0 + tile * kXmKeys));  // keys of this tile inside the split

    // ---- S = Q_h K_h^T for the 64 keys: 8 n-blocks of 8 keys ----
    float sc[2][8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      // B fragment: B[k = d][n = key]: b0 = K[key nb*8 + g][h*16 + 2t, +1], b1 = same row, d + 8
      const uint32_t* row = reinterpret_cast<const uint32_t*>(tK + (size_t)(nb * 8 + g) * kXmPitch + h * HD);
      const uint32_t b0 = row[tq], b1 = row[tq + 4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        sc[mt][nb][0] = sc[mt][nb][1] = sc[mt][nb][2] = sc[mt][nb][3] = 0.f;
        mma_bf16_16816(sc[mt][nb], qa[mt], b0, b1);
      }
    }
    // ---- mask + online softmax (rows g / g+8 of each m-tile; this thread's keys: nb*8 + 2t, +1) ----
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const uint32_t w0 = tM[nb * 8 + tq * 2], w1 = tM[nb * 8 + tq * 2 + 1];
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const int q = mt * 16 + g + hi * 8;
          const bool um = (row_use_mask >> (mt * 2 + hi)) & 1u;
          // the blocked bit applies only while the query's mask is in use; keys past the end of the split never take part
          const int key = nb * 8 + tq * 2;
          const bool x0 = key >= nvalid || (um && ((w0 >> q) & 1u));
          const bool x1 = key + 1 >= nvalid || (um && ((w1 >> q) & 1u));
          if (x0) sc[mt][nb][hi * 2] = -INFINITY;
          if (x1) sc[mt][nb][hi * 2 + 1] = -INFINITY;
          tmax[hi] = fmaxf(tmax[hi], fmaxf(sc[mt][nb][hi * 2], sc[mt][nb][hi * 2 + 1]));
        }
      }
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        tmax[hi] = fmaxf(tmax[hi], __shfl_xor_sync(0xffffffffu, tmax[hi], 1));
        tmax[hi] = fmaxf(tmax[hi], __shfl_xor_sync(0xffffffffu, tmax[hi], 2));
        const float mn = fmaxf(m_run[mt][hi], tmax[hi]);
        const float ms = mn == -INFINITY ? 0.f : mn;           // all keys so far excluded: keep everything at zero
        const float corr = exp2f(m_run[mt][hi] - ms);          // m_run = -inf -> 0
        float rs = 0.f;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const float p0 = exp2f(sc[mt][nb][hi * 2] - ms), p1 = exp2f(sc[mt][nb][hi * 2 + 1] - ms);
          sc[mt][nb][hi * 2] = p0;
          sc[mt][nb][hi * 2 + 1] = p1;
          rs += p0 + p1;
        }
        rs += __shfl_xor_sync(0xffffffffu, rs, 1);
        rs += __shfl_xor_sync(0xffffffffu, rs, 2);
        l_run[mt][hi] = l_run[mt][hi] * corr + rs;
        m_run[mt][hi] = mn;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          o[mt][nb][hi * 2] *= corr;
          o[mt][nb][hi * 2 + 1] *= corr;
        }
      }
    }
    // ---- O += P V_h: 4 k-steps of 16 keys, 2 d-blocks of 8 channels ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t pa[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        pa[mt][0] = pack_bf16x2(sc[mt][2 * ks][0], sc[mt][2 * ks][1]);
        pa[mt][1] = pack_bf16x2(sc[mt][2 * ks][2], sc[mt][2 * ks][3]);
        pa[mt][2] = pack_bf16x2(sc[mt][2 * ks + 1][0], sc[mt][2 * ks + 1][1]);
        pa[mt][3] = pack_bf16x2(sc[mt][2 * ks + 1][2], sc[mt][2 * ks + 1][3]);
      }
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        // B[k = key][n = d]: V rows (keys) ks*16 .. +15, channels h*16 + nb*8 .. +7, transposed on load
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(tV + (size_t)(ks * 16 + (lane & 15)) * kXmPitch + h * HD + nb * 8);
        uint32_t b0, b1;
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(o[mt][nb], pa[mt], b0, b1);
      }
    }
    __syncthreads();  // the buffer is refilled by the load issued in the next iteration
  }

  // partial record of (query, head): {m, l, acc[16]} at partial[((b * nsplit + split) * Q * H + q * H + h) * (2 + 32)]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int q = mt * 16 + g + hi * 8;
      if (q >= Q) continue;
      float* rec = partial + (((int64_t)b * nsplit + split) * Q * H + (int64_t)q * H + h) * (2 + kXaMaxHd);
      if (tq == 0) {
        rec[0] = m_run[mt][hi];
        rec[1] = l_run[mt][hi];
      }
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        rec[2 + nb * 8 + tq * 2] = o[mt][nb][hi * 2];
        rec[2 + nb * 8 + tq * 2 + 1] = o[mt][nb][hi * 2 + 1];
      }
    }
}

// grid (ceil(Q*H / 8), B), block 256: warp w handles (query, head) pair 8*blockIdx.x + w; its 32 lanes fold the splits
// lane, lane+32, ... in ascending order (online-softmax merge), then the lanes are merged by a fixed shuffle tree.
__global__ void __launch_bounds__(256) k_xattn_combine(const float* __restrict__ partial, int nsplit, int Q, int D, int H,
                                                       float* __restrict__ ctx, float* __restrict__ lse) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= Q * H) return;
  const int hd = D / H, q = t / H, h = t % H;
  float m = -INFINITY, l = 0.f, acc[kXaMaxHd];
#pragma unroll
  for (int j = 0; j < kXaMaxHd; ++j) acc[j] = 0.f;
  for (int s = lane; s < nsplit; s += 32) {
    const float* p = partial + (((int64_t)b * nsplit + s) * Q * H + t) * (2 + kXaMaxHd);
    const float pm = p[0], pl = p[1];
    if (pl == 0.f) continue;
    const float mn = fmaxf(m, pm);
    const float c0 = exp2f(m - mn), c1 = exp2f(pm - mn);
    l = l * c0 + pl * c1;
#pragma unroll
    for (int j = 0; j < kXaMaxHd; ++j)
      if (j < hd) acc[j] = acc[j] * c0 + p[2 + j] * c1;
    m = mn;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float om = __shfl_down_sync(0xffffffffu, m, d), ol = __shfl_down_sync(0xffffffffu, l, d);
    const float mn = fmaxf(m, om);
    const float c0 = (m == -INFINITY) ? 0.f : exp2f(m - mn), c1 = (om == -INFINITY) ? 0.f : exp2f(om - mn);
    l = l * c0 + ol * c1;
#pragma unroll
    for (int j = 0; j < kXaMaxHd; ++j) {
      const float oa = __shfl_down_sync(0xffffffffu, acc[j], d);
      if (j < hd) acc[j] = acc[j] * c0 + oa * c1;
    }
    m = mn;
  }
  if (lane == 0) {
    const float inv = 1.f / l;
#pragma unroll
    for (int j = 0; j < kXaMaxHd; ++j)
      if (j < hd) ctx[((int64_t)b * Q + q) * D + h * hd + j] = acc[j] * inv;
    if (lse != nullptr) lse[(int64_t)b * Q * H + t] = m + log2f(l);  // log-sum-exp of the scaled scores in log2 units
  }
}

// ---------------------------------------------------------------------------------------------------
// backward of the masked cross-attention core.  With P = softmax(masked scale * q k^T) per (scene, query, head):
//   dV = P^T dO,   dP = dO V^T,   dS = P * (dP - delta),  delta = rowsum(dO * O),   dQ = scale dS K,   dK = scale dS^T Q.
// P is recomputed from the saved log-sum-exp.  Two kernels, neither needs atomics: a KEY-stationary one (thread = (key, head),
// all Q queries of the scene in shared memory) for dK / dV, and a QUERY-stationary one shaped like the forward (thread =
// (query, head), K / V tiles in shared memory, split over keys) for dQ.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_xattn_delta(const float* __restrict__ ctx, const float* __restrict__ dctx, int total, int hd,
                                                     float* __restrict__ delta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (b, q, h) flattened; heads are contiguous hd-blocks of D
  if (i >= total) return;
  float a = 0.f;
  for (int j = 0; j < hd; ++j) a = fmaf(ctx[(int64_t)i * hd + j], dctx[(int64_t)i * hd + j], a);
  delta[i] = a;
}

constexpr int kXbKeysPerBlock = 256;

template <typename T, int HD>
__global__ void __launch_bounds__(256)
    k_xattn_bwd_kv(const float* __restrict__ qp, const T* __restrict__ kp, const T* __restrict__ vp, const uint8_t* __restrict__ blocked,
                   const int32_t* __restrict__ any_open, const int32_t* __restrict__ scene_offsets, const float* __restrict__ dctx,
                   const float* __restrict__ lse, const float* __restrict__ delta, int Q, int D, int H, float scale,
                   float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  const int hd = D / H, HP = hd + 4, RS = H * HP;
  float* sQ = sm;                          // [Q][RS]  scale * q
  float* sO = sQ + (size_t)Q * RS;         // [Q][RS]  dO
  float* sL = sO + (size_t)Q * RS;         // [Q*H]
  float* sDl = sL + Q * H;                 // [Q*H]
  int* sOpen = (int*)(sDl + Q * H);        // [Q]
  uint8_t* sM = (uint8_t*)(sOpen + Q);     // [32][Q]
  const int b = blockIdx.y;
  const int64_t s0 = (int64_t)scene_offsets[b] + (int64_t)blockIdx.x * kXbKeysPerBlock;
  const int64_t s1 = min((int64_t)scene_offsets[b + 1], s0 + kXbKeysPerBlock);
  if (s0 >= s1) return;
  const int t = threadIdx.x, r = t / H, h = t - r * H;
  for (int i = t; i < Q * (D >> 2); i += blockDim.x) {
    const int q = i / (D >> 2), c = (i - q * (D >> 2)) * 4;
    const int hh = c / hd, jj = c - hh * hd;
    float4 a = load4(qp + ((int64_t)b * Q + q) * D + c);
    a.x *= scale, a.y *= scale, a.z *= scale, a.w *= scale;
    store4(sQ + q * RS + hh * HP + jj, a);
    store4(sO + q * RS + hh * HP + jj, load4(dctx + ((int64_t)b * Q + q) * D + c));
  }
  for (int i = t; i < Q * H; i += blockDim.x) {
    sL[i] = lse[(int64_t)b * Q * H + i];
    sDl[i] = delta[(int64_t)b * Q * H + i];
  }
  for (int i = t; i < Q; i += blockDim.x) sOpen[i] = any_open[b * Q + i];
  for (int64_t k0 = s0; k0 < s1; k0 += kXaKeysPerTile) {
    const int nk = (int)min((int64_t)kXaKeysPerTile, s1 - k0);
    __syncthreads();
    for (int i = t; i < nk * Q; i += blockDim.x) sM[i] = blocked[k0 * Q + i];
    __syncthreads();
    if (r >= nk) continue;
    const int64_t row = k0 + r;
    float kk[HD], vv[HD], dkr[HD], dvr[HD];
#pragma unroll
    for (int j = 0; j < HD; j += 4) {
      if (j < hd) {
        const float4 a = load4(kp + row * D + h * hd + j), c = load4(vp + row * D + h * hd + j);
        kk[j] = a.x, kk[j + 1] = a.y, kk[j + 2] = a.z, kk[j + 3] = a.w;
        vv[j] = c.x, vv[j + 1] = c.y, vv[j + 2] = c.z, vv[j + 3] = c.w;
      } else {
        kk[j] = kk[j + 1] = kk[j + 2] = kk[j + 3] = 0.f;
        vv[j] = vv[j + 1] = vv[j + 2] = vv[j + 3] = 0.f;
      }
      dkr[j] = dkr[j + 1] = dkr[j + 2] = dkr[j + 3] = 0.f;
      dvr[j] = dvr[j + 1] = dvr[j + 2] = dvr[j + 3] = 0.f;
    }
    for (int q = 0; q < Q; ++q) {
      if (sOpen[q] && sM[r * Q + q]) continue;
      const float4* qr = reinterpret_cast<const float4*>(sQ + q * RS + h * HP);
      const float4* orow = reinterpret_cast<const float4*>(sO + q * RS + h * HP);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int j = 0; j < HD; j += 4) {
        if (j < hd) {
          const float4 a = qr[j >> 2], c = orow[j >> 2];
          s = fmaf(a.x, kk[j], s), s = fmaf(a.y, kk[j + 1], s), s = fmaf(a.z, kk[j + 2], s), s = fmaf(a.w, kk[j + 3], s);
          dp = fmaf(c.x, vv[j], dp), dp = fmaf(c.y, vv[j + 1], dp), dp = fmaf(c.z, vv[j + 2], dp), dp = fmaf(c.w, vv[j + 3], dp);
        }
      }
      const float p = exp2f(s * 1.4426950408889634f - sL[q * H + h]);
      const float ds = p * (dp - sDl[q * H + h]);
#pragma unroll
      for (int j = 0; j < HD; j += 4) {
        if (j < hd) {
          const float4 a = qr[j >> 2], c = orow[j >> 2];
          dvr[j] = fmaf(p, c.x, dvr[j]), dvr[j + 1] = fmaf(p, c.y, dvr[j + 1]);
          dvr[j + 2] = fmaf(p, c.z, dvr[j + 2]), dvr[j + 3] = fmaf(p, c.w, dvr[j + 3]);
          dkr[j] = fmaf(ds, a.x, dkr[j]), dkr[j + 1] = fmaf(ds, a.y, dkr[j + 1]);
          dkr[j + 2] = fmaf(ds, a.z, dkr[j + 2]), dkr[j + 3] = fmaf(ds, a.w, dkr[j + 3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < HD; j += 4) {
      if (j < hd) {
        store4(dk + row * D + h * hd + j, make_float4(dkr[j], dkr[j + 1], dkr[j + 2], dkr[j + 3]));
        store4(dv + row * D + h * hd + j, make_float4(dvr[j], dvr[j + 1], dvr[j + 2], dvr[j + 3]));
      }
    }
  }
}

// grid (splits, B), thread = (query, head): the forward's loop with dQ in place of the output accumulator.
// partial[(b*splits + split)][q*H + h][hd] = sum over the split's keys of dS * k
template <typename T, int HD>
__global__ void __launch_bounds__(512)
    k_xattn_bwd_q(const float* __restrict__ qp, const T* __restrict__ kp, const T* __restrict__ vp, const uint8_t* __restrict__ blocked,
                  const int32_t* __restrict__ any_open, const int32_t* __restrict__ scene_offsets, const float* __restrict__ dctx,
                  const float* __restrict__ lse, const float* __restrict__ delta, int Q, int D, int H, float scale, int keys_per_split,
                  float* __restrict__ partial) {
  extern __shared__ __align__(16) float sm[];
  const int hd = D / H;
  const int HP = hd + 4, RS = H * HP;
  float* sK = sm;
  float* sV = sm + (size_t)kXaKeysPerTile * RS;
  uint8_t* sM = (uint8_t*)(sV + (size_t)kXaKeysPerTile * RS);  // [tile][Q]
  const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
  const int64_t s0 = (int64_t)scene_offsets[b] + (int64_t)split * keys_per_split;
  const int64_t s1 = min((int64_t)scene_offsets[b + 1], s0 + keys_per_split);
  const int t = threadIdx.x;
  const int qb = blockDim.x / H;  // queries per block; blockIdx.z = query chunk
  const int q = blockIdx.z * qb + t / H, h = t % H;
  const bool active = t < qb * H && q < Q;
  float qv[HD], ov[HD], acc[HD];
#pragma unroll
  for (int j = 0; j < HD; ++j) {
    const bool in = active && j < hd;
    qv[j] = in ? qp[((int64_t)b * Q + q) * D + h * hd + j] * (scale * 1.4426950408889634f) : 0.f;
    ov[j] = in ? dctx[((int64_t)b * Q + q) * D + h * hd + j] : 0.f;
    acc[j] = 0.f;
  }
  const bool use_mask = active && any_open[b * Q + q] != 0;
  const float L = active ? lse[(int64_t)b * Q * H + q * H + h] : 0.f, dl = active ? delta[(int64_t)b * Q * H + q * H + h] : 0.f;
  for (int64_t k0 = s0; k0 < s1; k0 += kXaKeysPerTile) {
    const int nk = (int)min((int64_t)kXaKeysPerTile, s1 - k0);
    __syncthreads();
    for (int i = t; i < nk * (D >> 2); i += blockDim.x) {
      const int r = i / (D >> 2), c = (i - r * (D >> 2)) * 4;
      const int hh = c / hd, jj = c - hh * hd;
      store4(sK + r * RS + hh * HP + jj, load4(kp + (k0 + r) * D + c));
      store4(sV + r * RS + hh * HP + jj, load4(vp + (k0 + r) * D + c));
    }
    for (int i = t; i < nk * Q; i += blockDim.x) sM[i] = blocked[k0 * Q + i];
    __syncthreads();
    if (active) {
      for (int r = 0; r < nk; ++r) {
        if (use_mask && sM[r * Q + q]) continue;
        const float4* kr = reinterpret_cast<const float4*>(sK + r * RS + h * HP);
        const float4* vr = reinterpret_cast<const float4*>(sV + r * RS + h * HP);
        float s = 0.f, dp = 0.f;
#pragma unroll
        for (int j = 0; j < HD; j += 4) {
          if (j < hd) {
            const float4 a = kr[j >> 2], c = vr[j >> 2];
            s = fmaf(qv[j], a.x, s), s = fmaf(qv[j + 1], a.y, s), s = fmaf(qv[j + 2], a.z, s), s = fmaf(qv[j + 3], a.w, s);
            dp = fmaf(ov[j], c.x, dp), dp = fmaf(ov[j + 1], c.y, dp), dp = fmaf(ov[j + 2], c.z, dp), dp = fmaf(ov[j + 3], c.w, dp);
          }
        }
        const float ds = exp2f(s - L) * (dp - dl);
#pragma unroll
        for (int j = 0; j < HD; j += 4) {
          if (j < hd) {
            const float4 a = kr[j >> 2];
            acc[j] = fmaf(ds, a.x, acc[j]), acc[j + 1] = fmaf(ds, a.y, acc[j + 1]);
            acc[j + 2] = fmaf(ds, a.z, acc[j + 2]), acc[j + 3] = fmaf(ds, a.w, acc[j + 3]);
          }
        }
      }
    }
  }
  if (active) {
    float* o = partial + (((int64_t)b * nsplit + split) * Q * H + q * H + h) * kXaMaxHd;
#pragma unroll
    for (int j = 0; j < HD; ++j)
      if (j < hd) o[j] = acc[j];
  }
}

// dq[b, q, h*hd + j] = scale * sum over splits (ascending: deterministic)
__global__ void __launch_bounds__(256) k_xattn_bwd_q_combine(const float* __restrict__ partial, int nsplit, int B, int Q, int D, int H,
                                                             float scale, float* __restrict__ dq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Q * D) return;
  const int hd = D / H;
  const int b = i / (Q * D), rem = i - b * Q * D, qq = rem / D, c = rem - qq * D, h = c / hd, j = c - h * hd;
  float a = 0.f;
  for (int s = 0; s < nsplit; ++s) a += partial[(((int64_t)b * nsplit + s) * Q * H + qq * H + h) * kXaMaxHd + j];
  dq[i] = a * scale;
}

int g_xattn_simt = 0;  // 1: force the CUDA-core forward kernel (A/B measurements, mk_set_xattn_simt)

inline int xattn_splits(int64_t n, int B) {
  // ~6 resident blocks (7 warps each, 34 KB of shared memory) per SM over the batch so that one block's tile load /
  // barrier hides behind the others' arithmetic; at least 128 keys per split
  int64_t s = ceil_div(6 * (int64_t)num_sms(), B > 0 ? B : 1);
  int64_t by_keys = ceil_div(n > 0 ? n : 1, 128);
  if (s > by_keys) s = by_keys;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace mk

using namespace mk;

extern "C" {

int mk_mask_logits(const void* feat, int x_dtype, int64_t n, int C, const int32_t* scene_offsets, int B, const float* embed,
                   int Q, float* logits, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(C % 4 == 0, "C must be a multiple of 4");
  MK_REQUIRE(Q >= 1 && Q <= 4 * kMlMaxQPerThread, "Q must be <= 128");
  if (n == 0) return 0;
  const size_t smem = ((size_t)Q * C + (size_t)kMlRows * (C + 1)) * sizeof(float);
  MK_REQUIRE(smem <= 200 * 1024, "Q*C too large for shared memory");
  dim3 grid((unsigned)ceil_div(n, kMlRows), (unsigned)B);
  if (x_dtype == MK_F32) {
    cudaFuncSetAttribute(k_mask_logits<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_mask_logits<float><<<grid, kMlThreads, smem, st>>>((const float*)feat, C, scene_offsets, embed, Q, logits);
  } else {
    cudaFuncSetAttribute(k_mask_logits<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_mask_logits<__nv_bfloat16><<<grid, kMlThreads, smem, st>>>((const __nv_bfloat16*)feat, C, scene_offsets, embed, Q, logits);
  }
  return check_launch(__func__);
}

int mk_scene_minmax(const float* xyz, int64_t n, const int32_t* scene_offsets, int B, float* cmin, float* cmax, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0) return 0;
  int64_t chunks = ceil_div(2 * (int64_t)num_sms(), B);
  int64_t rpc = ceil_div(n > 0 ? n : 1, chunks);
  if (rpc < 1024) rpc = 1024;
  chunks = ceil_div(n > 0 ? n : 1, rpc);
  k_minmax_init<<<(unsigned)ceil_div(3 * B, 128), 128, 0, st>>>((unsigned*)cmin, (unsigned*)cmax, 3 * B);
  k_scene_minmax<<<dim3((unsigned)chunks, (unsigned)B), 256, 0, st>>>(xyz, scene_offsets, (int)rpc, (unsigned*)cmin, (unsigned*)cmax);
  k_minmax_finish<<<(unsigned)ceil_div(3 * B, 128), 128, 0, st>>>((unsigned*)cmin, (unsigned*)cmax, 3 * B);
  return check_launch(__func__);
}

int mk_fourier_pe(const float* xyz, int64_t n, const int32_t* scene_offsets, int B, const float* cmin, const float* cmax,
                  const float* gauss_B, int D, const void* base, int base_dtype, void* out, int out_dtype, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(D % 2 == 0, "D must be even");
  if (n == 0) return 0;
  int64_t blocks = ceil_div(n * (D / 2), 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  const bool vec8 = (D % 16) == 0 && (((uintptr_t)out | (uintptr_t)base) & 15) == 0;
  int64_t blocks8 = ceil_div(n * (D / 16), 256);
  if (blocks8 > 16 * num_sms()) blocks8 = 16 * num_sms();
#define MK_PE(TB, TO)                                                                                                                    \
  do {                                                                                                                                  \
    if (vec8) k_fourier_pe8<TB, TO><<<(unsigned)blocks8, 256, 0, st>>>(xyz, n, scene_offsets, B, cmin, cmax, gauss_B, D, (const TB*)base, (TO*)out); \
    else k_fourier_pe<TB, TO><<<(unsigned)blocks, 256, 0, st>>>(xyz, n, scene_offsets, B, cmin, cmax, gauss_B, D, (const TB*)base, (TO*)out);        \
  } while (0)
  if (base_dtype == MK_F32 && out_dtype == MK_F32) MK_PE(float, float);
  else if (base_dtype == MK_F32 && out_dtype == MK_BF16) MK_PE(float, __nv_bfloat16);
  else if (base_dtype == MK_BF16 && out_dtype == MK_F32) MK_PE(__nv_bfloat16, float);
  else MK_PE(__nv_bfloat16, __nv_bfloat16);
#undef MK_PE
  return check_launch(__func__);
}

int mk_set_xattn_simt(int on) {
  g_xattn_simt = on ? 1 : 0;
  return 0;
}

int64_t mk_xattn_workspace_bytes(int B, int Q, int H, int64_t n_keys) {
  const int splits = xattn_splits(n_keys, B);
  return ((int64_t)B * splits * Q * H * (2 + kXaMaxHd)) * (int64_t)sizeof(float) + (int64_t)B * Q * sizeof(int32_t) + 256;
}

int mk_masked_xattn_fwd(const float* qp, const void* kp, const void* vp, int kv_dtype, int64_t n, const uint8_t* blocked,
                        const int32_t* scene_offsets, int B, int Q, int D, int H, float scale, float* ctx, void* workspace,
                        int64_t workspace_bytes, void* stream_) {
  return mk_masked_xattn_fwd_lse(qp, kp, vp, kv_dtype, n, blocked, scene_offsets, B, Q, D, H, scale, ctx, nullptr, workspace,
                                 workspace_bytes, stream_);
}

int mk_masked_xattn_fwd_lse(const float* qp, const void* kp, const void* vp, int kv_dtype, int64_t n, const uint8_t* blocked,
                            const int32_t* scene_offsets, int B, int Q, int D, int H, float scale, float* ctx, float* lse,
                            void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(D % H == 0 && D / H <= kXaMaxHd && (D / H) % 4 == 0, "head dim must be a multiple of 4, <= 32");
  MK_REQUIRE(Q * H <= 1024, "Q*H must be <= 1024");
  MK_REQUIRE(workspace_bytes >= mk_xattn_workspace_bytes(B, Q, H, n), "workspace too small");
  const int splits = xattn_splits(n, B);
  int32_t* any_open = (int32_t*)workspace;
  float* partial = (float*)((uint8_t*)workspace + (((int64_t)B * Q * sizeof(int32_t) + 255) / 256) * 256);
  cudaMemsetAsync(any_open, 0, (size_t)B * Q * sizeof(int32_t), st);
  if (n > 0) {
    int64_t blocks = ceil_div(ceil_div(n * Q, 16), 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    k_any_open<<<(unsigned)blocks, 256, 0, st>>>(blocked, n, Q, scene_offsets, B, any_open);
  }
  const int keys_per_split = (int)ceil_div(n > 0 ? n : 1, splits);  // upper bound on any scene's size / splits
  const int threads = ((Q * H + 31) / 32) * 32;
  const size_t smem = (size_t)2 * kXaKeysPerTile * (D + 4 * H) * sizeof(float) + (size_t)kXaKeysPerTile * Q;
  dim3 grid((unsigned)splits, (unsigned)B);
#define MK_XA(T, HD)                                                                                                        \
  k_xattn_partial<T, HD><<<grid, threads, smem, st>>>(qp, (const T*)kp, (const T*)vp, blocked, any_open, scene_offsets, Q, D, H, \
                                                      scale, keys_per_split, partial)
  const bool hd16 = D / H <= 16;
  if (kv_dtype == MK_BF16 && D == 128 && H == 8 && Q <= 32 && !g_xattn_simt) {
    // tensor-core path (flash-attention style, warp-level softmax); same partial records, same combine
    const size_t smem_mma = (size_t)4 * kXmKeys * kXmPitch * sizeof(__nv_bfloat16) + 2 * kXmKeys * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(k_xattn_partial_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
      attr_set = true;
    }
    k_xattn_partial_mma<<<grid, kXmThreads, smem_mma, st>>>(qp, (const __nv_bfloat16*)kp, (const __nv_bfloat16*)vp, blocked, any_open,
                                                            scene_offsets, Q, scale, keys_per_split, partial);
  } else if (kv_dtype == MK_F32) {
    if (hd16) MK_XA(float, 16); else MK_XA(float, 32);
  } else {
    if (hd16) MK_XA(__nv_bfloat16, 16); else MK_XA(__nv_bfloat16, 32);
  }
#undef MK_XA
  k_xattn_combine<<<dim3((unsigned)ceil_div(Q * H, 8), (unsigned)B), 256, 0, st>>>(partial, splits, Q, D, H, ctx, lse);
  return check_launch(__func__);
}

int64_t mk_xattn_bwd_workspace_bytes(int B, int Q, int H, int64_t n_keys) {
  const int splits = xattn_splits(n_keys, B);
  return ((int64_t)B * splits * Q * H * kXaMaxHd + (int64_t)B * Q * H) * (int64_t)sizeof(float) + (int64_t)B * Q * sizeof(int32_t) + 512;
}

int mk_masked_xattn_bwd(const float* qp, const void* kp, const void* vp, int kv_dtype, int64_t n, const uint8_t* blocked,
                        const int32_t* scene_offsets, int B, int Q, int D, int H, float scale, const float* ctx, const float* lse,
                        const float* dctx, int64_t max_scene_rows, float* dqp, float* dkp, float* dvp, void* workspace,
                        int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(D % H == 0 && D / H <= kXaMaxHd && (D / H) % 4 == 0, "head dim must be a multiple of 4, <= 32");
  MK_REQUIRE(H >= 1 && H <= 8, "backward: at most 8 heads");
  MK_REQUIRE(workspace_bytes >= mk_xattn_bwd_workspace_bytes(B, Q, H, n), "workspace too small");
  MK_REQUIRE(max_scene_rows >= 0 && max_scene_rows <= n, "max_scene_rows out of range");
  const int splits = xattn_splits(n, B);
  const int hd = D / H;
  uint8_t* w = (uint8_t*)workspace;
  int32_t* any_open = (int32_t*)w;
  w += (((int64_t)B * Q * sizeof(int32_t) + 255) / 256) * 256;
  float* delta = (float*)w;
  w += (((int64_t)B * Q * H * sizeof(float) + 255) / 256) * 256;
  float* partial = (float*)w;
  cudaMemsetAsync(any_open, 0, (size_t)B * Q * sizeof(int32_t), st);
  if (n > 0) {
    int64_t blocks = ceil_div(ceil_div(n * Q, 16), 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    k_any_open<<<(unsigned)blocks, 256, 0, st>>>(blocked, n, Q, scene_offsets, B, any_open);
  }
  k_xattn_delta<<<(unsigned)ceil_div((int64_t)B * Q * H, 256), 256, 0, st>>>(ctx, dctx, B * Q * H, hd, delta);
  const bool hd16 = hd <= 16;
  if (n > 0 && max_scene_rows > 0) {
    const size_t smem = (size_t)2 * Q * (D + 4 * H) * sizeof(float) + (size_t)2 * Q * H * sizeof(float) + (size_t)Q * sizeof(int) +
                        (size_t)kXaKeysPerTile * Q;
    dim3 grid((unsigned)ceil_div(max_scene_rows, kXbKeysPerBlock), (unsigned)B);
#define MK_XB(T, HD)                                                                                                             \
  do {                                                                                                                           \
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_xattn_bwd_kv<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    k_xattn_bwd_kv<T, HD><<<grid, 32 * H, smem, st>>>(qp, (const T*)kp, (const T*)vp, blocked, any_open, scene_offsets, dctx, lse, \
                                                     delta, Q, D, H, scale, dkp, dvp);                                          \
  } while (0)
    if (kv_dtype == MK_F32) {
      if (hd16) MK_XB(float, 16); else MK_XB(float, 32);
    } else {
      if (hd16) MK_XB(__nv_bfloat16, 16); else MK_XB(__nv_bfloat16, 32);
    }
#undef MK_XB
    if (check_launch(__func__)) return 1;
  }
  const int keys_per_split = (int)ceil_div(n > 0 ? n : 1, splits);
  const int qb = Q < 512 / H ? Q : 512 / H;  // queries per block (<= 512 threads)
  const int threads = ((qb * H + 31) / 32) * 32;
  const size_t smem_q = (size_t)2 * kXaKeysPerTile * (D + 4 * H) * sizeof(float) + (size_t)kXaKeysPerTile * Q;
  dim3 gq((unsigned)splits, (unsigned)B, (unsigned)ceil_div(Q, qb));
#define MK_XQ(T, HD)                                                                                                           \
  k_xattn_bwd_q<T, HD><<<gq, threads, smem_q, st>>>(qp, (const T*)kp, (const T*)vp, blocked, any_open, scene_offsets, dctx, lse, \
                                                    delta, Q, D, H, scale, keys_per_split, partial)
  if (kv_dtype == MK_F32) {
    if (hd16) MK_XQ(float, 16); else MK_XQ(float, 32);
  } else {
    if (hd16) MK_XQ(__nv_bfloat16, 16); else MK_XQ(__nv_bfloat16, 32);
  }
#undef MK_XQ
  k_xattn_bwd_q_combine<<<(unsigned)ceil_div((int64_t)B * Q * D, 256), 256, 0, st>>>(partial, splits, B, Q, D, H, scale, dqp);
  return check_launch(__func__);
}

}  // extern "C"
