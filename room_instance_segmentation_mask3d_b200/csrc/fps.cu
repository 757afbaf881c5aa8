// Furthest point sampling, all scenes in one launch, tie-breaking identical to the reference kernel.
//
// Replaces pointnet2._ext.furthest_point_sampling (reference: mask3d/third_party/pointnet2/_ext_src/src/
// sampling_gpu.cu:72-176, one 512-thread block per call, called once per scene from mask3d/models/mask3d.py:309-317).
// The reference's result is a function of its block size Bv = min(2^floor(log2 n), 512) (cuda_utils.h:17-21):
// per-thread strict '>' scan (lines 111-112) followed by a left-biased shared-memory tree (lines 62-68: at stride s
// position p keeps its own value on ties with p+s, s = Bv/2 ... 1) selects
//     argmax over candidates of (d2 descending, bitreverse(k mod Bv) ascending, k ascending)
// — the last tree level prefers even threads, the one before threads = 0 mod 4, and so on.
// This kernel evaluates that total order directly, so any thread count / scan order gives the same indices.
// Points with |p|^2 <= 1e-3 are skipped exactly as line 104 does; temp starts at 1e10 (sampling.cpp:76-78).
#include "common.cuh"

namespace mk {

struct Cand {
  float d;
  unsigned kmod;  // __brev(k mod Bv): only the relative order matters
  int k;
};

__device__ __forceinline__ bool better(const Cand& a, const Cand& b) {  // a strictly better than b
  if (a.d != b.d) return a.d > b.d;
  if (a.kmod != b.kmod) return a.kmod < b.kmod;
  return a.k < b.k;
}

__device__ __forceinline__ Cand shfl_down(const Cand& c, int d) {
  Cand r;
  r.d = __shfl_down_sync(0xffffffffu, c.d, d);
  r.kmod = __shfl_down_sync(0xffffffffu, c.kmod, d);
  r.k = __shfl_down_sync(0xffffffffu, c.k, d);
  return r;
}

constexpr int kFpsThreads = 1024;

__global__ void __launch_bounds__(kFpsThreads) k_fps(const float* __restrict__ xyz, const int32_t* __restrict__ scene_offsets, int m,
                                                     float* __restrict__ temp, int32_t* __restrict__ idx) {
  const int b = blockIdx.x;
  const int64_t s = scene_offsets[b];
  const int n = scene_offsets[b + 1] - scene_offsets[b];
  if (m <= 0 || n <= 0) return;
  const float* pts = xyz + s * 3;
  float* tmp = temp + s;
  int32_t* out = idx + (int64_t)b * m;
  // virtual block size of the reference launch
  int pow2 = 0;
  while ((2 << pow2) <= n) ++pow2;  // floor(log2 n)
  const int Bv = min(1 << pow2, 512);
  __shared__ Cand swarp[kFpsThreads / 32];
  __shared__ int s_old;
  for (int k = threadIdx.x; k < n; k += kFpsThreads) tmp[k] = 1e10f;
  int old = 0;
  if (threadIdx.x == 0) out[0] = 0;
  __syncthreads();
  for (int j = 1; j < m; ++j) {
    const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
    Cand best{-1.f, 0u, 0};  // a reference thread that saw no candidate contributes (best = -1, besti = 0)
    for (int k = threadIdx.x; k < n; k += kFpsThreads) {
      const float x2 = pts[k * 3 + 0], y2 = pts[k * 3 + 1], z2 = pts[k * 3 + 2];
      const float mag = (x2 * x2) + (y2 * y2) + (z2 * z2);
      if (mag <= 1e-3f) continue;
      const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
      const float d2 = fminf(d, tmp[k]);
      tmp[k] = d2;
      Cand c{d2, __brev((unsigned)(k % Bv)), k};
      if (better(c, best)) best = c;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      Cand o = shfl_down(best, d);
      if (better(o, best)) best = o;
    }
    if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
      Cand c = swarp[threadIdx.x];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        Cand o = shfl_down(c, d);
        if (better(o, c)) c = o;
      }
      if (threadIdx.x == 0) {
        s_old = c.k;
        out[j] = c.k;
      }
    }
    __syncthreads();
    old = s_old;
  }
}

}  // namespace mk

using namespace mk;

extern "C" int mk_fps(const float* xyz, const int32_t* scene_offsets, int B, int m, float* temp, int32_t* idx, void* stream_) {
  if (B <= 0) return 0;
  k_fps<<<B, kFpsThreads, 0, (cudaStream_t)stream_>>>(xyz, scene_offsets, m, temp, idx);
  return check_launch(__func__);
}
