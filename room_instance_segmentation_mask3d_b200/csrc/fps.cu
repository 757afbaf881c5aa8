// Furthest point sampling, all scenes in one launch sequence, tie-breaking identical to the reference kernel.
//
// Replaces pointnet2._ext.furthest_point_sampling (reference: mask3d/third_party/pointnet2/_ext_src/src/
// sampling_gpu.cu:72-176 — ONE 512-thread block per call, called once per scene from mask3d/models/mask3d.py:309-317,
// i.e. a single SM walking all n points m-1 times).  The reference's result is a function of its block size
// Bv = min(2^floor(log2 n), 512) (cuda_utils.h:17-21): a per-thread strict '>' scan (lines 111-112) followed by a
// left-biased shared-memory tree (lines 62-68: at stride s position p keeps its own value on ties with p+s,
// s = Bv/2 ... 1) selects
//     argmax over candidates of (d2 descending, bitreverse(k mod Bv) ascending, k ascending)
// — the last tree level prefers even threads, the one before threads = 0 mod 4, and so on.  That total order is packed
// into one 64-bit key  (float bits of d2) << 32 | ~rank,  rank = bitreverse(k mod Bv) * ceil(n / Bv) + k / Bv,  so a round
// is a grid-wide atomicMax: many CTAs per scene, one launch per round, HBM/L2-bound (20 B per point per round).
// Points with |p|^2 <= 1e-3 are skipped exactly as line 104 does; temp starts at 1e10 (sampling.cpp:76-78).
#include "common.cuh"

namespace mk {

constexpr int kFpsThreads = 256;

__device__ __forceinline__ int fps_virtual_block(int n, int* log2_out) {
  int pow2 = 0;
  while ((2 << pow2) <= n) ++pow2;  // floor(log2 n)
  if (pow2 > 9) pow2 = 9;
  *log2_out = pow2;
  return 1 << pow2;
}

// grid (blocks_per_scene, B); rounds 1 .. m-1.  Round r reads the winner of round r-1 (sample r-1), updates temp and
// atomicMax-es the key of sample r into best[b, r].
__global__ void __launch_bounds__(kFpsThreads) k_fps_round(const float* __restrict__ xyz, const int32_t* __restrict__ scene_offsets, int m,
                                                           int round, float* __restrict__ temp, unsigned long long* __restrict__ best,
                                                           int32_t* __restrict__ idx) {
  const int b = blockIdx.y;
  const int64_t s = scene_offsets[b];
  const int n = scene_offsets[b + 1] - scene_offsets[b];
  if (n <= 0 || round >= m) return;
  const float* pts = xyz + s * 3;
  float* tmp = temp + s;
  int lb;
  const int Bv = fps_virtual_block(n, &lb);
  const int R = (n + Bv - 1) / Bv;
  // winner of the previous round
  int old = 0;
  if (round > 1) {
    const unsigned long long key = best[(int64_t)b * m + round - 1];
    if (key != 0ull) {
      const unsigned rank = ~(unsigned)(key & 0xffffffffull);
      const unsigned brevk = rank / (unsigned)R, kdiv = rank % (unsigned)R;
      const unsigned kmod = lb ? (__brev(brevk) >> (32 - lb)) : 0u;
      old = (int)(kdiv * (unsigned)Bv + kmod);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) idx[(int64_t)b * m + round - 1] = old;  // sample round-1 (0 for round 1)
  const float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
  unsigned long long kbest = 0ull;
  for (int k = blockIdx.x * kFpsThreads + threadIdx.x; k < n; k += gridDim.x * kFpsThreads) {
    const float x2 = pts[k * 3 + 0], y2 = pts[k * 3 + 1], z2 = pts[k * 3 + 2];
    const float mag = (x2 * x2) + (y2 * y2) + (z2 * z2);
    if (mag <= 1e-3f) continue;
    const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
    const float t = round == 1 ? 1e10f : tmp[k];
    const float d2 = fminf(d, t);
    tmp[k] = d2;
    const unsigned kmod = (unsigned)(k % Bv);
    const unsigned brevk = lb ? (__brev(kmod) >> (32 - lb)) : 0u;
    const unsigned rank = brevk * (unsigned)R + (unsigned)(k / Bv);
    const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(~rank);
    kbest = key > kbest ? key : kbest;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_down_sync(0xffffffffu, kbest, d);
    kbest = o > kbest ? o : kbest;
  }
  __shared__ unsigned long long sw[kFpsThreads / 32];
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = kbest;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long v = sw[0];
    for (int i = 1; i < kFpsThreads / 32; ++i) v = sw[i] > v ? sw[i] : v;
    if (v != 0ull) atomicMax(&best[(int64_t)b * m + round], v);
  }
}

// decode the last round's winner
__global__ void k_fps_finish(const int32_t* __restrict__ scene_offsets, int B, int m, const unsigned long long* __restrict__ best,
                             int32_t* __restrict__ idx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || m < 2) return;
  const int n = scene_offsets[b + 1] - scene_offsets[b];
  if (n <= 0) return;
  int lb;
  const int Bv = fps_virtual_block(n, &lb);
  const int R = (n + Bv - 1) / Bv;
  const unsigned long long key = best[(int64_t)b * m + m - 1];
  int old = 0;
  if (key != 0ull) {
    const unsigned rank = ~(unsigned)(key & 0xffffffffull);
    const unsigned brevk = rank / (unsigned)R, kdiv = rank % (unsigned)R;
    const unsigned kmod = lb ? (__brev(brevk) >> (32 - lb)) : 0u;
    old = (int)(kdiv * (unsigned)Bv + kmod);
  }
  idx[(int64_t)b * m + m - 1] = old;
}

}  // namespace mk

using namespace mk;

extern "C" int64_t mk_fps_workspace_bytes(int B, int m) { return (int64_t)B * m * (int64_t)sizeof(unsigned long long); }

// round r (1..m-1) computes the winner stored in best[b, r]; the kernel of round r+1 (or k_fps_finish) decodes it to idx[b, r].
extern "C" int mk_fps(const float* xyz, const int32_t* scene_offsets, int B, int m, int64_t n_total, float* temp, int32_t* idx,
                      void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (B <= 0 || m <= 0) return 0;
  MK_REQUIRE(workspace_bytes >= mk_fps_workspace_bytes(B, m), "workspace too small");
  unsigned long long* best = (unsigned long long*)workspace;
  cudaMemsetAsync(best, 0, (size_t)B * m * sizeof(unsigned long long), st);
  cudaMemsetAsync(idx, 0, (size_t)B * m * sizeof(int32_t), st);
  int64_t bps = ceil_div(n_total > 0 ? n_total : 1, (int64_t)kFpsThreads * 8);  // ~8 points per thread
  int64_t cap = (4 * (int64_t)num_sms() + B - 1) / B;
  if (bps > cap) bps = cap;
  if (bps < 1) bps = 1;
  dim3 grid((unsigned)bps, (unsigned)B);
  for (int r = 1; r < m; ++r) k_fps_round<<<grid, kFpsThreads, 0, st>>>(xyz, scene_offsets, m, r, temp, best, idx);
  k_fps_finish<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(scene_offsets, B, m, best, idx);
  return check_launch(__func__);
}
