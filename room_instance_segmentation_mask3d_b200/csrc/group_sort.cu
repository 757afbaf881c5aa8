// Stable grouping of rows by a 32-bit key with FEW distinct values: the tile-plan sort of the tensor-core convolution.
//
// The conv groups the output rows of a kernel map by their offset mask (the 27-bit set of present neighbours) so that a
// 128-row tile has one mask (engine.KernelMap.sorted_plan; subsystem (2) "kernel-map construction" of BASELINE.json's north
// star).  A scene has a few hundred distinct masks, so a general radix sort (4 passes over keys + payload; it was torch.sort =
// CUB, 7.5 % of the step's kernel time and linear in the batch size) does far more work than needed.  Here:
//
//   k_gs_insert    distinct keys -> open-addressing hash set (warp-aggregated: one atomicCAS per distinct key of a warp)
//   k_gs_rank      one block: compact the set, bitonic-sort the <= 1024 distinct keys, write each key's rank (its group id, in
//                  ascending key order = the order a stable sort by key produces) back into the table
//   k_gs_hist      one warp per 1024 consecutive rows: group id of every row (table lookup) + the block's histogram over groups
//   k_gs_scan      per group: exclusive scan of its counts over the blocks (one block per group);  k_gs_base: scan over groups
//   k_gs_scatter   one warp per 1024 rows, 32 rows at a time in row order: order[base[g] + offset[block][g] + rank among the
//                  earlier rows of the block with group g] = row   (match_any + popc: stable inside the step, counters carry on)
//
// Deterministic (no result depends on atomic ordering: the hash set's CONTENT is order independent, ranks come from sorted
// keys), stable, 3 reads of the 4-byte key per row and one 4-byte write.  More than 1024 distinct keys sets status bit 0
// (checked by the caller at its next host round trip).
#include "common.cuh"

namespace mk {

constexpr int kGsMaxGroups = 1024;
constexpr int kGsTableCap = 4096;          // hash set capacity (load factor <= 0.25)
constexpr uint32_t kGsEmpty = 0xFFFFFFFFu; // never a valid key: offset masks use at most 27 bits, small integer keys are >= 0
constexpr int kGsRowsPerWarp = 1024;

__device__ __forceinline__ uint32_t gs_hash(uint32_t k) {
  k ^= k >> 16;
  k *= 0x85ebca6bu;
  k ^= k >> 13;
  k *= 0xc2b2ae35u;
  k ^= k >> 16;
  return k;
}

__global__ void __launch_bounds__(256) k_gs_insert(const uint32_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ tkeys) {
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + (threadIdx.x & 31);
    const bool in = i < n;
    const uint32_t k = in ? keys[i] : kGsEmpty;
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    if (in && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) {  // one lane per distinct key of the warp
      uint32_t h = gs_hash(k) & (kGsTableCap - 1);
      for (int probe = 0; probe < kGsTableCap; ++probe) {
        const uint32_t cur = tkeys[h];
        if (cur == k) break;
        if (cur == kGsEmpty) {
          const uint32_t prev = atomicCAS(&tkeys[h], kGsEmpty, k);
          if (prev == kGsEmpty || prev == k) break;
        }
        h = (h + 1) & (kGsTableCap - 1);
      }
    }
  }
}

// one block of 1024 threads
__global__ void __launch_bounds__(1024) k_gs_rank(const uint32_t* __restrict__ tkeys, uint32_t* __restrict__ tids, int32_t* __restrict__ meta /* [0] = D, [1] = status */) {
  __shared__ uint32_t s_keys[kGsMaxGroups];
  __shared__ int s_count;
  const int t = threadIdx.x;
  if (t == 0) s_count = 0;
  s_keys[t] = kGsEmpty;
  __syncthreads();
  for (int s = t; s < kGsTableCap; s += 1024) {
    const uint32_t k = tkeys[s];
    if (k != kGsEmpty) {
      const int pos = atomicAdd(&s_count, 1);
      if (pos < kGsMaxGroups) s_keys[pos] = k;
    }
  }
  __syncthreads();
  const int D = s_count;
  if (t == 0) {
    meta[0] = D < kGsMaxGroups ? D : kGsMaxGroups;
    if (D > kGsMaxGroups) atomicOr(&meta[1], 1);
  }
  // bitonic sort of 1024 keys ascending (empty = 0xFFFFFFFF sorts last); the multiset is deterministic, so is the result
  for (int k = 2; k <= kGsMaxGroups; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int ixj = t ^ j;
      if (ixj > t) {
        const uint32_t a = s_keys[t], b = s_keys[ixj];
        const bool up = (t & k) == 0;
        if ((a > b) == up) {
          s_keys[t] = b;
          s_keys[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
  // rank of each key = its position; write it next to the key's table slot
  const uint32_t k = s_keys[t];
  if (k != kGsEmpty) {
    uint32_t h = gs_hash(k) & (kGsTableCap - 1);
    while (tkeys[h] != k) h = (h + 1) & (kGsTableCap - 1);
    tids[h] = (uint32_t)t;
  }
}

__device__ __forceinline__ int gs_lookup(const uint32_t* __restrict__ tkeys, const uint32_t* __restrict__ tids, uint32_t k) {
  uint32_t h = gs_hash(k) & (kGsTableCap - 1);
  for (int probe = 0; probe < kGsTableCap; ++probe) {
    const uint32_t cur = __ldg(tkeys + h);
    if (cur == k) return (int)__ldg(tids + h);
    if (cur == kGsEmpty) break;
    h = (h + 1) & (kGsTableCap - 1);
  }
  return 0;  // more than kGsMaxGroups distinct keys (status bit set by k_gs_rank): keep indices in range
}

// block = 4 warps, warp w handles rows [(4 * blockIdx.x + w) * 1024, +1024); hist[(warp block) * kGsMaxGroups + g].  The 4096-slot
// table is copied to shared memory first: a warp makes 32 dependent steps, and with the probes going to L2 the kernel sat at
// 6 % of the warp slots for 33 us per 889 k rows (ncu, profiles/r02_families_infer_summary.txt).
__global__ void __launch_bounds__(128) k_gs_hist(const uint32_t* __restrict__ keys, int64_t n, const uint32_t* __restrict__ tkeys,
                                                 const uint32_t* __restrict__ tids, const int32_t* __restrict__ meta, int16_t* __restrict__ gid,
                                                 int32_t* __restrict__ hist) {
  __shared__ int s_hist[4][kGsMaxGroups];
  __shared__ uint32_t s_tk[kGsTableCap];
  __shared__ uint16_t s_ti[kGsTableCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = meta[0];
  for (int i = threadIdx.x; i < kGsTableCap; i += 128) {
    s_tk[i] = tkeys[i];
    s_ti[i] = (uint16_t)tids[i];
  }
  for (int g = lane; g < D; g += 32) s_hist[warp][g] = 0;
  __syncthreads();
  const int64_t wb = (int64_t)blockIdx.x * 4 + warp;
  const int64_t r0 = wb * kGsRowsPerWarp;
  if (r0 >= n) return;
  auto lookup = [&](uint32_t k) -> int {
    uint32_t h = gs_hash(k) & (kGsTableCap - 1);
    for (int probe = 0; probe < kGsTableCap; ++probe) {
      const uint32_t cur = s_tk[h];
      if (cur == k) return (int)s_ti[h];
      if (cur == kGsEmpty) break;
      h = (h + 1) & (kGsTableCap - 1);
    }
    return 0;  // more than kGsMaxGroups distinct keys (status bit set by k_gs_rank): keep indices in range
  };
  for (int step = 0; step < kGsRowsPerWarp / 32; step += 4) {
    uint32_t kk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = r0 + (step + u) * 32 + lane;
      kk[u] = i < n ? keys[i] : kGsEmpty;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = r0 + (step + u) * 32 + lane;
      const bool in = i < n;
      const int g = in ? lookup(kk[u]) : -1;
      if (in) gid[i] = (int16_t)g;
      const unsigned peers = __match_any_sync(0xffffffffu, g);
      if (in && (int)(__ffs(peers) - 1) == lane) s_hist[warp][g] += __popc(peers);
      __syncwarp();
    }
  }
  for (int g = lane; g < D; g += 32) hist[wb * kGsMaxGroups + g] = s_hist[warp][g];
}

// grid = kGsMaxGroups blocks; block g scans hist[b][g] over the warp blocks b (exclusive, in place) and writes total[g]
__global__ void __launch_bounds__(256) k_gs_scan(int32_t* __restrict__ hist, int64_t nblocks, const int32_t* __restrict__ meta,
                                                 int32_t* __restrict__ total) {
  const int g = blockIdx.x;
  if (g >= meta[0]) return;
  __shared__ int s_part[256];
  const int t = threadIdx.x;
  const int64_t per = (nblocks + 255) / 256;
  const int64_t b0 = t * per, b1 = min(nblocks, b0 + per);
  int sum = 0;
  for (int64_t b = b0; b < b1; ++b) sum += hist[b * kGsMaxGroups + g];
  s_part[t] = sum;
  __syncthreads();
  // exclusive scan of the 256 partial sums (Hillis-Steele on shared memory)
  for (int d = 1; d < 256; d <<= 1) {
    const int v = t >= d ? s_part[t - d] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  int run = s_part[t] - sum;
  if (t == 255) total[g] = s_part[255];
  for (int64_t b = b0; b < b1; ++b) {
    const int c = hist[b * kGsMaxGroups + g];
    hist[b * kGsMaxGroups + g] = run;
    run += c;
  }
}

// one block of 1024 threads: base[g] = exclusive scan of total[g]
__global__ void __launch_bounds__(1024) k_gs_base(const int32_t* __restrict__ total, const int32_t* __restrict__ meta, int32_t* __restrict__ base) {
  __shared__ int s[kGsMaxGroups];
  const int t = threadIdx.x;
  const int D = meta[0];
  const int mine = t < D ? total[t] : 0;
  s[t] = mine;
  __syncthreads();
  for (int d = 1; d < kGsMaxGroups; d <<= 1) {
    const int v = t >= d ? s[t - d] : 0;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  if (t < D) base[t] = s[t] - mine;
}

__global__ void __launch_bounds__(128) k_gs_scatter(const int16_t* __restrict__ gid, int64_t n, const int32_t* __restrict__ meta,
                                                    const int32_t* __restrict__ hist, const int32_t* __restrict__ base,
                                                    int32_t* __restrict__ order) {
  __shared__ int s_off[4][kGsMaxGroups];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = meta[0];
  const int64_t wb = (int64_t)blockIdx.x * 4 + warp;
  const int64_t r0 = wb * kGsRowsPerWarp;
  if (r0 >= n) return;
  for (int g = lane; g < D; g += 32) s_off[warp][g] = base[g] + hist[wb * kGsMaxGroups + g];
  __syncwarp();
  for (int step = 0; step < kGsRowsPerWarp / 32; step += 4) {
    int gg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // the four loads are independent: one memory latency per four steps
      const int64_t i = r0 + (step + u) * 32 + lane;
      gg[u] = i < n ? (int)gid[i] : -1;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = r0 + (step + u) * 32 + lane;
      const bool in = i < n;
      const int g = gg[u];
      const unsigned peers = __match_any_sync(0xffffffffu, g);
      if (in) {
        const int rank = __popc(peers & ((1u << lane) - 1u));  // earlier rows of this step with the same group
        order[s_off[warp][g] + rank] = (int32_t)i;
      }
      __syncwarp();
      if (in && (int)(__ffs(peers) - 1) == lane) s_off[warp][g] += __popc(peers);
      __syncwarp();
    }
  }
}

}  // namespace mk

using namespace mk;

extern "C" {

int64_t mk_group_rows_workspace_bytes(int64_t n) {
  const int64_t wblocks = (n + kGsRowsPerWarp - 1) / kGsRowsPerWarp + 4;
  return (int64_t)kGsTableCap * 8 + 64 + wblocks * kGsMaxGroups * 4 + (int64_t)kGsMaxGroups * 8 + ((n * 2 + 255) / 256) * 256 + 1024;
}

int mk_group_rows(const uint32_t* keys, int64_t n, int32_t* order, int32_t* status, void* workspace, int64_t workspace_bytes, void* stream_) {
  MK_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
  MK_REQUIRE(workspace_bytes >= mk_group_rows_workspace_bytes(n), "workspace too small");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  uint8_t* w = (uint8_t*)workspace;
  uint32_t* tkeys = (uint32_t*)w;
  w += (size_t)kGsTableCap * 4;
  uint32_t* tids = (uint32_t*)w;
  w += (size_t)kGsTableCap * 4;
  int32_t* meta = (int32_t*)w;  // [0] = number of groups, [1] = status
  w += 64;
  const int64_t wblocks = (n + kGsRowsPerWarp - 1) / kGsRowsPerWarp;
  int32_t* hist = (int32_t*)w;
  w += (size_t)(wblocks + 4) * kGsMaxGroups * 4;
  int32_t* total = (int32_t*)w;
  w += (size_t)kGsMaxGroups * 4;
  int32_t* base = (int32_t*)w;
  w += (size_t)kGsMaxGroups * 4;
  int16_t* gid = (int16_t*)w;
  cudaMemsetAsync(tkeys, 0xFF, (size_t)kGsTableCap * 4, st);
  cudaMemsetAsync(tids, 0, (size_t)kGsTableCap * 4, st);  // keys beyond the 1024 ranked ones (status bit 0) fall into group 0
  cudaMemsetAsync(meta, 0, 64, st);
  int64_t bx = ceil_div(n, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  k_gs_insert<<<(unsigned)bx, 256, 0, st>>>(keys, n, tkeys);
  k_gs_rank<<<1, 1024, 0, st>>>(tkeys, tids, meta);
  const unsigned hb = (unsigned)ceil_div(wblocks, 4);
  k_gs_hist<<<hb, 128, 0, st>>>(keys, n, tkeys, tids, meta, gid, hist);
  k_gs_scan<<<kGsMaxGroups, 256, 0, st>>>(hist, wblocks, meta, total);
  k_gs_base<<<1, 1024, 0, st>>>(total, meta, base);
  k_gs_scatter<<<hb, 128, 0, st>>>(gid, n, meta, hist, base, order);
  if (status) cudaMemcpyAsync(status, meta + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
  return check_launch(__func__);
}

}  // extern "C"
