// Coordinate maps and kernel maps: GPU open-addressing hash over packed 64-bit keys.
//
// Replaces MinkowskiEngine 0.5.4 CoordinateMapGPU::{insert_and_map, stride, kernel_map, origin_map}
//  — HBM/L2-bound integer work: coalesced int4 coordinate loads, one random probe
// per lookup into an L2-resident table, coalesced int32 stores.  Results are deterministic: the winner
// of duplicate keys is the lowest source row (atomicMin), and new rows are ranked by an exclusive scan
// over first-occurrence flags, i.e. ME's CPU order, never ME's racy GPU order.
#include "common.cuh"

namespace mk {

thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------------------
// exclusive scan of int32 flags (three small kernels; n <= 2^31)
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total_out) {
  // returns exclusive prefix of v within the block; total in *total_out (all threads)
  __shared__ int warp_sums[kScanThreads / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;
    if (lane == kScanThreads / 32 - 1) block_total = wi;
  }
  __syncthreads();
  int res = incl - v + warp_sums[warp];
  *total_out = block_total;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_block_sums(const int32_t* __restrict__ in, int64_t n,
                                                                  int32_t* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_sums(int32_t* __restrict__ block_sums, int nblocks,
                                                            int32_t* __restrict__ total_out) {
  // single block: sequential over chunks of kScanThreads
  int carry = 0;
  for (int base = 0; base < nblocks; base += kScanThreads) {
    int idx = base + threadIdx.x;
    int v = idx < nblocks ? block_sums[idx] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    if (idx < nblocks) block_sums[idx] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const int32_t* __restrict__ in, int64_t n,
                                                             const int32_t* __restrict__ block_sums,
                                                             int32_t* __restrict__ out) {
  // thread-contiguous items so that the scan order is the row order
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  int total;
  int ex = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// hash build
// ------------------------------------------------------------------------------------------------
__global__ void k_hash_insert(const int4* __restrict__ coords, int64_t n, int sx, int sy, int sz,
                              unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals, uint64_t mask,
                              int32_t* __restrict__ slot_of, int32_t* __restrict__ status) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int4 c = coords[r];
  // status bit 1: the batch column is not ascending -- the scene row ranges (scene_offsets, decomposed_*, the instance-norm
  // segments) assume batch-sorted rows, which is what sparse_collate produces (datasets/utils.py:126)
  if (r > 0 && coords[r - 1].x > c.x) atomicOr(status, 2);
  if (sx > 1) c.y = floor_div(c.y, sx) * sx;
  if (sy > 1) c.z = floor_div(c.z, sy) * sy;
  if (sz > 1) c.w = floor_div(c.w, sz) * sz;
  if (!coord_in_range(c.x, c.y, c.z, c.w)) {
    atomicOr(status, 1);
    slot_of[r] = -1;
    return;
  }
  const unsigned long long key = pack_key(c.x, c.y, c.z, c.w);
  uint64_t h = hash_key(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&keys[h], kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) {
      atomicMin(&vals[h], (uint32_t)r);
      slot_of[r] = (int32_t)h;
      return;
    }
    h = (h + 1) & mask;
  }
}

__global__ void k_hash_flag(int64_t n, const int32_t* __restrict__ slot_of, const uint32_t* __restrict__ vals,
                            int32_t* __restrict__ flag) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int s = slot_of[r];
  flag[r] = (s >= 0 && vals[s] == (uint32_t)r) ? 1 : 0;
}

__global__ void k_hash_finalize(const int4* __restrict__ coords, int64_t n, int sx, int sy, int sz,
                                const int32_t* __restrict__ slot_of, const uint32_t* __restrict__ vals,
                                const int32_t* __restrict__ rank, const int32_t* __restrict__ flag,
                                int4* __restrict__ out_coords, int32_t* __restrict__ unique_index,
                                int32_t* __restrict__ inverse) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int s = slot_of[r];
  if (s < 0) {
    if (inverse) inverse[r] = -1;
    return;
  }
  if (inverse) inverse[r] = rank[vals[s]];
  if (flag[r]) {
    int4 c = coords[r];
    if (sx > 1) c.y = floor_div(c.y, sx) * sx;
    if (sy > 1) c.z = floor_div(c.z, sy) * sy;
    if (sz > 1) c.w = floor_div(c.w, sz) * sz;
    int o = rank[r];
    out_coords[o] = c;
    if (unique_index) unique_index[o] = (int32_t)r;
  }
}

__global__ void k_hash_setvals(int64_t n, const int32_t* __restrict__ slot_of, const int32_t* __restrict__ rank,
                               const int32_t* __restrict__ flag, uint32_t* __restrict__ vals) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (flag[r]) vals[slot_of[r]] = (uint32_t)rank[r];
}

__device__ __forceinline__ int hash_lookup(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals,
                                           uint64_t mask, unsigned long long key) {
  uint64_t h = hash_key(key) & mask;
  while (true) {
    unsigned long long k = keys[h];
    if (k == key) return vals[h];
    if (k == kEmptyKey) return -1;
    h = (h + 1) & mask;
  }
}

__global__ void k_hash_query(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals, uint64_t mask,
                             const int4* __restrict__ q, int64_t n, int32_t* __restrict__ rows) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int4 c = q[r];
  rows[r] = coord_in_range(c.x, c.y, c.z, c.w) ? hash_lookup(keys, vals, mask, pack_key(c.x, c.y, c.z, c.w)) : -1;
}

// ------------------------------------------------------------------------------------------------
// kernel map: nbr[k, o]
// ------------------------------------------------------------------------------------------------
__global__ void k_kernel_map(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals, uint64_t mask,
                             const int4* __restrict__ out_coords, int64_t n_out, const int32_t* __restrict__ offsets,
                             int32_t* __restrict__ nbr) {
  const int k = blockIdx.y;
  const int ox = offsets[3 * k], oy = offsets[3 * k + 1], oz = offsets[3 * k + 2];
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += (int64_t)gridDim.x * blockDim.x) {
    int4 c = out_coords[o];
    c.y += ox;
    c.z += oy;
    c.w += oz;
    int row = -1;
    if (coord_in_range(c.x, c.y, c.z, c.w)) row = hash_lookup(keys, vals, mask, pack_key(c.x, c.y, c.z, c.w));
    nbr[(int64_t)k * n_out + o] = row;
  }
}

// Same-map kernel map (in map == out map, odd kernel: offsets[K-1-k] == -offsets[k]): nbr[k, o] = i  <=>  nbr[K-1-k, i] = o, so
// only the first K/2 offsets are probed; a hit writes both entries, the centre offset is the row itself, and the upper half is
// pre-filled with -1 by the caller.  Each thread takes kSymK offsets of one row (one read of the row's coordinates per kSymK
// probes).  Bit-identical to k_kernel_map; half the hash probes (the 5^3 stem map: 62 instead of 125 per voxel).
constexpr int kSymK = 4;
__global__ void __launch_bounds__(256) k_kernel_map_sym(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals,
                                                        uint64_t mask, const int4* __restrict__ coords, int64_t n,
                                                        const int32_t* __restrict__ offsets, int K, int32_t* __restrict__ nbr) {
  const int half = K >> 1;
  const int k0 = blockIdx.y * kSymK;
  int ox[kSymK], oy[kSymK], oz[kSymK];
#pragma unroll
  for (int j = 0; j < kSymK; ++j) {
    const int k = min(k0 + j, half);
    ox[j] = offsets[3 * k], oy[j] = offsets[3 * k + 1], oz[j] = offsets[3 * k + 2];
  }
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = coords[o];
#pragma unroll
    for (int j = 0; j < kSymK; ++j) {
      const int k = k0 + j;
      if (k > half) break;
      if (k == half) {
        nbr[(int64_t)k * n + o] = (int)o;
        break;
      }
      const int x = c.y + ox[j], y = c.z + oy[j], z = c.w + oz[j];
      int row = -1;
      if (coord_in_range(c.x, x, y, z)) row = hash_lookup(keys, vals, mask, pack_key(c.x, x, y, z));
      nbr[(int64_t)k * n + o] = row;
      if (row >= 0) nbr[(int64_t)(K - 1 - k) * n + row] = (int)o;
    }
  }
}

// 5^3 same-map kernel map COMPOSED from the 3^3 table of the same map: a target d with |d|_inf = 2 is reached through an
// intermediate voxel o + s (|s|_inf = 1, s takes the sign of every axis with |d_a| = 2) and the remainder r = d - s (|r|_inf <= 1):
// if the intermediate exists, nbr5[d][o] = nbr3[r][nbr3[s][o]] EXACTLY (also when that is -1), two array reads instead of a hash
// probe.  The axes with |d_a| = 1 may go into s or into r, which gives up to four intermediates to try; only when none of them
// exists is the hash table probed.  Same symmetric half / reverse write scheme as k_kernel_map_sym, bit-identical result.  On
// surface scenes > 90 % of the 62 probes per voxel of the stem's 5^3 map become table reads (the probes are random DRAM sector
// reads once the table of a batch no longer fits in L2: 5.4 of the 7.2 ms of kernel-map time in a batch-of-8 step).
__global__ void __launch_bounds__(256) k_kernel_map_compose5(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals,
                                                             uint64_t mask, const int4* __restrict__ coords, int64_t n, int sx, int sy, int sz,
                                                             const int32_t* __restrict__ nbr3, int32_t* __restrict__ nbr5, int sym) {
  const int k = blockIdx.y;  // sym: 0 .. 62 = lower half of the 125 offsets + the centre, hits also write the mirrored entry; else all 125
  const int ix = k % 5 - 2, iy = (k / 5) % 5 - 2, iz = k / 25 - 2;
  const int ax = ix < 0 ? -ix : ix, ay = iy < 0 ? -iy : iy, az = iz < 0 ? -iz : iz;
  auto k3 = [](int x, int y, int z) { return (x + 1) + 3 * (y + 1) + 9 * (z + 1); };
  // forced part of the intermediate step (axes at distance 2) and the optional axes (distance 1)
  const int fx = ax == 2 ? (ix > 0 ? 1 : -1) : 0, fy = ay == 2 ? (iy > 0 ? 1 : -1) : 0, fz = az == 2 ? (iz > 0 ? 1 : -1) : 0;
  const int ox = ax == 1 ? ix : 0, oy = ay == 1 ? iy : 0, oz = az == 1 ? iz : 0;
  const bool direct = ax <= 1 && ay <= 1 && az <= 1;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    if (k == 62) {
      nbr5[(int64_t)k * n + o] = (int)o;
      continue;
    }
    int row = -1;
    if (direct) {
      row = nbr3[(int64_t)k3(ix, iy, iz) * n + o];
    } else {
      bool resolved = false;
#pragma unroll
      for (int c = 0; c < 8 && !resolved; ++c) {  // subsets of the optional axes that go into the intermediate step
        const int tx = (c & 1) ? ox : 0, ty = (c & 2) ? oy : 0, tz = (c & 4) ? oz : 0;
        if (((c & 1) && !ox) || ((c & 2) && !oy) || ((c & 4) && !oz)) continue;  // not a distinct subset
        const int s0 = fx + tx, s1 = fy + ty, s2 = fz + tz;
        const int m = nbr3[(int64_t)k3(s0, s1, s2) * n + o];
        if (m >= 0) {
          row = nbr3[(int64_t)k3(ix - s0, iy - s1, iz - s2) * n + m];
          resolved = true;
        }
      }
      if (!resolved) {
        const int4 cc = coords[o];
        const int x = cc.y + ix * sx, y = cc.z + iy * sy, z = cc.w + iz * sz;
        if (coord_in_range(cc.x, x, y, z)) row = hash_lookup(keys, vals, mask, pack_key(cc.x, x, y, z));
      }
    }
    nbr5[(int64_t)k * n + o] = row;
    if (sym && row >= 0) nbr5[(int64_t)(124 - k) * n + row] = (int)o;
  }
}

__global__ void __launch_bounds__(MK_TILE_ROWS) k_tile_kmask(const int32_t* __restrict__ nbr, int K, int64_t n_slots,
                                                             uint32_t* __restrict__ tile_kmask,
                                                             unsigned long long* __restrict__ pair_count) {
  const int64_t slot = (int64_t)blockIdx.x * MK_TILE_ROWS + threadIdx.x;
  uint32_t m = 0;
  int cnt = 0;
  for (int k = 0; k < K; ++k) {
    bool present = slot < n_slots && nbr[(int64_t)k * n_slots + slot] >= 0;
    cnt += present;
    if (__syncthreads_or(present) && k < 32) m |= 1u << k;
  }
  // block pair count
  __shared__ int wsum[MK_TILE_ROWS / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, d);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < MK_TILE_ROWS / 32; ++i) t += wsum[i];
    if (tile_kmask) tile_kmask[blockIdx.x] = m;
    if (pair_count && t) atomicAdd(pair_count, (unsigned long long)t);
  }
}

// per-slot sort key for mask-sorted tile plans:  (offset bitmask) << 32 | (batch & 3) << 30 | 30-bit Morton code.
// bit k of the mask is set iff nbr[k, slot] >= 0 (K <= 32).  Slots with equal masks are grouped into the same tensor-core
// tiles (a tile's active-offset set then equals its rows' set) and, inside a mask group, ordered along a Z-curve so that a
// tile is a compact spatial patch whose neighbour gathers overlap (L1 / L2 reuse).
__device__ __forceinline__ unsigned spread3(unsigned v) {  // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void k_row_sort_keys(const int32_t* __restrict__ nbr, int K, int64_t n, const int4* __restrict__ coords, int sx, int sy,
                                int sz, unsigned long long* __restrict__ keys) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    unsigned m = 0;
    for (int k = 0; k < K; ++k) m |= (nbr[(int64_t)k * n + o] >= 0 ? 1u : 0u) << k;
    const int4 c = coords[o];
    const unsigned morton = spread3((unsigned)floor_div(c.y, sx)) | (spread3((unsigned)floor_div(c.z, sy)) << 1) |
                            (spread3((unsigned)floor_div(c.w, sz)) << 2);
    keys[o] = ((unsigned long long)m << 32) | ((unsigned long long)((unsigned)c.x & 3u) << 30) | (unsigned long long)morton;
  }
}

// mask-only sort key (int32): a STABLE sort by it groups the slots by offset mask and leaves the rows of a group in their
// original (lexicographic voxel) order — measured as good for the gathers as the Z-curve order, with a 32-bit instead of a
// 64-bit radix sort (4 instead of 8 passes over half the bytes).
__global__ void k_row_mask_keys(const int32_t* __restrict__ nbr, int K, int64_t n, int32_t* __restrict__ keys) {
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
    unsigned m = 0;
    for (int k = 0; k < K; ++k) m |= (nbr[(int64_t)k * n + o] >= 0 ? 1u : 0u) << k;
    keys[o] = (int32_t)m;
  }
}

// nbr_sorted[k, s] = nbr[k, order[s]]
// Lexicographic sort keys of (batch, x, y, z) rows: the packed 64-bit coordinate key (biased, so unsigned key order ==
// signed lexicographic row order, what np.unique(axis=0) sorts by).  status |= 1 for rows outside the packable range.
__global__ void k_pack_keys(const int4* __restrict__ coords, int64_t n, unsigned long long* __restrict__ keys, int32_t* __restrict__ status) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int4 c = coords[i];
    if (!coord_in_range(c.x, c.y, c.z, c.w)) {
      *status = 1;
      keys[i] = kEmptyKey;
    } else {
      keys[i] = pack_key(c.x, c.y, c.z, c.w);
    }
  }
}

__global__ void k_permute_table(const int32_t* __restrict__ nbr, int64_t n, const int32_t* __restrict__ order, int64_t n_slots,
                                int32_t* __restrict__ out) {
  const int k = blockIdx.y;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
    const int o = order[s];
    out[(int64_t)k * n_slots + s] = o >= 0 ? nbr[(int64_t)k * n + o] : -1;
  }
}

__global__ void k_transpose_map(const int32_t* __restrict__ nbr_down, int K, int64_t n_coarse,
                                int32_t* __restrict__ parent, int32_t* __restrict__ koff) {
  const int k = blockIdx.y;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_coarse; c += (int64_t)gridDim.x * blockDim.x) {
    int f = nbr_down[(int64_t)k * n_coarse + c];
    if (f >= 0) {
      parent[f] = (int32_t)c;
      koff[f] = k;
    }
  }
}

}  // namespace mk

using namespace mk;

extern "C" {

const char* mk_last_error(void) { return g_last_error.c_str(); }
int mk_version(void) { return 1; }

int mk_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? 1 : 0;
}

int64_t mk_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

int64_t mk_coordmap_workspace_bytes(int64_t n) {
  // slot_of[n] + flag[n] + rank[n] + block sums
  int64_t nb = ceil_div(n > 0 ? n : 1, kScanTile);
  return (3 * n + nb + 64) * (int64_t)sizeof(int32_t);
}

int mk_coordmap_build(const int32_t* in_coords, int64_t n, const int32_t* stride3_host, uint64_t* keys, int32_t* vals,
                      int64_t cap, int32_t* out_coords, int32_t* unique_index, int32_t* inverse, int32_t* n_out,
                      int32_t* status, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
  MK_REQUIRE(cap >= 2 * n && (cap & (cap - 1)) == 0, "cap must be a power of two >= 2n");
  MK_REQUIRE(workspace_bytes >= mk_coordmap_workspace_bytes(n), "workspace too small");
  const int sx = stride3_host ? stride3_host[0] : 1, sy = stride3_host ? stride3_host[1] : 1,
            sz = stride3_host ? stride3_host[2] : 1;
  MK_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1, "stride must be >= 1");
  cudaMemsetAsync(keys, 0xFF, (size_t)cap * sizeof(uint64_t), st);
  cudaMemsetAsync(vals, 0xFF, (size_t)cap * sizeof(int32_t), st);
  cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  if (n == 0) {
    cudaMemsetAsync(n_out, 0, sizeof(int32_t), st);
    return check_launch(__func__);
  }
  int32_t* slot_of = (int32_t*)workspace;
  int32_t* flag = slot_of + n;
  int32_t* rank = flag + n;
  int32_t* bsums = rank + n;
  const int threads = 256;
  const unsigned blocks = (unsigned)ceil_div(n, threads);
  const unsigned sblocks = (unsigned)ceil_div(n, kScanTile);
  k_hash_insert<<<blocks, threads, 0, st>>>((const int4*)in_coords, n, sx, sy, sz, (unsigned long long*)keys,
                                             (uint32_t*)vals, (uint64_t)(cap - 1), slot_of, status);
  k_hash_flag<<<blocks, threads, 0, st>>>(n, slot_of, (const uint32_t*)vals, flag);
  k_scan_block_sums<<<sblocks, kScanThreads, 0, st>>>(flag, n, bsums);
  k_scan_sums<<<1, kScanThreads, 0, st>>>(bsums, (int)sblocks, n_out);
  k_scan_apply<<<sblocks, kScanThreads, 0, st>>>(flag, n, bsums, rank);
  k_hash_finalize<<<blocks, threads, 0, st>>>((const int4*)in_coords, n, sx, sy, sz, slot_of, (const uint32_t*)vals, rank,
                                               flag, (int4*)out_coords, unique_index, inverse);
  k_hash_setvals<<<blocks, threads, 0, st>>>(n, slot_of, rank, flag, (uint32_t*)vals);
  return check_launch(__func__);
}

int mk_hash_query(const uint64_t* keys, const int32_t* vals, int64_t cap, const int32_t* query_coords, int64_t n,
                  int32_t* rows, void* stream_) {
  if (n == 0) return 0;
  k_hash_query<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream_>>>(
      (const unsigned long long*)keys, vals, (uint64_t)(cap - 1), (const int4*)query_coords, n, rows);
  return check_launch(__func__);
}

int mk_kernel_map(const uint64_t* in_keys, const int32_t* in_vals, int64_t cap, const int32_t* out_coords, int64_t n_out,
                  const int32_t* offsets, int K, int32_t* nbr, void* stream_) {
  MK_REQUIRE(K >= 1 && K <= 65535, "K out of range");
  if (n_out == 0) return 0;
  int64_t bx = ceil_div(n_out, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  dim3 grid((unsigned)bx, (unsigned)K);
  k_kernel_map<<<grid, 256, 0, (cudaStream_t)stream_>>>((const unsigned long long*)in_keys, in_vals, (uint64_t)(cap - 1),
                                                         (const int4*)out_coords, n_out, offsets, nbr);
  return check_launch(__func__);
}

int mk_kernel_map_same(const uint64_t* keys, const int32_t* vals, int64_t cap, const int32_t* coords, int64_t n, const int32_t* offsets,
                       int K, int32_t* nbr, void* stream_) {
  MK_REQUIRE(K >= 1 && K <= 65535 && (K & 1) == 1, "same-map kernel maps have an odd number of offsets");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  const int half = K >> 1;
  if (half > 0) cudaMemsetAsync(nbr + (int64_t)(half + 1) * n, 0xFF, (size_t)half * n * sizeof(int32_t), st);
  int64_t bx = ceil_div(n, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  dim3 grid((unsigned)bx, (unsigned)ceil_div(half + 1, kSymK));
  k_kernel_map_sym<<<grid, 256, 0, st>>>((const unsigned long long*)keys, vals, (uint64_t)(cap - 1), (const int4*)coords, n, offsets, K, nbr);
  return check_launch(__func__);
}

int mk_kernel_map_compose5(const uint64_t* keys, const int32_t* vals, int64_t cap, const int32_t* coords, int64_t n, const int32_t* step3_host,
                           const int32_t* nbr3, int32_t* nbr5, void* stream_) {
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  // lower half + mirrored writes (sym = 1): computing all 125 entries independently (no mirrored scatter, no pre-fill) was measured
  // at 7.9 ms against 4.1 ms for a batch of 7.1 M voxels -- the fallback probes of the directions normal to the surfaces dominate
  // and they double
  const int sym = 1;
  if (sym) cudaMemsetAsync(nbr5 + (int64_t)63 * n, 0xFF, (size_t)62 * n * sizeof(int32_t), st);
  int64_t bx = ceil_div(n, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  dim3 grid((unsigned)bx, sym ? 63u : 125u);
  k_kernel_map_compose5<<<grid, 256, 0, st>>>((const unsigned long long*)keys, vals, (uint64_t)(cap - 1), (const int4*)coords, n, step3_host[0],
                                              step3_host[1], step3_host[2], nbr3, nbr5, sym);
  return check_launch(__func__);
}

int mk_tile_kmask(const int32_t* nbr, int K, int64_t n_slots, uint32_t* tile_kmask, int64_t* pair_count, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (pair_count) cudaMemsetAsync(pair_count, 0, sizeof(int64_t), st);
  if (n_slots == 0) return check_launch(__func__);
  k_tile_kmask<<<(unsigned)ceil_div(n_slots, MK_TILE_ROWS), MK_TILE_ROWS, 0, st>>>(nbr, K, n_slots, tile_kmask,
                                                                                   (unsigned long long*)pair_count);
  return check_launch(__func__);
}

int mk_row_sort_keys(const int32_t* nbr, int K, int64_t n, const int32_t* coords, const int32_t* stride3_host, uint64_t* keys,
                     void* stream_) {
  MK_REQUIRE(K >= 1 && K <= 32, "K must be <= 32");
  if (n == 0) return 0;
  int64_t bx = ceil_div(n, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  const int sx = stride3_host ? stride3_host[0] : 1, sy = stride3_host ? stride3_host[1] : 1, sz = stride3_host ? stride3_host[2] : 1;
  k_row_sort_keys<<<(unsigned)bx, 256, 0, (cudaStream_t)stream_>>>(nbr, K, n, (const int4*)coords, sx, sy, sz,
                                                                  (unsigned long long*)keys);
  return check_launch(__func__);
}

int mk_row_mask_keys(const int32_t* nbr, int K, int64_t n, int32_t* keys, void* stream_) {
  MK_REQUIRE(K >= 1 && K <= 32, "K must be <= 32");
  if (n == 0) return 0;
  int64_t bx = ceil_div(n, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  k_row_mask_keys<<<(unsigned)bx, 256, 0, (cudaStream_t)stream_>>>(nbr, K, n, keys);
  return check_launch(__func__);
}

int mk_pack_keys(const int32_t* coords, int64_t n, uint64_t* keys, int32_t* status, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  if (n == 0) return check_launch(__func__);
  int64_t bx = ceil_div(n, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  k_pack_keys<<<(unsigned)bx, 256, 0, st>>>((const int4*)coords, n, (unsigned long long*)keys, status);
  return check_launch(__func__);
}

int mk_permute_table(const int32_t* nbr, int K, int64_t n, const int32_t* order, int64_t n_slots, int32_t* out, void* stream_) {
  if (n_slots == 0) return 0;
  int64_t bx = ceil_div(n_slots, 256);
  if (bx > 8 * num_sms()) bx = 8 * num_sms();
  dim3 grid((unsigned)bx, (unsigned)K);
  k_permute_table<<<grid, 256, 0, (cudaStream_t)stream_>>>(nbr, n, order, n_slots, out);
  return check_launch(__func__);
}

int mk_transpose_map(const int32_t* nbr_down, int K, int64_t n_coarse, int64_t n_fine, int32_t* parent, int32_t* koff,
                     void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  cudaMemsetAsync(parent, 0xFF, (size_t)n_fine * sizeof(int32_t), st);
  cudaMemsetAsync(koff, 0xFF, (size_t)n_fine * sizeof(int32_t), st);
  if (n_coarse == 0) return check_launch(__func__);
  int64_t bx = ceil_div(n_coarse, 256);
  dim3 grid((unsigned)bx, (unsigned)K);
  k_transpose_map<<<grid, 256, 0, st>>>(nbr_down, K, n_coarse, parent, koff);
  return check_launch(__func__);
}

}  // extern "C"
