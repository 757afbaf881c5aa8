// tcgen05 weight-gradient of the sparse convolution (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
//   dW[k][ci][co] = sum over output slots o with nbr[k,o] >= 0 of  x[nbr[k,o]][ci] * dy[row(o)][co]
//
// i.e. per kernel offset k one GEMM  D[M = Cin, N = Cout] = A[M, pairs] * B[pairs, N]  whose reduction runs over the PAIRS of
// the kernel map.  ME computes it as K x (gather, gather, GEMM) launches (ConvolutionBackwardGPU); here it is ONE persistent,
// warp-specialised kernel over the same mask-sorted tile plan the forward kernel uses:
//
//   work item  = (tile block of T tiles, offset k): the tiles of the block that have offset k.  Items are ordered by (block, k)
//                and dealt round-robin to the CTAs, so the grid works on a few neighbouring tile blocks at a time (their
//                x / dy rows stay in L2 while all their offsets are processed) and a CTA keeps ONE accumulator alive across
//                the tiles of an item and flushes it once per item
//   stage      = 64 pair rows: the gathered x rows (A) and the matching dy rows (B), 16-byte cp.async (LDGSTS, zero-fill for
//                absent pairs) into 128-byte rows with the 128B swizzle.  Both operands have the reduction index as their ROW,
//                i.e. they are MN-major for the tensor core; the MN-major SWIZZLE_128B canonical layout
//                ((8,n),(8,k)) : ((1,LBO),(8,SBO)) [16-byte units] is exactly "rows of 128 bytes, 16-byte piece XOR (row & 7)",
//                the same staging the forward gather produces, with SBO = 1024 B (8 rows) and LBO = the distance between
//                64-channel chunks
//   MMA        = tcgen05.mma kind::f16, M = 128 (Cin block; rows beyond Cin are zero), N = Cout, K = 16 pair rows, both
//                operands MN-major (instruction-descriptor bits 15/16), accumulator fp32 in TMEM, double buffered
//   flush      = tcgen05.ld -> shared-memory transpose -> coalesced fp32 RED into dW (dW is zeroed by the entry point)
//
// Warp roles as in conv_umma.cu: warps 0-3 and 5 own whole stages round-robin (producers), warp 4 issues the MMAs,
// warps 6-9 flush.
#include "umma.cuh"

namespace mk {

constexpr int kWgRows = 64;                       // pair rows per stage
constexpr int kWgChunkBytes = kWgRows * 128;      // one 64-channel chunk of a stage: 8 KB
constexpr int kWgAChunks = 2;                     // Cin block = 128 channels
constexpr int kWgMaxBChunks = 4;                  // Cout <= 256
constexpr int kWgMaxProducers = 5;
constexpr int kWgItemShift = 26;                  // item = k << 26 | tile

struct WgradParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  const int32_t* nbr;
  const int32_t* out_rows;
  const int32_t* plan;  // see mk_conv_wgrad_plan
  float* dw;
  float* partial;     // deterministic mode: per-item accumulators [n_items * nmb][128][cout] (plain stores, merged in item order); else nullptr
  int64_t n_slots;
  int64_t n_out;
  int cin, cout, K;
  int nmb;            // 128-channel blocks of Cin
  int nb_chunks;      // ceil(cout / 64)
  int num_stages;
  int num_producers;
  int max_items;      // capacity of the item array inside `plan`
};

constexpr int kWgPlanHeader = 4;   // plan[0] = n_items, [1] = n_units, [2] = tiles per block
constexpr int kWgFlushPitch = 17;  // fp32 pitch of the per-warp flush staging rows (16 columns + 1: conflict free)

// instruction descriptor: as make_idesc, plus A and B MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) { return make_idesc(M, N) | (1u << 15) | (1u << 16); }

// Work item = (tile block b, offset k): the tiles of block b that have offset k, a contiguous run `units[start, start + cnt)`.
// Items are ordered by (b, k) and CTA c takes items c, c + grid, ...: at any time the grid works on ~grid consecutive items,
// i.e. on ~grid / (active offsets per block) neighbouring tile blocks, whose x / dy rows (a few tens of MB) stay in L2 while
// all their offsets are processed.  (Ordering by offset first made every offset sweep the whole 340 MB of x + dy through the
// 126 MB L2: ncu measured 3.0 GB of DRAM reads per launch against 0.34 GB of operands.)
struct WgItem {
  int start, cnt, k, mb;
};
__device__ __forceinline__ WgItem wg_load_item(const int32_t* __restrict__ items, int n_items, int64_t j) {
  WgItem it;
  it.mb = (int)(j / n_items);
  const int i = (int)(j - (int64_t)it.mb * n_items);
  const int2 v = __ldg(reinterpret_cast<const int2*>(items) + i);
  it.start = v.x;
  it.k = v.y >> kWgItemShift;
  it.cnt = v.y & ((1 << kWgItemShift) - 1);
  return it;
}

// Producer-side position in the CTA's concatenated iteration sequence (2 iterations = two 64-row halves per unit).
struct WgCursor {
  const int32_t* items;
  int n_items;
  int64_t J, j;
  int stride, local;
  WgItem it;
  __device__ WgCursor(const int32_t* items_, int n_items_, int64_t J_, int first, int stride_)
      : items(items_), n_items(n_items_), J(J_), j(first), stride(stride_), local(0) {
    it.start = it.cnt = it.k = it.mb = 0;
    if (j < J) it = wg_load_item(items, n_items, j);
  }
  __device__ bool valid() const { return j < J; }
  __device__ void advance(int n) {
    local += n;
    while (j < J && local >= 2 * it.cnt) {
      local -= 2 * it.cnt;
      j += stride;
      if (j < J) it = wg_load_item(items, n_items, j);
    }
  }
};

__global__ void __launch_bounds__(kNumThreads, 1) k_conv_wgrad_umma(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int S = p.num_stages;
  const int a_bytes = kWgAChunks * kWgChunkBytes;
  const int b_bytes = p.nb_chunks * kWgChunkBytes;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = (uint64_t*)(smem + (size_t)S * stage_bytes);
  uint64_t* full_bar = bars;                         // [S] 32 producer lanes (copy completion)
  uint64_t* empty_bar = bars + kMaxStages;           // [S] one tcgen05.commit
  uint64_t* tmem_full = bars + 2 * kMaxStages;       // [2] one tcgen05.commit
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2] 128 flush threads
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);
  float* flush_stage = (float*)(bars + 2 * kMaxStages + 6);  // [4 warps][32 rows][kWgFlushPitch]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the staging area once: pieces beyond a narrow Cin block are never written by the gathers but are read by the
  // M = 128 MMA (they only reach accumulator rows that are not flushed)
  for (int i = threadIdx.x; i < S * stage_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 32);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 4 * 32);
    }
    fence_barrier_init();
  }
  if (warp == kLoadWarp) {
    tmem_alloc(smem_u32(tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int n_items = p.plan[0];
  const int64_t J = (int64_t)n_items * p.nmb;
  const int32_t* items = p.plan + kWgPlanHeader;
  const int32_t* units = items + 2 * (int64_t)p.max_items;

  if (warp < kNumProducerWarps || warp == kLoadWarp) {
    // ===================== producers: iteration i of the CTA belongs to producer i mod NP =====================
    const int pi = warp < kNumProducerWarps ? warp : kNumProducerWarps;
    const int NP = p.num_producers;
    const int piece = lane & 7, rsub = lane >> 3;
    const uint32_t sw_even = (uint32_t)((piece ^ rsub) << 4), sw_odd = (uint32_t)((piece ^ (4 + rsub)) << 4);
    const uint32_t x_row_bytes = (uint32_t)p.cin * 2u, dy_row_bytes = (uint32_t)p.cout * 2u;
    struct Idx {
      int in[2], out[2];  // pair rows lane, lane + 32 of the stage: x row / dy row (-1 = absent -> zero fill)
      int mb;
    };
    auto load_idx = [&](const WgCursor& c) -> Idx {
      Idx r;
      r.mb = c.it.mb;
      r.in[0] = r.in[1] = r.out[0] = r.out[1] = -1;
      if (c.valid()) {
        const int64_t tile = __ldg(units + c.it.start + (c.local >> 1));
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int64_t slot = tile * kTileM + (c.local & 1) * kWgRows + q * 32 + lane;
          if (slot < p.n_slots) {
            const int in_row = __ldg(p.nbr + (int64_t)c.it.k * p.n_slots + slot);
            int64_t o = p.out_rows ? (int64_t)__ldg(p.out_rows + slot) : slot;
            if (o >= p.n_out) o = -1;
            if (in_row >= 0 && o >= 0) {
              r.in[q] = in_row;
              r.out[q] = (int)o;
            }
          }
        }
      }
      return r;
    };
    const bool active = pi < NP;
    WgCursor cur(items, n_items, J, blockIdx.x, gridDim.x);
    cur.advance(pi);
    WgCursor nxt = cur;
    nxt.advance(NP);
    Idx ci = load_idx(cur), ni = load_idx(nxt);
    int stage = pi % S;
    uint32_t phase = (uint32_t)((pi / S) & 1);
    while (active && cur.valid()) {
      WgCursor nn = nxt;
      nn.advance(NP);
      const Idx nni = load_idx(nn);
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      const uint32_t a_base = smem_u32(smem + (size_t)stage * stage_bytes);
      const uint32_t b_base = a_base + (uint32_t)a_bytes;
      const int cinb = min(128, p.cin - ci.mb * 128);  // valid channels of this Cin block
      uint64_t xsrc = (uint64_t)p.x + (uint64_t)(ci.mb * 128 + piece * 8) * 2u;
      uint64_t dsrc = (uint64_t)p.dy + (uint64_t)(piece * 8) * 2u;
      asm volatile("" : "+l"(xsrc));
      asm volatile("" : "+l"(dsrc));
      // all 32 index shuffles first, then the copies: `asm volatile` statements keep their order, so interleaving one SHFL per
      // group of copies made every group wait ~30 cycles on its own shuffle (ncu source view of the forward kernel, round 2)
      int ii[2][8], oo[2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          ii[q][jj] = __shfl_sync(0xffffffffu, ci.in[q], 4 * jj + rsub);
          oo[q][jj] = __shfl_sync(0xffffffffu, ci.out[q], 4 * jj + rsub);
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int row = q * 32 + 4 * jj + rsub;
          const uint32_t roff = (uint32_t)(row * 128) + ((jj & 1) ? sw_odd : sw_even);
          const uint32_t nbytes = ii[q][jj] < 0 ? 0u : 16u;
          const uint64_t xs = xsrc + (uint64_t)(ii[q][jj] < 0 ? 0u : (uint32_t)ii[q][jj]) * x_row_bytes;
          const uint64_t ds = dsrc + (uint64_t)(oo[q][jj] < 0 ? 0u : (uint32_t)oo[q][jj]) * dy_row_bytes;
#pragma unroll
          for (int c = 0; c < kWgAChunks; ++c)
            if (c * 64 + piece * 8 < cinb) cp_async_16(a_base + (uint32_t)(c * kWgChunkBytes) + roff, (const void*)(xs + c * 128), nbytes);
#pragma unroll
          for (int c = 0; c < kWgMaxBChunks; ++c)
            if (c * 64 + piece * 8 < p.cout) cp_async_16(b_base + (uint32_t)(c * kWgChunkBytes) + roff, (const void*)(ds + c * 128), nbytes);
        }
      }
      cp_async_arrive_noinc(bar);  // completion-triggered arrive, no thread-side wait
      __syncwarp();
      ci = ni;
      ni = nni;
      cur = nxt;
      nxt = nn;
      stage += NP;
      while (stage >= S) {
        stage -= S;
        phase ^= 1;
      }
    }
    cp_async_wait<0>();
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    Ring ring(S);
    const uint32_t idesc = make_idesc_mn(kTileM, p.cout);
    // MN-major descriptors: LBO = chunk stride (8 KB), SBO = 1024 B (8 pair rows)
    const uint32_t lbo_field = (uint32_t)(kWgChunkBytes >> 4) << 16;
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    for (int64_t j = blockIdx.x; j < J; j += gridDim.x) {
      const WgItem item = wg_load_item(items, n_items, j);
      mbar_wait(smem_u32(&tmem_empty[acc_buf]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc_buf * 256);
      uint32_t accumulate = 0;
      for (int it = 0; it < 2 * item.cnt; ++it) {
        mbar_wait(smem_u32(&full_bar[ring.stage]), ring.phase);
        fence_proxy_async_acquire();
        tc_fence_after();
        const uint32_t sbase = smem_u32(smem) + (uint32_t)(ring.stage * stage_bytes);
        const uint32_t a_lo = ((sbase >> 4) & 0x3FFFu) | lbo_field;
        const uint32_t b_lo = (((sbase + (uint32_t)a_bytes) >> 4) & 0x3FFFu) | lbo_field;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < kWgRows / 16; ++kk) {
            // 16 pair rows = 2048 bytes along the reduction
            umma_bf16_lohi(tmem_d, a_lo + (uint32_t)(kk * (2048 >> 4)), b_lo + (uint32_t)(kk * (2048 >> 4)), kSmemDescHi, idesc,
                           (kk > 0) ? 1u : accumulate);
          }
          umma_commit(smem_u32(&empty_bar[ring.stage]));
        }
        __syncwarp();
        accumulate = 1;
        ring.advance();
      }
      if (elect_one_sync()) umma_commit(smem_u32(&tmem_full[acc_buf]));
      __syncwarp();
      if (++acc_buf == 2) {
        acc_buf = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== flush (4 warps): TMEM -> shared-memory transpose -> coalesced fp32 RED into dW =====================
    const int q = warp & 3;
    const uint32_t stg = smem_u32(flush_stage) + (uint32_t)(q * 32 * kWgFlushPitch) * 4u;  // explicit shared-space accesses (umma.cuh)
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    for (int64_t j = blockIdx.x; j < J; j += gridDim.x) {
      const WgItem item = wg_load_item(items, n_items, j);
      mbar_wait(smem_u32(&tmem_full[acc_buf]), acc_phase);
      tc_fence_after();
      const int row0 = item.mb * 128 + q * 32;  // first Cin row of this warp's TMEM lane quarter
      for (int c0 = 0; c0 < p.cout; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_buf * 256 + c0), r);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) sts_f32(stg + (uint32_t)(lane * kWgFlushPitch + jj) * 4u, __uint_as_float(r[jj]));
        __syncwarp();
        // 32 rows x 16 columns as 16-byte vector REDs: lane = (row within a group of 8, 4-column quad) => one instruction adds
        // 8 rows x 64 contiguous bytes; 4 instructions per chunk instead of 16 scalar ones
        const int quad = lane & 3, rsel = lane >> 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + rsel;
          if (p.partial != nullptr) {
            // deterministic mode: this item's accumulator goes to its own slot with plain stores; k_wg_merge adds the slots of an
            // offset in item order (no floating-point atomics, run-to-run bit-identical)
            const uint32_t sv = stg + (uint32_t)(rr * kWgFlushPitch + 4 * quad) * 4u;
            float* dst = p.partial + ((int64_t)j * 128 + q * 32 + rr) * p.cout + c0 + 4 * quad;
            *reinterpret_cast<float4*>(dst) = make_float4(lds_f32(sv), lds_f32(sv + 4u), lds_f32(sv + 8u), lds_f32(sv + 12u));
          } else if (row0 + rr < p.cin) {
            const uint32_t sv = stg + (uint32_t)(rr * kWgFlushPitch + 4 * quad) * 4u;
            float* dst = p.dw + ((int64_t)item.k * p.cin + row0 + rr) * p.cout + c0 + 4 * quad;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(lds_f32(sv)), "f"(lds_f32(sv + 4u)),
                         "f"(lds_f32(sv + 8u)), "f"(lds_f32(sv + 12u))
                         : "memory");
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[acc_buf]));
      if (++acc_buf == 2) {
        acc_buf = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kLoadWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------
// plan: items (tile block, offset) and their unit lists, from a tile plan's per-tile offset masks
// ---------------------------------------------------------------------------------------------------
// pair = (block b, offset k), index b * K + k.  scratch: cnt[pairs], ustart[pairs]
__global__ void __launch_bounds__(256) k_wg_count(const uint32_t* __restrict__ kmask, int num_tiles, int K, int T, int npairs,
                                                  int32_t* __restrict__ cnt) {
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= npairs) return;
  const int b = pr / K, k = pr - b * K;
  int c = 0;
  const int t1 = min(num_tiles, (b + 1) * T);
  for (int t = b * T; t < t1; ++t) c += (kmask[t] >> k) & 1u;
  cnt[pr] = c;
}

// single block: exclusive scans of cnt (unit starts) and of (cnt > 0) (item slots); writes the items and the header
__global__ void __launch_bounds__(1024) k_wg_scan(const int32_t* __restrict__ cnt, int npairs, int K, int T, int32_t* __restrict__ plan,
                                                  int32_t* __restrict__ ustart) {
  __shared__ int wsum_u[32], wsum_i[32];
  __shared__ int run_u, run_i;
  if (threadIdx.x == 0) run_u = run_i = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t* items = plan + kWgPlanHeader;
  for (int base = 0; base < npairs; base += 1024) {
    const int pr = base + threadIdx.x;
    const int c = pr < npairs ? cnt[pr] : 0;
    const int f = c > 0 ? 1 : 0;
    int su = c, si = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int tu = __shfl_up_sync(0xffffffffu, su, d), ti = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d) {
        su += tu;
        si += ti;
      }
    }
    if (lane == 31) {
      wsum_u[w] = su;
      wsum_i[w] = si;
    }
    __syncthreads();
    int off_u = run_u, off_i = run_i;
    for (int j = 0; j < w; ++j) {
      off_u += wsum_u[j];
      off_i += wsum_i[j];
    }
    if (pr < npairs) {
      const int us = off_u + su - c;
      ustart[pr] = us;
      if (f) {
        const int slot = off_i + si - 1;
        items[2 * slot] = us;
        items[2 * slot + 1] = ((pr % K) << kWgItemShift) | c;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tu = 0, ti = 0;
      for (int j = 0; j < 32; ++j) {
        tu += wsum_u[j];
        ti += wsum_i[j];
      }
      run_u += tu;
      run_i += ti;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    plan[0] = run_i;
    plan[1] = run_u;
    plan[2] = T;
    plan[3] = 0;
  }
}

__global__ void __launch_bounds__(256) k_wg_fill(const uint32_t* __restrict__ kmask, int num_tiles, int K, int T, int npairs,
                                                 const int32_t* __restrict__ ustart, int32_t* __restrict__ units) {
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= npairs) return;
  const int b = pr / K, k = pr - b * K;
  int o = ustart[pr];
  const int t1 = min(num_tiles, (b + 1) * T);
  for (int t = b * T; t < t1; ++t)
    if ((kmask[t] >> k) & 1u) units[o++] = t;
}

// deterministic merge: dW[k][mb * 128 + r][c] = sum over the items i of offset k, in item order, of partial[mb * n_items + i][r][c]
__global__ void __launch_bounds__(256) k_wg_merge(const float* __restrict__ partial, const int32_t* __restrict__ plan, int max_items, int K,
                                                  int cin, int cout, int nmb, float* __restrict__ dw) {
  const int n_items = plan[0];
  const int32_t* items = plan + kWgPlanHeader;
  const int k = blockIdx.y;
  const int64_t per_k = (int64_t)cin * cout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per_k; e += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / cout), c = (int)(e - (int64_t)row * cout);
    const int mb = row >> 7, r = row & 127;
    float acc = 0.f;
    for (int i = 0; i < n_items; ++i) {
      const int2 v = __ldg(reinterpret_cast<const int2*>(items) + i);
      if ((v.y >> kWgItemShift) == k && (v.y & ((1 << kWgItemShift) - 1)) > 0)
        acc += partial[(((int64_t)mb * n_items + i) * 128 + r) * cout + c];
    }
    dw[(int64_t)k * per_k + e] = acc;
  }
  (void)max_items;
  (void)K;
  (void)nmb;
}

extern int g_deterministic;  // conv_umma.cu (mk_set_deterministic)

inline int wgrad_block_tiles(int64_t num_tiles) {
  // ~2 tile blocks per SM: long enough accumulation runs (one flush per (block, offset)), enough items to balance the grid
  int64_t t = ceil_div(num_tiles, 2 * (int64_t)num_sms());
  if (t < 8) t = 8;
  if (t > 64) t = 64;
  return (int)t;
}
inline int64_t wgrad_max_items(int64_t num_tiles, int K) { return ceil_div(num_tiles, wgrad_block_tiles(num_tiles)) * K; }

inline size_t wgrad_smem_bytes(int nb_chunks, int stages) {
  return 1024 + (size_t)stages * (size_t)(kWgAChunks + nb_chunks) * kWgChunkBytes + (2 * kMaxStages + 6) * sizeof(uint64_t) + 16 +
         (size_t)4 * 32 * kWgFlushPitch * sizeof(float);
}

}  // namespace mk

using namespace mk;

extern "C" {

int64_t mk_conv_wgrad_plan_elems(int64_t num_tiles, int K) {
  const int64_t mi = wgrad_max_items(num_tiles > 0 ? num_tiles : 1, K);
  // header | items [mi][2] | units [num_tiles * K] | scratch: cnt [mi], ustart [mi]
  return kWgPlanHeader + 2 * mi + num_tiles * (int64_t)K + 2 * mi;
}

int mk_conv_wgrad_plan(const uint32_t* tile_kmask, int64_t num_tiles, int K, int32_t* plan, void* stream_) {
  MK_REQUIRE(K >= 1 && K <= 32, "K must be <= 32 (per-tile offset bitmask)");
  MK_REQUIRE(num_tiles < ((int64_t)1 << 24), "too many tiles");
  cudaStream_t st = (cudaStream_t)stream_;
  if (num_tiles == 0) {
    cudaMemsetAsync(plan, 0, (size_t)kWgPlanHeader * sizeof(int32_t), st);
    return check_launch(__func__);
  }
  const int T = wgrad_block_tiles(num_tiles);
  const int64_t mi = wgrad_max_items(num_tiles, K);
  int32_t* units = plan + kWgPlanHeader + 2 * mi;
  int32_t* cnt = units + num_tiles * (int64_t)K;
  int32_t* ustart = cnt + mi;
  const int npairs = (int)mi;
  k_wg_count<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(tile_kmask, (int)num_tiles, K, T, npairs, cnt);
  k_wg_scan<<<1, 1024, 0, st>>>(cnt, npairs, K, T, plan, ustart);
  k_wg_fill<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(tile_kmask, (int)num_tiles, K, T, npairs, ustart, units);
  return check_launch(__func__);
}

int64_t mk_conv_wgrad_det_workspace_bytes(int64_t n_slots, int K, int cin, int cout) {
  const int64_t items = wgrad_max_items(ceil_div(n_slots > 0 ? n_slots : 1, kTileM), K) * ((cin + 127) / 128);
  return items * 128 * (int64_t)cout * (int64_t)sizeof(float) + 256;
}

int mk_conv_wgrad_bf16_ws(const void* x, int64_t n_in, int cin, const void* dy, int cout, const int32_t* nbr, int K, int64_t n_slots,
                          const int32_t* out_rows, const int32_t* plan, int64_t n_out, float* dW, void* workspace, int64_t workspace_bytes,
                          void* stream_);

int mk_conv_wgrad_bf16(const void* x, int64_t n_in, int cin, const void* dy, int cout, const int32_t* nbr, int K, int64_t n_slots,
                       const int32_t* out_rows, const int32_t* plan, int64_t n_out, float* dW, void* stream_) {
  return mk_conv_wgrad_bf16_ws(x, n_in, cin, dy, cout, nbr, K, n_slots, out_rows, plan, n_out, dW, nullptr, 0, stream_);
}

int mk_conv_wgrad_bf16_ws(const void* x, int64_t n_in, int cin, const void* dy, int cout, const int32_t* nbr, int K, int64_t n_slots,
                          const int32_t* out_rows, const int32_t* plan, int64_t n_out, float* dW, void* workspace, int64_t workspace_bytes,
                          void* stream_) {
  (void)n_in;
  MK_REQUIRE(mk_device_supported(), "tcgen05 path needs an sm_100 device (B200)");
  MK_REQUIRE(cin % 8 == 0 && cin >= 8, "cin must be a multiple of 8");
  MK_REQUIRE(cout % 16 == 0 && cout >= 16 && cout <= 256, "cout must be a multiple of 16 in [16,256]");
  MK_REQUIRE(K >= 1 && K <= 32, "K must be <= 32");
  MK_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0, "16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream_;
  cudaMemsetAsync(dW, 0, (size_t)K * cin * cout * sizeof(float), st);
  if (n_slots == 0) return check_launch(__func__);
  WgradParams p;
  p.x = (const __nv_bfloat16*)x;
  p.dy = (const __nv_bfloat16*)dy;
  p.nbr = nbr;
  p.out_rows = out_rows;
  p.plan = plan;
  p.dw = dW;
  // deterministic mode (mk_set_deterministic(1)) with a workspace: per-item accumulators + fixed-order merge instead of REDs
  const bool det = g_deterministic && workspace != nullptr && workspace_bytes >= mk_conv_wgrad_det_workspace_bytes(n_slots, K, cin, cout);
  p.partial = det ? (float*)workspace : nullptr;
  p.n_slots = n_slots;
  p.n_out = n_out;
  p.cin = cin;
  p.cout = cout;
  p.K = K;
  p.nmb = (cin + 127) / 128;
  p.nb_chunks = (cout + 63) / 64;
  int stages = kMaxStages;
  while (stages > 2 && wgrad_smem_bytes(p.nb_chunks, stages) > 222 * 1024) --stages;
  p.num_stages = stages;
  p.num_producers = stages < kWgMaxProducers ? stages : kWgMaxProducers;
  const size_t smem = wgrad_smem_bytes(p.nb_chunks, stages);
  cudaFuncSetAttribute(k_conv_wgrad_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  p.max_items = (int)wgrad_max_items(ceil_div(n_slots, kTileM), K);
  const int64_t max_units = (int64_t)p.max_items * p.nmb;
  int grid = num_sms();
  if (grid > max_units) grid = (int)max_units;
  k_conv_wgrad_umma<<<grid, kNumThreads, smem, st>>>(p);
  if (det) {
    const int64_t per_k = (int64_t)cin * cout;
    dim3 mg((unsigned)(ceil_div(per_k, 256) < 64 ? ceil_div(per_k, 256) : 64), (unsigned)K);
    k_wg_merge<<<mg, 256, 0, st>>>(p.partial, plan, p.max_items, K, cin, cout, p.nmb, dW);
  }
  return check_launch(__func__);
}

}  // extern "C"
