// tcgen05 weight-gradient of the sparse convolution (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
//   dW[k][ci][co] = sum over output slots o with nbr[k,o] >= 0 of  x[nbr[k,o]][ci] * dy[row(o)][co]
//
// i.e. per kernel offset k one GEMM  D[M = Cin, N = Cout] = A[M, pairs] * B[pairs, N]  whose reduction runs over the PAIRS of
// the kernel map.  ME computes it as K x (gather, gather, GEMM) launches (ConvolutionBackwardGPU); here it is ONE persistent,
// warp-specialised kernel over the same mask-sorted tile plan the forward kernel uses:
//
//   work unit  = (128-channel block of Cin, offset k, 128-slot tile that has offset k) — the units, ordered by (block, k, tile),
//                are cut into gridDim.x equal contiguous ranges, so a CTA keeps ONE accumulator alive across a long run of
//                tiles of the same k and flushes it once per (block, k) it touches
//   stage      = 64 pair rows: the gathered x rows (A) and the matching dy rows (B), 16-byte cp.async (LDGSTS, zero-fill for
//                absent pairs) into 128-byte rows with the 128B swizzle.  Both operands have the reduction index as their ROW,
//                i.e. they are MN-major for the tensor core; the MN-major SWIZZLE_128B canonical layout
//                ((8,n),(8,k)) : ((1,LBO),(8,SBO)) [16-byte units] is exactly "rows of 128 bytes, 16-byte piece XOR (row & 7)",
//                the same staging the forward gather produces, with SBO = 1024 B (8 rows) and LBO = the distance between
//                64-channel chunks
//   MMA        = tcgen05.mma kind::f16, M = 128 (Cin block; rows beyond Cin are zero), N = Cout, K = 16 pair rows, both
//                operands MN-major (instruction-descriptor bits 15/16), accumulator fp32 in TMEM, double buffered
//   flush      = tcgen05.ld -> fp32 atomicAdd into dW (a handful of flushes per CTA; dW is zeroed by the entry point)
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
  const int32_t* plan;  // [0] = number of items I, [1 .. K] per-offset counts, [1 + K ...] items ordered by (k, tile)
  float* dw;
  int64_t n_slots;
  int64_t n_out;
  int cin, cout, K;
  int nmb;            // 128-channel blocks of Cin
  int nb_chunks;      // ceil(cout / 64)
  int num_stages;
  int num_producers;
};

// instruction descriptor: as make_idesc, plus A and B MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) { return make_idesc(M, N) | (1u << 15) | (1u << 16); }

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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the staging area once: pieces beyond a narrow Cin block are never written by the gathers but are read by the
  // M = 128 MMA (they only reach accumulator rows that are not flushed, but must not be NaN-producing garbage for tidy runs)
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

  // this CTA's contiguous range of units
  const int64_t I = p.plan[0];
  const int64_t U = I * p.nmb;
  const int64_t lo = U * (int64_t)blockIdx.x / (int64_t)gridDim.x;
  const int64_t hi = U * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;
  const int32_t* items = p.plan + 1 + p.K;
  const int64_t n_iter = 2 * (hi - lo);  // two 64-row halves per unit

  if (warp < kNumProducerWarps || warp == kLoadWarp) {
    // ===================== producers: stage i belongs to producer i mod NP =====================
    const int pi = warp < kNumProducerWarps ? warp : kNumProducerWarps;
    const int NP = p.num_producers;
    const int piece = lane & 7, rsub = lane >> 3;
    const uint32_t sw_even = (uint32_t)((piece ^ rsub) << 4), sw_odd = (uint32_t)((piece ^ (4 + rsub)) << 4);
    const uint32_t x_row_bytes = (uint32_t)p.cin * 2u, dy_row_bytes = (uint32_t)p.cout * 2u;
    struct Idx {
      int in[2], out[2];  // pair rows lane, lane + 32 of the stage: x row / dy row (-1 = absent -> zero fill)
      int mb;
    };
    auto load_idx = [&](int64_t it) -> Idx {
      Idx r;
      r.mb = 0;
      r.in[0] = r.in[1] = r.out[0] = r.out[1] = -1;
      if (it < n_iter) {
        const int64_t u = lo + (it >> 1);
        r.mb = (int)(u / I);
        const int32_t item = __ldg(items + (u - (int64_t)r.mb * I));
        const int k = item >> kWgItemShift;
        const int64_t tile = item & ((1 << kWgItemShift) - 1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int64_t slot = tile * kTileM + (it & 1) * kWgRows + q * 32 + lane;
          if (slot < p.n_slots) {
            const int in_row = __ldg(p.nbr + (int64_t)k * p.n_slots + slot);
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
    int64_t it = pi;
    Idx ci = load_idx(it), ni = load_idx(it + NP);
    int stage = pi % S;
    uint32_t phase = (uint32_t)((pi / S) & 1);
    while (active && it < n_iter) {
      const Idx nni = load_idx(it + 2 * NP);
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      const uint32_t bar = smem_u32(&full_bar[stage]);
      const uint32_t a_base = smem_u32(smem + (size_t)stage * stage_bytes);
      const uint32_t b_base = a_base + (uint32_t)a_bytes;
      const int cinb = min(128, p.cin - ci.mb * 128);  // valid channels of this Cin block
      uint64_t xsrc = (uint64_t)p.x + (uint64_t)(ci.mb * 128 + piece * 8) * 2u;
      uint64_t dsrc = (uint64_t)p.dy + (uint64_t)(piece * 8) * 2u;
      asm volatile("" : "+l"(xsrc));
      asm volatile("" : "+l"(dsrc));
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int ii = __shfl_sync(0xffffffffu, ci.in[q], 4 * jj + rsub);
          const int oo = __shfl_sync(0xffffffffu, ci.out[q], 4 * jj + rsub);
          const int row = q * 32 + 4 * jj + rsub;
          const uint32_t roff = (uint32_t)(row * 128) + ((jj & 1) ? sw_odd : sw_even);
          const uint32_t nbytes = ii < 0 ? 0u : 16u;
          const uint64_t xs = xsrc + (uint64_t)(ii < 0 ? 0u : (uint32_t)ii) * x_row_bytes;
          const uint64_t ds = dsrc + (uint64_t)(oo < 0 ? 0u : (uint32_t)oo) * dy_row_bytes;
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
      it += NP;
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
    int64_t u = lo;
    while (u < hi) {
      // run of units with the same (Cin block, offset)
      const int mb = (int)(u / I);
      const int k = items[u - (int64_t)mb * I] >> kWgItemShift;
      int64_t u_end = u + 1;
      const int64_t blk_end = min(hi, ((int64_t)mb + 1) * I);
      while (u_end < blk_end && (items[u_end - (int64_t)mb * I] >> kWgItemShift) == k) ++u_end;
      mbar_wait(smem_u32(&tmem_empty[acc_buf]), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc_buf * 256);
      uint32_t accumulate = 0;
      for (int64_t it = 2 * (u - lo); it < 2 * (u_end - lo); ++it) {
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
      u = u_end;
    }
  } else {
    // ===================== flush (4 warps): TMEM -> atomicAdd into dW =====================
    const int q = warp & 3;
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    int64_t u = lo;
    while (u < hi) {
      const int mb = (int)(u / I);
      const int k = items[u - (int64_t)mb * I] >> kWgItemShift;
      int64_t u_end = u + 1;
      const int64_t blk_end = min(hi, ((int64_t)mb + 1) * I);
      while (u_end < blk_end && (items[u_end - (int64_t)mb * I] >> kWgItemShift) == k) ++u_end;
      mbar_wait(smem_u32(&tmem_full[acc_buf]), acc_phase);
      tc_fence_after();
      const int ci_row = mb * 128 + q * 32 + lane;
      float* dst = p.dw + ((int64_t)k * p.cin + ci_row) * p.cout;
      for (int c0 = 0; c0 < p.cout; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_buf * 256 + c0), r);
        tmem_ld_wait();
        if (ci_row < p.cin) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + c0 + j, __uint_as_float(r[j]));
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[acc_buf]));
      if (++acc_buf == 2) {
        acc_buf = 0;
        acc_phase ^= 1;
      }
      u = u_end;
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
// unit list of a tile plan: items ordered by (k, tile) for every (tile, k) with bit k set in tile_kmask[tile]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wgrad_count(const uint32_t* __restrict__ kmask, int num_tiles, int32_t* __restrict__ plan) {
  const int k = blockIdx.x;
  int c = 0;
  for (int t = threadIdx.x; t < num_tiles; t += blockDim.x) c += (kmask[t] >> k) & 1u;
  __shared__ int red[8];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    plan[1 + k] = s;
  }
}

__global__ void __launch_bounds__(256) k_wgrad_fill(const uint32_t* __restrict__ kmask, int num_tiles, int K, int32_t* __restrict__ plan) {
  const int k = blockIdx.x;
  int base = 0;
  for (int j = 0; j < k; ++j) base += plan[1 + j];
  if (k == K - 1 && threadIdx.x == 0) plan[0] = base + plan[1 + k];
  int32_t* items = plan + 1 + K;
  __shared__ int wsum[8];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int start = 0; start < num_tiles; start += blockDim.x) {
    const int t = start + threadIdx.x;
    const bool has = t < num_tiles && ((kmask[t] >> k) & 1u);
    const unsigned bal = __ballot_sync(0xffffffffu, has);
    if (lane == 0) wsum[w] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int j = 0; j < w; ++j) off += wsum[j];
    if (has) items[base + off + __popc(bal & ((1u << lane) - 1u))] = (k << kWgItemShift) | t;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int j = 0; j < 8; ++j) s += wsum[j];
      running += s;
    }
    __syncthreads();
  }
}

inline size_t wgrad_smem_bytes(int nb_chunks, int stages) {
  return 1024 + (size_t)stages * (size_t)(kWgAChunks + nb_chunks) * kWgChunkBytes + (2 * kMaxStages + 4) * sizeof(uint64_t) + 16;
}

}  // namespace mk

using namespace mk;

extern "C" {

int64_t mk_conv_wgrad_plan_elems(int64_t num_tiles, int K) { return 1 + (int64_t)K + num_tiles * (int64_t)K; }

int mk_conv_wgrad_plan(const uint32_t* tile_kmask, int64_t num_tiles, int K, int32_t* plan, void* stream_) {
  MK_REQUIRE(K >= 1 && K <= 32, "K must be <= 32 (per-tile offset bitmask)");
  MK_REQUIRE(num_tiles < ((int64_t)1 << kWgItemShift), "too many tiles for the packed item format");
  cudaStream_t st = (cudaStream_t)stream_;
  if (num_tiles == 0) {
    cudaMemsetAsync(plan, 0, (size_t)(1 + K) * sizeof(int32_t), st);
    return check_launch(__func__);
  }
  k_wgrad_count<<<K, 256, 0, st>>>(tile_kmask, (int)num_tiles, plan);
  k_wgrad_fill<<<K, 256, 0, st>>>(tile_kmask, (int)num_tiles, K, plan);
  return check_launch(__func__);
}

int mk_conv_wgrad_bf16(const void* x, int64_t n_in, int cin, const void* dy, int cout, const int32_t* nbr, int K, int64_t n_slots,
                       const int32_t* out_rows, const int32_t* plan, int64_t n_out, float* dW, void* stream_) {
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
  const int64_t max_units = ceil_div(n_slots, kTileM) * K * p.nmb;
  int grid = num_sms();
  if (grid > max_units) grid = (int)max_units;
  k_conv_wgrad_umma<<<grid, kNumThreads, smem, st>>>(p);
  return check_launch(__func__);
}

}  // extern "C"
