// tcgen05 gather-GEMM sparse convolution (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
// Replaces ME's per-offset gather -> GEMM -> atomic scatter launches (K launches x 3 per conv, SURVEY.md
// section 3.4) with ONE persistent, warp-specialised, output-stationary kernel:
//
//   tile      = 128 output slots (rows of the accumulator = TMEM lanes)
//   iteration = (kernel offset k with at least one neighbour in the tile, 64-channel chunk of Cin)
//   A operand = the 128 gathered input rows of the chunk, staged by 4 producer warps with 16-byte
//               cp.async (zero-fill for absent neighbours) straight into the 128B-swizzled K-major
//               shared-memory layout that tcgen05.mma reads
//   B operand = W[k] chunk, pre-swizzled on the device once per weight ("weight image"), so a stage is one
//               linear cp.async.bulk (TMA engine, UBLKCP) completing on the stage's mbarrier
//   D         = fp32 accumulator in TMEM, two buffers of `cout` columns: the epilogue of tile t
//               (tcgen05.ld -> bias -> cast -> 16-byte stores, each output row written exactly once,
//               no atomics) overlaps the MMAs of tile t+1
//
// Offsets whose neighbour column is empty for the whole tile are skipped via the per-tile bitmask from
// mk_tile_kmask, which is what keeps the MMA work close to the algorithmic 2*P*Cin*Cout on surface data.
//
// Warp roles (320 threads): warps 0-3 gather producers, warp 4 MMA issuer (one elected lane), warp 5 weight
// loader + TMEM allocator, warps 6-9 epilogue (TMEM lane quarter = warp % 4).
#include "common.cuh"

namespace mk {

constexpr int kTileM = MK_TILE_ROWS;       // 128
constexpr int kChunkElems = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kABytes = kTileM * 128;      // 16 KB per stage
constexpr int kNumProducerWarps = 4;
constexpr int kNumProducerThreads = kNumProducerWarps * 32;
constexpr int kMmaWarp = 4;
constexpr int kLoadWarp = 5;
constexpr int kFirstEpiWarp = 6;
constexpr int kNumThreads = 320;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread = lane = accumulator row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 format, version 1):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) | [32,46) SBO >> 4 (1024 B between
//   8-row groups) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct ConvParams {
  const __nv_bfloat16* in;
  const __nv_bfloat16* wimg;
  const int32_t* nbr;
  const int32_t* out_rows;
  const uint32_t* tile_kmask;
  const int32_t* kidx;
  const float* bias;
  void* out;
  int64_t n_slots;
  int64_t n_out;
  int cin;
  int cout;
  int K;
  int num_tiles;
  int num_stages;
  int nchunks;
  int out_cols;    // columns actually stored (<= cout; the rest of the accumulator is padding)
  int out_stride;  // elements between consecutive output rows
};

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  int n;
  __device__ explicit Ring(int n_) : n(n_) {}
  __device__ void advance() {
    if (++stage == n) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// kLag: cp.async groups a producer thread keeps in flight before it signals the oldest one (gather latency hiding)
template <typename TOut, int kLag>
__global__ void __launch_bounds__(kNumThreads, 1) k_conv_umma(const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A 16 KB][stages x B cout*128][barriers]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int S = p.num_stages;
  const int b_bytes = p.cout * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)S * kABytes;
  uint64_t* bars = (uint64_t*)(sB + (size_t)S * b_bytes);
  uint64_t* full_bar = bars;                       // [S]  producers (128) + weight loader (1, with tx bytes)
  uint64_t* empty_bar = bars + kMaxStages;         // [S]  one tcgen05.commit
  uint64_t* tmem_full = bars + 2 * kMaxStages;     // [2]  one tcgen05.commit
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]  128 epilogue threads
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), kNumProducerThreads + 1);
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

  const int nchunks = p.nchunks;

  if (warp < kNumProducerWarps) {
    // ===================== gather producers =====================
    Ring ring(S);
    int issued = 0, arrived = 0;
    int arr_stage = 0;
    const int piece = lane & 7;     // 16-byte piece of the 128-byte row
    const int rsub = lane >> 3;     // 4 rows per warp instruction
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t kmask = p.tile_kmask[tile];
      const int64_t slot = (int64_t)tile * kTileM + warp * 32 + lane;
      for (uint32_t m = kmask; m; m &= m - 1) {
        const int k = __ffs(m) - 1;
        const int my_idx = slot < p.n_slots ? p.nbr[(int64_t)k * p.n_slots + slot] : -1;
        int idxs[8];  // input rows of the 8 tile rows this lane copies a piece of (rows 4j + rsub of the warp)
#pragma unroll
        for (int j = 0; j < 8; ++j) idxs[j] = __shfl_sync(0xffffffffu, my_idx, 4 * j + rsub);
        for (int ch = 0; ch < nchunks; ++ch) {
          const int width = min(kChunkElems, p.cin - ch * kChunkElems);  // valid elements in this chunk
          mbar_wait(smem_u32(&empty_bar[ring.stage]), ring.phase ^ 1);
          const uint32_t a_base = smem_u32(sA + (size_t)ring.stage * kABytes);
          if (piece * 8 < width) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int idx = idxs[j];
              const int row = warp * 32 + 4 * j + rsub;  // row within the tile
              const uint32_t dst = a_base + (uint32_t)(row * 128 + ((piece ^ (row & 7)) << 4));
              const __nv_bfloat16* src = p.in + (int64_t)(idx < 0 ? 0 : idx) * p.cin + ch * kChunkElems + piece * 8;
              cp_async_16(dst, src, idx < 0 ? 0u : 16u);
            }
          }
          cp_async_commit();
          ++issued;
          ring.advance();
          if (issued - arrived > kLag) {
            cp_async_wait<kLag>();
            fence_proxy_async();
            mbar_arrive(smem_u32(&full_bar[arr_stage]));
            ++arrived;
            if (++arr_stage == S) arr_stage = 0;
          }
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (arrived < issued) {
      mbar_arrive(smem_u32(&full_bar[arr_stage]));
      ++arrived;
      if (++arr_stage == S) arr_stage = 0;
    }
  } else if (warp == kLoadWarp) {
    // ===================== weight loader (one lane) =====================
    if (lane == 0) {
      Ring ring(S);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const uint32_t kmask = p.tile_kmask[tile];
        for (uint32_t m = kmask; m; m &= m - 1) {
          const int k = __ffs(m) - 1;
          const int kw = p.kidx ? p.kidx[k] : k;
          for (int ch = 0; ch < nchunks; ++ch) {
            mbar_wait(smem_u32(&empty_bar[ring.stage]), ring.phase ^ 1);
            const uint32_t bar = smem_u32(&full_bar[ring.stage]);
            mbar_arrive_expect_tx(bar, (uint32_t)b_bytes);
            const __nv_bfloat16* src = p.wimg + ((int64_t)kw * nchunks + ch) * (int64_t)p.cout * kChunkElems;
            bulk_copy_g2s(smem_u32(sB + (size_t)ring.stage * b_bytes), src, (uint32_t)b_bytes, bar);
            ring.advance();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      Ring ring(S);
      const uint32_t idesc = make_idesc(kTileM, p.cout);
      int acc_buf = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const uint32_t kmask = p.tile_kmask[tile];
        if (kmask == 0) continue;
        mbar_wait(smem_u32(&tmem_empty[acc_buf]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc_buf * 256);
        uint32_t accumulate = 0;
        for (uint32_t m = kmask; m; m &= m - 1) {
          for (int ch = 0; ch < nchunks; ++ch) {
            const int width = min(kChunkElems, p.cin - ch * kChunkElems);
            mbar_wait(smem_u32(&full_bar[ring.stage]), ring.phase);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(smem_u32(sA + (size_t)ring.stage * kABytes));
            const uint64_t bdesc = make_smem_desc(smem_u32(sB + (size_t)ring.stage * b_bytes));
            for (int kk = 0; kk < width / 16; ++kk) {
              // +32 bytes along K inside the 128-byte swizzle row == +2 in the (addr >> 4) field
              umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(smem_u32(&empty_bar[ring.stage]));  // frees the stage when these MMAs retire
            ring.advance();
          }
        }
        umma_commit(smem_u32(&tmem_full[acc_buf]));
        if (++acc_buf == 2) {
          acc_buf = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (4 warps) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t kmask = p.tile_kmask[tile];
      const int64_t slot = (int64_t)tile * kTileM + q * 32 + lane;
      int64_t orow = -1;
      if (slot < p.n_slots) orow = p.out_rows ? (int64_t)p.out_rows[slot] : slot;
      if (orow >= p.n_out) orow = -1;
      TOut* optr = orow >= 0 ? (TOut*)p.out + orow * p.out_stride : nullptr;
      const bool vec_ok = ((p.out_stride * (int)sizeof(TOut)) & 15) == 0;
      if (kmask != 0) {
        mbar_wait(smem_u32(&tmem_full[acc_buf]), acc_phase);
        tc_fence_after();
      }
      for (int c0 = 0; c0 < p.out_cols; c0 += 16) {
        float v[16];
        if (kmask != 0) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_buf * 256 + c0), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += (c0 + j < p.out_cols) ? p.bias[c0 + j] : 0.f;
        }
        if (optr) {
          if (vec_ok && c0 + 16 <= p.out_cols) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) store4(optr + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.out_cols) optr[c0 + j] = from_f32<TOut>(v[j]);
          }
        }
      }
      if (kmask != 0) {
        tc_fence_before();
        mbar_arrive(smem_u32(&tmem_empty[acc_buf]));
        if (++acc_buf == 2) {
          acc_buf = 0;
          acc_phase ^= 1;
        }
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
// weight images
// ---------------------------------------------------------------------------------------------------
// img[((k * nchunks + ch) * rows + n) * 64 + swz(n, e)] = W'[k][ch*64 + e][n],  W' = W (rows = cout) or W^T.
__global__ void __launch_bounds__(256) k_weight_image(const float* __restrict__ W, int K, int cin, int cout, int transpose,
                                                      __nv_bfloat16* __restrict__ img) {
  const int red = transpose ? cout : cin;     // reduction length
  const int rows = transpose ? cin : cout;    // N of the GEMM
  const int nchunks = (red + kChunkElems - 1) / kChunkElems;
  const int64_t total = (int64_t)K * nchunks * rows * kChunkElems;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % kChunkElems);
    int64_t t = i / kChunkElems;
    const int n = (int)(t % rows);
    t /= rows;
    const int ch = (int)(t % nchunks);
    const int k = (int)(t / nchunks);
    const int r = ch * kChunkElems + e;
    float v = 0.f;
    if (r < red) v = transpose ? W[((int64_t)k * cin + n) * cout + r] : W[((int64_t)k * cin + r) * cout + n];
    const int piece = e >> 3, within = e & 7;
    const int64_t dst = ((int64_t)(k * nchunks + ch) * rows + n) * kChunkElems + (((piece ^ (n & 7)) << 3) | within);
    img[dst] = __float2bfloat16_rn(v);
  }
}

inline size_t conv_smem_bytes(int cout, int stages) {
  return 1024 + (size_t)stages * (kABytes + (size_t)cout * 128) + (2 * kMaxStages + 4) * sizeof(uint64_t) + 16;
}

}  // namespace mk

using namespace mk;

extern "C" {

int64_t mk_conv_weight_image_elems(int K, int cin, int cout, int transpose) {
  const int red = transpose ? cout : cin, rows = transpose ? cin : cout;
  return (int64_t)K * ((red + kChunkElems - 1) / kChunkElems) * rows * kChunkElems;
}

int mk_conv_weight_image_bf16(const float* W, int K, int cin, int cout, int transpose, void* img, void* stream_) {
  const int64_t total = mk_conv_weight_image_elems(K, cin, cout, transpose);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  k_weight_image<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(W, K, cin, cout, transpose, (__nv_bfloat16*)img);
  return check_launch(__func__);
}

int mk_conv_fwd_bf16(const void* in, int cin, const void* wimg, int cout, const int32_t* nbr, int K, int64_t n_slots,
                     const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                     const float* bias, void* out, int out_dtype, int out_cols, int out_stride, void* stream_) {
  MK_REQUIRE(mk_device_supported(), "tcgen05 path needs an sm_100 device (B200)");
  MK_REQUIRE(cin % 16 == 0 && cin >= 16, "cin must be a multiple of 16");
  MK_REQUIRE(cout % 16 == 0 && cout >= 16 && cout <= 256, "cout must be a multiple of 16 in [16,256]");
  MK_REQUIRE(K >= 1 && K <= 32, "K must be <= 32 (per-tile offset bitmask)");
  if (out_cols <= 0) out_cols = cout;
  if (out_stride <= 0) out_stride = out_cols;
  MK_REQUIRE(out_cols <= cout && out_stride >= out_cols, "out_cols <= cout and out_stride >= out_cols required");
  MK_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)wimg & 15) == 0 && ((uintptr_t)out & 15) == 0, "16-byte alignment");
  if (n_slots == 0) return 0;
  ConvParams p;
  p.in = (const __nv_bfloat16*)in;
  p.wimg = (const __nv_bfloat16*)wimg;
  p.nbr = nbr;
  p.out_rows = out_rows;
  p.tile_kmask = tile_kmask;
  p.kidx = kidx;
  p.bias = bias;
  p.out = out;
  p.n_slots = n_slots;
  p.n_out = n_out;
  p.cin = cin;
  p.cout = cout;
  p.K = K;
  p.num_tiles = (int)ceil_div(n_slots, kTileM);
  p.nchunks = (cin + kChunkElems - 1) / kChunkElems;
  p.out_cols = out_cols;
  p.out_stride = out_stride;
  int stages = kMaxStages;
  while (stages > 3 && conv_smem_bytes(cout, stages) > 210 * 1024) --stages;
  MK_REQUIRE(conv_smem_bytes(cout, stages) <= 227 * 1024, "shared memory budget exceeded");
  p.num_stages = stages;
  const size_t smem = conv_smem_bytes(cout, stages);
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  cudaStream_t st = (cudaStream_t)stream_;
  // producers keep (lag + 1) stages of gathers in flight; the ring must be strictly deeper than that
  const int lag = stages >= 8 ? 5 : (stages >= 6 ? 4 : (stages >= 5 ? 3 : (stages >= 4 ? 2 : 1)));
#define MK_LAUNCH(T, L)                                                                                     \
  do {                                                                                                      \
    cudaFuncSetAttribute(k_conv_umma<T, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    k_conv_umma<T, L><<<grid, kNumThreads, smem, st>>>(p);                                                  \
  } while (0)
#define MK_LAUNCH_T(T)              \
  do {                              \
    switch (lag) {                  \
      case 5: MK_LAUNCH(T, 5); break; \
      case 4: MK_LAUNCH(T, 4); break; \
      case 3: MK_LAUNCH(T, 3); break; \
      case 2: MK_LAUNCH(T, 2); break; \
      default: MK_LAUNCH(T, 1); break; \
    }                               \
  } while (0)
  if (out_dtype == MK_F32) MK_LAUNCH_T(float);
  else MK_LAUNCH_T(__nv_bfloat16);
#undef MK_LAUNCH_T
#undef MK_LAUNCH
  return check_launch(__func__);
}

}  // extern "C"
