// tcgen05 gather-GEMM sparse convolution, TMA-gather variant (the default tensor-core path).
//
// Same tiling / pipeline contract as conv_umma.cu (128-slot output tiles, (offset k, 64-channel chunk) iterations, fp32
// accumulators double-buffered in TMEM, per-tile offset masks, mask-sorted tile plans), but the A operand is staged by the
// TMA engine instead of the LSU:  every lane of ONE producer warp issues a
//     cp.async.bulk.tensor.2d.tile::gather4   (SASS: UTMALDG)
// that fetches 4 arbitrary rows x 64 channels of the input feature matrix through a {64, 1}-box tensor map with the
// 128-byte swizzle tcgen05.mma expects; 32 lanes x 4 rows = the 128 rows of a stage.  Absent neighbours are row index -1:
// out-of-range coordinates are zero-filled by TMA, which is exactly the "no contribution" the convolution needs.
// Measured motivation (profiles/): with 16-byte cp.async the LSU retires ~1 gathered 128-byte row per 8 cycles per SM, which
// capped every channel width <= 128 at ~0.6 us per stage regardless of bytes moved.
//
// Warp roles (256 threads): warp 0 TMA producer (A gathers on all lanes, weight-image bulk copy on lane 0), warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-7 epilogue (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <cstring>

#include "umma.cuh"

namespace mk {

constexpr int kTmaThreads = 256;
constexpr int kTmaProducerWarp = 0;
constexpr int kTmaMmaWarp = 1;
constexpr int kTmaAllocWarp = 2;
constexpr int kTmaFirstEpiWarp = 4;

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tmap, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"((uint64_t)tmap), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

template <typename TOut>
__global__ void __launch_bounds__(kTmaThreads, 1) k_conv_tma(const ConvParams p, const __grid_constant__ CUtensorMap tmap_in) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int S = p.num_stages;
  const int b_bytes = p.cout * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)S * kABytes;
  uint64_t* bars = (uint64_t*)(sB + (size_t)S * b_bytes);
  uint64_t* full_bar = bars;                          // [S]  1 arrive.expect_tx (A gathers + weight image bytes)
  uint64_t* empty_bar = bars + kMaxStages;            // [S]  one tcgen05.commit
  uint64_t* tmem_full = bars + 2 * kMaxStages;        // [2]  one tcgen05.commit
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;   // [2]  128 epilogue threads
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 4 * 32);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_in);
  }
  if (warp == kTmaAllocWarp) {
    tmem_alloc(smem_u32(tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int nchunks = p.nchunks;

  if (warp == kTmaProducerWarp) {
    // ===================== TMA producer =====================
    Ring ring(S);
    // lane l gathers tile rows 4l .. 4l+3; their input-row indices are prefetched 3 work items ahead (the index load ->
    // gather address dependency would otherwise expose an L2 round trip per (tile, offset))
    auto load_idx = [&](const WorkCursor& c) -> int4 {
      int4 r = make_int4(-1, -1, -1, -1);
      if (c.valid()) {
        const int64_t slot = (int64_t)c.tile * kTileM + 4 * lane;
        const int32_t* src = p.nbr + (int64_t)c.k() * p.n_slots + slot;
        if (slot + 3 < p.n_slots) {
          r.x = __ldg(src); r.y = __ldg(src + 1); r.z = __ldg(src + 2); r.w = __ldg(src + 3);
        } else {
          if (slot < p.n_slots) r.x = __ldg(src);
          if (slot + 1 < p.n_slots) r.y = __ldg(src + 1);
          if (slot + 2 < p.n_slots) r.z = __ldg(src + 2);
        }
      }
      return r;
    };
    WorkCursor cur(p.tile_kmask, blockIdx.x, gridDim.x, p.num_tiles), pf = cur;
    int4 q0 = load_idx(pf); pf.next();
    int4 q1 = load_idx(pf); pf.next();
    int4 q2 = load_idx(pf); pf.next();
    const uint32_t stage_tx = (uint32_t)(kABytes + b_bytes);
    while (cur.valid()) {
      const int4 idx = q0;
      q0 = q1; q1 = q2;
      q2 = load_idx(pf);
      pf.next();
      const int k = cur.k();
      const int kw = p.kidx ? p.kidx[k] : k;
      for (int ch = 0; ch < nchunks; ++ch) {
        const uint32_t bar = smem_u32(&full_bar[ring.stage]);
        if (lane == 0) {
          mbar_wait(smem_u32(&empty_bar[ring.stage]), ring.phase ^ 1);
          mbar_arrive_expect_tx(bar, stage_tx);
          const __nv_bfloat16* wsrc = p.wimg + ((int64_t)kw * nchunks + ch) * (int64_t)p.cout * kChunkElems;
          bulk_copy_g2s(smem_u32(sB + (size_t)ring.stage * b_bytes), wsrc, (uint32_t)b_bytes, bar);
        }
        __syncwarp();
        tma_gather4(smem_u32(sA + (size_t)ring.stage * kABytes) + (uint32_t)(lane * 512), &tmap_in, ch * kChunkElems, idx.x, idx.y, idx.z,
                    idx.w, bar);
        ring.advance();
      }
      cur.next();
    }
  } else if (warp == kTmaMmaWarp) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      Ring ring(S);
      const uint32_t idesc = make_idesc(kTileM, p.cout);
      int acc_buf = 0;
      uint32_t acc_phase = 0;
      bool tile_open = false;
      uint32_t accumulate = 0;
      for (WorkCursor cur(p.tile_kmask, blockIdx.x, gridDim.x, p.num_tiles); cur.valid(); cur.next()) {
        if (!tile_open) {
          mbar_wait(smem_u32(&tmem_empty[acc_buf]), acc_phase ^ 1);
          tc_fence_after();
          tile_open = true;
          accumulate = 0;
        }
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc_buf * 256);
        for (int ch = 0; ch < nchunks; ++ch) {
          const int width = min(kChunkElems, p.cin - ch * kChunkElems);
          mbar_wait(smem_u32(&full_bar[ring.stage]), ring.phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(sA + (size_t)ring.stage * kABytes));
          const uint64_t bdesc = make_smem_desc(smem_u32(sB + (size_t)ring.stage * b_bytes));
          for (int kk = 0; kk < width / 16; ++kk) {
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(smem_u32(&empty_bar[ring.stage]));
          ring.advance();
        }
        if (cur.last_k_of_tile()) {
          umma_commit(smem_u32(&tmem_full[acc_buf]));
          tile_open = false;
          if (++acc_buf == 2) {
            acc_buf = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= kTmaFirstEpiWarp) {
    // ===================== epilogue (4 warps) =====================
    const int q = warp & 3;
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    auto load_row = [&](int tile) -> int64_t {
      const int64_t slot = (int64_t)tile * kTileM + q * 32 + lane;
      int64_t r = -1;
      if (tile < p.num_tiles && slot < p.n_slots) r = p.out_rows ? (int64_t)__ldg(p.out_rows + slot) : slot;
      return r >= p.n_out ? -1 : r;
    };
    uint32_t kmask_next = blockIdx.x < p.num_tiles ? p.tile_kmask[blockIdx.x] : 0u;
    int64_t orow_next = load_row(blockIdx.x);
    const bool vec_ok = ((p.out_stride * (int)sizeof(TOut)) & 15) == 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t kmask = kmask_next;
      const int64_t orow = orow_next;
      if (tile + (int)gridDim.x < p.num_tiles) kmask_next = p.tile_kmask[tile + gridDim.x];
      orow_next = load_row(tile + gridDim.x);
      TOut* optr = orow >= 0 ? (TOut*)p.out + orow * p.out_stride : nullptr;
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
  if (warp == kTmaAllocWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------
// host: tensor map over the input feature matrix [n_in, cin] bf16 (driver entry point resolved at run time, no -lcuda)
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

// tensor map {cin, n_in} of the rows at `base` (elem = 2: bf16, 4: fp32), box {128 bytes, 1}, SWIZZLE_128B -> 128 bytes at `out`
int encode_rows_tmap(void* out, const void* base, int elem, int cin, int64_t n_in) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -1;
  const cuuint64_t gdim[2] = {(cuuint64_t)cin, (cuuint64_t)(n_in > 0 ? n_in : 1)};
  const cuuint64_t gstride[1] = {(cuuint64_t)cin * (cuuint64_t)elem};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elem), 1u};
  const cuuint32_t estride[2] = {1u, 1u};
  CUtensorMap tmap;
  CUresult r = enc(&tmap, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return -2;
  memcpy(out, &tmap, sizeof(tmap));
  return 0;
}

int conv_tma_launch(const ConvParams& p_in, int64_t n_in, int out_dtype, cudaStream_t st) {
  ConvParams p = p_in;
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail("mk_conv_fwd_bf16", "cuTensorMapEncodeTiled entry point not available");
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)p.cin, (cuuint64_t)(n_in > 0 ? n_in : 1)};
  const cuuint64_t gstride[1] = {(cuuint64_t)p.cin * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {(cuuint32_t)kChunkElems, 1u};
  const cuuint32_t estride[2] = {1u, 1u};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)p.in, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("mk_conv_fwd_bf16", "cuTensorMapEncodeTiled failed");
  int stages = kMaxStages;
  auto smem_bytes = [&](int s) { return 1024 + (size_t)s * (kABytes + (size_t)p.cout * 128) + (2 * kMaxStages + 4) * sizeof(uint64_t) + 16; };
  while (stages > 2 && smem_bytes(stages) > 216 * 1024) --stages;
  if (smem_bytes(stages) > 227 * 1024) return fail("mk_conv_fwd_bf16", "shared memory budget exceeded");
  p.num_stages = stages;
  const size_t smem = smem_bytes(stages);
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  if (out_dtype == MK_F32) {
    cudaFuncSetAttribute(k_conv_tma<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_conv_tma<float><<<grid, kTmaThreads, smem, st>>>(p, tmap);
  } else {
    cudaFuncSetAttribute(k_conv_tma<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_conv_tma<__nv_bfloat16><<<grid, kTmaThreads, smem, st>>>(p, tmap);
  }
  return check_launch("mk_conv_fwd_bf16");
}

}  // namespace mk
