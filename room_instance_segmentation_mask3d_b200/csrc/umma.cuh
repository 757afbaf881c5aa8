// Shared device helpers for the tcgen05 kernels: PTX wrappers (mbarrier, cp.async, bulk copies, TMEM, tcgen05.mma),
// shared-memory / instruction descriptors, the pipeline ring and the (tile, offset) work cursor.
#pragma once

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
// Same, with a suspend-time hint (ns): the hardware may park the thread until the phase completes or the limit passes.  Without
// a hint try_wait returns after a few tens of cycles, and the 6 waiting warps of k_conv_umma (4 epilogue, 2 MMA) then spin
// through ~10 M try_wait + branch instructions per launch (ncu source view), stealing issue slots and MIO bandwidth from
// the producers.
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(ns)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Consumer-side (acquire) half of the generic -> async proxy hand-over: executed by the MMA thread after it has observed the
// stage's mbarrier phase, before tcgen05.mma reads the tile through the async proxy.  The bidirectional fence.proxy.async
// compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC; the MEMBAR queues behind every in-flight LDGSTS of the SM's LSU FIFO
// (~500-1000 cycles per pipeline stage, measured with clock64 around it), which serialised the whole gather pipeline.
// The unidirectional acquire form needs no instruction on sm_100a: the release half is the completion-triggered
// cp.async.mbarrier.arrive (the arrive fires after the copies have been written to shared memory).
__device__ __forceinline__ void fence_proxy_async_acquire() {
  asm volatile("fence.proxy.async::generic.acquire.sync_restrict::shared::cluster.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// true on exactly one lane of a converged warp.  A branch on this predicate is how single-thread tcgen05 / bulk-copy
// instructions must be issued: the compiler then keeps their operands on the uniform datapath and emits ONE UTCHMMA / UTCBAR /
// UBLKCP per warp.  Guarding with `lane == 0` instead makes it treat the code as divergent and wrap every such instruction in
// an ELECT / BRA.U.ANY waterfall loop with R2UR moves (~15 dependent instructions, ~100 cycles per MMA on the issuing thread).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// explicit shared-state-space accesses by 32-bit address (a pointer carved out of the dynamic shared-memory block is generic to
// the compiler, which then emits generic ST.E / LD.E with 64-bit address arithmetic per element)
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ int lds_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int4 lds_v4i(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_s32(uint32_t addr, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  // .ca: allocate in L1 — a tile is a compact spatial patch (mask + Z-curve sorted), so the neighbour rows gathered for its
  // different offsets overlap heavily and are re-read from L1 instead of L2
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` when all cp.async issued so far by this thread have landed (no thread-side wait: the stage stays in flight)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
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

// TMA row gather (sm_100): FOUR rows r0..r3 of a 2-D tensor map (box = {row chunk, 1}), columns [c0, c0 + box), land as four consecutive
// box rows at dst in the map's shared-memory swizzle; rows / columns outside the tensor (r < 0: absent neighbours) are zero-filled
// and still count towards the barrier's transaction bytes.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* tmap, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
               : "memory");
}
__device__ __forceinline__ int4 ldg_nc_v4(const int32_t* p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
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
// Same, with the descriptors given as (lo, hi) words: only the low word (start address field) changes between stages and
// K steps, so the issuing warp does 32-bit uniform arithmetic instead of 64-bit adds.
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tf32 x tf32 -> fp32 (fp32 operands in shared memory, K = 8 per instruction): the reference's own precision
// (torch.set_float32_matmul_precision("high"), mask3d/main_instance_segmentation.py:52)
__device__ __forceinline__ void umma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTf32) umma_tf32_lohi(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate);
  else umma_bf16_lohi(tmem_d, a_lo, b_lo, desc_hi, idesc, accumulate);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t kSmemDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B

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

// instruction descriptor, kind::tf32: D fp32, A/B tf32 (format 2), both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
  int elem;        // bytes per operand element: 2 = bf16 (kind::f16), 4 = fp32 read as tf32 (kind::tf32)
  int num_tiles;
  int num_stages;
  int group;       // member tiles per super-tile (1 or 2)
  int num_producers;  // producer groups (<= num_stages: a group may be at most one ring phase ahead of the consumer)
  int split;          // producer warps per group (2: one warp per member tile of a super-tile)
  int nchunks;
  int out_cols;    // columns actually stored (<= cout; the rest of the accumulator is padding)
  int out_stride;  // elements between consecutive output rows
  int stage_out;   // 1: epilogue transposes each warp's 32 x out_cols block through shared memory and stores it as one
                   // contiguous run (rows are contiguous: out_rows == nullptr, out_stride == out_cols, row pitch not 16-byte aligned)
  int chunk_major; // iteration order inside a super-tile (see SuperCursor)
  int ksplit;      // > 1: gridDim.x = super-tiles x ksplit; CTA (st, part) takes every ksplit-th active offset of its super-tile
                   // and ADDS its partial result into the (zeroed, fp32) output with vector REDs
  int debug;       // timing ablation flags (MK3D_DEBUG), 0 in production
  unsigned long long* trace;  // MK3D_TRACE: per-CTA role timers [grid][8], nullptr in production
  int vec32;       // 1: output rows are 32-byte aligned: the direct epilogue stores whole sectors (STG.256)
  int epi_stage;   // 1: epilogue transposes 32 x 64-byte blocks through shared memory and stores whole sectors (rows in slot order)
  int gather4;     // 1: the A rows are gathered by the TMA engine (tile::gather4 through `tmap`) instead of LDGSTS
  alignas(64) unsigned long long tmap[16];  // CUtensorMap of the input rows: {cin, n_in}, box {128 bytes of channels, 1}, SWIZZLE_128B
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


// Enumerates this CTA's (tile, offset k) work items in execution order: tiles blockIdx.x, +gridDim.x, ...; inside a tile the
// set bits of its offset mask, ascending.  The mask of the NEXT tile is loaded when the current one is entered, so the
// global-load latency of tile_kmask never sits on a role's critical path.
struct WorkCursor {
  const uint32_t* kmask;
  int tile, stride, num_tiles;
  uint32_t m, m_next;
  __device__ WorkCursor(const uint32_t* kmask_, int first, int stride_, int num_tiles_)
      : kmask(kmask_), tile(first), stride(stride_), num_tiles(num_tiles_) {
    m = tile < num_tiles ? kmask[tile] : 0u;
    m_next = tile + stride < num_tiles ? kmask[tile + stride] : 0u;
    skip_empty();
  }
  __device__ void skip_empty() {
    while (m == 0u && tile < num_tiles) {
      tile += stride;
      m = m_next;
      m_next = tile + stride < num_tiles ? kmask[tile + stride] : 0u;
    }
  }
  __device__ bool valid() const { return tile < num_tiles; }
  __device__ int k() const { return __ffs(m) - 1; }
  __device__ bool last_k_of_tile() const { return (m & (m - 1)) == 0u; }
  __device__ void next() {
    m &= m - 1;
    skip_empty();
  }
};


}  // namespace mk
