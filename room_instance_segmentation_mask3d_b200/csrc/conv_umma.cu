// tcgen05 gather-GEMM sparse convolution (bf16 or tf32 operands, fp32 accumulation in TMEM) for sm_100a.
//
// Replaces ME's per-offset gather -> GEMM -> atomic scatter launches (K launches x 3 per conv, SURVEY.md
// section 3.4) with ONE persistent, warp-specialised, output-stationary kernel:
//
//   tile       = 128 output slots (rows of the accumulator = TMEM lanes); a SUPER-TILE is G = 2 consecutive tiles when
//                cout <= 128 (two accumulators share every weight chunk)
//   iteration  = (super-tile, 128-byte chunk of the input row, kernel offset k with a neighbour in the super-tile)
//              = one ring slot
//   A operand  = the gathered input rows of the chunk, staged with 16-byte cp.async (LDGSTS, zero-fill for absent
//                neighbours) straight into the 128B-swizzled K-major shared-memory layout that tcgen05.mma reads
//   B operand  = W[k] chunk, pre-swizzled on the device once per weight ("weight image"), so a slot's weights are one
//                linear cp.async.bulk (TMA engine, UBLKCP) completing on the slot's mbarrier
//   D          = fp32 accumulators in TMEM, double-buffered: the epilogue of super-tile t (tcgen05.ld -> bias -> cast ->
//                whole-sector stores, each output row written exactly once, no atomics) overlaps the MMAs of t + 1
//
// Offsets whose neighbour column is empty for the whole tile are skipped via the per-tile bitmask from mk_tile_kmask, which
// is what keeps the MMA work close to the algorithmic 2*P*Cin*Cout on surface data (mask-sorted tile plans: engine.py).
//
// Warp roles (576 threads): warps 0-7 gather PRODUCERS (a group of `split` warps owns every NP-th iteration; a warp owns the
// 32-row units of one member tile), warps 8-9 MMA ISSUERS (one per member tile, one elected lane each; warp 9 also owns the
// TMEM allocation), warps 10-17 EPILOGUE (two per TMEM lane quarter: member tile 1 / upper column half go to the second set).
// DESIGN.md section 5 has the measurements behind every one of these choices (role timers, ncu, ablations, and the A/B
// records of the alternatives: TMA tile::gather4, hybrid gathers, 16 producer warps, blocked vs direct epilogue stores).
#include <cstdlib>
#include <vector>

#include "umma.cuh"

namespace mk {

// ---- thread layout of k_conv_umma (576 threads = 18 warps) ----
// A wider layout (768 threads: 16 gather warps, registers rebalanced with setmaxnreg) was measured and dropped: the same
// times within 3 % on every shape (profiles/r02_conv_sweep.txt) -- the gather is bound by the LSU's LDGSTS rate and the
// depth of the 4-slot ring, not by the number of warps issuing.
constexpr int kCvProducerWarps = 8;   // warps 0-7: gather producers (stage ownership in groups of `split` warps)
constexpr int kCvMmaWarp = 8;         // MMA issuer
constexpr int kCvAllocWarp = 9;       // TMEM allocation; second MMA issuer when G == 2
constexpr int kCvEpiWarp0 = 10;       // warps 10-17: epilogue (TMEM lane quarter = warp & 3), two warps per quarter
constexpr int kCvEpiWarps = 8;        // set 0 = warps 10-13, set 1 = warps 14-17: member tile 1 of a super-tile, or the upper columns when G == 1
constexpr int kCvThreads = 576;
constexpr int kMaxUW = 4;             // most 32-row units a producer warp gathers per stage

// keep the set bits of m whose ordinal is congruent to `part` modulo `ksplit` (offset split of small problems)
__device__ __forceinline__ uint32_t split_bits(uint32_t m, int part, int ksplit) {
  uint32_t r = 0;
  for (int i = 0; m; ++i) {
    const uint32_t b = m & (0u - m);
    if (i % ksplit == part) r |= b;
    m ^= b;
  }
  return r;
}

// Work enumeration of one CTA, identical in every role.  A "super-tile" is G consecutive 128-slot tiles (G = 2 when cout <= 128:
// the two accumulators share every weight chunk, and mask-sorted plans give neighbouring tiles the same offset set).  An
// iteration is (super-tile, 64-channel chunk c, offset k of the union of the member tiles' masks), ordered super-tile, chunk,
// offset.  Inside a chunk pass the offsets of a 3^3 kernel are visited x-plane by x-plane (all offsets with dx = -1, then
// dx = 0, then dx = +1): voxel rows are in lexicographic (x, y, z) order, so the 9 offsets of one dx gather (nearly) the SAME
// input rows, shifted by a few rows -- consecutive iterations then re-read rows that are still in the L1 left next to the
// pipeline stages, instead of re-fetching them from L2 (the kernel is bound by L2 -> SM bandwidth once its issue overheads
// are gone: ncu l1tex__m_xbar2l1tex_read_bytes, profiles/r02_ncu_conv96_summary.txt).
//
// The order is materialised once per super-tile and warp as a per-lane vector: lane j holds the j-th offset (`kvec`), so an
// iteration costs one shuffle + a handful of ALU instructions and no data-dependent loop (the first version of this kernel
// spent ~190 cycles per step in a bit-walking cursor, 4 steps per producer stage).
struct SuperTile {
  uint32_t m0, m1;  // member tile masks
  int n;            // offsets of this CTA's share of the union
  int kvec;         // lane j < n: j-th offset in visiting order
};
__device__ __forceinline__ void load_masks(const uint32_t* __restrict__ km, int s, int num_st, int num_tiles, bool two, uint32_t& a, uint32_t& b) {
  a = 0u;
  b = 0u;
  if (s < num_st) {
    const int t = two ? 2 * s : s;
    a = __ldg(km + t);
    if (two && t + 1 < num_tiles) b = __ldg(km + t + 1);
  }
}
__device__ __forceinline__ SuperTile make_super(uint32_t m0, uint32_t m1, int K, int part, int ks, int lane) {
  SuperTile t;
  t.m0 = m0;
  t.m1 = m1;
  uint32_t mall = m0 | m1;
  if (ks > 1) mall = split_bits(mall, part, ks);
  const uint32_t M0 = K == 27 ? 0x1249249u : 0xFFFFFFFFu;  // ix == 0 (k = ix + 3 * iy + 9 * iz)
  const uint32_t M1 = K == 27 ? 0x2492492u : 0u;           // ix == 1
  const uint32_t g0 = mall & M0, g1 = mall & M1, g2 = mall & ~(M0 | M1);
  const int n0 = __popc(g0), n1 = __popc(g1);
  t.n = __popc(mall);
  uint32_t g = g2;
  int r = lane - n0 - n1;
  if (lane < n0) {
    g = g0;
    r = lane;
  } else if (lane < n0 + n1) {
    g = g1;
    r = lane - n0;
  }
  t.kvec = lane < t.n ? (int)__fns(g, 0, r + 1) : 0;
  return t;
}

template <typename TOut, bool kDbg, bool kTf32>
__global__ void __launch_bounds__(kCvThreads, 1) k_conv_umma(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x G x A 16 KB][stages x B cout*128][barriers]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int S = p.num_stages;
  const int G = p.group;
  const int a_bytes = G * kABytes;
  const int b_bytes = p.cout * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)S * a_bytes;
  uint64_t* bars = (uint64_t*)(sB + (size_t)S * b_bytes);
  uint64_t* full_bar = bars;                       // [S]  32 * split producer lanes (copy completion) + 1 expect_tx arrive
  uint64_t* empty_bar = bars + kMaxStages;         // [S]  one tcgen05.commit per MMA issuer (G of them)
  uint64_t* tmem_full = bars + 2 * kMaxStages;     // [2]  one tcgen05.commit per MMA issuer
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]  128 epilogue threads
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);
  int32_t* idx_stage = (int32_t*)(bars + 2 * kMaxStages + 6);  // [producer warps][2][4 units][32] neighbour-table entries (1 KB per warp)
  float* out_stage = (float*)(idx_stage + kCvProducerWarps * 2 * 4 * 32);  // [8 epilogue warps][32 rows x out_cols] when p.stage_out

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dbg = kDbg ? p.debug : 0;
  const int acc_cols = G == 2 ? 128 : 256;  // TMEM columns per accumulator: [buffer][member tile] x acc_cols = 512
  const int split = p.split;                // producer warps per stage (1: one warp gathers both member tiles, 2: one warp per tile)
  // bias staged once per CTA: the epilogue adds it per element, and reading it from global memory there cost one load
  // instruction per output element and thread (the 1x1 convs with bias ran at a quarter of the HBM rate)
  __shared__ float s_bias[256];
  if (p.bias) {
    for (int i = threadIdx.x; i < p.out_cols; i += blockDim.x) s_bias[i] = p.bias[i];
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), p.gather4 ? 1 : 32 * split + 1);  // gather4: ONE arrive.expect_tx, the TMA engine completes the bytes
      mbar_init(smem_u32(&empty_bar[s]), G);  // one tcgen05.commit per MMA issuer
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), G);
      mbar_init(smem_u32(&tmem_empty[b]), kCvEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == kCvAllocWarp) {
    tmem_alloc(smem_u32(tmem_base_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int nchunks = p.nchunks;
  const int ksplit = p.ksplit;
  const int cta_first = (int)blockIdx.x / ksplit, cta_stride = (int)gridDim.x / ksplit, cta_part = (int)blockIdx.x % ksplit;

  if (warp < kCvProducerWarps) {
    // ===================== producers =====================
    // Iteration i is staged by producer group i mod NP into ring slot i mod S (NP <= S: the parity wait on a slot's empty
    // barrier is only unambiguous while a producer is at most ONE ring phase ahead of the consumer).  A group is `split` warps;
    // a stage has 4 * G units of 32 rows (member tile g, row quarter q), warp `sub` of the group gathers units
    // [sub * UW, (sub + 1) * UW), UW = 4 * G / split <= 4.
    //
    // Per unit a lane issues 8 LDGSTS: lane = (rsub, piece) copies the 16-byte piece `piece` of unit rows 4 * jj + rsub, jj = 0..7,
    // i.e. one warp instruction moves 4 consecutive rows x 128 B.  The unit's 32 table entries are staged in shared memory one own
    // stage ahead (4-byte cp.async per lane) and each lane reads its 8 entries back (broadcast LDS) -- no index shuffles: ncu's source view of the shuffle-based version showed every
    // LDGSTS waiting ~30 cycles on its own SHFL (volatile asm is not reordered), and the 64 SHFL per iteration took a quarter
    // of the LSU data pipe.
    const int NP = p.num_producers;
    const int gi = warp / split;
    const int sub = warp - gi * split;
    if (gi < NP) {
      const int UW = 4 * G / split;  // units per warp: 4 or 2
      const int piece = lane & 7;    // 16-byte piece of the 128-byte row
      const int rsub = lane >> 3;    // row octet of the unit this lane copies
      const uint32_t row_bytes = (uint32_t)p.cin * (uint32_t)p.elem;  // bf16 (2) or fp32 / tf32 (4) activations
      const int cin_bytes = p.cin * p.elem;
      const int64_t n_slots = p.n_slots;
      const int32_t* __restrict__ nbr = p.nbr;
      const int32_t* __restrict__ kidx = p.kidx;
      const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
      // One warp instruction copies unit rows 4 * jj + rsub (rsub = lane / 8): FOUR CONSECUTIVE rows.  Their swizzle phases (row & 7)
      // differ, so the 32-byte sectors of the four rows that arrive together land in disjoint banks and share a shared-memory
      // wavefront; with rows 8 apart (same phase) ncu counted 10.0 M LDGSTS bank conflicts per 96->96 launch instead of 5.9 M.
      const bool two = G == 2;
      const int num_st = two ? (p.num_tiles + 1) >> 1 : p.num_tiles;
      const uint32_t* __restrict__ kmask = p.tile_kmask;
      const int K = p.K;
      const uint32_t my_idx = smem_u32(idx_stage) + (uint32_t)(warp * 2 * 4 * 128);  // this warp's [2 buffers][4 units][32 entries]
      long long tp_wait = 0, tp_start = kDbg ? clock64() : 0;
      // ---- lookahead walker: position of this group's NEXT own iteration (may be in a later super-tile) ----
      // kept incrementally as (super-tile, chunk pass, offset ordinal): one iteration costs two compares, a shuffle and the table
      // address (the first version re-derived chunk and ordinal with a division loop: ~550 cycles per iteration in MK3D_TRACE)
      struct Own {
        int st, k, ch;
        uint32_t has;
        bool valid, tail;
        const int32_t* col;  // table entries of this warp's first unit: nbr[k][st * G * 128 + sub * UW * 32 ...]
      };
      int w_st = cta_first;   // super-tile of the walker
      int w_j = gi;           // offset ordinal inside the current chunk pass (may exceed T.n: carried on)
      int w_ch = 0;           // chunk pass
      uint32_t nm0, nm1;
      load_masks(kmask, cta_first, num_st, p.num_tiles, two, nm0, nm1);
      SuperTile T = make_super(nm0, nm1, K, cta_part, ksplit, lane);
      load_masks(kmask, cta_first + cta_stride, num_st, p.num_tiles, two, nm0, nm1);
      const int64_t n_full = n_slots - (int64_t)(G * kTileM);  // super-tiles starting beyond this are partial
      auto take = [&]() -> Own {  // resolve the walker's position to an iteration, then advance it by NP
        Own o;
        while (w_st < num_st && w_j >= T.n) {
          w_j -= T.n;
          if (T.n == 0 || ++w_ch == nchunks) {
            w_ch = 0;
            w_st += cta_stride;
            T = make_super(nm0, nm1, K, cta_part, ksplit, lane);
            load_masks(kmask, w_st + cta_stride, num_st, p.num_tiles, two, nm0, nm1);
            if (T.n <= 2 && sub == 0) {  // short passes (1x1 convs, mask logits): the table rows stream from DRAM, pull them into L2 early
              const int k0 = __shfl_sync(0xffffffffu, T.kvec, 0);
              const int64_t slot = ((int64_t)w_st + 6 * cta_stride) * (G * kTileM) + lane * 32;
              if (lane < 4 * G && slot < n_slots) asm volatile("prefetch.global.L2 [%0];" ::"l"(nbr + (int64_t)k0 * n_slots + slot));
            }
          }
        }
        o.valid = w_st < num_st;
        o.st = w_st;
        o.ch = w_ch;
        o.k = __shfl_sync(0xffffffffu, T.kvec, o.valid ? w_j : 0);
        o.has = o.valid ? (((T.m0 >> o.k) & 1u) | (((T.m1 >> o.k) & 1u) << 1)) : 0u;
        const int64_t slot0 = (int64_t)o.st * (G * kTileM);
        o.tail = slot0 > n_full;
        o.col = nbr + ((int64_t)o.k * n_slots + slot0 + sub * UW * 32);
        w_j += NP;
        return o;
      };
      if (p.gather4) {
        // ---- TMA row gather: the owning warp (split == 1) issues, per member tile, ONE tile::gather4 per lane = rows 4 * lane .. + 3
        // of the tile (their four table entries are one 16-byte load, issued one own iteration ahead and held in registers).  The TMA
        // engine fetches the four 128-byte row pieces from L2, writes them in the 128B swizzle and zero-fills absent neighbours
        // (index -1 is outside the tensor map), so the gather costs the SM 2 instructions per 4 rows instead of 8 LDGSTS + their
        // address arithmetic, and none of the LSU data-pipe wavefronts that bounded the LDGSTS version
        // (profiles/r02_ncu_conv96_summary.txt: l1tex lsu wavefronts 65 % with the tensor pipe at 28 %).
        const bool vec = (n_slots & 3) == 0;
        auto load_idx = [&](const Own& o, int g) -> int4 {
          int4 v = make_int4(-1, -1, -1, -1);
          if (o.valid && ((o.has >> g) & 1u)) {
            const int64_t slot = (int64_t)o.st * (G * kTileM) + g * kTileM + 4 * lane;
            const int32_t* col = nbr + (int64_t)o.k * n_slots;
            if (vec) {
              if (slot < n_slots) v = ldg_nc_v4(col + slot);
            } else {
              if (slot + 0 < n_slots) v.x = __ldg(col + slot + 0);
              if (slot + 1 < n_slots) v.y = __ldg(col + slot + 1);
              if (slot + 2 < n_slots) v.z = __ldg(col + slot + 2);
              if (slot + 3 < n_slots) v.w = __ldg(col + slot + 3);
            }
          }
          return v;
        };
        int stage = gi % S;
        uint32_t phase = (uint32_t)((gi / S) & 1);
        const int box = 128 / p.elem;  // channels per 128-byte chunk
        Own cur = take();
        int4 c0 = load_idx(cur, 0), c1 = two ? load_idx(cur, 1) : make_int4(-1, -1, -1, -1);
        while (cur.valid) {
          const Own nxt = take();
          const int4 n0 = load_idx(nxt, 0), n1 = two ? load_idx(nxt, 1) : make_int4(-1, -1, -1, -1);
          const int k = cur.k, ch = cur.ch;
          const uint32_t has = cur.has;
          const int kw = kidx ? __ldg(kidx + k) : k;
          const long long tp0 = kDbg ? clock64() : 0;
          mbar_wait_hint(smem_u32(&empty_bar[stage]), phase ^ 1, 20000u);
          if (kDbg) tp_wait += clock64() - tp0;
          const uint32_t bar = smem_u32(&full_bar[stage]);
          const uint32_t a_dst = sA_u32 + (uint32_t)(stage * a_bytes) + (uint32_t)(lane * 512);
          const bool gather = !(kDbg && (dbg & 1));
          if (lane == 0) {
            const uint32_t a_tx = gather ? (uint32_t)__popc(has) * (uint32_t)kABytes : 0u;
            if (kDbg && (dbg & 2)) {
              mbar_arrive_expect_tx(bar, a_tx);
            } else {
              mbar_arrive_expect_tx(bar, a_tx + (uint32_t)b_bytes);
              const uint8_t* wsrc = (const uint8_t*)p.wimg + ((int64_t)kw * nchunks + ch) * (int64_t)b_bytes;
              bulk_copy_g2s(sB_u32 + (uint32_t)(stage * b_bytes), wsrc, (uint32_t)b_bytes, bar);
            }
          }
          if (gather) {
            if (has & 1u) tma_gather4(a_dst, p.tmap, ch * box, c0.x, c0.y, c0.z, c0.w, bar);
            if (has & 2u) tma_gather4(a_dst + (uint32_t)kABytes, p.tmap, ch * box, c1.x, c1.y, c1.z, c1.w, bar);
          }
          __syncwarp();
          stage += NP;
          while (stage >= S) {
            stage -= S;
            phase ^= 1;
          }
          cur = nxt;
          c0 = n0;
          c1 = n1;
        }
        if (kDbg && p.trace && warp == 0 && lane == 0) {
          p.trace[blockIdx.x * 16 + 0] = (unsigned long long)tp_wait;
          p.trace[blockIdx.x * 16 + 1] = (unsigned long long)(clock64() - tp_start);
        }
      } else {
        // ---- LDGSTS gather ----
        // All units of a warp belong to ONE member tile (UW <= 4), so "does this warp have work in the iteration" is one
        // warp-uniform test, and everything inside is straight-line code: MK3D_TRACE of the previous version (per-unit predicates,
        // 8 scalar index reads per unit, bounds checks in every unit) showed 59 % of a producer iteration outside the gather issue.
        const int member = (sub * UW) >> 2;
        // a lane stages the table entry of unit row L = lane at position perm(L) = (L & 3) * 8 + (L >> 2): the eight rows
        // 4 * jj + rsub a lane copies (wide chunks) are then the 8 consecutive entries rsub * 8 .., read with two LDS.128; for
        // chunks of <= 64 bytes (rows 8 * jj + q, q = lane >> 2) they are entries (q & 3) * 8 + 2 * jj + (q >> 2)
        const uint32_t st_off = (uint32_t)((((lane & 3) << 3) | (lane >> 2)) << 2);
        const int q8 = lane >> 2;
        auto stage_indices = [&](const Own& o, int buf) {
          if ((o.has >> member) & 1u) {
            const uint32_t dst = my_idx + (uint32_t)(buf * 512) + st_off;
            if (!o.tail) {
#pragma unroll
              for (int u = 0; u < kMaxUW; ++u)
                if (u < UW) cp_async_4(dst + (uint32_t)(u * 128), o.col + u * 32 + lane, 4u);
            } else {  // partial super-tile at the end of the plan: rows past its end gather nothing
              const int64_t slot = (int64_t)o.st * (G * kTileM) + sub * UW * 32 + lane;
#pragma unroll
              for (int u = 0; u < kMaxUW; ++u) {
                if (u < UW) {
                  if (slot + u * 32 < n_slots) cp_async_4(dst + (uint32_t)(u * 128), o.col + u * 32 + lane, 4u);
                  else sts_s32(dst + (uint32_t)(u * 128), -1);
                }
              }
            }
          }
          cp_async_commit();
        };
        // pull the table lines of an iteration into L1 (not tracked by cp.async groups / the stage barrier: see below)
        auto prefetch_indices = [&](const Own& o) {
          if (o.valid && !o.tail && lane < UW && ((o.has >> member) & 1u)) asm volatile("prefetch.global.L1 [%0];" ::"l"(o.col + lane * 32));
        };
        int stage = gi % S;
        uint32_t phase = (uint32_t)((gi / S) & 1);
        int ibuf = 0;
        // Three iterations are in flight per producer: `cur` is gathered now; the table entries of `nxt` are copied to shared memory
        // (cp.async, 4 B per lane and unit) while cur's rows are gathered; the table LINES of `nx2` are prefetched into L1 one
        // iteration earlier still.  The cp.async of nxt's entries precede cur's gathers in this thread's cp.async order, so the
        // stage's completion-triggered mbarrier arrive also waits for them: they must be L1 hits (~50 cycles), not L2 / DRAM round
        // trips (ncu + MK3D_TRACE: with the entries fetched from L2 every stage completed ~one memory latency late, and with only 4
        // ring slots that latency is paid once per 4 iterations by the whole pipeline).
        Own cur = take();
        Own nxt = take();
        prefetch_indices(nxt);
        if (cur.valid) stage_indices(cur, ibuf);
        long long tq_take = 0, tq_stage = 0, tq_lds = 0, tq_issue = 0, tq_arr = 0, tq0 = 0;
        // per-lane constants of the two copy shapes
        const uint32_t a_wide = sA_u32 + (uint32_t)(sub * UW * 32 * 128 + rsub * 128);  // + stage, unit, jj * 512, swizzle
        const uint32_t sw_even = (uint32_t)((piece ^ rsub) << 4), sw_odd = (uint32_t)((piece ^ (4 + rsub)) << 4);  // row & 7 = 4 * (jj & 1) + rsub
        const uint32_t a_nar = sA_u32 + (uint32_t)(sub * UW * 32 * 128 + q8 * 128) + (uint32_t)(((lane & 3) ^ q8) << 4);  // row & 7 = q8 for every jj
        const uint32_t rd_wide = my_idx + (uint32_t)(rsub * 32), rd_nar = my_idx + (uint32_t)((q8 & 3) * 32);
        const bool odd_entry = (q8 >> 2) != 0;
        while (cur.valid) {
          if (kDbg) tq0 = clock64();
          const Own nx2 = take();
          prefetch_indices(nx2);
          if (kDbg) {
            const long long t = clock64();
            tq_take += t - tq0;
            tq0 = t;
          }
          if (nxt.valid) {
            stage_indices(nxt, ibuf ^ 1);
            cp_async_wait<1>();  // everything but the group just committed: this iteration's entries have landed
          } else {
            cp_async_wait<0>();
          }
          __syncwarp();  // the entries were written by other lanes of this warp
          if (kDbg) {
            const long long t = clock64();
            tq_stage += t - tq0;
            tq0 = t;
          }
          const int ch = cur.ch;
          const bool mine = ((cur.has >> member) & 1u) != 0;
          const int width = min(128, cin_bytes - ch * 128);  // valid bytes of this 128-byte chunk
          // chunks of <= 64 valid bytes (32-channel layers, the second chunk of 96 channels): a warp instruction copies EIGHT
          // consecutive rows x 64 B (lane = (row, 16-byte piece 0..3)) -- 4 LDGSTS per unit instead of 8 half-empty ones
          const bool narrow = width <= 64;
          const uint32_t rd = (narrow ? rd_nar : rd_wide) + (uint32_t)(ibuf * 512);
          int4 e0[kMaxUW], e1[kMaxUW];
          if (mine) {
#pragma unroll
            for (int u = 0; u < kMaxUW; ++u) {
              if (u < UW) {
                e0[u] = lds_v4i(rd + (uint32_t)(u * 128));
                e1[u] = lds_v4i(rd + (uint32_t)(u * 128 + 16));
              }
            }
          }
          ibuf ^= 1;
          const long long tp0 = kDbg ? clock64() : 0;
          if (kDbg) tq_lds += tp0 - tq0;
          mbar_wait_hint(smem_u32(&empty_bar[stage]), phase ^ 1, 20000u);
          if (kDbg) {
            tq0 = clock64();
            tp_wait += tq0 - tp0;
          }
          const uint32_t bar = smem_u32(&full_bar[stage]);
          if (sub == 0 && elect_one_sync()) {  // the weight chunk first: its bulk copy runs beside the gathers
            if (kDbg && (dbg & 2)) {
              mbar_arrive(bar);
            } else {
              const int kw = kidx ? __ldg(kidx + cur.k) : cur.k;
              mbar_arrive_expect_tx(bar, (uint32_t)b_bytes);
              const uint8_t* wsrc = (const uint8_t*)p.wimg + (uint64_t)(uint32_t)(kw * nchunks + ch) * (uint64_t)(uint32_t)b_bytes;
              bulk_copy_g2s(sB_u32 + (uint32_t)(stage * b_bytes), wsrc, (uint32_t)b_bytes, bar);
            }
          }
          __syncwarp();
          const int pc = narrow ? (lane & 3) : piece;
          // lanes whose piece lies beyond a narrow chunk (or all lanes in the gather-off ablation) issue nothing
          if (mine && pc * 16 < width && !(dbg & 1)) {
            // source address = per-lane 64-bit base + row * row_bytes as ONE 32x32+64 multiply-add (IMAD.WIDE.U32)
            uint64_t src0 = (uint64_t)p.in + (uint64_t)(ch * 128 + pc * 16);
            asm volatile("" : "+l"(src0));  // keep the base in one register pair (otherwise it is re-derived from the constant bank per copy)
            if (narrow) {
              const uint32_t a0 = a_nar + (uint32_t)(stage * a_bytes);
#pragma unroll
              for (int u = 0; u < kMaxUW; ++u) {
                if (u < UW) {
                  const int r4[4] = {odd_entry ? e0[u].y : e0[u].x, odd_entry ? e0[u].w : e0[u].z, odd_entry ? e1[u].y : e1[u].x,
                                     odd_entry ? e1[u].w : e1[u].z};
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    const int r = r4[jj];
                    cp_async_16(a0 + (uint32_t)(u * 32 * 128 + jj * 1024), (const void*)(src0 + (uint64_t)(r < 0 ? 0u : (uint32_t)r) * row_bytes),
                                r < 0 ? 0u : 16u);
                  }
                }
              }
            } else {
              const uint32_t a0 = a_wide + (uint32_t)(stage * a_bytes);
#pragma unroll
              for (int u = 0; u < kMaxUW; ++u) {
                if (u < UW) {
                  const int r8[8] = {e0[u].x, e0[u].y, e0[u].z, e0[u].w, e1[u].x, e1[u].y, e1[u].z, e1[u].w};
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) {
                    const int r = r8[jj];
                    cp_async_16(a0 + (uint32_t)(u * 32 * 128 + jj * 512) + ((jj & 1) ? sw_odd : sw_even),
                                (const void*)(src0 + (uint64_t)(r < 0 ? 0u : (uint32_t)r) * row_bytes), r < 0 ? 0u : 16u);
                  }
                }
              }
            }
          }
          if (kDbg) {
            const long long t = clock64();
            tq_issue += t - tq0;
            tq0 = t;
          }
          // completion-triggered arrive: no wait_group / fence in the producer; the proxy hand-over is completed by the MMA
          // thread's acquire fence after it has observed the barrier phase
          cp_async_arrive_noinc(bar);
          stage += NP;
          if (stage >= S) {  // NP <= S
            stage -= S;
            phase ^= 1;
          }
          if (kDbg) tq_arr += clock64() - tq0;
          cur = nxt;
          nxt = nx2;
        }
        if (kDbg && p.trace && warp == 0 && lane == 0) {
          p.trace[blockIdx.x * 16 + 0] = (unsigned long long)tp_wait;
          p.trace[blockIdx.x * 16 + 1] = (unsigned long long)(clock64() - tp_start);
          p.trace[blockIdx.x * 16 + 8] = (unsigned long long)tq_take;
          p.trace[blockIdx.x * 16 + 9] = (unsigned long long)tq_stage;
          p.trace[blockIdx.x * 16 + 10] = (unsigned long long)tq_lds;
          p.trace[blockIdx.x * 16 + 11] = (unsigned long long)tq_issue;
          p.trace[blockIdx.x * 16 + 12] = (unsigned long long)tq_arr;
        }
        cp_async_wait<0>();
      }  // LDGSTS gather
    }
  } else if (warp < kCvEpiWarp0) {
    if (warp == kCvMmaWarp || (warp == kCvAllocWarp && G == 2)) {
    // ===================== MMA issuers (whole warp converged, one elected lane issues) =====================
    // Every lane runs the loop so that all control flow and descriptor arithmetic stays warp-uniform; only the tcgen05
    // instructions sit under elect_one_sync() (see umma.cuh).  With super-tiles (G == 2) there are TWO issuers, one per member
    // tile (independent accumulators): the issuing warp is a serial chain of ~8-cycle instructions (barrier wait, descriptor
    // arithmetic on the uniform datapath, 4 UTCHMMA, commit: ~500 cycles per iteration for 8 MMAs, ncu source view), above the
    // 384 cycles the 8 MMAs of a 96-column iteration take on the tensor pipe; split over two warps both halves overlap.  Both
    // wait on the stage's full barrier and both commit to its empty barrier (count = number of issuers).
    const int mg = warp == kCvMmaWarp ? 0 : 1;  // member tile this warp issues for (G == 1: the only one)
    const uint32_t idesc = kTf32 ? make_idesc_tf32(kTileM, p.cout) : make_idesc(kTileM, p.cout);
    const int cin_bytes = p.cin * p.elem;
    const uint32_t a_lo_base = smem_desc_lo(smem_u32(sA)) + (uint32_t)(mg * (kABytes >> 4)), b_lo_base = smem_desc_lo(smem_u32(sB));
    const uint32_t a_step = (uint32_t)(a_bytes >> 4), b_step = (uint32_t)(b_bytes >> 4);
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    const int nk_last = min(128, cin_bytes - (nchunks - 1) * 128) >> 5;  // MMAs of the last chunk (one per 32 bytes of K: 16 bf16 / 8 tf32)
    int slot = 0;
    uint32_t ph = 0;
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    long long tw_full = 0, tw_tmem = 0, t_start = kDbg ? clock64() : 0, n_it = 0;
    const bool two = G == 2;
    const int num_st = two ? (p.num_tiles + 1) >> 1 : p.num_tiles;
    const uint32_t* __restrict__ kmask = p.tile_kmask;
    const int K = p.K;
    uint32_t tmem_d = 0, started = 0, tfull_bar = 0;
    // one pipeline stage: wait for it, issue this member tile's MMAs, free it.  Called with a LITERAL slot in the 4-stage loop
    // below, so that barrier addresses and descriptor bases are base + immediate (the issuing warp is a serial chain of
    // ~8-cycle uniform-datapath instructions: every instruction removed from an iteration is ~8 cycles of issue latency)
    auto issue = [&](const int s, const uint32_t phs, const uint32_t mine, const int nk, const bool last) {
      const long long t1 = kDbg ? clock64() : 0;
      mbar_wait_hint(full0 + 8u * s, phs, 5000u);
      if (kDbg) {
        tw_full += clock64() - t1;
        ++n_it;
      }
      fence_proxy_async_acquire();  // producers' cp.async (generic proxy) writes -> visible to tcgen05.mma (async proxy)
      tc_fence_after();
      const uint32_t b_lo = b_lo_base + (uint32_t)s * b_step;
      const uint32_t a_lo = a_lo_base + (uint32_t)s * a_step;
      if (elect_one_sync()) {
        if (mine) {
          if (kDbg && (dbg & 4)) {
            if (!started) umma_lohi<kTf32>(tmem_d, a_lo, b_lo, kSmemDescHi, idesc, 0u);
          } else if (nk == 4) {
            // +32 bytes along K inside the 128-byte swizzle row == +2 in the (addr >> 4) field
            umma_lohi<kTf32>(tmem_d, a_lo, b_lo, kSmemDescHi, idesc, started);
            umma_lohi<kTf32>(tmem_d, a_lo + 2u, b_lo + 2u, kSmemDescHi, idesc, 1u);
            umma_lohi<kTf32>(tmem_d, a_lo + 4u, b_lo + 4u, kSmemDescHi, idesc, 1u);
            umma_lohi<kTf32>(tmem_d, a_lo + 6u, b_lo + 6u, kSmemDescHi, idesc, 1u);
          } else {
            for (int kk = 0; kk < nk; ++kk)
              umma_lohi<kTf32>(tmem_d, a_lo + 2u * kk, b_lo + 2u * kk, kSmemDescHi, idesc, (kk > 0) ? 1u : started);
          }
        }
        umma_commit(empty0 + 8u * s);  // frees the stage when these MMAs (and the other issuer's) retire
        if (last) umma_commit(tfull_bar);
      }
      __syncwarp();
      started |= mine;
    };
    uint32_t nm0, nm1;
    load_masks(kmask, cta_first, num_st, p.num_tiles, two, nm0, nm1);
    for (int st = cta_first; st < num_st; st += cta_stride) {
      const SuperTile T = make_super(nm0, nm1, K, cta_part, ksplit, lane);
      load_masks(kmask, st + cta_stride, num_st, p.num_tiles, two, nm0, nm1);
      if (T.n == 0) continue;
      {
        const long long t0 = kDbg ? clock64() : 0;
        mbar_wait_hint(tempty0 + 8u * acc_buf, acc_phase ^ 1, 20000u);
        if (kDbg) tw_tmem += clock64() - t0;
        tc_fence_after();
      }
      started = 0;  // 1: this warp's accumulator holds a partial sum
      tmem_d = tmem_base + (uint32_t)((acc_buf * G + mg) * acc_cols);
      tfull_bar = tfull0 + 8u * acc_buf;
      // bit j: this warp's member tile has the j-th offset of the pass
      const uint32_t mine_vec = __ballot_sync(0xffffffffu, lane < T.n && (((mg ? T.m1 : T.m0) >> T.kvec) & 1u));
      const int total = T.n * nchunks;
      const int full_its = total - T.n;  // iterations before the last chunk (4 MMAs each)
      int i = 0, j = 0;
#define MK_STEP(SLOT)                                                              \
  {                                                                                \
    issue(SLOT, ph, (mine_vec >> j) & 1u, i < full_its ? 4 : nk_last, i == total - 1); \
    ++i;                                                                           \
    if (++j == T.n) j = 0;                                                         \
  }
      if (S == 4) {
        while (i < total) {
          switch (slot) {
            case 0:
              MK_STEP(0);
              if (i == total) {
                slot = 1;
                break;
              }
            case 1:
              MK_STEP(1);
              if (i == total) {
                slot = 2;
                break;
              }
            case 2:
              MK_STEP(2);
              if (i == total) {
                slot = 3;
                break;
              }
            default:
              MK_STEP(3);
              slot = 0;
              ph ^= 1;
          }
        }
      } else {
        while (i < total) {
          MK_STEP(slot);
          if (++slot == S) {
            slot = 0;
            ph ^= 1;
          }
        }
      }
#undef MK_STEP
      if (++acc_buf == 2) {
        acc_buf = 0;
        acc_phase ^= 1;
      }
    }
    if (kDbg && p.trace && lane == 0 && mg == 0) {
      p.trace[blockIdx.x * 16 + 2] = (unsigned long long)tw_full;
      p.trace[blockIdx.x * 16 + 3] = (unsigned long long)tw_tmem;
      p.trace[blockIdx.x * 16 + 4] = (unsigned long long)(clock64() - t_start);
      p.trace[blockIdx.x * 16 + 5] = (unsigned long long)n_it;
    }
    }  // MMA issuer warps
  } else {
    // ===================== epilogue (4 warps) =====================
    // MK3D_DEBUG=8 (no stores) showed the dense launches (1x1 convs, mask logits) bound by the epilogue: one warp per TMEM lane
    // quarter walked both member tiles and all columns as ONE serial chain of TMEM load -> convert -> store (96->128: 0.130 ms
    // with stores, 0.072 without; the MMA issuers waited 64 % of their time for a free accumulator).  Two warps per quarter halve
    // the chain: set 1 takes member tile 1 of a super-tile, or the upper half of the columns when there is one tile per stage.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int eset = (warp - kCvEpiWarp0) >> 2;
    const int col_half = ((p.out_cols + 31) >> 5) << 4;  // G == 1: columns [0, col_half) to set 0, the rest to set 1
    const int c_lo = (G == 1 && eset == 1) ? col_half : 0;
    const int c_hi = (G == 1 && eset == 0) ? min(col_half, p.out_cols) : p.out_cols;
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    const int num_st = (p.num_tiles + G - 1) / G;
    auto load_row = [&](int tile) -> int64_t {
      const int64_t slot = (int64_t)tile * kTileM + q * 32 + lane;
      int64_t r = -1;
      if (tile < p.num_tiles && slot < p.n_slots) r = p.out_rows ? (int64_t)__ldg(p.out_rows + slot) : slot;
      return r >= p.n_out ? -1 : r;
    };
    auto load_mask = [&](int tile) -> uint32_t { return tile < p.num_tiles ? p.tile_kmask[tile] : 0u; };
    uint32_t km_next[2];
    int64_t orow_next[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int t = g < G ? cta_first * G + g : p.num_tiles;
      km_next[g] = cta_first < num_st ? load_mask(t) : 0u;
      orow_next[g] = cta_first < num_st ? load_row(t) : -1;
    }
    const bool vec_ok = ((p.out_stride * (int)sizeof(TOut)) & 15) == 0;
    long long te_wait = 0, te_start = kDbg ? clock64() : 0;
    for (int st = cta_first; st < num_st; st += cta_stride) {
      uint32_t km[2];
      int64_t orow[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        km[g] = km_next[g];
        orow[g] = orow_next[g];
        const int tn = g < G ? (st + cta_stride) * G + g : p.num_tiles;
        const bool vn = st + cta_stride < num_st;
        km_next[g] = vn ? load_mask(tn) : 0u;
        orow_next[g] = vn ? load_row(tn) : -1;
      }
      if (ksplit > 1) {
        // this CTA's share of the super-tile's offsets (same filter as the cursor of the other roles): a member tile takes
        // part iff one of ITS offsets is in the share; an empty share means no accumulator and nothing to add
        const uint32_t share = split_bits(km[0] | km[1], cta_part, ksplit);
        km[0] &= share;
        km[1] &= share;
      }
      const bool any = (km[0] | km[1]) != 0u;
      if (any) {
        const long long t0 = kDbg ? clock64() : 0;
        mbar_wait_hint(smem_u32(&tmem_full[acc_buf]), acc_phase, 100000u);
        if (kDbg) te_wait += clock64() - t0;
        tc_fence_after();
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g >= G || st * G + g >= p.num_tiles) continue;
        if (G == 2 && g != eset) continue;
        TOut* optr = orow[g] >= 0 ? (TOut*)p.out + orow[g] * p.out_stride : nullptr;
        const uint32_t tcol = (uint32_t)((acc_buf * G + g) * acc_cols);
        if (p.stage_out) {
          // Row pitch of e.g. 25 fp32 (mask logits): per-thread row stores would touch 32 sectors per instruction.  The
          // warp's 32 rows are one contiguous run of 32 * out_cols elements in global memory, so transpose through shared
          // memory (odd pitch => conflict free) and store the run with fully coalesced 128-byte instructions.
          if (G == 1 && eset == 1) continue;  // whole rows are staged by one warp
          const uint32_t stg = smem_u32(out_stage) + (uint32_t)(((eset << 2) + q) * 32 * p.out_cols) * 4u;  // this warp's staging block
          const uint32_t my_row = stg + (uint32_t)(lane * p.out_cols) * 4u;
          for (int c0 = 0; c0 < p.out_cols; c0 += 16) {
            uint32_t r[16];
            if (km[g] != 0u) {
              tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)c0, r);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.out_cols) r[j] = __float_as_uint(__uint_as_float(r[j]) + s_bias[c0 + j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.out_cols) sts_f32(my_row + (uint32_t)(c0 + j) * 4u, __uint_as_float(r[j]));
          }
          __syncwarp();
          const int64_t row0 = ((int64_t)(st * G + g)) * kTileM + q * 32;
          const int64_t nrows = min((int64_t)32, p.n_out - row0);
          if (nrows > 0 && !(dbg & 8)) {
            TOut* dst = (TOut*)p.out + row0 * p.out_cols;
            const int total = (int)nrows * p.out_cols;
            if constexpr (sizeof(TOut) == 4) {
              // the run starts 16-byte aligned (row0 is a multiple of 32 rows): 16-byte shared-memory reads and global stores,
              // ~7 of each per lane for a 32 x 25 block instead of 25 + 25 scalar ones (the epilogue, not the MMAs or the
              // gathers, bounded the 25-column mask-logit product: ~780 instructions per tile and warp)
              const int nvec = total >> 2;
              for (int i = lane; i < nvec; i += 32) reinterpret_cast<float4*>(dst)[i] = lds_v4(stg + (uint32_t)i * 16u);
              for (int i = (nvec << 2) + lane; i < total; i += 32) dst[i] = from_f32<TOut>(lds_f32(stg + (uint32_t)i * 4u));
            } else {
              for (int i = lane; i < total; i += 32) dst[i] = from_f32<TOut>(lds_f32(stg + (uint32_t)i * 4u));
            }
          }
          __syncwarp();
          continue;
        }
        if (p.epi_stage) {
          // Rows in slot order (dense launches: 1x1 convs, mask logits): per-thread row stores put 16 bytes per lane into 32
          // different rows, i.e. 32 half-filled sectors per instruction, and the MK3D_TRACE / MK3D_DEBUG=8 runs showed these
          // launches bound by exactly that (96->128: 0.130 ms with stores, 0.072 without; with 8 epilogue warps the gather
          // producers then stalled behind the store traffic in the LSU).  Blocks of 32 rows x 64 bytes go through shared memory
          // (row pitch 80 B) and are stored 8 rows x 64 contiguous bytes per instruction: whole sectors, half the L2 write
          // transactions.
          const int EB = p.epi_stage;                 // bytes of a row per block: 64 or 128 (what the shared-memory budget allows)
          const int CB = EB / (int)sizeof(TOut);      // columns per block
          const int pitch_s = EB + 16;                // staging row pitch: conflict-free 16-byte accesses both ways
          const uint32_t stg = smem_u32(out_stage) + (uint32_t)(((eset << 2) + q) * 32 * pitch_s);
          const uint32_t my_row = stg + (uint32_t)(lane * pitch_s);
          const int64_t row0 = ((int64_t)(st * G + g)) * kTileM + q * 32;
          const int lpr = EB >> 4;                    // lanes per row of a store instruction (4 or 8)
          const int rsub = lane / lpr, piece = lane - rsub * lpr, rpi = 32 / lpr;  // rows per instruction (8 or 4)
          const int64_t pitch = (int64_t)p.out_stride * (int)sizeof(TOut);
          for (int c0 = c_lo; c0 < c_hi; c0 += CB) {
            const int width = min(CB, c_hi - c0);  // a multiple of 16 columns (host)
            for (int h = 0; h < width; h += 32) {  // two TMEM loads in flight per wait
              uint32_t r[2][16];
              const bool second = h + 16 < width;
              if (km[g] != 0u) {
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)(c0 + h), r[0]);
                if (second) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)(c0 + h + 16), r[1]);
                tmem_ld_wait();
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[0][j] = r[1][j] = 0u;
              }
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                if (u == 0 || second) {
                  const int cb = h + 16 * u;
                  float v[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[u][j]) + (p.bias ? s_bias[c0 + cb + j] : 0.f);
                  if constexpr (sizeof(TOut) == 4) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) sts_v4(my_row + (uint32_t)((cb + j) * 4), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; j += 8) {
                      uint32_t w[4];
#pragma unroll
                      for (int t = 0; t < 4; ++t) {
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                        w[t] = *reinterpret_cast<uint32_t*>(&b2);
                      }
                      sts_v4(my_row + (uint32_t)((cb + j) * 2),
                             make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3])));
                    }
                  }
                }
              }
            }
            __syncwarp();
            if (piece * 16 < width * (int)sizeof(TOut) && !(dbg & 8)) {
              uint8_t* dst0 = (uint8_t*)p.out + ((int64_t)c0 * (int)sizeof(TOut) + piece * 16);
              for (int rr = rsub; rr < 32; rr += rpi) {
                if (row0 + rr < p.n_out) {
                  const float4 w4 = lds_v4(stg + (uint32_t)(rr * pitch_s + piece * 16));
                  *reinterpret_cast<float4*>(dst0 + (row0 + rr) * pitch) = w4;
                }
              }
            }
            __syncwarp();
          }
          continue;
        }
        // the TMEM load of the next 16 columns is in flight while these 16 are converted and stored
        uint32_t rn[16];
        if (km[g] != 0u && c_lo < c_hi) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)c_lo, rn);
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
          float v[16];
          if (km[g] != 0u) {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rn[j]);
            if (c0 + 16 < c_hi) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)(c0 + 16), rn);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
          if (p.bias && (ksplit == 1 || cta_part == 0)) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += (c0 + j < p.out_cols) ? s_bias[c0 + j] : 0.f;
          }
          if (ksplit > 1) {
            // offset split: add this CTA's partial sums (the bias comes from part 0 alone, also for tiles part 0 has no offset of;
            // rows a part did not touch add nothing)
            if constexpr (sizeof(TOut) == 4) {
              if (optr && (km[g] != 0u || (cta_part == 0 && p.bias))) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  if (c0 + j + 4 <= p.out_cols)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"((float*)optr + c0 + j), "f"(v[j]), "f"(v[j + 1]),
                                 "f"(v[j + 2]), "f"(v[j + 3])
                                 : "memory");
                }
              }
            }
          } else if (optr && !(dbg & 8)) {
            if (p.vec32 && c0 + 16 <= p.out_cols) {
              store16_sectors(optr + c0, v);  // a whole 32-byte sector per lane: half the L2 write transactions of 16-byte stores
            } else if (vec_ok && c0 + 16 <= p.out_cols) {
              store16(optr + c0, v);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.out_cols) optr[c0 + j] = from_f32<TOut>(v[j]);
            }
          }
        }
      }
      if (any) {
        tc_fence_before();
        mbar_arrive(smem_u32(&tmem_empty[acc_buf]));
        if (++acc_buf == 2) {
          acc_buf = 0;
          acc_phase ^= 1;
        }
      }
    }
    if (kDbg && p.trace && warp == kCvEpiWarp0 && lane == 0) {
      p.trace[blockIdx.x * 16 + 6] = (unsigned long long)te_wait;
      p.trace[blockIdx.x * 16 + 7] = (unsigned long long)(clock64() - te_start);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kCvAllocWarp) {
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

// fp32 (tf32) weight image: 32-element (128-byte) chunks, same 16-byte-piece XOR (row & 7) swizzle:
// img[((k * nchunks + ch) * rows + n) * 32 + swz(n, e)] = W'[k][ch*32 + e][n]
__global__ void __launch_bounds__(256) k_weight_image_tf32(const float* __restrict__ W, int K, int cin, int cout, int transpose,
                                                           float* __restrict__ img) {
  const int red = transpose ? cout : cin;
  const int rows = transpose ? cin : cout;
  const int nchunks = (red + 31) / 32;
  const int64_t total = (int64_t)K * nchunks * rows * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % 32);
    int64_t t = i / 32;
    const int n = (int)(t % rows);
    t /= rows;
    const int ch = (int)(t % nchunks);
    const int k = (int)(t / nchunks);
    const int r = ch * 32 + e;
    float v = 0.f;
    if (r < red) v = transpose ? W[((int64_t)k * cin + n) * cout + r] : W[((int64_t)k * cin + r) * cout + n];
    const int piece = e >> 2, within = e & 3;  // 16-byte piece = 4 fp32
    const int64_t dst = ((int64_t)(k * nchunks + ch) * rows + n) * 32 + (((piece ^ (n & 7)) << 2) | within);
    img[dst] = v;
  }
}

inline size_t conv_smem_bytes(int cout, int stages, int group = 1, int stage_out_cols = 0, int epi_stage = 0) {
  const size_t out_stage = epi_stage ? (size_t)kCvEpiWarps * 32 * (epi_stage + 16) : (size_t)kCvEpiWarps * 32 * stage_out_cols * sizeof(float);
  return 1024 + (size_t)stages * ((size_t)group * kABytes + (size_t)cout * 128) + (2 * kMaxStages + 6) * sizeof(uint64_t) + 16 +
         (size_t)kCvProducerWarps * 2 * 4 * 32 * sizeof(int32_t) + out_stage;
}

int conv_tma_launch(const ConvParams& p, int64_t n_in, int out_dtype, cudaStream_t st);  // conv_tma.cu
int encode_rows_tmap(void* out, const void* base, int elem, int cin, int64_t n_in);          // conv_tma.cu

// 0 = LSU producer: 16-byte cp.async (LDGSTS) into the swizzled A tile (default); 1 = TMA gather4 producer (mk_set_gather_mode(1);
// measured ~2.5x slower on random rows, kept for A/B measurements)
static int g_gather_mode = 0;
inline int gather_mode() { return g_gather_mode; }

// Pipeline-shape overrides and timing ablations for measurement sweeps.  Read from the environment ONLY in a tuning build
// (nvcc -DMK3D_TUNING, `MK3D_BUILD_TUNING=1 python -m room_instance_segmentation_mask3d_b200.build`); the production library
// reads no environment variable on this path.
//   MK3D_DEBUG bit mask (results are wrong when non-zero): 1 skip gathers, 2 skip weight copies, 4 one MMA per tile, 8 skip stores
//   MK3D_TRACE=1 per-role cycle counters; MK3D_GROUP, MK3D_KSPLIT, MK3D_STAGES, MK3D_PRODUCERS, MK3D_SPLIT, MK3D_CARVEOUT shape overrides
struct Tuning {
  int debug = 0, trace = 0, group = 0, ksplit = 0, stages = 0, producers = 0, split = 0, carveout = -1, epi_stage = 1, vec32 = 1;
};
inline const Tuning& tuning() {
  static Tuning t;
#ifdef MK3D_TUNING
  static bool init = false;
  if (!init) {
    auto geti = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
    t.debug = geti("MK3D_DEBUG", 0);
    t.trace = getenv("MK3D_TRACE") ? 1 : 0;
    t.group = geti("MK3D_GROUP", 0);
    t.ksplit = geti("MK3D_KSPLIT", 0);
    t.stages = geti("MK3D_STAGES", 0);
    t.producers = geti("MK3D_PRODUCERS", 0);
    t.split = geti("MK3D_SPLIT", 0);
    t.carveout = geti("MK3D_CARVEOUT", -1);
    t.epi_stage = geti("MK3D_EPI_STAGE", 1);
    t.vec32 = geti("MK3D_VEC32", 1);
    init = true;
  }
#endif
  return t;
}

// 1: no run-to-run variation of results (no offset split with REDs in the forward / dgrad); set by mk_set_deterministic
int g_deterministic = 0;

}  // namespace mk

using namespace mk;

extern "C" {

int mk_set_deterministic(int on) {
  g_deterministic = on ? 1 : 0;
  return 0;
}

int mk_set_gather_mode(int mode) {
  g_gather_mode = mode;
  return 0;
}

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

}  // extern "C"

// shared launcher of mk_conv_fwd_bf16 / mk_conv_fwd_tf32 (`in_dtype`: MK_BF16 -> kind::f16 on bf16 rows, MK_F32 -> kind::tf32 on fp32 rows)
static int conv_fwd_launch(const void* in, int in_dtype, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K,
                           int64_t n_slots, const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                           const float* bias, void* out, int out_dtype, int out_cols, int out_stride, void* stream_) {
  MK_REQUIRE(mk_device_supported(), "tcgen05 path needs an sm_100 device (B200)");
  MK_REQUIRE(in_dtype == MK_F32 ? (cin % 8 == 0 && cin >= 8) : (cin % 16 == 0 && cin >= 16), "cin must be a multiple of 16 (bf16) / 8 (tf32)");
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
  p.elem = in_dtype == MK_F32 ? 4 : 2;
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
  p.nchunks = (cin * p.elem + 127) / 128;
  p.out_cols = out_cols;
  p.out_stride = out_stride;
  const int elem = out_dtype == MK_F32 ? 4 : 2;
  p.stage_out = (out_rows == nullptr && out_stride == out_cols && ((out_stride * elem) & 15) != 0 && out_cols <= 64) ? 1 : 0;
  const int soc = p.stage_out ? out_cols : 0;
  p.vec32 = ((((uintptr_t)out) & 31) == 0 && ((out_stride * elem) & 31) == 0 && tuning().vec32 != 0) ? 1 : 0;
  // coalescing epilogue for launches whose rows are in slot order (dense: 1x1 convs, mask logits, unsorted plans)
  p.epi_stage = (!p.stage_out && out_rows == nullptr && (out_cols & 15) == 0 && ((out_stride * elem) & 15) == 0 && tuning().epi_stage != 0) ? 1 : 0;
  const Tuning& tn = tuning();
  p.debug = tn.debug;
  p.trace = nullptr;
  p.chunk_major = 1;
  if (gather_mode() == 1 && p.elem == 2) return conv_tma_launch(p, n_in, out_dtype, (cudaStream_t)stream_);
  // super-tiles of two tiles when both accumulators fit next to their double buffer (4 x 128 TMEM columns) and there are
  // enough tiles to keep every SM busy with the coarser work unit
  int group = (cout <= 128 && p.num_tiles >= 4 * num_sms()) ? 2 : 1;
  if (tn.group == 1 || (tn.group == 2 && cout <= 128)) group = tn.group;
  p.group = group;
  // Offset split for the coarse levels: a level with fewer super-tiles than SMs (L3 / L4 of a single scene: 7-27 tiles for 148
  // SMs) ran as 7-27 CTAs walking all 27 offsets each, pure latency.  With an fp32 output the offsets of a super-tile are dealt
  // to `ksplit` CTAs that add their partial sums with vector REDs (the output is zeroed here first).  The accumulation order of
  // those REDs varies from run to run: mk_set_deterministic(1) turns the split off (the trainer's `deterministic: true`).
  p.ksplit = 1;
  {
    const int num_super0 = (p.num_tiles + group - 1) / group;
    if (out_dtype == MK_F32 && K > 1 && !p.stage_out && (out_cols & 3) == 0 && ((out_stride * 4) & 15) == 0 && tn.ksplit != 1 &&
        !g_deterministic && 2 * num_super0 <= num_sms()) {
      int ks = num_sms() / num_super0;
      if (ks > K) ks = K;
      if (ks > 9) ks = 9;
      if (tn.ksplit > 1 && ks > tn.ksplit) ks = tn.ksplit;
      p.ksplit = ks;
    }
  }
  if (p.ksplit > 1) p.epi_stage = 0;  // partial sums are added with REDs from registers
  if (p.epi_stage) {
    // 128-byte blocks, for rows that are whole blocks: measured on the 889 k-voxel map (profiles/r02_conv_ablation.txt) 96->128
    // 0.149 -> 0.113 ms, 256->128 0.195 -> 0.133, 384->256 0.362 -> 0.251, 128->160 fp32 0.284 -> 0.199 -- also where the larger
    // staging area costs the ring its fourth slot; 64-byte blocks gain a third of that, and rows of 192 bytes (96 bf16 channels)
    // are 13 % slower than with direct stores, so they keep those.
    // Against direct stores of whole sectors (STG.256, added later): still ahead for bf16 rows of >= 256 bytes (96->128 0.113 vs
    // 0.120 ms, 256->128 0.133 vs 0.161, 384->256 0.248 vs 0.297), behind for fp32 rows (128->160: 0.199 vs 0.185) and 128-byte rows.
    p.epi_stage = (((out_cols * elem) & 127) == 0 && elem == 2 && out_cols * elem >= 256) ? 128 : 0;
    if (tuning().epi_stage == 64 || tuning().epi_stage == 128) p.epi_stage = tuning().epi_stage;
  }
  int stages = kMaxStages;
  while (stages > 2 && conv_smem_bytes(cout, stages, group, soc, p.epi_stage) > 222 * 1024) --stages;
  // Super-tiles: 4 stages (<= 176 KB) instead of the 5 that fit: the smaller carve-out leaves ~50 KB of L1 for the neighbour
  // rows a compact tile re-gathers across offsets (L1 sector hit rate 13 % -> 50+ %).
  if (group == 2 && stages > 4) stages = 4;
  MK_REQUIRE(conv_smem_bytes(cout, stages, group, soc, p.epi_stage) <= 227 * 1024, "shared memory budget exceeded");
  if (tn.stages >= 2 && tn.stages < stages) stages = tn.stages;  // tuning sweeps only
  // producer groups: one per ring slot (a group may be at most one ring phase ahead); two warps per group (one per member
  // tile) when there are member tiles to split and warps to spare
  const int pw = kCvProducerWarps;
  p.split = group;  // a producer warp gathers at most 4 units of 32 rows
  while (p.split < 4 && 2 * p.split * stages <= pw) p.split *= 2;  // 4 stages: 2 warps per stage
  if (tn.split >= group && tn.split <= 4) p.split = tn.split;
  p.num_producers = stages < pw / p.split ? stages : pw / p.split;
  if (tn.producers >= 1 && tn.producers < p.num_producers) p.num_producers = tn.producers;
  p.gather4 = 0;
  if (gather_mode() == 2 && ((cin * p.elem) & 15) == 0 && ((uintptr_t)nbr & 15) == 0) {  // TMA row gather inside this kernel (A/B switch)
    MK_REQUIRE(encode_rows_tmap(p.tmap, in, p.elem, cin, n_in) == 0, "cuTensorMapEncodeTiled failed");
    p.gather4 = 1;
    p.split = 1;
    p.num_producers = stages < 4 ? stages : 4;
    if (tn.producers >= 1 && tn.producers < p.num_producers) p.num_producers = tn.producers;
  }
  p.num_stages = stages;
  const size_t smem = conv_smem_bytes(cout, stages, group, soc, p.epi_stage);
  int grid = num_sms();
  const int num_super = (p.num_tiles + group - 1) / group;
  if (grid > num_super) grid = num_super;
  if (p.ksplit > 1) {
    grid = num_super * p.ksplit;
    // only the columns this call owns (the output may be a column slice of a wider row, e.g. the sliced dgrad of 384 channels)
    cudaMemset2DAsync(out, (size_t)out_stride * sizeof(float), 0, (size_t)out_cols * sizeof(float), (size_t)n_out, (cudaStream_t)stream_);
  }
  cudaStream_t st = (cudaStream_t)stream_;
#define MK_LAUNCH1(T, D, F)                                                                                  \
  do {                                                                                                      \
    static size_t smem_set = 0; /* per instantiation: raise the limit only when a launch needs more */      \
    if (smem > smem_set) {                                                                                  \
      cudaFuncSetAttribute(k_conv_umma<T, D, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
      smem_set = smem;                                                                                      \
    }                                                                                                       \
    if (tn.carveout >= 0) cudaFuncSetAttribute(k_conv_umma<T, D, F>, cudaFuncAttributePreferredSharedMemoryCarveout, tn.carveout); \
    k_conv_umma<T, D, F><<<grid, kCvThreads, smem, st>>>(p);                                                \
  } while (0)
#define MK_LAUNCH(T, D)                 \
  do {                                  \
    if (p.elem == 4) MK_LAUNCH1(T, D, true); \
    else MK_LAUNCH1(T, D, false);       \
  } while (0)
#ifdef MK3D_TUNING
  if (tn.trace) {
    cudaMalloc(&p.trace, (size_t)grid * 16 * sizeof(unsigned long long));
    cudaMemset(p.trace, 0, (size_t)grid * 16 * sizeof(unsigned long long));
  }
  const bool dbg_build = tn.trace || p.debug != 0;
  if (out_dtype == MK_F32) {
    if (dbg_build) MK_LAUNCH(float, true);
    else MK_LAUNCH(float, false);
  } else {
    if (dbg_build) MK_LAUNCH(__nv_bfloat16, true);
    else MK_LAUNCH(__nv_bfloat16, false);
  }
  if (tn.trace) {
    std::vector<unsigned long long> h((size_t)grid * 16);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), p.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    double a[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int j = 0; j < 16; ++j) a[j] += (double)h[(size_t)b * 16 + j] / grid;
    fprintf(stderr, "[mk3d trace] cin %d cout %d K %d tiles %d stages %d producers %dx%d | producer0: wait_empty %.0f / total %.0f | mma: wait_full %.0f wait_tmem %.0f / total %.0f, iters %.0f | epilogue: wait_full %.0f / total %.0f (cycles, CTA mean)\n",
            cin, cout, K, p.num_tiles, stages, p.num_producers, p.split, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
    fprintf(stderr, "[mk3d trace]   producer0 phases: walker %.0f | index staging + wait_group %.0f | index reads %.0f | gather issue %.0f | arrive + weight copy %.0f\n",
            a[8], a[9], a[10], a[11], a[12]);
  }
#else
  if (out_dtype == MK_F32) MK_LAUNCH(float, false);
  else MK_LAUNCH(__nv_bfloat16, false);
#endif
#undef MK_LAUNCH
#undef MK_LAUNCH1
  return check_launch(__func__);
}

extern "C" {

int mk_conv_fwd_bf16(const void* in, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K, int64_t n_slots,
                     const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                     const float* bias, void* out, int out_dtype, int out_cols, int out_stride, void* stream_) {
  return conv_fwd_launch(in, MK_BF16, n_in, cin, wimg, cout, nbr, K, n_slots, out_rows, tile_kmask, kidx, n_out, bias, out, out_dtype, out_cols,
                         out_stride, stream_);
}

int mk_conv_fwd_tf32(const void* in, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K, int64_t n_slots,
                     const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                     const float* bias, void* out, int out_dtype, int out_cols, int out_stride, void* stream_) {
  return conv_fwd_launch(in, MK_F32, n_in, cin, wimg, cout, nbr, K, n_slots, out_rows, tile_kmask, kidx, n_out, bias, out, out_dtype, out_cols,
                         out_stride, stream_);
}

int64_t mk_conv_weight_image_tf32_elems(int K, int cin, int cout, int transpose) {
  const int red = transpose ? cout : cin, rows = transpose ? cin : cout;
  return (int64_t)K * ((red + 31) / 32) * rows * 32;
}

int mk_conv_weight_image_tf32(const float* W, int K, int cin, int cout, int transpose, void* img, void* stream_) {
  const int64_t total = mk_conv_weight_image_tf32_elems(K, cin, cout, transpose);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  k_weight_image_tf32<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(W, K, cin, cout, transpose, (float*)img);
  return check_launch(__func__);
}

}  // extern "C"
