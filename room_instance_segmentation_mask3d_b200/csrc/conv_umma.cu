// tcgen05 gather-GEMM sparse convolution (bf16 operands, fp32 accumulation in TMEM) for sm_100a.
//
// Replaces ME's per-offset gather -> GEMM -> atomic scatter launches (K launches x 3 per conv, SURVEY.md
// section 3.4) with ONE persistent, warp-specialised, output-stationary kernel:
//
//   tile      = 128 output slots (rows of the accumulator = TMEM lanes)
//   iteration = (kernel offset k with at least one neighbour in the tile, 64-channel chunk of Cin) = one pipeline stage
//   A operand = the 128 gathered input rows of the chunk, staged with 16-byte cp.async (LDGSTS, zero-fill for absent
//               neighbours) straight into the 128B-swizzled K-major shared-memory layout that tcgen05.mma reads
//   B operand = W[k] chunk, pre-swizzled on the device once per weight ("weight image"), so a stage is one
//               linear cp.async.bulk (TMA engine, UBLKCP) completing on the stage's mbarrier
//   D         = fp32 accumulator in TMEM, two buffers of `cout` columns: the epilogue of tile t
//               (tcgen05.ld -> bias -> cast -> 16-byte stores, each output row written exactly once,
//               no atomics) overlaps the MMAs of tile t+1
//
// Offsets whose neighbour column is empty for the whole tile are skipped via the per-tile bitmask from
// mk_tile_kmask, which is what keeps the MMA work close to the algorithmic 2*P*Cin*Cout on surface data.
//
// Warp roles (320 threads): warps 0-3 and 5 are PRODUCERS, warp 4 the MMA issuer (one elected lane), warps 6-9 the
// epilogue (TMEM lane quarter = warp % 4); warp 5 also owns the TMEM allocation.
//
// Producers own whole stages round-robin (stage i belongs to producer i mod 5): ncu's per-instruction stall samples of the
// previous design (all 4 producer warps cooperating on EVERY stage, profiles/r01_conv96_*) showed the producer loop —
// mbarrier try_wait, index shuffles, LDGSTS issue throttling, exposed index loads — at ~1100 cycles per stage on every
// producer warp, i.e. the gather front-end, not the tensor pipe or L2, bounded the kernel.  With stage ownership the same
// fixed costs overlap across five stages in flight.  Per stage the owning warp: waits for the stage to be free, gathers
// 128 rows x 128 B with 32 LDGSTS per lane (neighbour indices held 4 per lane, distributed by shuffle, prefetched two of its
// own stages ahead), lets the copies arrive on the stage's full barrier (cp.async.mbarrier.arrive.noinc: no thread-side
// wait), and lane 0 adds the weight-image bulk copy with its expect_tx on the same barrier.
#include <cstdlib>
#include <vector>

#include "umma.cuh"

namespace mk {

constexpr int kMaxProducers = 5;  // warps 0-3 and 5

// Work enumeration of one CTA.  A "super-tile" is G consecutive 128-slot tiles (G = 2 when cout <= 128: the two accumulators
// share every weight chunk, which halves the weight traffic and the per-stage bookkeeping of all three roles; mask-sorted
// plans give neighbouring tiles the same offset set, so the union costs nothing).  Items are the set bits of the union of the
// member tiles' offset masks, ascending; an iteration is (item, 64-channel chunk).
// keep the set bits of m whose ordinal is congruent to `part` modulo `ksplit` (offset split of small problems)
__device__ __forceinline__ uint32_t split_bits(uint32_t m, int part, int ksplit) {
  if (ksplit <= 1) return m;
  uint32_t r = 0;
  for (int i = 0; m; ++i) {
    const uint32_t b = m & (0u - m);
    if (i % ksplit == part) r |= b;
    m ^= b;
  }
  return r;
}

struct SuperCursor {
  const uint32_t* kmask;
  int G, st, stride, num_st, num_tiles;
  int part, ksplit;
  uint32_t m0, m1, mu;      // member masks of the current super-tile, remaining union bits (of the current chunk pass)
  uint32_t n0, n1;          // masks of the next super-tile (prefetched)
  int ch, nchunks;
  int chunk_major;          // 1: all offsets of chunk 0, then all offsets of chunk 1, ... ; 0: (offset, chunk) with chunk inner
  __device__ uint32_t mask_of(int t) const { return t < num_tiles ? kmask[t] : 0u; }
  __device__ SuperCursor(const uint32_t* kmask_, int G_, int first, int stride_, int num_tiles_, int nchunks_, int chunk_major_,
                         int part_ = 0, int ksplit_ = 1)
      : kmask(kmask_), G(G_), st(first), stride(stride_), num_tiles(num_tiles_), part(part_), ksplit(ksplit_), ch(0), nchunks(nchunks_),
        chunk_major(chunk_major_) {
    num_st = (num_tiles + G - 1) / G;
    load(st, m0, m1);
    load(st + stride, n0, n1);
    mu = split_bits(m0 | m1, part, ksplit);
    skip_empty();
  }
  __device__ uint32_t all_bits() const { return split_bits(m0 | m1, part, ksplit); }
  __device__ void load(int s, uint32_t& a, uint32_t& b) const {
    a = s < num_st ? mask_of(G * s) : 0u;
    b = (s < num_st && G == 2) ? mask_of(2 * s + 1) : 0u;
  }
  __device__ void next_super() {
    st += stride;
    m0 = n0;
    m1 = n1;
    mu = all_bits();
    ch = 0;
    load(st + stride, n0, n1);
  }
  __device__ void skip_empty() {
    while (mu == 0u && st < num_st) next_super();
  }
  __device__ bool valid() const { return st < num_st; }
  __device__ int k() const { return __ffs(mu) - 1; }
  __device__ bool has(int g) const { return (((g ? m1 : m0) >> k()) & 1u) != 0u; }
  // true on the last iteration of the super-tile (its accumulators are complete after it)
  __device__ bool last_iter() const { return (mu & (mu - 1)) == 0u && ch == nchunks - 1; }
  __device__ void step() {  // one iteration
    if (chunk_major) {
      mu &= mu - 1;
      if (mu == 0u) {
        if (++ch < nchunks) mu = all_bits();
        else {
          next_super();
          skip_empty();
        }
      }
    } else {
      if (++ch == nchunks) {
        ch = 0;
        mu &= mu - 1;
        if (mu == 0u) {
          next_super();
          skip_empty();
        }
      }
    }
  }
  __device__ void advance(int n) {  // n iterations
    for (int i = 0; i < n && valid(); ++i) step();
  }
};

template <typename TOut, bool kDbg>
__global__ void __launch_bounds__(kNumThreads, 1) k_conv_umma(const ConvParams p) {
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
  uint64_t* full_bar = bars;                       // [S]  32 producer lanes (copy completion) + 1 expect_tx arrive
  uint64_t* empty_bar = bars + kMaxStages;         // [S]  one tcgen05.commit
  uint64_t* tmem_full = bars + 2 * kMaxStages;     // [2]  one tcgen05.commit
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;  // [2]  128 epilogue threads
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 2 * kMaxStages + 4);
  float* out_stage = (float*)(bars + 2 * kMaxStages + 6);  // [4 epilogue warps][32 rows x out_cols] when p.stage_out (16-byte aligned)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dbg = kDbg ? p.debug : 0;
  const int acc_cols = G == 2 ? 128 : 256;  // TMEM columns per accumulator: [buffer][member tile] x acc_cols = 512
  // bias staged once per CTA: the epilogue adds it per element, and reading it from global memory there cost one load
  // instruction per output element and thread (the 1x1 convs with bias ran at a quarter of the HBM rate)
  __shared__ float s_bias[256];
  if (p.bias) {
    for (int i = threadIdx.x; i < p.out_cols; i += blockDim.x) s_bias[i] = p.bias[i];
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 32 + 1);
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
  const int ksplit = p.ksplit;
  const int cta_first = (int)blockIdx.x / ksplit, cta_stride = (int)gridDim.x / ksplit, cta_part = (int)blockIdx.x % ksplit;

  if (warp < kNumProducerWarps || warp == kLoadWarp) {
    // ===================== producers (up to 5 warps, one stage each, round-robin) =====================
    const int pi = warp < kNumProducerWarps ? warp : kNumProducerWarps;
    // Stage ownership: iteration i is staged by producer i mod NP into ring slot i mod S.  NP <= S is required: the
    // parity wait on a slot's empty barrier is only unambiguous while a producer is at most ONE ring phase ahead of the
    // consumer; having passed the wait of iteration i - NP it knows i - NP - S and everything earlier was consumed, which
    // covers i - 2S iff S >= NP.
    const int NP = p.num_producers;
    const int piece = lane & 7;  // 16-byte piece of the 128-byte row
    const int rsub = lane >> 3;  // 4 rows per warp instruction
    // swizzled piece offset inside the row: row & 7 == 4 * (jj & 1) + rsub for row = 4 * jj + rsub
    const uint32_t sw_even = (uint32_t)((piece ^ rsub) << 4), sw_odd = (uint32_t)((piece ^ (4 + rsub)) << 4);
    const uint32_t row_bytes = (uint32_t)p.cin * (uint32_t)sizeof(__nv_bfloat16);
    struct Idx {
      int r[2][4];  // per member tile: input rows gathered into tile rows lane, lane + 32, lane + 64, lane + 96
      int kw;       // weight-image index of the item's offset
      int ch;       // 64-channel chunk of the iteration
      uint32_t has; // bit g: member tile g has this offset
      bool valid;
    };
    auto load_idx = [&](const SuperCursor& c) -> Idx {
      Idx x;
      const bool v = c.valid();
      const int k = v ? c.k() : 0;
      x.valid = v;
      x.ch = c.ch;
      x.has = v ? ((c.has(0) ? 1u : 0u) | (c.has(1) ? 2u : 0u)) : 0u;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t slot = ((int64_t)c.st * G + g) * kTileM + q * 32 + lane;
          x.r[g][q] = (((x.has >> g) & 1u) && slot < p.n_slots) ? __ldg(p.nbr + (int64_t)k * p.n_slots + slot) : -1;
        }
      }
      x.kw = (v && p.kidx) ? __ldg(p.kidx + k) : k;
      return x;
    };
    const bool active = pi < NP;  // surplus producer warps idle
    int stage = pi % S;
    uint32_t phase = (uint32_t)((pi / S) & 1);
    long long tp_wait = 0, tp_start = kDbg ? clock64() : 0;
    auto stage_body = [&](const Idx& ci) {
      const int ch = ci.ch;
      const int width = min(kChunkElems, p.cin - ch * kChunkElems);  // valid elements in this chunk
      const long long tp0 = kDbg ? clock64() : 0;
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      if (kDbg) tp_wait += clock64() - tp0;
      const uint32_t bar = smem_u32(&full_bar[stage]);
      // lanes whose piece lies beyond a narrow chunk (or all lanes in the gather-off ablation) still take part in the shuffles
      const bool act = piece * 8 < width && !(dbg & 1);
      // source address = per-lane 64-bit base + row * row_bytes as ONE 32x32+64 multiply-add (IMAD.WIDE.U32)
      uint64_t src0 = (uint64_t)p.in + (uint64_t)(ch * kChunkElems + piece * 8) * sizeof(__nv_bfloat16);
      asm volatile("" : "+l"(src0));  // keep the base in one register pair (otherwise it is re-derived from the constant bank per copy)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if ((ci.has >> g) & 1u) {  // warp-uniform
          const uint32_t a_base = smem_u32(sA + (size_t)stage * a_bytes + (size_t)g * kABytes);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int idx = __shfl_sync(0xffffffffu, ci.r[g][q], 4 * jj + rsub);
              const int row = q * 32 + 4 * jj + rsub;  // row within the tile
              const uint32_t dst = a_base + (uint32_t)(row * 128) + ((jj & 1) ? sw_odd : sw_even);
              const uint32_t idxc = idx < 0 ? 0u : (uint32_t)idx;
              if (act) cp_async_16(dst, (const void*)(src0 + (uint64_t)idxc * row_bytes), idx < 0 ? 0u : 16u);
            }
          }
        }
      }
      // completion-triggered arrive: no wait_group / fence in the producer; the proxy hand-over is completed by the MMA
      // thread's acquire fence after it has observed the barrier phase
      cp_async_arrive_noinc(bar);
      if (elect_one_sync()) {
        if (dbg & 2) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, (uint32_t)b_bytes);
          const __nv_bfloat16* wsrc = p.wimg + ((int64_t)ci.kw * nchunks + ch) * (int64_t)p.cout * kChunkElems;
          bulk_copy_g2s(smem_u32(sB + (size_t)stage * b_bytes), wsrc, (uint32_t)b_bytes, bar);
        }
      }
      __syncwarp();
      stage += NP;
      while (stage >= S) {
        stage -= S;
        phase ^= 1;
      }
    };
    // Three index sets rotate WITHOUT register copies (the loop is unrolled by three): a `ci = ni; ni = nni;` rotation makes
    // the compiler move the just-loaded index registers at the end of the same iteration, i.e. wait for the table loads it
    // was meant to hide for two stages (1x1 and 2^3 convs, whose stages are short, gained ~25 %).
    SuperCursor la(p.tile_kmask, G, cta_first, cta_stride, p.num_tiles, nchunks, p.chunk_major, cta_part, ksplit);  // lookahead cursor
    if (active) la.advance(pi);
    Idx i0 = load_idx(la);
    la.advance(NP);
    Idx i1 = load_idx(la);
    la.advance(NP);
    Idx i2;
    while (active) {
      if (!i0.valid) break;
      i2 = load_idx(la);
      la.advance(NP);
      stage_body(i0);
      if (!i1.valid) break;
      i0 = load_idx(la);
      la.advance(NP);
      stage_body(i1);
      if (!i2.valid) break;
      i1 = load_idx(la);
      la.advance(NP);
      stage_body(i2);
    }
    if (kDbg && p.trace && warp == 0 && lane == 0) {
      p.trace[blockIdx.x * 16 + 0] = (unsigned long long)tp_wait;
      p.trace[blockIdx.x * 16 + 1] = (unsigned long long)(clock64() - tp_start);
    }
    cp_async_wait<0>();
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    // Every lane runs the loop so that all control flow and descriptor arithmetic stays warp-uniform; only the tcgen05
    // instructions sit under elect_one_sync() (see umma.cuh).
    Ring ring(S);
    const uint32_t idesc = make_idesc(kTileM, p.cout);
    int acc_buf = 0;
    uint32_t acc_phase = 0;
    bool tile_open = false;
    uint32_t started = 0;  // bit g: accumulator of member tile g holds a partial sum
    long long tw_full = 0, tw_tmem = 0, t_start = kDbg ? clock64() : 0, n_it = 0;
    for (SuperCursor cur(p.tile_kmask, G, cta_first, cta_stride, p.num_tiles, nchunks, p.chunk_major, cta_part, ksplit); cur.valid(); cur.step()) {
      if (!tile_open) {
        const long long t0 = kDbg ? clock64() : 0;
        mbar_wait(smem_u32(&tmem_empty[acc_buf]), acc_phase ^ 1);
        if (kDbg) tw_tmem += clock64() - t0;
        tc_fence_after();
        tile_open = true;
        started = 0;
      }
      const uint32_t has = (cur.has(0) ? 1u : 0u) | (cur.has(1) ? 2u : 0u);
      {
        const int ch = cur.ch;
        const int width = min(kChunkElems, p.cin - ch * kChunkElems);
        const long long t1 = kDbg ? clock64() : 0;
        mbar_wait(smem_u32(&full_bar[ring.stage]), ring.phase);
        if (kDbg) {
          tw_full += clock64() - t1;
          ++n_it;
        }
        if (dbg & 64) fence_proxy_async();  // A/B only: the bidirectional fence (MEMBAR.ALL.CTA) the acquire form replaces
        else fence_proxy_async_acquire();   // producers' cp.async (generic proxy) writes -> visible to tcgen05.mma (async proxy)
        tc_fence_after();
        const uint32_t b_lo = smem_desc_lo(smem_u32(sB) + (uint32_t)(ring.stage * b_bytes));
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA) + (uint32_t)(ring.stage * a_bytes));
        const int nk = width >> 4;
        if (elect_one_sync()) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if ((has >> g) & 1u) {
              const uint32_t a_lo = a_lo0 + (uint32_t)(g * (kABytes >> 4));
              const uint32_t tmem_d = tmem_base + (uint32_t)((acc_buf * G + g) * acc_cols);
              const uint32_t first_acc = (started >> g) & 1u;
              const int nkg = (dbg & 4) ? (first_acc ? 0 : 1) : nk;
              for (int kk = 0; kk < nkg; ++kk) {
                // +32 bytes along K inside the 128-byte swizzle row == +2 in the (addr >> 4) field
                umma_bf16_lohi(tmem_d, a_lo + 2u * kk, b_lo + 2u * kk, kSmemDescHi, idesc, (kk > 0) ? 1u : first_acc);
              }
            }
          }
          umma_commit(smem_u32(&empty_bar[ring.stage]));  // frees the stage when these MMAs retire
        }
        __syncwarp();
        started |= has;
        ring.advance();
      }
      if (cur.last_iter()) {
        if (elect_one_sync()) umma_commit(smem_u32(&tmem_full[acc_buf]));
        __syncwarp();
        tile_open = false;
        if (++acc_buf == 2) {
          acc_buf = 0;
          acc_phase ^= 1;
        }
      }
    }
    if (kDbg && p.trace && lane == 0) {
      p.trace[blockIdx.x * 16 + 2] = (unsigned long long)tw_full;
      p.trace[blockIdx.x * 16 + 3] = (unsigned long long)tw_tmem;
      p.trace[blockIdx.x * 16 + 4] = (unsigned long long)(clock64() - t_start);
      p.trace[blockIdx.x * 16 + 5] = (unsigned long long)n_it;
    }
  } else {
    // ===================== epilogue (4 warps) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
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
        mbar_wait(smem_u32(&tmem_full[acc_buf]), acc_phase);
        if (kDbg) te_wait += clock64() - t0;
        tc_fence_after();
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g >= G || st * G + g >= p.num_tiles) continue;
        TOut* optr = orow[g] >= 0 ? (TOut*)p.out + orow[g] * p.out_stride : nullptr;
        const uint32_t tcol = (uint32_t)((acc_buf * G + g) * acc_cols);
        if (p.stage_out) {
          // Row pitch of e.g. 25 fp32 (mask logits): per-thread row stores would touch 32 sectors per instruction.  The
          // warp's 32 rows are one contiguous run of 32 * out_cols elements in global memory, so transpose through shared
          // memory (odd pitch => conflict free) and store the run with fully coalesced 128-byte instructions.
          const uint32_t stg = smem_u32(out_stage) + (uint32_t)(q * 32 * p.out_cols) * 4u;  // this warp's staging block
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
        for (int c0 = 0; c0 < p.out_cols; c0 += 16) {
          float v[16];
          if (km[g] != 0u) {
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
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
            if (vec_ok && c0 + 16 <= p.out_cols) {
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
    if (kDbg && p.trace && warp == kFirstEpiWarp && lane == 0) {
      p.trace[blockIdx.x * 16 + 6] = (unsigned long long)te_wait;
      p.trace[blockIdx.x * 16 + 7] = (unsigned long long)(clock64() - te_start);
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

inline size_t conv_smem_bytes(int cout, int stages, int group = 1, int stage_out_cols = 0) {
  return 1024 + (size_t)stages * ((size_t)group * kABytes + (size_t)cout * 128) + (2 * kMaxStages + 6) * sizeof(uint64_t) + 16 +
         (size_t)4 * 32 * stage_out_cols * sizeof(float);
}

int conv_tma_launch(const ConvParams& p, int64_t n_in, int out_dtype, cudaStream_t st);  // conv_tma.cu

// 0 = LSU producer: LDG.128 -> registers -> swizzled STS.128 (default); 1 = TMA gather4 producer (MK3D_GATHER=tma; measured
// ~2.5x slower on random rows, kept for A/B measurements)
static int g_gather_mode = -1;
inline int gather_mode() {
  if (g_gather_mode < 0) {
    const char* e = getenv("MK3D_GATHER");
    g_gather_mode = (e && e[0] == 't') ? 1 : 0;
  }
  return g_gather_mode;
}

}  // namespace mk

using namespace mk;

extern "C" {

int mk_set_gather_mode(int mode) {
  g_gather_mode = mode ? 1 : 0;
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

int mk_conv_fwd_bf16(const void* in, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K, int64_t n_slots,
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
  const int elem = out_dtype == MK_F32 ? 4 : 2;
  p.stage_out = (out_rows == nullptr && out_stride == out_cols && ((out_stride * elem) & 15) != 0 && out_cols <= 64) ? 1 : 0;
  const int soc = p.stage_out ? out_cols : 0;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("MK3D_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;  // timing ablations only (results are wrong when non-zero): 1 skip gathers, 2 skip weight copies, 4 one MMA per tile, 8 skip stores
  }
  p.trace = nullptr;
  {
    // chunk-major iteration order inside a super-tile (all offsets of channel chunk 0, then of chunk 1): the rows a compact
    // tile re-gathers across its offsets are then re-touched within ~10 iterations on HALF the bytes, which fits the L1 left
    // next to the pipeline stages (L1 sector hit rate 44 % -> 60 %, L2 -> L1 traffic 1.90 -> 1.57 GB at 96 channels).
    // MK3D_CHUNK_MAJOR=0 restores (offset, chunk) order for A/B runs.
    static int cm = -1;
    if (cm < 0) { const char* e = getenv("MK3D_CHUNK_MAJOR"); cm = e ? atoi(e) : 1; }
    p.chunk_major = cm;
  }
  if (gather_mode() == 1) return conv_tma_launch(p, n_in, out_dtype, (cudaStream_t)stream_);
  // super-tiles of two tiles when both accumulators fit next to their double buffer (4 x 128 TMEM columns) and there are
  // enough tiles to keep every SM busy with the coarser work unit
  static int grp_env = -1;
  if (grp_env < 0) { const char* e = getenv("MK3D_GROUP"); grp_env = e ? atoi(e) : 0; }
  int group = (cout <= 128 && p.num_tiles >= 4 * num_sms()) ? 2 : 1;
  if (grp_env == 1 || (grp_env == 2 && cout <= 128)) group = grp_env;
  p.group = group;
  // Offset split for the coarse levels: a level with fewer super-tiles than SMs (L3 / L4 of a single scene: 7-27 tiles for 148
  // SMs) ran as 7-27 CTAs walking all 27 offsets each, pure latency.  With an fp32 output the offsets of a super-tile are dealt
  // to `ksplit` CTAs that add their partial sums with vector REDs (the output is zeroed here first).
  p.ksplit = 1;
  {
    static int ks_env = -1;
    if (ks_env < 0) { const char* e = getenv("MK3D_KSPLIT"); ks_env = e ? atoi(e) : 0; }
    const int num_super0 = (p.num_tiles + group - 1) / group;
    if (out_dtype == MK_F32 && K > 1 && !p.stage_out && (out_cols & 3) == 0 && ((out_stride * 4) & 15) == 0 && ks_env != 1 &&
        2 * num_super0 <= num_sms()) {
      int ks = num_sms() / num_super0;
      if (ks > K) ks = K;
      if (ks > 9) ks = 9;
      if (ks_env > 1 && ks > ks_env) ks = ks_env;
      p.ksplit = ks;
    }
  }
  int stages = kMaxStages;
  while (stages > 2 && conv_smem_bytes(cout, stages, group, soc) > 222 * 1024) --stages;
  // Super-tiles: 4 stages (<= 176 KB) instead of the 5 that fit.  The kernel is bound by shared-memory bandwidth (LDGSTS
  // writes + tcgen05 operand reads, ncu: l1tex data-pipe wavefronts ~59 %, tensor pipe ~22 % at 96 channels), not by
  // latency, so the fifth stage buys nothing, while the smaller carve-out leaves ~60 KB of L1 for the neighbour rows a
  // compact tile re-gathers across offsets (L1 sector hit rate 13 % -> 50+ %, 96->96: 486 -> 541 TFLOP/s; 128 unchanged).
  if (group == 2 && stages > 4) stages = 4;
  MK_REQUIRE(conv_smem_bytes(cout, stages, group, soc) <= 227 * 1024, "shared memory budget exceeded");
  {
    static int st_env = -1, np_env = -1;
    if (st_env < 0) { const char* e = getenv("MK3D_STAGES"); st_env = e ? atoi(e) : 0; }
    if (np_env < 0) { const char* e = getenv("MK3D_PRODUCERS"); np_env = e ? atoi(e) : 0; }
    if (st_env >= 2 && st_env < stages) stages = st_env;  // tuning sweeps only
    p.num_producers = stages < kMaxProducers ? stages : kMaxProducers;
    if (np_env >= 1 && np_env < p.num_producers) p.num_producers = np_env;
  }
  p.num_stages = stages;
  const size_t smem = conv_smem_bytes(cout, stages, group, soc);
  int grid = num_sms();
  const int num_super = (p.num_tiles + group - 1) / group;
  if (grid > num_super) grid = num_super;
  if (p.ksplit > 1) {
    grid = num_super * p.ksplit;
    // only the columns this call owns (the output may be a column slice of a wider row, e.g. the sliced dgrad of 384 channels)
    cudaMemset2DAsync(out, (size_t)out_stride * sizeof(float), 0, (size_t)out_cols * sizeof(float), (size_t)n_out, (cudaStream_t)stream_);
  }
  cudaStream_t st = (cudaStream_t)stream_;
  static int carve_env = -2;
  if (carve_env == -2) { const char* e = getenv("MK3D_CARVEOUT"); carve_env = e ? atoi(e) : -1; }
#define MK_LAUNCH(T, D)                                                                                     \
  do {                                                                                                      \
    static size_t smem_set_##D = 0; /* per instantiation: raise the limit only when a launch needs more */  \
    if (smem > smem_set_##D) {                                                                              \
      cudaFuncSetAttribute(k_conv_umma<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
      smem_set_##D = smem;                                                                                  \
    }                                                                                                       \
    if (carve_env >= 0) cudaFuncSetAttribute(k_conv_umma<T, D>, cudaFuncAttributePreferredSharedMemoryCarveout, carve_env); \
    k_conv_umma<T, D><<<grid, kNumThreads, smem, st>>>(p);                                                  \
  } while (0)
  static int trace_on = -1;
  if (trace_on < 0) trace_on = getenv("MK3D_TRACE") ? 1 : 0;
  if (trace_on) {
    cudaMalloc(&p.trace, (size_t)grid * 16 * sizeof(unsigned long long));
    cudaMemset(p.trace, 0, (size_t)grid * 16 * sizeof(unsigned long long));
  }
  const bool dbg_build = trace_on || p.debug != 0;
  if (out_dtype == MK_F32) {
    if (dbg_build) MK_LAUNCH(float, true);
    else MK_LAUNCH(float, false);
  } else {
    if (dbg_build) MK_LAUNCH(__nv_bfloat16, true);
    else MK_LAUNCH(__nv_bfloat16, false);
  }
  if (trace_on) {
    std::vector<unsigned long long> h((size_t)grid * 16);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), p.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    double a[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int j = 0; j < 16; ++j) a[j] += (double)h[(size_t)b * 16 + j] / grid;
    fprintf(stderr, "[mk3d trace] cin %d cout %d K %d tiles %d stages %d | producer0: wait_empty %.0f / total %.0f | mma: wait_full %.0f wait_tmem %.0f / total %.0f, iters %.0f | epilogue: wait_full %.0f / total %.0f (cycles, CTA mean)\n",
            cin, cout, K, p.num_tiles, stages, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
  }
#undef MK_LAUNCH
  return check_launch(__func__);
}

}  // extern "C"
