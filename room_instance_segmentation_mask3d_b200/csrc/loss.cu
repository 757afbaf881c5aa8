// Loss side of the mask logits: Hungarian-matcher cost matrices, the matched mask losses and the
// post-processing mask score, all from ONE pass over each [N, Q] logit tensor.
//
// Reference (paths relative to /root/reference/mask3d):
//   models/matcher.py:17-56      batch_dice_loss / batch_sigmoid_ce_loss  — einsum("nc,mc->nm") over [Q,N] x [T,N]
//   models/matcher.py:120-160    cost matrix C = w_mask * cost_mask + w_class * cost_class + w_dice * cost_dice
//   models/criterion.py:27-72    dice_loss / sigmoid_ce_loss of the matched (query, target) pairs
//   models/criterion.py:168-213  loss_masks
//   trainer/trainer.py:373-402   get_mask_and_scores: sum(sigmoid(x) * (x > 0)) / (sum(x > 0) + 1e-6) per query
//
// The reference converts the bool targets to fp32 [T, N], evaluates two BCE maps and a sigmoid over [Q, N] and runs three
// dense einsums per decoder output and scene (13 outputs per step), then repeats the sigmoid / BCE for the matched pairs.
// Here: with  softplus(-x) - softplus(x) = -x  the BCE sum of query q against target t is
//     sum_n softplus(x[n,q]) - sum_{n in t} x[n,q]
// and the dice terms need sum_{n in t} sigmoid(x[n,q]), sum_n sigmoid(x[n,q]) and |t|.  Targets are per-voxel BIT WORDS
// (bit t of word t/32 = voxel belongs to instance t), so the work per voxel is Q x (instances containing it, normally one)
// instead of Q x T, and the logits are read once: HBM-bound (4*Q + 4*W bytes per voxel).  The matched losses are gathers from
// the same cost matrices (criterion.py computes exactly cost_mask[q_i,t_i] and cost_dice[q_i,t_i] again), so the forward
// of the loss costs nothing more; the backward is one elementwise pass.
// Deterministic: fixed row ranges per block, warp-private shared-memory accumulators, fixed-order fp64 merge.
#include "common.cuh"

namespace mk {

constexpr int kLossWarps = 8;     // warps per block up to 64 queries; 4 above (the warp-private accumulators must fit in shared memory)
constexpr int kLossRows = 1024;   // rows of one scene handled by one block
constexpr int kLossU = 8;                                 // rows per load group (two groups in flight per warp)

// bits[n, w] bit j = masks[(32 w + j), n]   (masks = ItemTargets.instances_masks as bytes, [T, n_scene])
__global__ void __launch_bounds__(256) k_target_bits(const uint8_t* __restrict__ masks, int T, int64_t n_scene, int W,
                                                     uint32_t* __restrict__ bits) {
  const int64_t total = n_scene * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads take consecutive rows of the same word: the T byte streams are read coalesced
    const int w = (int)(i / n_scene);
    const int64_t n = i - (int64_t)w * n_scene;
    uint32_t word = 0;
    const int t1 = min(T, 32 * w + 32);
    for (int t = 32 * w; t < t1; ++t) word |= (masks[(int64_t)t * n_scene + n] != 0 ? 1u : 0u) << (t & 31);
    bits[n * W + w] = word;
  }
}

__device__ __forceinline__ float fast_rcp(float v) {  // MUFU.RCP, 1 ulp; v in [1, 2]
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

struct SigSp {
  float s, sp;
};
// sigmoid(x) and softplus(x) = max(x,0) + log1p(exp(-|x|)) from one exponential
__device__ __forceinline__ SigSp sig_softplus(float x) {
  const float e = __expf(-fabsf(x));
  const float inv = fast_rcp(1.f + e);
  // log1p(e), e in (0, 1]: three series terms below 2^-6 (relative error < 1e-6), MUFU log above (absolute error < 4e-7 on
  // values > 0.0155)
  const float l = e < 0.015625f ? e * (1.f - e * (0.5f - e * (1.f / 3.f))) : __logf(1.f + e);
  SigSp r;
  r.s = x >= 0.f ? inv : e * inv;
  r.sp = fmaxf(x, 0.f) + l;
  return r;
}

template <int QC>
__device__ __forceinline__ void flush_run(float2* acc_w, uint32_t cur, const float (&ax)[QC], const float (&as)[QC], int lane) {
  while (cur) {
    const int t = __ffs(cur) - 1;
    cur &= cur - 1;
#pragma unroll
    for (int j = 0; j < QC; ++j) {
      float2* p = acc_w + t * (QC * 32) + j * 32 + lane;
      float2 v = *p;
      v.x += ax[j];
      v.y += as[j];
      *p = v;
    }
  }
}

// A group of kLossU consecutive rows held by a warp, lane = query (QC chunks of 32): logits, target word, row flags.
template <int QC>
struct RowGroup {
  float x[kLossU][QC];
  uint32_t word[kLossU];
  uint32_t keep;  // bit u: row exists and is selected
};

// px = logits + row * Q + lane, pb = bits + row * W + w, ps = row_select + row (or null) of the group's first row; `rows` of the
// group exist (kLossU for every group but the last).  The full-group path carries no per-row predicates or 64-bit multiplies.
template <int QC>
__device__ __forceinline__ void load_group(RowGroup<QC>& g, const float* __restrict__ px, int Q, const uint32_t* __restrict__ pb, int W,
                                           const uint8_t* __restrict__ ps, int rows, const bool (&qv)[QC]) {
  if (rows == kLossU) {
    g.keep = (1u << kLossU) - 1;
#pragma unroll
    for (int u = 0; u < kLossU; ++u) {
      g.word[u] = pb[u * W];
#pragma unroll
      for (int j = 0; j < QC; ++j) g.x[u][j] = qv[j] ? px[u * Q + j * 32] : 0.f;
    }
    if (ps != nullptr) {
      g.keep = 0;
#pragma unroll
      for (int u = 0; u < kLossU; ++u) g.keep |= (ps[u] != 0 ? 1u : 0u) << u;
    }
  } else {
    g.keep = 0;
#pragma unroll
    for (int u = 0; u < kLossU; ++u) {
      const bool in = u < rows;
      g.word[u] = in ? pb[u * W] : 0u;
      if (in && (ps == nullptr || ps[u] != 0)) g.keep |= 1u << u;
#pragma unroll
      for (int j = 0; j < QC; ++j) g.x[u][j] = (in && qv[j]) ? px[u * Q + j * 32] : 0.f;
    }
  }
}

// Row layout of the per-(scene, range) partial sums, all fp32 (counts <= 1024 are exact):
//   [0, 64 W Q)                (sum_{n in t} x, sum_{n in t} sig) per (word, bit, query)
//   [64 W Q, +4 Q)             (softplus sum, sig sum, sum sig [x>0], sum [x>0]) per query        (written by word 0)
//   [.., +32 W)                |t| per (word, bit)
//   [.., +1)                   rows that took part                                                 (written by word 0)
__host__ __device__ inline int loss_off_q(int Q, int W) { return 64 * W * Q; }
__host__ __device__ inline int loss_off_t(int Q, int W) { return 64 * W * Q + 4 * Q; }
__host__ __device__ inline int loss_off_n(int Q, int W) { return 64 * W * Q + 4 * Q + 32 * W; }
__host__ __device__ inline int loss_row_len(int Q, int W) { return (loss_off_n(Q, W) + 1 + 3) & ~3; }

// grid (ranges, scenes, words).  Block = 8 warps, each walking 128 consecutive rows with lane = query.  A lane keeps the sums
// of the current RUN of rows with identical target word in registers and adds them to the warp's private [32 targets][Q]
// shared-memory accumulator only when the word changes (rows are in lexicographic voxel order, so runs are long).  Two groups
// of 8 rows are in flight per warp (the next group is loaded before the current one is processed).
template <int QC, int WARPS, bool SCORE>
__global__ void __launch_bounds__(WARPS * 32)
    k_mask_stats(const float* __restrict__ logits, int Q, const uint32_t* __restrict__ bits, int W, const uint8_t* __restrict__ sel,
                 const int32_t* __restrict__ scene_offsets, float* __restrict__ part, int R_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int QP = QC * 32;
  float2* acc = reinterpret_cast<float2*>(smem_raw);                       // [warps][32][QP]
  float4* stat = reinterpret_cast<float4*>(acc + WARPS * 32 * QP);    // [warps][QP]
  int32_t* scnt = reinterpret_cast<int32_t*>(stat + WARPS * QP);      // [warps][32] + [warps]

  const int r = blockIdx.x, b = blockIdx.y, w = blockIdx.z;
  const int64_t s0 = scene_offsets[b], s1 = scene_offsets[b + 1];
  const int64_t row0 = s0 + (int64_t)r * kLossRows;
  if (row0 >= s1) return;
  const int64_t row1 = min(s1, row0 + kLossRows);
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* acc_w = acc + wi * 32 * QP;
  for (int i = lane; i < 32 * QP; i += 32) acc_w[i] = make_float2(0.f, 0.f);

  float ax[QC], as[QC], G[QC], S[QC], HS[QC], HC[QC];
  bool qv[QC];
#pragma unroll
  for (int j = 0; j < QC; ++j) {
    ax[j] = as[j] = G[j] = S[j] = HS[j] = HC[j] = 0.f;
    qv[j] = j * 32 + lane < Q;
  }
  uint32_t cur = 0;
  int cnt = 0, nsel = 0;
  const int64_t a0 = row0 + (int64_t)wi * (kLossRows / WARPS);
  const int nrows = (int)max((int64_t)0, min(row1 - a0, (int64_t)(kLossRows / WARPS)));
  const float* px = logits + a0 * Q + lane;
  const uint32_t* pb = bits + a0 * W + w;
  const uint8_t* ps = sel != nullptr ? sel + a0 : nullptr;
  const int stepx = kLossU * Q, stepb = kLossU * W;
  RowGroup<QC> g, nx;
  if (nrows > 0) load_group<QC>(g, px, Q, pb, W, ps, min(nrows, kLossU), qv);
  for (int done = 0; done < nrows; done += kLossU) {
    px += stepx, pb += stepb;
    if (ps != nullptr) ps += kLossU;
    const int left = nrows - done - kLossU;
    if (left > 0) load_group<QC>(nx, px, Q, pb, W, ps, min(left, kLossU), qv);
    uint32_t diff = g.keep ^ ((1u << kLossU) - 1);
#pragma unroll
    for (int u = 0; u < kLossU; ++u) diff |= g.word[u] ^ cur;
    if (diff == 0) {  // the whole group continues the current run (warp-uniform): straight-line accumulation
      nsel += kLossU;
      cnt += kLossU * ((cur >> lane) & 1u);
#pragma unroll
      for (int u = 0; u < kLossU; ++u) {
#pragma unroll
        for (int j = 0; j < QC; ++j) {
          const float xv = g.x[u][j];
          const SigSp v = sig_softplus(xv);
          G[j] += v.sp;
          S[j] += v.s;
          if (SCORE) {
            const bool pos = xv > 0.f;
            HS[j] += pos ? v.s : 0.f;
            HC[j] += pos ? 1.f : 0.f;
          }
          ax[j] += xv;
          as[j] += v.s;
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < kLossU; ++u) {
        if (!((g.keep >> u) & 1u)) continue;  // warp-uniform
        ++nsel;
        const uint32_t wd = g.word[u];
        cnt += (wd >> lane) & 1u;
        if (wd != cur) {
          flush_run<QC>(acc_w, cur, ax, as, lane);
          cur = wd;
#pragma unroll
          for (int j = 0; j < QC; ++j) ax[j] = as[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < QC; ++j) {
          const float xv = g.x[u][j];
          const SigSp v = sig_softplus(xv);
          G[j] += v.sp;
          S[j] += v.s;
          if (SCORE) {
            const bool pos = xv > 0.f;
            HS[j] += pos ? v.s : 0.f;
            HC[j] += pos ? 1.f : 0.f;
          }
          ax[j] += xv;
          as[j] += v.s;
        }
      }
    }
    g = nx;
  }
  flush_run<QC>(acc_w, cur, ax, as, lane);
#pragma unroll
  for (int j = 0; j < QC; ++j) stat[wi * QP + j * 32 + lane] = make_float4(G[j], S[j], HS[j], HC[j]);
  scnt[wi * 32 + lane] = cnt;
  if (lane == 0) scnt[WARPS * 32 + wi] = nsel;
  __syncthreads();

  float* prow = part + ((int64_t)b * R_max + r) * loss_row_len(Q, W);
  float2* pq = reinterpret_cast<float2*>(prow) + (int64_t)w * 32 * Q;
  for (int i = threadIdx.x; i < 32 * Q; i += blockDim.x) {
    const int t = i / Q, q = i - t * Q;
    float2 v = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < WARPS; ++k) {
      const float2 a = acc[(k * 32 + t) * QP + q];
      v.x += a.x;
      v.y += a.y;
    }
    pq[i] = v;
  }
  if (threadIdx.x < 32) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < WARPS; ++k) c += scnt[k * 32 + threadIdx.x];
    prow[loss_off_t(Q, W) + w * 32 + threadIdx.x] = (float)c;
  }
  if (w == 0) {
    float4* p4 = reinterpret_cast<float4*>(prow + loss_off_q(Q, W));
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < WARPS; ++k) {
        const float4 a = stat[k * QP + q];
        v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
      }
      p4[q] = v;
    }
    if (threadIdx.x == 0) {
      int c = 0;
      for (int k = 0; k < WARPS; ++k) c += scnt[WARPS * 32 + k];
      prow[loss_off_n(Q, W)] = (float)c;
    }
  }
}

// sums[b, c] = sum over the scene's ranges of part[b, r, c] in fp64: block = 32 columns x 32 range lanes, every lane adds its
// ranges in ascending order, then a fixed-order tree over the lanes (deterministic).
__global__ void __launch_bounds__(1024) k_loss_merge(const float* __restrict__ part, const int32_t* __restrict__ scene_offsets, int L,
                                                     int R_max, double* __restrict__ sums) {
  __shared__ double sm[32][33];
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t nb = scene_offsets[b + 1] - scene_offsets[b];
  const int R = (int)((nb + kLossRows - 1) / kLossRows);
  double a = 0.0;
  if (c < L)
    for (int r = threadIdx.y; r < R; r += 32) a += (double)part[((int64_t)b * R_max + r) * L + c];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  for (int h = 16; h > 0; h >>= 1) {
    if (threadIdx.y < h) sm[threadIdx.y][threadIdx.x] += sm[threadIdx.y + h][threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.y == 0 && c < L) sums[(int64_t)b * L + c] = sm[0][threadIdx.x];
}

// One thread per (scene, query, target):
//   cost_mask = (sum_n softplus(x) - sum_{n in t} x) / n            matcher.py:39-56 (== sigmoid_ce_loss of the pair)
//   cost_dice = 1 - (2 sum_{n in t} sig + 1) / (sum_n sig + |t| + 1)  matcher.py:17-33 (== dice_loss of the pair)
// rows written by word w > 0 blocks only hold their own columns; the word-0 columns (per-query sums, n) come from word 0.
__global__ void __launch_bounds__(128)
    k_mask_costs_final(const double* __restrict__ sums, int B, int Q, int W, float* __restrict__ cost_mask, float* __restrict__ cost_dice,
                       float* __restrict__ dice_den, float* __restrict__ mask_score, float* __restrict__ n_rows) {
  const int Tp = 32 * W;
  const int64_t total = (int64_t)B * Q * Tp;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int t = (int)(i % Tp);
  const int q = (int)((i / Tp) % Q);
  const int b = (int)(i / ((int64_t)Tp * Q));
  const double* s = sums + (int64_t)b * loss_row_len(Q, W);
  const double sx = s[2 * (t * Q + q)], ss = s[2 * (t * Q + q) + 1];
  const double* sq = s + loss_off_q(Q, W) + 4 * q;
  const double G = sq[0], S = sq[1], HS = sq[2], HC = sq[3];
  const double cnt = s[loss_off_t(Q, W) + t], n = s[loss_off_n(Q, W)];
  const double den = S + cnt + 1.0;
  cost_mask[i] = n > 0 ? (float)((G - sx) / n) : 0.f;
  cost_dice[i] = (float)(1.0 - (2.0 * ss + 1.0) / den);
  dice_den[i] = (float)den;
  if (t == 0) {
    if (mask_score != nullptr) mask_score[b * Q + q] = (float)(HS / (HC + 1e-6));
    if (q == 0) n_rows[b] = (float)n;
  }
}

// d(sum_i coef-weighted losses)/dx: for the query's matched target t (match[b,q], -1 = unmatched)
//   grad = a (sig - tgt) + sig (1 - sig) (c1 - tgt c2),   (a, c1, c2) = coef[b,q,:]
// with a = g_mask / (M n), c1 = g_dice num / (M den^2), c2 = 2 g_dice / (M den)   (criterion.py:27-72 differentiated).
// Same shape as k_mask_stats: lane = query, so the query's target and coefficients live in registers for the whole scene and
// the target word of a row is one broadcast load.
template <int QC>
__global__ void __launch_bounds__(kLossWarps * 32)
    k_mask_loss_bwd(const float* __restrict__ logits, int Q, const uint32_t* __restrict__ bits, int W, const uint8_t* __restrict__ sel,
                    const int32_t* __restrict__ scene_offsets, const int32_t* __restrict__ match, const float* __restrict__ coef,
                    float* __restrict__ grad) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int64_t s0 = scene_offsets[b], s1 = scene_offsets[b + 1];
  const int64_t row0 = s0 + (int64_t)r * kLossRows;
  if (row0 >= s1) return;
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t a0 = row0 + (int64_t)wi * (kLossRows / kLossWarps);
  const int nrows = (int)max((int64_t)0, min(min(s1, row0 + kLossRows) - a0, (int64_t)(kLossRows / kLossWarps)));
  int tw[QC], tb[QC];
  float ca[QC], c1[QC], c2[QC];
  bool qv[QC];
#pragma unroll
  for (int j = 0; j < QC; ++j) {
    const int q = j * 32 + lane;
    qv[j] = q < Q;
    const int t = qv[j] ? match[b * Q + q] : -1;
    tw[j] = t >= 0 ? (t >> 5) : 0;
    tb[j] = t & 31;
    const float* c = coef + (int64_t)(b * Q + (qv[j] ? q : 0)) * 3;
    const bool live = t >= 0;  // unmatched query: all coefficients 0 -> gradient 0
    ca[j] = live ? c[0] : 0.f, c1[j] = live ? c[1] : 0.f, c2[j] = live ? c[2] : 0.f;
  }
  const float* px = logits + a0 * Q + lane;
  float* pg = grad + a0 * Q + lane;
  const uint32_t* pb = bits + a0 * W;
  const uint8_t* ps = sel != nullptr ? sel + a0 : nullptr;
  for (int done = 0; done < nrows; done += kLossU) {
    const int rows = min(kLossU, nrows - done);
    float x[kLossU][QC];
    uint32_t wd[kLossU][QC];
    uint32_t keep = 0;
    if (rows == kLossU) {
#pragma unroll
      for (int u = 0; u < kLossU; ++u) {
#pragma unroll
        for (int j = 0; j < QC; ++j) {
          x[u][j] = qv[j] ? px[u * Q + j * 32] : 0.f;
          wd[u][j] = pb[u * W + tw[j]];
        }
      }
      keep = (1u << kLossU) - 1;
      if (ps != nullptr) {
        keep = 0;
#pragma unroll
        for (int u = 0; u < kLossU; ++u) keep |= (ps[u] != 0 ? 1u : 0u) << u;
      }
    } else {
#pragma unroll
      for (int u = 0; u < kLossU; ++u) {
        const bool in = u < rows;
        if (in && (ps == nullptr || ps[u] != 0)) keep |= 1u << u;
#pragma unroll
        for (int j = 0; j < QC; ++j) {
          x[u][j] = (in && qv[j]) ? px[u * Q + j * 32] : 0.f;
          wd[u][j] = in ? pb[u * W + tw[j]] : 0u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kLossU; ++u) {
      if (u >= rows) break;
      const float k = (keep >> u) & 1u ? 1.f : 0.f;
#pragma unroll
      for (int j = 0; j < QC; ++j) {
        const float e = __expf(-fabsf(x[u][j]));
        const float inv = fast_rcp(1.f + e);
        const float s = x[u][j] >= 0.f ? inv : e * inv;
        const float tgt = (float)((wd[u][j] >> tb[j]) & 1u);
        const float gv = k * (ca[j] * (s - tgt) + s * (1.f - s) * (c1[j] - tgt * c2[j]));
        if (qv[j]) pg[u * Q + j * 32] = gv;
      }
    }
    px += kLossU * Q, pg += kLossU * Q, pb += kLossU * W;
    if (ps != nullptr) ps += kLossU;
  }
}

inline int64_t loss_ranges(int64_t max_scene_rows) { return max_scene_rows > 0 ? ceil_div(max_scene_rows, kLossRows) : 1; }

template <int QC, bool SCORE>
int launch_mask_stats(dim3 grid, cudaStream_t st, const float* logits, int Q, const uint32_t* bits, int W, const uint8_t* sel,
                      const int32_t* scene_offsets, float* part, int R_max) {
  constexpr int QP = QC * 32;
  constexpr int WARPS = QC <= 2 ? kLossWarps : 4;
  const size_t smem = (size_t)WARPS * 32 * QP * sizeof(float2) + (size_t)WARPS * QP * sizeof(float4) +
                      (size_t)(WARPS * 32 + WARPS) * sizeof(int32_t);
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    if (cudaFuncSetAttribute(k_mask_stats<QC, WARPS, SCORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail("mk_mask_costs", "cannot reserve shared memory");
    attr_set = true;
  }
  k_mask_stats<QC, WARPS, SCORE><<<grid, WARPS * 32, smem, st>>>(logits, Q, bits, W, sel, scene_offsets, part, R_max);
  return 0;
}

}  // namespace mk

using namespace mk;

extern "C" {

int mk_target_bits(const uint8_t* masks, int T, int64_t n_scene, int W, uint32_t* bits, void* stream_) {
  MK_REQUIRE(T >= 0 && W >= 1 && T <= 32 * W, "T must fit in W words");
  if (n_scene == 0) return 0;
  int64_t blocks = ceil_div(n_scene * W, 256);
  const int64_t cap = 8 * (int64_t)num_sms();
  k_target_bits<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream_>>>(masks, T, n_scene, W, bits);
  return check_launch(__func__);
}

int64_t mk_mask_costs_workspace_bytes(int B, int64_t max_scene_rows, int Q, int W) {
  const int64_t L = loss_row_len(Q, W);
  return (int64_t)B * loss_ranges(max_scene_rows) * L * 4 + (int64_t)B * L * 8 + 64;
}

int mk_mask_costs(const float* logits, int64_t n, int Q, const uint32_t* bits, int W, const uint8_t* row_select,
                  const int32_t* scene_offsets, int B, int64_t max_scene_rows, float* cost_mask, float* cost_dice, float* dice_den,
                  float* mask_score, float* n_rows, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(Q >= 1 && Q <= 128, "1 <= Q <= 128");
  MK_REQUIRE(W >= 1 && W <= 64, "1 <= W <= 64 target words");
  MK_REQUIRE(B >= 1 && B <= 65535, "1 <= B <= 65535");
  MK_REQUIRE(max_scene_rows >= 0 && max_scene_rows <= n, "max_scene_rows out of range");
  MK_REQUIRE(workspace_bytes >= mk_mask_costs_workspace_bytes(B, max_scene_rows, Q, W), "workspace too small");
  const int R_max = (int)loss_ranges(max_scene_rows);
  const int L = loss_row_len(Q, W);
  double* sums = (double*)workspace;  // [B, L]
  float* part = (float*)((unsigned char*)workspace + (((int64_t)B * L * 8 + 15) & ~15ll));  // [B, R_max, L]
  if (n > 0 && max_scene_rows > 0) {
    dim3 grid((unsigned)R_max, (unsigned)B, (unsigned)W);
    int rc = 0;
#define MK_STATS(QC)                                                                                                       \
  (mask_score != nullptr ? launch_mask_stats<QC, true>(grid, st, logits, Q, bits, W, row_select, scene_offsets, part, R_max) \
                         : launch_mask_stats<QC, false>(grid, st, logits, Q, bits, W, row_select, scene_offsets, part, R_max))
    switch ((Q + 31) / 32) {
      case 1: rc = MK_STATS(1); break;
      case 2: rc = MK_STATS(2); break;
      case 3: rc = MK_STATS(3); break;
      default: rc = MK_STATS(4); break;
    }
#undef MK_STATS
    if (rc) return rc;
    if (check_launch(__func__)) return 1;
  }
  k_loss_merge<<<dim3((unsigned)ceil_div(L, 32), (unsigned)B), dim3(32, 32), 0, st>>>(part, scene_offsets, L, R_max, sums);
  const int64_t total = (int64_t)B * Q * 32 * W;
  k_mask_costs_final<<<(unsigned)ceil_div(total, 128), 128, 0, st>>>(sums, B, Q, W, cost_mask, cost_dice, dice_den, mask_score, n_rows);
  return check_launch(__func__);
}

int mk_mask_loss_bwd(const float* logits, int64_t n, int Q, const uint32_t* bits, int W, const uint8_t* row_select,
                     const int32_t* scene_offsets, int B, int64_t max_scene_rows, const int32_t* match, const float* coef, float* grad,
                     void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(Q >= 1 && Q <= 128 && W >= 1 && B >= 1 && B <= 65535, "bad sizes");
  MK_REQUIRE(max_scene_rows >= 0 && max_scene_rows <= n, "max_scene_rows out of range");
  if (n == 0 || max_scene_rows == 0) return 0;
  dim3 grid((unsigned)loss_ranges(max_scene_rows), (unsigned)B);
  switch ((Q + 31) / 32) {
    case 1: k_mask_loss_bwd<1><<<grid, kLossWarps * 32, 0, st>>>(logits, Q, bits, W, row_select, scene_offsets, match, coef, grad); break;
    case 2: k_mask_loss_bwd<2><<<grid, kLossWarps * 32, 0, st>>>(logits, Q, bits, W, row_select, scene_offsets, match, coef, grad); break;
    case 3: k_mask_loss_bwd<3><<<grid, kLossWarps * 32, 0, st>>>(logits, Q, bits, W, row_select, scene_offsets, match, coef, grad); break;
    default: k_mask_loss_bwd<4><<<grid, kLossWarps * 32, 0, st>>>(logits, Q, bits, W, row_select, scene_offsets, match, coef, grad); break;
  }
  return check_launch(__func__);
}

}  // extern "C"
