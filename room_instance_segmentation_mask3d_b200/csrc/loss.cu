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

constexpr int kLossRows = 2048;  // rows of one scene handled by one block
constexpr int kLossWarps = 4;
constexpr int kLossRowsPerWarp = kLossRows / kLossWarps;

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

struct SigSp {
  float s, sp;
};
// sigmoid(x) and softplus(x) = max(x,0) + log1p(exp(-|x|)) from one exponential
__device__ __forceinline__ SigSp sig_softplus(float x) {
  const float e = __expf(-fabsf(x));
  const float inv = __fdividef(1.f, 1.f + e);
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

// grid (ranges, scenes, words).  Block = 4 warps, each walking a contiguous run of rows with lane = query (QC chunks of 32).
// A lane keeps the sums of the current RUN of rows with identical target word in registers and adds them to the warp's
// private [32 targets][Q] shared-memory accumulator only when the word changes (rows are in lexicographic voxel order, so
// runs are long).
template <int QC>
__global__ void __launch_bounds__(kLossWarps * 32)
    k_mask_stats(const float* __restrict__ logits, int Q, const uint32_t* __restrict__ bits, int W, const uint8_t* __restrict__ sel,
                 const int32_t* __restrict__ scene_offsets, float2* __restrict__ part_qt, float4* __restrict__ part_q,
                 int32_t* __restrict__ part_t, int32_t* __restrict__ part_n, int R_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int QP = QC * 32;
  float2* acc = reinterpret_cast<float2*>(smem_raw);                       // [warps][32][QP]
  float4* stat = reinterpret_cast<float4*>(acc + kLossWarps * 32 * QP);    // [warps][QP]
  int32_t* scnt = reinterpret_cast<int32_t*>(stat + kLossWarps * QP);      // [warps][32] + [warps]

  const int r = blockIdx.x, b = blockIdx.y, w = blockIdx.z;
  const int64_t s0 = scene_offsets[b], s1 = scene_offsets[b + 1];
  const int64_t row0 = s0 + (int64_t)r * kLossRows;
  if (row0 >= s1) return;
  const int64_t row1 = min(s1, row0 + kLossRows);
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* acc_w = acc + wi * 32 * QP;
  for (int i = lane; i < 32 * QP; i += 32) acc_w[i] = make_float2(0.f, 0.f);

  float ax[QC], as[QC], G[QC], S[QC], HS[QC], HC[QC];
#pragma unroll
  for (int j = 0; j < QC; ++j) ax[j] = as[j] = G[j] = S[j] = HS[j] = HC[j] = 0.f;
  uint32_t cur = 0;
  int cnt = 0, nsel = 0;
  const int64_t a0 = row0 + (int64_t)wi * kLossRowsPerWarp, a1 = min(row1, a0 + kLossRowsPerWarp);
  constexpr int U = 4;
  for (int64_t base = a0; base < a1; base += U) {
    float x[U][QC];
    uint32_t word[U];
    bool keep[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = base + u;
      const bool in = row < a1;
      word[u] = in ? bits[row * W + w] : 0u;
      keep[u] = in && (sel == nullptr || sel[row] != 0);
#pragma unroll
      for (int j = 0; j < QC; ++j) {
        const int q = j * 32 + lane;
        x[u][j] = (in && q < Q) ? logits[row * Q + q] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!keep[u]) continue;  // warp-uniform
      ++nsel;
      const uint32_t wd = word[u];
      cnt += (wd >> lane) & 1u;
      if (wd != cur) {
        flush_run<QC>(acc_w, cur, ax, as, lane);
        cur = wd;
#pragma unroll
        for (int j = 0; j < QC; ++j) ax[j] = as[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < QC; ++j) {
        const SigSp v = sig_softplus(x[u][j]);
        G[j] += v.sp;
        S[j] += v.s;
        const bool pos = x[u][j] > 0.f;
        HS[j] += pos ? v.s : 0.f;
        HC[j] += pos ? 1.f : 0.f;
        ax[j] += x[u][j];
        as[j] += v.s;
      }
    }
  }
  flush_run<QC>(acc_w, cur, ax, as, lane);
#pragma unroll
  for (int j = 0; j < QC; ++j) stat[wi * QP + j * 32 + lane] = make_float4(G[j], S[j], HS[j], HC[j]);
  scnt[wi * 32 + lane] = cnt;
  if (lane == 0) scnt[kLossWarps * 32 + wi] = nsel;
  __syncthreads();

  const int64_t slot = ((int64_t)b * R_max + r) * W + w;
  float2* pq = part_qt + slot * 32 * Q;
  for (int i = threadIdx.x; i < 32 * Q; i += blockDim.x) {
    const int t = i / Q, q = i - t * Q;
    float2 v = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kLossWarps; ++k) {
      const float2 a = acc[(k * 32 + t) * QP + q];
      v.x += a.x;
      v.y += a.y;
    }
    pq[i] = v;
  }
  if (threadIdx.x < 32) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < kLossWarps; ++k) c += scnt[k * 32 + threadIdx.x];
    part_t[slot * 32 + threadIdx.x] = c;
  }
  if (w == 0) {
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < kLossWarps; ++k) {
        const float4 a = stat[k * QP + q];
        v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
      }
      part_q[((int64_t)b * R_max + r) * Q + q] = v;
    }
    if (threadIdx.x == 0) {
      int c = 0;
      for (int k = 0; k < kLossWarps; ++k) c += scnt[kLossWarps * 32 + k];
      part_n[(int64_t)b * R_max + r] = c;
    }
  }
}

// One thread per (scene, query, target): fixed-order fp64 merge of the range partials, then
//   cost_mask = (sum_n softplus(x) - sum_{n in t} x) / n            matcher.py:39-56 (== sigmoid_ce_loss of the pair)
//   cost_dice = 1 - (2 sum_{n in t} sig + 1) / (sum_n sig + |t| + 1)  matcher.py:17-33 (== dice_loss of the pair)
__global__ void __launch_bounds__(128)
    k_mask_costs_final(const float2* __restrict__ part_qt, const float4* __restrict__ part_q, const int32_t* __restrict__ part_t,
                       const int32_t* __restrict__ part_n, const int32_t* __restrict__ scene_offsets, int B, int Q, int W, int R_max,
                       float* __restrict__ cost_mask, float* __restrict__ cost_dice, float* __restrict__ dice_den,
                       float* __restrict__ mask_score, float* __restrict__ n_rows) {
  const int Tp = 32 * W;
  const int64_t total = (int64_t)B * Q * Tp;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int t = (int)(i % Tp);
  const int q = (int)((i / Tp) % Q);
  const int b = (int)(i / ((int64_t)Tp * Q));
  const int64_t nb = scene_offsets[b + 1] - scene_offsets[b];
  const int R = (int)((nb + kLossRows - 1) / kLossRows);
  const int w = t >> 5, tb = t & 31;
  double sx = 0, ss = 0, G = 0, S = 0, HS = 0, HC = 0, cnt = 0, n = 0;
  for (int r = 0; r < R; ++r) {
    const int64_t rr = (int64_t)b * R_max + r;
    const float2 p = part_qt[((rr * W + w) * 32 + tb) * Q + q];
    sx += p.x, ss += p.y;
    const float4 g = part_q[rr * Q + q];
    G += g.x, S += g.y, HS += g.z, HC += g.w;
    cnt += part_t[(rr * W + w) * 32 + tb];
    n += part_n[rr];
  }
  const double den = S + cnt + 1.0;
  cost_mask[i] = n > 0 ? (float)((G - sx) / n) : 0.f;
  cost_dice[i] = (float)(1.0 - (2.0 * ss + 1.0) / den);
  dice_den[i] = (float)den;
  if (t == 0) {
    mask_score[b * Q + q] = (float)(HS / (HC + 1e-6));
    if (q == 0) n_rows[b] = (float)n;
  }
}

// d(sum_i coef-weighted losses)/dx: for the query's matched target t (match[b,q], -1 = unmatched)
//   grad = a (sig - tgt) + sig (1 - sig) (c1 - tgt c2),   (a, c1, c2) = coef[b,q,:]
// with a = g_mask / (M n), c1 = g_dice num / (M den^2), c2 = 2 g_dice / (M den)   (criterion.py:27-72 differentiated).
template <typename I>
__global__ void __launch_bounds__(256)
    k_mask_loss_bwd(const float* __restrict__ logits, I total, int Q, const uint32_t* __restrict__ bits, int W,
                    const uint8_t* __restrict__ sel, const int32_t* __restrict__ scene_offsets, int B, const int32_t* __restrict__ match,
                    const float* __restrict__ coef, float* __restrict__ grad) {
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const I row = i / (I)Q;
    const int q = (int)(i - row * (I)Q);
    const int b = scene_of_row(scene_offsets, B, (int64_t)row);
    const int t = match[b * Q + q];
    float g = 0.f;
    if (t >= 0 && (sel == nullptr || sel[row] != 0)) {
      const float x = logits[i];
      const float e = __expf(-fabsf(x));
      const float inv = __fdividef(1.f, 1.f + e);
      const float s = x >= 0.f ? inv : e * inv;
      const float tgt = (float)((bits[(int64_t)row * W + (t >> 5)] >> (t & 31)) & 1u);
      const float* c = coef + (b * Q + q) * 3;
      g = c[0] * (s - tgt) + s * (1.f - s) * (c[1] - tgt * c[2]);
    }
    grad[i] = g;
  }
}

inline int64_t loss_ranges(int64_t max_scene_rows) { return max_scene_rows > 0 ? ceil_div(max_scene_rows, kLossRows) : 1; }

template <int QC>
int launch_mask_stats(dim3 grid, cudaStream_t st, const float* logits, int Q, const uint32_t* bits, int W, const uint8_t* sel,
                      const int32_t* scene_offsets, float2* part_qt, float4* part_q, int32_t* part_t, int32_t* part_n, int R_max) {
  constexpr int QP = QC * 32;
  const size_t smem = (size_t)kLossWarps * 32 * QP * sizeof(float2) + (size_t)kLossWarps * QP * sizeof(float4) +
                      (size_t)(kLossWarps * 32 + kLossWarps) * sizeof(int32_t);
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    if (cudaFuncSetAttribute(k_mask_stats<QC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail("mk_mask_costs", "cannot reserve shared memory");
    attr_set = true;
  }
  k_mask_stats<QC><<<grid, kLossWarps * 32, smem, st>>>(logits, Q, bits, W, sel, scene_offsets, part_qt, part_q, part_t, part_n, R_max);
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
  const int64_t slots = (int64_t)B * loss_ranges(max_scene_rows);
  return slots * ((int64_t)W * 32 * Q * 8 + (int64_t)Q * 16 + (int64_t)W * 32 * 4 + 4) + 64;
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
  const int64_t slots = (int64_t)B * R_max;
  unsigned char* ws = (unsigned char*)workspace;
  float4* part_q = (float4*)ws;  // 16-byte aligned first
  ws += slots * Q * 16;
  float2* part_qt = (float2*)ws;
  ws += slots * W * 32 * Q * 8;
  int32_t* part_t = (int32_t*)ws;
  ws += slots * W * 32 * 4;
  int32_t* part_n = (int32_t*)ws;
  if (n > 0 && max_scene_rows > 0) {
    dim3 grid((unsigned)R_max, (unsigned)B, (unsigned)W);
    int rc = 0;
    switch ((Q + 31) / 32) {
      case 1: rc = launch_mask_stats<1>(grid, st, logits, Q, bits, W, row_select, scene_offsets, part_qt, part_q, part_t, part_n, R_max); break;
      case 2: rc = launch_mask_stats<2>(grid, st, logits, Q, bits, W, row_select, scene_offsets, part_qt, part_q, part_t, part_n, R_max); break;
      case 3: rc = launch_mask_stats<3>(grid, st, logits, Q, bits, W, row_select, scene_offsets, part_qt, part_q, part_t, part_n, R_max); break;
      default: rc = launch_mask_stats<4>(grid, st, logits, Q, bits, W, row_select, scene_offsets, part_qt, part_q, part_t, part_n, R_max); break;
    }
    if (rc) return rc;
    if (check_launch(__func__)) return 1;
  }
  const int64_t total = (int64_t)B * Q * 32 * W;
  k_mask_costs_final<<<(unsigned)ceil_div(total, 128), 128, 0, st>>>(part_qt, part_q, part_t, part_n, scene_offsets, B, Q, W, R_max,
                                                                    cost_mask, cost_dice, dice_den, mask_score, n_rows);
  return check_launch(__func__);
}

int mk_mask_loss_bwd(const float* logits, int64_t n, int Q, const uint32_t* bits, int W, const uint8_t* row_select,
                     const int32_t* scene_offsets, int B, const int32_t* match, const float* coef, float* grad, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  MK_REQUIRE(Q >= 1 && W >= 1 && B >= 1, "bad sizes");
  const int64_t total = n * Q;
  if (total == 0) return 0;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = 16 * (int64_t)num_sms();
  const unsigned g = (unsigned)(blocks > cap ? cap : blocks);
  if (total < (1ll << 31) - (int64_t)g * 256)
    k_mask_loss_bwd<int32_t><<<g, 256, 0, st>>>(logits, (int32_t)total, Q, bits, W, row_select, scene_offsets, B, match, coef, grad);
  else
    k_mask_loss_bwd<int64_t><<<g, 256, 0, st>>>(logits, total, Q, bits, W, row_select, scene_offsets, B, match, coef, grad);
  return check_launch(__func__);
}

}  // extern "C"
