/*
 * mk3d.h — C ABI of libmk3d.so: the B200 (sm_100a) sparse-convolution / masked-attention hot path
 * behind the MinkowskiEngine surface that mvg-inatech/room-instance-segmentation-mask3d consumes.
 *
 * The reference has no C-level plugin interface for this path: its lower boundary is MinkowskiEngine
 * 0.5.4's pybind11 module `MinkowskiEngineBackend._C` (un-vendored; pinned at
 * mask3d/requirements_pip_freeze.txt:62).  Each entry point below names the reference call site whose
 * work it replaces (paths relative to /root/reference).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); nothing synchronises;
 *   - no allocation inside: scratch comes from the caller (`*_workspace_bytes` query functions);
 *   - return value: 0 = ok, non-zero = error, message via mk_last_error() (thread-local);
 *   - rows of a batch item are contiguous and batch items ascend (what ME.utils.sparse_collate yields,
 *     mask3d/datasets/utils.py:126); `scene_offsets` are int32 [B+1] row offsets;
 *   - dtype codes: MK_F32 = 0, MK_BF16 = 1.
 */
#ifndef MK3D_H_
#define MK3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MK_F32 0
#define MK_BF16 1
#define MK_TILE_ROWS 128 /* output rows per tensor-core tile */

const char* mk_last_error(void);
int mk_version(void);
/* 1 if the current device is sm_100 (B200); the tcgen05 entry points refuse to run otherwise. */
int mk_device_supported(void);

/* ---------------------------------------------------------------------------------------------
 * (1) Coordinate maps: GPU open-addressing hash table over packed 64-bit (batch,x,y,z) keys.
 * Replaces ME CoordinateMapGPU::insert_and_map / ::stride, reached from
 *   ME.SparseTensor(coordinates=...)            mask3d/datasets/utils.py:249-255
 *   every stride-2 conv / pool (implicit)        mask3d/models/res16unet.py:55-63,248 ; mask3d/models/mask3d.py:297
 *   ME.utils.sparse_quantize                     
 * Semantics: first occurrence wins; output rows in ascending first-occurrence order (ME CPU order).
 * ------------------------------------------------------------------------------------------- */
int64_t mk_hash_capacity(int64_t n);                 /* power of two >= 2n */
int64_t mk_coordmap_workspace_bytes(int64_t n);
/* in_coords int32 [n,4] (batch,x,y,z); stride3_host: 3 ints (1,1,1 = plain insert; else coords are floored to
 * multiples of the stride first).  Outputs: hash (keys uint64[cap], vals int32[cap]) of the NEW map,
 * out_coords int32 [<=n,4], unique_index int32 [<=n] (source row of each new row), inverse int32 [n]
 * (new row of every source row), n_out (device int32), status (device int32, non-zero = coordinate out of the
 * packable range [-2^17, 2^17) or batch >= 1023). */
int mk_coordmap_build(const int32_t* in_coords, int64_t n, const int32_t* stride3_host, uint64_t* keys, int32_t* vals,
                      int64_t cap, int32_t* out_coords, int32_t* unique_index, int32_t* inverse, int32_t* n_out,
                      int32_t* status, void* workspace, int64_t workspace_bytes, void* stream);
/* Voxel dedup in the reference's own order (np.unique(coords, axis=0, return_index=True) per scene followed by
 * ME.utils.sparse_collate, mask3d/datasets/utils.py:96-126): keys[i] = packed (batch,x,y,z) of row i, whose unsigned order is the
 * lexicographic order of the signed rows; the caller sorts (key, row) pairs stably and feeds the sorted rows to
 * mk_coordmap_build, whose first-occurrence order then IS the lexicographic order and whose unique_index picks the
 * lowest original row of every voxel.  status (device int32): non-zero = a coordinate outside the packable range. */
int mk_pack_keys(const int32_t* coords, int64_t n, uint64_t* keys, int32_t* status, void* stream);
/* rows[i] = row of query coordinate i in the map, or -1. */
int mk_hash_query(const uint64_t* keys, const int32_t* vals, int64_t cap, const int32_t* query_coords, int64_t n,
                  int32_t* rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (2) Kernel maps.  Replaces ME CoordinateMapGPU::kernel_map (implicit in every conv/pool; Python-visible as
 * CoordinateManager.kernel_map).  Output-stationary form: nbr[k*n_out + o] = input row of
 * (out_coords[o] + offsets[k]) or -1; offsets int32 [K,3] already scaled by dilation * input tensor stride,
 * enumerated x-fastest.
 * ------------------------------------------------------------------------------------------- */
int mk_kernel_map(const uint64_t* in_keys, const int32_t* in_vals, int64_t cap, const int32_t* out_coords, int64_t n_out,
                  const int32_t* offsets, int K, int32_t* nbr, void* stream);
/* The same table for a map onto ITSELF (stride-1 convolutions with an odd kernel: in map == out map, offsets[K-1-k] ==
 * -offsets[k], which the caller guarantees): nbr[k,o] = i <=> nbr[K-1-k,i] = o, so half the offsets are probed and a hit writes
 * both entries.  Bit-identical to mk_kernel_map. */
int mk_kernel_map_same(const uint64_t* keys, const int32_t* vals, int64_t cap, const int32_t* coords, int64_t n, const int32_t* offsets,
                       int K, int32_t* nbr, void* stream);
/* The 5x5x5 table of a map onto itself composed from its 3x3x3 table `nbr3` [27, n] (same offset step: `step3` = the three
 * host ints dilation * tensor stride): two table reads per entry wherever an intermediate voxel exists, a hash probe otherwise.
 * Bit-identical to mk_kernel_map_same with the 125 x-fastest offsets. */
int mk_kernel_map_compose5(const uint64_t* keys, const int32_t* vals, int64_t cap, const int32_t* coords, int64_t n, const int32_t* step3,
                           const int32_t* nbr3, int32_t* nbr5, void* stream);
/* Per 128-slot tile, bit k set iff some slot of the tile has a neighbour for offset k (K <= 32); also the
 * total pair count (device int64).  `n_slots` is the slot count of nbr (row stride of the table). */
int mk_tile_kmask(const int32_t* nbr, int K, int64_t n_slots, uint32_t* tile_kmask, int64_t* pair_count, void* stream);
/* Transposed-conv plan (coarse -> fine): given the stride-s table nbr_down [K, n_coarse] (fine row or -1) build
 * parent[f] and koff[f] for every fine row f. */
int mk_transpose_map(const int32_t* nbr_down, int K, int64_t n_coarse, int64_t n_fine, int32_t* parent, int32_t* koff,
                     void* stream);
/* Mask-sorted tile plans: keys[o] = (bitmask over k of present neighbours of slot o, K <= 32) << 32 | (batch & 3) << 30 |
 * 30-bit Morton code of the slot's coordinate / stride.  The host sorts slots by key so that every 128-slot tile has (almost)
 * a single mask and is a compact spatial patch, then permutes the table: out[k, s] = nbr[k, order[s]]  (order[s] < 0 -> -1). */
int mk_row_sort_keys(const int32_t* nbr, int K, int64_t n, const int32_t* coords, const int32_t* stride3_host, uint64_t* keys,
                     void* stream);
/* keys[o] = offset bitmask of slot o only (int32).  A STABLE sort by it groups equal masks and keeps each group in row order
 * (the default plan: as fast for the gathers as the Z-curve order, half the sort). */
int mk_row_mask_keys(const int32_t* nbr, int K, int64_t n, int32_t* keys, void* stream);
/* Stable grouping of rows by a 32-bit key with at most 1024 distinct values (the tile-plan sort: rows grouped by their offset
 * mask, groups in ascending key order, rows of a group in their original order -- exactly the permutation a stable sort by key
 * gives).  keys uint32 [n] (0xFFFFFFFF is reserved), order int32 [n] out, status int32 (device, may be NULL): bit 0 set when
 * there were more than 1024 distinct keys (order is then a permutation but not grouped).  Deterministic.  Replaces the
 * thrust / CUB sorts of MinkowskiEngine's kernel-map construction. */
int64_t mk_group_rows_workspace_bytes(int64_t n);
int mk_group_rows(const uint32_t* keys, int64_t n, int32_t* order, int32_t* status, void* workspace, int64_t workspace_bytes, void* stream);
int mk_permute_table(const int32_t* nbr, int K, int64_t n, const int32_t* order, int64_t n_slots, int32_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (3) Sparse convolution.  Replaces ME ConvolutionForwardGPU / ConvolutionTransposeForwardGPU reached from
 * MinkowskiConvolution.forward / MinkowskiConvolutionTranspose.forward (mask3d/models/modules/common.py:151-160,
 * 186-195; call sites mask3d/models/res16unet.py:244-307, mask3d/models/modules/resnet_block.py:52-68).
 *   out[o,:] = sum_k in[nbr[k,o],:] @ W[k]   (+ bias)
 * ------------------------------------------------------------------------------------------- */
/* fp32 CUDA-core path (exact fp32 FMA; the "fp32-accumulate parity mode").  W fp32 [K,cin,cout]. */
int mk_conv_fwd_f32(const float* in, int cin, const float* W, int cout, const int32_t* nbr, int K, int64_t n_out,
                    const float* bias, float* out, void* stream);
/* Weight images for the tensor-core path: W fp32 [K,cin,cout] -> bf16 images, one per (k, 64-channel chunk),
 * each [n_rows x 128 B] in the 128-byte-swizzled K-major shared-memory layout tcgen05.mma reads, so that a
 * stage's B operand is one linear bulk copy.  transpose=0: rows = cout, reduction = cin (forward);
 * transpose=1: rows = cin, reduction = cout (dgrad). */
int64_t mk_conv_weight_image_elems(int K, int cin, int cout, int transpose);
int mk_conv_weight_image_bf16(const float* W, int K, int cin, int cout, int transpose, void* img, void* stream);
/* tcgen05 path: bf16 operands, fp32 accumulation in TMEM.  `in` bf16 [n_in,cin] (cin % 16 == 0), nbr int32
 * [K,n_slots], out_rows int32 [n_slots] or NULL (slot == output row), tile_kmask from mk_tile_kmask,
 * kidx int32 [K] or NULL: weight-image index used for table row k (dgrad passes the mirrored offset),
 * bias fp32 [cout] or NULL, out [n_out,out_stride] of out_dtype of which the first out_cols (<= cout) columns are
 * written (0 = cout; used by the mask-logit product whose Q = 25 queries are padded to 32 accumulator columns).
 * Persistent: grid = #SMs. */
int mk_conv_fwd_bf16(const void* in, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K, int64_t n_slots,
                     const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                     const float* bias, void* out, int out_dtype, int out_cols, int out_stride, void* stream);
/* TF32 tensor-core mode = the reference's own precision (fp32 tensors with torch.set_float32_matmul_precision("high"),
 * mask3d/main_instance_segmentation.py:52): same kernel, fp32 rows gathered as 128-byte chunks of 32 channels,
 * tcgen05.mma.kind::tf32 (K = 8 per instruction), fp32 accumulation.  `in` fp32 [n_in,cin] (cin % 8 == 0), `wimg` from
 * mk_conv_weight_image_tf32 (fp32, 32-channel chunks); everything else as mk_conv_fwd_bf16. */
int64_t mk_conv_weight_image_tf32_elems(int K, int cin, int cout, int transpose);
int mk_conv_weight_image_tf32(const float* W, int K, int cin, int cout, int transpose, void* img, void* stream);
int mk_conv_fwd_tf32(const void* in, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K, int64_t n_slots,
                     const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                     const float* bias, void* out, int out_dtype, int out_cols, int out_stride, void* stream);
/* A-operand producer of mk_conv_fwd_bf16 / _tf32, for A/B measurements (bit-identical results): 0 = LSU (16-byte cp.async into the
 * swizzled tile, default), 1 = the round-1 TMA tile::gather4 kernel (bf16 only), 2 = tile::gather4 producers inside the current
 * kernel.  Both TMA forms measured slower on B200 (profiles/r02_gather_mechanisms.txt). */
int mk_set_gather_mode(int mode);
/* 1: results are bit-identical from run to run (the reference trains with `deterministic: true`,
 * mask3d/conf/trainer/trainer.yaml:1): no offset split with floating-point REDs in mk_conv_fwd_bf16 (forward / dgrad), fixed-order
 * reduction in mk_conv_wgrad_bf16.  0 (default): fastest. */
int mk_set_deterministic(int on);
/* Stem: Cin == 1, K up to 125 (5^3): pure lookup, CUDA cores.  in fp32 [n_in], W fp32 [K,1,cout]. */
int mk_conv_stem_f32(const float* in, const float* W, int cout, const int32_t* nbr, int K, int64_t n_out, void* out,
                     int out_dtype, void* stream);
/* Stem weight gradient (cin == 1, cout <= 32, K <= 128): dW[k,c] = sum_o in[nbr[k,o]] * dout[o,c]; register accumulators per
 * warp, per-block partials in `workspace`, fixed-order merge (deterministic). */
int64_t mk_conv_stem_wgrad_workspace_bytes(int64_t n_out, int K, int cout);
int mk_conv_stem_wgrad_f32(const float* in, const float* dout, int cout, const int32_t* nbr, int K, int64_t n_out, float* dW,
                           void* workspace, int64_t workspace_bytes, void* stream);
/* Backward of the fp32 path (dgrad + wgrad); CUDA cores, deterministic (no atomics on dgrad). */
int mk_conv_dgrad_f32(const float* dout, int cout, const float* W, int cin, const int32_t* nbr_t, int K, int64_t n_in,
                      const int32_t* kidx, float* din, void* stream);
int mk_conv_wgrad_f32(const float* in, int cin, const float* dout, int cout, const int32_t* nbr, int K, int64_t n_out,
                      float* dW, void* stream);
/* Backward of the tensor-core path (ME ConvolutionBackwardGPU: K x (gather, gather, GEMM) launches, reached through the
 * autograd Function of MinkowskiConvolution / MinkowskiConvolutionTranspose).
 *   dgrad = mk_conv_fwd_bf16 over the transposed table with weight images built with transpose=1 (and `kidx` mirroring the
 *           offsets for same-map convs), no separate entry point;
 *   wgrad = dW[k] = sum over slots o of x[nbr[k,o]]^T dy[row(o)], one persistent tcgen05 kernel over the forward's tile plan:
 *           both operands MN-major, fp32 accumulation in TMEM across a run of tiles of one offset, fp32 atomic flush.
 * mk_conv_wgrad_plan turns a plan's tile_kmask into the work items the kernel deals to its CTAs: item = (tile block, offset) ->
 * the tiles of the block that have the offset; plan int32 [mk_conv_wgrad_plan_elems(num_tiles, K)] =
 * {n_items, n_units, tiles per block, 0 | items [start, k << 26 | count] ordered by (block, k) | unit tile ids | scratch}.
 * x bf16 [n_in,cin] (cin % 8 == 0), dy bf16 [n_out,cout] (cout % 16 == 0, <= 256), nbr / out_rows as mk_conv_fwd_bf16,
 * dW fp32 [K,cin,cout] (zeroed inside). */
int64_t mk_conv_wgrad_plan_elems(int64_t num_tiles, int K);
int mk_conv_wgrad_plan(const uint32_t* tile_kmask, int64_t num_tiles, int K, int32_t* plan, void* stream);
int mk_conv_wgrad_bf16(const void* x, int64_t n_in, int cin, const void* dy, int cout, const int32_t* nbr, int K, int64_t n_slots,
                       const int32_t* out_rows, const int32_t* plan, int64_t n_out, float* dW, void* stream);
/* Same with a workspace: when mk_set_deterministic(1) is active and `workspace` holds mk_conv_wgrad_det_workspace_bytes() bytes, every
 * (tile block, offset) work item writes its accumulator to its own slot and the slots of an offset are added in item order
 * (run-to-run bit-identical, no floating-point REDs); otherwise identical to mk_conv_wgrad_bf16. */
int64_t mk_conv_wgrad_det_workspace_bytes(int64_t n_slots, int K, int cin, int cout);
int mk_conv_wgrad_bf16_ws(const void* x, int64_t n_in, int cin, const void* dy, int cout, const int32_t* nbr, int K, int64_t n_slots,
                          const int32_t* out_rows, const int32_t* plan, int64_t n_out, float* dW, void* workspace, int64_t workspace_bytes,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * (4) Normalisation / activation.  Replaces MinkowskiInstanceNorm (global avg-pool + broadcast kernels,
 * mask3d/models/modules/common.py:24), MinkowskiBatchNorm (:22), MinkowskiReLU and SparseTensor += (resnet_block.py:65).
 * Instance norm: per scene, per channel, biased variance, eps 1e-8.
 * ------------------------------------------------------------------------------------------- */
int64_t mk_instnorm_workspace_bytes(int64_t n, int C, int B);
/* mean/rstd fp32 [B,C].  Deterministic two-stage reduction (fixed merge order, fp64 merge). */
int mk_instnorm_stats(const void* x, int x_dtype, int64_t n, int C, const int32_t* scene_offsets, int B, float eps,
                      float* mean, float* rstd, void* workspace, int64_t workspace_bytes, void* stream);
/* y = (x - mean[b]) * rstd[b] * gamma + beta  (+ residual)  (relu) ; gamma/beta fp32 [C]; residual/out share out_dtype. */
int mk_norm_apply(const void* x, int x_dtype, int64_t n, int C, const int32_t* scene_offsets, int B, const float* mean,
                  const float* rstd, const float* gamma, const float* beta, const void* residual, int relu, void* out,
                  int out_dtype, void* stream);
/* dx of instance norm (ME's custom backward): given dy (already multiplied by gamma) and y_hat = (x-mean)*rstd. */
int mk_instnorm_bwd(const float* x, const float* dy, int64_t n, int C, const int32_t* scene_offsets, int B,
                    const float* mean, const float* rstd, const float* gamma, float* dx, float* dgamma, float* dbeta,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Fused backward of (instance norm -> [+ residual] -> [ReLU]) for fp32 or bf16 tensors: dy = gradient w.r.t. the unit's output,
 * out = that output (NULL: no ReLU), x = the un-normalised input of the norm (the conv output).  Writes dx (gradient w.r.t. x;
 * dtype `dx_dtype`, fp32 allowed for bf16 inputs), dres (NULL or the gradient of the residual branch = masked dy), dgamma /
 * dbeta fp32 [C].  One statistics pass (16-byte loads, deterministic merge) + one apply pass; replaces ME's five broadcast /
 * pooling launches of MinkowskiInstanceNormFunction.backward plus the ReLU backward.  workspace as mk_instnorm_workspace_bytes. */
int mk_instnorm_relu_bwd(const void* dy, const void* out, const void* x, int dtype, int64_t n, int C, const int32_t* scene_offsets,
                         int B, const float* mean, const float* rstd, const float* gamma, void* dx, int dx_dtype, void* dres,
                         float* dgamma, float* dbeta, void* workspace, int64_t workspace_bytes, void* stream);

/* Composite: conv (tensor cores) -> instance norm -> (+ residual) -> (ReLU) in one call, the unit Res16UNet34C repeats per
 * layer (mask3d/models/res16unet.py:244-307, mask3d/models/modules/resnet_block.py:52-68).  Conv arguments as
 * mk_conv_fwd_bf16; tmp = scratch of tmp_bytes >= 2 * n_out * cout holding the un-normalised conv output as bf16 — or as fp32 when
 * tmp_bytes >= 4 * n_out * cout and the level is small enough for the conv's offset split (then several CTAs per tile); stats = fp32 [2,B,cout]
 * (mean, rstd: outputs); residual/out share out_dtype; workspace as mk_instnorm_workspace_bytes(n_out, cout, B). */
int mk_conv_instnorm_bf16(const void* in, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K,
                          int64_t n_slots, const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                          const float* bias, const int32_t* scene_offsets, int B, float eps, const float* gamma, const float* beta,
                          const void* residual, int relu, void* tmp, int64_t tmp_bytes, float* stats, void* out, int out_dtype,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (5) Pooling.  Replaces MinkowskiAvgPooling / MinkowskiSumPooling (mask3d/models/mask3d.py:209,297,547).
 * out[o] = sum over existing children / (#children if average).  nbr int32 [K,n_out].
 * ------------------------------------------------------------------------------------------- */
int mk_pool_fwd(const float* in, int C, const int32_t* nbr, int K, int64_t n_out, int average, float* out, void* stream);
/* same, input fp32 or bf16 (in_dtype), output fp32: pools the bf16 mask features without a cast pass */
int mk_pool_fwd_dt(const void* in, int in_dtype, int C, const int32_t* nbr, int K, int64_t n_out, int average, float* out, void* stream);
int mk_pool_bwd(const float* dout, int C, const int32_t* nbr, int K, int64_t n_out, int average, int64_t n_in, float* din,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * (6) Decoder.  mask-logit product of Mask3D.mask_module (mask3d/models/mask3d.py:534-537), its pooled
 * sigmoid<0.5 attention mask (:544-553) and the masked cross-attention (:397-478, CrossAttentionLayer :686-705).
 * ------------------------------------------------------------------------------------------- */
/* logits[n,q] = feat[n,:] . embed[b(n),q,:]   feat [n,C] (x_dtype), embed fp32 [B,Q,C], logits fp32 [n,Q]. */
int mk_mask_logits(const void* feat, int x_dtype, int64_t n, int C, const int32_t* scene_offsets, int B,
                   const float* embed, int Q, float* logits, void* stream);
/* One pooling step on logits followed (optionally) by the threshold: blocked[o,q] = sigmoid(pooled) < 0.5. */
/* Compact inference result of one decoder output (what trainer/trainer.py:373-402 derives from the mask logits): bits uint32 [n]
 * (bit q = logit[n, q] > 0, i.e. sigmoid > 0.5; Q <= 32) and score fp32 [B, Q] = sum_n sigmoid(x) [x > 0] / (sum_n [x > 0] + 1e-6) per
 * scene, in ONE pass over logits fp32 [n, Q]; deterministic (fixed-order merge of per-block partial sums). */
int64_t mk_mask_compact_workspace_bytes(int64_t n, int B);
int mk_mask_compact(const float* logits, int64_t n, int Q, const int32_t* scene_offsets, int B, uint32_t* bits, float* score, void* workspace,
                    int64_t workspace_bytes, void* stream);
int mk_pool_threshold(const float* pooled, int64_t n, int Q, uint8_t* blocked, void* stream);
/* Per-scene min / max of fp32 [n,3] coordinates (input_range of the fourier embedding, mask3d.py:267-268,323-328). */
int mk_scene_minmax(const float* xyz, int64_t n, const int32_t* scene_offsets, int B, float* cmin, float* cmax, void* stream);
/* Fourier positional encoding with normalisation (mask3d/models/position_embedding.py:127-156):
 * out[n, :] = (base[n, :] if base else 0) + [sin(p), cos(p)], p = ((xyz - min_b) / (max_b - min_b) * 2pi) @ gauss_B.
 * gauss_B fp32 [3, D/2]; base / out [n, D]. */
int mk_fourier_pe(const float* xyz, int64_t n, const int32_t* scene_offsets, int B, const float* cmin, const float* cmax,
                  const float* gauss_B, int D, const void* base, int base_dtype, void* out, int out_dtype, void* stream);
/* 1: force the CUDA-core forward kernel of the masked cross-attention (A/B measurements); default 0: bf16 keys / values with
 * D = 128, H = 8, Q <= 32 run on the tensor-core kernel (mma.sync, warp-level online softmax). */
int mk_set_xattn_simt(int on);
int64_t mk_xattn_workspace_bytes(int B, int Q, int H, int64_t n_keys);
/* Masked cross-attention core of one decoder layer, all scenes, no padding (varlen over scene_offsets):
 *   ctx[b,q,h,:] = softmax_k( scale * qp[b,q,h,:] . kp[k,h,:]  masked ) @ vp[k,h,:]
 * qp fp32 [B,Q,D] (projected queries), kp / vp [n,D] (kv_dtype; projected keys / values of the level),
 * blocked uint8 [n,Q] (1 = key blocked for the query).  Per (scene, query): blocked keys get -inf unless ALL keys of
 * the scene are blocked for that query, in which case none is (mask3d/models/mask3d.py:456); pads do not exist.
 * ctx fp32 [B,Q,D] = concatenated heads BEFORE out_proj. */
int mk_masked_xattn_fwd(const float* qp, const void* kp, const void* vp, int kv_dtype, int64_t n, const uint8_t* blocked,
                        const int32_t* scene_offsets, int B, int Q, int D, int H, float scale, float* ctx, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* Same, additionally writing lse fp32 [B,Q,H] = log-sum-exp of the masked scaled scores in log2 units (what the backward
 * recomputes the attention weights from); lse may be NULL. */
int mk_masked_xattn_fwd_lse(const float* qp, const void* kp, const void* vp, int kv_dtype, int64_t n, const uint8_t* blocked,
                            const int32_t* scene_offsets, int B, int Q, int D, int H, float scale, float* ctx, float* lse,
                            void* workspace, int64_t workspace_bytes, void* stream);
int64_t mk_xattn_bwd_workspace_bytes(int B, int Q, int H, int64_t n_keys);
/* Backward of the core (what autograd does through nn.MultiheadAttention's softmax(QK^T)V, mask3d/models/mask3d.py:686-705):
 * given dctx [B,Q,D] writes dqp fp32 [B,Q,D], dkp / dvp fp32 [n,D] (every row written).  ctx, lse: the forward's outputs.
 * max_scene_rows (host): largest scene of the batch (n is always a valid bound).  No atomics, deterministic. */
int mk_masked_xattn_bwd(const float* qp, const void* kp, const void* vp, int kv_dtype, int64_t n, const uint8_t* blocked,
                        const int32_t* scene_offsets, int B, int Q, int D, int H, float scale, const float* ctx, const float* lse,
                        const float* dctx, int64_t max_scene_rows, float* dqp, float* dkp, float* dvp, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (7) Furthest point sampling.  Replaces pointnet2._ext.furthest_point_sampling
 * (mask3d/third_party/pointnet2/_ext_src/src/sampling_gpu.cu:72-176; call site mask3d/models/mask3d.py:309-317)
 * with identical tie-breaking; all scenes together, many CTAs per scene, one launch per sampling round.
 * xyz fp32 [n_total,3], temp fp32 [n_total] scratch, idx int32 [B,m] (indices local to the scene).
 * ------------------------------------------------------------------------------------------- */
int64_t mk_fps_workspace_bytes(int B, int m);
int mk_fps(const float* xyz, const int32_t* scene_offsets, int B, int m, int64_t n_total, float* temp, int32_t* idx,
           void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (8) Loss side of the mask logits.  Replaces, per decoder output and scene,
 *   batch_sigmoid_ce_loss / batch_dice_loss         mask3d/models/matcher.py:17-56 (cost matrices, called at :143-146)
 *   sigmoid_ce_loss / dice_loss of matched pairs     mask3d/models/criterion.py:27-72 (called at :195-196)
 *   mask score of the post-processing                mask3d/trainer/trainer.py:389-392
 * with one pass over the fp32 logits [n,Q] (all scenes of the batch, rows of a scene contiguous).
 * Targets are per-voxel bit words: bits uint32 [n,W], bit (t & 31) of word (t >> 5) = voxel belongs to instance t of ITS
 * scene (ItemTargets.instances_masks[t, voxel], mask3d/datasets/utils.py:149-160).
 * ------------------------------------------------------------------------------------------- */
/* masks uint8 [T, n_scene] (the bool instance masks of ONE scene) -> bits [n_scene, W] (pass bits + row0 * W). */
int mk_target_bits(const uint8_t* masks, int T, int64_t n_scene, int W, uint32_t* bits, void* stream);
int64_t mk_mask_costs_workspace_bytes(int B, int64_t max_scene_rows, int Q, int W);
/* Outputs, all fp32, Tp = 32 W (columns t >= T of a scene are those of an empty target):
 *   cost_mask [B,Q,Tp]  mean_n BCEWithLogits(x[n,q], tgt[t,n])                        (matcher.py:39-56)
 *   cost_dice [B,Q,Tp]  1 - (2 sum_n sig(x) tgt + 1) / (sum_n sig(x) + sum_n tgt + 1)  (matcher.py:17-33)
 *   dice_den  [B,Q,Tp]  the dice denominator (backward needs it)
 *   mask_score [B,Q]    sum_n sig(x) [x > 0] / (sum_n [x > 0] + 1e-6)                  (trainer.py:389-392); NULL = not wanted
 *   n_rows [B]          rows that took part (scene size, or the selected rows)
 * row_select uint8 [n] or NULL: rows with 0 are left out (the matcher's / criterion's `point_idx` subset when
 * num_points != -1, matcher.py:125-127, criterion.py:181-182).  max_scene_rows (host) = largest scene of the batch. */
int mk_mask_costs(const float* logits, int64_t n, int Q, const uint32_t* bits, int W, const uint8_t* row_select,
                  const int32_t* scene_offsets, int B, int64_t max_scene_rows, float* cost_mask, float* cost_dice, float* dice_den,
                  float* mask_score, float* n_rows, void* workspace, int64_t workspace_bytes, void* stream);
/* Gradient of the matched mask losses w.r.t. the logits: match int32 [B,Q] = target matched to the query or -1,
 * coef fp32 [B,Q,3] = (a, c1, c2):  grad[n,q] = a (sig - tgt) + sig (1 - sig) (c1 - tgt c2);  unmatched / unselected = 0. */
int mk_mask_loss_bwd(const float* logits, int64_t n, int Q, const uint32_t* bits, int W, const uint8_t* row_select,
                     const int32_t* scene_offsets, int B, int64_t max_scene_rows, const int32_t* match, const float* coef, float* grad,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Casts (plumbing between fp32 module boundaries and the bf16 tensor-core path).
 * ------------------------------------------------------------------------------------------- */
int mk_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n_elems, void* stream);
/* dst[n, c_off : c_off + C] = src[n, :]  (channel concat without a separate pass over the other operand) */
int mk_copy_channels(const void* src, int src_dtype, int C, void* dst, int dst_dtype, int C_dst, int c_off, int64_t n,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MK3D_H_ */
