// Composite entry points: the launch sequences the backbone repeats per layer, enqueued by ONE C-ABI call so that the host
// side (Python here, Python in the reference) pays one call instead of one per kernel.
//
// mk_conv_instnorm_bf16 = MinkowskiConvolution -> MinkowskiInstanceNorm -> (+= residual) -> MinkowskiReLU, the unit
// Res16UNet34C applies 62 times per forward (reference: mask3d/models/res16unet.py:244-307, modules/resnet_block.py:52-68).
#include "common.cuh"

using namespace mk;

extern "C" {

int mk_conv_instnorm_bf16(const void* in, int64_t n_in, int cin, const void* wimg, int cout, const int32_t* nbr, int K,
                          int64_t n_slots, const int32_t* out_rows, const uint32_t* tile_kmask, const int32_t* kidx, int64_t n_out,
                          const float* bias, const int32_t* scene_offsets, int B, float eps, const float* gamma, const float* beta,
                          const void* residual, int relu, void* tmp, int64_t tmp_bytes, float* stats, void* out, int out_dtype,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  MK_REQUIRE(tmp != nullptr && stats != nullptr && out != nullptr, "tmp, stats and out are required");
  MK_REQUIRE(tmp_bytes >= n_out * (int64_t)cout * 2, "tmp too small");
  // coarse levels (fewer tiles than half the SMs): fp32 intermediate, which lets the conv split a tile's offsets over several
  // CTAs (mk_conv_fwd_bf16); the statistics and the apply pass read either dtype
  const bool small = K > 1 && 2 * ((n_slots + MK_TILE_ROWS - 1) / MK_TILE_ROWS) <= num_sms();
  const int tmp_dtype = (small && tmp_bytes >= n_out * (int64_t)cout * 4) ? MK_F32 : MK_BF16;
  int rc = mk_conv_fwd_bf16(in, n_in, cin, wimg, cout, nbr, K, n_slots, out_rows, tile_kmask, kidx, n_out, bias, tmp, tmp_dtype, 0, 0,
                            stream);
  if (rc) return rc;
  float* mean = stats;
  float* rstd = stats + (int64_t)B * cout;
  rc = mk_instnorm_stats(tmp, tmp_dtype, n_out, cout, scene_offsets, B, eps, mean, rstd, workspace, workspace_bytes, stream);
  if (rc) return rc;
  return mk_norm_apply(tmp, tmp_dtype, n_out, cout, scene_offsets, B, mean, rstd, gamma, beta, residual, relu, out, out_dtype, stream);
}

}  // extern "C"
