"""Functional wrappers over the C ABI: one Python function per kernel family, raw torch tensors in and out.

Three compute modes for the convolution (BASELINE.json north_star):
  "bf16"  tcgen05 tensor-core path (mk_conv_fwd_bf16): bf16 operands, fp32 accumulation in TMEM — the benchmarked mode;
  "tf32"  the same tcgen05 kernel on fp32 activations with tcgen05.mma.kind::tf32 (mk_conv_fwd_tf32) — the REFERENCE's own
          precision (fp32 tensors, torch.set_float32_matmul_precision("high"), mask3d/main_instance_segmentation.py:52), <= 1e-3
          relative vs the oracle per op;
  "fp32"  exact-fp32 CUDA-core path (mk_conv_fwd_f32) — a test-only checker (no tensor cores).
"""
from __future__ import annotations

import weakref
from typing import Dict, Optional, Tuple

import torch

from . import lib
from .engine import KernelMap

_MODE = "bf16"
_DETERMINISTIC = False
SORT_TILES = True  # mask-sorted tile plans for the tensor-core conv (engine.KernelMap.sorted_plan)
USE_COMPOSITE = True  # conv -> instance norm -> act through one C-ABI call; bench.py turns it off to time conv launches alone


def set_compute_mode(mode: str) -> None:
    global _MODE
    assert mode in ("fp32", "bf16", "tf32")
    _MODE = mode


def get_compute_mode() -> str:
    return _MODE


def set_deterministic(on: bool) -> None:
    """Run-to-run bit-identical results (what the reference's `deterministic: true`, conf/trainer/trainer.yaml:1, asks of a
    backend): the tensor-core conv stops splitting the offsets of small levels over CTAs (floating-point REDs in varying order).
    Off by default (fastest)."""
    global _DETERMINISTIC
    _DETERMINISTIC = bool(on)
    lib.call("mk_set_deterministic", int(bool(on)))


# ----------------------------------------------------------------------------------------------------
# weight images (bf16, pre-swizzled) — cached per parameter version
# ----------------------------------------------------------------------------------------------------
_wimg_cache: Dict[Tuple[int, int], Tuple[int, torch.Tensor, "weakref.ref"]] = {}


def weight_image(W: torch.Tensor, transpose: bool = False, tf32: bool = False) -> torch.Tensor:
    """Pre-swizzled weight image of W fp32 [K, cin, cout] (or [cin, cout]): bf16 for mk_conv_fwd_bf16, fp32 for mk_conv_fwd_tf32."""
    key = (W.data_ptr(), int(transpose) + 2 * int(tf32), tuple(W.shape))
    ent = _wimg_cache.get(key)
    if ent is not None and ent[0] == W._version and ent[2]() is not None:
        lib.cache_use(ent[3])
        return ent[1]
    W3 = W.detach()
    if W3.dim() == 2:
        W3 = W3.unsqueeze(0)
    W3 = W3.contiguous().float()
    K, cin, cout = W3.shape
    if tf32:
        n = lib.query("mk_conv_weight_image_tf32_elems", K, cin, cout, int(transpose))
        img = torch.empty(n, dtype=torch.float32, device=W.device)
        lib.call("mk_conv_weight_image_tf32", lib.p(W3), K, cin, cout, int(transpose), lib.p(img), lib.stream())
    else:
        n = lib.query("mk_conv_weight_image_elems", K, cin, cout, int(transpose))
        img = torch.empty(n, dtype=torch.bfloat16, device=W.device)
        lib.call("mk_conv_weight_image_bf16", lib.p(W3), K, cin, cout, int(transpose), lib.p(img), lib.stream())
    if len(_wimg_cache) > 4096:
        _wimg_cache.clear()
    _wimg_cache[key] = (W._version, img, weakref.ref(W), lib.cache_built())
    return img


_kidx_cache: Dict[Tuple[int, str], torch.Tensor] = {}


def reversed_offsets(K: int, device) -> torch.Tensor:
    """int32 [K-1, ..., 0]: the offset permutation of the same-map dgrad (offset k of the transposed table reads W[K-1-k]);
    cached per (K, device) — one tiny kernel per backward node otherwise."""
    key = (K, str(device))
    t = _kidx_cache.get(key)
    if t is None:
        t = _kidx_cache[key] = torch.arange(K - 1, -1, -1, device=device, dtype=torch.int32)
    return t


def _flat32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """fp32 1-D view of a (small) parameter; no copy when it already is contiguous fp32."""
    if t is None:
        return None
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.reshape(-1) if t.is_contiguous() else t.contiguous().reshape(-1)


def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if x.dtype == dtype:
        return x
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    lib.call("mk_cast", lib.p(x.contiguous()), lib.dt(x), lib.p(out), lib.dt(out), x.numel(), lib.stream())
    return out


# ----------------------------------------------------------------------------------------------------
# convolution
# ----------------------------------------------------------------------------------------------------
def conv_forward(x: torch.Tensor, W: torch.Tensor, km: KernelMap, bias: Optional[torch.Tensor] = None, mode: Optional[str] = None,
                 out_dtype: Optional[torch.dtype] = None, kidx: Optional[torch.Tensor] = None, transpose_w: bool = False,
                 out_cols: int = 0) -> torch.Tensor:
    """out[o] = sum_k x[nbr[k, o]] @ W[k] (+ bias).  W fp32 [K, cin, cout]; x [n_in, cin] fp32 or bf16."""
    mode = mode or _MODE
    W3 = W if W.dim() == 3 else W.unsqueeze(0)
    K, cin, cout = W3.shape
    if transpose_w:
        cin, cout = cout, cin
    assert K == km.K and x.shape[1] == cin, f"shape mismatch: W {tuple(W3.shape)}, x {tuple(x.shape)}, K {km.K}"
    bias_flat = _flat32(bias)
    if cin == 1 and not transpose_w and km.out_rows is None:
        out_dtype = out_dtype or (torch.bfloat16 if mode == "bf16" else torch.float32)
        out = torch.empty((km.n_out, cout), dtype=out_dtype, device=x.device)
        lib.call("mk_conv_stem_f32", lib.p(cast(x, torch.float32).contiguous()), lib.p(W3.detach().float().contiguous()), cout, lib.p(km.nbr),
                 K, km.n_out, lib.p(out), lib.dt(out), lib.stream())
        if bias_flat is not None:
            out += bias_flat.to(out.dtype)
        return out
    if mode == "fp32":
        assert km.out_rows is None and not transpose_w
        out = torch.empty((km.n_out, cout), dtype=torch.float32, device=x.device)
        lib.call("mk_conv_fwd_f32", lib.p(cast(x, torch.float32).contiguous()), cin, lib.p(W3.detach().float().contiguous()), cout,
                 lib.p(km.nbr), K, km.n_out, lib.p(bias_flat), lib.p(out), lib.stream())
        return out if out_dtype in (None, torch.float32) else cast(out, out_dtype)
    tf32 = mode == "tf32"
    out_dtype = out_dtype or (torch.float32 if tf32 else torch.bfloat16)
    xb = cast(x.contiguous(), torch.float32 if tf32 else torch.bfloat16)
    img = weight_image(W, transpose=transpose_w, tf32=tf32)
    ocols = out_cols or cout
    out = torch.empty((km.n_out, ocols), dtype=out_dtype, device=x.device)
    if lib._timed is not None:  # bench.py instrumentation: algorithmic FLOPs = 2 * P * Cin * Cout
        lib.set_tag((2 * km.num_pairs * cin * ocols, K, cin, cout, km.n_out, "dense" if (K == 1 or out_cols) else "conv"))
    plan = km.sorted_plan() if SORT_TILES else km
    lib.call("mk_conv_fwd_tf32" if tf32 else "mk_conv_fwd_bf16", lib.p(xb), xb.shape[0], cin, lib.p(img), cout, lib.p(plan.nbr), K, plan.n_slots, lib.p(plan.out_rows), lib.p(plan.tile_kmask),
             lib.p(kidx), km.n_out, lib.p(bias_flat), lib.p(out), lib.dt(out), ocols, ocols, lib.stream())
    return out


def conv_dgrad_bf16(dout: torch.Tensor, W: torch.Tensor, km_t: KernelMap, kidx: Optional[torch.Tensor] = None,
                    out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """din[i] = sum_k dout[nbr_t[k, i]] @ W[kidx[k]]^T on the tensor-core kernel (fp32 out).  W fp32 [K, cin, cout]; the
    kernel's accumulator holds at most 256 columns, so wider inputs (block5's 384-channel concat) go in column slices."""
    W3 = W if W.dim() == 3 else W.unsqueeze(0)
    K, cin, cout = W3.shape
    if cin <= 256:
        return conv_forward(dout, W3, km_t, mode="bf16", out_dtype=out_dtype, kidx=kidx, transpose_w=True)
    db = cast(dout.contiguous(), torch.bfloat16)
    din = torch.empty((km_t.n_out, cin), dtype=out_dtype, device=dout.device)
    plan = km_t.sorted_plan() if SORT_TILES else km_t
    nsplit = (cin + 255) // 256
    step = ((cin + nsplit - 1) // nsplit + 15) // 16 * 16
    for c0 in range(0, cin, step):
        c1 = min(cin, c0 + step)
        img = weight_image(W3[:, c0:c1, :], transpose=True)
        lib.call("mk_conv_fwd_bf16", lib.p(db), db.shape[0], cout, lib.p(img), c1 - c0, lib.p(plan.nbr), K, plan.n_slots, lib.p(plan.out_rows),
                 lib.p(plan.tile_kmask), lib.p(kidx), km_t.n_out, None, din.data_ptr() + din.element_size() * c0, lib.dt(din), c1 - c0, cin,
                 lib.stream())
    return din


def conv_wgrad_bf16(x: torch.Tensor, dout: torch.Tensor, km: KernelMap, K: int, cin: int, cout: int) -> torch.Tensor:
    """dW[k] = sum_o x[nbr[k, o]]^T dout[o] on tcgen05 (mk_conv_wgrad_bf16) over the forward's tile plan.  fp32 [K, cin, cout]."""
    xb = cast(x.contiguous(), torch.bfloat16)
    db = cast(dout.contiguous(), torch.bfloat16)
    plan = km.sorted_plan() if SORT_TILES else km
    dW = torch.empty((K, cin, cout), dtype=torch.float32, device=x.device)
    if lib._timed is not None:
        lib.set_tag((2 * km.num_pairs * cin * cout, K, cin, cout, km.n_out, "dense" if K == 1 else "conv"))
    if _DETERMINISTIC:  # per-item accumulators + fixed-order merge instead of floating-point REDs
        wsb = lib.query("mk_conv_wgrad_det_workspace_bytes", plan.n_slots, K, cin, cout)
        ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
        lib.call("mk_conv_wgrad_bf16_ws", lib.p(xb), xb.shape[0], cin, lib.p(db), cout, lib.p(plan.nbr), K, plan.n_slots, lib.p(plan.out_rows),
                 lib.p(plan.wgrad_plan), km.n_out, lib.p(dW), lib.p(ws), wsb, lib.stream())
        return dW
    lib.call("mk_conv_wgrad_bf16", lib.p(xb), xb.shape[0], cin, lib.p(db), cout, lib.p(plan.nbr), K, plan.n_slots, lib.p(plan.out_rows),
             lib.p(plan.wgrad_plan), km.n_out, lib.p(dW), lib.stream())
    return dW


def conv_instnorm_bf16(x: torch.Tensor, W: torch.Tensor, km: KernelMap, scene_offsets: torch.Tensor, B: int, gamma, beta, bias=None,
                       residual: Optional[torch.Tensor] = None, relu: bool = True, eps: float = 1e-8,
                       out_dtype: torch.dtype = torch.bfloat16, keep: bool = False):
    """conv (tcgen05) -> instance norm -> (+ residual) -> (ReLU) through ONE C-ABI call (mk_conv_instnorm_bf16)."""
    W3 = W if W.dim() == 3 else W.unsqueeze(0)
    K, cin, cout = W3.shape
    assert K == km.K and x.shape[1] == cin and x.dtype == torch.bfloat16 and x.is_contiguous()
    img = weight_image(W)
    plan = km.sorted_plan() if SORT_TILES else km
    n = km.n_out
    out = torch.empty((n, cout), dtype=out_dtype, device=x.device)
    stats = torch.empty((2, B, cout), dtype=torch.float32, device=x.device)
    wsb = lib.query("mk_instnorm_workspace_bytes", n, cout, B)
    # keep=True (training): the un-normalised conv output is a tensor of its own, returned with the statistics for the backward
    tmp = torch.empty((n, cout), dtype=torch.bfloat16, device=x.device) if keep else None
    # scratch tmp: 4 bytes per element on small levels so that the composite may use an fp32 intermediate (offset split)
    small = K > 1 and 2 * ((plan.n_slots + lib.TILE - 1) // lib.TILE) <= 148
    tmp_bytes = 0 if keep else (n * cout * (4 if small else 2) + 255) // 256 * 256
    ws = lib.scratch(tmp_bytes + wsb, x.device)
    if residual is not None:
        assert residual.dtype == out_dtype and residual.is_contiguous()
    if lib._timed is not None:
        lib.set_tag((2 * km.num_pairs * cin * cout, K, cin, cout, n, "dense" if K == 1 else "conv"))
    base = ws.data_ptr()
    lib.call("mk_conv_instnorm_bf16", x.data_ptr(), x.shape[0], cin, img.data_ptr(), cout, plan.nbr.data_ptr(), K, plan.n_slots,
             lib.p(plan.out_rows), plan.tile_kmask.data_ptr(), None, n, lib.p(_flat32(bias)), scene_offsets.data_ptr(), B, eps,
             lib.p(_flat32(gamma)), lib.p(_flat32(beta)), lib.p(residual), int(relu), tmp.data_ptr() if keep else base,
             n * cout * 2 if keep else tmp_bytes, stats.data_ptr(), out.data_ptr(), lib.dt(out), base + tmp_bytes, ws.numel() - tmp_bytes,
             lib.stream())
    return (out, tmp, stats) if keep else out


# ---- lean inference path of the composite call: per-layer constants resolved once, per-call work = two allocations + one C call
_layer_plans: Dict[int, tuple] = {}
_wsb_cache: Dict[tuple, int] = {}


def layer_plan(W: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """(weight image, gamma fp32 [C], beta fp32 [C], K, cin, cout) of one conv + instance-norm layer, cached until a parameter changes."""
    key = id(W)
    ver = (W._version, gamma._version, beta._version, W.data_ptr(), gamma.data_ptr())
    ent = _layer_plans.get(key)
    if ent is not None and ent[0] == ver:
        lib.cache_use(ent[2])
        return ent[1]
    W3 = W if W.dim() == 3 else W.unsqueeze(0)
    K, cin, cout = W3.shape
    lp = (weight_image(W), _flat32(gamma), _flat32(beta), K, cin, cout)
    if len(_layer_plans) > 4096:
        _layer_plans.clear()
    _layer_plans[key] = (ver, lp, lib.cache_built())
    return lp


def conv_instnorm_fast(x: torch.Tensor, lp: tuple, km: KernelMap, scene_offsets: torch.Tensor, B: int, residual: Optional[torch.Tensor],
                       relu: bool, out_dtype: torch.dtype, eps: float = 1e-8) -> torch.Tensor:
    """conv_instnorm_bf16 for inference with the per-layer constants of `layer_plan` (no bias, no kept intermediates)."""
    img, g, b, K, cin, cout = lp
    plan = km.sorted_plan() if SORT_TILES else km
    n = km.n_out
    out = torch.empty((n, cout), dtype=out_dtype, device=x.device)
    wk = (n, cout, B)
    wsb = _wsb_cache.get(wk)
    if wsb is None:
        if len(_wsb_cache) > 4096:
            _wsb_cache.clear()
        wsb = _wsb_cache[wk] = lib.query("mk_instnorm_workspace_bytes", n, cout, B)
    small = K > 1 and 2 * ((plan.n_slots + lib.TILE - 1) // lib.TILE) <= 148
    tmp_bytes = (n * cout * (4 if small else 2) + 255) // 256 * 256
    stats_bytes = (2 * B * cout * 4 + 255) // 256 * 256
    ws = lib.scratch(tmp_bytes + stats_bytes + wsb, x.device)
    base = ws.data_ptr()
    if lib._timed is not None:
        lib.set_tag((2 * km.num_pairs * cin * cout, K, cin, cout, n, "dense" if K == 1 else "conv"))
    out_rows = plan.out_rows
    lib.call("mk_conv_instnorm_bf16", x.data_ptr(), x.shape[0], cin, img.data_ptr(), cout, plan.nbr.data_ptr(), K, plan.n_slots,
             out_rows.data_ptr() if out_rows is not None else None, plan.tile_kmask.data_ptr(), None, n, None, scene_offsets.data_ptr(), B, eps,
             g.data_ptr(), b.data_ptr(), residual.data_ptr() if residual is not None else None, int(relu), base, tmp_bytes,
             base + tmp_bytes, out.data_ptr(), lib.BF16 if out_dtype == torch.bfloat16 else lib.F32, base + tmp_bytes + stats_bytes,
             ws.numel() - tmp_bytes - stats_bytes, lib.stream())
    return out


def conv_transpose_forward_fp32(x: torch.Tensor, W: torch.Tensor, plan: KernelMap, bias=None) -> torch.Tensor:
    """fp32 parity path of the transposed conv: out[f] = x[parent[f]] @ W[koff[f]], expressed through the generic
    output-stationary kernel with the plan's table restricted to real rows (slot -> row permutation undone)."""
    W3 = W if W.dim() == 3 else W.unsqueeze(0)
    K, cin, cout = W3.shape
    # identity-ordered table [K, n_fine]: nbr[k, f] = parent[f] if koff[f] == k else -1
    ks = torch.arange(K, device=x.device, dtype=torch.int32)[:, None]
    nbr = torch.where(plan.koff[None, :] == ks, plan.parent[None, :], torch.full_like(plan.parent[None, :], -1)).contiguous()
    out = torch.empty((plan.n_out, cout), dtype=torch.float32, device=x.device)
    bias_flat = bias.detach().reshape(-1).float().contiguous() if bias is not None else None
    lib.call("mk_conv_fwd_f32", lib.p(cast(x, torch.float32).contiguous()), cin, lib.p(W3.detach().float().contiguous()), cout, lib.p(nbr), K,
             plan.n_out, lib.p(bias_flat), lib.p(out), lib.stream())
    return out


# ----------------------------------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------------------------------
def instance_norm_stats(x: torch.Tensor, scene_offsets: torch.Tensor, B: int, eps: float = 1e-8):
    n, C = x.shape
    stats = torch.empty((2, B, C), dtype=torch.float32, device=x.device)
    mean, rstd = stats[0], stats[1]
    wsb = lib.query("mk_instnorm_workspace_bytes", n, C, B)
    ws = lib.scratch(wsb, x.device)
    lib.call("mk_instnorm_stats", lib.p(x), lib.dt(x), n, C, lib.p(scene_offsets), B, eps, stats.data_ptr(), stats.data_ptr() + 4 * B * C,
             lib.p(ws), ws.numel(), lib.stream())
    return mean, rstd


def norm_apply(x, scene_offsets, B, mean, rstd, gamma, beta, residual=None, relu=False, out_dtype=None) -> torch.Tensor:
    n, C = x.shape
    out_dtype = out_dtype or x.dtype
    out = torch.empty((n, C), dtype=out_dtype, device=x.device)
    if residual is not None:
        residual = cast(residual.contiguous(), out_dtype)
    g, b = _flat32(gamma), _flat32(beta)
    lib.call("mk_norm_apply", lib.p(x), lib.dt(x), n, C, lib.p(scene_offsets), B, lib.p(mean), lib.p(rstd), lib.p(g), lib.p(b), lib.p(residual),
             int(relu), lib.p(out), lib.dt(out), lib.stream())
    return out


def instance_norm(x, scene_offsets, B, gamma, beta, residual=None, relu=False, out_dtype=None, eps=1e-8) -> torch.Tensor:
    x = x.contiguous()
    mean, rstd = instance_norm_stats(x, scene_offsets, B, eps)
    return norm_apply(x, scene_offsets, B, mean, rstd, gamma, beta, residual, relu, out_dtype)


def instnorm_relu_bwd(dy: torch.Tensor, out: Optional[torch.Tensor], x: torch.Tensor, scene_offsets: torch.Tensor, B: int, mean, rstd,
                      gamma, want_dres: bool = False, dx_dtype: Optional[torch.dtype] = None):
    """Backward of (instance norm -> [+ residual] -> [ReLU]) in one statistics + one apply pass (mk_instnorm_relu_bwd).
    dy / out / x share a dtype (fp32 or bf16); `out` = the unit's output when it ends in a ReLU, else None.
    Returns (dx, dres or None, dgamma fp32 [C], dbeta fp32 [C])."""
    n, C = x.shape
    assert dy.dtype == x.dtype and dy.is_contiguous() and x.is_contiguous() and (out is None or (out.dtype == x.dtype and out.is_contiguous()))
    dx = torch.empty((n, C), dtype=dx_dtype or x.dtype, device=x.device)
    dres = torch.empty((n, C), dtype=x.dtype, device=x.device) if want_dres else None
    dgb = torch.empty((2, C), dtype=torch.float32, device=x.device)
    wsb = lib.query("mk_instnorm_workspace_bytes", n, C, B)
    ws = lib.scratch(wsb, x.device)
    lib.call("mk_instnorm_relu_bwd", dy.data_ptr(), lib.p(out), x.data_ptr(), lib.dt(x), n, C, scene_offsets.data_ptr(), B, mean.data_ptr(),
             rstd.data_ptr(), lib.p(_flat32(gamma)), dx.data_ptr(), lib.dt(dx), lib.p(dres), dgb.data_ptr(), dgb.data_ptr() + 4 * C,
             ws.data_ptr(), ws.numel(), lib.stream())
    return dx, dres, dgb[0], dgb[1]


# ----------------------------------------------------------------------------------------------------
# pooling, decoder pieces, FPS
# ----------------------------------------------------------------------------------------------------
def pool_forward(x: torch.Tensor, km: KernelMap, average: bool = True) -> torch.Tensor:
    """fp32 [n_out, C] = avg / sum over the existing children of every output row; x fp32 or bf16 (read as is)."""
    x = x.contiguous()
    out = torch.empty((km.n_out, x.shape[1]), dtype=torch.float32, device=x.device)
    lib.call("mk_pool_fwd_dt", lib.p(x), lib.dt(x), x.shape[1], lib.p(km.nbr), km.K, km.n_out, int(average), lib.p(out), lib.stream())
    return out


def pool_backward(dout: torch.Tensor, km: KernelMap, n_in: int, average: bool = True) -> torch.Tensor:
    dout = dout.contiguous().float()
    din = torch.empty((n_in, dout.shape[1]), dtype=torch.float32, device=dout.device)
    lib.call("mk_pool_bwd", lib.p(dout), dout.shape[1], lib.p(km.nbr), km.K, km.n_out, int(average), n_in, lib.p(din), lib.stream())
    return din


def mask_logits(feat: torch.Tensor, scene_offsets: torch.Tensor, B: int, embed: torch.Tensor, select: Optional[KernelMap] = None,
                padded: bool = False) -> torch.Tensor:
    """logits[n, q] = feat[n] . embed[b(n), q]   (mask3d.py:534-537).

    `padded` (tensor-core path only): return [n, Qp] with Qp = Q rounded up to 16 (zero columns at the end) -- a row pitch the
    epilogue can store with 16-byte vectors; used for the stacked products of several decoder outputs (`mask_logits_stacked`).

    bf16 features + a scene-select table (engine.CoordinateManager.scene_select_table): the product runs on the tcgen05
    kernel as a conv with one "offset" per scene, W[b] = embed[b]^T zero-padded to 32 accumulator columns per 32 queries,
    storing only the Q real columns.  fp32 features: exact-fp32 CUDA-core kernel (parity mode)."""
    n, C = feat.shape
    Q = embed.shape[1]
    tc_mode = "bf16" if feat.dtype == torch.bfloat16 else ("tf32" if _MODE == "tf32" else None)
    if tc_mode is not None and select is not None and select.K <= 32 and C % 16 == 0 and Q <= 256:
        Qp = ((Q + 15) // 16) * 16
        W = torch.zeros((embed.shape[0], C, Qp), dtype=torch.float32, device=feat.device)
        W[:, :, :Q] = embed.detach().float().transpose(1, 2)
        return conv_forward(feat, W, select, mode=tc_mode, out_dtype=torch.float32, out_cols=Qp if padded else Q)
    assert not padded, "padded mask logits are a tensor-core path feature"
    out = torch.empty((n, Q), dtype=torch.float32, device=feat.device)
    lib.call("mk_mask_logits", lib.p(feat.contiguous()), lib.dt(feat), n, C, lib.p(scene_offsets), B, lib.p(embed.detach().float().contiguous()),
             Q, lib.p(out), lib.stream())
    return out


def mask_logits_stacked(feat: torch.Tensor, scene_offsets: torch.Tensor, B: int, embeds, select: KernelMap, group: int = 6):
    """Mask logits of SEVERAL decoder outputs in one pass over the features per `group` outputs: the embeddings [B, Q, C] of a
    group are stacked to [B, group * Q, C], one product writes [n, pad16(group * Q)] and output l is the column slice
    [:, l * Q:(l + 1) * Q] (a strided view).  Every logit is the same K = C dot product on the same tensor-core path as in
    `mask_logits`, so the values are bit-identical; the features (the largest tensor of the model: [N0, 128]) are read once per
    group instead of once per output."""
    Q = embeds[0].shape[1]
    outs = []
    for g0 in range(0, len(embeds), group):
        part = embeds[g0:g0 + group]
        big = mask_logits(feat, scene_offsets, B, torch.cat(part, dim=1), select=select, padded=True)
        outs += [big[:, i * Q:(i + 1) * Q] for i in range(len(part))]
    return outs


class _MaskLogitsFn(torch.autograd.Function):
    """Differentiable mask-logit product (mask3d.py:534-537): logits = feat . embed[b]^T.  bf16 mode: forward, d(feat) and
    d(embed) all run on the tcgen05 conv kernels over the scene-select table (one "offset" per scene; the gradient w.r.t. the
    logits is zero-padded to a multiple of 16 columns in bf16).  fp32 parity mode: per-scene fp32 matrix products."""

    @staticmethod
    def forward(ctx, feat, embed, scene_offsets, B, select, offs_host):
        ctx.save_for_backward(feat, embed, scene_offsets)
        ctx.B, ctx.select, ctx.offs_host = B, select, offs_host
        return mask_logits(feat.detach(), scene_offsets, B, embed.detach(), select=select)

    @staticmethod
    def backward(ctx, dlogits):
        feat, embed, scene_offsets = ctx.saved_tensors
        B, select = ctx.B, ctx.select
        n, C = feat.shape
        Q = embed.shape[1]
        dfeat = demb = None
        if feat.dtype == torch.bfloat16 and select is not None and select.K <= 32 and C % 16 == 0:
            Qp = ((Q + 15) // 16) * 16
            dp = torch.zeros((n, Qp), dtype=torch.bfloat16, device=feat.device)
            dl = dlogits.contiguous().float()
            lib.call("mk_copy_channels", lib.p(dl), lib.F32, Q, lib.p(dp), lib.BF16, Qp, 0, n, lib.stream())
            W = torch.zeros((embed.shape[0], C, Qp), dtype=torch.float32, device=feat.device)
            W[:, :, :Q] = embed.detach().float().transpose(1, 2)
            if ctx.needs_input_grad[0]:
                dfeat = conv_forward(dp, W, select, mode="bf16", out_dtype=feat.dtype, transpose_w=True)
            if ctx.needs_input_grad[1]:
                dW = conv_wgrad_bf16(feat.detach(), dp, select, select.K, C, Qp)  # [B, C, Qp]
                demb = dW[:, :, :Q].transpose(1, 2).to(embed.dtype)
            return dfeat, demb, None, None, None, None
        dl = dlogits.float()
        f32 = feat.detach().float()
        oh = ctx.offs_host
        if ctx.needs_input_grad[0]:
            dfeat = torch.cat([dl[oh[b]:oh[b + 1]] @ embed[b].detach().float() for b in range(B)]).to(feat.dtype)
        if ctx.needs_input_grad[1]:
            demb = torch.stack([dl[oh[b]:oh[b + 1]].t() @ f32[oh[b]:oh[b + 1]] for b in range(B)]).to(embed.dtype)
        return dfeat, demb, None, None, None, None


def mask_embed_image(embed: torch.Tensor, img: Optional[torch.Tensor] = None, W: Optional[torch.Tensor] = None):
    """Weight image of the mask-logit product for the tensor-core path: W[b] = embed[b]^T zero-padded to a multiple of 16
    queries.  Optional preallocated `W` / `img` (CUDA-graph capture: static buffers, no caching by parameter version)."""
    B, Q, C = embed.shape
    Qp = ((Q + 15) // 16) * 16
    if W is None:
        W = torch.zeros((B, C, Qp), dtype=torch.float32, device=embed.device)
    W[:, :, :Q] = embed.detach().float().transpose(1, 2)
    if img is None:
        img = torch.empty(lib.query("mk_conv_weight_image_elems", B, C, Qp, 0), dtype=torch.bfloat16, device=embed.device)
    lib.call("mk_conv_weight_image_bf16", lib.p(W), B, C, Qp, 0, lib.p(img), lib.stream())
    return img, W


def mask_logits_from_image(feat: torch.Tensor, img: torch.Tensor, select: KernelMap, Q: int) -> torch.Tensor:
    """logits [n, Q] fp32 = feat (bf16) x the per-scene embedding image built by `mask_embed_image`."""
    n, C = feat.shape
    Qp = ((Q + 15) // 16) * 16
    assert feat.dtype == torch.bfloat16 and feat.is_contiguous()
    out = torch.empty((n, Q), dtype=torch.float32, device=feat.device)
    if lib._timed is not None:
        lib.set_tag((2 * n * C * Q, select.K, C, Qp, n, "dense"))
    lib.call("mk_conv_fwd_bf16", lib.p(feat), n, C, lib.p(img), Qp, lib.p(select.nbr), select.K, select.n_slots, lib.p(select.out_rows),
             lib.p(select.tile_kmask), None, n, None, lib.p(out), lib.F32, Q, Q, lib.stream())
    return out


def mask_compact(logits: torch.Tensor, scene_offsets: torch.Tensor, B: int):
    """Compact result of a decoder output: (bits int32 [n]: bit q = logits[n, q] > 0, score fp32 [B, Q]: the per-query mask score
    of the reference's post-processing, trainer/trainer.py:389-392) in one pass over the fp32 logits [n, Q], Q <= 32."""
    n, Q = logits.shape
    bits = torch.empty(n, dtype=torch.int32, device=logits.device)
    score = torch.empty((B, Q), dtype=torch.float32, device=logits.device)
    ws = lib.scratch(lib.query("mk_mask_compact_workspace_bytes", n, B), logits.device)
    lib.call("mk_mask_compact", lib.p(logits.contiguous()), n, Q, lib.p(scene_offsets), B, lib.p(bits), lib.p(score), lib.p(ws), ws.numel(),
             lib.stream())
    return bits, score


def threshold_blocked(pooled: torch.Tensor) -> torch.Tensor:
    """uint8 [n, Q]: sigmoid(pooled) < 0.5   (mask3d.py:550)."""
    out = torch.empty(pooled.shape, dtype=torch.uint8, device=pooled.device)
    lib.call("mk_pool_threshold", lib.p(pooled.contiguous()), pooled.shape[0], pooled.shape[1], lib.p(out), lib.stream())
    return out


def scene_minmax(xyz: torch.Tensor, scene_offsets: torch.Tensor, B: int):
    cmin = torch.empty((B, 3), dtype=torch.float32, device=xyz.device)
    cmax = torch.empty((B, 3), dtype=torch.float32, device=xyz.device)
    lib.call("mk_scene_minmax", lib.p(xyz.contiguous()), xyz.shape[0], lib.p(scene_offsets), B, lib.p(cmin), lib.p(cmax), lib.stream())
    return cmin, cmax


def fourier_pe(xyz, scene_offsets, B, cmin, cmax, gauss_B, base: Optional[torch.Tensor] = None, out_dtype=torch.float32) -> torch.Tensor:
    n = xyz.shape[0]
    D = gauss_B.shape[1] * 2
    out = torch.empty((n, D), dtype=out_dtype, device=xyz.device)
    lib.call("mk_fourier_pe", lib.p(xyz.contiguous()), n, lib.p(scene_offsets), B, lib.p(cmin), lib.p(cmax), lib.p(gauss_B.detach().float().contiguous()), D,
             lib.p(base.contiguous()) if base is not None else None, lib.dt(base) if base is not None else lib.F32, lib.p(out), lib.dt(out),
             lib.stream())
    return out


class _MaskedXattnFn(torch.autograd.Function):
    """Masked cross-attention core with its backward (mk_masked_xattn_fwd_lse / mk_masked_xattn_bwd): what autograd does through
    the softmax(QK^T)V of the reference's stock nn.MultiheadAttention (mask3d/models/mask3d.py:686-705)."""

    @staticmethod
    def forward(ctx_, qp, kp, vp, blocked, scene_offsets, num_heads, max_scene_rows):
        B, Q, D = qp.shape
        n = kp.shape[0]
        q32, k_, v_, bl = qp.detach().float().contiguous(), kp.detach().contiguous(), vp.detach().contiguous(), blocked.contiguous()
        ws = lib.scratch(lib.query("mk_xattn_workspace_bytes", B, Q, num_heads, n), qp.device)
        out = torch.empty((B, Q, D), dtype=torch.float32, device=qp.device)
        lse = torch.empty((B, Q, num_heads), dtype=torch.float32, device=qp.device)
        scale = float((D // num_heads) ** -0.5)
        lib.call("mk_masked_xattn_fwd_lse", lib.p(q32), lib.p(k_), lib.p(v_), lib.dt(k_), n, lib.p(bl), lib.p(scene_offsets), B, Q, D, num_heads,
                 scale, lib.p(out), lib.p(lse), lib.p(ws), ws.numel(), lib.stream())
        ctx_.save_for_backward(q32, k_, v_, bl, scene_offsets, out, lse)
        ctx_.meta = (num_heads, scale, int(max_scene_rows) if max_scene_rows is not None else n, qp.dtype, kp.dtype, vp.dtype)
        return out

    @staticmethod
    def backward(ctx_, dout):
        q32, k_, v_, bl, scene_offsets, out, lse = ctx_.saved_tensors
        H, scale, max_rows, qd, kd, vd = ctx_.meta
        B, Q, D = q32.shape
        n = k_.shape[0]
        dout = dout.float().contiguous()
        dq = torch.empty_like(q32)
        dk = torch.empty((n, D), dtype=torch.float32, device=q32.device)
        dv = torch.empty((n, D), dtype=torch.float32, device=q32.device)
        ws = lib.scratch(lib.query("mk_xattn_bwd_workspace_bytes", B, Q, H, n), q32.device)
        lib.call("mk_masked_xattn_bwd", lib.p(q32), lib.p(k_), lib.p(v_), lib.dt(k_), n, lib.p(bl), lib.p(scene_offsets), B, Q, D, H, scale,
                 lib.p(out), lib.p(lse), lib.p(dout), min(max_rows, n), lib.p(dq), lib.p(dk), lib.p(dv), lib.p(ws), ws.numel(), lib.stream())
        return dq.to(qd), dk.to(kd), dv.to(vd), None, None, None, None


def masked_cross_attention(qp: torch.Tensor, kp: torch.Tensor, vp: torch.Tensor, blocked: torch.Tensor, scene_offsets: torch.Tensor,
                           num_heads: int, out: Optional[torch.Tensor] = None, max_scene_rows: Optional[int] = None) -> torch.Tensor:
    """ctx [B, Q, D] (heads concatenated, before out_proj).  Differentiable w.r.t. qp, kp, vp."""
    B, Q, D = qp.shape
    n = kp.shape[0]
    assert kp.dtype == vp.dtype
    if torch.is_grad_enabled() and (qp.requires_grad or kp.requires_grad or vp.requires_grad):
        assert out is None, "out= is an inference-only argument"
        return _MaskedXattnFn.apply(qp, kp, vp, blocked, scene_offsets, num_heads, max_scene_rows)
    wsb = lib.query("mk_xattn_workspace_bytes", B, Q, num_heads, n)
    ws = lib.scratch(wsb, qp.device)
    ctx = out if out is not None else torch.empty((B, Q, D), dtype=torch.float32, device=qp.device)
    assert ctx.is_contiguous() and ctx.dtype == torch.float32
    scale = float((D // num_heads) ** -0.5)
    lib.call("mk_masked_xattn_fwd", lib.p(qp.detach().float().contiguous()), lib.p(kp.contiguous()), lib.p(vp.contiguous()), lib.dt(kp), n,
             lib.p(blocked.contiguous()), lib.p(scene_offsets), B, Q, D, num_heads, scale, lib.p(ctx), lib.p(ws), wsb, lib.stream())
    return ctx


def furthest_point_sample(xyz: torch.Tensor, scene_offsets: torch.Tensor, B: int, m: int) -> torch.Tensor:
    """int32 [B, m] scene-local indices; tie-breaking identical to pointnet2's kernel (csrc/fps.cu)."""
    xyz = xyz.float().contiguous()
    temp = torch.empty(xyz.shape[0], dtype=torch.float32, device=xyz.device)
    idx = torch.empty((B, m), dtype=torch.int32, device=xyz.device)
    wsb = lib.query("mk_fps_workspace_bytes", B, m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=xyz.device)
    lib.call("mk_fps", lib.p(xyz), lib.p(scene_offsets), B, m, xyz.shape[0], lib.p(temp), lib.p(idx), lib.p(ws), wsb, lib.stream())
    return idx
