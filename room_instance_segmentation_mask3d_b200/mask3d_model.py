"""Host-side mirror of the reference model on the hot path: Res16UNet34C backbone + Mask3D query decoder.

State-dict names and shapes are identical to the reference's modules (reference: mask3d/models/res16unet.py:28-314,
mask3d/models/resnet.py:99-152, mask3d/models/modules/resnet_block.py:7-68, mask3d/models/mask3d.py:73-259), so a
reference checkpoint loads with `load_state_dict`.  Two execution paths:

  * `forward`        composes the MinkowskiEngine-compatible modules of `me_api` exactly like the reference does
                     (autograd-capable; what the unmodified reference code would execute on the shim);
  * `forward_fused`  inference fast path: calls the kernels directly, keeps activations in the compute dtype
                     (bf16 on the tcgen05 path), fuses InstanceNorm + residual + ReLU + cast into one pass, replaces the
                     decoder's per-scene Python loops / padding / dense masks by varlen kernels.

The architecture is written from the layer table, not from the reference's source text.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib
from . import me_api as ME
from . import ops
from .engine import CoordinateMapKey

LAYERS = (2, 3, 4, 6, 2, 2, 2, 2)
PLANES = (32, 64, 128, 256, 256, 128, 96, 96)
INIT_DIM = 32


def _conv(cin, cout, k, stride=1, bias=False):
    return ME.MinkowskiConvolution(cin, cout, kernel_size=k, stride=stride, dilation=1, bias=bias, dimension=3)


def _conv_tr(cin, cout):
    return ME.MinkowskiConvolutionTranspose(cin, cout, kernel_size=2, stride=2, dilation=1, bias=False, dimension=3)


class BasicBlock(nn.Module):
    """conv3 -> IN -> ReLU -> conv3 -> IN -> (+ 1x1 conv + IN of the input when widths differ) -> ReLU."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1, self.norm1 = _conv(cin, cout, 3), ME.MinkowskiInstanceNorm(cout)
        self.conv2, self.norm2 = _conv(cout, cout, 3), ME.MinkowskiInstanceNorm(cout)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.downsample = nn.Sequential(_conv(cin, cout, 1), ME.MinkowskiInstanceNorm(cout)) if cin != cout else None

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        out += x if self.downsample is None else self.downsample(x)
        return self.relu(out)


def _stage(cin, cout, n):
    return nn.Sequential(*[BasicBlock(cin if i == 0 else cout, cout) for i in range(n)])


class _ConvNormActFn(torch.autograd.Function):
    """conv (tcgen05) -> instance norm -> (+ residual) -> (ReLU) as ONE autograd node on bf16 activations: the training unit of
    the backbone (reference: the MinkowskiConvolution / MinkowskiInstanceNorm / `+=` / MinkowskiReLU sequence of
    res16unet.py:244-307 and resnet_block.py:52-68, four autograd nodes and fp32 tensors there).  Forward = one C call
    (mk_conv_instnorm_bf16), backward = mk_instnorm_relu_bwd + tcgen05 dgrad + tcgen05 wgrad.  Saved for backward: the bf16
    input, the bf16 un-normalised conv output, the statistics, and the bf16 output (ReLU mask)."""

    @staticmethod
    def forward(ctx, x, W, gamma, beta, residual, km, kind, aux, offs, B, relu):
        x = x.contiguous()
        out, tmp, stats = ops.conv_instnorm_bf16(x, W, km, offs, B, gamma, beta, residual=residual, relu=relu, keep=True)
        ctx.km, ctx.kind, ctx.aux, ctx.B, ctx.relu, ctx.has_res = km, kind, aux, B, relu, residual is not None
        ctx.save_for_backward(x, W, gamma, tmp, stats, out if relu else None, offs)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, W, gamma, tmp, stats, out, offs = ctx.saved_tensors
        km, kind, B = ctx.km, ctx.kind, ctx.B
        W3 = W if W.dim() == 3 else W.unsqueeze(0)
        K, cin, cout = W3.shape
        dy = dy.contiguous()
        dtmp, dres, dgamma, dbeta = ops.instnorm_relu_bwd(dy, out, tmp, offs, B, stats[0], stats[1], gamma, want_dres=ctx.has_res)
        dx = dW = None
        if ctx.needs_input_grad[0]:
            if kind == 0:
                dx = ops.conv_dgrad_bf16(dtmp, W3, km, kidx=ops.reversed_offsets(K, x.device), out_dtype=torch.bfloat16)
            else:
                dx = ops.conv_dgrad_bf16(dtmp, W3, ctx.aux, out_dtype=torch.bfloat16)
        if ctx.needs_input_grad[1]:
            dW = ops.conv_wgrad_bf16(x, dtmp, km, K, cin, cout).reshape(W.shape)
        return dx, dW, dgamma.view_as(gamma), dbeta.view_as(gamma), dres, None, None, None, None, None, None


class _LinFn(torch.autograd.Function):
    """1x1 convolution + bias on the rows of one level with the output in the requested dtype, as ONE autograd node (the
    mask-feature head, lin_squeeze and the K / V in-projections of the training forward): forward = the tcgen05 conv kernel with
    fused bias and cast, backward = tcgen05 dgrad + wgrad + a column sum for the bias."""

    @staticmethod
    def forward(ctx, x, W, bias, km, out_dtype):
        x = x.contiguous()
        ctx.km = km
        ctx.save_for_backward(x, W)
        return ops.conv_forward(x.detach(), W, km, bias=bias, mode="bf16", out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        km = ctx.km
        W3 = W.detach().unsqueeze(0) if W.dim() == 2 else W.detach()
        _, cin, cout = W3.shape
        db = ops.cast(dy.contiguous(), torch.bfloat16)
        dx = dW = dbias = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv_dgrad_bf16(db, W3, km, kidx=ops.reversed_offsets(1, x.device), out_dtype=x.dtype)
        if ctx.needs_input_grad[1]:
            dW = ops.conv_wgrad_bf16(x, db, km, 1, cin, cout).reshape(W.shape)
        if ctx.needs_input_grad[2]:
            dbias = dy.float().sum(0)
        return dx, dW, dbias, None, None


class Res16UNet34C(nn.Module):
    """Sparse U-Net: 4 stride-2 encoder levels, 4 transposed-conv decoder levels with skip concatenation."""

    PLANES, LAYERS = PLANES, LAYERS

    def __init__(self, in_channels=1, out_channels=1, conv1_kernel_size=5, out_fpn=True):
        super().__init__()
        self.out_fpn, self.D = out_fpn, 3
        P, L = PLANES, LAYERS
        self.conv0p1s1, self.bn0 = _conv(in_channels, INIT_DIM, conv1_kernel_size), ME.MinkowskiInstanceNorm(INIT_DIM)
        widths = [INIT_DIM, P[0], P[1], P[2], P[3]]  # block output width per level 0..4
        for lvl in range(1, 5):
            cin = widths[lvl - 1]
            setattr(self, f"conv{lvl}p{2 ** (lvl - 1)}s2", _conv(cin, cin, 2, stride=2))
            setattr(self, f"bn{lvl}", ME.MinkowskiInstanceNorm(cin))
            setattr(self, f"block{lvl}", _stage(cin, P[lvl - 1], L[lvl - 1]))
        cur = P[3]
        for j, lvl in enumerate(range(4, 0, -1)):
            up = P[4 + j]
            setattr(self, f"convtr{4 + j}p{2 ** lvl}s2", _conv_tr(cur, up))
            setattr(self, f"bntr{4 + j}", ME.MinkowskiInstanceNorm(up))
            setattr(self, f"block{5 + j}", _stage(up + widths[lvl - 1], up, L[4 + j]))
            cur = up
        self.final = _conv(P[7], out_channels, 1, bias=True)  # constructed, never called (res16unet.py:231-238)
        self.relu = ME.MinkowskiReLU(inplace=True)

    # ---- module path (reference composition, res16unet.py:241-314)
    def forward(self, x):
        skips = [self.relu(self.bn0(self.conv0p1s1(x)))]
        out = skips[0]
        for lvl in range(1, 5):
            out = self.relu(getattr(self, f"bn{lvl}")(getattr(self, f"conv{lvl}p{2 ** (lvl - 1)}s2")(out)))
            out = getattr(self, f"block{lvl}")(out)
            skips.append(out)
        fmaps = [out]
        for j, lvl in enumerate(range(4, 0, -1)):
            out = self.relu(getattr(self, f"bntr{4 + j}")(getattr(self, f"convtr{4 + j}p{2 ** lvl}s2")(out)))
            out = ME.cat(out, skips[lvl - 1])
            out = getattr(self, f"block{5 + j}")(out)
            fmaps.append(out)
        return (out, fmaps) if self.out_fpn else out

    # ---- fused training path: same composition, every conv + norm + activation one autograd node on bf16 activations
    def forward_train(self, x: ME.SparseTensor):
        """Autograd-capable fused forward (bf16 tensor-core mode).  Returns what `forward` returns; `.F` of the results is
        bf16.  The stem (Cin = 1) runs through the fp32 modules."""
        assert ops.get_compute_mode() == "bf16", "forward_train is the bf16 tensor-core path"
        with lib.stream_scope():
            mgr = x.coordinate_manager
            B = mgr.number_of_unique_batch_indices()
            keys = mgr.prebuild(x.coordinate_map_key, 5, self.conv0p1s1.kernel_generator.kernel_size[0])
            offs = [mgr.scene_offsets(k) for k in keys]

            def unit(f, conv, norm, km, lvl, kind=0, aux=None, relu=True, residual=None):
                assert conv.bias is None
                return _ConvNormActFn.apply(f, conv.kernel, norm.weight, norm.bias, residual, km, kind, aux, offs[lvl], B, relu)

            def block(f, blk, lvl):
                km3 = mgr.neighbor_table(keys[lvl], keys[lvl], 3, 1)
                h = unit(f, blk.conv1, blk.norm1, km3, lvl)
                res = f
                if blk.downsample is not None:
                    res = unit(f, blk.downsample[0], blk.downsample[1], mgr.neighbor_table(keys[lvl], keys[lvl], 1, 1), lvl, relu=False)
                return unit(h, blk.conv2, blk.norm2, km3, lvl, residual=res)

            f = self.relu(self.bn0(self.conv0p1s1(x))).F.to(torch.bfloat16)
            skips = [f]
            for lvl in range(1, 5):
                km = mgr.neighbor_table(keys[lvl - 1], keys[lvl], 2, 1)
                tplan = mgr.transpose_plan(keys[lvl], keys[lvl - 1], 2, 1)
                f = unit(f, getattr(self, f"conv{lvl}p{2 ** (lvl - 1)}s2"), getattr(self, f"bn{lvl}"), km, lvl, kind=1, aux=tplan)
                for blk in getattr(self, f"block{lvl}"):
                    f = block(f, blk, lvl)
                skips.append(f)
            fmaps = [f]
            for j, lvl in enumerate(range(4, 0, -1)):
                plan = mgr.transpose_plan(keys[lvl], keys[lvl - 1], 2, 1)
                down = mgr.neighbor_table(keys[lvl - 1], keys[lvl], 2, 1)
                f = unit(f, getattr(self, f"convtr{4 + j}p{2 ** lvl}s2"), getattr(self, f"bntr{4 + j}"), plan, lvl - 1, kind=2, aux=down)
                f = torch.cat([f, skips[lvl - 1]], dim=1)
                for blk in getattr(self, f"block{5 + j}"):
                    f = block(f, blk, lvl - 1)
                fmaps.append(f)
            wrap = lambda t, k: ME.SparseTensor(t, coordinate_map_key=k, coordinate_manager=mgr)
            out = wrap(f, keys[0])
            fm = [wrap(t, keys[4 - i]) for i, t in enumerate(fmaps)]
            return (out, fm) if self.out_fpn else out

    # ---- fused inference path
    @torch.no_grad()
    def forward_fused(self, x: ME.SparseTensor):
        with lib.stream_scope():
            return self._forward_fused(x)

    def _forward_fused(self, x: ME.SparseTensor):
        mgr = x.coordinate_manager
        mode = ops.get_compute_mode()
        act = torch.bfloat16 if mode == "bf16" else torch.float32
        B = mgr.number_of_unique_batch_indices()
        keys = mgr.prebuild(x.coordinate_map_key, 5, self.conv0p1s1.kernel_generator.kernel_size[0])
        offs = [mgr.scene_offsets(k) for k in keys]

        def cn(f, conv, norm, km, lvl, relu=True, residual=None, kind=0):
            if ops.USE_COMPOSITE and mode == "bf16" and f.dtype == torch.bfloat16 and f.shape[1] % 16 == 0 and conv.bias is None:
                return ops.conv_instnorm_fast(f, ops.layer_plan(conv.kernel, norm.weight, norm.bias), km, offs[lvl], B, residual, relu, act)
            if kind == 2 and mode == "fp32":
                y = ops.conv_transpose_forward_fp32(f, conv.kernel, km)
            else:
                # small levels: fp32 conv output, as the composite call uses (lets the kernel split a tile's offsets over CTAs)
                small = mode != "fp32" and km.K > 1 and 2 * ((km.n_slots + lib.TILE - 1) // lib.TILE) <= 148
                y = ops.conv_forward(f, conv.kernel, km, bias=conv.bias, mode=mode, out_dtype=torch.float32 if small else act)
            return ops.instance_norm(y, offs[lvl], B, norm.weight, norm.bias, residual=residual, relu=relu, out_dtype=act)

        def block(f, blk, lvl):
            km3 = mgr.neighbor_table(keys[lvl], keys[lvl], 3, 1)
            h = cn(f, blk.conv1, blk.norm1, km3, lvl)
            res = f
            if blk.downsample is not None:
                res = cn(f, blk.downsample[0], blk.downsample[1], mgr.neighbor_table(keys[lvl], keys[lvl], 1, 1), lvl, relu=False)
            return cn(h, blk.conv2, blk.norm2, km3, lvl, relu=True, residual=res)

        ks0 = self.conv0p1s1.kernel_generator.kernel_size
        f = cn(x.F, self.conv0p1s1, self.bn0, mgr.neighbor_table(keys[0], keys[0], ks0, 1), 0)
        skips = [f]
        for lvl in range(1, 5):
            km = mgr.neighbor_table(keys[lvl - 1], keys[lvl], 2, 1)
            f = cn(f, getattr(self, f"conv{lvl}p{2 ** (lvl - 1)}s2"), getattr(self, f"bn{lvl}"), km, lvl)
            for blk in getattr(self, f"block{lvl}"):
                f = block(f, blk, lvl)
            skips.append(f)
        fmaps = [f]
        for j, lvl in enumerate(range(4, 0, -1)):
            plan = mgr.transpose_plan(keys[lvl], keys[lvl - 1], 2, 1)
            f = cn(f, getattr(self, f"convtr{4 + j}p{2 ** lvl}s2"), getattr(self, f"bntr{4 + j}"), plan, lvl - 1, kind=2)
            f = torch.cat([f, skips[lvl - 1]], dim=1)
            for blk in getattr(self, f"block{5 + j}"):
                f = block(f, blk, lvl - 1)
            fmaps.append(f)
        wrap = lambda t, k: ME.SparseTensor(t, coordinate_map_key=k, coordinate_manager=mgr)
        out = wrap(f, keys[0])
        fm = [wrap(t, keys[4 - i]) for i, t in enumerate(fmaps)]
        return (out, fm) if self.out_fpn else out


# ----------------------------------------------------------------------------------------------------
# decoder
# ----------------------------------------------------------------------------------------------------
class _GaussPE(nn.Module):
    """Holds the `pos_enc.gauss_B` buffer [3, 64] (position_embedding.py:72-74)."""

    def __init__(self, d_pos=128, gauss_scale=1.0):
        super().__init__()
        self.register_buffer("gauss_B", torch.empty(3, d_pos // 2).normal_() * gauss_scale)


class _MLP(nn.Module):
    """GenericMLP(use_conv=True, hidden=[d], output_use_activation=True): Conv1d-ReLU-Conv1d-ReLU under `.layers`."""

    def __init__(self, d):
        super().__init__()
        self.layers = nn.Sequential(nn.Conv1d(d, d, 1, bias=True), nn.ReLU(), nn.Conv1d(d, d, 1, bias=True), nn.ReLU())


class CrossAttentionLayer(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.multihead_attn, self.norm = nn.MultiheadAttention(d, h, dropout=0.0), nn.LayerNorm(d)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class SelfAttentionLayer(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.self_attn, self.norm = nn.MultiheadAttention(d, h, dropout=0.0), nn.LayerNorm(d)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class FFNLayer(nn.Module):
    def __init__(self, d, dff):
        super().__init__()
        self.linear1, self.linear2, self.norm = nn.Linear(d, dff), nn.Linear(dff, d), nn.LayerNorm(d)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class _QueryGraphs:
    """CUDA graphs of the query-side work of the decoder (everything that touches only the [B, Q, D] query tensors).

    Between two of the big varlen kernels (mask-logit product / pooling / masked cross-attention over all voxels) the
    reference runs ~45 tiny ops on 25 x 128 tensors per layer (mask3d.py:384-495: out-proj, 3 LayerNorms, self-attention,
    FFN, mask-embed MLP, class head, query projection).  Their shapes depend only on (B, Q, D), never on the scene, so they
    are captured ONCE per batch size into one graph per decoder level and replayed: the host pays one cudaGraphLaunch per
    layer instead of ~45 launches, and the GPU runs them back to back.  Static buffers: queries, query_pos, ctx (written by
    mk_masked_xattn_fwd), qp, embed / its bf16 weight image (read by the mask-logit kernel), cls."""

    def __init__(self, model: "Mask3D", B: int, mode: str, device):
        Q, D = model.num_queries, model.mask_dim
        f32 = dict(dtype=torch.float32, device=device)
        self.queries, self.query_pos = torch.zeros((B, Q, D), **f32), torch.zeros((B, Q, D), **f32)
        self.ctx, self.qp = torch.zeros((B, Q, D), **f32), torch.zeros((B, Q, D), **f32)
        self.emb = torch.zeros((B, Q, D), **f32)
        self.cls = torch.zeros((B, Q, model.num_classes + 1), **f32)
        self.mode = mode
        Qp = ((Q + 15) // 16) * 16
        self.W = torch.zeros((B, D, Qp), **f32)
        self.img = torch.empty(lib.query("mk_conv_weight_image_elems", B, D, Qp, 0), dtype=torch.bfloat16, device=device)
        L = len(model.hlevels)
        fns = [lambda: self._head(model, 0)] + [(lambda i=i: (self._post(model, i), self._head(model, (i + 1) % L))) for i in range(L)]
        saved = lib._stream_cache
        lib._stream_cache = None  # capture runs on torch's capture stream, not on the stream pinned by stream_scope
        try:
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):  # eager warm-up (cuBLAS handles / workspaces) before capture
                for f in fns:
                    f()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            self.graphs = []
            pool = None
            for f in fns:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    f()
                pool = g.pool()
                self.graphs.append(g)
        finally:
            lib._stream_cache = saved
        self.queries.zero_()

    def _head(self, m: "Mask3D", i_next: int):
        """mask_module head (mask3d.py:520-533) + query projection of the next cross-attention (:686-697)."""
        D = m.mask_dim
        qn = m.decoder_norm(self.queries)
        self.emb.copy_(m.mask_embed_head(qn))
        self.cls.copy_(m.class_embed_head(qn))
        if self.mode == "bf16":
            ops.mask_embed_image(self.emb, img=self.img, W=self.W)
        mha = m.cross_attention[0][i_next].multihead_attn
        self.qp.copy_(F.linear(self.queries + self.query_pos, mha.in_proj_weight[:D], mha.in_proj_bias[:D]))

    def _post(self, m: "Mask3D", i: int):
        """out-proj + LN of the cross-attention, self-attention, FFN (mask3d.py:478-495)."""
        self.queries.copy_(m._query_update(i, self.queries, self.query_pos, self.ctx))

    def first(self):
        self.graphs[0].replay()

    def layer(self, i: int):
        self.graphs[1 + i].replay()


class Mask3D(nn.Module):
    """Default reference configuration (mask3d/conf/model/mask3d.yaml): non-parametric queries (FPS + fourier PE), shared
    decoder, post-norm, dropout 0, no level embedding, is_eval=True (keys never subsampled, trainer.py:274)."""

    def __init__(self, backbone: Optional[Res16UNet34C] = None, hidden_dim=128, num_queries=25, num_heads=8, dim_feedforward=1024,
                 num_classes=1, num_decoders=3, hlevels=(0, 1, 2, 3), gauss_scale=1.0):
        super().__init__()
        self.backbone = backbone if backbone is not None else Res16UNet34C()
        self.mask_dim, self.num_queries, self.num_heads = hidden_dim, num_queries, num_heads
        self.num_decoders, self.hlevels, self.num_classes = num_decoders, tuple(hlevels), num_classes
        sizes = PLANES[-5:]
        d = hidden_dim
        self.mask_features_head = _conv(PLANES[7], d, 1, bias=True)
        self.query_projection = _MLP(d)
        self.mask_embed_head = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, d))
        self.class_embed_head = nn.Linear(d, num_classes + 1)
        self.pos_enc = _GaussPE(d, gauss_scale)
        self.pooling = ME.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)
        mk = lambda f: nn.ModuleList([nn.ModuleList([f(i) for i in range(len(self.hlevels))])])
        self.cross_attention = mk(lambda i: CrossAttentionLayer(d, num_heads))
        self.self_attention = mk(lambda i: SelfAttentionLayer(d, num_heads))
        self.ffn_attention = mk(lambda i: FFNLayer(d, dim_feedforward))
        self.lin_squeeze = mk(lambda i: nn.Linear(sizes[self.hlevels[i]], d))
        self.decoder_norm = nn.LayerNorm(d)
        self._derived: Dict[Tuple[int, str], Tuple[int, torch.Tensor]] = {}
        self.use_cuda_graphs = os.environ.get("MK3D_NO_GRAPHS", "0") != "1"   # query-side work of the decoder as CUDA graphs (_QueryGraphs)
        self._qgraphs: Dict[tuple, _QueryGraphs] = {}
        self._side_streams: Dict[torch.device, torch.cuda.Stream] = {}
        self._ev_backbone: Optional[torch.cuda.Event] = None

    def _d(self, p: torch.Tensor, tag: str, fn):
        """cache of derived weight layouts (transposes / slices), invalidated by the parameter's version counter"""
        k = (p.data_ptr(), tag)
        e = self._derived.get(k)
        if e is None or e[0] != p._version:
            e = (p._version, fn(p.detach()).contiguous(), lib.cache_built())
            self._derived[k] = e
        else:
            lib.cache_use(e[2])  # built on another stream (two-lane inference): wait for its producer once
        return e[1]

    def _query_update(self, i: int, queries, query_pos, ctx):
        """Everything of decoder layer i after the attention core: cross-attention out-proj + LN (CrossAttentionLayer.forward_post,
        mask3d.py:686-705), self-attention over the Q queries of each scene (SelfAttentionLayer.forward_post, :624-636: q = k =
        out + query_pos, v = out), FFN (:776-783)."""
        B, Q, D = queries.shape
        H = self.num_heads
        hd = D // H
        ca, sa, ff = self.cross_attention[0][i], self.self_attention[0][i], self.ffn_attention[0][i]
        mha, sam = ca.multihead_attn, sa.self_attn
        out = ca.norm(queries + F.linear(ctx, mha.out_proj.weight, mha.out_proj.bias))
        qk = out + query_pos
        wq, wk_, wv = sam.in_proj_weight[:D], sam.in_proj_weight[D:2 * D], sam.in_proj_weight[2 * D:]
        bq, bk_, bv = sam.in_proj_bias[:D], sam.in_proj_bias[D:2 * D], sam.in_proj_bias[2 * D:]
        q_ = F.linear(qk, wq, bq).view(B, Q, H, hd).transpose(1, 2)
        k_ = F.linear(qk, wk_, bk_).view(B, Q, H, hd).transpose(1, 2)
        v_ = F.linear(out, wv, bv).view(B, Q, H, hd).transpose(1, 2)
        t2 = F.scaled_dot_product_attention(q_, k_, v_).transpose(1, 2).reshape(B, Q, D)
        out = sa.norm(out + F.linear(t2, sam.out_proj.weight, sam.out_proj.bias))
        return ff.norm(out + ff.linear2(F.relu(ff.linear1(out))))

    def _query_graphs(self, B: int, mode: str, device) -> _QueryGraphs:
        key = (B, mode, str(device), self.decoder_norm.weight.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        g = self._qgraphs.get(key)
        if g is None:
            g = self._qgraphs[key] = _QueryGraphs(self, B, mode, device)
        return g

    # ---- coordinate phase of the NEXT scene on a side stream -----------------------------------------------------------------
    @torch.no_grad()
    def prepare(self, coordinates: torch.Tensor, features: torch.Tensor, wait_event: Optional[torch.cuda.Event] = None) -> ME.SparseTensor:
        """Everything of a forward pass that depends only on the coordinates — H2D copy, coordinate hashing, the 4 strided
        maps, 14 kernel maps, tile plans, scene offsets (all the host round trips for data-dependent sizes) — enqueued on a
        side stream.  Called for scene i+1 while scene i's decoder is still running on the main stream, it hides the
        coordinate phase (five host syncs with an otherwise idle GPU, ~1.5 ms at 0.9 M voxels) behind compute.  Returns a
        SparseTensor carrying an event; `forward` makes the main stream wait for it.

        Allocator safety: the side stream only starts after `wait_event` (recorded on the main stream behind the backbone of
        the current scene), i.e. after every main-stream kernel that may still read map tensors of EARLIER scenes whose blocks
        the side stream's pool is about to recycle."""
        dev = next(self.parameters()).device
        side = self.side_stream(dev)
        if wait_event is not None:
            side.wait_event(wait_event)
        with torch.cuda.stream(side):
            c = coordinates.to(dev, non_blocking=True)
            f = features.to(dev, non_blocking=True)
            x = ME.SparseTensor(features=f, coordinates=c)
            mgr = x.coordinate_manager
            keys = mgr.prebuild(x.coordinate_map_key, 5, self.backbone.conv0p1s1.kernel_generator.kernel_size[0])
            mgr.scene_offsets_host(keys[0])
            mgr.scene_select_table(keys[0]).tile_kmask
            ev = torch.cuda.Event()
            ev.record(side)
        x._ready_event = ev
        return x

    def side_stream(self, device=None) -> torch.cuda.Stream:
        """The stream `prepare` runs on (created on first use)."""
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        side = self._side_streams.get(dev)
        if side is None:
            side = self._side_streams[dev] = torch.cuda.Stream(device=dev)
        return side

    @torch.no_grad()
    def forward(self, x: ME.SparseTensor, is_eval: bool = True, prefetch=None, compact: bool = False, aux_outputs: bool = True) -> dict:
        """Fused inference forward.  Returns the fields of the reference's ModelOutput (mask3d.py:28-47).

        `prefetch = (coordinates, features)` of the NEXT scene (host pinned or device tensors): its coordinate phase is
        enqueued on a side stream behind this scene's backbone and the prepared SparseTensor is returned under
        "prefetched" — pass it as `x` of the next call.

        The forward runs on the CURRENT torch stream and keeps its static CUDA-graph buffers per stream, so consecutive scenes may
        alternate between two streams (`with torch.cuda.stream(lane[i % 2]): model(x, prefetch=...)`): the latency-bound
        decoder chain of one scene then overlaps the bandwidth- and tensor-bound backbone of the next (bench.py does this).

        `aux_outputs=False`: the full-resolution mask logits of the 12 intermediate decoder outputs (the reference's `aux_outputs`,
        mask3d.py:405-441, read by the criterion only) are not materialised -- the attention masks of those layers come from the
        pooled mask features either way; `pred_mask_all_decoders` then holds the final prediction only."""
        ev = getattr(x, "_ready_event", None)
        if ev is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            # built on the side stream, consumed here: keep the allocator from recycling the maps under this stream's feet
            # (matters once scenes alternate between two streams, see `forward` docstring)
            x.coordinate_manager.record_stream(cur)
            x.F.record_stream(cur)
            x.C.record_stream(cur)
        with lib.stream_scope():
            out = self._forward(x, aux_outputs=aux_outputs)
            if compact:
                # what the reference's post-processing keeps of the final mask logits (trainer/trainer.py:373-402): the binary masks
                # (bit q of word n = voxel n belongs to query q) and the per-query mask scores, built on the device in one pass --
                # 4 bytes per voxel to copy to the host instead of 4 * Q
                off0 = x.coordinate_manager.scene_offsets(out["backbone_features"].coordinate_map_key)
                B = x.coordinate_manager.number_of_unique_batch_indices()
                out["mask_bits"], out["mask_score"] = ops.mask_compact(out["pred_mask_all_decoders"][-1], off0, B)
        if prefetch is not None:
            out["prefetched"] = self.prepare(prefetch[0], prefetch[1], wait_event=self._ev_backbone)
        return out

    def forward_train(self, x: ME.SparseTensor) -> dict:
        """Autograd-capable forward of the whole model (bf16 tensor-core mode): backbone through `Res16UNet34C.forward_train`,
        1x1 convolutions (mask-feature head, lin_squeeze, K / V in-projections) through the tcgen05 conv forward / dgrad /
        wgrad, mask logits and the masked cross-attention core through their own backward kernels, the [B, Q, 128] query-side
        layers through PyTorch autograd.  Returns what `forward` returns; feed it to `criterion.SetCriterion`."""
        assert ops.get_compute_mode() == "bf16", "forward_train is the bf16 tensor-core path"
        with lib.stream_scope():
            return self._forward(x, train=True)

    def _forward(self, x: ME.SparseTensor, train: bool = False, aux_outputs: bool = True) -> dict:
        mgr = x.coordinate_manager
        mode = ops.get_compute_mode()
        act = torch.bfloat16 if mode == "bf16" else torch.float32
        B, Q, H, D = mgr.number_of_unique_batch_indices(), self.num_queries, self.num_heads, self.mask_dim
        backbone_out, fmaps = self.backbone.forward_train(x) if train else self.backbone.forward_fused(x)

        def lin(xf, W, km, bias, out_dtype):
            """1x1 convolution + bias on the rows of a level; differentiable when training (W: [cin, cout] view of a parameter)"""
            if not train:
                return ops.conv_forward(xf, W, km, bias=bias, mode=mode, out_dtype=out_dtype)
            return _LinFn.apply(xf, W, bias.view(-1), km, out_dtype)

        self._ev_backbone = torch.cuda.Event()
        self._ev_backbone.record(torch.cuda.current_stream())
        keys = [fm.coordinate_map_key for fm in fmaps]  # coarse -> fine (L4 .. L0)
        nl = len(fmaps)
        offs = [mgr.scene_offsets(k) for k in keys]
        gB = self.pos_enc.gauss_B

        # coordinate pyramid (avg-pooled voxel coordinates, mask3d.py:286-299) and per-level fourier PE (:261-278)
        cf = [None] * nl
        cf[nl - 1] = x.C[:, 1:4].float().contiguous()
        pool_km = [None] * nl  # pool_km[i]: level i+1 (finer) -> level i
        for i in range(nl - 2, -1, -1):
            pool_km[i] = mgr.neighbor_table(keys[i + 1], keys[i], 2, 1)
            cf[i] = ops.pool_forward(cf[i + 1], pool_km[i], average=True)
        # the reference builds the PE of every level (mask3d.py:261-278) but only the levels the decoder attends to are ever read
        # (hlevels = L4..L1): the [N0, 128] fp32 PE of the finest level (1.8 GB written at 3.5 M voxels) is never used -- only its
        # coordinate range is, for the query PE
        # (mk_fourier_pe adds the PE onto the key source `src` in the same pass, see the K / V projections below)
        used = set(self.hlevels)
        rng = [ops.scene_minmax(cf[i], offs[i], B) if (i in used or i == nl - 1) else None for i in range(nl)]
        mins0, maxs0 = rng[nl - 1]

        km1 = mgr.neighbor_table(keys[-1], keys[-1], 1, 1)
        mask_features = lin(backbone_out.F, self.mask_features_head.kernel, km1, self.mask_features_head.bias, act)

        # non-parametric queries: FPS over the input voxel coordinates + fourier PE of the sampled points (:308-338)
        fps_idx = ops.furthest_point_sample(cf[nl - 1], offs[-1], B, Q).long()  # scene-local
        sampled = cf[nl - 1][(fps_idx + offs[-1][:B, None].long()).reshape(-1)]  # [B*Q, 3]
        q_off = torch.arange(0, (B + 1) * Q, Q, dtype=torch.int32, device=sampled.device)
        qpos = ops.fourier_pe(sampled, q_off, B, mins0, maxs0, gB, out_dtype=torch.float32)  # [B*Q, D]
        l0, l2 = self.query_projection.layers[0], self.query_projection.layers[2]
        qpos = F.relu(F.linear(F.relu(F.linear(qpos, l0.weight.squeeze(-1), l0.bias)), l2.weight.squeeze(-1), l2.bias))
        query_pos = qpos.view(B, Q, D)
        queries = torch.zeros_like(query_pos)

        pred_class, pred_masks = [], []
        sel0 = mgr.scene_select_table(keys[-1])
        off0_host = mgr.scene_offsets_host(keys[-1]) if train else None

        # Attention masks from POOLED MASK FEATURES.  The reference pools the [N0, Q] mask logits `nl - hl - 1` times with
        # MinkowskiAvgPooling and thresholds the result (mask3d.py:544-553), 12 times per forward, each time starting from the
        # full-resolution logits.  Average pooling is a linear map on rows, so pool^n(F_mask @ E^T) == pool^n(F_mask) @ E^T: the
        # pooled mask features are query independent and are built ONCE per forward (4 pooling launches); a decoder layer then
        # only multiplies the coarse level's [N_l, 128] features with the mask embeddings and thresholds [N_l, Q] -- no pass over
        # the [N0, Q] logits at all (they are still produced at full resolution: all 13 are model outputs).
        mf_lvl = [None] * nl
        mf_lvl[nl - 1] = mask_features.detach()
        for i in range(nl - 2, -1, -1):
            pooled_f = ops.pool_forward(mf_lvl[i + 1], pool_km[i], average=True)  # fp32 accumulation over the existing children
            mf_lvl[i] = ops.cast(pooled_f, act)
        sel_lvl = [mgr.scene_select_table(k) for k in keys]

        def blocked_from(emb, hl, img=None):
            """uint8 [N_hl, Q]: sigmoid(pooled mask logits) < 0.5 on level hl (True = the query may not attend to the voxel)"""
            if img is not None:
                pooled = ops.mask_logits_from_image(mf_lvl[hl], img, sel_lvl[hl], Q)
            else:
                pooled = ops.mask_logits(mf_lvl[hl], offs[hl], B, emb, select=sel_lvl[hl] if (mode != "fp32" and B <= 32) else None)
            return ops.threshold_blocked(pooled)

        # Inference: the full-resolution mask logits of the 12 intermediate decoder outputs are results only (the attention masks
        # come from the pooled mask features, `blocked_from`), so their products are deferred to the end and stacked six outputs per
        # pass over the mask features (ops.mask_logits_stacked: same values, a third of the HBM traffic of 12 separate passes)
        defer_group = int(os.environ.get("MK3D_DEFER_GROUP", "6"))  # tuning switch: 0 = one product per decoder output
        defer = (not train) and aux_outputs and mode in ("bf16", "tf32") and sel0 is not None and sel0.K <= 32 and defer_group > 0
        deferred = []

        def mask_module(queries, hl, want_mask):
            qn = self.decoder_norm(queries)
            emb = self.mask_embed_head(qn)
            cls = self.class_embed_head(qn)
            if train:
                logits = ops._MaskLogitsFn.apply(mask_features, emb, offs[-1], B, sel0, off0_host)
            elif want_mask and (not aux_outputs or defer):
                logits = None
                if defer:
                    deferred.append(emb.detach().clone())
            else:
                logits = ops.mask_logits(mask_features, offs[-1], B, emb, select=sel0)  # [N0, Q] fp32
            blocked = blocked_from(emb.detach(), hl) if want_mask else None  # the attention mask carries no gradient (mask3d.py:549)
            return cls, logits, blocked

        # keys / values of every level do not depend on the queries and the decoder weights are shared between the
        # `num_decoders` repetitions (shared_decoder: true, conf/model/mask3d.yaml:28): project them once per level instead of
        # once per layer (the reference recomputes them 3x, mask3d.py:467-478)
        kv = {}
        for i, hl in enumerate(self.hlevels):
            ca, sq = self.cross_attention[0][i], self.lin_squeeze[0][i]
            kml = mgr.neighbor_table(keys[hl], keys[hl], 1, 1)
            mha = ca.multihead_attn
            if train:  # views of the parameters: gradients flow back through them
                Wsq, Wk, Wv = sq.weight.t(), mha.in_proj_weight[D:2 * D].t(), mha.in_proj_weight[2 * D:].t()
            else:
                Wsq = self._d(sq.weight, "T", lambda w: w.t())                                  # [C_l, D]
                Wk = self._d(mha.in_proj_weight, "k", lambda w: w[D:2 * D].t())                   # [D, D] (in, out)
                Wv = self._d(mha.in_proj_weight, "v", lambda w: w[2 * D:].t())
            # inference: `src` and `src + pos` directly in the activation dtype the K / V projections read (bf16 on the tensor-core
            # path): no fp32 round trip + cast pass over [N_l, 128]; training keeps fp32 (autograd through the PE add)
            sdt = torch.float32 if train else act
            src = lin(fmaps[hl].F, Wsq, kml, sq.bias, sdt)
            kin = ops.fourier_pe(cf[hl], offs[hl], B, rng[hl][0], rng[hl][1], gB, base=src, out_dtype=sdt)  # src + pos
            kv[i] = (lin(kin, Wk, kml, mha.in_proj_bias[D:2 * D], act), lin(src, Wv, kml, mha.in_proj_bias[2 * D:], act))
        max_rows = None
        if train:
            max_rows = []
            for hl in self.hlevels:
                oh = mgr.scene_offsets_host(keys[hl])
                max_rows.append(max(int(oh[b + 1]) - int(oh[b]) for b in range(B)))

        if self.use_cuda_graphs and not train:
            # query-side work replayed from CUDA graphs; the varlen voxel kernels in between run eagerly
            qg = self._query_graphs(B, mode, query_pos.device)
            qg.query_pos.copy_(query_pos)
            qg.queries.zero_()
            qg.first()

            def masks_from_graph(hl, want_mask):
                cls = qg.cls.clone()
                if want_mask and (not aux_outputs or defer):
                    logits = None
                    if defer:
                        deferred.append(qg.emb.clone())
                elif mode == "bf16" and sel0.K <= 32:  # more than 32 scenes: the per-tile offset mask is 32 bits -> CUDA-core kernel
                    logits = ops.mask_logits_from_image(mask_features, qg.img, sel0, Q)
                elif mode == "bf16":
                    logits = ops.mask_logits(mask_features, offs[-1], B, qg.emb, select=None)
                else:  # tf32: the tensor-core kernel on fp32 features; fp32: the exact CUDA-core checker
                    logits = ops.mask_logits(mask_features, offs[-1], B, qg.emb, select=sel0 if mode == "tf32" else None)
                blocked = blocked_from(qg.emb, hl, img=qg.img if (mode == "bf16" and sel0.K <= 32) else None) if want_mask else None
                return cls, logits, blocked

            for _ in range(self.num_decoders):
                for i, hl in enumerate(self.hlevels):
                    cls, logits, blocked = masks_from_graph(hl, True)
                    kp, vp = kv[i]
                    ops.masked_cross_attention(qg.qp, kp, vp, blocked, offs[hl], H, out=qg.ctx)
                    qg.layer(i)
                    pred_class.append(cls)
                    if logits is not None:
                        pred_masks.append(logits)
            cls, logits, _ = masks_from_graph(0, False)
            queries = qg.queries.clone()
        else:
            for _ in range(self.num_decoders):
                for i, hl in enumerate(self.hlevels):
                    cls, logits, blocked = mask_module(queries, hl, True)
                    mha = self.cross_attention[0][i].multihead_attn
                    kp, vp = kv[i]
                    qp = F.linear(queries + query_pos, mha.in_proj_weight[:D], mha.in_proj_bias[:D])
                    ctx = ops.masked_cross_attention(qp, kp, vp, blocked, offs[hl], H, max_scene_rows=max_rows[i] if train else None)
                    queries = self._query_update(i, queries, query_pos, ctx)
                    pred_class.append(cls)
                    if logits is not None:
                        pred_masks.append(logits)
            cls, logits, _ = mask_module(queries, 0, False)
        pred_class.append(cls)
        if deferred:
            pred_masks = ops.mask_logits_stacked(mask_features, offs[-1], B, deferred, sel0, group=defer_group) + pred_masks
        pred_masks.append(logits)
        off0 = mgr.scene_offsets_host(keys[-1])
        split = lambda t: [t[off0[b]:off0[b + 1]] for b in range(B)]
        return dict(pred_class=pred_class[-1], pred_class_all_decoders=pred_class, pred_mask=split(pred_masks[-1]),
                    pred_mask_all_decoders=pred_masks, sampled_coords=sampled.view(B, Q, 3), backbone_features=backbone_out,
                    queries=queries, fps_idx=fps_idx)
