"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

Functional CPU restatement (plain PyTorch on top of oracle/me_oracle.py) of the reference model
forward on the hot path, keyed by the reference's own state-dict names so the same weights can be
fed to the reference modules, to this oracle and to the CUDA path:

  * `res16unet34c_forward`   follows mask3d/models/res16unet.py:241-314 (composition),
    :327-329,390-391 (LAYERS / PLANES of Res16UNet34C), mask3d/models/resnet.py:99-152
    (`_make_layer`, 1x1 downsample branch) and mask3d/models/modules/resnet_block.py:52-68 (BasicBlock);
  * `mask3d_forward`         follows mask3d/models/mask3d.py:280-518 (forward), :520-561 (mask_module),
    :261-278 (get_pos_encs), :686-705 / :624-636 / :780-784 (post-norm attention / FFN layers) and
    mask3d/models/position_embedding.py:13-43,127-156 (fourier embedding with normalisation).
    Attention is evaluated with torch.nn.functional.multi_head_attention_forward — the very function the
    reference's nn.MultiheadAttention calls — on tensors padded exactly as mask3d.py:400-465 pads them.

Pinned by tests/golden/*.npz, which were produced by running the reference's UNMODIFIED modules
(tests/golden/make_golden.py); tests/test_oracle.py replays them through this file.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import me_oracle as ME
from .fps_oracle import furthest_point_sample

LAYERS = (2, 3, 4, 6, 2, 2, 2, 2)  # Res16UNet34 (res16unet.py:327-329)
PLANES = (32, 64, 128, 256, 256, 128, 96, 96)  # Res16UNet34C (res16unet.py:390-391)
INIT_DIM = 32


# ----------------------------------------------------------------------------------------------
# backbone
# ----------------------------------------------------------------------------------------------


def _conv(sd, name, x: ME.SparseTensor, ksize, stride=1, transpose=False) -> ME.SparseTensor:
    W = sd[name + ".kernel"]
    mgr = x.coordinate_manager
    if W.dim() == 2:
        out_key, Fo = x.coordinate_map_key, x.F @ W
    elif transpose:
        ts = tuple(a // stride for a in x.coordinate_map_key.tensor_stride)
        out_key = ME.CoordinateMapKey(ts, "")
        nbr = mgr.neighbor_table(out_key, x.coordinate_map_key, ksize, 1)
        Fo = ME.conv_transpose_from_table(x.F, W, nbr, mgr.size(out_key))
    else:
        out_key = mgr.stride(x.coordinate_map_key, stride)
        nbr = mgr.neighbor_table(x.coordinate_map_key, out_key, ksize, 1)
        Fo = ME.conv_from_table(x.F, W, nbr, mgr.size(out_key))
    if name + ".bias" in sd:
        Fo = Fo + sd[name + ".bias"]
    return ME.SparseTensor(Fo, coordinate_map_key=out_key, coordinate_manager=mgr)


def _inorm(sd, name, x: ME.SparseTensor, relu=False, residual=None) -> ME.SparseTensor:
    rows = x.coordinate_manager.batch_rows(x.coordinate_map_key)
    Fo = ME.instance_norm(x.F, rows, sd[name + ".weight"], sd[name + ".bias"])
    if residual is not None:
        Fo = Fo + residual.F
    if relu:
        Fo = torch.relu(Fo)
    return ME.SparseTensor(Fo, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


def _basic_block(sd, name, x):
    out = _conv(sd, name + ".conv1", x, 3)
    out = _inorm(sd, name + ".norm1", out, relu=True)
    out = _conv(sd, name + ".conv2", out, 3)
    residual = x
    if name + ".downsample.0.kernel" in sd:
        residual = _inorm(sd, name + ".downsample.1", _conv(sd, name + ".downsample.0", x, 1))
    return _inorm(sd, name + ".norm2", out, relu=True, residual=residual)


def _stage(sd, name, x, n):
    for i in range(n):
        x = _basic_block(sd, f"{name}.{i}", x)
    return x


def res16unet34c_forward(sd: Dict[str, torch.Tensor], x: ME.SparseTensor, prefix: str = ""):
    """Returns (out, feature_maps) like Res16UNetBase.forward with out_fpn=True."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    ks0 = round(sd["conv0p1s1.kernel"].shape[0] ** (1 / 3))
    out_p1 = _inorm(sd, "bn0", _conv(sd, "conv0p1s1", x, ks0), relu=True)
    skips = [out_p1]
    out = out_p1
    for lvl in range(1, 5):
        out = _inorm(sd, f"bn{lvl}", _conv(sd, f"conv{lvl}p{2 ** (lvl - 1)}s2", out, 2, stride=2), relu=True)
        out = _stage(sd, f"block{lvl}", out, LAYERS[lvl - 1])
        skips.append(out)
    fmaps = [out]
    for j, lvl in enumerate(range(4, 0, -1)):
        out = _inorm(sd, f"bntr{4 + j}", _conv(sd, f"convtr{4 + j}p{2 ** lvl}s2", out, 2, stride=2, transpose=True), relu=True)
        out = ME.cat(out, skips[lvl - 1])
        out = _stage(sd, f"block{5 + j}", out, LAYERS[4 + j])
        fmaps.append(out)
    return out, fmaps


# ----------------------------------------------------------------------------------------------
# decoder
# ----------------------------------------------------------------------------------------------


def fourier_pos_enc(xyz: torch.Tensor, gauss_B: torch.Tensor, mins: torch.Tensor, maxs: torch.Tensor) -> torch.Tensor:
    """position_embedding.py:127-156 with normalize=True: [B,N,3] -> [B, d_pos, N]."""
    x = (xyz - mins[:, None, :]) / (maxs[:, None, :] - mins[:, None, :])
    x = x * (2 * np.pi)
    proj = torch.mm(x.reshape(-1, 3), gauss_B).view(xyz.shape[0], xyz.shape[1], -1)
    return torch.cat([proj.sin(), proj.cos()], dim=2).permute(0, 2, 1)


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"])


def _mha(sd, name, q, k, v, nheads, attn_mask=None):
    out, _ = F.multi_head_attention_forward(
        q, k, v, q.shape[-1], nheads,
        sd[name + ".in_proj_weight"], sd[name + ".in_proj_bias"], None, None, False, 0.0,
        sd[name + ".out_proj.weight"], sd[name + ".out_proj.bias"],
        training=False, key_padding_mask=None, need_weights=True, attn_mask=attn_mask,
    )
    return out


def mask_module(sd, queries, mask_features: ME.SparseTensor, num_pooling_steps: int, ret_attn_mask=True):
    """mask3d.py:520-561."""
    qf = _ln(sd, "decoder_norm", queries)
    me = F.linear(F.relu(F.linear(qf, sd["mask_embed_head.0.weight"], sd["mask_embed_head.0.bias"])),
                  sd["mask_embed_head.2.weight"], sd["mask_embed_head.2.bias"])
    cls = F.linear(qf, sd["class_embed_head.weight"], sd["class_embed_head.bias"])
    dec = mask_features.decomposed_features
    masks = torch.cat([dec[i] @ me[i].T for i in range(len(dec))])
    om = ME.SparseTensor(masks, coordinate_map_key=mask_features.coordinate_map_key,
                         coordinate_manager=mask_features.coordinate_manager)
    if not ret_attn_mask:
        return cls, om.decomposed_features, None
    pool = ME.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)
    am = om
    for _ in range(num_pooling_steps):
        am = pool(am.float())
    am = ME.SparseTensor(am.F.detach().sigmoid() < 0.5, coordinate_map_key=am.coordinate_map_key,
                         coordinate_manager=am.coordinate_manager)
    return cls, om.decomposed_features, am


def mask3d_forward(sd: Dict[str, torch.Tensor], x: ME.SparseTensor, num_queries=25, num_heads=8, num_decoders=3,
                   hlevels=(0, 1, 2, 3)):
    """Mask3D.forward(is_eval=True) with the default config (conf/model/mask3d.yaml): non-parametric queries
    (FPS + fourier PE), shared decoder, post-norm, no level embedding, dropout 0."""
    bsd = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    backbone_out, fmaps = res16unet34c_forward(bsd, x)
    mgr = x.coordinate_manager
    pool = ME.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)

    coordinates = ME.SparseTensor(x.C[:, 1:4].float(), coordinate_map_key=fmaps[-1].coordinate_map_key, coordinate_manager=mgr)
    clist = [coordinates]
    for _ in range(len(fmaps) - 1):
        clist.append(pool(clist[-1]))
    clist.reverse()

    gB = sd["pos_enc.gauss_B"]
    pos_encs = []
    for lvl in range(len(clist)):
        per = []
        for c in clist[lvl].decomposed_features:
            mn, mx = c.min(0)[0][None], c.max(0)[0][None]
            per.append(fourier_pos_enc(c[None].float(), gB, mn, mx).squeeze(0).permute(1, 0))
        pos_encs.append(per)

    mask_features = _conv(sd, "mask_features_head", backbone_out, 1)

    dc = x.decomposed_coordinates
    cf = coordinates.decomposed_features
    fps_idx = [torch.from_numpy(furthest_point_sample(dc[b].float().numpy(), num_queries).astype(np.int64)) for b in range(len(dc))]
    sampled = torch.stack([cf[b][fps_idx[b]] for b in range(len(dc))])
    mins = torch.stack([c.min(0)[0] for c in cf])
    maxs = torch.stack([c.max(0)[0] for c in cf])
    qp = fourier_pos_enc(sampled.float(), gB, mins, maxs)  # [B, 128, Q]
    qp = F.relu(F.conv1d(F.relu(F.conv1d(qp, sd["query_projection.layers.0.weight"], sd["query_projection.layers.0.bias"])),
                         sd["query_projection.layers.2.weight"], sd["query_projection.layers.2.bias"]))
    queries = torch.zeros_like(qp).permute(0, 2, 1)  # [B, Q, 128]
    query_pos = qp.permute(2, 0, 1)  # [Q, B, 128]

    pred_class, pred_masks, attn_masks = [], [], []
    for _ in range(num_decoders):
        for i, hl in enumerate(hlevels):
            cls, om, am = mask_module(sd, queries, mask_features, len(fmaps) - hl - 1, True)
            feats = fmaps[hl].decomposed_features
            dam = am.decomposed_features
            Kmax = max(f.shape[0] for f in feats)
            idxs, pads = [], []
            for f in feats:  # mask3d.py:410-441 (eval: never subsampled)
                idx = torch.zeros(Kmax, dtype=torch.long)
                pad = torch.ones(Kmax, dtype=torch.bool)
                idx[: f.shape[0]] = torch.arange(f.shape[0])
                pad[: f.shape[0]] = False
                idxs.append(idx)
                pads.append(pad)
            aux = torch.stack([feats[b][idxs[b]] for b in range(len(feats))])
            bam = torch.stack([dam[b][idxs[b]] for b in range(len(feats))])  # [B, K, Q]
            bam.permute(0, 2, 1)[bam.sum(1) == Kmax] = False  # mask3d.py:456
            bam = torch.logical_or(bam, torch.stack(pads)[..., None])
            pe = torch.stack([pos_encs[hl][b][idxs[b]] for b in range(len(feats))])
            src = F.linear(aux.permute(1, 0, 2), sd[f"lin_squeeze.0.{i}.weight"], sd[f"lin_squeeze.0.{i}.bias"])
            tgt = queries.permute(1, 0, 2)
            ca = f"cross_attention.0.{i}"
            t2 = _mha(sd, ca + ".multihead_attn", tgt + query_pos, src + pe.permute(1, 0, 2), src, num_heads,
                      attn_mask=bam.repeat_interleave(num_heads, dim=0).permute(0, 2, 1))
            out = _ln(sd, ca + ".norm", tgt + t2)
            sa = f"self_attention.0.{i}"
            t2 = _mha(sd, sa + ".self_attn", out + query_pos, out + query_pos, out, num_heads)
            out = _ln(sd, sa + ".norm", out + t2)
            ff = f"ffn_attention.0.{i}"
            t2 = F.linear(F.relu(F.linear(out, sd[ff + ".linear1.weight"], sd[ff + ".linear1.bias"])),
                          sd[ff + ".linear2.weight"], sd[ff + ".linear2.bias"])
            queries = _ln(sd, ff + ".norm", out + t2).permute(1, 0, 2)
            pred_class.append(cls)
            pred_masks.append(om)
            attn_masks.append(am.F)
    cls, om, _ = mask_module(sd, queries, mask_features, 0, False)
    pred_class.append(cls)
    pred_masks.append(om)
    return dict(pred_class=pred_class, pred_masks=pred_masks, attn_masks=attn_masks, queries=queries,
                sampled_coords=sampled, backbone_out=backbone_out, feature_maps=fmaps, mask_features=mask_features,
                fps_idx=fps_idx)
