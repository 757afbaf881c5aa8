"""The MinkowskiEngine 0.5.4 Python surface that mask3d/models consumes, backed by libmk3d.so.

Same names, argument meaning, state-dict layouts and error behaviour as ME, so that the reference's
`mask3d/models/*.py`, `mask3d/datasets/utils.py` and the Lightning trainer run unchanged when
`room_instance_segmentation_mask3d_b200/compat` is on sys.path (it exposes this module as `MinkowskiEngine`).
Module boundaries carry fp32 features like ME does; in "bf16" compute mode convolutions run on the tcgen05
path (bf16 operands, fp32 accumulate) and still hand back fp32 rows.  There is no CPU path: constructing a
SparseTensor without CUDA raises.
"""
from __future__ import annotations

from enum import Enum
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import lib, ops
from .engine import CoordinateManager, CoordinateMapKey, KernelMap, _tuple3

__all__ = [
    "SparseTensor", "CoordinateManager", "CoordinateMapKey", "RegionType", "SparseTensorQuantizationMode", "KernelGenerator",
    "MinkowskiNetwork", "MinkowskiConvolution", "MinkowskiConvolutionTranspose", "MinkowskiBatchNorm", "MinkowskiInstanceNorm",
    "MinkowskiReLU", "MinkowskiSigmoid", "MinkowskiLinear", "MinkowskiAvgPooling", "MinkowskiSumPooling", "MinkowskiAvgUnpooling",
    "MinkowskiPoolingTranspose", "MinkowskiGlobalPooling", "MinkowskiBroadcastMultiplication", "cat", "sparse_collate", "sparse_quantize",
    "set_compute_mode", "get_compute_mode",
]

set_compute_mode = ops.set_compute_mode
get_compute_mode = ops.get_compute_mode


class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class SparseTensorQuantizationMode(Enum):
    RANDOM_SUBSAMPLE = 0
    UNWEIGHTED_AVERAGE = 1
    UNWEIGHTED_SUM = 2
    NO_QUANTIZATION = 3
    MAX_POOL = 4


# ----------------------------------------------------------------------------------------------------
# SparseTensor
# ----------------------------------------------------------------------------------------------------
class SparseTensor:
    """ME.SparseTensor (call sites: mask3d/datasets/utils.py:250; mask3d/models/mask3d.py:287,538,549)."""

    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None, coordinate_manager=None,
                 quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE, allocator_type=None, minkowski_algorithm=None,
                 requires_grad=None, device=None):
        lib.require_cuda()
        assert isinstance(features, torch.Tensor), "features must be a torch.Tensor"
        assert features.dim() == 2, f"features must be [N, C], got {tuple(features.shape)}"
        if coordinate_map_key is None:
            assert coordinates is not None, "either coordinates or coordinate_map_key must be given"
            dev = torch.device(device) if device is not None else (features.device if features.is_cuda else torch.device("cuda"))
            if dev.type != "cuda":
                raise RuntimeError("the B200 build has no CPU backend (device must be cuda)")
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager(D=coordinates.shape[1] - 1, device=dev)
            assert coordinates.shape[0] == features.shape[0], "coordinates and features must have the same number of rows"
            coordinate_map_key, (uniq, inv) = coordinate_manager.insert_and_map(coordinates, tensor_stride)
            self.unique_index, self.inverse_mapping = uniq, inv
            features = features.to(dev)
            if uniq.shape[0] != features.shape[0]:
                features = features[uniq]  # RANDOM_SUBSAMPLE: the first occurrence of each voxel
        else:
            assert coordinate_manager is not None, "coordinate_map_key needs its coordinate_manager"
            if coordinate_manager.size(coordinate_map_key) != features.shape[0]:
                raise RuntimeError(f"feature rows ({features.shape[0]}) != coordinate map size ({coordinate_manager.size(coordinate_map_key)})")
            features = features.to(coordinate_manager.device)
        if requires_grad is not None:
            features.requires_grad_(requires_grad)
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self._manager = coordinate_manager

    @property
    def F(self):
        return self._F

    features = F

    @property
    def C(self):
        return self._manager.get_coordinates(self.coordinate_map_key)

    coordinates = C

    @property
    def coordinate_manager(self):
        return self._manager

    @property
    def tensor_stride(self):
        return list(self.coordinate_map_key.tensor_stride)

    @property
    def device(self):
        return self._F.device

    @property
    def dtype(self):
        return self._F.dtype

    @property
    def shape(self):
        return self._F.shape

    @property
    def D(self):
        return self._manager.D

    def size(self, *a):
        return self._F.size(*a)

    def __len__(self):
        return self._F.shape[0]

    def _offsets(self) -> List[int]:
        return self._manager.scene_offsets_host(self.coordinate_map_key)

    @property
    def decomposed_features(self):
        off = self._offsets()
        return [self._F[off[b]:off[b + 1]] for b in range(len(off) - 1)]

    @property
    def decomposed_coordinates(self):
        off, C = self._offsets(), self.C
        return [C[off[b]:off[b + 1], 1:] for b in range(len(off) - 1)]

    def _like(self, F):
        return SparseTensor(F, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self._manager)

    def float(self):
        return self._like(self._F.float())

    def detach(self):
        return self._like(self._F.detach())

    def _same_map(self, other):
        if self._manager is not other._manager or self.coordinate_map_key != other.coordinate_map_key:
            raise RuntimeError("SparseTensors live on different coordinate maps")

    def __iadd__(self, other):
        self._same_map(other)
        self._F = self._F + other._F
        return self

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            self._same_map(other)
            other = other._F
        return self._like(self._F + other)

    def __repr__(self):
        return f"SparseTensor(F={tuple(self._F.shape)} {self._F.dtype}, key={self.coordinate_map_key}, device={self._F.device})"


def cat(*tensors):
    """MinkowskiOps.cat (mask3d/models/res16unet.py:276-306): same map required."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    t0 = tensors[0]
    for t in tensors[1:]:
        t0._same_map(t)
    return t0._like(torch.cat([t.F for t in tensors], dim=1))


# ----------------------------------------------------------------------------------------------------
# autograd Functions (forward: mode-dependent kernels; backward: deterministic fp32 CUDA-core kernels)
# ----------------------------------------------------------------------------------------------------
def _identity_table(plan: KernelMap) -> torch.Tensor:
    ks = torch.arange(plan.K, device=plan.koff.device, dtype=torch.int32)[:, None]
    return torch.where(plan.koff[None, :] == ks, plan.parent[None, :], torch.full_like(plan.parent[None, :], -1)).contiguous()


class _ConvFn(torch.autograd.Function):
    """kind: 0 = same-map conv (odd kernel, stride 1); 1 = strided conv fine->coarse; 2 = transposed conv coarse->fine."""

    @staticmethod
    def forward(ctx, x, W, km_fwd, kind, n_in, aux):
        ctx.km, ctx.kind, ctx.n_in, ctx.aux = km_fwd, kind, n_in, aux
        ctx.save_for_backward(x, W)
        mode = ops.get_compute_mode()
        if kind == 2 and mode == "fp32":
            out = ops.conv_transpose_forward_fp32(x, W, km_fwd)
        else:
            out = ops.conv_forward(x, W, km_fwd, mode=mode, out_dtype=torch.float32)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W = ctx.saved_tensors
        km, kind, n_in = ctx.km, ctx.kind, ctx.n_in
        W3 = (W if W.dim() == 3 else W.unsqueeze(0)).detach().float().contiguous()
        K, cin, cout = W3.shape
        dx = dW = None
        if ops.get_compute_mode() == "bf16" and K <= 32 and cin % 16 == 0 and cout % 16 == 0 and cout <= 256:
            # tensor-core backward: dgrad = the forward kernel over the transposed table with W^T, wgrad = tcgen05 kernel over
            # the forward's own table (x rows gathered through nbr, dy rows through the plan's slot -> row map)
            db = ops.cast(dout.contiguous(), torch.bfloat16)
            if ctx.needs_input_grad[0]:
                if kind == 0:
                    dx = ops.conv_dgrad_bf16(db, W3, km, kidx=ops.reversed_offsets(K, x.device))
                else:
                    dx = ops.conv_dgrad_bf16(db, W3, ctx.aux)  # kind 1: transpose plan; kind 2: fine -> coarse table
            if ctx.needs_input_grad[1]:
                dW = ops.conv_wgrad_bf16(x.detach(), db, km, K, cin, cout).reshape(W.shape)
            return dx, dW, None, None, None, None
        dout = dout.contiguous().float()
        xf = x.detach().float().contiguous()
        if kind == 0:
            nbr_t, kidx, nbr_w = km.nbr, ops.reversed_offsets(K, x.device), km.nbr
        elif kind == 1:
            tplan = ctx.aux  # transpose plan of the same pair of maps (parent / koff per fine row)
            nbr_t, kidx, nbr_w = _identity_table(tplan), None, km.nbr
        else:
            down = ctx.aux  # KernelMap fine -> coarse (nbr_down[k, coarse] = fine)
            nbr_t, kidx, nbr_w = down.nbr, None, _identity_table(km)
        if ctx.needs_input_grad[0]:
            dx = torch.empty((n_in, cin), dtype=torch.float32, device=x.device)
            lib.call("mk_conv_dgrad_f32", lib.p(dout), cout, lib.p(W3), cin, lib.p(nbr_t), K, n_in, lib.p(kidx), lib.p(dx), lib.stream())
        if ctx.needs_input_grad[1] and cin == 1 and cout <= 32 and K <= 128 and kind == 0:
            dW = torch.empty_like(W3)
            wsb = lib.query("mk_conv_stem_wgrad_workspace_bytes", dout.shape[0], K, cout)
            ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
            lib.call("mk_conv_stem_wgrad_f32", lib.p(xf), lib.p(dout), cout, lib.p(nbr_w), K, dout.shape[0], lib.p(dW), lib.p(ws), wsb,
                     lib.stream())
            dW = dW.reshape(W.shape)
        elif ctx.needs_input_grad[1]:
            dW = torch.empty_like(W3)
            lib.call("mk_conv_wgrad_f32", lib.p(xf), cin, lib.p(dout), cout, lib.p(nbr_w), K, dout.shape[0], lib.p(dW), lib.stream())
            dW = dW.reshape(W.shape)
        return dx, dW, None, None, None, None


class _InstanceNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, offsets, B, eps):
        x = x.contiguous()
        mean, rstd = ops.instance_norm_stats(x, offsets, B, eps)
        ctx.save_for_backward(x, weight, mean, rstd, offsets)
        ctx.B = B
        return ops.norm_apply(x, offsets, B, mean, rstd, weight, bias, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd, offsets = ctx.saved_tensors
        n, C = x.shape
        xf, dy = x.float().contiguous(), dy.contiguous().float()
        if C % 4 == 0:
            dx, _, dg, db = ops.instnorm_relu_bwd(dy, None, xf, offsets, ctx.B, mean, rstd, weight)
        else:  # odd widths: scalar kernels
            dx = torch.empty_like(xf)
            dg = torch.empty(C, dtype=torch.float32, device=x.device)
            db = torch.empty(C, dtype=torch.float32, device=x.device)
            wsb = lib.query("mk_instnorm_workspace_bytes", n, C, ctx.B)
            ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
            lib.call("mk_instnorm_bwd", lib.p(xf), lib.p(dy), n, C, lib.p(offsets), ctx.B, lib.p(mean), lib.p(rstd),
                     lib.p(weight.detach().reshape(-1).float().contiguous()), lib.p(dx), lib.p(dg), lib.p(db), lib.p(ws), wsb, lib.stream())
        return dx, dg.reshape(weight.shape), db.reshape(weight.shape), None, None, None


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, km, average, n_in):
        ctx.km, ctx.average, ctx.n_in = km, average, n_in
        return ops.pool_forward(x, km, average)

    @staticmethod
    def backward(ctx, dout):
        return ops.pool_backward(dout, ctx.km, ctx.n_in, ctx.average), None, None, None


# ----------------------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------------------
class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose: bool = False, region_type=RegionType.HYPER_CUBE,
                 region_offsets=None, expand_coordinates: bool = False, axis_types=None, dimension=-1):
        assert dimension > 0, "dimension must be a positive integer"
        self.dimension = dimension
        self.kernel_size = list(_tuple3(kernel_size, dimension))
        self.kernel_stride = list(_tuple3(stride, dimension))
        self.kernel_dilation = list(_tuple3(dilation, dimension))
        if region_type not in (RegionType.HYPER_CUBE, 0):
            raise NotImplementedError("only RegionType.HYPER_CUBE is on the hot path (mask3d/models/modules/common.py:61-70,147)")
        self.region_type, self.region_offsets, self.axis_types = region_type, region_offsets, axis_types
        self.is_transpose, self.expand_coordinates = is_transpose, expand_coordinates
        self.kernel_volume = int(np.prod(self.kernel_size))
        self.requires_strided_coordinates = any(s > 1 for s in self.kernel_stride)


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class _ConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator, is_transpose, dimension):
        super().__init__()
        assert dimension is not None and dimension > 0, "dimension must be a positive integer"
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, is_transpose=is_transpose, dimension=dimension)
        self.kernel_generator = kernel_generator
        self.in_channels, self.out_channels, self.dimension, self.is_transpose = in_channels, out_channels, dimension, is_transpose
        kv = kernel_generator.kernel_volume
        self.use_mm = kv == 1 and not kernel_generator.requires_strided_coordinates
        shape = (in_channels, out_channels) if self.use_mm else (kv, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters(is_transpose)

    def reset_parameters(self, is_transpose=False):
        with torch.no_grad():
            n = (self.out_channels if is_transpose else self.in_channels) * self.kernel_generator.kernel_volume
            stdv = 1.0 / float(np.sqrt(n))
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def extra_repr(self):
        kg = self.kernel_generator
        return f"in={self.in_channels}, out={self.out_channels}, kernel_size={kg.kernel_size}, stride={kg.kernel_stride}, dilation={kg.kernel_dilation}"


class MinkowskiConvolution(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator, False, dimension)

    def forward(self, x: SparseTensor, coordinates=None) -> SparseTensor:
        mgr, kg = x.coordinate_manager, self.kernel_generator
        in_key = x.coordinate_map_key
        if self.use_mm:
            out_key = in_key
            km = mgr.neighbor_table(in_key, in_key, 1, 1)
            F = _ConvFn.apply(x.F, self.kernel, km, 0, x.F.shape[0], None)
        else:
            out_key = mgr.stride(in_key, kg.kernel_stride)
            km = mgr.neighbor_table(in_key, out_key, kg.kernel_size, kg.kernel_dilation)
            if out_key == in_key:
                assert all(k % 2 == 1 for k in kg.kernel_size), "even kernels need stride > 1 on this path"
                F = _ConvFn.apply(x.F, self.kernel, km, 0, x.F.shape[0], None)
            else:
                # backward of a strided conv needs the (offset, parent) map: raises NotImplementedError unless kernel == stride
                aux = mgr.transpose_plan(out_key, in_key, kg.kernel_size, kg.kernel_dilation) if (torch.is_grad_enabled() and x.F.requires_grad) else None
                F = _ConvFn.apply(x.F, self.kernel, km, 1, x.F.shape[0], aux)
        if self.bias is not None:
            F = F + self.bias
        return SparseTensor(F, coordinate_map_key=out_key, coordinate_manager=mgr)


class MinkowskiConvolutionTranspose(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator, True, dimension)

    def forward(self, x: SparseTensor, coordinates=None) -> SparseTensor:
        mgr, kg = x.coordinate_manager, self.kernel_generator
        in_key = x.coordinate_map_key
        if self.use_mm:
            km = mgr.neighbor_table(in_key, in_key, 1, 1)
            out_key, F = in_key, _ConvFn.apply(x.F, self.kernel, km, 0, x.F.shape[0], None)
        else:
            ts = tuple(a // b for a, b in zip(in_key.tensor_stride, kg.kernel_stride))
            out_key = CoordinateMapKey(ts, "")
            if not mgr.exists(out_key):
                raise RuntimeError(f"transposed conv needs the encoder's coordinate map {out_key} (generate_new_coords=False)")
            plan = mgr.transpose_plan(in_key, out_key, kg.kernel_size, kg.kernel_dilation)
            down = mgr.neighbor_table(out_key, in_key, kg.kernel_size, kg.kernel_dilation)
            F = _ConvFn.apply(x.F, self.kernel, plan, 2, x.F.shape[0], down)
        if self.bias is not None:
            F = F + self.bias
        return SparseTensor(F, coordinate_map_key=out_key, coordinate_manager=mgr)


class _PoolBase(nn.Module):
    AVERAGE = True

    def __init__(self, kernel_size=-1, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, dimension=dimension)
        self.kernel_generator, self.dimension = kernel_generator, dimension

    def forward(self, x: SparseTensor, coordinates=None) -> SparseTensor:
        mgr, kg = x.coordinate_manager, self.kernel_generator
        out_key = mgr.stride(x.coordinate_map_key, kg.kernel_stride)
        km = mgr.neighbor_table(x.coordinate_map_key, out_key, kg.kernel_size, kg.kernel_dilation)
        if torch.is_grad_enabled() and x.F.requires_grad and (list(kg.kernel_size) != list(kg.kernel_stride) or any(d != 1 for d in kg.kernel_dilation)):
            # mk_pool_bwd writes each input row's gradient once: only right when every input row has one parent
            raise NotImplementedError("pooling backward needs kernel_size == stride and dilation 1 (non-overlapping windows)")
        F = _PoolFn.apply(x.F, km, self.AVERAGE, x.F.shape[0])
        return SparseTensor(F, coordinate_map_key=out_key, coordinate_manager=mgr)


class MinkowskiAvgPooling(_PoolBase):
    AVERAGE = True


class MinkowskiSumPooling(_PoolBase):
    AVERAGE = False


class _NotOnHotPath(nn.Module):
    """Constructed only by helpers / backbones the reference never instantiates."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is outside the hot path (SURVEY.md section 8)")


class MinkowskiAvgUnpooling(_NotOnHotPath):
    pass


class MinkowskiPoolingTranspose(_NotOnHotPath):
    pass


class MinkowskiGlobalPooling(_NotOnHotPath):
    pass


class MinkowskiBroadcastMultiplication(_NotOnHotPath):
    pass


class MinkowskiInstanceNorm(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.num_features, self.eps = num_features, 1e-8
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, x: SparseTensor) -> SparseTensor:
        mgr = x.coordinate_manager
        off = mgr.scene_offsets(x.coordinate_map_key)
        F = _InstanceNormFn.apply(x.F, self.weight, self.bias, off, mgr.number_of_unique_batch_indices(), self.eps)
        return x._like(F)

    def extra_repr(self):
        return f"nchannels={self.num_features}"


class MinkowskiBatchNorm(nn.Module):
    """MinkowskiBatchNorm (mask3d/models/modules/common.py:22): batch norm over all rows of `.F`, parameters and running
    statistics under the reference's state-dict keys (`bn.weight`, `bn.bias`, `bn.running_mean`, `bn.running_var`,
    `bn.num_batches_tracked`).  The `nn.BatchNorm1d` child only HOLDS that state; the arithmetic is the instance-norm kernels
    with the whole batch as one segment (mk_instnorm_stats / mk_norm_apply / mk_instnorm_relu_bwd): statistics in one read,
    apply in one read + one write.  Not instantiated by the reference's default model (res16unet.py:16 selects instance norm)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor) -> SparseTensor:
        lib.require_cuda()
        bn, f = self.bn, x.F
        n, C = f.shape
        dev = f.device
        off = torch.tensor([0, n], dtype=torch.int32, device=dev)
        w = bn.weight.view(1, C) if bn.affine else torch.ones(1, C, device=dev)
        b = bn.bias.view(1, C) if bn.affine else torch.zeros(1, C, device=dev)
        if bn.training or not bn.track_running_stats or bn.running_mean is None:
            if bn.training and bn.track_running_stats and n > 0:
                with torch.no_grad():
                    mean, rstd = ops.instance_norm_stats(f.detach().contiguous(), off, 1, bn.eps)
                    var = (1.0 / (rstd[0] * rstd[0]) - bn.eps).clamp_min(0.0) * (n / max(n - 1, 1))  # unbiased, as BatchNorm1d stores it
                    bn.num_batches_tracked += 1
                    mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                    bn.running_mean.mul_(1 - mom).add_(mean[0], alpha=mom)
                    bn.running_var.mul_(1 - mom).add_(var, alpha=mom)
            return x._like(_InstanceNormFn.apply(f, w, b, off, 1, bn.eps))
        rstd = torch.rsqrt(bn.running_var.float() + bn.eps)
        if torch.is_grad_enabled() and (f.requires_grad or (bn.affine and bn.weight.requires_grad)):
            # frozen statistics under autograd: an affine map per channel
            return x._like((f.float() - bn.running_mean.float()) * (rstd * w) + b)
        return x._like(ops.norm_apply(f.contiguous(), off, 1, bn.running_mean.float().view(1, C).contiguous(), rstd.view(1, C).contiguous(),
                                      w, b, out_dtype=torch.float32))


class _Elementwise(nn.Module):
    MODULE = None

    def __init__(self, *a, **k):
        super().__init__()
        self.module = self.MODULE(*a, **k)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(self.module(x.F))


class MinkowskiReLU(_Elementwise):
    MODULE = nn.ReLU


class MinkowskiSigmoid(_Elementwise):
    MODULE = nn.Sigmoid


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(self.linear(x.F))


# ----------------------------------------------------------------------------------------------------
# utils
# ----------------------------------------------------------------------------------------------------
def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """ME.utils.sparse_collate (mask3d/datasets/utils.py:126): prepend the batch index, concatenate in list order."""
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
    bc = []
    for b, c in enumerate(coords):
        c = as_t(c)
        bc.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device), c.to(dtype)], dim=1))
    bcoords = torch.cat(bc, 0)
    f = torch.cat([as_t(x) for x in feats], 0)
    if device is not None:
        bcoords, f = bcoords.to(device), f.to(device)
    if labels is None:
        return bcoords, f
    l = torch.cat([as_t(x) for x in labels], 0)
    return bcoords, f, (l.to(device) if device is not None else l)


def voxelize_collate(coords, feats=None, labels=None, device="cuda"):
    """GPU form of the reference's per-scene dedup + collate (mask3d/datasets/utils.py:96-126):

        for every scene:  coords, unique_indices = np.unique(coords.int(), axis=0, return_index=True);  feats = feats[unique_indices] ...
        ME.utils.sparse_collate(coords, feats, labels)

    `coords`: list of per-scene [N_b, 3] arrays / tensors (floats are truncated like `.int()` does).  Returns
    (bcoords int32 [sum N', 4] batch-first, rows of a scene lexicographically sorted exactly as np.unique orders them,
    feats, labels, unique_indices: list of int64 tensors = np.unique's first-occurrence indices, and the SparseTensor-ready
    (coordinate_manager, coordinate_map_key) whose hash map was built on the way, so SparseTensor(...) needs no second insert)."""
    lib.require_cuda()
    dev = torch.device(device)
    cs = [torch.as_tensor(c).to(dev).to(torch.int32) for c in coords]
    counts = [int(c.shape[0]) for c in cs]
    bc = torch.cat([torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32, device=dev), c], 1) for b, c in enumerate(cs)], 0).contiguous() \
        if cs else torch.zeros((0, 4), dtype=torch.int32, device=dev)
    n = bc.shape[0]
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.call("mk_pack_keys", lib.p(bc), n, lib.p(keys), lib.p(status), lib.stream())
    # unsigned key order: the packed key uses 64 bits only through bit 63 = 0 (10 + 3 * 18 = 64 bits, batch < 512 keeps the sign
    # bit clear), so the signed int64 sort below is the unsigned order for every batch index the collate can produce here
    assert len(cs) < 512, "voxelize_collate: at most 511 scenes per batch"
    order = torch.sort(keys, stable=True).indices
    sorted_rows = bc[order].contiguous()
    mgr = CoordinateManager(D=3, device=dev)
    key, (uniq, _) = mgr.insert_and_map(sorted_rows)  # raises on out-of-range coordinates
    if int(status.item()) != 0:
        raise RuntimeError("coordinate out of the packable range [-2^17, 2^17)")
    bcoords = mgr.get_coordinates(key)
    src = order[uniq]  # original (concatenated) row of every voxel = lowest index among its duplicates (stable sort)
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    off = mgr.coord_map(key).scene_offsets_host(len(cs)) if len(cs) else [0]
    unique_indices = [src[off[b]:off[b + 1]] - starts[b] for b in range(len(cs))]

    def gather(parts):
        if parts is None:
            return None
        cat = torch.cat([torch.as_tensor(p).to(dev) for p in parts], 0)
        return cat[src]

    return bcoords, gather(feats), gather(labels), unique_indices, (mgr, key)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False, return_inverse=False,
                    return_maps_only=False, quantization_size=None, device="cuda"):
    """ME.utils.sparse_quantize on the GPU hash: floor(coords / quantization_size), first-occurrence
    unique voxels in first-occurrence order; voxels whose points disagree on the label get `ignore_label`."""
    lib.require_cuda()
    is_np = isinstance(coordinates, np.ndarray)
    c = torch.as_tensor(coordinates).to("cuda")
    if quantization_size is not None:
        q = torch.as_tensor(quantization_size, dtype=torch.float64, device=c.device)
        dc = torch.floor(c.double() / q).to(torch.int32)
    else:
        dc = torch.floor(c.double()).to(torch.int32) if c.is_floating_point() else c.to(torch.int32)
    c4 = torch.cat([torch.zeros((dc.shape[0], 1), dtype=torch.int32, device=dc.device), dc], 1) if dc.shape[1] == 3 else dc
    mgr = CoordinateManager(D=3, device=c4.device)
    _, (mapping, inverse) = mgr.insert_and_map(c4.contiguous())
    conv = (lambda t: t.cpu().numpy()) if is_np else (lambda t: t)
    if return_maps_only:
        return (conv(mapping), conv(inverse)) if return_inverse else conv(mapping)
    out = [conv(dc[mapping])]
    if features is not None:
        out.append(features[conv(mapping) if is_np else mapping.to(torch.as_tensor(features).device)])
    if labels is not None:
        lab = torch.as_tensor(labels).to(c4.device)
        first = lab[mapping].clone()
        conflict = torch.zeros(mapping.shape[0], dtype=torch.int32, device=c4.device)
        conflict.index_add_(0, inverse, (lab != first[inverse]).to(torch.int32))
        first[conflict > 0] = ignore_label
        out.append(conv(first))
    if return_index:
        out.append(conv(mapping))
    if return_inverse:
        out.append(conv(inverse))
    return out[0] if len(out) == 1 else tuple(out)
