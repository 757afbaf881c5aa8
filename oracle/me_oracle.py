"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (numpy + plain PyTorch, fp32/fp64) of the MinkowskiEngine 0.5.4 semantics that the
reference's hot path consumes.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
leg may import this file.

PARITY UNPINNED at the MinkowskiEngine boundary: the reference pins `MinkowskiEngine==0.5.4`
(reference: mask3d/requirements_pip_freeze.txt:62) but does not vendor it (mask3d/.gitignore:5),
it is not installable in this container, and the reference ships no tests or golden vectors
for this path (SURVEY.md sections 4 and 8c).  The semantics below restate ME 0.5.4's published algorithm
 and are anchored instead on
  * the reference's own call sites: mask3d/models/modules/common.py:128-195 (conv / conv_tr),
    mask3d/models/res16unet.py:241-314, mask3d/models/mask3d.py:286-299,520-561,
    mask3d/datasets/utils.py:96-126,249-255;
  * self-consistency against stock PyTorch dense ops, which ARE the reference's own arithmetic
    for the same maths (tests/test_oracle.py): sparse conv == F.conv3d on the densified grid,
    conv-transpose == F.conv_transpose3d, avg-pool == dense sum / dense count,
    instance norm == F.instance_norm per scene, kernel map == brute-force dictionary lookup;
  * golden vectors produced by running the reference's UNMODIFIED model code
    (mask3d/models/*.py) on top of this module (tests/golden/make_golden.py).

Every function cites the reference file:line (or the ME behaviour recalled in SURVEY.md) it follows.
The module exposes the ME Python surface listed in SURVEY.md section 8b so that `import MinkowskiEngine`
can be aliased to it (oracle/shims.py) when generating fixtures.
"""
from __future__ import annotations

from enum import Enum
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

# --------------------------------------------------------------------------------------
# coordinate keys / packing
# --------------------------------------------------------------------------------------

_BIAS = 1 << 17  # coordinates must lie in [-2^17, 2^17)
_BITS = 18


def pack_keys(coords: np.ndarray) -> np.ndarray:
    """(batch, x, y, z) int32 rows -> one int64 key per row (order-preserving on each field)."""
    c = np.asarray(coords, dtype=np.int64)
    assert c.ndim == 2 and c.shape[1] == 4
    if c.size:
        assert c[:, 0].min() >= 0 and c[:, 0].max() < 1023, "batch index out of range"
        assert c[:, 1:].min() >= -_BIAS and c[:, 1:].max() < _BIAS, "coordinate out of range"
    return (c[:, 0] << (3 * _BITS)) | ((c[:, 1] + _BIAS) << (2 * _BITS)) | ((c[:, 2] + _BIAS) << _BITS) | (c[:, 3] + _BIAS)


def _to_tuple(v, D=3) -> Tuple[int, ...]:
    if isinstance(v, (int, np.integer)):
        return (int(v),) * D
    if isinstance(v, torch.Tensor):
        v = v.tolist()
    v = tuple(int(x) for x in v)
    assert len(v) == D
    return v


class CoordinateMapKey:
    """(tensor_stride, string_id) — ME's key of a coordinate map."""

    def __init__(self, tensor_stride, string_id: str = ""):
        self.tensor_stride = (tensor_stride,) * 3 if isinstance(tensor_stride, int) else tuple(int(v) for v in tensor_stride)
        self.string_id = string_id

    def get_key(self):
        return (list(self.tensor_stride), self.string_id)

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def get_coordinate_size(self):
        return len(self.tensor_stride) + 1

    def __hash__(self):
        return hash((self.tensor_stride, self.string_id))

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and (self.tensor_stride, self.string_id) == (other.tensor_stride, other.string_id)

    def __repr__(self):
        return f"CoordinateMapKey(stride={list(self.tensor_stride)}, id={self.string_id!r})"


class RegionType(Enum):
    """Values 0,1,2 — constructed from ints at import time by the reference
    (mask3d/models/modules/common.py:73)."""

    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class SparseTensorQuantizationMode(Enum):
    RANDOM_SUBSAMPLE = 0
    UNWEIGHTED_AVERAGE = 1
    UNWEIGHTED_SUM = 2
    NO_QUANTIZATION = 3
    MAX_POOL = 4


# --------------------------------------------------------------------------------------
# pure functions on coordinate arrays (these are what the CUDA path is compared with)
# --------------------------------------------------------------------------------------


def insert_and_map(coords: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """ME CoordinateMap insert (SURVEY.md appendix A `insert_and_map`): first occurrence wins, map rows
    in ascending source-row order.  Returns (unique_index, inverse_mapping), both int64."""
    keys = pack_keys(coords)
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    # np.unique orders by key; ME orders by first occurrence => rank the first-occurrence rows
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return first[order].astype(np.int64), rank[inv.reshape(-1)].astype(np.int64)


def stride_coords(coords: np.ndarray, out_stride: Sequence[int]) -> np.ndarray:
    """ME stride map (SURVEY.md appendix A `stride`): floor(c / ts_out) * ts_out per axis, batch kept,
    rows of the new map in first-occurrence order over ascending input rows."""
    c = np.asarray(coords, dtype=np.int64).copy()
    ts = np.asarray(out_stride, dtype=np.int64)
    c[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
    uniq, _ = insert_and_map(c)
    return c[uniq].astype(np.int32)


def kernel_offsets(kernel_size: Sequence[int], dilation: Sequence[int], tensor_stride: Sequence[int]) -> np.ndarray:
    """HYPER_CUBE region offsets, x fastest: k = ix + Kx*(iy + Ky*iz).  Odd K is centred
    (i - K//2); even K is anchored at the output coordinate (i)."""
    K = list(kernel_size)
    offs = []
    for iz in range(K[2]):
        for iy in range(K[1]):
            for ix in range(K[0]):
                o = []
                for a, i in enumerate((ix, iy, iz)):
                    off = (i - K[a] // 2) if K[a] % 2 == 1 else i
                    o.append(off * dilation[a] * tensor_stride[a])
                offs.append(o)
    return np.asarray(offs, dtype=np.int64).reshape(-1, 3)


def neighbor_table(in_coords: np.ndarray, out_coords: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """nbr[k, o] = row of (out_coords[o] + offsets[k]) in the input map, or -1.
    This is the output-stationary form of ME's kernel map (SURVEY.md appendix A `kernel_map`)."""
    in_keys = pack_keys(in_coords)
    order = np.argsort(in_keys, kind="stable")
    sorted_keys = in_keys[order]
    K = len(offsets)
    No = len(out_coords)
    nbr = np.full((K, No), -1, dtype=np.int32)
    oc = np.asarray(out_coords, dtype=np.int64)
    lim = _BIAS
    for k in range(K):
        q = oc.copy()
        q[:, 1:] += offsets[k]
        ok = np.all((q[:, 1:] >= -lim) & (q[:, 1:] < lim), axis=1)
        qk = pack_keys(np.where(ok[:, None], q, oc))
        pos = np.searchsorted(sorted_keys, qk)
        pos_c = np.minimum(pos, len(sorted_keys) - 1) if len(sorted_keys) else pos
        hit = ok & (pos < len(sorted_keys))
        if len(sorted_keys):
            hit &= sorted_keys[pos_c] == qk
            nbr[k, hit] = order[pos_c[hit]].astype(np.int32)
    return nbr


def kernel_map_from_table(nbr: np.ndarray, transpose: bool = False) -> Dict[int, np.ndarray]:
    """Canonical ME-visible kernel map: {k: int32[2, n_k]} rows (in, out), pairs of each k sorted by
    out row (by in row after a transpose swap keeps the same list).  Offsets with no pair are omitted,
    as ME's `kernel_map` does."""
    out = {}
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0].astype(np.int32)
        if len(o) == 0:
            continue
        i = nbr[k, o].astype(np.int32)
        out[k] = np.stack([o, i] if transpose else [i, o]).astype(np.int32)
    return out


# --------------------------------------------------------------------------------------
# CoordinateManager
# --------------------------------------------------------------------------------------


class CoordinateManager:
    """Owns every coordinate map and kernel map derived from one input SparseTensor
    (ME ownership rule)."""

    def __init__(self, D: int = 3, **_):
        self.D = D
        self._maps: Dict[CoordinateMapKey, np.ndarray] = {}
        self._nbr: Dict[tuple, np.ndarray] = {}
        self._batch_rows: Dict[CoordinateMapKey, List[np.ndarray]] = {}

    # -- maps ---------------------------------------------------------------------------
    def insert_and_map(self, coords, tensor_stride=1, string_id: str = ""):
        c = np.ascontiguousarray(coords.detach().cpu().numpy() if isinstance(coords, torch.Tensor) else coords).astype(np.int32)
        uniq, inv = insert_and_map(c)
        key = CoordinateMapKey(_to_tuple(tensor_stride, self.D), string_id)
        assert key not in self._maps, "coordinate map key already exists"
        self._maps[key] = c[uniq]
        return key, (torch.from_numpy(uniq), torch.from_numpy(inv))

    def stride(self, in_key: CoordinateMapKey, stride) -> CoordinateMapKey:
        s = _to_tuple(stride, self.D)
        out_ts = tuple(a * b for a, b in zip(in_key.tensor_stride, s))
        key = CoordinateMapKey(out_ts, "")
        if all(x == 1 for x in s):
            return in_key
        if key not in self._maps:
            self._maps[key] = stride_coords(self._maps[in_key], out_ts)
        return key

    def exists(self, key) -> bool:
        return key in self._maps

    def get_coordinates(self, key: CoordinateMapKey) -> torch.Tensor:
        return torch.from_numpy(self._maps[key])

    def size(self, key) -> int:
        return len(self._maps[key])

    def number_of_unique_batch_indices(self) -> int:
        k0 = next(iter(self._maps))
        return int(self._maps[k0][:, 0].max()) + 1 if len(self._maps[k0]) else 0

    def batch_rows(self, key: CoordinateMapKey) -> List[np.ndarray]:
        """origin_map: ascending rows of every batch item."""
        if key not in self._batch_rows:
            b = self._maps[key][:, 0]
            nb = self.number_of_unique_batch_indices()
            self._batch_rows[key] = [np.nonzero(b == i)[0] for i in range(nb)]
        return self._batch_rows[key]

    origin_map = batch_rows

    # -- kernel maps -------------------------------------------------------------------
    def neighbor_table(self, in_key, out_key, kernel_size, dilation=1) -> np.ndarray:
        ks = _to_tuple(kernel_size, self.D)
        dl = _to_tuple(dilation, self.D)
        ck = (in_key, out_key, ks, dl)
        if ck not in self._nbr:
            offs = kernel_offsets(ks, dl, in_key.tensor_stride)
            self._nbr[ck] = neighbor_table(self._maps[in_key], self._maps[out_key], offs)
        return self._nbr[ck]

    def kernel_map(self, in_key, out_key, stride=1, kernel_size=3, dilation=1, region_type=RegionType.HYPER_CUBE,
                   region_offset=None, is_transpose: bool = False, is_pool: bool = False) -> Dict[int, torch.Tensor]:
        """{k: int32[2, n_k]}; transposed = the forward map between the same two coordinate maps with the
        columns swapped."""
        assert region_type in (RegionType.HYPER_CUBE, 0)
        if is_transpose:
            nbr = self.neighbor_table(out_key, in_key, kernel_size, dilation)
        else:
            nbr = self.neighbor_table(in_key, out_key, kernel_size, dilation)
        return {k: torch.from_numpy(v) for k, v in kernel_map_from_table(nbr, transpose=is_transpose).items()}


# --------------------------------------------------------------------------------------
# SparseTensor
# --------------------------------------------------------------------------------------


class SparseTensor:
    """ME.SparseTensor as consumed at mask3d/datasets/utils.py:250, mask3d/models/mask3d.py:287,538,549."""

    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None, coordinate_manager=None,
                 quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE, allocator_type=None,
                 minkowski_algorithm=None, requires_grad=None, device=None):
        assert isinstance(features, torch.Tensor)
        if coordinate_map_key is None:
            assert coordinates is not None
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager(D=coordinates.shape[1] - 1)
            coordinate_map_key, (uniq, inv) = coordinate_manager.insert_and_map(coordinates, tensor_stride)
            self.unique_index, self.inverse_mapping = uniq, inv
            if len(uniq) != len(features):
                features = features[uniq]  # RANDOM_SUBSAMPLE: first occurrence
        else:
            assert coordinate_manager is not None
            assert coordinate_manager.size(coordinate_map_key) == features.shape[0], "row count must match the map"
        if device is not None:
            features = features.to(device)
        if requires_grad is not None:
            features.requires_grad_(requires_grad)
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self._manager = coordinate_manager

    # --- properties the reference touches
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

    @property
    def decomposed_features(self):
        return [self._F[torch.from_numpy(r)] for r in self._manager.batch_rows(self.coordinate_map_key)]

    @property
    def decomposed_coordinates(self):
        C = self.C
        return [C[torch.from_numpy(r), 1:] for r in self._manager.batch_rows(self.coordinate_map_key)]

    def float(self):
        return SparseTensor(self._F.float(), coordinate_map_key=self.coordinate_map_key, coordinate_manager=self._manager)

    def double(self):
        return SparseTensor(self._F.double(), coordinate_map_key=self.coordinate_map_key, coordinate_manager=self._manager)

    def detach(self):
        return SparseTensor(self._F.detach(), coordinate_map_key=self.coordinate_map_key, coordinate_manager=self._manager)

    def _same_map(self, other):
        assert self._manager is other._manager and self.coordinate_map_key == other.coordinate_map_key, "different coordinate maps"

    def __iadd__(self, other):
        self._same_map(other)
        self._F = self._F + other._F  # out-of-place: keeps autograd simple, same values
        return self

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            self._same_map(other)
            other = other._F
        return SparseTensor(self._F + other, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self._manager)

    def __repr__(self):
        return f"SparseTensor(oracle, F={tuple(self._F.shape)}, key={self.coordinate_map_key})"


def cat(*tensors):
    """MinkowskiOps.cat (reference call sites: mask3d/models/res16unet.py:276-306)."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    t0 = tensors[0]
    for t in tensors[1:]:
        t0._same_map(t)
    return SparseTensor(torch.cat([t.F for t in tensors], dim=1), coordinate_map_key=t0.coordinate_map_key,
                        coordinate_manager=t0.coordinate_manager)


# --------------------------------------------------------------------------------------
# functional ops (plain differentiable torch)
# --------------------------------------------------------------------------------------


def conv_from_table(x: torch.Tensor, W: torch.Tensor, nbr: np.ndarray, n_out: int) -> torch.Tensor:
    """out[o] = sum_k in[nbr[k,o]] @ W[k]  — ME's per-offset gather / GEMM / scatter-add
    (SURVEY.md appendix A `MinkowskiConvolution`), written with index_select + mm + index_add."""
    out = torch.zeros((n_out, W.shape[-1]), dtype=x.dtype, device=x.device)
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o) == 0:
            continue
        i = torch.from_numpy(nbr[k, o].astype(np.int64))
        out = out.index_add(0, torch.from_numpy(o), x.index_select(0, i) @ W[k])
    return out


def conv_transpose_from_table(x: torch.Tensor, W: torch.Tensor, nbr_down: np.ndarray, n_out: int) -> torch.Tensor:
    """Transposed conv between coarse (in) and fine (out): uses the stride-s conv's table
    nbr_down[k, coarse] = fine row, with in/out swapped and the same k."""
    out = torch.zeros((n_out, W.shape[-1]), dtype=x.dtype, device=x.device)
    for k in range(nbr_down.shape[0]):
        c = np.nonzero(nbr_down[k] >= 0)[0]
        if len(c) == 0:
            continue
        f = torch.from_numpy(nbr_down[k, c].astype(np.int64))
        out = out.index_add(0, f, x.index_select(0, torch.from_numpy(c)) @ W[k])
    return out


def avg_pool_from_table(x: torch.Tensor, nbr: np.ndarray, n_out: int, average: bool = True) -> torch.Tensor:
    """MinkowskiAvgPooling: mean over EXISTING inputs (sum / nnz); MinkowskiSumPooling: no divide
    (SURVEY.md section 8a-a13)."""
    out = torch.zeros((n_out, x.shape[1]), dtype=x.dtype, device=x.device)
    cnt = np.zeros(n_out, dtype=np.int64)
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o) == 0:
            continue
        out = out.index_add(0, torch.from_numpy(o), x.index_select(0, torch.from_numpy(nbr[k, o].astype(np.int64))))
        cnt[o] += 1
    if average:
        out = out / torch.from_numpy(np.maximum(cnt, 1)).to(x.dtype)[:, None]
    return out


def instance_norm(x: torch.Tensor, batch_rows: List[np.ndarray], weight=None, bias=None, eps: float = 1e-8) -> torch.Tensor:
    """MinkowskiInstanceNorm: per scene, per channel, biased variance, eps 1e-8, then affine with
    weight/bias [1, C]."""
    pieces = []
    for rows in batch_rows:
        r = torch.from_numpy(rows)
        xb = x.index_select(0, r)
        mu = xb.mean(0, keepdim=True)
        xc = xb - mu
        var = (xc * xc).mean(0, keepdim=True)
        pieces.append((r, xc / torch.sqrt(var + eps)))
    out = torch.zeros_like(x)
    for r, y in pieces:
        out = out.index_copy(0, r, y)
    if weight is not None:
        out = out * weight + bias
    return out


# --------------------------------------------------------------------------------------
# nn.Modules of the ME surface
# --------------------------------------------------------------------------------------


class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose: bool = False, region_type=RegionType.HYPER_CUBE,
                 region_offsets=None, expand_coordinates: bool = False, axis_types=None, dimension=-1):
        assert dimension > 0
        self.dimension = dimension
        self.kernel_size = list(_to_tuple(kernel_size, dimension))
        self.kernel_stride = list(_to_tuple(stride, dimension))
        self.kernel_dilation = list(_to_tuple(dilation, dimension))
        self.region_type = region_type
        self.region_offsets = region_offsets
        self.axis_types = axis_types
        self.is_transpose = is_transpose
        self.expand_coordinates = expand_coordinates
        self.kernel_volume = int(np.prod(self.kernel_size))
        self.requires_strided_coordinates = any(s > 1 for s in self.kernel_stride)


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class _ConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 is_transpose=False, expand_coordinates=False, convolution_mode=None, dimension=-1):
        super().__init__()
        assert dimension > 0
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, is_transpose=is_transpose, dimension=dimension)
        self.kernel_generator = kernel_generator
        self.in_channels, self.out_channels, self.dimension = in_channels, out_channels, dimension
        self.is_transpose = is_transpose
        kv = kernel_generator.kernel_volume
        self.use_mm = kv == 1 and not kernel_generator.requires_strided_coordinates
        shape = (in_channels, out_channels) if self.use_mm else (kv, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters(is_transpose)

    def reset_parameters(self, is_transpose=False):
        # ME: n = (out if transpose else in) * kernel_volume ; U(-1/sqrt(n), 1/sqrt(n))  
        with torch.no_grad():
            n = (self.out_channels if is_transpose else self.in_channels) * self.kernel_generator.kernel_volume
            stdv = 1.0 / float(np.sqrt(n))
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)


class MinkowskiConvolution(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator,
                         is_transpose=False, dimension=dimension)

    def forward(self, x: SparseTensor, coordinates=None) -> SparseTensor:
        mgr, kg = x.coordinate_manager, self.kernel_generator
        if self.use_mm:
            out_key = x.coordinate_map_key
            F = x.F @ self.kernel
        else:
            out_key = mgr.stride(x.coordinate_map_key, kg.kernel_stride)
            nbr = mgr.neighbor_table(x.coordinate_map_key, out_key, kg.kernel_size, kg.kernel_dilation)
            F = conv_from_table(x.F, self.kernel, nbr, mgr.size(out_key))
        if self.bias is not None:
            F = F + self.bias
        return SparseTensor(F, coordinate_map_key=out_key, coordinate_manager=mgr)


class MinkowskiConvolutionTranspose(_ConvBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator,
                         is_transpose=True, dimension=dimension)

    def forward(self, x: SparseTensor, coordinates=None) -> SparseTensor:
        mgr, kg = x.coordinate_manager, self.kernel_generator
        in_key = x.coordinate_map_key
        if self.use_mm:
            out_key, F = in_key, x.F @ self.kernel
        else:
            ts = tuple(a // b for a, b in zip(in_key.tensor_stride, kg.kernel_stride))
            out_key = CoordinateMapKey(ts, "")
            assert mgr.exists(out_key), "transposed conv needs the encoder's coordinate map (generate_new_coords=False)"
            nbr_down = mgr.neighbor_table(out_key, in_key, kg.kernel_size, kg.kernel_dilation)
            F = conv_transpose_from_table(x.F, self.kernel, nbr_down, mgr.size(out_key))
        if self.bias is not None:
            F = F + self.bias
        return SparseTensor(F, coordinate_map_key=out_key, coordinate_manager=mgr)


class _PoolBase(nn.Module):
    AVERAGE = True

    def __init__(self, kernel_size=-1, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, dimension=dimension)
        self.kernel_generator = kernel_generator
        self.dimension = dimension

    def forward(self, x: SparseTensor, coordinates=None) -> SparseTensor:
        mgr, kg = x.coordinate_manager, self.kernel_generator
        out_key = mgr.stride(x.coordinate_map_key, kg.kernel_stride)
        nbr = mgr.neighbor_table(x.coordinate_map_key, out_key, kg.kernel_size, kg.kernel_dilation)
        F = avg_pool_from_table(x.F, nbr, mgr.size(out_key), average=self.AVERAGE)
        return SparseTensor(F, coordinate_map_key=out_key, coordinate_manager=mgr)


class MinkowskiAvgPooling(_PoolBase):
    AVERAGE = True


class MinkowskiSumPooling(_PoolBase):
    AVERAGE = False


class _NotOnHotPath(nn.Module):
    """Constructors the reference only reaches from unused helpers / backbones."""

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
        self.num_features = num_features
        self.eps = 1e-8
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, x: SparseTensor) -> SparseTensor:
        rows = x.coordinate_manager.batch_rows(x.coordinate_map_key)
        F = instance_norm(x.F, rows, self.weight, self.bias, self.eps)
        return SparseTensor(F, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return SparseTensor(self.bn(x.F), coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


class _Elementwise(nn.Module):
    MODULE = None

    def __init__(self, *a, **k):
        super().__init__()
        self.module = self.MODULE(*a, **k)

    def forward(self, x: SparseTensor) -> SparseTensor:
        F = self.module(x.F.clone() if getattr(self.module, "inplace", False) else x.F)
        return SparseTensor(F, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


class MinkowskiReLU(_Elementwise):
    MODULE = nn.ReLU


class MinkowskiSigmoid(_Elementwise):
    MODULE = nn.Sigmoid


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return SparseTensor(self.linear(x.F), coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


# --------------------------------------------------------------------------------------
# utils
# --------------------------------------------------------------------------------------


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """ME.utils.sparse_collate (reference call: mask3d/datasets/utils.py:126): prepend the batch index,
    concatenate in list order."""
    bc = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(np.asarray(c)) if not isinstance(c, torch.Tensor) else c
        bc.append(torch.cat([torch.full((len(c), 1), b, dtype=dtype), c.to(dtype)], dim=1))
    bcoords = torch.cat(bc, 0)
    f = torch.cat([torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x for x in feats], 0)
    if labels is None:
        return bcoords, f
    l = torch.cat([torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x for x in labels], 0)
    return bcoords, f, l


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False, return_inverse=False,
                    return_maps_only=False, quantization_size=None, device="cpu"):
    """ME.utils.sparse_quantize: floor(coords / qsize), first-occurrence unique rows
    in first-occurrence order; voxels whose points disagree on the label get `ignore_label`."""
    is_t = isinstance(coordinates, torch.Tensor)
    c = coordinates.detach().cpu().numpy() if is_t else np.asarray(coordinates)
    if quantization_size is not None:
        dc = np.floor(c / np.asarray(quantization_size)).astype(np.int32)
    else:
        dc = np.floor(c).astype(np.int32)
    c4 = np.concatenate([np.zeros((len(dc), 1), np.int32), dc], 1) if dc.shape[1] == 3 else dc
    mapping, inverse = insert_and_map(c4)
    wrap = (lambda a: torch.from_numpy(np.ascontiguousarray(a))) if is_t else (lambda a: a)
    if return_maps_only:
        return (wrap(mapping), wrap(inverse)) if return_inverse else wrap(mapping)
    out = [wrap(dc[mapping])]
    if features is not None:
        out.append(features[mapping])
    if labels is not None:
        lab = labels.detach().cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
        first = lab[mapping].copy()
        conflict = np.zeros(len(mapping), dtype=bool)
        np.logical_or.at(conflict, inverse, lab != first[inverse])
        first[conflict] = ignore_label
        out.append(wrap(first) if isinstance(labels, torch.Tensor) else first)
    if return_index:
        out.append(wrap(mapping))
    if return_inverse:
        out.append(wrap(inverse))
    return out[0] if len(out) == 1 else tuple(out)
