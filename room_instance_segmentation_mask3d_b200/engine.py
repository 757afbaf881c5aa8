"""Coordinate manager: coordinate maps, kernel maps and tensor-core tile plans on the GPU.

Mirrors MinkowskiEngine 0.5.4's CoordinateManager as the reference uses it: one manager per
input SparseTensor, shared by every tensor derived from it; coordinate maps cached by (tensor_stride, string_id),
kernel maps cached by (in_key, out_key, kernel_size, dilation).  All arithmetic runs in libmk3d.so
(csrc/coords.cu); torch supplies device memory and the stream.
"""
from __future__ import annotations

import os

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib

I32 = torch.int32


def _tuple3(v, D=3) -> Tuple[int, ...]:
    if isinstance(v, int):
        return (v,) * D
    if isinstance(v, torch.Tensor):
        v = v.tolist()
    v = tuple(int(x) for x in v)
    assert len(v) == D, f"expected {D} values, got {v}"
    return v


class CoordinateMapKey:
    """(tensor_stride, string_id)."""

    __slots__ = ("tensor_stride", "string_id")

    def __init__(self, tensor_stride, string_id: str = ""):
        self.tensor_stride = _tuple3(tensor_stride)
        self.string_id = string_id

    def get_key(self):
        return (list(self.tensor_stride), self.string_id)

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def get_coordinate_size(self):
        return len(self.tensor_stride) + 1

    def __hash__(self):
        return hash((self.tensor_stride, self.string_id))

    def __eq__(self, o):
        return isinstance(o, CoordinateMapKey) and self.tensor_stride == o.tensor_stride and self.string_id == o.string_id

    def __repr__(self):
        return f"CoordinateMapKey(stride={list(self.tensor_stride)}, id={self.string_id!r})"


class CoordMap:
    """One coordinate map: rows + its hash table + the per-scene row ranges."""

    def __init__(self, coords: torch.Tensor, keys: torch.Tensor, vals: torch.Tensor, cap: int):
        self.coords, self.keys, self.vals, self.cap = coords, keys, vals, cap
        self.n = coords.shape[0]
        self._offsets: Optional[torch.Tensor] = None
        self._offsets_host: Optional[List[int]] = None

    def scene_offsets(self, B: int) -> torch.Tensor:
        """int32 [B+1]: rows of scene b are [off[b], off[b+1]) — batch items are contiguous and ascending."""
        if self._offsets is None or self._offsets.numel() != B + 1:
            b = self.coords[:, 0].contiguous()
            bounds = torch.arange(B + 1, device=b.device, dtype=I32)
            self._offsets = torch.searchsorted(b, bounds, right=False).to(I32)
            self._offsets_host = None
        return self._offsets

    def scene_offsets_host(self, B: int) -> List[int]:
        if self._offsets_host is None:
            self._offsets_host = self.scene_offsets(B).tolist()
        return self._offsets_host


_SORT_MODE = os.environ.get("MK3D_SORT", "mask")


def group_rows(keys: torch.Tensor) -> torch.Tensor:
    """int32 [n] = the permutation a stable sort by `keys` (int32, few distinct values, all >= 0) gives, through mk_group_rows
    (hash set of the distinct keys + counting sort: csrc/group_sort.cu).  No host sync; more than 1024 distinct keys (never seen:
    a scene has a few hundred distinct offset masks) would leave the rows ungrouped, which only costs tile efficiency."""
    n = keys.shape[0]
    order = torch.empty(n, dtype=I32, device=keys.device)
    if n == 0:
        return order
    wsb = lib.query("mk_group_rows_workspace_bytes", n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=keys.device)
    lib.call("mk_group_rows", lib.p(keys), n, lib.p(order), None, lib.p(ws), wsb, lib.stream())
    return order


class KernelMap:
    """Output-stationary kernel map nbr[K, n_slots] (+ what the tensor-core kernel needs)."""

    def __init__(self, nbr: torch.Tensor, K: int, n_out: int, out_rows: Optional[torch.Tensor] = None,
                 out_coords: Optional[torch.Tensor] = None, out_stride: Tuple[int, int, int] = (1, 1, 1)):
        self.out_coords, self.out_stride = out_coords, out_stride  # coordinates of the output rows (for spatial tile sorting)
        self.nbr = nbr                    # int32 [K, n_slots]
        self.K = K
        self.n_slots = nbr.shape[1]
        self.n_out = n_out
        self.out_rows = out_rows          # int32 [n_slots] or None (slot == row)
        self._kmask: Optional[torch.Tensor] = None
        self._pairs_dev: Optional[torch.Tensor] = None
        self._pairs: Optional[int] = None
        self._sorted: Optional["KernelMap"] = None
        self._sorted_is_self = False
        self._wgrad_plan: Optional[torch.Tensor] = None

    def _build_kmask(self):
        ntiles = (self.n_slots + lib.TILE - 1) // lib.TILE
        self._kmask = torch.empty(max(ntiles, 1), dtype=I32, device=self.nbr.device)
        self._pairs_dev = torch.zeros(1, dtype=torch.int64, device=self.nbr.device)
        lib.call("mk_tile_kmask", lib.p(self.nbr), self.K, self.n_slots, lib.p(self._kmask) if self.K <= 32 else None,
                 lib.p(self._pairs_dev), lib.stream())

    def sorted_plan(self) -> "KernelMap":
        """Mask-sorted tile plan: output slots are reordered so that rows with the same set of present offsets share
        tensor-core tiles (a lexicographic tile of a surface scene touches ~23 of 27 offsets, a mask-sorted one ~10, i.e.
        2.3x less gather + MMA work); inside a mask group the rows keep their lexicographic voxel order (a stable 32-bit sort by
        the mask alone; ordering each group along a Z-curve with 64-bit keys — MK3D_SORT=morton — costs twice the sort and
        measured no faster gathers: 96->96 at 889 k voxels 0.313 vs 0.321 ms).  Slots map back to rows through `out_rows`."""
        # NB: no `self._sorted = self`: a self-reference is a cycle, and cyclic garbage is only freed by Python's generational
        # collector — every kernel map (up to 444 MB for the 5^3 table of a 0.9 M-voxel scene) then outlives its step, the
        # caching allocator keeps growing by cudaMalloc, and the eventual collection is a 100-300 ms pause.
        if self._sorted_is_self:
            return self
        if self._sorted is None:
            if self.K == 1 or self.K > 32 or self.out_rows is not None or self.n_out == 0 or self.out_coords is None:
                self._sorted_is_self = True
                return self
            else:
                n, dev = self.n_out, self.nbr.device
                if _SORT_MODE == "morton":  # A/B: (mask, batch, Z-curve) 64-bit keys
                    keys = torch.empty(n, dtype=torch.int64, device=dev)
                    st = (ctypes.c_int32 * 3)(*self.out_stride)
                    lib.call("mk_row_sort_keys", lib.p(self.nbr), self.K, n, lib.p(self.out_coords), ctypes.cast(st, ctypes.c_void_p),
                             lib.p(keys), lib.stream())
                else:  # default: group by offset mask, rows of a group stay in lexicographic order (32-bit stable sort)
                    keys = torch.empty(n, dtype=I32, device=dev)
                    lib.call("mk_row_mask_keys", lib.p(self.nbr), self.K, n, lib.p(keys), lib.stream())
                if _SORT_MODE == "morton":
                    order = torch.sort(keys, stable=True).indices.to(I32)
                else:
                    order = group_rows(keys)
                nbr_s = torch.empty_like(self.nbr)
                lib.call("mk_permute_table", lib.p(self.nbr), self.K, n, lib.p(order), n, lib.p(nbr_s), lib.stream())
                self._sorted = KernelMap(nbr_s, self.K, n, order)
                self._sorted._sorted_is_self = True
        return self._sorted

    @property
    def tile_kmask(self) -> torch.Tensor:
        if self._kmask is None:
            self._build_kmask()
        return self._kmask

    @property
    def wgrad_plan(self) -> torch.Tensor:
        """(offset, tile) unit list of this plan for the tensor-core weight gradient (mk_conv_wgrad_plan); K <= 32."""
        if self._wgrad_plan is None:
            ntiles = (self.n_slots + lib.TILE - 1) // lib.TILE
            self._wgrad_plan = torch.empty(lib.query("mk_conv_wgrad_plan_elems", ntiles, self.K), dtype=I32, device=self.nbr.device)
            lib.call("mk_conv_wgrad_plan", lib.p(self.tile_kmask), ntiles, self.K, lib.p(self._wgrad_plan), lib.stream())
        return self._wgrad_plan

    @property
    def num_pairs(self) -> int:
        """Actual pair count P = sum_k n_k (the algorithmic-FLOP figure).  Host sync."""
        if self._pairs is None:
            if self._pairs_dev is None:
                self._build_kmask()
            self._pairs = int(self._pairs_dev.item())
        return self._pairs


def kernel_offsets(kernel_size: Sequence[int], dilation: Sequence[int], tensor_stride: Sequence[int]) -> List[List[int]]:
    """HYPER_CUBE offsets, x fastest; odd K centred, even K anchored at the output coordinate."""
    out = []
    for iz in range(kernel_size[2]):
        for iy in range(kernel_size[1]):
            for ix in range(kernel_size[0]):
                o = []
                for a, i in enumerate((ix, iy, iz)):
                    off = (i - kernel_size[a] // 2) if kernel_size[a] % 2 == 1 else i
                    o.append(off * dilation[a] * tensor_stride[a])
                out.append(o)
    return out


_OFFSETS_CACHE: Dict[tuple, torch.Tensor] = {}  # (kernel_size, dilation, tensor_stride, device) -> int32 [K, 3]; scene independent


class CoordinateManager:
    def __init__(self, D: int = 3, device=None, **_):
        lib.require_cuda()
        assert D == 3, "the hot path is 3-D (SURVEY.md section 8)"
        self.D = D
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._maps: Dict[CoordinateMapKey, CoordMap] = {}
        self._kmaps: Dict[tuple, KernelMap] = {}
        self._tplans: Dict[tuple, KernelMap] = {}
        self._offs_cache: Dict[tuple, torch.Tensor] = {}
        self._B: Optional[int] = None

    # ------------------------------------------------------------------ coordinate maps
    def _build_map(self, in_coords: torch.Tensor, stride: Tuple[int, int, int]):
        n = in_coords.shape[0]
        dev = in_coords.device
        cap = lib.query("mk_hash_capacity", n)
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        vals = torch.empty(cap, dtype=I32, device=dev)
        out_coords = torch.empty((n, 4), dtype=I32, device=dev)
        uniq = torch.empty(n, dtype=I32, device=dev)
        inv = torch.empty(n, dtype=I32, device=dev)
        meta = torch.zeros(2, dtype=I32, device=dev)  # [n_out, status]
        wsb = lib.query("mk_coordmap_workspace_bytes", n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        stride_host = (ctypes.c_int32 * 3)(*stride)
        lib.call("mk_coordmap_build", lib.p(in_coords), n, ctypes.cast(stride_host, ctypes.c_void_p), lib.p(keys), lib.p(vals), cap,
                 lib.p(out_coords), lib.p(uniq), lib.p(inv), ctypes.c_void_p(meta.data_ptr()), ctypes.c_void_p(meta.data_ptr() + 4),
                 lib.p(ws), wsb, lib.stream())
        n_out, status = meta.tolist()  # one host sync per map: sizes are data-dependent
        if status & 1:
            raise RuntimeError("coordinate out of the packable range [-2^17, 2^17) or batch index >= 1023")
        if status & 2:
            raise RuntimeError("rows are not sorted by batch index: this backend keeps every scene's rows contiguous (what "
                               "ME.utils.sparse_collate produces, mask3d/datasets/utils.py:126); sort the rows by coordinates[:, 0] first")
        return CoordMap(out_coords[:n_out], keys, vals, cap), uniq[:n_out], inv

    def insert_and_map(self, coords: torch.Tensor, tensor_stride=1, string_id: str = ""):
        key = CoordinateMapKey(_tuple3(tensor_stride), string_id)
        if key in self._maps:
            raise RuntimeError(f"coordinate map {key} already exists")
        c = coords.to(device=self.device, dtype=I32).contiguous()
        assert c.dim() == 2 and c.shape[1] == 4, "coordinates must be [N, 4] = (batch, x, y, z)"
        cmap, uniq, inv = self._build_map(c, (1, 1, 1))
        self._maps[key] = cmap
        return key, (uniq.long(), inv.long())

    def stride(self, in_key: CoordinateMapKey, stride) -> CoordinateMapKey:
        s = _tuple3(stride)
        if all(x == 1 for x in s):
            return in_key
        out_ts = tuple(a * b for a, b in zip(in_key.tensor_stride, s))
        key = CoordinateMapKey(out_ts, "")
        if key not in self._maps:
            cmap, _, _ = self._build_map(self._maps[in_key].coords, out_ts)
            self._maps[key] = cmap
        return key

    def exists(self, key) -> bool:
        return key in self._maps

    def get_coordinates(self, key) -> torch.Tensor:
        return self._maps[key].coords

    def size(self, key) -> int:
        return self._maps[key].n

    def coord_map(self, key) -> CoordMap:
        return self._maps[key]

    def number_of_unique_batch_indices(self) -> int:
        if self._B is None:
            k0 = next(iter(self._maps))
            c = self._maps[k0].coords
            self._B = int(c[-1, 0].item()) + 1 if c.shape[0] else 0  # batch indices ascend (sparse_collate)
        return self._B

    def scene_offsets(self, key) -> torch.Tensor:
        return self._maps[key].scene_offsets(self.number_of_unique_batch_indices())

    def scene_offsets_host(self, key) -> List[int]:
        return self._maps[key].scene_offsets_host(self.number_of_unique_batch_indices())

    def origin_map(self, key) -> List[torch.Tensor]:
        off = self.scene_offsets_host(key)
        return [torch.arange(off[b], off[b + 1], device=self.device) for b in range(len(off) - 1)]

    # ------------------------------------------------------------------ kernel maps
    def _offsets_dev(self, ks, dil, ts) -> torch.Tensor:
        ck = (ks, dil, ts, self.device)
        t = _OFFSETS_CACHE.get(ck)
        if t is None:
            t = _OFFSETS_CACHE[ck] = torch.tensor(kernel_offsets(ks, dil, ts), dtype=I32, device=self.device).contiguous()
        return t

    def neighbor_table(self, in_key, out_key, kernel_size, dilation=1) -> KernelMap:
        ks, dl = _tuple3(kernel_size), _tuple3(dilation)
        ck = (in_key, out_key, ks, dl)
        km = self._kmaps.get(ck)
        if km is None:
            im, om = self._maps[in_key], self._maps[out_key]
            K = ks[0] * ks[1] * ks[2]
            nbr = torch.empty((K, om.n), dtype=I32, device=self.device)
            offs = self._offsets_dev(ks, dl, in_key.tensor_stride)
            same = in_key == out_key and all(k % 2 == 1 for k in ks) and os.environ.get("MK3D_KMAP_SYM", "1") != "0"
            if same and ks == (5, 5, 5) and len(set(dl)) == 1 and os.environ.get("MK3D_KMAP_COMPOSE", "1") != "0":
                # the stem's 5^3 table composed from the 3^3 table of the same map (which the last decoder stage needs anyway):
                # table reads instead of 62 hash probes per voxel, bit-identical (csrc/coords.cu k_kernel_map_compose5)
                km3 = self.neighbor_table(in_key, out_key, 3, dl)
                step = (ctypes.c_int32 * 3)(*[d * t for d, t in zip(dl, in_key.tensor_stride)])
                lib.call("mk_kernel_map_compose5", lib.p(im.keys), lib.p(im.vals), im.cap, lib.p(om.coords), om.n,
                         ctypes.cast(step, ctypes.c_void_p), lib.p(km3.nbr), lib.p(nbr), lib.stream())
            else:
                lib.call("mk_kernel_map_same" if same else "mk_kernel_map", lib.p(im.keys), lib.p(im.vals), im.cap, lib.p(om.coords), om.n,
                         lib.p(offs), K, lib.p(nbr), lib.stream())
            km = KernelMap(nbr, K, om.n, out_coords=om.coords, out_stride=out_key.tensor_stride)
            self._kmaps[ck] = km
        return km

    def scene_select_table(self, key) -> KernelMap:
        """Table with one "offset" per scene: nbr[b, r] = r if row r belongs to scene b else -1.  Lets the tensor-core conv
        kernel apply a different weight per scene (the mask-logit product, mask3d.py:534-535) with its per-tile offset mask."""
        ck = ("scene_select", key)
        km = self._kmaps.get(ck)
        if km is None:
            cm = self._maps[key]
            B = self.number_of_unique_batch_indices()
            rows = torch.arange(cm.n, device=self.device, dtype=I32)
            if B <= 1:
                nbr = rows[None, :].contiguous()
            else:
                b_of = cm.coords[:, 0]
                nbr = torch.where(b_of[None, :] == torch.arange(B, device=self.device, dtype=I32)[:, None], rows[None, :],
                                  torch.full((1, 1), -1, dtype=I32, device=self.device)).contiguous()
            km = KernelMap(nbr, max(B, 1), cm.n)
            self._kmaps[ck] = km
        return km

    def transpose_plan(self, coarse_key, fine_key, kernel_size, dilation=1) -> KernelMap:
        """Tile plan of the transposed conv coarse -> fine.  Every fine row has exactly one (offset k, parent) pair, so
        slots are the fine rows stably sorted by k and padded per k to a multiple of 128: every tile then has a single
        active offset and the tensor-core kernel does no wasted MMA work."""
        ks, dl = _tuple3(kernel_size), _tuple3(dilation)
        ratio = tuple(c // f for c, f in zip(coarse_key.tensor_stride, fine_key.tensor_stride))
        if ks != ratio or any(d != 1 for d in dl):
            # one (offset, parent) pair per fine row only holds for non-overlapping kernels (kernel_size == stride, dilation 1:
            # the 2^3 / stride-2 convs and pools of Res16UNet34C); overlapping kernels (3/2, 1/2) would race on parent[f] / koff[f]
            raise NotImplementedError(f"transposed / strided-backward maps need kernel_size == stride and dilation 1 (got kernel "
                                      f"{list(ks)}, stride {list(ratio)}, dilation {list(dl)}): outside the hot path (SURVEY.md section 8)")
        ck = (coarse_key, fine_key, ks, dl)
        plan = self._tplans.get(ck)
        if plan is None:
            down = self.neighbor_table(fine_key, coarse_key, ks, dl)  # nbr_down[k, coarse] = fine row
            n_fine, n_coarse, K = self._maps[fine_key].n, self._maps[coarse_key].n, down.K
            parent = torch.empty(n_fine, dtype=I32, device=self.device)
            koff = torch.empty(n_fine, dtype=I32, device=self.device)
            lib.call("mk_transpose_map", lib.p(down.nbr), K, n_coarse, n_fine, lib.p(parent), lib.p(koff), lib.stream())
            # slots = fine rows grouped by k (stable: ascending row inside a group), every group padded to a multiple of 128.
            # No host sync: the slot count is the upper bound ceil(n_fine / 128) + K tiles; surplus tiles have an empty offset
            # mask and are skipped by the kernel.
            order = group_rows((koff + 1).contiguous()).long()  # koff in [-1, K): keys >= 0
            k_sorted = koff[order].long()
            ks_ = torch.arange(K, device=self.device)
            start = torch.searchsorted(k_sorted, ks_)  # k_sorted ascends: group boundaries without a host sync
            counts = torch.searchsorted(k_sorted, ks_, right=True) - start
            padded = (counts + (lib.TILE - 1)) // lib.TILE * lib.TILE
            pstart = torch.cumsum(padded, 0) - padded
            slot = pstart[k_sorted] + (torch.arange(n_fine, device=self.device) - start[k_sorted])
            n_slots = ((n_fine + lib.TILE - 1) // lib.TILE + K) * lib.TILE
            out_rows = torch.full((n_slots,), -1, dtype=I32, device=self.device)
            out_rows[slot] = order.to(I32)
            nbr_t = torch.full((K, n_slots), -1, dtype=I32, device=self.device)
            nbr_t[k_sorted, slot] = parent[order]
            plan = KernelMap(nbr_t, K, n_fine, out_rows)
            plan.parent, plan.koff = parent, koff
            self._tplans[ck] = plan
        return plan

    def record_stream(self, stream) -> None:
        """Tell the caching allocator that every tensor of this manager is (also) used on `stream` — needed when the maps were
        built on one stream (Mask3D.prepare's side stream) and are consumed on another: their blocks must not be handed out again
        before the consumer's work that was enqueued up to the moment they are freed has finished."""
        def rec(t):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(stream)

        for m in self._maps.values():
            for t in (m.coords, m.keys, m.vals, m._offsets):
                rec(t)
        seen = set()

        def walk(km):
            if km is None or id(km) in seen:
                return
            seen.add(id(km))
            for t in (km.nbr, km.out_rows, km._kmask, km._pairs_dev, km._wgrad_plan, km.out_coords, getattr(km, "parent", None),
                      getattr(km, "koff", None)):
                rec(t)
            walk(km._sorted)

        for km in list(self._kmaps.values()) + list(self._tplans.values()):
            walk(km)

    def prebuild(self, in_key, levels: int = 5, first_kernel: int = 5) -> List[CoordinateMapKey]:
        """Build every coordinate map, kernel map and tile plan a Res16UNet-style network needs, up front: the host syncs
        of the data-dependent map sizes then happen while the GPU queue is still empty instead of draining it mid-network."""
        keys = [in_key]
        for _ in range(levels - 1):
            keys.append(self.stride(keys[-1], 2))
        self.neighbor_table(keys[0], keys[0], first_kernel, 1)
        for lvl, k in enumerate(keys):
            self.neighbor_table(k, k, 3, 1).sorted_plan().tile_kmask
            self.neighbor_table(k, k, 1, 1).tile_kmask
            self.scene_offsets(k)
            if lvl > 0:
                self.neighbor_table(keys[lvl - 1], k, 2, 1).sorted_plan().tile_kmask
                self.transpose_plan(k, keys[lvl - 1], 2, 1).tile_kmask
        return keys

    def kernel_map(self, in_key, out_key, stride=1, kernel_size=3, dilation=1, region_type=None, region_offset=None,
                   is_transpose: bool = False, is_pool: bool = False) -> Dict[int, torch.Tensor]:
        """ME-visible kernel map {k: int32 [2, n_k]} rows (in, out), pairs of each k sorted by out row (the canonical
        form of SURVEY.md section 8a-a6); transposed = the forward map of the same two coordinate maps, columns swapped."""
        km = self.neighbor_table(out_key, in_key, kernel_size, dilation) if is_transpose else \
            self.neighbor_table(in_key, out_key, kernel_size, dilation)
        res = {}
        present = km.nbr >= 0
        for k in range(km.K):
            o = torch.nonzero(present[k]).flatten().to(I32)
            if o.numel() == 0:
                continue
            i = km.nbr[k, o.long()]
            res[k] = torch.stack([o, i] if is_transpose else [i, o]).to(I32)
        return res
