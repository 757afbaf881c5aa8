"""ctypes binding of libmk3d.so (include/mk3d.h).  Fails loudly: there is no CPU / eager fallback.

Every wrapper takes torch CUDA tensors, checks dtype / contiguity, and passes raw device pointers plus the
current torch CUDA stream to the C ABI.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from typing import Optional

import torch

from . import build as _build

F32, BF16 = 0, 1
TILE = 128

_lib = None

_SIGS = {
    "mk_last_error": (c_char_p, []),
    "mk_version": (c_int, []),
    "mk_device_supported": (c_int, []),
    "mk_hash_capacity": (c_int64, [c_int64]),
    "mk_coordmap_workspace_bytes": (c_int64, [c_int64]),
    "mk_coordmap_build": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_pack_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "mk_hash_query": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "mk_kernel_map": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "mk_kernel_map_same": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "mk_kernel_map_compose5": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mk_tile_kmask": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "mk_transpose_map": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "mk_conv_fwd_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "mk_conv_weight_image_elems": (c_int64, [c_int, c_int, c_int, c_int]),
    "mk_conv_weight_image_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mk_conv_fwd_bf16": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                 c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mk_conv_weight_image_tf32_elems": (c_int64, [c_int, c_int, c_int, c_int]),
    "mk_conv_weight_image_tf32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mk_conv_fwd_tf32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                 c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mk_set_gather_mode": (c_int, [c_int]),
    "mk_set_deterministic": (c_int, [c_int]),
    "mk_row_sort_keys": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mk_row_mask_keys": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "mk_group_rows_workspace_bytes": (c_int64, [c_int64]),
    "mk_group_rows": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_permute_table": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "mk_conv_stem_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p]),
    "mk_conv_stem_wgrad_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "mk_conv_stem_wgrad_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_conv_dgrad_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "mk_conv_wgrad_f32": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "mk_conv_wgrad_plan_elems": (c_int64, [c_int64, c_int]),
    "mk_conv_wgrad_plan": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mk_conv_wgrad_bf16": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                   c_void_p]),
    "mk_conv_wgrad_det_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int]),
    "mk_conv_wgrad_bf16_ws": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_int64, c_void_p]),
    "mk_instnorm_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "mk_instnorm_stats": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_norm_apply": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p, c_int, c_void_p]),
    "mk_instnorm_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_instnorm_relu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_conv_instnorm_bf16": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "mk_pool_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "mk_pool_fwd_dt": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "mk_pool_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int, c_int64, c_void_p, c_void_p]),
    "mk_mask_logits": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "mk_mask_compact_workspace_bytes": (c_int64, [c_int64, c_int]),
    "mk_mask_compact": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_pool_threshold": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mk_scene_minmax": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mk_fourier_pe": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                              c_void_p]),
    "mk_set_xattn_simt": (c_int, [c_int]),
    "mk_xattn_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int64]),
    "mk_masked_xattn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                    c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_masked_xattn_fwd_lse": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                        c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_xattn_bwd_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int64]),
    "mk_masked_xattn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                    c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_fps_workspace_bytes": (c_int64, [c_int, c_int]),
    "mk_fps": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_target_bits": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p]),
    "mk_mask_costs_workspace_bytes": (c_int64, [c_int, c_int64, c_int, c_int]),
    "mk_mask_costs": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mk_mask_loss_bwd": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mk_cast": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p]),
    "mk_copy_channels": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building with nvcc first if the .so is absent or stale).  Raises if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(_build.LIB):
        raise RuntimeError(f"libmk3d.so not found at {_build.LIB}; run `python -m room_instance_segmentation_mask3d_b200.build`")
    lib = ctypes.CDLL(_build.LIB)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {load().mk_last_error().decode()}")


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("room_instance_segmentation_mask3d_b200 needs a CUDA device: the hot path has no CPU fallback")


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def tdtype(code: int) -> torch.dtype:
    return torch.float32 if code == F32 else torch.bfloat16


def p(t: Optional[torch.Tensor]):
    """Device pointer of a contiguous CUDA tensor as a plain int (ctypes converts it for c_void_p parameters)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return t.data_ptr()


_stream_cache = None  # (cuda_stream handle) pinned for the duration of a `stream_scope`


def stream():
    if _stream_cache is not None:
        return _stream_cache
    return torch.cuda.current_stream().cuda_stream


class stream_scope:
    """Resolve the current torch CUDA stream once for a whole forward pass (torch.cuda.current_stream() costs several
    microseconds per call, i.e. milliseconds per step at ~2000 launches).  Callers must not switch streams inside."""

    def __enter__(self):
        global _stream_cache
        self.prev = _stream_cache
        _stream_cache = torch.cuda.current_stream().cuda_stream
        return self

    def __exit__(self, *exc):
        global _stream_cache
        _stream_cache = self.prev
        return False


# kernels launched per C-ABI call (memsets not counted) — bench.py reports the total as `gpu_launches`
def cache_built():
    """Guard of a cached device tensor built on the current stream: [event behind its producer, that stream, streams that have waited].
    Cached weight images / derived layouts are handed to whichever stream runs the next forward (bench.py alternates two): a
    consumer on another stream must not read an image the building stream is still writing."""
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(cur)
    return [ev, cur.cuda_stream, set()]


def cache_use(guard) -> None:
    """Make the current stream wait (once) for the producer of a cached tensor built on another stream."""
    cur = torch.cuda.current_stream()
    sid = cur.cuda_stream
    if sid != guard[1] and sid not in guard[2]:
        cur.wait_event(guard[0])
        guard[2].add(sid)


_KERNELS_PER_CALL = {"mk_coordmap_build": 7, "mk_instnorm_stats": 2, "mk_masked_xattn_fwd": 3, "mk_masked_xattn_fwd_lse": 3, "mk_masked_xattn_bwd": 5, "mk_instnorm_bwd": 3, "mk_instnorm_relu_bwd": 3, "mk_scene_minmax": 3,
                     "mk_fps": 25, "mk_mask_compact": 2, "mk_group_rows": 6, "mk_mask_costs": 3, "mk_conv_instnorm_bf16": 4, "mk_conv_wgrad_plan": 3, "mk_conv_wgrad_bf16_ws": 2, "mk_conv_stem_wgrad_f32": 2}
kernel_launches = 0
_timed = None  # optional {entry-point name: [(start_event, end_event, tag), ...]} filled when bench.py instruments a step
_tag = None


def instrument(names=None):
    """Start collecting CUDA-event pairs around the named entry points (None = stop)."""
    global _timed
    _timed = {n: [] for n in names} if names else None
    return _timed


def set_tag(tag) -> None:
    global _tag
    _tag = tag


_fns = {}
_scratch = {}


def scratch(nbytes: int, device) -> torch.Tensor:
    """Reusable per-device scratch buffer for kernels that need a workspace only for the duration of one C-ABI call
    (stream-ordered: the next user is enqueued behind the previous one on the same stream)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream())
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _scratch[key] = torch.empty(max(int(nbytes * 1.5), 1 << 20), dtype=torch.uint8, device=device)
    return buf


def reserve_device_pool(gigabytes: float = 32.0, device=None, stream=None) -> None:
    """Make torch's caching allocator hold ONE large splittable segment up front.

    Scenes differ in size, so a step's allocation pattern never repeats exactly; with the default growth policy the allocator
    keeps finding no fit every few steps and calls cudaMalloc for another segment (measured: 91 -> 145 segments over 30
    inference steps), each a 10-100 ms stall with an implicit device synchronisation — the source of occasional 2x slower
    benchmark phases.  A B200 has 180 GB: reserving a few tens of GB once and releasing them to the cache turns every later
    request into a split of that segment.  The allocator caches blocks PER STREAM: pass `stream` to reserve for a side stream
    (every stream a forward pass runs on needs its own reservation)."""
    require_cuda()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    free, _ = torch.cuda.mem_get_info(dev)
    nbytes = int(min(gigabytes * (1 << 30), 0.25 * free))
    if nbytes > 0:
        if stream is None:
            t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        else:
            with torch.cuda.stream(stream):
                t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        del t


def call(name: str, *args) -> None:
    global kernel_launches
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = (getattr(load(), name), _KERNELS_PER_CALL.get(name, 1))
    kernel_launches += fn[1]
    if _timed is not None and name in _timed:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _check(fn[0](*args), name)
        e.record()
        _timed[name].append((s, e, _tag))
        return
    if fn[0](*args) != 0:
        _check(1, name)


def query(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))
