"""Deterministic, name-keyed synthetic weights.

There is no network for checkpoints, and a 39.6 M-parameter state dict does not fit in a test
fixture.  Every tensor is therefore generated from a numpy Generator seeded by CRC32 of its
state-dict name (platform-stable), so the build container (which fills the reference's own modules
when writing tests/golden/*.npz) and the GPU box (which fills this package's model) obtain
bit-identical weights from the name + shape alone.

Layouts follow the reference checkpoint format: ME conv `kernel` [K, Cin, Cout]
(`[Cin, Cout]` for 1x1), `bias` [1, Cout]; MinkowskiInstanceNorm `weight` / `bias` [1, C];
`pos_enc.gauss_B` [3, 64].
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np


def _fan_in(name: str, shape: Tuple[int, ...]) -> int:
    if name.endswith(".kernel"):
        return int(np.prod(shape[:-1]))  # K * Cin  (or Cin)
    if len(shape) >= 2:
        return int(np.prod(shape[1:]))  # Linear [out, in], Conv1d [out, in, 1]
    return 1


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng([zlib.crc32(name.encode()), seed])
    shape = tuple(int(s) for s in shape)
    if name.endswith("gauss_B"):
        return rng.standard_normal(shape).astype(np.float32)
    if name.endswith("num_batches_tracked"):
        return np.zeros(shape, dtype=np.int64)
    if name.endswith("running_var"):
        return (1.0 + 0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)
    if name.endswith("running_mean"):
        return (0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)
    is_matrix = sum(1 for d in shape if d > 1) >= 2 or name.endswith(".kernel")
    if is_matrix:
        b = 1.0 / np.sqrt(max(_fan_in(name, shape), 1))
        return rng.uniform(-b, b, shape).astype(np.float32)
    if name.endswith("weight"):  # norm scale ([1, C] or [C])
        return (1.0 + 0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)
    return (0.1 * rng.uniform(-1, 1, shape)).astype(np.float32)  # biases


def synth_state_dict(spec: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0) -> Dict[str, np.ndarray]:
    return {n: synth_tensor(n, s, seed) for n, s in spec}


def load_synth(module, seed: int = 0) -> None:
    """Overwrite every parameter/buffer of a torch module (in place) with its synthetic value."""
    import torch

    with torch.no_grad():
        for name, t in module.state_dict().items():
            v = torch.from_numpy(synth_tensor(name, tuple(t.shape), seed))
            t.copy_(v.to(t.dtype))
