"""Import shim: the reference binds scatter_mean/max/min at import time (mask3d/models/mask3d.py:12,138-143) and never
calls them on the hot path."""


def _never(*a, **k):
    raise NotImplementedError("torch_scatter is bound but never called on the hot path (mask3d.py:138-143)")


scatter_mean = scatter_max = scatter_min = _never
