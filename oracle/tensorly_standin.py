"""Stand-in for the two third-party modules the reference imports but this image
does not have: ``tensorly`` (version unpinned by the reference: no requirements /
setup / lock file) and ``matplotlib``.

Only the three tensorly entry points the hot path calls are provided, restated from
tensorly's published semantics:

  tl.set_backend('pytorch')                         std:364,451  mn:12,353,445
  tl.cp_tensor.cp_to_tensor((weights, factors))     std:124      mn:182
  tl.tenalg.inner(a, b, n_modes)                    std:123      mn:181

Test infrastructure only (see oracle/__init__.py).
"""
import sys
import types

import numpy as np
import torch


def khatri_rao(mats):
    """Column-wise Kronecker product, first matrix's row index slowest."""
    res = mats[0]
    rank = res.shape[1]
    for e in mats[1:]:
        res = (res[:, None, :] * e[None, :, :]).reshape(-1, rank)
    return res


def cp_to_tensor(cp_tensor):
    weights, factors = cp_tensor
    factors = list(factors)
    shape = tuple(f.shape[0] for f in factors)
    if isinstance(weights, np.ndarray):
        weights = torch.as_tensor(weights)
    if weights is None:
        weights = torch.ones(factors[0].shape[1], dtype=factors[0].dtype, device=factors[0].device)
    if len(factors) == 1:
        return torch.sum(factors[0] * weights, dim=1)
    dt = torch.promote_types(weights.dtype, factors[0].dtype)
    f0 = factors[0].to(dt) * weights.to(dt)
    kr = khatri_rao([f.to(dt) for f in factors[1:]])
    return (f0 @ kr.T).reshape(shape)


def inner(a, b, n_modes=None):
    if n_modes is None:
        if a.shape != b.shape:
            raise ValueError('shapes must match when n_modes is None')
        return torch.sum(a * b)
    if tuple(a.shape[a.ndim - n_modes:]) != tuple(b.shape[:n_modes]):
        raise ValueError(f'inner: trailing {n_modes} dims of {tuple(a.shape)} != leading dims of {tuple(b.shape)}')
    s = 1
    for d in b.shape[:n_modes]:
        s *= int(d)
    out_shape = tuple(a.shape[:a.ndim - n_modes]) + tuple(b.shape[n_modes:])
    return (a.reshape(-1, s) @ b.reshape(s, -1)).reshape(out_shape)


def install():
    """Inject stand-in ``tensorly`` and empty ``matplotlib`` modules (idempotent)."""
    if 'tensorly' not in sys.modules:
        tl = types.ModuleType('tensorly')
        tl.__standin__ = True
        tl.set_backend = lambda name: None
        tl.cp_tensor = types.ModuleType('tensorly.cp_tensor')
        tl.cp_tensor.cp_to_tensor = cp_to_tensor
        tl.tenalg = types.ModuleType('tensorly.tenalg')
        tl.tenalg.inner = inner
        tl.tenalg.khatri_rao = khatri_rao
        sys.modules['tensorly'] = tl
        sys.modules['tensorly.cp_tensor'] = tl.cp_tensor
        sys.modules['tensorly.tenalg'] = tl.tenalg
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType('matplotlib')
        plt = types.ModuleType('matplotlib.pyplot')
        mpl.pyplot = plt
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = plt
