"""Version counters for parameters that kernels update through raw pointers (torch's own `_version` only sees torch ops).
Kept OUTSIDE the tensors — attributes set on a Parameter leak into `state_dict()` pickles.  Keyed by object identity
(tensors compare element-wise, so they cannot be weak-dictionary keys); the entry dies with the tensor."""
from __future__ import annotations

import weakref
from typing import Dict

_versions: Dict[int, int] = {}


def bump(p) -> None:
    k = id(p)
    if k not in _versions:
        _versions[k] = 0
        weakref.finalize(p, _versions.pop, k, None)
    _versions[k] += 1


def get(p) -> int:
    return _versions.get(id(p), 0)
