"""Drop-in for the numeric callables of the reference's `phyclone/tree/utils.py`.

Same names, argument meaning and return conventions; the arithmetic runs in the sm_100a
kernels through the C ABI.  The reference wraps `compute_log_S` and `_convolve_two_children`
in content-hash LRU caches whose keys ignore child order (utils/utils.py:70-149) although the
floored fold is order-sensitive; this implementation always evaluates in
the order given.  `cache_info` / `cache_clear` / `__wrapped__` are kept so that
`phyclone/utils/dev.py` keeps working.
"""

from collections import namedtuple

import numpy as np

from .. import ops

_CacheInfo = namedtuple("CacheInfo", ["hits", "misses", "maxsize", "currsize"])


def _np_conv_dims(child_1, child_2):
    """reference tree/utils.py:60-85"""
    return ops.np_conv_dims(child_1, child_2)


def _convolve_two_children(child_1, child_2):
    """reference tree/utils.py:50-57.  The reference switches to an FFT at G >= 1000 (utils/math.py:241-263,
    same max-shift / floor contract, plus FFT round-off of ~1e-16 of the row maximum); the device kernel evaluates
    the direct sum for every supported grid (G <= 8192; several strip pairs per lane beyond 1088), so large grids get the exact value of that contract."""
    return ops.np_conv_dims(child_1, child_2)


_convolve_two_children.cache_info = lambda: _CacheInfo(0, 0, 0, 0)
_convolve_two_children.cache_clear = lambda: None


def _sub_compute_S(log_D):
    """reference tree/utils.py:25-30"""
    return ops.sub_compute_S(log_D)


def compute_log_D(child_log_R_values):
    """reference tree/utils.py:33-47"""
    n = len(child_log_R_values)
    if n == 0:
        return 0
    if n == 1:
        return child_log_R_values[0]
    acc = ops.np_conv_dims(child_log_R_values[0], child_log_R_values[1])
    for j in range(2, n):
        acc = ops.np_conv_dims(child_log_R_values[j], acc)
    return acc


def _compute_log_S_body(child_log_R_values):
    if len(child_log_R_values) == 0:
        return 0.0
    return np.ascontiguousarray(ops.compute_log_S_batch([np.asarray(child_log_R_values)])[0])


def compute_log_S(child_log_R_values):
    """reference tree/utils.py:7-22: prefix-LSE of the left fold of the children, fused on device."""
    return _compute_log_S_body(child_log_R_values)


compute_log_S.__wrapped__ = _compute_log_S_body
compute_log_S.cache_info = lambda: _CacheInfo(0, 0, 0, 0)
compute_log_S.cache_clear = lambda: None
