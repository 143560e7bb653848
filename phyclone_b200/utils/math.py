"""Scalar / tiny-vector host helpers with the names of the reference's `phyclone/utils/math.py`.

These are O(particles) host-side scalars between kernel launches (prior terms, proposal
probabilities, RNG draws); the tile arithmetic is on the device.  log-gamma goes through the
C library exactly like the reference's Numba `math.lgamma` (utils/math.py:126-151).
"""

import ctypes
import ctypes.util
import math
from functools import lru_cache

import numpy as np

from .._hostmod import load as _load_hostmod

_hm = _load_hostmod()
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.lgamma.restype = ctypes.c_double
_libm.lgamma.argtypes = [ctypes.c_double]


def log_gamma(x):
    return _libm.lgamma(float(x))


def log_factorial(x):
    return _libm.lgamma(float(x) + 1.0)


@lru_cache(maxsize=None)
def cached_log_factorial(x):
    return log_factorial(x)


def log_binomial_coefficient(n, x):
    return log_factorial(n) - log_factorial(x) - log_factorial(n - x)


@lru_cache(maxsize=2048)
def cached_log_binomial_coefficient(n, x):
    return log_binomial_coefficient(n, x)


def log_sum_exp(log_X):
    """utils/math.py:102-118 (max shift, terms summed left to right with libm exp / log): host extension."""
    return _hm.log_sum_exp(np.ascontiguousarray(log_X, dtype=np.float64))


def log_normalize(log_p):
    """utils/math.py:121-123"""
    log_p = np.asarray(log_p, dtype=np.float64)
    return log_p - log_sum_exp(log_p)


def exp_normalize(log_p):
    """utils/math.py:37-59"""
    log_p = np.asarray(log_p, dtype=np.float64)
    log_norm = log_sum_exp(log_p)
    p = np.array([math.exp(v) for v in (log_p - log_norm).tolist()])
    total = 0.0
    for v in p.tolist():
        total += v
    return p / total, log_norm


def discrete_rvs(p, rng):
    """utils/math.py:19-21"""
    p = p / np.sum(p)
    return rng.multinomial(1, p).argmax()


def conv_log(log_x, log_y, ans=None):
    """utils/math.py:212-238: direct log-space convolution with a per-output max shift (no floor); device kernel.
    `ans`, if given, is filled in place like the reference's output argument."""
    from .. import ops

    out = ops.conv_log(log_x, log_y)
    if ans is None:
        return out
    ans[...] = out
    return ans


def fft_convolve_two_children(child_1, child_2):
    """utils/math.py:241-263: the reference's FFT evaluation of the `_np_conv_dims` contract (per-row max shift,
    truncation to G terms, `<= 0 -> 1e-100` floor) for G >= 1000.  Here the same contract is evaluated by the
    direct-sum kernel, i.e. without the FFT's round-off noise; G <= 8192."""
    from .. import ops

    return ops.np_conv_dims(child_1, child_2)


def non_log_conv(child_log_R, prev_log_D_n):
    """utils/math.py:266-282: the 1-D form of `_np_conv_dims`."""
    from .. import ops

    a = np.asarray(child_log_R, dtype=np.float64)
    b = np.asarray(prev_log_D_n, dtype=np.float64)
    return ops.np_conv_dims(b[None, None, :], a[None, None, :])[0, 0]  # np.convolve(R, D): R is the second operand
