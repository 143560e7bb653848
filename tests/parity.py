"""Shared comparison helpers for parity tests (SURVEY.md Appendix A.6 buckets)."""

import numpy as np

RTOL = 1e-9  # BASELINE.json north_star: node log-likelihoods within 1e-9 relative in FP64


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    both_inf = np.isinf(got) & np.isinf(ref) & (np.sign(got) == np.sign(ref))
    with np.errstate(invalid="ignore"):
        err = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    return np.where(both_inf, 0.0, err)


def assert_close(got, ref, rtol=RTOL, what=""):
    err = rel_err(got, ref)
    bad = ~(err <= rtol)
    assert not bad.any(), "%s: %d/%d entries beyond %.1e (max %.3e)" % (what, bad.sum(), bad.size, rtol, np.nanmax(err))


def assert_close_bucketed(got, ref, rtol=RTOL, what="", max_boundary=0, boundary_nats_below_max=690.0):
    """Elementwise rtol; entries far enough below their row max to sit in the denormal /
    floor band of the max-shifted convolution (Appendix A.6) may differ and are only counted."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = rel_err(got, ref)
    bad = ~(err <= rtol)
    if not bad.any():
        return 0
    row_max = ref.max(axis=-1, keepdims=True)
    in_band = (row_max - ref) > boundary_nats_below_max
    hard = bad & ~in_band
    assert not hard.any(), "%s: %d normal-range entries beyond %.1e (max %.3e)" % (what, hard.sum(), rtol, np.nanmax(err[hard]))
    n = int((bad & in_band).sum())
    assert n <= max_boundary, "%s: %d boundary-band entries differ (allowed %d)" % (what, n, max_boundary)
    return n
