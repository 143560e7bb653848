"""NumPy-facing batched wrappers over the HOST-buffer C entry points.

Each function copies its inputs to the GPU, runs the sm_100a kernels and copies the result
back (that is the `e2e` path of bench.py).  Shapes: tile = [S, G] float64.
"""

import ctypes as C

import numpy as np

from . import _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def np_conv_dims(c1, c2):
    """Batched `_np_conv_dims` (reference tree/utils.py:60-85).  c1, c2: [B,S,G] or [S,G]."""
    lib = _lib.require_device()
    c1, c2 = _f64(c1), _f64(c2)
    if c1.shape != c2.shape or c1.ndim not in (2, 3):
        raise ValueError("np_conv_dims: operands must both be [S,G] or [B,S,G]")
    single = c1.ndim == 2
    a, b = (c1[None], c2[None]) if single else (c1, c2)
    B, S, G = a.shape
    out = np.empty_like(a)
    _lib.check(lib.pcb_np_conv_dims_host(_ptr(a), _ptr(b), _ptr(out), B, S, G))
    return out[0] if single else out


def sub_compute_S(log_D):
    """Batched `_sub_compute_S` (reference tree/utils.py:25-30)."""
    lib = _lib.require_device()
    x = _f64(log_D)
    single = x.ndim == 2
    a = x[None] if single else x
    B, S, G = a.shape
    out = np.empty_like(a)
    _lib.check(lib.pcb_sub_compute_S_host(_ptr(a), _ptr(out), B, S, G))
    return out[0] if single else out


def _ragged(child_lists, S=None, G=None):
    off = np.zeros(len(child_lists) + 1, dtype=np.int64)
    flat = []
    for b, ch in enumerate(child_lists):
        ch = _f64(ch)
        if ch.size:
            if ch.ndim != 3:
                raise ValueError("each child list must be [C,S,G]")
            flat.append(ch)
        off[b + 1] = off[b] + (ch.shape[0] if ch.size else 0)
    if flat:
        kids = np.concatenate(flat, axis=0)
    else:
        kids = np.zeros((0, S or 1, G or 1))
    return np.ascontiguousarray(kids), off


def compute_log_S_batch(child_lists):
    """`compute_log_S.__wrapped__` for a batch of child lists (reference tree/utils.py:7-22)."""
    lib = _lib.require_device()
    kids, off = _ragged(child_lists)
    B = len(child_lists)
    _, S, G = kids.shape
    out = np.empty((B, S, G))
    _lib.check(lib.pcb_compute_log_S_host(_ptr(kids), _ptr(off), _ptr(out), B, S, G))
    return out


def node_update_batch(log_p, child_lists):
    """`TreeNode.update_node_from_child_r_vals` for a batch (reference tree/tree_node.py:61-71)."""
    lib = _lib.require_device()
    log_p = _f64(log_p)
    B, S, G = log_p.shape
    kids, off = _ragged(child_lists, S, G)
    out = np.empty_like(log_p)
    _lib.check(lib.pcb_node_update_host(_ptr(log_p), _ptr(kids), _ptr(off), _ptr(out), B, S, G))
    return out


def root_terms(root_tiles):
    """(sum_s LSE_g tile[s,:], sum_s tile[s,-1]) per tile (reference tree/distributions.py:197-200)."""
    lib = _lib.require_device()
    x = _f64(root_tiles)
    single = x.ndim == 2
    a = x[None] if single else x
    B, S, G = a.shape
    lse, last = np.empty(B), np.empty(B)
    _lib.check(lib.pcb_root_terms_host(_ptr(a), _ptr(lse), _ptr(last), B, S, G))
    return (lse[0], last[0]) if single else (lse, last)


DENSITY = {"binomial": 0, "beta-binomial": 1}


def likelihood_grid(ref_counts, alt_counts, tumour_content, cn, mu, log_pi, n_geno, grid_size, density, precision):
    """`_compute_liklihood_grid` over [M,S] (mutation, sample) pairs (reference data/pyclone.py:323-436)."""
    lib = _lib.require_device()
    if density not in DENSITY:
        raise ValueError("unknown density %r" % (density,))
    if density == "beta-binomial" and precision is None:
        raise Exception("Precision must be set when using Beta-Binomial.")
    a = np.ascontiguousarray(ref_counts, dtype=np.int64)
    b = np.ascontiguousarray(alt_counts, dtype=np.int64)
    t = _f64(tumour_content)
    cn = np.ascontiguousarray(cn, dtype=np.int64)
    mu = _f64(mu)
    log_pi = _f64(log_pi)
    ng = np.ascontiguousarray(n_geno, dtype=np.int32)
    M, S = a.shape
    Cmax = cn.shape[2]
    out = np.empty((M, S, grid_size))
    _lib.check(
        lib.pcb_likelihood_grid_host(
            _ptr(a), _ptr(b), _ptr(t), _ptr(cn), _ptr(mu), _ptr(log_pi), _ptr(ng),
            M, S, Cmax, grid_size, DENSITY[density], float(precision if precision is not None else 0.0), _ptr(out),
        )
    )
    return out


def cluster_sum(grids, cluster_off):
    """value[k] = sum of the member grids of cluster k (reference data/pyclone.py:89-94)."""
    lib = _lib.require_device()
    g = _f64(grids)
    off = np.ascontiguousarray(cluster_off, dtype=np.int64)
    K = len(off) - 1
    _, S, G = g.shape
    out = np.empty((K, S, G))
    _lib.check(lib.pcb_cluster_sum_host(_ptr(g), _ptr(off), K, S, G, _ptr(out)))
    return out


def outlier_marginal(values):
    """`DataPoint.outlier_marginal_prob` for [K,S,G] values (reference data/base.py:31-37)."""
    lib = _lib.require_device()
    v = _f64(values)
    single = v.ndim == 2
    a = v[None] if single else v
    K, S, G = a.shape
    out = np.empty(K)
    _lib.check(lib.pcb_outlier_marginal_host(_ptr(a), K, S, G, _ptr(out)))
    return out[0] if single else out


def map_trees(node_log_p, tree_off, post_order, child_off, child_idx, want_log_r_max=False):
    """MAP prevalence indices of a batch of trees (reference process_trace/map.py:48-166).
    node_log_p [N,S,G]; tree_off [B+1]; post_order [N] (children first, each tree's root last); child_off [N+1];
    child_idx [E] in fold order.  Returns max_idx [N,S] int32 (and log_R_max [N,S,G] if asked)."""
    lib = _lib.require_device()
    lp = _f64(node_log_p)
    N, S, G = lp.shape
    to, po, co, ci = (np.ascontiguousarray(x, dtype=np.int32) for x in (tree_off, post_order, child_off, child_idx))
    if len(po) != N or len(co) != N + 1:
        raise ValueError("map_trees: post_order / child_off do not match the number of nodes")
    max_idx = np.zeros((N, S), dtype=np.int32)
    lrm = np.empty((N, S, G)) if want_log_r_max else None
    _lib.check(lib.pcb_map_trees_host(_ptr(lp), _ptr(to), _ptr(po), _ptr(co), _ptr(ci) if len(ci) else None,
                                      len(to) - 1, N, len(ci), S, G, _ptr(max_idx), None if lrm is None else _ptr(lrm)))
    return (max_idx, lrm) if want_log_r_max else max_idx


def conv_log(log_x, log_y):
    """`conv_log` (reference utils/math.py:212-238) for [n] or [B,n] vectors."""
    lib = _lib.require_device()
    x, y = _f64(log_x), _f64(log_y)
    if x.shape != y.shape:
        raise ValueError("conv_log: shape mismatch")
    single = x.ndim == 1
    a, b = (x[None], y[None]) if single else (x, y)
    out = np.empty_like(a)
    _lib.check(lib.pcb_conv_log_host(_ptr(a), _ptr(b), _ptr(out), a.shape[0], a.shape[1]))
    return out[0] if single else out


def swarm_weights(log_w):
    """(log_Z, log_W, W, ess) of `ParticleSwarm` (reference smc/swarm/swarm.py:21-61)."""
    lib = _lib.require_device()
    x = _f64(log_w)
    N = x.shape[0]
    log_Z, ess = np.empty(1), np.empty(1)
    log_W, W = np.empty(N), np.empty(N)
    _lib.check(lib.pcb_swarm_weights_host(_ptr(x), N, _ptr(log_Z), _ptr(log_W), _ptr(W), _ptr(ess)))
    return float(log_Z[0]), log_W, W, float(ess[0])


def normalize_segments(log_p, seg_off):
    """Per-segment (log_q, q, log_norm) (reference smc/kernels/semi_adapted.py:151-163, utils/math.py:37-59)."""
    lib = _lib.require_device()
    x = _f64(log_p)
    off = np.ascontiguousarray(seg_off, dtype=np.int64)
    U = len(off) - 1
    log_q, q, ln = np.empty_like(x), np.empty_like(x), np.empty(U)
    _lib.check(lib.pcb_normalize_segments_host(_ptr(x), _ptr(off), U, _ptr(log_q), _ptr(q), _ptr(ln)))
    return log_q, q, ln


def resample(W, uniforms):
    """Optional device resampling mode: inverse-CDF ancestors for the supplied uniforms."""
    lib = _lib.require_device()
    w, u = _f64(W), _f64(uniforms)
    anc = np.empty(u.shape[0], dtype=np.int64)
    _lib.check(lib.pcb_resample_host(_ptr(w), w.shape[0], _ptr(u), u.shape[0], _ptr(anc)))
    return anc


def fp64_peak(iters=4096):
    """Measured FP64 FMA throughput (FLOP/s) of the current device: the roofline denominator."""
    lib = _lib.require_device()
    flops, ms = C.c_double(), C.c_double()
    _lib.check(lib.pcb_fp64_peak(iters, C.byref(flops), C.byref(ms)))
    return flops.value, ms.value


def launch_count():
    return int(_lib.load().pcb_launch_count())


def timing_begin(max_launches=1 << 16, stride=1):
    """Bracket every `stride`-th node-jobs launch with a CUDA event pair (library-owned) until `timing_end`."""
    lib = _lib.require_device()
    _lib.check(lib.pcb_timing_sample(int(stride), None))
    _lib.check(lib.pcb_timing_begin(int(max_launches)))


def timing_seen():
    """Node-jobs launches since `timing_begin`, bracketed or not (read it before `timing_end`)."""
    n = C.c_int64(0)
    _lib.check(_lib.require_device().pcb_timing_sample(0, C.byref(n)))
    return n.value


def timing_end():
    """(summed kernel milliseconds, number of launches recorded) since `timing_begin`; synchronises the device."""
    ms, n = C.c_double(0.0), C.c_int64(0)
    _lib.check(_lib.require_device().pcb_timing_end(C.byref(ms), C.byref(n)))
    return ms.value, n.value


def plan_stats():
    """(seconds enqueueing, seconds waiting for the stream, flushes) over all synchronising flushes of this process."""
    a, b, n = C.c_double(0.0), C.c_double(0.0), C.c_int64(0)
    _lib.check(_lib.require_device().pcb_plan_stats(C.byref(a), C.byref(b), C.byref(n)))
    return a.value, b.value, n.value
