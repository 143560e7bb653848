"""CPU oracle for the PhyClone tree-node likelihood path  --  TEST INFRASTRUCTURE ONLY.

A plain-NumPy restatement of the reference arithmetic for the hot path named in
BASELINE.json.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this
module; nothing under `phyclone_b200/` does.

Parity status: PINNED.  Every function below is checked against the reference's
own implementation (imported from /root/reference in the build container, with
the two content-hash LRU caches bypassed -- SURVEY.md Appendix A.9) by
`tests/golden/make_golden.py`; the resulting vectors are committed under
`tests/golden/` and re-checked by `tests/test_oracle_golden.py` on every run.

All `file:line` citations are relative to /root/reference/phyclone/.
"""

import math

import numpy as np
from scipy.special import gammaln

FLOOR = 1e-100  # tree/utils.py:77


# --------------------------------------------------------------------------
# a10  _np_conv_dims                                   tree/utils.py:60-85
# --------------------------------------------------------------------------
def conv_pair(c1, c2):
    """Row-wise max-shifted truncated linear convolution of two [S,G] log tiles."""
    c1 = np.asarray(c1, dtype=np.float64)
    c2 = np.asarray(c2, dtype=np.float64)
    S, G = c1.shape
    m1 = c1.max(axis=1, keepdims=True)
    m2 = c2.max(axis=1, keepdims=True)
    x1 = np.exp(c1 - m1)
    x2 = np.exp(c2 - m2)
    y = np.empty((S, G))
    for s in range(S):
        y[s] = np.convolve(x2[s], x1[s])[:G]
    y[y <= 0] = FLOOR
    np.log(y, out=y)
    y += m1
    y += m2
    return y


# --------------------------------------------------------------------------
# a11  _sub_compute_S                                  tree/utils.py:25-30
# --------------------------------------------------------------------------
def prefix_lse(log_d):
    """Sequential prefix log-sum-exp along the grid axis."""
    log_d = np.asarray(log_d, dtype=np.float64)
    out = np.empty_like(log_d)
    for s in range(log_d.shape[0]):
        np.logaddexp.accumulate(log_d[s], out=out[s])
    return out


# --------------------------------------------------------------------------
# a8 / a7  compute_log_D, compute_log_S               tree/utils.py:7-47
# --------------------------------------------------------------------------
def log_d_chain(children):
    """Left fold conv(c0,c1), conv(c2,acc), ... in the GIVEN child order."""
    n = len(children)
    if n == 0:
        return 0.0
    if n == 1:
        return np.asarray(children[0], dtype=np.float64)
    acc = conv_pair(children[0], children[1])
    for j in range(2, n):
        acc = conv_pair(children[j], acc)
    return acc


def log_s(children):
    """Uncached body of compute_log_S (order-sensitive)."""
    if len(children) == 0:
        return 0.0
    return np.ascontiguousarray(prefix_lse(log_d_chain(children)))


# --------------------------------------------------------------------------
# a6  TreeNode.update_node_from_child_r_vals          tree/tree_node.py:61-71
# --------------------------------------------------------------------------
def node_log_r(log_p, children):
    log_p = np.asarray(log_p, dtype=np.float64)
    if len(children) == 0:
        return log_p.copy()
    return log_p + log_s(children)


# --------------------------------------------------------------------------
# a15 (tile part)  Sum_s LSE_g root.log_r, Sum_s root.log_r[s,-1]
#                                                 tree/distributions.py:197-200
# --------------------------------------------------------------------------
def lse_1d(x):
    """utils/math.py:102-118 (log_sum_exp).  The reference body is Numba-compiled, i.e. scalar
    libm exp/log, which is what math.exp/math.log call; NumPy's SIMD exp can differ in the last
    bit, and one ulp in a particle weight is enough to change how many uniforms
    Generator.multinomial consumes ("RNG-stream identity")."""
    x = np.asarray(x, dtype=np.float64)
    m = float(np.max(x))
    if math.isinf(m):
        return m
    total = 0.0
    for v in x.tolist():
        total += math.exp(v - m)
    return math.log(total) + m


def root_terms(root_tile):
    root_tile = np.asarray(root_tile, dtype=np.float64)
    lse_sum = 0.0
    last_sum = 0.0
    for s in range(root_tile.shape[0]):
        lse_sum += lse_1d(root_tile[s])
        last_sum += root_tile[s, -1]
    return lse_sum, last_sum


# --------------------------------------------------------------------------
# a17 / a20  log_normalize, exp_normalize            utils/math.py:37-59,121
# --------------------------------------------------------------------------
def log_normalize(log_p):
    log_p = np.asarray(log_p, dtype=np.float64)
    return log_p - lse_1d(log_p)


def exp_normalize(log_p):
    """utils/math.py:37-59; Numba body: elementwise libm exp and a sequential sum."""
    log_p = np.asarray(log_p, dtype=np.float64)
    log_norm = lse_1d(log_p)
    p = np.array([math.exp(v) for v in (log_p - log_norm).tolist()])
    total = 0.0
    for v in p.tolist():
        total += v
    p = p / total
    return p, log_norm


def candidate_q(log_p_cand):
    """smc/kernels/semi_adapted.py:151-163 -> (log_q, q)."""
    log_q = log_normalize(log_p_cand)
    q = np.exp(log_q)
    q_sum = np.sum(q)
    assert abs(1 - q_sum) < 1e-6
    return log_q, q / q_sum


# --------------------------------------------------------------------------
# a18  ParticleSwarm                                  smc/swarm/swarm.py:21-61
# --------------------------------------------------------------------------
def swarm_stats(log_w):
    log_w = np.asarray(log_w, dtype=np.float64)
    log_z = lse_1d(log_w)
    log_W = log_w - log_z
    W = np.exp(log_W)
    W /= W.sum()
    ess = 1.0 / np.square(W).sum()
    return log_z, log_W, W, ess


# --------------------------------------------------------------------------
# a4  DataPoint.outlier_marginal_prob                 data/base.py:31-37
# --------------------------------------------------------------------------
def outlier_marginal(value):
    value = np.asarray(value, dtype=np.float64)
    lp = -np.log(value.shape[1])
    sub = prefix_lse(value + lp) + lp
    total = 0.0
    for s in range(value.shape[0]):
        m = sub[s].max()
        total += np.log(np.exp(sub[s] - m).sum()) + m  # scipy logsumexp
    return total


# --------------------------------------------------------------------------
# a1  get_major_cn_prior                              data/pyclone.py:269-307
# --------------------------------------------------------------------------
def major_cn_prior(major_cn, minor_cn, normal_cn, error_rate=1e-3):
    if major_cn < minor_cn:
        raise ValueError("major_cn < minor_cn")
    total = major_cn + minor_cn
    cn, mu = [], []
    for x in range(1, major_cn + 1):
        cn.append((normal_cn, normal_cn, total))
        mu.append((error_rate, error_rate, min(1 - error_rate, x / total)))
    after = (normal_cn, total, total)
    if after not in cn:
        cn.append(after)
        mu.append((error_rate, error_rate, min(1 - error_rate, 1 / total)))
    n = len(cn)
    log_pi = np.zeros(n) - math.log(n)  # log_normalize of zeros
    return np.array(cn, dtype=np.int64), np.array(mu, dtype=np.float64), log_pi


# --------------------------------------------------------------------------
# a2  likelihood grid                      data/pyclone.py:323-436, utils/math.py:126-209
# --------------------------------------------------------------------------
def _log_beta(a, b):
    with np.errstate(invalid="ignore", divide="ignore"):
        out = gammaln(a) + gammaln(b) - gammaln(a + b)
    return np.where((a <= 0) | (b <= 0), -np.inf, out)


def _log_binom_coeff(n, x):
    return gammaln(n + 1.0) - gammaln(x + 1.0) - gammaln(n - x + 1.0)


def likelihood_grid(ref_counts, alt_counts, tumour_content, cn, mu, log_pi, n_geno, grid_size, density, precision):
    """Per-(mutation, sample) log-likelihood over the CCF grid.

    Shapes: ref_counts/alt_counts/tumour_content/n_geno [M,S]; cn, mu [M,S,C,3];
    log_pi [M,S,C].  Returns [M,S,G].  `ref_counts` is SampleDataPoint.a and
    `alt_counts` is SampleDataPoint.b; the pdf is evaluated at n=a+b, x=b
    (data/pyclone.py:400).
    """
    a = np.asarray(ref_counts, dtype=np.float64)[..., None, None]
    b = np.asarray(alt_counts, dtype=np.float64)[..., None, None]
    t = np.asarray(tumour_content, dtype=np.float64)[..., None, None]
    cn = np.asarray(cn, dtype=np.float64)[:, :, None, :, :]  # [M,S,1,C,3]
    mu = np.asarray(mu, dtype=np.float64)[:, :, None, :, :]
    log_pi = np.asarray(log_pi, dtype=np.float64)[:, :, None, :]
    n_geno = np.asarray(n_geno)
    f = np.linspace(0, 1, grid_size)[None, None, :, None]  # [1,1,G,1]
    w = [1 - t + 0 * f, t * (1 - f), t * f]
    e_vaf = 0.0
    norm = 0.0
    for i in range(3):
        e_cn = w[i] * cn[..., i]
        e_vaf = e_vaf + e_cn * mu[..., i]
        norm = norm + e_cn
    with np.errstate(invalid="ignore", divide="ignore"):
        e_vaf = e_vaf / norm  # [M,S,G,C]; padded genotype slots are 0/0 and masked below
    n = a + b
    x = b
    if density == "beta-binomial":
        if precision is None:
            raise Exception("Precision must be set when using Beta-Binomial.")
        aa = e_vaf * precision
        bb = precision - aa
        ll = _log_binom_coeff(n, x) + _log_beta(aa + x, bb + n - x) - _log_beta(aa, bb)
    elif density == "binomial":
        p = e_vaf
        with np.errstate(invalid="ignore", divide="ignore"):
            body = x * np.log(p) + (n - x) * np.log(1 - p)
        body = np.where(p == 0, np.where(x == 0, 0.0, -np.inf), body)
        body = np.where(p == 1, np.where(x == n, 0.0, -np.inf), body)
        ll = _log_binom_coeff(n, x) + body
    else:
        raise ValueError(density)
    ll = log_pi + ll
    C = ll.shape[-1]
    valid = np.arange(C)[None, None, None, :] < n_geno[:, :, None, None]
    ll = np.where(valid, ll, -np.inf)
    m = ll.max(axis=-1)
    with np.errstate(invalid="ignore"):
        tot = np.where(valid, np.exp(ll - m[..., None]), 0.0).sum(axis=-1)
    out = np.log(tot) + m
    return np.where(np.isinf(m), m, out)


def cluster_sum(grids, cluster_offsets):
    """a3 data/pyclone.py:89-94: value[k] = sum of member grids (np.sum over axis 0)."""
    K = len(cluster_offsets) - 1
    out = np.empty((K,) + grids.shape[1:])
    for k in range(K):
        out[k] = np.sum(grids[cluster_offsets[k] : cluster_offsets[k + 1]], axis=0)
    return out


# --------------------------------------------------------------------------
# optional device-resampling mode: inverse-CDF ancestors from supplied uniforms
# --------------------------------------------------------------------------
def ancestors_from_uniforms(W, uniforms):
    cdf = np.cumsum(np.asarray(W, dtype=np.float64))
    cdf[-1] = np.inf
    return np.searchsorted(cdf, np.asarray(uniforms, dtype=np.float64), side="right").astype(np.int64)


def conv_log(log_x, log_y):
    """utils/math.py:212-238: out[k] = LSE_{j<=k}(x[j] + y[k-j]) with a per-output max shift, no floor."""
    n = len(log_x)
    out = np.empty(n)
    for k in range(n):
        v = log_x[: k + 1] + log_y[k::-1]
        m = v.max()
        acc = 0.0
        for t in np.exp(v - m):
            acc += t
        out[k] = np.log(acc) + m
    return out


def non_log_conv(child_log_r, prev_log_d):
    """utils/math.py:266-282: conv_pair for one row; np.convolve(R, D) there is np.convolve(x2, x1) of
    conv_pair(c1=D, c2=R), including the order in which the two maxima are added back."""
    return conv_pair(np.asarray(prev_log_d)[None, :], np.asarray(child_log_r)[None, :])[0]
