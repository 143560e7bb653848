"""Generate golden vectors from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only mount) and the rustworkx stand-in under
oracle/refshim (rustworkx itself is not installable here; ordering parity with
the real wheel is therefore unpinned -- the arithmetic below does not depend on
it).  The two content-hash LRU caches around compute_log_S /
_convolve_two_children are bypassed (`__wrapped__` bodies / _np_conv_dims), as
SURVEY.md Appendix A.9 requires for an order-faithful oracle.

Outputs small .npz files next to this script; they are committed, the GPU box
never sees /root/reference.
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, "oracle", "refshim"), "/root/reference"]

import numpy as np  # noqa: E402

import phyclone.tree.utils as tu  # noqa: E402
from phyclone.data.base import DataPoint as RefDataPoint  # noqa: E402
from phyclone.data.pyclone import load_data, load_pyclone_data  # noqa: E402
from phyclone.smc.swarm.swarm import ParticleSwarm  # noqa: E402
from phyclone.tree.tree_node import TreeNode  # noqa: E402
from phyclone.utils.math import exp_normalize, log_normalize, log_sum_exp  # noqa: E402


def uncached_log_s(children):
    """compute_log_S body with every pair conv routed to _np_conv_dims (no LRU)."""
    saved = tu._convolve_two_children
    tu._convolve_two_children = tu._np_conv_dims
    try:
        return tu.compute_log_S.__wrapped__(np.array(children, order="C"))
    finally:
        tu._convolve_two_children = saved


def benign_tile(rng, S, G):
    # recipe of tests/test_fft_convolve.py:52-53
    return np.log(rng.uniform(0.1, 1000, size=(S, G)))


def peaked_tile(rng, S, G, width):
    """Adversarial: row range >> 745 nats so exp() underflows and floors are hit."""
    g = np.arange(G)[None, :]
    centre = rng.uniform(0, G - 1, size=(S, 1))
    scale = rng.uniform(0.5, 2.0, size=(S, 1)) * width
    return -((g - centre) ** 2) * scale + rng.uniform(-50, 50, size=(S, 1))


def mixing_tiles():
    rng = np.random.default_rng(0)
    data, samples = load_data(
        "/root/reference/examples/data/mixing.tsv",
        rng,
        1e-4,
        0.4,
        False,
        cluster_file="/root/reference/examples/data/mixing_clusters.tsv",
        density="beta-binomial",
        grid_size=101,
        outlier_prob=0,
        precision=400,
    )
    return data


def gen_conv(out):
    rng = np.random.default_rng(12345)
    cases = {}
    for name, (S, G) in {"s4g101": (4, 101), "s8g101": (8, 101), "s20g201": (20, 201), "s1g11": (1, 11), "s3g37": (3, 37)}.items():
        a, b = benign_tile(rng, S, G), benign_tile(rng, S, G)
        cases["benign_" + name] = (a, b)
        a, b = peaked_tile(rng, S, G, 3.0), peaked_tile(rng, S, G, 3.0)
        cases["peaked_" + name] = (a, b)
    data = mixing_tiles()
    for i, j in [(0, 1), (1, 2), (3, 7), (9, 0), (5, 5)]:
        cases["mixing_%d_%d" % (i, j)] = (data[i].value, data[j].value)
    for k, (a, b) in cases.items():
        out["conv_in1_" + k] = a
        out["conv_in2_" + k] = b
        out["conv_out_" + k] = tu._np_conv_dims(a, b)
        out["scan_out_" + k] = tu._sub_compute_S(out["conv_out_" + k])
    return data


def gen_log_s(out, data):
    rng = np.random.default_rng(777)
    tiles = [d.value for d in data]
    k = 0
    for n_children in (1, 2, 3, 4, 5, 6, 9):
        for rep in range(3):
            idx = rng.permutation(len(tiles))[:n_children]
            children = [tiles[i] for i in idx]
            log_p = tiles[int(rng.integers(len(tiles)))] - np.log(101)
            node = TreeNode((4, 101), -np.log(101), 0)
            node.log_p[:] = log_p
            saved_node_fn = sys.modules["phyclone.tree.tree_node"].compute_log_S
            sys.modules["phyclone.tree.tree_node"].compute_log_S = uncached_log_s
            try:
                node.update_node_from_child_r_vals(children)
            finally:
                sys.modules["phyclone.tree.tree_node"].compute_log_S = saved_node_fn
            out["logs_idx_%d" % k] = idx.astype(np.int64)
            out["logs_out_%d" % k] = uncached_log_s(children)
            out["logs_logp_%d" % k] = log_p
            out["logs_logr_%d" % k] = node.log_r.copy()
            k += 1
    # benign 8x101 chains
    rng = np.random.default_rng(4242)
    for n_children in (2, 3, 7):
        children = [benign_tile(rng, 8, 101) for _ in range(n_children)]
        out["logsb_in_%d" % n_children] = np.array(children)
        out["logsb_out_%d" % n_children] = uncached_log_s(children)
    out["logs_count"] = np.array(k)
    out["mixing_tiles"] = np.array(tiles)


def gen_root_terms(out):
    rng = np.random.default_rng(99)
    tiles = rng.normal(-500, 200, size=(6, 8, 101))
    lse = np.array([sum(log_sum_exp(t[s, :]) for s in range(t.shape[0])) for t in tiles])
    last = np.array([sum(t[s, -1] for s in range(t.shape[0])) for t in tiles])
    out["root_in"] = tiles
    out["root_lse"] = lse
    out["root_last"] = last


def gen_likelihood(out):
    pyclone_data, samples = load_pyclone_data("/root/reference/examples/data/mixing.tsv")
    muts = list(pyclone_data.keys())[:40]
    S = len(samples)
    Cmax = 4
    M = len(muts)
    a = np.zeros((M, S), dtype=np.int64)
    b = np.zeros((M, S), dtype=np.int64)
    t = np.zeros((M, S))
    cn = np.zeros((M, S, Cmax, 3), dtype=np.int64)
    mu = np.zeros((M, S, Cmax, 3))
    log_pi = np.full((M, S, Cmax), -np.inf)
    ncfg = np.zeros((M, S), dtype=np.int32)
    bb = np.zeros((M, S, 101))
    bn = np.zeros((M, S, 101))
    bb21 = np.zeros((M, S, 21))
    for m, mut in enumerate(muts):
        dp = pyclone_data[mut]
        for s, sdp in enumerate(dp.sample_data_points):
            a[m, s], b[m, s], t[m, s] = sdp.a, sdp.b, sdp.t
            C = len(sdp.cn)
            ncfg[m, s] = C
            cn[m, s, :C] = sdp.cn
            mu[m, s, :C] = sdp.mu
            log_pi[m, s, :C] = sdp.log_pi
        bb[m] = dp.to_likelihood_grid("beta-binomial", 101, precision=400)
        bn[m] = dp.to_likelihood_grid("binomial", 101, precision=400)
        bb21[m] = dp.to_likelihood_grid("beta-binomial", 21, precision=50.0)
    out.update(lik_a=a, lik_b=b, lik_t=t, lik_cn=cn, lik_mu=mu, lik_log_pi=log_pi, lik_ncfg=ncfg)
    out.update(lik_bb101=bb, lik_bin101=bn, lik_bb21=bb21)
    # synthetic multi-genotype cases (major_cn up to 3) through the reference classes
    from phyclone.data.pyclone import DataPoint as RefMut, SampleDataPoint, get_major_cn_prior

    rng = np.random.default_rng(5)
    M2, S2 = 12, 3
    a2 = rng.integers(5, 400, size=(M2, S2))
    b2 = rng.integers(0, 200, size=(M2, S2))
    t2 = rng.uniform(0.3, 1.0, size=(M2, S2))
    cn2 = np.zeros((M2, S2, Cmax, 3), dtype=np.int64)
    mu2 = np.zeros((M2, S2, Cmax, 3))
    lp2 = np.full((M2, S2, Cmax), -np.inf)
    nc2 = np.zeros((M2, S2), dtype=np.int32)
    cnspec = np.zeros((M2, S2, 3), dtype=np.int64)
    o_bb = np.zeros((M2, S2, 101))
    o_bn = np.zeros((M2, S2, 101))
    for m in range(M2):
        sdps = []
        for s in range(S2):
            major = int(rng.integers(1, 4))
            minor = int(rng.integers(0, major + 1))
            cnspec[m, s] = (major, minor, 2)
            c, u, lp = get_major_cn_prior(major, minor, 2, 1e-3)
            C = len(c)
            nc2[m, s] = C
            cn2[m, s, :C], mu2[m, s, :C], lp2[m, s, :C] = c, u, lp
            sdps.append(SampleDataPoint(int(a2[m, s]), int(b2[m, s]), c, u, lp, float(t2[m, s])))
        dp = RefMut(["s%d" % i for i in range(S2)], sdps)
        o_bb[m] = dp.to_likelihood_grid("beta-binomial", 101, precision=400)
        o_bn[m] = dp.to_likelihood_grid("binomial", 101, precision=400)
    out.update(lik2_a=a2, lik2_b=b2, lik2_t=t2, lik2_cn=cn2, lik2_mu=mu2, lik2_log_pi=lp2, lik2_ncfg=nc2)
    out.update(lik2_bb101=o_bb, lik2_bin101=o_bn, lik2_cnspec=cnspec)


def gen_datapoint(out, data):
    out["dp_outlier_marginal"] = np.array([d.outlier_marginal_prob for d in data])
    rng = np.random.default_rng(31)
    vals = rng.normal(-300, 150, size=(5, 8, 101))
    out["dp2_values"] = vals
    out["dp2_outlier_marginal"] = np.array([RefDataPoint(i, v).outlier_marginal_prob for i, v in enumerate(vals)])


def gen_swarm(out):
    rng = np.random.default_rng(2024)
    for k, n in enumerate((100, 500, 7)):
        logw = rng.normal(-1000, 30 if k < 2 else 300, size=n)
        sw = ParticleSwarm()
        for w in logw:
            sw.add_particle(float(w), None)
        out["swarm_logw_%d" % k] = logw
        out["swarm_logZ_%d" % k] = np.array(sw.log_norm_const)
        out["swarm_logW_%d" % k] = sw.log_weights
        out["swarm_W_%d" % k] = sw.weights
        out["swarm_ess_%d" % k] = np.array(sw.ess)
    lp = rng.normal(-4000, 40, size=17)
    out["norm_in"] = lp
    out["norm_log"] = log_normalize(lp)
    p, ln = exp_normalize(lp)
    out["norm_p"] = p
    out["norm_lognorm"] = np.array(ln)


def main():
    out = {}
    data = gen_conv(out)
    gen_log_s(out, data)
    gen_root_terms(out)
    gen_likelihood(out)
    gen_datapoint(out, data)
    gen_swarm(out)
    path = os.path.join(HERE, "numerics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
