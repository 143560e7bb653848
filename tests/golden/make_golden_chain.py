"""Record traces of the UNMODIFIED reference sampler (build container only).

Run:  python tests/golden/make_golden_chain.py
Reference configuration for parity: `compute_log_S` and
`_convolve_two_children` content-hash caches bypassed (they are keyed order-insensitively
and change the reference's own trajectory); graph container = oracle/refshim/rustworkx.
Writes tests/golden/chain_*.npz: inputs (data tiles) + per-trace-entry alpha, log_p_one,
parent vector (by data point) and outlier flags.
"""

import io
import os
import sys
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle", "refshim"), "/root/reference"]

import numpy as np  # noqa: E402

import phyclone.tree.tree_node as ref_node  # noqa: E402
import phyclone.tree.utils as tu  # noqa: E402
from phyclone.data.base import DataPoint as RefDataPoint  # noqa: E402
from phyclone.data.pyclone import load_data  # noqa: E402
from phyclone.run import run_phyclone_chain  # noqa: E402

from oracle.synthetic import make_synthetic  # noqa: E402


def bypass_tile_caches():
    tu._convolve_two_children = tu._np_conv_dims
    body = tu.compute_log_S.__wrapped__

    def uncached(children):
        return body(np.array(children, order="C"))

    tu.compute_log_S = uncached
    ref_node.compute_log_S = uncached


def tree_signature(tree_dict, n_data):
    """Topology as arrays indexed by data point: node label of each data point (-1 outlier) and,
    per node, the smallest data index of its parent node (or -1 for roots)."""
    node_of = np.full(n_data, -2, dtype=np.int64)
    for node, dps in tree_dict["node_data"].items():
        if node == "root":
            continue
        for dp in dps:
            node_of[dp.idx] = node
    rev = tree_dict["node_idx_rev"]
    parent_node = {}
    for s, t in tree_dict["graph"]:
        parent_node[rev[t]] = rev[s]
    # canonical: per data point, the min data idx in its node and in its parent's node
    mins = {}
    for node, dps in tree_dict["node_data"].items():
        if node not in ("root", -1) and dps:
            mins[node] = min(dp.idx for dp in dps)
    own = np.array([mins.get(node_of[i], -1) for i in range(n_data)], dtype=np.int64)
    par = np.array([
        -1 if node_of[i] == -1 else mins.get(parent_node.get(node_of[i], "root"), -1) for i in range(n_data)
    ], dtype=np.int64)
    return own, par


def record(name, data, seed, iters, particles, outlier_prob, proposal="semi-adapted", burnin=1, subtree_update_prob=0.0):
    rng = np.random.default_rng(seed)
    with redirect_stdout(io.StringIO()):
        res = run_phyclone_chain(burnin, True, 1.0, data, float("inf"), iters, particles, 1, 1, outlier_prob, 10**9,
                                 proposal, 0.5, rng, ["s"], 1, 0, subtree_update_prob)
    n = len(data)
    tr = res["trace"]
    own, par = zip(*[tree_signature(e["tree"], n) for e in tr])
    out = {
        "values": np.array([d.value for d in data]),
        "outlier_prob": np.array([d.outlier_prob for d in data], dtype=np.float64),
        "outlier_prob_not": np.array([d.outlier_prob_not for d in data], dtype=np.float64),
        "alpha": np.array([e["alpha"] for e in tr]),
        "log_p_one": np.array([e["log_p_one"] for e in tr]),
        "own": np.array(own),
        "par": np.array(par),
        "meta": np.array([seed, iters, particles, burnin], dtype=np.int64),
        "outlier_prob_run": np.array(outlier_prob),
        "proposal": np.array(proposal),
        "rng_after": np.array(rng.random()),
        "subtree_update_prob": np.array(subtree_update_prob),
    }
    path = os.path.join(HERE, "chain_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; final log_p_one", tr[-1]["log_p_one"])


def main_subtree():
    """Traces with the subtree particle-Gibbs move on (mcmc/particle_gibbs.py:51-109, run.py:268-271)."""
    bypass_tile_caches()
    syn = make_synthetic(n_clusters=12, n_samples=3, grid_size=41, muts_per_cluster=4, depth=800, seed=0)
    data = [RefDataPoint(i, v) for i, v in enumerate(syn["values"][:10])]
    record("syn10_subtree", data, seed=2, iters=12, particles=12, outlier_prob=0, subtree_update_prob=0.5)
    lo, lon = np.log(1e-3) * 4, np.log1p(-1e-3) * 4
    data = [RefDataPoint(i, v, outlier_prob=lo, outlier_prob_not=lon) for i, v in enumerate(syn["values"][:8])]
    record("syn8_subtree_outliers", data, seed=4, iters=10, particles=10, outlier_prob=1e-3, subtree_update_prob=0.5)


def main():
    bypass_tile_caches()
    rng = np.random.default_rng(0)
    with redirect_stdout(io.StringIO()):
        data, _ = load_data("/root/reference/examples/data/mixing.tsv", rng, 1e-4, 0.4, False,
                            cluster_file="/root/reference/examples/data/mixing_clusters.tsv",
                            density="beta-binomial", grid_size=101, outlier_prob=0, precision=400)
    record("mixing", data, seed=1, iters=12, particles=20, outlier_prob=0)
    # reduced config #2: synthetic clusters, floor-exercising depth
    syn = make_synthetic(n_clusters=12, n_samples=3, grid_size=41, muts_per_cluster=4, depth=800, seed=0)
    data = [RefDataPoint(i, v) for i, v in enumerate(syn["values"])]
    record("syn12", data, seed=1, iters=8, particles=16, outlier_prob=0)
    # outlier modelling on (config #4 flavour)
    lo, lon = np.log(1e-3) * 4, np.log1p(-1e-3) * 4
    data = [RefDataPoint(i, v, outlier_prob=lo, outlier_prob_not=lon) for i, v in enumerate(syn["values"][:8])]
    record("syn8_outliers", data, seed=3, iters=8, particles=12, outlier_prob=1e-3)
    data = [RefDataPoint(i, v) for i, v in enumerate(syn["values"][:6])]
    record("syn6_fully", data, seed=5, iters=5, particles=8, outlier_prob=0, proposal="fully-adapted")
    record("syn6_boot", data, seed=6, iters=5, particles=8, outlier_prob=0, proposal="bootstrap")


if __name__ == "__main__":
    if "--subtree" in sys.argv:
        main_subtree()
    else:
        main()
