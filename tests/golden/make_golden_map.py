"""Generate MAP-assignment golden vectors from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden_map.py
Needs /root/reference, networkx and the rustworkx stand-in under oracle/refshim.  For a set of trees built with
the reference `Tree` API it stores the inputs (node log_p tiles, children in the order the reference folds them:
networkx successor order of `convert_rustworkx_to_networkx`, i.e. ascending edge index) and what
`process_trace/map.py` computed from them: log_R_max, max_idx, ccf and clonal-prevalence per node.
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(REPO, "oracle", "refshim"), "/root/reference"]

import numpy as np  # noqa: E402

from phyclone.data.base import DataPoint  # noqa: E402
from phyclone.process_trace import map as ref_map  # noqa: E402
from phyclone.tree import Tree  # noqa: E402


def random_tree(rng, data, grid_size, p_outlier=0.0, flat=False):
    tree = Tree(grid_size)
    roots = []
    for dp in data:
        u = rng.random()
        if u < p_outlier:
            tree.add_data_point_to_outliers(dp)
        elif roots and u < 0.35:
            tree.add_data_point_to_node(dp, roots[rng.integers(len(roots))])
        else:
            k = 0 if (flat or not roots) else int(rng.integers(0, min(len(roots), 3) + 1))
            kids = list(rng.choice(roots, k, replace=False)) if k else []
            node = tree.create_root_node(children=kids, data=[dp])
            roots = [r for r in roots if r not in kids] + [node]
    return tree


def case(rng, T, S, G, kind, **kw):
    if kind == "benign":
        values = np.log(rng.uniform(0.1, 1000, size=(T, S, G)))
    elif kind == "peaked":
        g = np.arange(G)[None, None, :]
        centre = rng.uniform(0, (G - 1) / 3, size=(T, S, 1))
        values = -((g - centre) ** 2) * rng.uniform(0.05, 2.0, size=(T, S, 1)) + rng.uniform(-50, 50, size=(T, S, 1))
    else:  # ties: coarse integers, so equal maxima are common and the tie-breaking rules decide
        values = rng.integers(-3, 3, size=(T, S, G)).astype(np.float64)
    data = [DataPoint(i, values[i], outlier_prob=1e-3, outlier_prob_not=1 - 1e-3) for i in range(T)]
    tree = random_tree(rng, data, (S, G), **kw)
    root = tree.root_node_name
    graph = ref_map.compute_map_tree_features(tree, root)
    ccf, prev = ref_map.get_map_node_ccfs_and_clonal_prev_dicts(tree)
    names = [root] + sorted(n for n in graph.nodes if n != root)
    index = {n: i for i, n in enumerate(names)}
    children = [[index[c] for c in graph.successors(n)] for n in names]
    out = {
        "values": values, "names": np.array([-7 if n == root else n for n in names]),
        "child_off": np.cumsum([0] + [len(c) for c in children]),
        "child_idx": np.array([c for cs in children for c in cs], dtype=np.int64),
        "log_p": np.array([graph.nodes[n]["log_p"] for n in names]),
        "log_r_max": np.array([graph.nodes[n]["log_R_max"] for n in names]),
        "max_idx": np.array([graph.nodes[n]["max_idx"] for n in names]),
        "ccf": np.array([ccf[n] for n in names[1:]]), "clonal_prev": np.array([prev[n] for n in names[1:]]),
    }
    d = tree.to_dict()
    out["edges"] = np.array([tuple(e) for e in d["graph"]], dtype=np.int64).reshape(-1, 2)
    out["node_idx_keys"] = np.array([-7 if k == root else k for k in d["node_idx"]])
    out["node_idx_vals"] = np.array(list(d["node_idx"].values()))
    nd = [(k, dp.idx) for k, lst in d["node_data"].items() for dp in lst]
    out["node_data"] = np.array(nd, dtype=np.int64).reshape(-1, 2)
    return out


def main():
    rng = np.random.default_rng(20240607)
    specs = [("benign", 8, 3, 21, {}), ("benign", 14, 4, 101, {}), ("peaked", 12, 2, 64, {}),
             ("ties", 10, 3, 17, {}), ("ties", 16, 2, 33, {"p_outlier": 0.2}), ("benign", 6, 2, 40, {"flat": True}),
             ("peaked", 20, 8, 101, {"p_outlier": 0.1})]
    for n, (kind, T, S, G, kw) in enumerate(specs):
        out = case(rng, T, S, G, kind, **kw)
        path = os.path.join(HERE, "map_%d_%s.npz" % (n, kind))
        np.savez_compressed(path, **out)
        print(path, "nodes", len(out["names"]), "max_idx range", out["max_idx"].min(), out["max_idx"].max())


if __name__ == "__main__":
    main()
