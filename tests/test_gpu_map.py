"""GPU: MAP assignment kernel (pcb_map_trees_*) against the reference fixtures and the oracle."""

import numpy as np
import pytest

from oracle import map as omap
from tests.map_utils import case_paths, load_case, tree_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", case_paths(), ids=lambda p: p.split("/")[-1][:-4])
def test_kernel_matches_reference(path):
    from phyclone_b200 import ops

    c = load_case(path)
    N = len(c["names"])
    max_idx, r_max = ops.map_trees(c["log_p"], [0, N], c["post_order"], c["child_off"], c["child_idx"],
                                   want_log_r_max=True)
    np.testing.assert_array_equal(r_max, c["log_r_max"])  # additions and comparisons only: bit-exact
    np.testing.assert_array_equal(max_idx, c["max_idx"])


def _random_forest(rng, n_trees, S, G, max_nodes):
    log_p, tree_off, post, child_off, child_idx, per_tree = [], [0], [], [0], [], []
    for _ in range(n_trees):
        n = int(rng.integers(1, max_nodes + 1))
        base = len(log_p)
        parent = [-1] + [int(rng.integers(0, k)) for k in range(1, n)]
        kids = [[] for _ in range(n)]
        for k in range(1, n):
            kids[parent[k]].append(k)
        for k in range(n):
            rng.shuffle(kids[k])
        order, stack = [], [(0, False)]
        while stack:
            v, seen = stack.pop()
            if seen:
                order.append(v)
            else:
                stack.append((v, True))
                stack.extend((k, False) for k in kids[v])
        mode = rng.integers(3)
        for _k in range(n):
            if mode == 0:
                log_p.append(np.log(rng.uniform(0.1, 1000, size=(S, G))))
            elif mode == 1:
                log_p.append(rng.integers(-2, 3, size=(S, G)).astype(np.float64))
            else:
                t = -((np.arange(G)[None] - rng.uniform(0, G / 4, (S, 1))) ** 2) * rng.uniform(0.1, 3, (S, 1))
                t[rng.random((S, G)) < 0.05] = -np.inf
                log_p.append(t)
        post.extend(base + v for v in order)
        tree_off.append(len(post))
        for k in range(n):
            child_idx.extend(base + c for c in kids[k])
            child_off.append(len(child_idx))
        per_tree.append((base, kids, order))
    return np.array(log_p), tree_off, post, child_off, child_idx, per_tree


@pytest.mark.parametrize("S,G,n_trees,max_nodes", [(4, 101, 40, 30), (8, 101, 12, 51), (20, 201, 3, 40), (1, 1, 3, 4),
                                                  (3, 2, 5, 6), (2, 1100, 2, 5)])
def test_batched_forest_matches_oracle(S, G, n_trees, max_nodes):
    from phyclone_b200 import ops

    rng = np.random.default_rng(S * 1000 + G)
    log_p, tree_off, post, child_off, child_idx, per_tree = _random_forest(rng, n_trees, S, G, max_nodes)
    max_idx, r_max = ops.map_trees(log_p, tree_off, post, child_off, child_idx, want_log_r_max=True)
    for base, kids, order in per_tree:
        n = len(kids)
        want_idx, want_r = omap.map_tree(log_p[base:base + n], kids, order)
        np.testing.assert_array_equal(r_max[base:base + n], want_r)
        np.testing.assert_array_equal(max_idx[base:base + n], want_idx)


def test_malformed_structure_is_rejected():
    from phyclone_b200 import ops

    lp = np.zeros((2, 2, 5))
    with pytest.raises(RuntimeError):
        ops.map_trees(lp, [0, 2], [1, 0], [0, 1, 1], [7])  # child index out of range
    with pytest.raises(RuntimeError):
        ops.map_trees(lp, [0, 1], [1, 0], [0, 1, 1], [1])  # tree_off does not cover the nodes


@pytest.mark.parametrize("path", case_paths()[:4], ids=lambda p: p.split("/")[-1][:-4])
def test_product_trees_on_device(path):
    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.process_trace.map import map_trees
    from phyclone_b200.tiles import TileStore
    from phyclone_b200.tree.tree import Tree

    c = load_case(path)
    values = c["values"]
    store = TileStore(values)
    data = make_data_points(values)
    tree = Tree.from_dict(tree_dict(c, data, store.prior), store)
    (ccf, prev), (ccf2, _) = map_trees([tree, tree.copy()])
    for k, n in enumerate(int(x) for x in c["names"][1:]):
        np.testing.assert_array_equal(ccf[n], c["ccf"][k])
        np.testing.assert_array_equal(prev[n], c["clonal_prev"][k])
        np.testing.assert_array_equal(ccf2[n], c["ccf"][k])
