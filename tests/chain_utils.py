"""Helpers shared by the chain-level parity tests."""

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_chain(name):
    return np.load(os.path.join(GOLDEN, "chain_%s.npz" % name))


def signature(tree_dict, n_data, name_of=lambda dp: dp.idx):
    """Same canonical topology arrays as tests/golden/make_golden_chain.py::tree_signature."""
    node_of = np.full(n_data, -2, dtype=np.int64)
    for node, dps in tree_dict["node_data"].items():
        if node == "root":
            continue
        for dp in dps:
            node_of[name_of(dp)] = node
    rev = tree_dict["node_idx_rev"]
    parent_node = {}
    for s, t in tree_dict["graph"]:
        parent_node[rev[t]] = rev[s]
    mins = {}
    for node, dps in tree_dict["node_data"].items():
        if node not in ("root", -1) and dps:
            mins[node] = min(name_of(dp) for dp in dps)
    own = np.array([mins.get(node_of[i], -1) for i in range(n_data)], dtype=np.int64)
    par = np.array(
        [-1 if node_of[i] == -1 else mins.get(parent_node.get(node_of[i], "root"), -1) for i in range(n_data)],
        dtype=np.int64,
    )
    return own, par


def compare_trace(trace, g, rtol, what):
    n = g["values"].shape[0]
    assert len(trace) == len(g["alpha"]), (len(trace), len(g["alpha"]))
    for i, e in enumerate(trace):
        own, par = signature(e["tree"], n)
        assert np.array_equal(own, g["own"][i]) and np.array_equal(par, g["par"][i]), "%s: topology differs at trace entry %d" % (what, i)
        assert abs(e["alpha"] - g["alpha"][i]) <= rtol * abs(g["alpha"][i]), "%s: alpha differs at entry %d" % (what, i)
        assert abs(e["log_p_one"] - g["log_p_one"][i]) <= rtol * abs(g["log_p_one"][i]), "%s: log_p_one differs at entry %d (%r vs %r)" % (
            what, i, e["log_p_one"], g["log_p_one"][i])
