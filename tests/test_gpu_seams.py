"""GPU: the small reference seams outside the sampler's own data path -- utils/math.py conv_log,
fft_convolve_two_children (G >= 1000), non_log_conv, and the host-buffer `TreeNode` drop-in -- against vectors
generated from the reference (tests/golden/seams.npz, numerics.npz)."""

import os

import numpy as np
import pytest

from oracle import numerics as onum
from tests.parity import assert_close, assert_close_bucketed

pytestmark = pytest.mark.gpu

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seams.npz"))


def test_conv_log_and_non_log_conv():
    from phyclone_b200.utils import math as pm

    for k in ("1", "2", "30", "101", "wide"):
        x, y = Z["convlog_x_" + k], Z["convlog_y_" + k]
        assert_close(pm.conv_log(x, y), Z["convlog_out_" + k], 1e-12, "conv_log " + k)
        ans = np.zeros(len(x))
        assert pm.conv_log(x, y, ans) is ans
        assert_close(ans, Z["convlog_out_" + k], 1e-12)
        assert_close_bucketed(pm.non_log_conv(x, y), Z["nonlog_out_" + k], what="non_log_conv " + k, max_boundary=8)
    xb = np.stack([Z["convlog_x_101"], Z["convlog_y_101"]])
    yb = np.stack([Z["convlog_y_101"], Z["convlog_x_101"]])
    from phyclone_b200 import ops

    both = ops.conv_log(xb, yb)
    assert_close(both[0], Z["convlog_out_101"], 1e-12)
    assert_close(both[1], Z["convlog_out_101"], 1e-12)  # commutative up to rounding


@pytest.mark.parametrize("G", [1024, 1088])
def test_large_grid_direct_sum_vs_reference_fft(G):
    """For G >= 1000 the reference evaluates the same contract with an FFT (utils/math.py:241-263); the kernel's
    direct sum must match the oracle to 1e-9 and the FFT result up to the FFT's own round-off."""
    from phyclone_b200.tree import utils as tu
    from phyclone_b200.utils import math as pm

    a, b, want = Z["fft_a_%d" % G], Z["fft_b_%d" % G], Z["fft_out_%d" % G]
    got = pm.fft_convolve_two_children(a, b)
    assert_close(got, onum.conv_pair(a, b), 1e-9, "direct vs oracle")
    shift = a.max(1, keepdims=True) + b.max(1, keepdims=True)
    assert np.abs(np.exp(got - shift) - np.exp(want - shift)).max() <= 1e-12 * np.exp(want - shift).max()
    np.testing.assert_array_equal(tu._convolve_two_children(a, b), got)
    kids = [a, b, a[::-1].copy()]
    assert_close(tu.compute_log_S(kids), onum.log_s(kids), 1e-9, "log_S at G=%d" % G)


@pytest.mark.parametrize("S,G", [(2, 1089), (3, 1500), (4, 2001), (2, 4096), (1, 8192)])
def test_grids_beyond_1088(S, G):
    """Grids wider than 2 x 32 strips: lanes walk several strip pairs per fold step (node_jobs_kernel<17, false, true>).
    Pair convolution, three-child fold + scan and the node update against the oracle (1e-9); at G = 2001 also
    against scipy's FFT evaluation, which is what the reference runs there (utils/math.py:241-263), up to the FFT's
    own round-off; peaked rows exercise the 1e-100 floor."""
    from phyclone_b200 import ops
    from phyclone_b200.tree import utils as tu

    rng = np.random.default_rng(G)
    a = np.log(rng.uniform(0.1, 1000, size=(S, G)))
    b = np.log(rng.uniform(0.1, 1000, size=(S, G)))
    c = np.log(rng.uniform(0.1, 1000, size=(S, G)))
    peaked = -0.5 * ((np.arange(G) - G * 0.3) / 3.0) ** 2 + np.zeros((S, 1))  # rows spanning >> 745 nats
    got = ops.np_conv_dims(a, b)
    assert_close(got, onum.conv_pair(a, b), 1e-9, "pair conv at G=%d" % G)
    assert_close_bucketed(ops.np_conv_dims(peaked, peaked + 1.0), onum.conv_pair(peaked, peaked + 1.0),
                          what="peaked pair at G=%d" % G, max_boundary=8)
    kids = [a, b, c, peaked]
    assert_close_bucketed(tu.compute_log_S(kids), onum.log_s(kids), what="log_S at G=%d" % G, max_boundary=8)
    log_p = np.log(rng.uniform(0.1, 10, size=(1, S, G)))
    upd = ops.node_update_batch(log_p, [np.stack([a, b, c])])[0]
    assert_close(upd, onum.node_log_r(log_p[0], [a, b, c]), 1e-9, "node update at G=%d" % G)
    if G == 2001:
        from scipy.signal import fftconvolve

        shift = a.max(1, keepdims=True) + b.max(1, keepdims=True)
        fft = np.stack([fftconvolve(np.exp(b[s] - b[s].max()), np.exp(a[s] - a[s].max()))[:G] for s in range(S)])
        assert np.abs(np.exp(got - shift) - fft).max() <= 1e-11 * fft.max()


def test_grid_beyond_supported_size_is_an_error():
    from phyclone_b200 import ops

    with pytest.raises(RuntimeError):
        ops.np_conv_dims(np.zeros((1, 8300)), np.zeros((1, 8300)))


def test_tree_node_dropin(golden):
    from phyclone_b200.tree.tree_node import TreeNode

    class DP:
        def __init__(self, idx, value):
            self.idx, self.value = idx, value

    tiles = golden["mixing_tiles"]
    S, G = tiles[0].shape
    lp = -np.log(G)
    kids = []
    for i in (1, 4, 6):
        n = TreeNode((S, G), lp, i)
        n.add_data_point(DP(i, tiles[i]))
        n.update_node_from_child_r_vals([])
        np.testing.assert_array_equal(n.log_r, lp + tiles[i])
        kids.append(n)
    parent = TreeNode((S, G), lp, 9)
    parent.add_data_point_list([DP(0, tiles[0]), DP(2, tiles[2])])
    parent.update_node_from_child_r_vals([k.log_r for k in kids])
    want = onum.node_log_r(parent.log_p, [k.log_r for k in kids])
    assert_close_bucketed(parent.log_r, want, what="TreeNode.log_r", max_boundary=8)
    with pytest.raises(AssertionError):
        parent.add_data_point(DP(0, tiles[0]))
    parent.remove_data_point(DP(2, tiles[2]))
    np.testing.assert_array_equal(parent.log_p, (np.full((S, G), lp) + tiles[0] + tiles[2]) - tiles[2])
    cp = parent.copy()
    assert cp == parent and cp.log_r is not parent.log_r and cp.data_points == {0}
    assert sorted(parent.to_dict()) == ["log_R", "log_p", "node_id"]
    for k in range(int(golden["logs_count"])):  # the reference's own node updates
        n = TreeNode((S, G), 0.0, k)
        n.log_p[...] = golden["logs_logp_%d" % k]
        n.update_node_from_child_r_vals([tiles[i] for i in golden["logs_idx_%d" % k]])
        assert_close_bucketed(n.log_r, golden["logs_logr_%d" % k], what="golden node %d" % k, max_boundary=8)


def test_the_ctypes_stub_printed_in_integration_md_works():
    """INTEGRATION.md shows the binding a PhyClone maintainer would paste into phyclone/tree/utils.py; run exactly that
    text against the built library so that the document cannot drift from the ABI."""
    import re

    from phyclone_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    stub = re.findall(r"```python\n(.*?)```", text, flags=re.S)[0]
    assert "pcb_np_conv_dims_host" in stub and "def compute_log_S" in stub
    stub = stub.replace('ctypes.CDLL("libphyclone_b200.so")', "ctypes.CDLL(%r)" % _lib.LIB_PATH)
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)  # noqa: S102 -- our own documentation
    rng = np.random.default_rng(0)
    a, b, c = (np.log(rng.uniform(0.1, 1000, (8, 101))) for _ in range(3))
    assert_close(ns["_np_conv_dims"](a, b), onum.conv_pair(a, b), 1e-9, "stub conv")
    assert_close(ns["_sub_compute_S"](a), onum.prefix_lse(a), 1e-9, "stub scan")
    assert_close(ns["compute_log_S"]([a, b, c]), onum.log_s([a, b, c]), 1e-9, "stub log_S")
    assert ns["compute_log_S"]([]) == 0.0


def test_the_map_stub_printed_in_integration_md_works():
    """Same for the MAP binding (second code block of INTEGRATION.md), on a tree solved by the reference."""
    import re

    import networkx as nx

    from phyclone_b200 import _lib
    from tests.map_utils import case_paths, load_case

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    blocks = re.findall(r"```python\n(.*?)```", open(os.path.join(root, "INTEGRATION.md")).read(), flags=re.S)
    ns = {"nx": nx}
    exec(compile(blocks[0].replace('ctypes.CDLL("libphyclone_b200.so")', "ctypes.CDLL(%r)" % _lib.LIB_PATH),
                 "INTEGRATION.md#1", "exec"), ns)  # noqa: S102
    exec(compile(blocks[1], "INTEGRATION.md#2", "exec"), ns)  # noqa: S102
    c = load_case(case_paths()[1])
    graph = nx.DiGraph()
    for v, kids in enumerate(c["children"]):
        graph.add_node(v, log_p=c["log_p"][v])
    for v, kids in enumerate(c["children"]):
        for k in kids:
            graph.add_edge(v, k)
    got = ns["map_indices"](graph, 0)
    for v in range(len(c["children"])):
        np.testing.assert_array_equal(got[v], c["max_idx"][v])
