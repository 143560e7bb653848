"""GPU parity tests: the sm_100a kernels, called through the C ABI (ctypes), against the
NumPy oracle (oracle/numerics.py) and against the golden vectors generated from the
reference.  FP64 tolerance: 1e-9 relative (BASELINE.json north_star), reported in the
normal-range / boundary-band buckets of SURVEY.md Appendix A.6."""

import os

import numpy as np
import pytest

from oracle import numerics as onum
from tests.parity import assert_close, assert_close_bucketed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from phyclone_b200 import ops as _ops

    return _ops


def _conv_cases(golden):
    return sorted(k[len("conv_out_"):] for k in golden.files if k.startswith("conv_out_"))


VARIANTS = ["", "516", "708", "716", "908", "1304", "1308", "1316", "732", "11304", "11704"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_conv_pair_golden(ops, golden, variant, monkeypatch):
    if variant:
        monkeypatch.setenv("PCB_VARIANT", variant)
    else:
        monkeypatch.delenv("PCB_VARIANT", raising=False)
    from phyclone_b200 import _lib

    boundary = ran = 0
    for k in _conv_cases(golden):
        try:
            got = ops.np_conv_dims(golden["conv_in1_" + k], golden["conv_in2_" + k])
        except _lib.PcbError as e:  # this strip/lane variant cannot cover that grid size
            assert "layout" in str(e)
            continue
        ran += 1
        boundary += assert_close_bucketed(got, golden["conv_out_" + k], what="conv " + k, max_boundary=8)
    assert ran >= 5
    print("variant %r: %d cases, %d boundary-band entries differ" % (variant, ran, boundary))


def test_conv_pair_batched_matches_oracle(ops):
    rng = np.random.default_rng(7)
    for (B, S, G) in [(64, 8, 101), (16, 20, 201), (33, 4, 101), (5, 1, 11), (3, 3, 250), (2, 2, 999)]:
        a = np.log(rng.uniform(0.1, 1000, size=(B, S, G)))
        b = np.log(rng.uniform(0.1, 1000, size=(B, S, G)))
        got = ops.np_conv_dims(a, b)
        ref = np.array([onum.conv_pair(a[i], b[i]) for i in range(B)])
        assert_close(got, ref, what="conv batch %s" % ((B, S, G),))


def test_sub_compute_S_golden(ops, golden):
    for k in _conv_cases(golden):
        got = ops.sub_compute_S(golden["conv_out_" + k])
        assert_close(got, golden["scan_out_" + k], what="scan " + k)


def test_sub_compute_S_wide_range(ops):
    """rows spanning >> 745 nats: a linear-domain cumsum would lose the prefix, the scan must not."""
    rng = np.random.default_rng(3)
    x = np.cumsum(rng.uniform(-40, 60, size=(6, 8, 101)), axis=-1) - 3000
    x[0, 0, :] = -np.inf
    x[1, 1, :5] = -np.inf
    got = ops.sub_compute_S(x)
    ref = np.array([onum.prefix_lse(t) for t in x])
    assert_close(got, ref, what="wide scan")


def test_compute_log_S_and_node_update_golden(ops, golden):
    tiles = golden["mixing_tiles"]
    n = int(golden["logs_count"])
    child_lists = [tiles[golden["logs_idx_%d" % k]] for k in range(n)]
    got = ops.compute_log_S_batch(child_lists)
    boundary = 0
    for k in range(n):
        boundary += assert_close_bucketed(got[k], golden["logs_out_%d" % k], what="log_S %d" % k, max_boundary=8)
    log_p = np.array([golden["logs_logp_%d" % k] for k in range(n)])
    got = ops.node_update_batch(log_p, child_lists)
    for k in range(n):
        boundary += assert_close_bucketed(got[k], golden["logs_logr_%d" % k], what="log_r %d" % k, max_boundary=8)
    for c in (2, 3, 7):
        got = ops.compute_log_S_batch([golden["logsb_in_%d" % c]])[0]
        assert_close(got, golden["logsb_out_%d" % c], what="benign chain %d" % c)
    print("boundary-band entries:", boundary)


def test_node_update_leaf_and_order(ops, golden):
    tiles = golden["mixing_tiles"]
    log_p = tiles[:3] - np.log(101)
    fwd = [tiles[[7, 9, 0, 3, 5]], np.zeros((0, 4, 101)), tiles[[2]]]
    got = ops.node_update_batch(log_p, fwd)
    np.testing.assert_array_equal(got[1], log_p[1])  # leaf: log_r <- log_p
    for b in (0, 2):
        assert_close_bucketed(got[b], onum.node_log_r(log_p[b], list(fwd[b])), what="node %d" % b, max_boundary=8)
    # child order is part of the answer once floors occur
    rev = ops.compute_log_S_batch([tiles[[5, 3, 0, 9, 7]]])[0]
    assert_close_bucketed(rev, onum.log_s(list(tiles[[5, 3, 0, 9, 7]])), what="reversed", max_boundary=8)
    assert np.abs(rev - ops.compute_log_S_batch([fwd[0]])[0]).max() > 1.0


def test_tree_utils_dropin(golden):
    from phyclone_b200.tree import utils as tu

    tiles = golden["mixing_tiles"]
    assert tu.compute_log_S([]) == 0.0
    kids = [tiles[1], tiles[4], tiles[6]]
    assert_close_bucketed(tu.compute_log_S(kids), onum.log_s(kids), what="dropin", max_boundary=4)
    assert_close_bucketed(tu.compute_log_S.__wrapped__(np.array(kids)), onum.log_s(kids), what="wrapped", max_boundary=4)
    assert_close_bucketed(tu.compute_log_D(kids), onum.log_d_chain(kids), what="log_D", max_boundary=4)
    assert_close_bucketed(tu._convolve_two_children(kids[0], kids[1]), onum.conv_pair(kids[0], kids[1]), what="pair", max_boundary=4)
    assert_close(tu._sub_compute_S(kids[0]), onum.prefix_lse(kids[0]))
    assert tu.compute_log_S.cache_info().hits == 0


def test_root_terms(ops, golden):
    lse, last = ops.root_terms(golden["root_in"])
    assert_close(lse, golden["root_lse"], rtol=1e-13)
    assert_close(last, golden["root_last"], rtol=1e-15)


@pytest.mark.parametrize("prefix", ["lik", "lik2"])
def test_likelihood_grid(ops, golden, prefix):
    args = [golden[prefix + "_" + n] for n in ("a", "b", "t", "cn", "mu", "log_pi", "ncfg")]
    got = ops.likelihood_grid(*args, 101, "beta-binomial", 400)
    assert_close(got, golden[prefix + "_bb101"], what="beta-binomial")
    got = ops.likelihood_grid(*args, 101, "binomial", None)
    assert_close(got, golden[prefix + "_bin101"], what="binomial")
    if prefix == "lik":
        got = ops.likelihood_grid(*args, 21, "beta-binomial", 50.0)
        assert_close(got, golden["lik_bb21"], what="bb21")
    with pytest.raises(Exception):
        ops.likelihood_grid(*args, 101, "beta-binomial", None)


def test_cluster_sum_and_outlier_marginal(ops, golden):
    grids = golden["lik_bb101"]
    off = np.array([0, 3, 4, 11, 40])
    got = ops.cluster_sum(grids, off)
    np.testing.assert_array_equal(got, onum.cluster_sum(grids, off))
    got = ops.outlier_marginal(golden["mixing_tiles"])
    assert_close(got, golden["dp_outlier_marginal"], rtol=1e-12)
    got = ops.outlier_marginal(golden["dp2_values"])
    assert_close(got, golden["dp2_outlier_marginal"], rtol=1e-12)


def test_swarm_weights_and_normalisers(ops, golden):
    for k in range(3):
        log_z, log_W, W, ess = ops.swarm_weights(golden["swarm_logw_%d" % k])
        assert log_z == pytest.approx(float(golden["swarm_logZ_%d" % k]), rel=1e-13)
        assert_close(log_W, golden["swarm_logW_%d" % k], rtol=1e-11)
        assert_close(W, golden["swarm_W_%d" % k], rtol=1e-11)
        assert ess == pytest.approx(float(golden["swarm_ess_%d" % k]), rel=1e-11)
    x = np.concatenate([golden["norm_in"], golden["swarm_logw_2"]])
    off = np.array([0, 17, 17 + 7])
    log_q, q, ln = ops.normalize_segments(x, off)
    assert_close(log_q[:17], golden["norm_log"], rtol=1e-12)
    assert_close(q[:17], golden["norm_p"], rtol=1e-11)
    assert ln[0] == pytest.approx(float(golden["norm_lognorm"]), rel=1e-14)
    assert_close(q[17:], golden["swarm_W_2"], rtol=1e-11)


def test_resample_matches_inverse_cdf(ops, golden):
    rng = np.random.default_rng(11)
    W = golden["swarm_W_0"]
    u = rng.random(1000)
    np.testing.assert_array_equal(ops.resample(W, u), onum.ancestors_from_uniforms(W, u))
    # identical multiplicities for the same uniforms
    np.testing.assert_array_equal(
        np.bincount(ops.resample(W, u), minlength=len(W)), np.bincount(onum.ancestors_from_uniforms(W, u), minlength=len(W))
    )


# ---- full-size, size-independent properties (BASELINE.json configs #2, #4, #5 tile shapes) ----
@pytest.mark.parametrize("shape", [(4096, 8, 101), (512, 20, 201), (8192, 4, 101)])
def test_full_size_properties(ops, shape):
    B, S, G = shape
    rng = np.random.default_rng(2)
    a = np.log(rng.uniform(0.1, 1000, size=shape))
    b = np.log(rng.uniform(0.1, 1000, size=shape))
    y = ops.np_conv_dims(a, b)
    # commutativity of the pair convolution (same set of products)
    assert_close(ops.np_conv_dims(b, a), y, rtol=1e-12, what="commutative")
    # shift equivariance: adding a per-row constant to one operand shifts the result by it
    c = rng.uniform(-500, 500, size=(B, S, 1))
    assert_close(ops.np_conv_dims(a + c, b), y + c, rtol=1e-11, what="shift")
    # y[0] = a[0] + b[0] exactly one product
    assert_close(y[:, :, 0], a[:, :, 0] + b[:, :, 0], rtol=1e-12, what="first column")
    # prefix-LSE is non-decreasing and dominates its input
    s = ops.sub_compute_S(y)
    assert (np.diff(s, axis=-1) >= 0).all() and (s >= y - 1e-9).all()
    # spot-check against the oracle
    for i in rng.integers(0, B, size=4):
        assert_close(y[i], onum.conv_pair(a[i], b[i]), what="spot conv")
        assert_close(s[i], onum.prefix_lse(onum.conv_pair(a[i], b[i])), what="spot scan")


def test_empty_and_bad_arguments(ops):
    from phyclone_b200 import _lib

    assert ops.np_conv_dims(np.zeros((0, 4, 11)), np.zeros((0, 4, 11))).shape == (0, 4, 11)
    with pytest.raises(ValueError):
        ops.np_conv_dims(np.zeros((4, 11)), np.zeros((4, 12)))
    with pytest.raises(_lib.PcbError):
        ops.compute_log_S_batch([np.zeros((0, 4, 11)), np.zeros((1, 4, 11))])


def test_device_arena_jobs_match_oracle(golden):
    """The device-resident path (arena + job table) used by the sampler."""
    import torch

    from phyclone_b200 import engine

    tiles = golden["mixing_tiles"]  # [10,4,101]
    ar = engine.DeviceArena(4, 101, 64)
    ids = list(range(10))
    ar.import_tiles(tiles, ids)
    prior = -np.log(101)
    # tile 10 = tiles[2] + tiles[3] + prior ; tile 11 = node with children (0,1,4) and log_p = tile 10
    ar.tile_add([2], [3], [10], prior)
    jobs = [
        engine.make_job(0, 3, base=10, out=11, flags=engine.JOB_SCAN, slot=0),
        engine.make_job(3, 2, base=-1, out=-1, flags=engine.JOB_SCAN | engine.JOB_ADD_PRIOR, slot=1),
        engine.make_job(5, 1, base=-1, out=-1, flags=engine.JOB_SCAN | engine.JOB_ADD_PRIOR, slot=2),
    ]
    kids = [0, 1, 4, 5, 6, 7]
    sc = ar.run_jobs(np.array(jobs), np.array(kids), n_slots=3).cpu().numpy()
    log_p = tiles[2] + tiles[3] + prior
    node = onum.node_log_r(log_p, [tiles[0], tiles[1], tiles[4]])
    got = ar.export_tiles([10, 11]).cpu().numpy()
    np.testing.assert_array_equal(got[0], log_p)
    assert_close_bucketed(got[1], node, what="arena node", max_boundary=4)
    ref = [onum.root_terms(node), onum.root_terms(prior + onum.log_s([tiles[5], tiles[6]])),
           onum.root_terms(prior + onum.log_s([tiles[7]]))]
    assert_close(sc[0], [r[0] for r in ref], rtol=1e-9, what="lse terms")
    assert_close(sc[1], [r[1] for r in ref], rtol=1e-9, what="last terms")
    # second generation: the stored linear form of tile 11 feeds a further fold
    jobs2 = [engine.make_job(0, 2, base=-1, out=-1, flags=engine.JOB_SCAN | engine.JOB_ADD_PRIOR, slot=0)]
    sc2 = ar.run_jobs(np.array(jobs2), np.array([11, 8]), n_slots=1).cpu().numpy()
    ref2 = onum.root_terms(prior + onum.log_s([node, tiles[8]]))
    assert_close(sc2[:, 0], ref2, rtol=1e-9, what="second generation")
    torch.cuda.synchronize()


@pytest.mark.parametrize("variant", [-1, 1304, 708, 11304, 10704, "latency", "latency-2"])
@pytest.mark.parametrize("S,G,n_jobs", [(4, 101, 37), (8, 101, 21), (2, 64, 50), (16, 101, 5), (8, 224, 9), (1, 7, 3)])
def test_mixed_job_batches_match_oracle(variant, S, G, n_jobs):
    """One launch with jobs of every kind side by side -- 1..6 children, dummy-root scores, nodes with and without
    log_p / output tile / score slot -- which in the unit-per-warp kernel (variants >= 10000) share CTAs and warps."""
    from phyclone_b200 import _lib, engine

    rng = np.random.default_rng(100 * S + G + n_jobs)
    n_in = 24
    tiles = np.log(rng.uniform(0.1, 1000, size=(n_in, S, G)))
    tiles[3] = -((np.arange(G)[None] - rng.uniform(0, G, (S, 1))) ** 2) * 4.0  # floors
    tiles[7, :, : G // 2] -= 900.0
    # "latency": the warp-per-row mapping (node_jobs_rows_kernel) on the default / the chain's layout
    latency = isinstance(variant, str)
    if latency:
        variant = -2 if variant.endswith("-2") else -1
    try:
        ar = engine.DeviceArena(S, G, n_in + n_jobs + 4, variant=variant)
    except _lib.PcbError as e:
        assert "layout" in str(e)
        pytest.skip("variant %d does not cover (%d, %d)" % (variant, S, G))
    if latency:
        if ar.layout.kernel & 3:
            pytest.skip("the unit-per-warp layout has no latency mapping")
        ar.set_latency(True)
    ar.import_tiles(tiles, list(range(n_in)))
    prior = ar.prior
    jobs, kids, want = [], [], []
    for j in range(n_jobs):
        C = int(rng.integers(1, 7))
        ch = [int(c) for c in rng.integers(0, n_in, C)]
        kind = int(rng.integers(4))
        off = len(kids)
        kids.extend(ch)
        out = n_in + j
        kid_tiles = [tiles[c] for c in ch]
        if kind == 0:  # dummy-root score only
            jobs.append(engine.make_job(off, C, base=-1, out=-1, flags=engine.JOB_SCAN | engine.JOB_ADD_PRIOR, slot=j))
            want.append((None, prior + onum.log_s(kid_tiles)))
        elif kind == 1:  # node: log_p + log_S, stored and scored
            base = int(rng.integers(0, n_in))
            jobs.append(engine.make_job(off, C, base=base, out=out, flags=engine.JOB_SCAN, slot=j))
            want.append((out, onum.node_log_r(tiles[base], kid_tiles)))
        elif kind == 2:  # log_S only, stored, no score
            jobs.append(engine.make_job(off, C, base=-1, out=out, flags=engine.JOB_SCAN, slot=-1))
            want.append((out, onum.log_s(kid_tiles)))
        else:  # log_D (no scan), stored
            jobs.append(engine.make_job(off, C, base=-1, out=out, flags=0, slot=j))
            want.append((out, onum.log_d_chain(kid_tiles) if C > 1 else kid_tiles[0]))
    sc = ar.run_jobs(np.array(jobs), np.array(kids), n_slots=n_jobs).cpu().numpy()
    for j, (out, tile) in enumerate(want):
        if out is not None:
            got = ar.export_tiles([out]).cpu().numpy()[0]
            assert_close_bucketed(got, tile, what="job %d tile" % j, max_boundary=16)
        if jobs[j][5] >= 0:
            lse, last = onum.root_terms(tile)
            assert sc[0, j] == pytest.approx(lse, rel=1e-9), j
            if np.isfinite(last) and tile[:, -1].min() > tile.max() - 600:
                assert sc[1, j] == pytest.approx(last, rel=1e-9), j


@pytest.mark.parametrize("S", [1, 3, 8])
def test_latency_mapping_across_grid_sizes(S):
    """The warp-per-row mapping deals the strips of a row to lanes by a rule that depends on the number of strips
    (csrc/pcb_kernels.cuh node_jobs_rows_kernel): every strip count 1 .. 32 at L = 7, plus wider strips, against the
    oracle and against the lane-per-unit mapping of the same layout."""
    from phyclone_b200 import _lib, engine

    rng = np.random.default_rng(7 + S)
    cases = [(-2, G) for G in (1, 2, 6, 7, 8, 13, 14, 15, 49, 50, 97, 98, 99, 105, 106, 111, 112, 113, 150, 196, 217, 224)]
    cases += [(1308, 101), (1308, 201), (1316, 300), (1732, 400), (1732, 540)]
    ran = 0
    for variant, G in cases:
        try:
            ar = engine.DeviceArena(S, G, 16, variant=variant)
        except _lib.PcbError:
            continue
        if ar.layout.kernel & 3:
            continue
        n_in = 6
        tiles = np.log(rng.uniform(0.1, 1000, size=(n_in, S, G)))
        tiles[2] = -((np.arange(G)[None] - rng.uniform(0, G, (S, 1))) ** 2) * 3.0  # floors
        ar.import_tiles(tiles, list(range(n_in)))
        kids = [0, 1, 2, 3, 4, 5, 3, 1]
        jobs = [
            engine.make_job(0, 5, base=-1, out=-1, flags=engine.JOB_SCAN | engine.JOB_ADD_PRIOR, slot=0),
            engine.make_job(2, 4, base=0, out=8, flags=engine.JOB_SCAN, slot=1),
            engine.make_job(5, 1, base=1, out=9, flags=engine.JOB_SCAN, slot=2),
            engine.make_job(6, 2, base=-1, out=10, flags=0, slot=-1),
        ]
        want = [
            (None, ar.prior + onum.log_s([tiles[c] for c in kids[0:5]])),
            (8, onum.node_log_r(tiles[0], [tiles[c] for c in kids[2:6]])),
            (9, onum.node_log_r(tiles[1], [tiles[5]])),
            (10, onum.log_d_chain([tiles[3], tiles[1]])),
        ]
        got = {}
        for latency in (False, True):
            ar.set_latency(latency)
            sc = ar.run_jobs(np.array(jobs), np.array(kids), n_slots=3).cpu().numpy()
            out = ar.export_tiles([8, 9, 10]).cpu().numpy()
            got[latency] = (sc, out)
            for j, (o, tile) in enumerate(want):
                if o is not None:
                    assert_close_bucketed(out[o - 8], tile, what="G=%d job %d latency=%s" % (G, j, latency), max_boundary=16)
                if jobs[j][5] >= 0:
                    lse, last = onum.root_terms(tile)
                    assert sc[0, jobs[j][5]] == pytest.approx(lse, rel=1e-9), (G, j, latency)
        # the two mappings differ by rounding only
        np.testing.assert_allclose(got[True][0][0], got[False][0][0], rtol=1e-12)
        # the arena invariant both kernels rely on: nothing is ever written outside the G payload columns of a row
        import torch

        pad = torch.ones(ar.layout.pitch, dtype=torch.bool, device=ar.lg.device)
        pad[ar.layout.lead: ar.layout.lead + G] = False
        assert bool((ar.lin[:, :, pad] == 0).all()) and bool((ar.lg[:, :, pad] == 0).all()), (variant, G)
        ran += 1
    assert ran >= 20


@pytest.mark.parametrize("latency", [False, True])
def test_jobs_with_more_children_than_a_warp_holds(latency):
    """Child ids are handed out by shuffles from one coalesced load per 32 children: folds of 33, 64 and 70 children
    cross that boundary (forests with many roots do)."""
    from phyclone_b200 import engine

    S, G, n_in = 8, 101, 12
    rng = np.random.default_rng(3)
    tiles = np.log(rng.uniform(0.5, 2.0, size=(n_in, S, G)))
    ar = engine.DeviceArena(S, G, n_in + 8, variant=-2)
    ar.set_latency(latency)
    ar.import_tiles(tiles, list(range(n_in)))
    jobs, kids, want = [], [], []
    for j, C in enumerate((33, 64, 70, 32)):
        ch = [int(c) for c in rng.integers(0, n_in, C)]
        jobs.append(engine.make_job(len(kids), C, base=-1, out=-1, flags=engine.JOB_SCAN | engine.JOB_ADD_PRIOR, slot=j))
        kids.extend(ch)
        want.append(onum.root_terms(ar.prior + onum.log_s([tiles[c] for c in ch])))
    base = 5
    ch = [int(c) for c in rng.integers(0, n_in, 40)]
    jobs.append(engine.make_job(len(kids), 40, base=base, out=n_in, flags=engine.JOB_SCAN, slot=4))
    kids.extend(ch)
    node = onum.node_log_r(tiles[base], [tiles[c] for c in ch])
    want.append(onum.root_terms(node))
    sc = ar.run_jobs(np.array(jobs), np.array(kids), n_slots=5).cpu().numpy()
    for j, (lse, last) in enumerate(want):
        assert sc[0, j] == pytest.approx(lse, rel=1e-9), j
        assert sc[1, j] == pytest.approx(last, rel=1e-9), j
    assert_close_bucketed(ar.export_tiles([n_in]).cpu().numpy()[0], node, what="40-child node", max_boundary=16)
