"""GPU: the product sampler with its real device tile store against the CPU oracle.

* replay: the oracle's RNG decisions (tests/tape.py) are replayed into the GPU sampler; every probability
  vector it asks to draw from must match the oracle's to 1e-9, every call must line up, and the resulting
  traces (topologies, alpha, log_p_one) must agree -- node log-likelihoods from the sm_100a kernels to 1e-9.
* free-running: same seed, real Generator.  Identical topologies are required for the committed cases; see
  tests/tape.py for why this cannot be promised for arbitrary seeds (knife-edge ties in multinomial).
* L1 tree level: every node tile of the final tree against an oracle rebuild of the same tree.
"""

import numpy as np
import pytest

from oracle import numerics as onum
from oracle import sampler as osm
from tests.chain_utils import compare_trace, load_chain
from tests.parity import assert_close_bucketed

pytestmark = pytest.mark.gpu

CASES = ["mixing", "syn12", "syn8_outliers", "syn6_fully", "syn6_boot", "syn10_subtree", "syn8_subtree_outliers"]


def _data(g):
    from phyclone_b200.data.base import make_data_points

    vals = g["values"]
    probs = [(float(g["outlier_prob"][i]) if g["outlier_prob"][i] != 0 else 0, float(g["outlier_prob_not"][i]))
             for i in range(len(vals))]
    return make_data_points(vals, outlier_probs=probs)


def _run(g, rng):
    from phyclone_b200.run import run_phyclone_chain

    seed, iters, particles, burnin = [int(x) for x in g["meta"]]
    return run_phyclone_chain(burnin, True, 1.0, _data(g), float("inf"), iters, particles, 1, 1,
                              float(g["outlier_prob_run"]), 10**9, str(g["proposal"]), 0.5, rng, ["s"], 1, 0,
                              float(g["subtree_update_prob"]) if "subtree_update_prob" in g else 0.0)


@pytest.mark.parametrize("name", CASES)
def test_replayed_chain_matches_oracle(name):
    from tests.tape import ReplayRNG
    from tests.test_tape_cpu import record_oracle

    g = load_chain(name)
    res_o, tape = record_oracle(g)
    rep = ReplayRNG(tape)
    res = _run(g, rep)
    assert rep.done()
    compare_trace(res["trace"], g, 1e-9, name + " (gpu, replayed)")
    print("%s: %d recorded draws replayed, max relative probability error %.2e, stats %s" % (
        name, len(tape), rep.max_p_err, res["stats"]))
    assert res["stats"]["node_jobs"] > 0 and res["stats"]["launches"] > 0


@pytest.mark.parametrize("name", CASES)
def test_free_running_chain_matches_reference_trace(name):
    """Same seed, real Generator.  The replay test above is the parity statement; this one additionally pins the
    committed cases end to end.  `mixing` (real data, 46 % floored outputs, many tied particles) sits on
    multinomial knife edges (tests/tape.py): a last-bit change of a device log-likelihood -- e.g. another strip
    width -- may legitimately send the free-running chain down another, equally valid path, so a divergence there
    is reported as an expected failure instead of an error."""
    g = load_chain(name)
    seed = int(g["meta"][0])
    res = _run(g, np.random.default_rng(seed))
    try:
        compare_trace(res["trace"], g, 1e-9, name + " (gpu, free-running)")
    except AssertionError as e:
        if name != "mixing":
            raise
        assert all(np.isfinite(t["log_p_one"]) for t in res["trace"])
        pytest.xfail("free-running chain left the recorded trajectory on a multinomial knife edge: %s; %s"
                     % (e, _first_divergence(g)))


def _first_divergence(g):
    """Where the free-running GPU chain leaves the oracle's: both run from the same seed with every draw recorded;
    report the first call whose inputs or outcome differ -- the two probability vectors, their largest difference in
    ulps, and numpy's knife-edge quantity p_k / remaining_p next to 0.5."""
    from tests.tape import RecordingRNG
    from tests.test_tape_cpu import record_oracle

    _, tape_o = record_oracle(g)
    rec = RecordingRNG(np.random.default_rng(int(g["meta"][0])))
    _run(g, rec)
    for k, (a, b) in enumerate(zip(tape_o, rec.tape)):
        if a[0] != b[0]:
            return "call %d: oracle drew %s, GPU sampler drew %s" % (k, a[0], b[0])
        if a[0] == "multinomial":
            (n0, p0), (n1, p1) = a[1], b[1]
            same_in = n0 == n1 and p0.shape == p1.shape
            if not same_in or not np.array_equal(a[2], b[2]):
                if not same_in:
                    return "call %d: multinomial shapes differ (%s vs %s)" % (k, p0.shape, p1.shape)
                ulp = np.abs(p0 - p1) / np.maximum(np.spacing(np.abs(p0)), 5e-324)
                j = int(np.argmax(ulp))
                rem = 1.0 - np.cumsum(p0)[:-1]
                ratio = p0[1:] / np.where(rem > 0, rem, np.nan)
                near = float(np.nanmin(np.abs(ratio - 0.5))) if ratio.size else float("nan")
                return ("first diverging draw = call %d, multinomial(n=%d) over %d weights: outcomes %s (oracle) vs %s "
                        "(GPU); probability vectors differ by at most %.1f ulp (entry %d: %.17g vs %.17g); closest "
                        "p_k/remaining_p to the 0.5 branch point of numpy's binomial: |ratio - 0.5| = %.3e"
                        % (k, n0, p0.size, np.flatnonzero(a[2]).tolist()[:8], np.flatnonzero(b[2]).tolist()[:8],
                           float(ulp.max()), j, p0[j], p1[j], near))
        elif a[0] in ("random", "integers") and a[2] != b[2]:
            return "call %d: %s outcomes differ although the generators were seeded alike" % (k, a[0])
    return "no diverging draw found in %d calls (traces differ only in floating point)" % min(len(tape_o), len(rec.tape))


def test_outlier_marginals_on_device():
    g = load_chain("syn12")
    got = np.array([dp.outlier_marginal_prob for dp in _data(g)])
    ref = np.array([onum.outlier_marginal(v) for v in g["values"]])
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_final_tree_tiles_match_oracle_rebuild():
    """L1: rebuild the last traced tree on both sides and compare every node's log_r."""
    from phyclone_b200.tiles import TileStore
    from phyclone_b200.tree.distributions import FSCRPDistribution, TreeJointDistribution
    from phyclone_b200.tree.tree import Tree

    g = load_chain("mixing")
    res = _run(g, np.random.default_rng(int(g["meta"][0])))
    frozen = res["trace"][-1]["tree"]
    store = TileStore(g["values"])
    tree = Tree.from_dict(frozen, store)
    got = tree.fetch_log_r()
    # oracle rebuild from the same frozen structure
    tm = osm.TileMath()
    odata = {i: osm.Datum(i, v) for i, v in enumerate(g["values"])}
    od = dict(frozen)
    od["node_data"] = {k: [odata[dp.idx] for dp in v] for k, v in frozen["node_data"].items()}
    ot = osm.Forest.thaw(od, tm)
    boundary = 0
    for node, tile in got.items():
        boundary += assert_close_bucketed(tile, ot.lr[ot.gi[node]], what="node %r" % (node,), max_boundary=8)
    joint = osm.Joint(res["trace"][-1]["alpha"])
    lp, lp1 = TreeJointDistribution(FSCRPDistribution(res["trace"][-1]["alpha"])).compute_both_log_p_and_log_p_one(tree)
    olp, olp1 = joint.both(ot)
    assert abs(lp - olp) <= 1e-9 * abs(olp) and abs(lp1 - olp1) <= 1e-9 * abs(olp1)
    print("nodes compared: %d, boundary-band entries: %d" % (len(got), boundary))


def test_arena_grows_on_demand():
    """A deliberately tiny arena must double itself (tile ids stay valid) instead of failing."""
    from phyclone_b200.run import run_phyclone_chain
    from phyclone_b200.tiles import TileStore

    g = load_chain("syn12")
    store = TileStore(g["values"], capacity=64)
    cap0 = store.arena.capacity
    res = run_phyclone_chain(1, True, 1.0, _data(g), float("inf"), 3, 16, 1, 1, 0.0, 10**9, "semi-adapted", 0.5,
                             np.random.default_rng(1), ["s"], 1, 0, 0.0, store=store)
    assert store.arena.capacity > cap0
    ref = run_phyclone_chain(1, True, 1.0, _data(g), float("inf"), 3, 16, 1, 1, 0.0, 10**9, "semi-adapted", 0.5,
                             np.random.default_rng(1), ["s"], 1, 0, 0.0)
    for a, b in zip(res["trace"], ref["trace"]):
        assert a["log_p_one"] == b["log_p_one"] and a["alpha"] == b["alpha"]


def _replay_against_oracle(vals, probs, particles, outlier_prob, proposal, burnin, iters, what, store=None, dp_factory=None):
    """One oracle chain with its draws recorded, replayed into the GPU sampler: the GPU sampler must ask for the same
    draws from the same distributions (1e-9) and produce the same topologies, alpha and log_p_one."""
    from phyclone_b200.run import run_phyclone_chain
    from tests.chain_utils import signature
    from tests.tape import RecordingRNG, ReplayRNG

    if dp_factory is None:
        from phyclone_b200.data.base import make_data_points
    rec = RecordingRNG(np.random.default_rng(1))
    if probs is None:
        odata = [osm.Datum(i, v) for i, v in enumerate(vals)]
    else:
        odata = [osm.Datum(i, v, outlier_prob=probs[i][0], outlier_prob_not=probs[i][1]) for i, v in enumerate(vals)]
    want = osm.run_chain(odata, rec, iters, burnin=burnin, num_particles=particles, proposal=proposal,
                         outlier_prob=outlier_prob)
    rep = ReplayRNG(rec.tape)
    data = make_data_points(vals, outlier_probs=probs) if dp_factory is None else dp_factory(vals, probs)
    res = run_phyclone_chain(burnin, True, 1.0, data, float("inf"), iters, particles,
                             1, 1, outlier_prob, 10**9, proposal, 0.5, rep, ["s"], 1, 0, 0.0, store=store)
    assert rep.done()
    assert len(res["trace"]) == len(want["trace"])
    for got, ref in zip(res["trace"], want["trace"]):
        assert got["alpha"] == pytest.approx(ref["alpha"], rel=1e-9)
        assert got["log_p_one"] == pytest.approx(ref["log_p_one"], rel=1e-9)
        a = signature(got["tree"], len(vals))
        b = signature(ref["tree"], len(vals), name_of=lambda dp: dp.idx)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    print("%s: %d draws replayed, max relative probability error %.2e, stats %s" % (
        what, len(rec.tape), rep.max_p_err, res["stats"]))
    return res


@pytest.mark.parametrize("depth", [200, 2000])
def test_replayed_chain_at_bench_config(depth):
    """BASELINE configs[1] at full size (50 clusters x 8 samples, grid 101, 100 particles; depth 2000 is the
    floor-exercising variant of SURVEY.md 8d): one burn-in sweep and one PG iteration of the CPU port, its
    draws replayed into the GPU sampler."""
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(50, 8, 101, 10, depth, seed=0, density="beta-binomial", precision=400)
    _replay_against_oracle(syn["values"], None, 100, 0, "semi-adapted", 1, 1, "config #2, depth %d" % depth)


def test_replayed_chain_at_config4_shape():
    """BASELINE configs[3] at full size: 100 clusters x 20 samples, grid 201, outlier modelling on (outlier_prob 1e-3:
    outlier candidates in every proposal, outlier marginals, the outlier Gibbs candidate), 100 particles; one burn-in
    sweep and one PG iteration (~20 k extensions), the (20,201) tile layout driven through the whole sampler."""
    from oracle.synthetic import make_synthetic
    from phyclone_b200 import workloads

    cfg = workloads.config_dict(4)
    syn = make_synthetic(cfg["clusters"], cfg["samples"], cfg["grid"], cfg["muts_per_cluster"], cfg["depth"], seed=0,
                         density=cfg["density"], precision=cfg["precision"])
    res = _replay_against_oracle(syn["values"], workloads.outlier_log_probs(cfg), cfg["particles"], cfg["outlier_prob"],
                                 "semi-adapted", 1, 1, "config #4 shape")
    assert res["stats"]["extensions"] >= 19000


def test_replayed_chain_at_reduced_config5():
    """Reduced BASELINE configs[4]: 300 unclustered mutations x 4 samples, grid 101, **500 particles** (the
    unit-per-warp kernel that S = 4 selects, the Gibbs sampler over 300 data points, ~300 k extensions)."""
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(300, 4, 101, 1, 200, seed=0, density="beta-binomial", precision=400)
    res = _replay_against_oracle(syn["values"], None, 500, 0, "semi-adapted", 1, 1, "reduced config #5")
    assert res["stats"]["extensions"] >= 290000


@pytest.mark.parametrize("proposal,points", [("fully-adapted", 20), ("bootstrap", 24)])
def test_replayed_other_proposals(proposal, points):
    """Fully-adapted (R + 2^R candidates per parent) and bootstrap proposals at >= 20 data points x 50 particles."""
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(points, 4, 41, 4, 400, seed=0)
    _replay_against_oracle(syn["values"], None, 50, 0, proposal, 1, 1, "%s, %d points x 50 particles" % (proposal, points))


def test_in_extension_flush_equals_generic_flush(monkeypatch):
    """`TileBook.run` (pack + stage + pcb_run_plan_dev inside the host extension) and the step-by-step Python flush
    must be the same computation: identical traces and identical launch / job counts on the same seed."""
    from phyclone_b200.tiles import TileStore

    g = load_chain("syn12")
    seed = int(g["meta"][0])
    fast = _run(g, np.random.default_rng(seed))
    assert fast["stats"]["staged_ints"] > 0  # the in-extension path did run

    def generic_only(self):
        if self.book.has_pending():
            self._flush_generic()

    monkeypatch.setattr(TileStore, "flush", generic_only)
    slow = _run(g, np.random.default_rng(seed))
    assert slow["stats"]["staged_ints"] == 0
    assert len(fast["trace"]) == len(slow["trace"])
    for a, b in zip(fast["trace"], slow["trace"]):
        assert a["log_p_one"] == b["log_p_one"] and a["alpha"] == b["alpha"]
        assert a["tree"]["graph"] == b["tree"]["graph"]
    for k in ("flushes", "launches", "node_jobs", "adds", "scores", "conv_pairs", "d2h_bytes"):
        assert fast["stats"][k] == slow["stats"][k], k


def test_chains_in_worker_processes_match_in_process_chains():
    """`run_chains_local` (one process per chain, sharing the GPU through MPS when the daemon can be started) must
    give every chain the trace it has when the chains run one after the other in this process."""
    from phyclone_b200.run import run_chains, run_chains_local

    g = load_chain("syn12")
    data = _data(g)
    kw = dict(num_chains=3, seed=11, burnin=1, num_iters=3, num_particles=10, print_freq=10**9)
    here = run_chains(data, ["s"], **kw)
    there = run_chains_local(data, ["s"], devices=[0], mps=True, **kw)
    assert sorted(there) == sorted(here) == [0, 1, 2]
    for c in here:
        assert len(there[c]["trace"]) == len(here[c]["trace"])
        for a, b in zip(there[c]["trace"], here[c]["trace"]):
            assert a["log_p_one"] == b["log_p_one"] and a["alpha"] == b["alpha"]
            assert a["tree"]["graph"] == b["tree"]["graph"]
