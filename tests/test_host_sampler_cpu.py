"""CPU: the HOST side of the product sampler (phyclone_b200 forest bookkeeping, proposals, samplers,
moves, RNG call order) driven by a test-only NumPy tile store must reproduce the reference traces.
The device arithmetic is covered by the -m gpu suite; this pins everything around it."""

import numpy as np
import pytest

from phyclone_b200.data.base import DataPoint
from phyclone_b200.run import run_phyclone_chain
from tests.chain_utils import compare_trace, load_chain
from tests.fake_store import FakeTileStore
from oracle import numerics as onum

CASES = ["mixing", "syn12", "syn8_outliers", "syn6_fully", "syn6_boot"]


def run_product(g, store_cls=FakeTileStore):
    vals = g["values"]
    data = [
        DataPoint(i, vals[i], outlier_prob=float(g["outlier_prob"][i]) if g["outlier_prob"][i] != 0 else 0,
                  outlier_prob_not=float(g["outlier_prob_not"][i]), outlier_marginal_prob=onum.outlier_marginal(vals[i]))
        for i in range(len(vals))
    ]
    seed, iters, particles, burnin = [int(x) for x in g["meta"]]
    rng = np.random.default_rng(seed)
    store = store_cls(vals)
    res = run_phyclone_chain(burnin, True, 1.0, data, float("inf"), iters, particles, 1, 1, float(g["outlier_prob_run"]),
                             10**9, str(g["proposal"]), 0.5, rng, ["s"], 1, 0, 0.0, store=store)
    return res, rng


@pytest.mark.parametrize("name", CASES)
def test_host_sampler_reproduces_reference_trace(name):
    g = load_chain(name)
    res, rng = run_product(g)
    compare_trace(res["trace"], g, 1e-9, name)
    # with the NumPy tile store the arithmetic is the oracle's, so even the Generator state must agree
    assert rng.random() == float(g["rng_after"])


def test_host_sampler_matches_oracle_with_many_outliers():
    """Outlier-heavy run (outlier_prob 0.3): exercises outlier proposals, outlier <-> node Gibbs moves and the
    outlier bookkeeping of the root-level forest states against the pinned oracle sampler.  The oracle's
    decisions are replayed (tests/tape.py): the product's prior terms are running sums, not bit-identical to the
    oracle's, and with many tied particles one ulp is enough to flip a free-running multinomial draw."""
    from oracle import sampler as osm
    from tests.chain_utils import signature
    from tests.tape import RecordingRNG, ReplayRNG

    g = load_chain("syn12")
    vals = g["values"][:9]
    lo, lon = float(np.log(0.3) * 4), float(np.log1p(-0.3) * 4)
    odata = [osm.Datum(i, vals[i], outlier_prob=lo, outlier_prob_not=lon) for i in range(len(vals))]
    rec = RecordingRNG(np.random.default_rng(5))
    ref = osm.run_chain(odata, rec, 10, burnin=2, num_particles=12, outlier_prob=0.3)
    assert any(len(e["outliers"]) > 0 for e in ref["trace"]), "the case must actually produce outliers"
    assert len({tuple(sorted(e["outliers"])) for e in ref["trace"]}) > 1, "and move points in and out of them"
    data = [DataPoint(i, vals[i], outlier_prob=lo, outlier_prob_not=lon,
                      outlier_marginal_prob=onum.outlier_marginal(vals[i])) for i in range(len(vals))]
    rep = ReplayRNG(rec.tape)
    res = run_phyclone_chain(2, True, 1.0, data, float("inf"), 10, 12, 1, 1, 0.3, 10**9, "semi-adapted", 0.5, rep,
                             ["s"], 1, 0, 0.0, store=FakeTileStore(vals))
    assert rep.done()
    n = len(vals)
    for i, (a, b) in enumerate(zip(res["trace"], ref["trace"])):
        own_a, par_a = signature(a["tree"], n)
        own_b, par_b = signature(b["tree"], n)
        assert np.array_equal(own_a, own_b) and np.array_equal(par_a, par_b), "topology differs at entry %d" % i
        assert abs(a["log_p_one"] - b["log_p_one"]) <= 1e-9 * abs(b["log_p_one"])
        assert abs(a["alpha"] - b["alpha"]) <= 1e-9 * abs(b["alpha"])


def test_genealogy_longer_than_the_recursion_limit():
    """Config #5 has 2000 data points: the replay of a particle's edits and the tree walks must not depend on
    Python's default recursion limit (a no-arithmetic executor keeps this fast)."""
    import inspect
    import sys

    from phyclone_b200.data.base import DataPoint
    from phyclone_b200.run import Chain
    from tests.null_store import NullTileStore

    T = 320
    vals = np.zeros((T, 2, 8))
    data = [DataPoint(i, vals[i], outlier_marginal_prob=-10.0) for i in range(T)]
    saved = sys.getrecursionlimit()
    try:
        chain = Chain(data, np.random.default_rng(3), num_particles=4, store=NullTileStore(vals))
        sys.setrecursionlimit(len(inspect.stack()) + 250)  # fewer frames than edits in a genealogy
        chain.burnin_step()
        assert sum(len(v) for v in chain.tree.to_dict()["node_data"].values()) == T
    finally:
        sys.setrecursionlimit(saved)


def test_chain_snapshot_restore_repeats_the_same_iterations():
    """bench.py rewinds the chain between its two timed legs so that both run identical iterations."""
    from phyclone_b200.data.base import DataPoint
    from phyclone_b200.run import Chain
    from tests.fake_store import FakeTileStore

    rng = np.random.default_rng(0)
    vals = np.log(rng.uniform(0.1, 1000, (7, 2, 15)))
    vals[:3, :, 8:] -= 30.0
    data = [DataPoint(i, vals[i], outlier_marginal_prob=-5.0) for i in range(7)]
    chain = Chain(data, np.random.default_rng(3), num_particles=6, store=FakeTileStore(vals))
    chain.burnin_step()
    chain.step()
    snap = chain.snapshot()
    first = []
    for _ in range(3):
        chain.step()
        first.append((chain.log_p_one(), chain.tree_dist.prior.alpha, chain.tree.to_newick_string()))
    chain.restore(snap)
    again = []
    for _ in range(3):
        chain.step()
        again.append((chain.log_p_one(), chain.tree_dist.prior.alpha, chain.tree.to_newick_string()))
    assert first == again
    assert len({f[1] for f in first}) > 1  # the concentration did move: the replay is not trivially constant
