"""CPU: the decision tape itself -- record an oracle run, replay it into the product host sampler
(NumPy tile store): every call must line up and the traces must agree."""

import numpy as np
import pytest

from oracle import sampler as osm
from tests.chain_utils import compare_trace, load_chain
from tests.tape import RecordingRNG, ReplayRNG
from tests.test_host_sampler_cpu import FakeTileStore
from tests.test_oracle_chain import run_oracle


def record_oracle(g):
    vals = g["values"]
    data = [osm.Datum(i, vals[i], outlier_prob=float(g["outlier_prob"][i]) if g["outlier_prob"][i] != 0 else 0,
                      outlier_prob_not=float(g["outlier_prob_not"][i])) for i in range(len(vals))]
    seed, iters, particles, burnin = [int(x) for x in g["meta"]]
    rec = RecordingRNG(np.random.default_rng(seed))
    res = osm.run_chain(data, rec, iters, burnin=burnin, num_particles=particles, proposal=str(g["proposal"]),
                        outlier_prob=float(g["outlier_prob_run"]),
                        subtree_update_prob=float(g["subtree_update_prob"]) if "subtree_update_prob" in g else 0.0)
    return res, rec.tape


def replay_product(g, tape, store):
    from phyclone_b200.data.base import DataPoint
    from phyclone_b200.run import run_phyclone_chain
    from oracle import numerics as onum

    vals = g["values"]
    data = [DataPoint(i, vals[i], outlier_prob=float(g["outlier_prob"][i]) if g["outlier_prob"][i] != 0 else 0,
                      outlier_prob_not=float(g["outlier_prob_not"][i]),
                      outlier_marginal_prob=onum.outlier_marginal(vals[i])) for i in range(len(vals))]
    seed, iters, particles, burnin = [int(x) for x in g["meta"]]
    rep = ReplayRNG(tape)
    res = run_phyclone_chain(burnin, True, 1.0, data, float("inf"), iters, particles, 1, 1, float(g["outlier_prob_run"]),
                             10**9, str(g["proposal"]), 0.5, rep, ["s"], 1, 0,
                             float(g["subtree_update_prob"]) if "subtree_update_prob" in g else 0.0, store=store)
    assert rep.done(), "product consumed %d of %d recorded draws" % (rep.pos, len(tape))
    return res, rep


@pytest.mark.parametrize("name", ["mixing", "syn12", "syn8_outliers", "syn6_fully", "syn6_boot", "syn10_subtree",
                                  "syn8_subtree_outliers"])
def test_tape_replay_cpu(name):
    g = load_chain(name)
    res, tape = record_oracle(g)
    compare_trace(res["trace"], g, 1e-12, name)  # recording does not disturb the oracle
    res2, rep = replay_product(g, tape, FakeTileStore(g["values"]))
    compare_trace(res2["trace"], g, 1e-9, name + " (replayed)")


@pytest.mark.parametrize("proposal,points", [("fully-adapted", 20), ("bootstrap", 24)])
def test_tape_replay_other_proposals_cpu(proposal, points):
    """Host logic of the GPU test `test_replayed_other_proposals` (>= 20 data points x 50 particles) on the NumPy
    executor: fully-adapted enumerates R + 2^R candidates per parent, bootstrap draws blind."""
    from oracle import numerics as onum
    from oracle.synthetic import make_synthetic
    from phyclone_b200.data.base import DataPoint
    from tests.test_gpu_chain import _replay_against_oracle

    syn = make_synthetic(points, 4, 41, 4, 400, seed=0)

    def cpu_points(vals, probs):
        return [DataPoint(i, v, outlier_marginal_prob=onum.outlier_marginal(v)) for i, v in enumerate(vals)]

    _replay_against_oracle(syn["values"], None, 50, 0, proposal, 1, 1, proposal, store=FakeTileStore(syn["values"]),
                           dp_factory=cpu_points)
