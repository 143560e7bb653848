"""CPU: the oracle sampler (oracle/sampler.py) must reproduce, decision for decision, the traces
recorded from the unmodified reference (tile caches bypassed) by tests/golden/make_golden_chain.py."""

import numpy as np
import pytest

from oracle import sampler as osm
from tests.chain_utils import compare_trace, load_chain

CASES = ["mixing", "syn12", "syn8_outliers", "syn6_fully", "syn6_boot", "syn10_subtree", "syn8_subtree_outliers"]


def run_oracle(g, **kw):
    vals = g["values"]
    data = [osm.Datum(i, vals[i], outlier_prob=float(g["outlier_prob"][i]) if g["outlier_prob"][i] != 0 else 0,
                      outlier_prob_not=float(g["outlier_prob_not"][i])) for i in range(len(vals))]
    seed, iters, particles, burnin = [int(x) for x in g["meta"]]
    rng = np.random.default_rng(seed)
    res = osm.run_chain(data, rng, iters, burnin=burnin, num_particles=particles, proposal=str(g["proposal"]),
                        outlier_prob=float(g["outlier_prob_run"]),
                        subtree_update_prob=float(g["subtree_update_prob"]) if "subtree_update_prob" in g else 0.0, **kw)
    return res, rng


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_trace(name):
    g = load_chain(name)
    res, rng = run_oracle(g)
    compare_trace(res["trace"], g, 1e-12, name)
    # the Generator must have been consumed identically
    assert rng.random() == float(g["rng_after"])
