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
