"""CPU: the uniform-driven mode of the SMC sweep as the oracle defines it (oracle/uniform.py) -- the specification the
device-resident sweep is checked against on the GPU (tests/test_gpu_device_sweep.py).

* the keyed uniform and the inverse-CDF / integers / choice conventions (known answers);
* the mode samples the right distribution: the reference's own exact-posterior check of its particle-Gibbs sampler
  (tests/test_pg_posterior.py:15-158: root-permutation density on, 10 particles, delta 0.03) passes with keyed
  uniforms in place of the NumPy Generator;
* switching the mode on does not disturb anything outside the sweep (sigma, Gibbs moves, alpha draw from the Generator).
"""

from collections import Counter

import numpy as np
import pytest

from oracle import sampler as osm
from oracle.uniform import FINAL_SLOT, RESAMPLE_SLOT, SweepRNG, inv_cdf, mix64, u01
from tests.posterior_utils import CASES, exact_posterior


def test_keyed_uniform_known_answers():
    # splitmix64 reference stream for seed 0 (the first outputs of the published generator)
    assert mix64(0x9E3779B97F4A7C15) == 0xE220A8397B1DCDAF
    assert mix64(2 * 0x9E3779B97F4A7C15 & ((1 << 64) - 1)) == 0x6E789E6AA1B965F4
    # u01(0, 0, 0, c) = c-th output of splitmix64(seed 0) as a 53-bit fraction
    assert u01(0, 0, 0, 0) == (0xE220A8397B1DCDAF >> 11) / 2.0 ** 53
    us = [u01(123, t, i, c) for t in range(3) for i in (0, 7, RESAMPLE_SLOT, FINAL_SLOT) for c in range(4)]
    assert len(set(us)) == len(us) and all(0.0 <= u < 1.0 for u in us)


def test_draw_conventions():
    assert inv_cdf([0.2, 0.3, 0.5], 0.0) == 0 and inv_cdf([0.2, 0.3, 0.5], 0.2) == 1
    assert inv_cdf([0.2, 0.3, 0.5], 0.4999) == 1 and inv_cdf([0.2, 0.3, 0.5], 0.9999999) == 2
    assert inv_cdf([1.0], 0.7) == 0
    rng = SweepRNG(np.random.default_rng(0))
    rng.begin(42)
    rng.at(3, 5)
    u = [u01(42, 3, 5, c) for c in range(8)]
    assert rng.random() == u[0]
    assert rng.integers(0, 7) == min(6, int(u[1] * 7))
    picked = rng.choice([10, 11, 12, 13], 2, replace=False)
    pos = [0, 1, 2, 3]
    j = 0 + min(3, int(u[2] * 4))
    pos[0], pos[j] = pos[j], pos[0]
    j = 1 + min(2, int(u[3] * 3))
    pos[1], pos[j] = pos[j], pos[1]
    assert list(picked) == [10 + pos[0], 10 + pos[1]]
    counts = rng.multinomial(3, [0.25, 0.25, 0.5])
    want = np.bincount([inv_cdf([0.25, 0.25, 0.5], x) for x in u[4:7]], minlength=3)
    assert np.array_equal(counts, want)
    rng.end()
    g = np.random.default_rng(0)
    assert SweepRNG(np.random.default_rng(0)).random() == g.random()  # outside a sweep: the Generator itself


@pytest.mark.parametrize("case", ["two_points_two_clusters", "three_points_non_informative"])
def test_uniform_mode_samples_the_exact_posterior(case):
    data = CASES[case](np.random.default_rng(12345))
    truth = exact_posterior(data)
    rng = SweepRNG(np.random.default_rng(7))
    tm, joint = osm.TileMath(), osm.Joint(1.0, perm_dist=True)
    kernel = osm.SemiAdaptedKernel(joint, rng, tm)
    tree = osm.Forest.single_node(data, tm)
    counts = Counter()
    for it in range(-100, 2000):
        kernel.clear()
        tree = osm.pg_move(tree, kernel, 10, 0.5)
        if it > 0:
            counts[tree.clades()] += 1
    total = sum(counts.values())
    err = max(abs(counts[k] / total - p) for k, p in truth.items())
    print("%s: %d forests, max |frequency - posterior| = %.4f" % (case, len(truth), err))
    assert err <= 0.03


def test_mode_only_changes_the_sweep_draws():
    """Same Generator seed with and without the mode: both chains draw sigma / seeds / local moves from the Generator,
    so they consume it differently, but each is reproducible and the uniform chain does not depend on NumPy's
    multinomial inside the sweep (its trace is a function of the seed alone)."""
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(8, 3, 41, 4, 800, seed=0)
    data = [osm.Datum(i, v) for i, v in enumerate(syn["values"])]
    a = osm.run_chain(data, SweepRNG(np.random.default_rng(3)), 3, burnin=1, num_particles=12)
    b = osm.run_chain(data, SweepRNG(np.random.default_rng(3)), 3, burnin=1, num_particles=12)
    assert [e["log_p_one"] for e in a["trace"]] == [e["log_p_one"] for e in b["trace"]]
    assert [e["clades"] for e in a["trace"]] == [e["clades"] for e in b["trace"]]
