"""GPU: the device-resident SMC sweep (csrc/pcb_sweep.cuh, smc/device_sweep.py) against its CPU definition, the oracle
sampler run with uniform-driven draws (oracle/uniform.py).

Both sides start from the same NumPy Generator (sigma, the sweep seeds, the Gibbs / prune-regraph moves and the
concentration update draw from it as in the reference); inside a sweep both use u(seed, step, slot, draw).  Required:
  * per sweep and SMC step: the same unnormalised particle log weights and log_p (1e-9), the same node every particle
    put its data point in (the decisions), the same resampling multiplicities (the ancestral indices), the same pick;
  * per iteration: identical topology, alpha, log_p_one (1e-9) -- whole chains, burn-in (unconditional sweep) included.
"""

import numpy as np
import pytest

from oracle import sampler as osm
from oracle.uniform import SweepRNG
from tests.chain_utils import load_chain, signature

pytestmark = pytest.mark.gpu


def _sweeps(log):
    """Split the oracle's flat record list into sweeps: [{"steps": [...], "resample": {t: counts}, "pick": idx}]."""
    out, cur = [], None
    for rec in log:
        if rec[0] == "step" and (cur is None or rec[1] == 0):
            if cur is not None and cur["pick"] is None:
                raise AssertionError("sweep without a pick")
            cur = {"steps": [], "resample": {}, "pick": None}
            out.append(cur)
        if rec[0] == "step":
            cur["steps"].append(rec)
        elif rec[0] == "resample":
            cur["resample"][rec[1]] = rec[2]
        else:
            cur["pick"] = rec[1]
    return out


def _close(a, b, what, tol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, "%s: max error %.3e at %d (%r vs %r)" % (what, err.max(), int(err.argmax()),
                                                                   a.flat[int(err.argmax())], b.flat[int(err.argmax())])
    return float(err.max())


def _compare_chain(vals, probs, particles, outlier_prob, burnin, iters, seed, what, perm_dist=False):
    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import Chain

    if probs is None:
        odata = [osm.Datum(i, v) for i, v in enumerate(vals)]
    else:
        odata = [osm.Datum(i, v, outlier_prob=probs[i][0], outlier_prob_not=probs[i][1]) for i, v in enumerate(vals)]
    log = []
    want = osm.run_chain(odata, SweepRNG(np.random.default_rng(seed)), iters, burnin=burnin, num_particles=particles,
                         outlier_prob=outlier_prob, sweep_log=log, perm_dist=perm_dist)
    sweeps = _sweeps(log)
    from phyclone_b200.smc.utils import RootPermutationDistribution

    chain = Chain(make_data_points(vals, outlier_probs=probs), np.random.default_rng(seed), num_particles=particles,
                  outlier_prob=outlier_prob, sweep="device",
                  perm_dist=RootPermutationDistribution() if perm_dist else None)
    hist_b, hist_m = [], []
    chain.burnin_sampler._device.history = hist_b
    chain.tree_sampler._device.history = hist_m
    trace = []
    for _ in range(burnin):
        chain.burnin_step()
    trace.append(chain.trace_entry(0, 0.0))
    for i in range(iters):
        chain.step()
        trace.append(chain.trace_entry(i, 0.0))
    st = chain.stats
    assert st["sweep"]["fallbacks"] == 0 and st["burnin_sweep"]["fallbacks"] == 0
    assert st["sweep"]["sweeps"] == iters and st["burnin_sweep"]["sweeps"] == burnin
    # ---- sweep level: weights, log_p, decisions, ancestors, pick
    hist = hist_b + hist_m
    assert len(hist) == len(sweeps)
    worst = 0.0
    for k, (dev, ora) in enumerate(zip(hist, sweeps)):
        conditional = k >= burnin
        first = 1 if conditional else 0
        T = len(ora["steps"])
        assert dev["raw"].shape[0] == T
        for t, (_, t_o, raw, logp, last_added) in enumerate(ora["steps"]):
            assert t_o == t
            where = "%s sweep %d step %d" % (what, k, t)
            worst = max(worst, _close(dev["raw"][t], raw, where + " log weights"))
            _close(dev["logp"][t][first:], logp[first:], where + " log_p")
            got = dev["name"][t][first:]
            ref = np.array([(-1 if n == osm.OUT else int(n)) for n in last_added[first:]])
            assert np.array_equal(got, ref), "%s: particles put the data point in different nodes" % where
            if t in ora["resample"]:
                assert dev["resampled"][t] == 1, where + ": oracle resampled, device did not"
                counts = np.bincount(dev["anc"][t][first:], minlength=particles)
                assert np.array_equal(counts, np.asarray(ora["resample"][t])), where + " resampling multiplicities"
            else:
                assert dev["resampled"][t] == 0, where + ": device resampled, oracle did not"
        assert int(dev["final"][0]) == ora["pick"], "%s sweep %d final pick" % (what, k)
    # ---- chain level
    assert len(trace) == len(want["trace"])
    for got, ref in zip(trace, want["trace"]):
        assert got["alpha"] == pytest.approx(ref["alpha"], rel=1e-9)
        assert got["log_p_one"] == pytest.approx(ref["log_p_one"], rel=1e-9)
        a = signature(got["tree"], len(vals))
        b = signature(ref["tree"], len(vals), name_of=lambda dp: dp.idx)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    print("%s: %d sweeps compared step by step, max log-weight error %.2e; device %s" % (
        what, len(hist), worst, st["sweep"]))
    return st


@pytest.mark.parametrize("mapping", ["latency", "default"])
@pytest.mark.parametrize("name,particles,iters", [("syn12", 16, 6), ("mixing", 20, 8), ("syn8_outliers", 12, 6)])
def test_device_sweep_matches_uniform_oracle(name, particles, iters, mapping, monkeypatch):
    """Both mappings of the job kernel: device-mode chains take the warp-per-row one for their small launches
    (run.py), PCB_LATENCY_KERNEL=0 keeps the lane-per-unit one."""
    monkeypatch.setenv("PCB_LATENCY_KERNEL", "1" if mapping == "latency" else "0")
    g = load_chain(name)
    vals = g["values"]
    probs = None
    if float(g["outlier_prob_run"]) > 0:
        probs = [(float(g["outlier_prob"][i]), float(g["outlier_prob_not"][i])) for i in range(len(vals))]
    _compare_chain(vals, probs, particles, float(g["outlier_prob_run"]), 1, iters, 11, name)


@pytest.mark.parametrize("name,particles,iters", [("syn12", 16, 6), ("syn8_outliers", 12, 6)])
def test_device_sweep_with_root_permutation_density(name, particles, iters):
    """The same comparison with `perm_dist` on: particle weights carry the root-permutation density
    (smc/kernels/base.py:49-55, smc/utils.py:18-66), as in the reference's posterior tests."""
    g = load_chain(name)
    vals = g["values"]
    probs = None
    if float(g["outlier_prob_run"]) > 0:
        probs = [(float(g["outlier_prob"][i]), float(g["outlier_prob_not"][i])) for i in range(len(vals))]
    _compare_chain(vals, probs, particles, float(g["outlier_prob_run"]), 1, iters, 13, name + " + perm_dist", perm_dist=True)


def test_device_mode_samples_the_exact_posterior():
    """The reference's own statistical check of its particle-Gibbs sampler (tests/test_pg_posterior.py:15-158: kernel
    with the root-permutation density, 10 particles, delta 0.03) run on the device-resident sweep: clade-set
    frequencies of 2000 PG sweeps against the exact FS-CRP posterior obtained by enumeration."""
    from collections import Counter

    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import Chain
    from phyclone_b200.smc.utils import RootPermutationDistribution
    from tests.posterior_utils import CASES, exact_posterior

    for name in ("two_points_non_informative", "two_points_two_clusters", "three_points_non_informative",
                 "two_points_2d_two_clusters"):
        data = CASES[name](np.random.default_rng(12345))
        truth = exact_posterior(data)
        chain = Chain(make_data_points(np.array([d.value for d in data])), np.random.default_rng(7), num_particles=10,
                      sweep="device", perm_dist=RootPermutationDistribution())
        tree = chain.tree
        counts = Counter()
        for it in range(-100, 2000):
            chain.kernel.clear_proposal_dist_caches()
            chain.store.reset()
            tree = chain.tree_sampler.sample_tree(tree)
            if it > 0:
                counts[_clades(tree)] += 1
        total = sum(counts.values())
        err = max(abs(counts[k] / total - p) for k, p in truth.items())
        print("%s: %d forests, max |device frequency - exact posterior| = %.4f" % (name, len(truth), err))
        assert err <= 0.03, (name, err)


def test_sweep_falls_back_to_the_host_sampler(monkeypatch):
    """Forests wider than the device sweep's root limit (64; forced down to 2 here) make a sweep fall back to the host
    sampler -- either before the call (the retained path does not fit) or after it (a particle overflowed: the
    sweep's outputs are void) -- and the chain carries on."""
    import phyclone_b200.smc.device_sweep as ds
    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import Chain

    monkeypatch.setattr(ds, "MAX_ROOTS", 2)
    g = load_chain("syn12")
    chain = Chain(make_data_points(g["values"]), np.random.default_rng(4), num_particles=12, sweep="device")
    chain.burnin_step()
    for _ in range(4):
        chain.step()
    st = chain.stats
    assert st["sweep"]["fallbacks"] + st["burnin_sweep"]["fallbacks"] >= 1
    assert np.isfinite(chain.log_p_one())
    assert sorted(dp.idx for dp in chain.tree.data) == list(range(len(g["values"])))


def test_device_sweep_at_bench_config():
    """BASELINE configs[1] at full size (50 x 8 x 101, 100 particles): burn-in sweep + two PG iterations."""
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(50, 8, 101, 10, 200, seed=0, density="beta-binomial", precision=400)
    st = _compare_chain(syn["values"], None, 100, 0.0, 1, 2, 1, "config #2")
    assert st["extensions"] >= 2 * 99 * 50 + 100 * 50


def test_device_sweep_wide_layout_many_particles():
    """S = 4 selects the unit-per-warp kernel; 200 particles over 40 unclustered points."""
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(40, 4, 101, 1, 200, seed=0, density="beta-binomial", precision=400)
    _compare_chain(syn["values"], None, 200, 0.0, 1, 2, 3, "40 points x 4 samples, 200 particles")


def test_device_sweep_500_particles():
    """BASELINE configs[4]'s swarm size (500 particles) on a reduced data set: 60 unclustered points x 4 samples."""
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(60, 4, 101, 1, 200, seed=0, density="beta-binomial", precision=400)
    _compare_chain(syn["values"], None, 500, 0.0, 1, 1, 7, "60 points x 4 samples, 500 particles")


def test_device_sweep_at_config4_shape():
    """BASELINE configs[3]'s tile shape and outlier modelling on a reduced data set: 20 clusters x 20 samples, grid 201,
    outlier_prob 1e-3, 40 particles."""
    from oracle.synthetic import make_synthetic
    from phyclone_b200 import workloads

    cfg = workloads.config_dict(4, clusters=20)
    syn = make_synthetic(20, cfg["samples"], cfg["grid"], cfg["muts_per_cluster"], cfg["depth"], seed=0,
                         density=cfg["density"], precision=cfg["precision"])
    _compare_chain(syn["values"], workloads.outlier_log_probs(cfg), 40, cfg["outlier_prob"], 1, 1, 9,
                   "20 clusters x 20 samples x grid 201, outliers")


def test_device_mode_has_the_host_samplers_stationary_distribution():
    """Statistical pin of the uniform-driven mode on the device, after the reference's own check of its particle-Gibbs
    sampler (tests/test_pg_posterior.py:15-158, delta 0.03).  That test gives the kernel the root-permutation
    density; production (run.py:403-417) and therefore the device sweep run without it, which is a different (biased)
    stationary distribution -- tests/test_uniform_mode_cpu.py checks the mode WITH the density against the exact
    posterior on the CPU.  Here: clade-set frequencies of 4000 device sweeps against those of the CPU port drawing
    from a NumPy Generator, same production configuration, on the reference's small cases."""
    from collections import Counter

    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import Chain
    from tests.posterior_utils import CASES

    for name in ("two_points_two_clusters", "three_points_non_informative", "two_points_2d_two_clusters"):
        data = CASES[name](np.random.default_rng(12345))
        gen = np.random.default_rng(5)
        tm, joint = osm.TileMath(), osm.Joint(1.0)
        kernel = osm.SemiAdaptedKernel(joint, gen, tm)
        tree = osm.Forest.single_node(data, tm)
        ref = Counter()
        for it in range(-100, 4000):
            kernel.clear()
            tree = osm.pg_move(tree, kernel, 10, 0.5)
            if it > 0:
                ref[tree.clades()] += 1
        chain = Chain(make_data_points(np.array([d.value for d in data])), np.random.default_rng(7), num_particles=10,
                      sweep="device")
        tree = chain.tree
        got = Counter()
        for it in range(-100, 4000):
            chain.kernel.clear_proposal_dist_caches()
            chain.store.reset()
            tree = chain.tree_sampler.sample_tree(tree)
            if it > 0:
                got[_clades(tree)] += 1
        n_ref, n_got = sum(ref.values()), sum(got.values())
        err = max(abs(got[k] / n_got - ref[k] / n_ref) for k in set(ref) | set(got))
        print("%s: %d forests seen, max |device frequency - host-sampler frequency| = %.4f" % (name, len(set(ref) | set(got)), err))
        assert err <= 0.03, (name, err)


def _clades(tree):
    """Clade set of a product Tree: one frozenset of data indices per node (its own data + all descendants')."""
    d = tree.to_dict()
    rev = d["node_idx_rev"]
    kids = {}
    for s, t in d["graph"]:
        kids.setdefault(rev[s], []).append(rev[t])
    out = set()

    def rec(node):
        acc = {dp.idx for dp in d["node_data"].get(node, ())}
        for c in kids.get(node, ()):
            acc |= rec(c)
        if node != "root":
            out.add(frozenset(acc))
        return acc

    rec("root")
    return frozenset(out)
