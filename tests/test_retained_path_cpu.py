"""CPU: the retained path of a conditional sweep as root-level states (smc/device_sweep.py::RetainedPath, built by the
host extension without any `Tree`) must be exactly what the tree-based builder (samplers.ConditionalSMCSampler, the
reference's conditional.py:19-74) produces: same packed records for the device sweep (states in both orders, scores,
structure hashes, edits) and the same replayed trees."""

import types

import numpy as np
import pytest

from oracle import numerics as onum
from tests.chain_utils import load_chain
from tests.test_host_sampler_cpu import FakeTileStore


def _chain(name, seed, particles, outliers=False):
    from phyclone_b200.data.base import DataPoint
    from phyclone_b200.run import Chain

    g = load_chain(name)
    vals = g["values"]
    data = []
    for i in range(len(vals)):
        op = float(g["outlier_prob"][i]) if g["outlier_prob"][i] != 0 else 0
        data.append(DataPoint(i, vals[i], outlier_prob=op, outlier_prob_not=float(g["outlier_prob_not"][i]),
                              outlier_marginal_prob=onum.outlier_marginal(vals[i])))
    ch = Chain(data, np.random.default_rng(seed), num_particles=particles, store=FakeTileStore(vals),
               outlier_prob=float(g["outlier_prob_run"]))
    ch.burnin_step()
    for _ in range(3):
        ch.step()
    return ch


@pytest.mark.parametrize("name", ["syn12", "mixing", "syn8_outliers"])
def test_state_path_equals_tree_path(name):
    from phyclone_b200.smc.device_sweep import DeviceSweep, RetainedPath
    from phyclone_b200.smc.samplers import ConditionalSMCSampler
    from phyclone_b200.smc.utils import RootPermutationDistribution

    ch = _chain(name, 5, 10)
    for rep in range(3):
        tree = ch.tree
        sigma = RootPermutationDistribution.sample(tree, np.random.default_rng(100 + rep))
        T = len(sigma)
        rmax = min(64, T + 1)
        ch.kernel.clear_proposal_dist_caches()
        ch.store.reset()
        sampler = ConditionalSMCSampler(tree, sigma, ch.kernel, num_particles=10, with_proposals=False)
        dummy = types.SimpleNamespace(perm=False)
        old = DeviceSweep._pack_retained(dummy, sampler.constrained_path, rmax)
        rp = RetainedPath(tree, sigma, ch.kernel)
        new = DeviceSweep._pack_states(dummy, rp, rmax)
        assert [tuple(e) for e in sampler.path_edits] == [tuple(e) for e in rp.edits]
        names = ["thawed ints", "as-built ints", "thawed doubles", "as-built doubles", "keys", "kids"]
        # child-list offsets depend on packing order only; compare the lists they point at
        for k in (0, 1):
            a = np.asarray(old[k]).reshape(T, -1).copy()
            b = np.asarray(new[k]).reshape(T, -1).copy()
            body = slice(5, 5 + 9 * rmax)
            fa, fb = a[:, body].reshape(T, 9, rmax), b[:, body].reshape(T, 9, rmax)
            for t in range(T):
                R = a[t, 0]
                assert R == b[t, 0]
                for r in range(R):
                    ka = old[5][fa[t, 7, r]: fa[t, 7, r] + fa[t, 8, r]]
                    kb = new[5][fb[t, 7, r]: fb[t, 7, r] + fb[t, 8, r]]
                    assert np.array_equal(ka, kb), (names[k], t, r)
            fa[:, 7, :] = 0
            fb[:, 7, :] = 0
            assert np.array_equal(a[:, :5], b[:, :5]), names[k]
            assert np.array_equal(fa, fb), names[k]
        for k in (2, 3):
            a, b = np.asarray(old[k]).reshape(T, 8), np.asarray(new[k]).reshape(T, 8)
            # crp, mult, oprior, omarg: running sums kept by the Tree (tree.prior_terms) on one side and by the host
            # extension's states on the other -- the same terms added in a different order
            assert np.allclose(a[:, :4], b[:, :4], rtol=1e-13, atol=1e-13), names[k]
        assert np.allclose(np.asarray(old[2]).reshape(T, 8)[:, 4:6], np.asarray(new[2]).reshape(T, 8)[:, 4:6], rtol=1e-13,
                           atol=0)
        assert np.array_equal(old[4], new[4]), "structure hashes"
        for t in (1, T // 2, T):
            want = sampler.constrained_path[t].tree.to_dict()
            got = rp.tree(t).to_dict()
            assert want["graph"] == got["graph"] and want["node_data"] == got["node_data"]
        ch.step()
