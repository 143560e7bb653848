"""Chain driver (reference run.py:177-380): burn-in, particle-Gibbs main loop, trace building.

`run_phyclone_chain` keeps the reference's positional signature; one chain drives one CUDA stream
and one TileStore (the device arena holding every tile of the chain).
"""

import time

import numpy as np

from .mcmc.concentration import GammaPriorConcentrationSampler
from .mcmc.gibbs_mh import DataPointSampler, PruneRegraphSampler
from .mcmc.particle_gibbs import ParticleGibbsTreeSampler
from .smc.kernels import BootstrapKernel, FullyAdaptedKernel, SemiAdaptedKernel
from .smc.samplers import UnconditionalSMCSampler
from .tiles import TileStore
from .tree.distributions import FSCRPDistribution, TreeJointDistribution
from .tree.tree import Tree

KERNELS = {"bootstrap": BootstrapKernel, "fully-adapted": FullyAdaptedKernel, "semi-adapted": SemiAdaptedKernel}


class Timer(object):
    """utils/utils.py:28-67"""

    def __init__(self):
        self.elapsed = 0.0
        self._start = None

    def __enter__(self):
        self._start = time.time()
        return self

    def __exit__(self, *args):
        self.elapsed += time.time() - self._start
        self._start = None


def setup_kernel(outlier_prob, proposal, rng, tree_dist, store):
    """run.py:403-417"""
    outlier_proposal_prob = 0.1 if outlier_prob > 0 else 0
    return KERNELS[proposal](tree_dist, rng, store, outlier_proposal_prob=outlier_proposal_prob)


def run_phyclone_chain(burnin, concentration_update, concentration_value, data, max_time, num_iters, num_particles,
                       num_samples_data_point, num_samples_prune_regraph, outlier_prob, print_freq, proposal,
                       resample_threshold, rng, samples, thin, chain_num, subtree_update_prob, device=None,
                       store=None, on_iteration=None):
    """run.py:177-232 (+ :235-380).  Returns {"data", "samples", "trace", "chain_num", "stats"}."""
    if subtree_update_prob != 0.0:
        raise NotImplementedError("the subtree particle-Gibbs move is outside the B200 hot path (DESIGN.md section 7)")
    if store is None:
        store = TileStore(np.array([dp.value for dp in data]), device=device)
    tree_dist = TreeJointDistribution(FSCRPDistribution(concentration_value))
    kernel = setup_kernel(outlier_prob, proposal, rng, tree_dist, store)
    dp_sampler = DataPointSampler(tree_dist, rng, outliers=(outlier_prob > 0))
    prg_sampler = PruneRegraphSampler(tree_dist, rng)
    conc_sampler = GammaPriorConcentrationSampler(0.01, 0.01, rng=rng)
    burnin_sampler = UnconditionalSMCSampler(kernel, num_particles=num_particles, resample_threshold=resample_threshold)
    tree_sampler = ParticleGibbsTreeSampler(kernel, rng, num_particles=num_particles,
                                            resample_threshold=resample_threshold)
    timer = Timer()
    tree = Tree.get_single_node_tree(data, store)

    def local_moves(tree):
        for _ in range(num_samples_data_point):
            tree = dp_sampler.sample_tree(tree)
        for _ in range(num_samples_prune_regraph):
            tree = prg_sampler.sample_tree(tree)
        tree.relabel_nodes()
        return tree

    for i in range(burnin):
        with timer:
            kernel.clear_proposal_dist_caches()
            store.reset()
            tree = local_moves(burnin_sampler.sample_tree(tree))
            if timer.elapsed > max_time:
                break

    def entry(i, tree):
        return {"iter": i, "time": timer.elapsed, "alpha": tree_dist.prior.alpha,
                "log_p_one": tree_dist.log_p_one(tree), "tree": tree.to_dict()}

    trace = [entry(0, tree)]
    for i in range(num_iters):
        with timer:
            kernel.clear_proposal_dist_caches()
            store.reset()
            rng.random()  # run.py:268: the subtree-move coin is drawn even when its probability is 0
            tree = local_moves(tree_sampler.sample_tree(tree))
            if concentration_update:
                node_sizes = [len(v) for k, v in tree.node_data.items() if k != tree.outlier_node_name]
                tree_dist.prior.alpha = conc_sampler.sample(tree_dist.prior.alpha, len(node_sizes), sum(node_sizes))
            if i % thin == 0:
                trace.append(entry(i, tree))
            if on_iteration is not None:
                on_iteration(i, trace[-1])
            if timer.elapsed >= max_time:
                break
    return {"data": data, "samples": samples, "trace": trace, "chain_num": chain_num,
            "stats": dict(store.stats, extensions=kernel.num_extensions)}
