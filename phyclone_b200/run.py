"""Chain driver (reference run.py:177-380): burn-in, particle-Gibbs main loop, trace building.

`run_phyclone_chain` keeps the reference's positional signature; one chain drives one CUDA stream
and one TileStore (the device arena holding every tile of the chain).
"""

import gc
import sys
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


class Chain(object):
    """One MCMC chain = one Python thread driving one CUDA stream and one device arena (run.py:177-380).
    `burnin_step()` / `step()` are one trip of the reference's burn-in / main loops."""

    def __init__(self, data, rng, num_particles=100, proposal="semi-adapted", resample_threshold=0.5, outlier_prob=0.0,
                 concentration_value=1.0, concentration_update=True, num_samples_data_point=1,
                 num_samples_prune_regraph=1, device=None, store=None):
        self.data = data
        self.rng = rng
        # the tree walks recurse once per level and a forest over T data points can be T deep
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * len(data) + 1000))
        if store is None:
            # a chain launches a handful of jobs at a time: the layout with the shortest launches (-2.7 % per step)
            store = TileStore(np.array([dp.value for dp in data]), device=device, variant=TileStore.LATENCY_LAYOUT)
        self.store = store
        self.tree_dist = TreeJointDistribution(FSCRPDistribution(concentration_value))
        self.kernel = setup_kernel(outlier_prob, proposal, rng, self.tree_dist, store)
        self.dp_sampler = DataPointSampler(self.tree_dist, rng, outliers=(outlier_prob > 0))
        self.prg_sampler = PruneRegraphSampler(self.tree_dist, rng)
        self.conc_sampler = GammaPriorConcentrationSampler(0.01, 0.01, rng=rng)
        self.burnin_sampler = UnconditionalSMCSampler(self.kernel, num_particles=num_particles,
                                                      resample_threshold=resample_threshold)
        self.tree_sampler = ParticleGibbsTreeSampler(self.kernel, rng, num_particles=num_particles,
                                                     resample_threshold=resample_threshold)
        self.concentration_update = concentration_update
        self.num_samples_data_point = num_samples_data_point
        self.num_samples_prune_regraph = num_samples_prune_regraph
        self.tree = Tree.get_single_node_tree(data, store)
        self.iteration = 0
        # everything alive now (torch, numpy, the data) stays alive for the whole run: keep the cyclic collector from
        # re-walking it on every full collection -- those pauses were 10 % swings on a 60 ms iteration
        gc.collect()
        gc.freeze()

    def _local_moves(self, tree):
        for _ in range(self.num_samples_data_point):
            tree = self.dp_sampler.sample_tree(tree)
        for _ in range(self.num_samples_prune_regraph):
            tree = self.prg_sampler.sample_tree(tree)
        tree.relabel_nodes()
        return tree

    def burnin_step(self):
        """run.py:323-361"""
        self.kernel.clear_proposal_dist_caches()
        self.store.reset()  # new arena generation: the carried tree is only read structurally from here on
        self.tree = self._local_moves(self.burnin_sampler.sample_tree(self.tree))

    def step(self):
        """run.py:261-288: PG sweep, data-point Gibbs, prune-regraph, relabel, concentration update."""
        self.kernel.clear_proposal_dist_caches()
        self.store.reset()
        self.rng.random()  # run.py:268: the subtree-move coin is drawn even when its probability is 0
        tree = self.tree = self._local_moves(self.tree_sampler.sample_tree(self.tree))
        if self.concentration_update:
            node_sizes = [len(v) for k, v in tree.node_data.items() if k != tree.outlier_node_name]
            self.tree_dist.prior.alpha = self.conc_sampler.sample(self.tree_dist.prior.alpha, len(node_sizes),
                                                                  sum(node_sizes))
        self.iteration += 1

    def snapshot(self):
        """State needed to repeat the coming iterations exactly (the carried tree is never mutated by a step)."""
        return {"tree": self.tree, "rng": self.rng.bit_generator.state, "alpha": self.tree_dist.prior.alpha,
                "iteration": self.iteration}

    def restore(self, snap):
        self.tree = snap["tree"]
        self.rng.bit_generator.state = snap["rng"]
        self.tree_dist.prior.alpha = snap["alpha"]
        self.iteration = snap["iteration"]

    def log_p_one(self):
        return self.tree_dist.log_p_one(self.tree)

    def trace_entry(self, i, elapsed):
        """run.py:293-302"""
        return {"iter": i, "time": elapsed, "alpha": self.tree_dist.prior.alpha, "log_p_one": self.log_p_one(),
                "tree": self.tree.to_dict()}

    @property
    def stats(self):
        return dict(self.store.stats, extensions=self.kernel.num_extensions)


def run_phyclone_chain(burnin, concentration_update, concentration_value, data, max_time, num_iters, num_particles,
                       num_samples_data_point, num_samples_prune_regraph, outlier_prob, print_freq, proposal,
                       resample_threshold, rng, samples, thin, chain_num, subtree_update_prob, device=None,
                       store=None, on_iteration=None):
    """run.py:177-232 (+ :235-380).  Returns {"data", "samples", "trace", "chain_num", "stats"}."""
    if subtree_update_prob != 0.0:
        raise NotImplementedError("the subtree particle-Gibbs move is outside the B200 hot path (DESIGN.md section 7)")
    chain = Chain(data, rng, num_particles=num_particles, proposal=proposal, resample_threshold=resample_threshold,
                  outlier_prob=outlier_prob, concentration_value=concentration_value,
                  concentration_update=concentration_update, num_samples_data_point=num_samples_data_point,
                  num_samples_prune_regraph=num_samples_prune_regraph, device=device, store=store)
    timer = Timer()
    for i in range(burnin):
        with timer:
            chain.burnin_step()
            if timer.elapsed > max_time:
                break
    trace = [chain.trace_entry(0, timer.elapsed)]
    for i in range(num_iters):
        with timer:
            chain.step()
            if i % thin == 0:
                trace.append(chain.trace_entry(i, timer.elapsed))
            if on_iteration is not None:
                on_iteration(i, trace[-1])
            if timer.elapsed >= max_time:
                break
    return {"data": data, "samples": samples, "trace": trace, "chain_num": chain_num, "stats": chain.stats}


# ------------------------------------------------------------------------------------------------
# several chains: one process per GPU, chains sharded over ranks, traces gathered at the end
# ------------------------------------------------------------------------------------------------
def chain_generators(seed, num_chains):
    """run.py:57,87-113: one chain uses the main Generator itself, several use its spawned children."""
    rng_main = np.random.default_rng(seed)
    return [rng_main] if num_chains == 1 else rng_main.spawn(num_chains)


def chains_of_rank(num_chains, rank, world_size):
    """Chains are independent (run.py:111-149), so they shard with no data-path collective: chain c runs on
    rank c % world_size."""
    return [c for c in range(num_chains) if c % world_size == rank]


def gather_results(results, rank, world_size, device=None):
    """The one collective of the run: every rank's {chain_num: result} dict, pickled, gathered on rank 0
    (NCCL over NVLink when the process group is NCCL; traces are MBs, so this is latency-irrelevant)."""
    import pickle

    import torch
    import torch.distributed as dist

    if world_size == 1:
        return results
    payload = np.frombuffer(pickle.dumps(results, protocol=pickle.HIGHEST_PROTOCOL), dtype=np.uint8).copy()
    on_gpu = dist.get_backend() == "nccl"
    dev = device if on_gpu else torch.device("cpu")
    n = torch.tensor([payload.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world_size)]
    dist.all_gather(sizes, n)
    cap = int(max(s.item() for s in sizes))
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    buf[: payload.size] = torch.from_numpy(payload).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    merged = {}
    for r in range(world_size):
        part = pickle.loads(bufs[r][: int(sizes[r].item())].cpu().numpy().tobytes())
        merged.update(part)
    return dict(sorted(merged.items()))


def run_chains(data, samples, num_chains=1, seed=None, burnin=100, num_iters=5000, num_particles=80,
               concentration_value=1.0, concentration_update=True, max_time=float("inf"), num_samples_data_point=1,
               num_samples_prune_regraph=1, outlier_prob=0, print_freq=100, proposal="semi-adapted",
               resample_threshold=0.5, thin=1, subtree_update_prob=0.0, rank=0, world_size=1, device=None,
               store_factory=None):
    """The chain part of the reference's `run()` (run.py:28-149) with its ProcessPoolExecutor replaced by
    rank sharding: call it from every rank of a torch.distributed job (or alone).  Returns {chain_num: result}
    on rank 0 and None elsewhere."""
    rngs = chain_generators(seed, num_chains)
    results = {}
    for c in chains_of_rank(num_chains, rank, world_size):
        store = None if store_factory is None else store_factory(np.array([dp.value for dp in data]))
        results[c] = run_phyclone_chain(burnin, concentration_update, concentration_value, data, max_time, num_iters,
                                        num_particles, num_samples_data_point, num_samples_prune_regraph, outlier_prob,
                                        print_freq, proposal, resample_threshold, rngs[c], samples, thin, c,
                                        subtree_update_prob, device=device, store=store)
    return gather_results(results, rank, world_size, device=device)
