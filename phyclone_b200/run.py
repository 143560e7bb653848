"""Chain driver (reference run.py:177-380): burn-in, particle-Gibbs main loop, trace building.

`run_phyclone_chain` keeps the reference's positional signature; one chain drives one CUDA stream
and one TileStore (the device arena holding every tile of the chain).
"""

import gc
import os
import sys
import time

import numpy as np

from .mcmc.concentration import GammaPriorConcentrationSampler
from .mcmc.gibbs_mh import DataPointSampler, PruneRegraphSampler
from .mcmc.particle_gibbs import ParticleGibbsSubtreeSampler, ParticleGibbsTreeSampler
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


def setup_kernel(outlier_prob, proposal, rng, tree_dist, store, perm_dist=None):
    """run.py:403-417"""
    outlier_proposal_prob = 0.1 if outlier_prob > 0 else 0
    return KERNELS[proposal](tree_dist, rng, store, outlier_proposal_prob=outlier_proposal_prob, perm_dist=perm_dist)


class Chain(object):
    """One MCMC chain = one Python thread driving one CUDA stream and one device arena (run.py:177-380).
    `burnin_step()` / `step()` are one trip of the reference's burn-in / main loops."""

    def __init__(self, data, rng, num_particles=100, proposal="semi-adapted", resample_threshold=0.5, outlier_prob=0.0,
                 concentration_value=1.0, concentration_update=True, num_samples_data_point=1,
                 num_samples_prune_regraph=1, device=None, store=None, sweep="host", subtree_update_prob=0.0,
                 perm_dist=None):
        self.data = data
        self.rng = rng
        # the tree walks recurse once per level and a forest over T data points can be T deep
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * len(data) + 1000))
        if store is None:
            # a chain launches a handful of jobs at a time: the layout with the shortest launches (-2.7 % per step)
            store = TileStore(np.array([dp.value for dp in data]), device=device, variant=TileStore.LATENCY_LAYOUT)
        self.store = store
        # a device-mode chain is defined up to rounding (uniform-driven draws): its launches of a few hundred jobs take
        # the latency mapping of the job kernel.  Host mode keeps the default mapping (bit-exact golden traces).
        if sweep == "device" and os.environ.get("PCB_LATENCY_KERNEL", "1") != "0" and hasattr(store.arena, "set_latency"):
            store.arena.set_latency(True)
        self.tree_dist = TreeJointDistribution(FSCRPDistribution(concentration_value))
        self.kernel = setup_kernel(outlier_prob, proposal, rng, self.tree_dist, store, perm_dist=perm_dist)
        self.dp_sampler = DataPointSampler(self.tree_dist, rng, outliers=(outlier_prob > 0))
        self.prg_sampler = PruneRegraphSampler(self.tree_dist, rng)
        self.conc_sampler = GammaPriorConcentrationSampler(0.01, 0.01, rng=rng)
        # sweep="device": the SMC sweeps run device-resident with uniform-driven draws (smc/device_sweep.py); only the
        # semi-adapted proposal has that form, the other proposals keep the host sampler
        if sweep == "device" and proposal != "semi-adapted":
            raise ValueError("sweep='device' needs the semi-adapted proposal")
        self.sweep = sweep
        self.burnin_sampler = UnconditionalSMCSampler(self.kernel, num_particles=num_particles,
                                                      resample_threshold=resample_threshold, sweep=sweep)
        self.tree_sampler = ParticleGibbsTreeSampler(self.kernel, rng, num_particles=num_particles,
                                                     resample_threshold=resample_threshold, sweep=sweep)
        self.subtree_sampler = ParticleGibbsSubtreeSampler(self.kernel, rng, num_particles=num_particles,
                                                           resample_threshold=resample_threshold)
        self.subtree_update_prob = subtree_update_prob
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
        if self.rng.random() < self.subtree_update_prob:  # run.py:268-271 (the coin is drawn even at probability 0)
            # the subtree move edits the carried tree itself: its tiles must survive into the new arena generation
            carried = self.tree.copy()
            self.store.flush()
            self.store.rebase(carried)
            tree = self.subtree_sampler.sample_tree(carried)
        else:
            self.store.reset()  # new arena generation: the carried tree is only read structurally from here on
            tree = self.tree_sampler.sample_tree(self.tree)
        tree = self.tree = self._local_moves(tree)
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
        d = dict(self.store.stats, extensions=self.kernel.num_extensions)
        for name, sampler in (("sweep", self.tree_sampler), ("burnin_sweep", self.burnin_sampler)):
            dev = getattr(sampler, "_device", None)
            if dev is not None:
                d[name] = dict(dev.stats)
        return d


def run_phyclone_chain(burnin, concentration_update, concentration_value, data, max_time, num_iters, num_particles,
                       num_samples_data_point, num_samples_prune_regraph, outlier_prob, print_freq, proposal,
                       resample_threshold, rng, samples, thin, chain_num, subtree_update_prob, device=None,
                       store=None, on_iteration=None, sweep="host"):
    """run.py:177-232 (+ :235-380).  Returns {"data", "samples", "trace", "chain_num", "stats"}."""
    chain = Chain(data, rng, num_particles=num_particles, proposal=proposal, resample_threshold=resample_threshold,
                  outlier_prob=outlier_prob, concentration_value=concentration_value,
                  concentration_update=concentration_update, num_samples_data_point=num_samples_data_point,
                  num_samples_prune_regraph=num_samples_prune_regraph, device=device, store=store, sweep=sweep,
                  subtree_update_prob=subtree_update_prob)
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


# ------------------------------------------------------------------------------------------------
# several chains on the GPUs of one box from ONE call: the reference's ProcessPoolExecutor (run.py:111-149), with the
# processes sharing each GPU through MPS
# ------------------------------------------------------------------------------------------------
class MpsServer(object):
    """Context manager around `nvidia-cuda-mps-control`.

    One chain keeps a B200 ~30 % busy with launches of a few dozen single-warp CTAs.  Chains in separate processes are
    separate CUDA contexts, which the GPU time-slices: their kernels never run side by side and the aggregate
    saturates at ~1.5x one chain (profiles/r1_multi_chain_one_gpu.md).  Under MPS the processes share one context and
    their launches overlap: 16 chains on one B200 ran at 311 it/s against 24 for one.  The daemon is started before
    any worker touches CUDA and shut down on exit; if the binary is missing or refuses to start, the chains simply
    run time-sliced."""

    def __init__(self, enable=True):
        self.enable = enable
        self.started = False
        self._dir = None
        self._saved = {}

    def __enter__(self):
        import shutil
        import subprocess
        import tempfile

        if not self.enable or shutil.which("nvidia-cuda-mps-control") is None:
            return self
        if os.environ.get("CUDA_MPS_PIPE_DIRECTORY") and os.path.exists(
                os.path.join(os.environ["CUDA_MPS_PIPE_DIRECTORY"], "control")):
            return self  # somebody else's daemon is already serving this environment: use it, do not stop it
        self._dir = tempfile.mkdtemp(prefix="pcb_mps_")
        for key, sub in (("CUDA_MPS_PIPE_DIRECTORY", "pipe"), ("CUDA_MPS_LOG_DIRECTORY", "log")):
            self._saved[key] = os.environ.get(key)
            os.environ[key] = os.path.join(self._dir, sub)
            os.makedirs(os.environ[key], exist_ok=True)
        try:
            rc = subprocess.run(["nvidia-cuda-mps-control", "-d"], capture_output=True, timeout=30).returncode
            self.started = rc == 0
        except Exception:  # noqa: BLE001
            self.started = False
        if not self.started:
            self._restore_env()
        return self

    def _restore_env(self):
        for key, val in self._saved.items():
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
        self._saved = {}

    def __exit__(self, *exc):
        import shutil
        import subprocess

        if self.started:
            try:
                subprocess.run(["nvidia-cuda-mps-control"], input=b"quit\n", capture_output=True, timeout=60)
                for _ in range(20):  # the daemon takes a second or two to go away: do not hand back a half-stopped one
                    probe = subprocess.run(["nvidia-cuda-mps-control"], input=b"get_server_list\n", capture_output=True,
                                           timeout=10)
                    if probe.returncode != 0:
                        break
                    time.sleep(0.25)
            except Exception:  # noqa: BLE001
                pass
            self.started = False
            self._restore_env()
        if self._dir is not None:
            shutil.rmtree(self._dir, ignore_errors=True)
            self._dir = None


def _chain_worker(args):
    (chain_num, num_chains, seed, device_index, data, samples, kw) = args
    import torch

    torch.cuda.set_device(device_index)
    rng = chain_generators(seed, num_chains)[chain_num]
    return chain_num, run_phyclone_chain(
        kw["burnin"], kw["concentration_update"], kw["concentration_value"], data, kw["max_time"], kw["num_iters"],
        kw["num_particles"], kw["num_samples_data_point"], kw["num_samples_prune_regraph"], kw["outlier_prob"],
        kw["print_freq"], kw["proposal"], kw["resample_threshold"], rng, samples, kw["thin"], chain_num,
        kw["subtree_update_prob"], device=torch.device("cuda", device_index))


def run_chains_local(data, samples, num_chains=1, seed=None, devices=None, mps=True, burnin=100, num_iters=5000,
                     num_particles=80, concentration_value=1.0, concentration_update=True, max_time=float("inf"),
                     num_samples_data_point=1, num_samples_prune_regraph=1, outlier_prob=0, print_freq=100,
                     proposal="semi-adapted", resample_threshold=0.5, thin=1, subtree_update_prob=0.0):
    """All chains of a run from one call, one worker process per chain (the reference's `run()`, run.py:111-149),
    chain c on GPU `devices[c % len(devices)]` (default: every visible GPU), the processes of a GPU sharing it
    through MPS.  Same chain seeding as `run_chains` / the reference (`default_rng(seed).spawn(num_chains)`, the main
    generator itself for a single chain), so a chain's trace does not depend on how the chains are placed.
    The calling process must not have initialised CUDA.  Returns {chain_num: result}."""
    import multiprocessing as mp

    if devices is None:
        from . import _lib

        devices = list(range(max(1, _lib.load().pcb_device_count())))
    kw = dict(burnin=burnin, num_iters=num_iters, num_particles=num_particles, concentration_value=concentration_value,
              concentration_update=concentration_update, max_time=max_time, num_samples_data_point=num_samples_data_point,
              num_samples_prune_regraph=num_samples_prune_regraph, outlier_prob=outlier_prob, print_freq=print_freq,
              proposal=proposal, resample_threshold=resample_threshold, thin=thin, subtree_update_prob=subtree_update_prob)
    jobs = [(c, num_chains, seed, devices[c % len(devices)], data, samples, kw) for c in range(num_chains)]
    with MpsServer(enable=mps and num_chains > len(devices)):
        ctx = mp.get_context("spawn")
        with ctx.Pool(processes=num_chains) as pool:
            results = dict(pool.map(_chain_worker, jobs, chunksize=1))
    return dict(sorted(results.items()))
