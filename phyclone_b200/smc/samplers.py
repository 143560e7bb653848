"""SMC samplers (reference smc/samplers/{base,standard,conditional,unconditional}.py).

Same control flow and RNG call order as the reference; per step the work is arranged as
  1. build the proposal of every distinct parent (host bookkeeping, recorded tile ops),
  2. draw every particle's move in swarm order (first read of q => ONE flush for step 1),
  3. read the incremental weights (ONE flush for all new-node trees),
so an SMC step costs two synchronising device round trips whatever the number of particles.
"""

import ctypes as C

import numpy as np

from .. import _hostlib
from .._hostmod import load as _load_hostmod
from ..tree.tree import Tree
from ..utils.math import discrete_rvs
from .forest_state import host_tuple
from .swarm import ParticleSwarm, StateHolder, TreeHolder
from .utils import RootPermutationDistribution

_hm = _load_hostmod()
_draws = []


def _draws_address():
    """Address of pcb_smc_draws (csrc/pcb_host.cpp) for the host extension to call."""
    if not _draws:
        _draws.append(C.cast(_hostlib.load().pcb_smc_draws, C.c_void_p).value)
    return _draws[0]


class AbstractSMCSampler(object):
    """smc/samplers/base.py"""

    def __init__(self, data_points, kernel, num_particles, resample_threshold=0.5):
        self.data_points = data_points
        self.kernel = kernel
        self.num_particles = num_particles
        self.resample_threshold = resample_threshold
        self.iteration = 0
        self.num_iterations = len(data_points)
        self.swarm = None
        self._rng = kernel.rng

    def sample(self):
        if self.kernel.perm_dist is not None:
            raise NotImplementedError("perm_dist is implemented by the device-resident sweep only (sweep='device')")
        self._init_swarm()
        self._resample_swarm()
        while self.iteration < self.num_iterations:
            self._update_swarm()
            if self.iteration < self.num_iterations - 1:
                self._resample_swarm()
            self.iteration += 1
        return self.swarm

    def _prepare_proposals(self, parents):
        """Step 1 of the module docstring: no RNG draw, no device synchronisation."""
        data_point = self.data_points[self.iteration]
        seen = set()
        for p in parents:
            key = None if p is None else p.holder.vkey
            if key not in seen:
                seen.add(key)
                self.kernel.get_proposal_distribution(data_point, p)

    def _propose_particle(self, parent_particle):
        return self.kernel.propose_particle(self.data_points[self.iteration], parent_particle)

    def _propose_all(self, parents):
        """`[propose_particle(dp, p) for p in parents]` (smc/kernels/base.py:59-67).  When the kernel's proposals
        are the state-based semi-adapted ones and the RNG is a real numpy Generator, all the draws of the step
        (coin, multinomial, integers, choice -- in swarm order, from the same bit generator) are made by ONE call
        into csrc/pcb_host.cpp instead of ~4 Python-level Generator calls per particle."""
        kernel = self.kernel
        data_point = self.data_points[self.iteration]
        addr = kernel.fast_draw_address()
        n = len(parents)
        if addr is None or n == 0:
            self._prepare_proposals(parents)
            return [self._propose_particle(p) for p in parents]
        # a resampled swarm lists the same Particle object many times: resolve each object once
        dists, index, by_obj, dist_l = [], {}, {}, []
        for p in parents:
            j = by_obj.get(id(p))
            if j is None:
                key = None if p is None else p.holder.vkey
                j = index.get(key)
                if j is None:
                    d = kernel.get_proposal_distribution(data_point, p)
                    if not getattr(d, "fast_draws", False):
                        return [self._propose_particle(q) for q in parents]
                    j = index[key] = len(dists)
                    dists.append(d)
                by_obj[id(p)] = j
            dist_l.append(j)
        for d in dists:
            if d._q_dist is None:
                d._finish()  # the first one resolves the scores of every proposal of the step: one flush
        meta = [(d.parent_is_empty_tree, float(d.log_half),
                 0.0 if d.parent_is_empty_tree else float(d._cached_log_old_num_roots), len(d.roots_array)) for d in dists]
        store = kernel.store
        make_new = (StateHolder, data_point, host_tuple(store, data_point), float(kernel.tree_dist.prior.log_alpha),
                    [(store, d._pstate, d._pholder) for d in dists])
        out = _hm.propose_step(_draws_address(), addr, parents, dist_l, [d._q_dist for d in dists],
                               [d.roots_array for d in dists], [d._curr_trees for d in dists],
                               [d._log_q for d in dists], meta, make_new)
        kernel.num_extensions += n
        return out

    def _get_log_ws(self, particles):
        """`_get_log_w` for a list of particles in one call into the host extension."""
        return _hm.particle_log_ws(particles, self.iteration >= self.num_iterations - 1)

    def _get_log_w(self, particle):
        """base.py:52-58"""
        if self.iteration < self.num_iterations - 1:
            return particle.log_w
        return particle.log_w - particle.log_p + particle.log_p_one


class SMCSampler(AbstractSMCSampler):
    """smc/samplers/standard.py"""

    def _init_swarm(self):
        self.swarm = ParticleSwarm()
        uniform_weight = -np.log(self.num_particles)
        for _ in range(self.num_particles):
            self.swarm.add_particle(uniform_weight, None)

    def _resample_swarm(self):
        if self.swarm.relative_ess <= self.resample_threshold:
            log_uniform_weight = -np.log(self.num_particles)
            multiplicities = self._rng.multinomial(self.num_particles, self.swarm.weights)
            particles = [p for p, m in zip(self.swarm.particles, multiplicities.tolist()) for _ in range(m)]
            self.swarm = ParticleSwarm.of(particles, [log_uniform_weight] * len(particles))

    def _update_swarm(self):
        log_W = self.swarm.log_weights
        particles = self._propose_all(self.swarm.particles)
        self.swarm = ParticleSwarm.of(particles, [a + b for a, b in zip(log_W, self._get_log_ws(particles))])


class ConditionalSMCSampler(AbstractSMCSampler):
    """smc/samplers/conditional.py"""

    def __init__(self, current_tree, data_points, kernel, num_particles, resample_threshold=0.5, with_proposals=True):
        """`with_proposals=False` (the device-resident sweep): only the retained trees are built; their proposal
        log-probabilities -- hence the particles' weights -- are computed on the device, from `path_edits`."""
        super().__init__(data_points, kernel, num_particles, resample_threshold=resample_threshold)
        self.path_edits = []  # per data point: (kind, node name | number of adopted roots) of the retained path
        self.constrained_path = self._get_constrained_path(current_tree, with_proposals)

    def _get_constrained_path(self, tree, with_proposals=True):
        """conditional.py:19-74.  All T retained trees and their proposals are built first (they do not
        depend on any score), the proposal probabilities are read afterwards: one flush for the whole path."""
        kernel = self.kernel
        data_to_node = tree.labels
        node_map = {}
        new_tree = Tree(tree.grid_size, kernel.store)
        outlier_node_name = tree.outlier_node_name
        holders = []
        # `thawed` follows `new_tree` through the same edits but is re-read (from_dict round trip) after every one:
        # it is what holder.tree must return, and building it incrementally keeps that O(touched nodes) per step --
        # thawing each as-built tree from scratch re-derives every node touched so far, O(T) per step (1 s per
        # iteration at 2000 data points).  Both start empty and have no vacancies, so the edits allocate the same
        # node names and slots in both, and tile ids are a function of structure (memoised ops): same result.
        thawed = new_tree.copy()
        for data_point in self.data_points:
            new_tree = new_tree.copy()
            thawed = thawed.copy()
            old_node = data_to_node[data_point.idx]
            if old_node == outlier_node_name:
                new_tree.add_data_point_to_outliers(data_point)
                thawed.add_data_point_to_outliers(data_point)
                self.path_edits.append((2, 0))
            elif old_node in node_map:
                new_tree.add_data_point_to_node(data_point, node_map[old_node])
                thawed.add_data_point_to_node(data_point, node_map[old_node])
                self.path_edits.append((0, int(node_map[old_node])))
            else:
                children = [node_map[child] for child in tree.get_children(old_node)]
                new_node = new_tree.create_root_node(children)
                node_map[old_node] = new_node
                new_tree.add_data_point_to_node(data_point, new_node)
                twin = thawed.create_root_node(children)
                assert twin == new_node
                thawed.add_data_point_to_node(data_point, new_node)
                self.path_edits.append((1, len(children)))
            thawed.outliers  # noqa: B018 -- same key-creation side effect as scoring the as-built tree
            thawed = thawed.thawed_copy()
            holder = TreeHolder(new_tree, kernel.tree_dist)
            holder._thawed = thawed
            holders.append(holder)
        constrained_path = [None]
        if not with_proposals:
            for holder in holders:
                constrained_path.append(kernel.create_particle(None, constrained_path[-1], holder))
            return constrained_path
        parent_tree = None
        pending = []
        for data_point, holder in zip(self.data_points, holders):
            parent_particle = constrained_path[-1]
            proposal_dist = kernel.get_proposal_distribution(data_point, parent_particle, parent_tree)
            particle = kernel.create_particle(None, parent_particle, holder)
            pending.append((particle, proposal_dist, holder))
            constrained_path.append(particle)
            parent_tree = holder._built
        for particle, proposal_dist, holder in pending:
            particle._log_q = proposal_dist.log_p(holder)
        return constrained_path

    def _init_swarm(self):
        self.swarm = ParticleSwarm()
        uniform_weight = -np.log(self.num_particles)
        self.swarm.add_particle(uniform_weight, self.constrained_path[1])
        for particle in self._propose_all([None] * (self.num_particles - 1)):
            self.swarm.add_particle(uniform_weight, particle)
        self.iteration += 1

    def _resample_swarm(self):
        if self.swarm.relative_ess <= self.resample_threshold:
            log_uniform_weight = -np.log(self.num_particles)
            multiplicities = self._rng.multinomial(self.num_particles - 1, self.swarm.weights)
            particles = [self.constrained_path[self.iteration + 1]]
            particles += [p for p, m in zip(self.swarm.particles, multiplicities.tolist()) for _ in range(m)]
            self.swarm = ParticleSwarm.of(particles, [log_uniform_weight] * len(particles))

    def _update_swarm(self):
        log_W = self.swarm.log_weights
        retained = self.constrained_path[self.iteration + 1]
        particles = [retained] + self._propose_all(self.swarm.particles[1:])
        self.swarm = ParticleSwarm.of(particles, [a + b for a, b in zip(log_W, self._get_log_ws(particles))])


class UnconditionalSMCSampler(object):
    """smc/samplers/unconditional.py"""

    def __init__(self, kernel, num_particles=20, resample_threshold=0.5, sweep="host"):
        self.kernel = kernel
        self.num_particles = num_particles
        self.resample_threshold = resample_threshold
        self._rng = kernel.rng
        self._device = None
        if sweep == "device":
            from .device_sweep import DeviceSweep

            self._device = DeviceSweep(kernel, num_particles, resample_threshold)
        elif sweep != "host":
            raise ValueError("sweep must be 'host' or 'device'")

    def sample_tree(self, tree):
        data_sigma = RootPermutationDistribution.sample(tree, self._rng)
        if self._device is not None:
            from .device_sweep import SweepFallback

            seed = int(self._rng.integers(0, 1 << 62))
            try:
                return self._device.run(data_sigma, seed, None)
            except SweepFallback:
                self._device.stats["fallbacks"] += 1
        smc_sampler = SMCSampler(data_sigma, self.kernel, num_particles=self.num_particles,
                                 resample_threshold=self.resample_threshold)
        swarm = smc_sampler.sample()
        idx = discrete_rvs(swarm.weights, self._rng)
        return swarm.particles[idx].tree
