"""SMC samplers (reference smc/samplers/{base,standard,conditional,unconditional}.py).

Same control flow and RNG call order as the reference; per step the work is arranged as
  1. build the proposal of every distinct parent (host bookkeeping, recorded tile ops),
  2. draw every particle's move in swarm order (first read of q => ONE flush for step 1),
  3. read the incremental weights (ONE flush for all new-node trees),
so an SMC step costs two synchronising device round trips whatever the number of particles.
"""

import numpy as np

from ..tree.tree import Tree
from ..utils.math import discrete_rvs
from .swarm import ParticleSwarm, TreeHolder
from .utils import RootPermutationDistribution


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
            new_swarm = ParticleSwarm()
            log_uniform_weight = -np.log(self.num_particles)
            multiplicities = self._rng.multinomial(self.num_particles, self.swarm.weights)
            for particle, multiplicity in zip(self.swarm.particles, multiplicities):
                for _ in range(multiplicity):
                    new_swarm.add_particle(log_uniform_weight, particle)
            self.swarm = new_swarm

    def _update_swarm(self):
        self._prepare_proposals(self.swarm.particles)
        log_W = self.swarm.log_weights
        particles = [self._propose_particle(p) for p in self.swarm.particles]
        new_swarm = ParticleSwarm()
        for parent_log_W, particle in zip(log_W, particles):
            new_swarm.add_particle(parent_log_W + self._get_log_w(particle), particle)
        self.swarm = new_swarm


class ConditionalSMCSampler(AbstractSMCSampler):
    """smc/samplers/conditional.py"""

    def __init__(self, current_tree, data_points, kernel, num_particles, resample_threshold=0.5):
        super().__init__(data_points, kernel, num_particles, resample_threshold=resample_threshold)
        self.constrained_path = self._get_constrained_path(current_tree)

    def _get_constrained_path(self, tree):
        """conditional.py:19-74.  All T retained trees and their proposals are built first (they do not
        depend on any score), the proposal probabilities are read afterwards: one flush for the whole path."""
        kernel = self.kernel
        data_to_node = tree.labels
        node_map = {}
        new_tree = Tree(tree.grid_size, kernel.store)
        outlier_node_name = tree.outlier_node_name
        holders = []
        for data_point in self.data_points:
            new_tree = new_tree.copy()
            old_node = data_to_node[data_point.idx]
            if old_node == outlier_node_name:
                new_tree.add_data_point_to_outliers(data_point)
            elif old_node in node_map:
                new_tree.add_data_point_to_node(data_point, node_map[old_node])
            else:
                children = [node_map[child] for child in tree.get_children(old_node)]
                new_node = new_tree.create_root_node(children)
                node_map[old_node] = new_node
                new_tree.add_data_point_to_node(data_point, new_node)
            holders.append(TreeHolder(new_tree, kernel.tree_dist))
        constrained_path = [None]
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
        self._prepare_proposals([None])
        for _ in range(self.num_particles - 1):
            self.swarm.add_particle(uniform_weight, self._propose_particle(None))
        self.iteration += 1

    def _resample_swarm(self):
        if self.swarm.relative_ess <= self.resample_threshold:
            new_swarm = ParticleSwarm()
            log_uniform_weight = -np.log(self.num_particles)
            multiplicities = self._rng.multinomial(self.num_particles - 1, self.swarm.weights)
            new_swarm.add_particle(log_uniform_weight, self.constrained_path[self.iteration + 1])
            for particle, multiplicity in zip(self.swarm.particles, multiplicities):
                for _ in range(multiplicity):
                    new_swarm.add_particle(log_uniform_weight, particle)
            self.swarm = new_swarm

    def _update_swarm(self):
        self._prepare_proposals(self.swarm.particles[1:])
        log_W = self.swarm.log_weights
        retained = self.constrained_path[self.iteration + 1]
        particles = [retained] + [self._propose_particle(p) for p in self.swarm.particles[1:]]
        new_swarm = ParticleSwarm()
        for parent_log_W, particle in zip(log_W, particles):
            new_swarm.add_particle(parent_log_W + self._get_log_w(particle), particle)
        self.swarm = new_swarm


class UnconditionalSMCSampler(object):
    """smc/samplers/unconditional.py"""

    def __init__(self, kernel, num_particles=20, resample_threshold=0.5):
        self.kernel = kernel
        self.num_particles = num_particles
        self.resample_threshold = resample_threshold
        self._rng = kernel.rng

    def sample_tree(self, tree):
        data_sigma = RootPermutationDistribution.sample(tree, self._rng)
        smc_sampler = SMCSampler(data_sigma, self.kernel, num_particles=self.num_particles,
                                 resample_threshold=self.resample_threshold)
        swarm = smc_sampler.sample()
        idx = discrete_rvs(swarm.weights, self._rng)
        return swarm.particles[idx].tree
