"""Particle Gibbs tree sampler (reference mcmc/particle_gibbs.py:7-48)."""

from ..smc.samplers import ConditionalSMCSampler
from ..smc.utils import RootPermutationDistribution
from ..utils.math import discrete_rvs


class ParticleGibbsTreeSampler(object):
    def __init__(self, kernel, rng, num_particles=10, resample_threshold=0.5):
        self.kernel = kernel
        self.num_particles = num_particles
        self.resample_threshold = resample_threshold
        self._rng = rng

    def sample_tree(self, tree):
        return self._sample_tree_from_swarm(self.sample_swarm(tree))

    def sample_swarm(self, tree):
        data_sigma = RootPermutationDistribution.sample(tree, self._rng)
        sampler = ConditionalSMCSampler(tree, data_sigma, self.kernel, num_particles=self.num_particles,
                                        resample_threshold=self.resample_threshold)
        return sampler.sample()

    def _sample_tree_from_swarm(self, swarm):
        particle_idx = discrete_rvs(swarm.weights, self._rng)
        return swarm.particles[particle_idx].tree
