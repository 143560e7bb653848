"""Particle Gibbs tree sampler (reference mcmc/particle_gibbs.py:7-48)."""

from ..smc.samplers import ConditionalSMCSampler
from ..smc.utils import RootPermutationDistribution
from ..utils.math import discrete_rvs


class ParticleGibbsTreeSampler(object):
    def __init__(self, kernel, rng, num_particles=10, resample_threshold=0.5, sweep="host"):
        self.kernel = kernel
        self.num_particles = num_particles
        self.resample_threshold = resample_threshold
        self._rng = rng
        self._device = None
        if sweep == "device":
            from ..smc.device_sweep import DeviceSweep

            self._device = DeviceSweep(kernel, num_particles, resample_threshold)
        elif sweep != "host":
            raise ValueError("sweep must be 'host' or 'device'")

    def sample_tree(self, tree):
        if self._device is not None:
            return self._sample_tree_on_device(tree)
        return self._sample_tree_from_swarm(self.sample_swarm(tree))

    def _sample_tree_on_device(self, tree):
        """The same move with the sweep itself on the device (smc/device_sweep.py): sigma and the retained path are
        built here, one seed is drawn from the chain's Generator, everything between the first proposal and the
        final pick happens in `pcb_smc_sweep_dev`."""
        from ..smc.device_sweep import SweepFallback

        data_sigma = RootPermutationDistribution.sample(tree, self._rng)
        seed = int(self._rng.integers(0, 1 << 62))
        sampler = ConditionalSMCSampler(tree, data_sigma, self.kernel, num_particles=self.num_particles,
                                        resample_threshold=self.resample_threshold)
        try:
            return self._device.run(data_sigma, seed, sampler.constrained_path)
        except SweepFallback:
            self._device.stats["fallbacks"] += 1
            return self._sample_tree_from_swarm(sampler.sample())

    def sample_swarm(self, tree):
        data_sigma = RootPermutationDistribution.sample(tree, self._rng)
        sampler = ConditionalSMCSampler(tree, data_sigma, self.kernel, num_particles=self.num_particles,
                                        resample_threshold=self.resample_threshold)
        return sampler.sample()

    def _sample_tree_from_swarm(self, swarm):
        particle_idx = discrete_rvs(swarm.weights, self._rng)
        return swarm.particles[particle_idx].tree
