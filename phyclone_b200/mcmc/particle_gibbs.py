"""Particle Gibbs tree samplers (reference mcmc/particle_gibbs.py:7-109)."""

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
        from ..smc.device_sweep import RetainedPath, SweepFallback

        data_sigma = RootPermutationDistribution.sample(tree, self._rng)
        seed = int(self._rng.integers(0, 1 << 62))
        try:
            if self._device.perm:
                # the root-permutation density is a function of whole trees: build the retained trees
                sampler = ConditionalSMCSampler(tree, data_sigma, self.kernel, num_particles=self.num_particles,
                                                resample_threshold=self.resample_threshold, with_proposals=False)
                return self._device.run(data_sigma, seed, sampler.constrained_path, sampler.path_edits)
            return self._device.run(data_sigma, seed, RetainedPath(tree, data_sigma, self.kernel))
        except SweepFallback:
            self._device.stats["fallbacks"] += 1
            sampler = ConditionalSMCSampler(tree, data_sigma, self.kernel, num_particles=self.num_particles,
                                            resample_threshold=self.resample_threshold)
            return self._sample_tree_from_swarm(sampler.sample())

    def sample_swarm(self, tree):
        data_sigma = RootPermutationDistribution.sample(tree, self._rng)
        sampler = ConditionalSMCSampler(tree, data_sigma, self.kernel, num_particles=self.num_particles,
                                        resample_threshold=self.resample_threshold)
        return sampler.sample()

    def _sample_tree_from_swarm(self, swarm):
        particle_idx = discrete_rvs(swarm.weights, self._rng)
        return swarm.particles[particle_idx].tree


class _Completed(object):
    """A particle of the subtree sweep after its subtree was put back into the rest of the tree."""

    __slots__ = ("holder",)

    def __init__(self, holder):
        self.holder = holder

    @property
    def tree(self):
        return self.holder.tree


class ParticleGibbsSubtreeSampler(ParticleGibbsTreeSampler):
    """mcmc/particle_gibbs.py:51-109: re-sample the subtree under the parent of a node picked through a random data
    point (the outliers travel with the subtree), then re-weight every particle by the full tree it completes.  The
    sweep itself is the host sampler's (the weight correction needs every particle's tree); all completed trees are
    scored by one batched flush."""

    def __init__(self, kernel, rng, num_particles=10, resample_threshold=0.5):
        super().__init__(kernel, rng, num_particles=num_particles, resample_threshold=resample_threshold, sweep="host")

    def sample_tree(self, tree):
        tree = tree.copy()  # the reference prunes its argument in place; callers here keep theirs (run.Chain.snapshot)
        outlier_node_name = tree.outlier_node_name
        nodes = [label for label in tree.labels.values() if label != outlier_node_name]
        subtree_root_child = self._rng.choice(nodes)
        subtree_root = tree.get_parent(subtree_root_child)
        parent = tree.get_parent(subtree_root)
        subtree = tree.get_subtree(subtree_root)
        tree.remove_subtree(subtree)
        for data_point in tree.outliers:
            tree.remove_data_point_from_outliers(data_point)
            subtree.add_data_point_to_outliers(data_point)
        swarm = self.sample_swarm(subtree)
        swarm = self._correct_weights(parent, swarm, tree)
        return self._sample_tree_from_swarm(swarm)

    def _correct_weights(self, parent, swarm, tree):
        """particle_gibbs.py:86-109"""
        from ..smc.swarm import ParticleSwarm, TreeHolder

        tree_dist = self.kernel.tree_dist
        weights, holders = [], []
        for p, w in zip(swarm.particles, swarm.unnormalized_log_weights):
            subtree = p.tree
            new_tree = tree.copy()
            new_tree.add_subtree(subtree, parent=parent)
            for data_point in subtree.outliers:
                new_tree.add_data_point_to_outliers(data_point)
            new_tree.update()
            weights.append(w - p.log_p_one)
            holders.append(TreeHolder(new_tree, tree_dist))
        # the first log_p_one read scores every completed tree in one flush
        return ParticleSwarm.of([_Completed(h) for h in holders], [w + h.log_p_one for w, h in zip(weights, holders)])
