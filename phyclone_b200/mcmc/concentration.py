"""Gibbs update of the CRP concentration (reference mcmc/concentration.py:11-60); scalar, host only."""

import numpy as np
from scipy.stats import bernoulli, beta, gamma


class GammaPriorConcentrationSampler(object):
    def __init__(self, a, b, rng):
        self.a = a
        self.b = b
        self._rng = rng

    def sample(self, old_value, num_clusters, num_data_points):
        rs = getattr(self._rng, "generator", self._rng)
        if num_clusters == 0:
            return gamma.rvs(self.a, scale=(1 / self.b), random_state=rs)
        a, b, k, n = self.a, self.b, num_clusters, num_data_points
        eta = beta.rvs(a=old_value + 1, b=n, random_state=rs)
        shape = a + k - 1
        rate = b - np.log(eta)
        x = shape / (n * rate)
        pi = x / (1 + x)
        shape += bernoulli.rvs(pi, random_state=rs)
        new_value = gamma.rvs(shape, scale=(1 / rate), random_state=rs)
        return max(new_value, 1e-10)
