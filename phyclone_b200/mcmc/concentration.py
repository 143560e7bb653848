"""Gibbs update of the CRP concentration alpha under a Gamma(shape, rate) prior: the auxiliary-variable scheme of
Escobar & West as the reference runs it (mcmc/concentration.py:11-60).  Scalar host work, three or four scipy draws
per MCMC iteration; the draws and the arithmetic keep the reference's order so that a chain stays on its RNG stream."""

import numpy as np
from scipy import stats

_FLOOR = 1e-10  # the reference clamps the draw away from zero


class GammaPriorConcentrationSampler(object):
    __slots__ = ("a", "b", "_rng")

    def __init__(self, a, b, rng):
        self.a, self.b, self._rng = a, b, rng

    def _generator(self):
        # the test-only recording / replaying RNGs wrap a numpy Generator; scipy needs the real thing
        return getattr(self._rng, "generator", self._rng)

    def sample(self, old_value, num_clusters, num_data_points):
        source = self._generator()
        if num_clusters == 0:  # nothing to condition on: a draw from the prior
            return stats.gamma.rvs(self.a, scale=(1 / self.b), random_state=source)
        # auxiliary eta | alpha ~ Beta(alpha + 1, n)
        eta = stats.beta.rvs(a=old_value + 1, b=num_data_points, random_state=source)
        rate = self.b - np.log(eta)
        shape = self.a + num_clusters - 1
        # alpha | eta is a two-component Gamma mixture; odds of the component with one more unit of shape
        odds = shape / (num_data_points * rate)
        shape += stats.bernoulli.rvs(odds / (1 + odds), random_state=source)
        return max(stats.gamma.rvs(shape, scale=(1 / rate), random_state=source), _FLOOR)
