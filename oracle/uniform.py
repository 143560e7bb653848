"""Uniform-driven draws for the SMC sweep  --  TEST INFRASTRUCTURE ONLY (oracle side of the device-resident sweep).

The device-resident conditional-SMC sweep (phyclone_b200/csrc/pcb_sweep.cuh) cannot consume a NumPy Generator:
`Generator.multinomial` / `integers` / `choice` use a data-dependent number of raw draws in swarm order.  Its draws
are therefore defined as inverse-CDF transforms of uniforms that are a pure function of
(sweep seed, SMC step t, particle slot i, draw number c):

    u(seed, t, i, c) = splitmix64(seed + GOLDEN * ((t << 40 | i << 20 | c) + 1)) >> 11, times 2**-53

and `SweepRNG` is that definition as an object with the Generator methods the oracle sampler calls
(oracle/sampler.py: `random`, `multinomial`, `integers`, `choice`).  Outside a sweep (`begin` .. `end`) every call
goes to the wrapped NumPy Generator, so the sigma permutation, the Gibbs / prune-regraph moves and the
concentration update draw exactly as in the reference; inside, the sampler announces which (step, slot) is drawing
with `at(t, i)`.  The oracle run with a `SweepRNG` IS the specification of the device mode: the CUDA sweep must
reproduce its decisions (ancestors, chosen candidates, adopted children) for the same seed.

Conventions (mirrored by the kernels):
  * inverse CDF: smallest k with p[0] + ... + p[k] > u, summed left to right in float64; the last index catches
    the remainder.
  * integers(0, n): min(n - 1, floor(u * n)).
  * choice(a, k, replace=False): partial Fisher-Yates on positions 0..len(a)-1, one uniform per pick:
    j = m + min(len - m - 1, floor(u * (len - m))), swap(m, j); result = a[first k positions].
  * multinomial(n, p): n inverse-CDF draws (draw numbers c .. c+n-1), returned as counts.
  * a new root adopts its children in DESCENDING root position of the thawed parent (`canonical_adoption`): every
    adopted edge goes to the head of the adjacency list, so the new node folds its children in root order, which is
    also the order `Tree.from_dict` gives later -- built and re-read trees coincide.  (The reference iterates a
    frozenset here, i.e. CPython's hash-table order: an accident, not part of the algorithm.)
  * the proposal / new-tree LRUs never evict (`unbounded_caches`): a particle whose tree equals the retained
    particle's always gets the retained path's proposal, as in the reference while its LRU still holds it.
"""

import numpy as np

MASK = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15
RESAMPLE_SLOT = (1 << 20) - 1  # pseudo slot of the resampling draws of a step
FINAL_SLOT = (1 << 20) - 2  # pseudo slot of the final pick


def mix64(z):
    z &= MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


def u01(seed, t, i, c):
    idx = ((t << 40) | (i << 20) | c) + 1
    return (mix64(seed + GOLDEN * idx) >> 11) * (1.0 / 9007199254740992.0)


def inv_cdf(p, u):
    acc = 0.0
    n = len(p)
    for k in range(n - 1):
        acc += float(p[k])
        if acc > u:
            return k
    return n - 1


class SweepRNG:
    canonical_adoption = True
    unbounded_caches = True

    def __init__(self, generator):
        self._g = generator
        self.seed = None
        self._t = self._i = self._c = 0
        self.log = None  # optional list of (t, i, c, kind, value) for decision-level comparisons

    # ---- sweep protocol
    def begin(self, seed):
        self.seed = int(seed) & MASK

    def end(self):
        self.seed = None

    def at(self, t, i):
        self._t, self._i, self._c = int(t), int(i), 0

    def _u(self):
        u = u01(self.seed, self._t, self._i, self._c)
        self._c += 1
        return u

    # ---- Generator interface
    @property
    def generator(self):
        return getattr(self._g, "generator", self._g)

    @property
    def bit_generator(self):
        return self._g.bit_generator

    def random(self):
        if self.seed is None:
            return self._g.random()
        return self._u()

    def multinomial(self, n, pvals):
        if self.seed is None:
            return self._g.multinomial(n, pvals)
        p = np.asarray(pvals, dtype=np.float64)
        counts = np.zeros(len(p), dtype=np.int64)
        for _ in range(int(n)):
            counts[inv_cdf(p, self._u())] += 1
        return counts

    def integers(self, low, high=None):
        if self.seed is None:
            return self._g.integers(low, high)
        if high is None:
            low, high = 0, low
        n = int(high) - int(low)
        return int(low) + min(n - 1, int(self._u() * n))

    def choice(self, a, size=None, replace=True):
        if self.seed is None:
            return self._g.choice(a, size, replace=replace)
        a = list(a)
        if size is None:
            return a[min(len(a) - 1, int(self._u() * len(a)))]
        assert not replace, "the sweep only draws without replacement"
        pos = list(range(len(a)))
        for m in range(int(size)):
            rest = len(a) - m
            j = m + min(rest - 1, int(self._u() * rest))
            pos[m], pos[j] = pos[j], pos[m]
        return np.asarray([a[k] for k in pos[: int(size)]])

    def shuffle(self, x):
        assert self.seed is None, "no shuffle inside a sweep"
        self._g.shuffle(x)
