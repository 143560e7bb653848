"""Decision tape: record every RNG call of one sampler run and replay it into another.

Why: the sampler's discrete decisions go through numpy's Generator.multinomial, whose outcome (and
even the number of uniforms it consumes) can flip on a ONE-ulp difference in a weight whenever two
tail particles tie (p / remaining_p = 0.5 +- eps picks a different binomial branch).  GPU and CPU
log-likelihoods agree to ~1e-13, not to the bit, so a free-running GPU chain need not follow the CPU
chain forever even though both are correct.  Replaying the oracle's decisions into the GPU sampler,
while checking that every probability vector it would have drawn from matches the oracle's to 1e-9,
proves the stronger per-step statement: same candidate enumeration, same call order, same
probabilities -- hence the same decision whenever the draw is not on a knife edge.
"""

import numpy as np


def _snap(x):
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, (list, tuple)):
        return list(x)
    return x


class RecordingRNG:
    def __init__(self, generator):
        self._g = generator
        self.tape = []

    @property
    def generator(self):
        self.tape.append(("generator", self._g.bit_generator.state))
        return self._g

    def random(self):
        r = self._g.random()
        self.tape.append(("random", None, r))
        return r

    def multinomial(self, n, pvals):
        r = self._g.multinomial(n, pvals)
        self.tape.append(("multinomial", (n, np.array(pvals, dtype=np.float64)), r.copy()))
        return r

    def integers(self, low, high=None):
        r = self._g.integers(low, high)
        self.tape.append(("integers", (low, high), r))
        return r

    def choice(self, a, size=None, replace=True):
        r = self._g.choice(a, size, replace=replace)
        self.tape.append(("choice", (_snap(np.asarray(a)), size, replace), _snap(r)))
        return r

    def shuffle(self, x):
        before = list(x)
        self._g.shuffle(x)
        self.tape.append(("shuffle", before, list(x)))


class ReplayRNG:
    """Returns the recorded results; asserts the call sequence and that real-valued arguments match to rtol."""

    def __init__(self, tape, rtol=1e-9):
        self.tape = tape
        self.pos = 0
        self.rtol = rtol
        self.max_p_err = 0.0

    def _next(self, kind):
        assert self.pos < len(self.tape), "replay ran past the end of the tape (%s)" % kind
        rec = self.tape[self.pos]
        assert rec[0] == kind, "call %d: expected %s, sampler asked for %s" % (self.pos, rec[0], kind)
        self.pos += 1
        return rec

    def done(self):
        return self.pos == len(self.tape)

    @property
    def generator(self):
        rec = self._next("generator")
        g = np.random.default_rng()
        g.bit_generator.state = rec[1]
        return g

    def random(self):
        return self._next("random")[2]

    def multinomial(self, n, pvals):
        rec = self._next("multinomial")
        n0, p0 = rec[1]
        p = np.asarray(pvals, dtype=np.float64)
        assert n == n0 and p.shape == p0.shape, "call %d: multinomial shape differs" % (self.pos - 1)
        err = np.abs(p - p0) / np.maximum(p0, 1e-300)
        # tolerance on probabilities that matter; entries below 1e-30 are checked absolutely
        bad = (err > self.rtol) & (np.abs(p - p0) > 1e-30)
        assert not bad.any(), "call %d: multinomial probabilities differ (max rel %.3e)" % (self.pos - 1, err[bad].max())
        self.max_p_err = max(self.max_p_err, float(err[np.abs(p - p0) > 1e-30].max()) if (np.abs(p - p0) > 1e-30).any() else 0.0)
        return rec[2].copy()

    def integers(self, low, high=None):
        rec = self._next("integers")
        assert (low, high) == rec[1], "call %d: integers range differs" % (self.pos - 1)
        return rec[2]

    def choice(self, a, size=None, replace=True):
        rec = self._next("choice")
        assert np.array_equal(np.asarray(a), np.asarray(rec[1][0])) and size == rec[1][1], "call %d: choice differs" % (self.pos - 1)
        return _snap(rec[2])

    def shuffle(self, x):
        rec = self._next("shuffle")
        assert list(x) == rec[1], "call %d: shuffle input differs" % (self.pos - 1)
        x[:] = rec[2]
