"""TEST/PROFILING-ONLY tile store: no arithmetic at all, scores are a cheap deterministic function of the
operand ids.  Lets the host bookkeeping of the product sampler be profiled on a CPU box."""

import numpy as np


class NullTileStore:
    def __init__(self, values):
        T, S, G = values.shape
        self.T, self.S, self.G = T, S, G
        self.prior = -float(np.log(G))
        self.value_id = list(range(T))
        self.single_id = list(range(T, 2 * T))
        self.prior_id = 2 * T
        self._perm = 2 * T + 1
        self.base = [float(-200.0 - 3.0 * (i % 7)) for i in range(T)]
        self.stats = {"flushes": 0, "adds": 0, "node_jobs": 0, "scores": 0, "launches": 0, "memo_hits": 0,
                      "d2h_bytes": 0, "conv_pairs": 0}
        self.reset()

    def reset(self):
        self._top = self._perm
        self._memo = {}
        self._score = {}
        self._w = {}

    def _weight(self, t):
        if t < self.T:
            return self.base[t]
        if t < 2 * self.T:
            return self.base[t - self.T]
        return self._w.get(t, 0.0)

    def _new(self, key, w):
        t = self._top
        self._top += 1
        self._memo[key] = t
        self._w[t] = w
        return t

    def add(self, a, b):
        key = ("a", a, b)
        t = self._memo.get(key)
        return t if t is not None else self._new(key, self._weight(a) + self._weight(b))

    def sub(self, a, b):
        key = ("s", a, b)
        t = self._memo.get(key)
        return t if t is not None else self._new(key, self._weight(a) - self._weight(b))

    def node(self, base, kids):
        key = ("n", base, kids)
        t = self._memo.get(key)
        if t is None:
            self.stats["node_jobs"] += 1
            t = self._new(key, self._weight(base) + sum(self._weight(k) for k in kids) - 2.0 * len(kids))
        return t

    def score(self, kids):
        h = self._score.get(kids)
        if h is None:
            w = sum(self._weight(k) for k in kids) - 1.5 * len(kids)
            h = self._score[kids] = [0, (w, w - 3.0)]
            self.stats["scores"] += 1
        return h

    def value(self, handle):
        return handle[1]

    def flush(self):
        self.stats["flushes"] += 1
