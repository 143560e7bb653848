"""TEST-ONLY stand-in for phyclone_b200.tiles.TileStore: same interface, tiles evaluated eagerly on
the CPU with the NumPy oracle.  It lets the `-m "not gpu"` suite exercise the whole host side of the
product sampler (forest bookkeeping, proposal enumeration, RNG call order, lazy scoring protocol)
without a device.  The product never imports this module."""

import numpy as np

from oracle import numerics as onum


class FakeTileStore:
    def __init__(self, values):
        values = np.ascontiguousarray(values, dtype=np.float64)
        T, S, G = values.shape
        self.T, self.S, self.G = T, S, G
        self.prior = -float(np.log(G))
        self.tiles = {}
        self.value_id = list(range(T))
        self.single_id = list(range(T, 2 * T))
        self.prior_id = 2 * T
        for i in range(T):
            self.tiles[i] = values[i]
            self.tiles[T + i] = np.full((S, G), self.prior) + values[i]
        self.tiles[self.prior_id] = np.full((S, G), self.prior)
        self._perm = 2 * T + 1
        self.stats = {"flushes": 0, "adds": 0, "node_jobs": 0, "scores": 0, "launches": 0, "memo_hits": 0,
                      "d2h_bytes": 0, "conv_pairs": 0}
        self.reset()

    def reset(self):
        for k in [k for k in self.tiles if k >= self._perm]:
            del self.tiles[k]
        self._top = self._perm
        self._memo = {}
        self._score = {}

    def _new(self, key, arr):
        t = self._top
        self._top += 1
        self.tiles[t] = arr
        self._memo[key] = t
        return t

    def add(self, a, b):
        key = ("a", a, b)
        t = self._memo.get(key)
        return t if t is not None else self._new(key, self.tiles[a] + self.tiles[b])

    def sub(self, a, b):
        key = ("s", a, b)
        t = self._memo.get(key)
        return t if t is not None else self._new(key, self.tiles[a] - self.tiles[b])

    def node(self, base, kids):
        key = ("n", base, kids)
        t = self._memo.get(key)
        if t is None:
            self.stats["node_jobs"] += 1
            t = self._new(key, self.tiles[base] + onum.log_s([self.tiles[k] for k in kids]))
        return t

    def score(self, kids):
        h = self._score.get(kids)
        if h is None:
            tile = self.prior + onum.log_s([self.tiles[k] for k in kids])
            h = self._score[kids] = [0, onum.root_terms(tile)]
            self.stats["scores"] += 1
        return h

    def value(self, handle):
        return handle[1]

    def flush(self):
        self.stats["flushes"] += 1

    def fetch(self, ids):
        return np.array([self.tiles[i] for i in ids])
