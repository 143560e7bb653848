"""TEST/PROFILING-ONLY executor: no tile arithmetic at all, scores are a cheap deterministic function of the slot.
Lets the host bookkeeping of the product sampler (Python + the _pcbhost extension) be profiled on a CPU box."""

import numpy as np

from phyclone_b200.engine import TileAllocator
from phyclone_b200.tiles import TileStore


class NullArena:
    def __init__(self, S, G, capacity):
        self.S, self.G = S, G
        self.capacity = int(capacity)
        self.alloc = TileAllocator(self.capacity)
        self.prior = -float(np.log(G))
        self.staged_bytes = 0
        self._n = 0

    def grow(self, new_capacity):
        self.capacity = self.alloc.capacity = int(new_capacity)

    def import_tiles(self, dense, ids):
        pass

    def upload_dense(self, values, ids):
        return np.asarray(values).nbytes

    def tile_add(self, *a, **k):
        pass

    def run_plan(self, packed, plan, n_slots, plan_raw=None):
        if not n_slots:
            return None
        self._n += 1
        base = -200.0 - 3.0 * ((np.arange(n_slots) * 7 + self._n) % 11)
        return np.ascontiguousarray(np.stack([base, base - 3.0]))


def NullTileStore(values):
    return TileStore(values, arena_factory=NullArena)
