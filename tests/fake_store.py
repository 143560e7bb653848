"""TEST-ONLY executor for phyclone_b200.tiles.TileStore: the same interface as engine.DeviceArena, tiles
evaluated on the CPU with the NumPy oracle.  It lets the `-m "not gpu"` suite exercise the whole host side of
the product sampler -- forest bookkeeping, proposal enumeration, RNG call order, the lazy scoring protocol AND
the C++ tile-store bookkeeping / flush packing (csrc/pcb_hostmod.cpp) -- without a device.
The product never imports this module."""

import numpy as np

from oracle import numerics as onum
from phyclone_b200.engine import JOB_ADD_PRIOR, JOB_SCAN, TileAllocator
from phyclone_b200.tiles import TileStore


class _ForgetfulAllocator(TileAllocator):
    """A new arena generation really forgets the old tiles here (the device arena merely overwrites them later): a
    tree that reads tile ids of a previous generation fails loudly in the CPU tests instead of working by luck."""

    def __init__(self, capacity, arena):
        super().__init__(capacity)
        self._arena = arena

    def reset(self):
        super().reset()
        for i in [i for i in self._arena.tiles if i >= self.permanent]:
            del self._arena.tiles[i]


class FakeArena:
    def __init__(self, S, G, capacity):
        self.S, self.G = S, G
        self.capacity = int(capacity)
        self.tiles = {}
        self.alloc = _ForgetfulAllocator(self.capacity, self)
        self.prior = -float(np.log(G))
        self.staged_bytes = 0

    def grow(self, new_capacity):
        self.capacity = self.alloc.capacity = int(new_capacity)

    def import_tiles(self, dense, ids):
        for t, i in zip(np.asarray(dense, dtype=np.float64), ids):
            self.tiles[int(i)] = t.copy()

    def upload_dense(self, values, ids):
        self.import_tiles(values, ids)
        return np.asarray(values).nbytes

    def export_tiles(self, ids):
        return np.array([self.tiles[int(i)] for i in ids])

    def tile_add(self, a_ids, b_ids, out_ids, addend=0.0, b_sign=1.0):
        for k, o in enumerate(out_ids):
            v = self.tiles[int(a_ids[k])]
            if b_ids is not None and b_ids[k] >= 0:
                v = v + b_sign * self.tiles[int(b_ids[k])]
            self.tiles[int(o)] = v + addend if addend != 0.0 else v + 0.0

    def map_trees(self, lp_ids, tree_off, post_order, child_off, child_idx):
        from oracle import map as omap

        lp = self.export_tiles(lp_ids)
        out = np.zeros((len(lp_ids), self.S), dtype=np.int32)
        kids = [list(child_idx[child_off[v]:child_off[v + 1]]) for v in range(len(lp_ids))]
        for t in range(len(tree_off) - 1):
            order = list(post_order[tree_off[t]:tree_off[t + 1]])
            idx, _ = omap.map_tree(lp, kids, order)
            out[order] = idx[order]
        return out

    def run_plan(self, packed, plan, n_slots, plan_raw=None):
        out = np.zeros((2, max(n_slots, 1)))
        for rec in plan:
            if rec[0] == "add":
                _, off, n, sign = rec
                ids = packed[off:off + 3 * n]
                self.tile_add(ids[:n], ids[n:2 * n], ids[2 * n:], 0.0, sign)
            else:
                _, off, n, n_kids, _ = rec
                tab = packed[off:off + 8 * n].reshape(n, 8)
                kids = packed[off + 8 * n:off + 8 * n + n_kids]
                for koff, c, base, outt, flags, slot, _, _ in tab:
                    assert flags & JOB_SCAN
                    tile = onum.log_s([self.tiles[int(k)] for k in kids[koff:koff + c]])
                    if base >= 0:
                        tile = self.tiles[int(base)] + tile
                    if flags & JOB_ADD_PRIOR:
                        tile = self.prior + tile
                    if outt >= 0:
                        self.tiles[int(outt)] = tile
                    if slot >= 0:
                        out[0, slot], out[1, slot] = onum.root_terms(tile)
        return np.ascontiguousarray(out[:, :n_slots]) if n_slots else None


def FakeTileStore(values, capacity=None):
    """The product's TileStore (C++ bookkeeping included) on the NumPy executor."""
    return TileStore(values, capacity=capacity, arena_factory=FakeArena)
