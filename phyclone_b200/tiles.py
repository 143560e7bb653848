"""Lazy, memoised tile algebra on top of the device arena.

Host trees never hold arrays: a node's `log_p` / `log_r` is an integer tile id.  Every tile
operation the reference performs eagerly on NumPy arrays --

    log_p += value ; log_r += value                (tree/tree_node.py:42-52)
    log_p -= value                                 (tree/tree_node.py:54-59)
    log_r  = log_p + compute_log_S(children)       (tree/tree_node.py:61-71)
    sum_s LSE_g root.log_r , sum_s root.log_r[:,-1] (tree/distributions.py:197-200)

-- is recorded here as a pending op and gets a tile id at once; `flush()` then runs all pending
ops of a whole SMC / Gibbs step as a few batched launches (one tile_add launch and one
node-jobs launch per dependency level).  Ops are memoised on (operation, ordered operand ids):
the same node refresh requested again -- resampled duplicates, the full `update()` after every
`from_dict`, the unchanged siblings of a Gibbs move -- costs a dict lookup instead of a kernel.
This replaces the reference's xxh3 content-hash LRUs (utils/utils.py:70-149); unlike those,
the key keeps the child ORDER, so the floored fold is always evaluated in the order asked for
(SURVEY.md App. A.9).
"""

import time

import numpy as np
import torch

from .engine import JOB_ADD_PRIOR, JOB_SCAN, DeviceArena


class TileStore:
    """Python face of the tile store.  The bookkeeping (memo tables, pending ops, dependency levels, flush
    packing) lives in the `_pcbhost.TileBook` C++ extension (csrc/pcb_hostmod.cpp); the arithmetic is executed
    by `arena` (an `engine.DeviceArena`; the CPU tests plug in a NumPy executor with the same interface)."""

    LATENCY_LAYOUT = -2  # narrow strips: shorter single launches (27 vs 33 us at config #2); what a chain uses
    THROUGHPUT_LAYOUT = -1  # wide strips: best FP64 pipe use on saturated batches

    def __init__(self, values, device=None, capacity=None, variant=-1, arena_factory=None):
        """values: [T,S,G] float64 host array with the data points' log-likelihood grids."""
        from ._hostmod import load as load_hostmod

        values = np.ascontiguousarray(values, dtype=np.float64)
        T, S, G = values.shape
        self.T, self.S, self.G = T, S, G
        tile_bytes = 2 * S * (G + 48) * 8
        if capacity is None:
            capacity = int(min(65536, max(4096, (2 << 30) // tile_bytes)))
        self.max_capacity = int(max(capacity, (64 << 30) // tile_bytes))  # the arena doubles on demand up to 64 GB
        if arena_factory is None:
            self.arena = DeviceArena(S, G, capacity + 2 * T + 1, device=device, variant=variant)
        else:
            self.arena = arena_factory(S, G, capacity + 2 * T + 1)
        self.device = getattr(self.arena, "device", None)
        self.prior = self.arena.prior
        # permanent tiles: V_d (data values), P_d = prior + V_d (log_p of a singleton node), PRIOR
        al = self.arena.alloc
        self.value_id = list(range(al.alloc(T), al.top))
        self.single_id = list(range(al.alloc(T), al.top))
        self.prior_id = al.alloc(1)
        self.arena.import_tiles(np.full((1, S, G), self.prior), [self.prior_id])
        al.freeze()
        self._values_host = None
        self.upload_values(values)
        self.book = load_hostmod().TileBook(al.permanent, self.arena.capacity, self.prior_id)
        self._stats = {"flushes": 0, "launches": 0, "d2h_bytes": 0, "flush_s": 0.0, "device_wait_s": 0.0}
        self._ctx = None
        self.generation = 0  # bumped by reset(): trees remember which generation their tile ids belong to
        self.reset()

    @property
    def stats(self):
        d = dict(self._stats)
        c = self.book.counters()
        d["flushes"] += c.pop("fast_flushes")
        d["launches"] += c.pop("fast_launches")
        d["d2h_bytes"] += c.pop("fast_d2h_bytes")
        run_s = c.pop("fast_run_s")
        d["flush_s"] += run_s
        d["device_wait_s"] += run_s
        d["staged_ints"] = c.pop("fast_staged_ints")
        d.update(c)
        return d

    def upload_values(self, values):
        """(Re)load the data points' grids: host -> HBM -> arena tiles V_d and P_d = prior + V_d.
        Returns the bytes copied host->device."""
        self.h2d_bytes = self.arena.upload_dense(values, self.value_id)
        self.arena.tile_add([self.prior_id] * self.T, self.value_id, self.single_id, 0.0)
        return self.h2d_bytes

    # ------------------------------------------------------------------ state
    def reset(self):
        """Start a new arena generation: everything but the permanent tiles is forgotten."""
        self.arena.alloc.reset()
        self.book.reset()
        self.generation += 1

    def rebase(self, tree):
        """Start a new arena generation that still holds `tree`'s tiles: they are copied to the bottom of the new
        generation and the tree's ids are rewritten in place.  A plain `reset()` leaves a carried tree's ids pointing
        at slots the new generation will overwrite, which is fine as long as the tree is only read structurally (the
        particle-Gibbs sweep, the Gibbs moves on the tree it returns) -- the subtree move prunes and re-grafts the
        carried tree itself, i.e. reads its tiles."""
        al = self.arena.alloc
        ids = sorted({i for i in list(tree._lp.values()) + list(tree._lr.values()) if i >= al.permanent})
        dense = self.arena.export_tiles(ids) if ids else None  # flushes nothing: a carried tree has no pending ops
        self.reset()
        tree._gen = self.generation
        if not ids:
            return
        base = self.book.reserve(len(ids))
        while base < 0:
            self._grow()
            base = self.book.reserve(len(ids))
        new_ids = list(range(base, base + len(ids)))
        self.arena.import_tiles(dense, new_ids)
        remap = dict(zip(ids, new_ids))
        tree._lp = {k: remap.get(v, v) for k, v in tree._lp.items()}
        tree._lr = {k: remap.get(v, v) for k, v in tree._lr.items()}
        tree._gen = self.generation

    def _grow(self):
        al = self.arena.alloc
        if al.capacity >= self.max_capacity:
            raise MemoryError("tile arena exhausted at %d tiles" % al.capacity)
        self.flush()  # pending ops hold pointers into the current tensors
        self.arena.grow(min(self.max_capacity, 2 * al.capacity))
        self.book.set_capacity(self.arena.capacity)
        self._ctx = None  # the arena descriptor points at new tensors

    # -------------------------------------------------------------------- ops
    def add(self, a, b):
        """tile(a) + tile(b)   (log_p += value)"""
        t = self.book.add(a, b)
        if t < 0:
            self._grow()
            t = self.book.add(a, b)
        return t

    def sub(self, a, b):
        """tile(a) - tile(b)   (log_p -= value)"""
        t = self.book.sub(a, b)
        if t < 0:
            self._grow()
            t = self.book.sub(a, b)
        return t

    def node(self, base, kids):
        """log_r = tile(base) + log_S(kids in order); kids must be non-empty."""
        t = self.book.node(base, kids)
        if t < 0:
            self._grow()
            t = self.book.node(base, kids)
        return t

    def score(self, kids):
        """Handle for the data terms of a tree whose dummy root folds `kids` (non-empty):
        (sum_s LSE_g, sum_s last column) of prior + log_S(kids).  Resolve with `value()`."""
        return self.book.score(kids)

    def value(self, handle):
        v = self.book.value(handle)
        if v is None:
            self.flush()
            v = self.book.value(handle)
        return v

    # ------------------------------------------------------------------ flush
    def flush(self):
        book = self.book
        if not book.has_pending():
            return
        ctx = self._ctx
        if ctx is not None and book.run(*ctx) is None:
            return  # packed, launched and read back inside the extension (one call into libphyclone_b200)
        self._flush_generic()
        make_ctx = getattr(self.arena, "flush_context", None)
        self._ctx = make_ctx() if make_ctx is not None else None  # buffers may have been (re)allocated

    def _flush_generic(self):
        """The same flush step by step from Python: first flush, buffer growth, batches without scores, and the
        executors of the CPU tests."""
        book = self.book
        st = self._stats
        t_begin = time.perf_counter()
        st["flushes"] += 1
        packed, plan, n_slots, plan_raw = book.pack()
        st["launches"] += len(plan) + (1 if n_slots else 0)
        t_run = time.perf_counter()
        host = self.arena.run_plan(np.frombuffer(packed, dtype=np.int32), plan, n_slots, plan_raw)
        st["device_wait_s"] += time.perf_counter() - t_run
        if n_slots:
            st["d2h_bytes"] += 16 * n_slots
        book.done(host)
        st["flush_s"] += time.perf_counter() - t_begin

    def fetch(self, ids):
        """Dense [n,S,G] host copy of tiles (debug / tests / trace export)."""
        self.flush()
        out = self.arena.export_tiles(list(ids))
        return out.cpu().numpy() if hasattr(out, "cpu") else out
