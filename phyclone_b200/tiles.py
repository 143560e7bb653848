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
    def __init__(self, values, device=None, capacity=None, variant=-1):
        """values: [T,S,G] float64 host array with the data points' log-likelihood grids."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        T, S, G = values.shape
        self.T, self.S, self.G = T, S, G
        tile_bytes = 2 * S * (G + 48) * 8
        if capacity is None:
            capacity = int(min(65536, max(4096, (2 << 30) // tile_bytes)))
        self.max_capacity = int(max(capacity, (64 << 30) // tile_bytes))  # the arena doubles on demand up to 64 GB
        self.arena = DeviceArena(S, G, capacity + 2 * T + 1, device=device, variant=variant)
        self.device = self.arena.device
        self.prior = self.arena.prior
        # permanent tiles: V_d (data values), P_d = prior + V_d (log_p of a singleton node), PRIOR
        al = self.arena.alloc
        self.value_id = list(range(al.alloc(T), al.top))
        self.single_id = list(range(al.alloc(T), al.top))
        self.prior_id = al.alloc(1)
        self.arena.import_tiles(np.full((1, S, G), self.prior), [self.prior_id])
        self._pinned = torch.empty((T, S, G), dtype=torch.float64, pin_memory=True)
        self._staged = torch.empty((T, S, G), dtype=torch.float64, device=self.device)
        al.freeze()
        self.upload_values(values)
        self.stats = {"flushes": 0, "adds": 0, "node_jobs": 0, "scores": 0, "launches": 0, "memo_hits": 0,
                      "d2h_bytes": 0, "conv_pairs": 0, "flush_s": 0.0, "sync_s": 0.0}
        self.reset()

    def upload_values(self, values):
        """(Re)load the data points' grids: pinned host buffer -> HBM -> arena tiles V_d and P_d = prior + V_d.
        Returns the bytes copied host->device."""
        self._pinned.numpy()[...] = values
        self._staged.copy_(self._pinned, non_blocking=True)
        self.arena.import_tiles(self._staged, self.value_id)
        self.arena.tile_add([self.prior_id] * self.T, self.value_id, self.single_id, 0.0)
        self.h2d_bytes = self._pinned.numel() * 8
        return self.h2d_bytes

    # ------------------------------------------------------------------ state
    def reset(self):
        """Start a new arena generation: everything but the permanent tiles is forgotten."""
        self.arena.alloc.reset()
        self._add_memo = {}
        self._node_memo = {}
        self._score_memo = {}  # kids tuple -> (lse_sum, last_sum) floats, or pending slot index
        self._level = {}  # tile id -> dependency level inside the pending batch
        self._pend_adds = []  # (level, a, b, out, sign)
        self._pend_jobs = []  # (level, base, kids, out, slot)
        self._pend_scores = []  # kids tuples in slot order

    def _alloc(self):
        al = self.arena.alloc
        if al.top >= al.capacity:
            if al.capacity >= self.max_capacity:
                raise MemoryError("tile arena exhausted at %d tiles" % al.capacity)
            self.flush()  # pending ops hold pointers into the current tensors
            self.arena.grow(min(self.max_capacity, 2 * al.capacity))
        return al.alloc()

    def _lvl(self, ids):
        lv = self._level
        m = 0
        for i in ids:
            x = lv.get(i, 0)
            if x > m:
                m = x
        return m + 1

    # -------------------------------------------------------------------- ops
    def add(self, a, b):
        """tile(a) + tile(b)   (log_p += value)"""
        key = (a, b)
        t = self._add_memo.get(key)
        if t is not None:
            self.stats["memo_hits"] += 1
            return t
        t = self._alloc()
        lv = self._lvl(key)
        self._level[t] = lv
        self._pend_adds.append((lv, a, b, t, 1.0))
        self._add_memo[key] = t
        return t

    def sub(self, a, b):
        """tile(a) - tile(b)   (log_p -= value)"""
        key = (a, -1 - b)
        t = self._add_memo.get(key)
        if t is not None:
            self.stats["memo_hits"] += 1
            return t
        t = self._alloc()
        lv = self._lvl((a, b))
        self._level[t] = lv
        self._pend_adds.append((lv, a, b, t, -1.0))
        self._add_memo[key] = t
        return t

    def node(self, base, kids):
        """log_r = tile(base) + log_S(kids in order); kids must be non-empty."""
        key = (base, kids)
        t = self._node_memo.get(key)
        if t is not None:
            self.stats["memo_hits"] += 1
            return t
        t = self._alloc()
        lv = self._lvl(kids + (base,))
        self._level[t] = lv
        self._pend_jobs.append((lv, base, kids, t, -1))
        self._node_memo[key] = t
        return t

    def score(self, kids):
        """Handle for the data terms of a tree whose dummy root folds `kids` (non-empty):
        (sum_s LSE_g, sum_s last column) of prior + log_S(kids).  Resolve with `value()`."""
        h = self._score_memo.get(kids)
        if h is None:
            slot = len(self._pend_scores)
            self._pend_scores.append(kids)
            self._pend_jobs.append((self._lvl(kids), -1, kids, -1, slot))
            h = self._score_memo[kids] = [slot, None]
        else:
            self.stats["memo_hits"] += 1
        return h

    def value(self, handle):
        if handle[1] is None:
            self.flush()
        return handle[1]

    # ------------------------------------------------------------------ flush
    def flush(self):
        if not (self._pend_adds or self._pend_jobs):
            return
        st = self.stats
        t_begin = time.perf_counter()
        st["flushes"] += 1
        scores = self._pend_scores
        n_slots = len(scores)
        # group by dependency level; within a level the tile adds go first (jobs of the level may not depend on
        # them, ops of one level never depend on each other)
        levels = {}
        for rec in self._pend_adds:
            levels.setdefault(rec[0], ([], [], []))[0 if rec[4] > 0 else 1].append(rec)
        for rec in self._pend_jobs:
            levels.setdefault(rec[0], ([], [], []))[2].append(rec)
        packed = []
        plan = []
        ext = packed.extend
        for lv in sorted(levels):
            pos, neg, jobs = levels[lv]
            for adds, sign in ((pos, 1.0), (neg, -1.0)):
                if adds:
                    plan.append(("add", len(packed), len(adds), sign))
                    ext([r[1] for r in adds])
                    ext([r[2] for r in adds])
                    ext([r[3] for r in adds])
                    st["adds"] += len(adds)
            if jobs:
                off = len(packed)
                koff = 0
                kids_all = []
                kext = kids_all.extend
                for _, base, ks, out, slot in jobs:
                    c = len(ks)
                    ext((koff, c, base, out, (JOB_SCAN | JOB_ADD_PRIOR) if slot >= 0 else JOB_SCAN, slot, 0, 0))
                    kext(ks)
                    koff += c
                ext(kids_all)
                n = len(jobs)
                plan.append(("jobs", off, n, koff, koff - n))
                st["node_jobs"] += n
                st["conv_pairs"] += koff - n
        st["launches"] += len(plan) + (2 if n_slots else 0)
        host = self.arena.run_plan(np.asarray(packed, dtype=np.int32), plan, n_slots)
        if n_slots:
            st["d2h_bytes"] += 16 * n_slots
            st["scores"] += n_slots
            lse, last = host[0].tolist(), host[1].tolist()
            memo = self._score_memo
            for slot, ks in enumerate(scores):
                memo[ks][1] = (lse[slot], last[slot])
        self._pend_adds = []
        self._pend_jobs = []
        self._pend_scores = []
        self._level = {}
        st["flush_s"] += time.perf_counter() - t_begin

    def fetch(self, ids):
        """Dense [n,S,G] host copy of tiles (debug / tests / trace export)."""
        self.flush()
        return self.arena.export_tiles(list(ids)).cpu().numpy()
