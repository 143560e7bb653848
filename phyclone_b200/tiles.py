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

import numpy as np
import torch

from .engine import JOB_ADD_PRIOR, JOB_SCAN, DeviceArena


class TileStore:
    def __init__(self, values, device=None, capacity=None, variant=-1):
        """values: [T,S,G] float64 host array with the data points' log-likelihood grids."""
        values = np.ascontiguousarray(values, dtype=np.float64)
        T, S, G = values.shape
        self.T, self.S, self.G = T, S, G
        if capacity is None:
            tile_bytes = 2 * S * (G + 48) * 8
            capacity = int(min(max(16384, 4096 * 50), max(4096, (8 << 30) // tile_bytes)))
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
                      "d2h_bytes": 0, "conv_pairs": 0}
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
        t = self.arena.alloc.alloc()
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
        t = self.arena.alloc.alloc()
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
        t = self.arena.alloc.alloc()
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
        ar = self.arena
        st = self.stats
        st["flushes"] += 1
        n_slots = len(self._pend_scores)
        max_level = 0
        for rec in self._pend_adds:
            if rec[0] > max_level:
                max_level = rec[0]
        for rec in self._pend_jobs:
            if rec[0] > max_level:
                max_level = rec[0]
        adds_by = [[] for _ in range(max_level + 1)]
        jobs_by = [[] for _ in range(max_level + 1)]
        for rec in self._pend_adds:
            adds_by[rec[0]].append(rec)
        for rec in self._pend_jobs:
            jobs_by[rec[0]].append(rec)
        rows = None
        if n_slots:
            rows = torch.empty((2, n_slots, self.S), dtype=torch.float64, device=self.device)
        for lv in range(1, max_level + 1):
            adds = adds_by[lv]
            if adds:
                pos = [r for r in adds if r[4] > 0]
                neg = [r for r in adds if r[4] < 0]
                if pos:
                    ar.tile_add([r[1] for r in pos], [r[2] for r in pos], [r[3] for r in pos], 0.0)
                    st["launches"] += 1
                if neg:
                    ar.tile_sub([r[1] for r in neg], [r[2] for r in neg], [r[3] for r in neg])
                    st["launches"] += 1
                st["adds"] += len(adds)
            jobs = jobs_by[lv]
            if jobs:
                n = len(jobs)
                tab = np.zeros((n, 8), dtype=np.int32)
                kids = []
                off = 0
                for i, (_, base, ks, out, slot) in enumerate(jobs):
                    c = len(ks)
                    tab[i, 0] = off
                    tab[i, 1] = c
                    tab[i, 2] = base
                    tab[i, 3] = out
                    tab[i, 4] = JOB_SCAN | (JOB_ADD_PRIOR if slot >= 0 else 0)
                    tab[i, 5] = slot
                    kids.extend(ks)
                    off += c
                    st["conv_pairs"] += c - 1
                ar.launch_jobs(tab, np.asarray(kids, dtype=np.int32), rows)
                st["launches"] += 1
                st["node_jobs"] += n
        if n_slots:
            sums = ar.reduce_rows(rows)  # [2, n_slots] on device
            st["launches"] += 1
            host = sums.cpu().numpy()  # the one synchronising D2H copy of the step
            st["d2h_bytes"] += host.nbytes
            st["scores"] += n_slots
            lse, last = host[0].tolist(), host[1].tolist()
            memo = self._score_memo
            for slot, ks in enumerate(self._pend_scores):
                memo[ks][1] = (lse[slot], last[slot])
        self._pend_adds = []
        self._pend_jobs = []
        self._pend_scores = []
        self._level = {}

    def fetch(self, ids):
        """Dense [n,S,G] host copy of tiles (debug / tests / trace export)."""
        self.flush()
        return self.arena.export_tiles(list(ids)).cpu().numpy()
