"""Device-resident SMC sweep (host side): hand the retained path to `pcb_smc_sweep_dev`, read the genealogy back, rebuild
the one tree that was picked.

This is the sampler of the reference (smc/samplers/conditional.py:76-125, standard.py:20-45,
smc/kernels/semi_adapted.py:36-194, mcmc/particle_gibbs.py:42-48) with its NumPy `Generator` draws inside the sweep
replaced by inverse-CDF transforms of keyed uniforms, so that proposal normalisation, draws, particle construction,
weights / ESS and resampling of all T steps run on the device with ONE synchronisation at the end
(csrc/pcb_sweep.cuh; the decision-for-decision CPU definition of the mode is oracle/uniform.py, a test-only module).
What stays on the host: the sigma permutation, the retained path (T trees, scored by one batched flush), the sweep
seed (one `integers` draw from the chain's Generator) and replaying the picked particle's T edits into a `Tree`.
"""

import ctypes as C
import ctypes.util

import numpy as np

from .. import _lib
from ..tree.tree import Tree
from .._hostmod import load as _load_hostmod
from .forest_state import ATTACH, NEW, ForestState, host_tuple

_hm = _load_hostmod()

IHEAD, NFIELDS, DREC = 5, 9, 8
FLAG_RETAINED_PARENT = 8
MAX_ROOTS = 64
STATUS = {1: "more than Rmax roots in a forest", 2: "tile range exhausted", 4: "child-list pool exhausted",
          8: "non-finite proposal or particle weights"}

_u64 = C.c_uint64


class SweepDesc(C.Structure):
    """include/phyclone_b200.h: pcb_sweep_desc"""

    _fields_ = (
        [(n, C.c_int32) for n in ("N", "T", "Rmax", "conditional", "outliers", "n_tab", "n_ret_kids", "perm_dist")]
        + [(n, C.c_double) for n in ("log_alpha", "threshold", "log_half", "log_uniform", "lw_uniform", "prior")]
        + [("seed", _u64), ("tile_base", C.c_int64), ("tile_limit", C.c_int64)]
        + [(n, C.c_void_p) for n in (
            "dp_idx", "value_id", "single_id", "dp_op", "dp_opn", "dp_marg", "lfact", "logn", "rterm", "ret_th_int",
            "ret_ba_int", "ret_th_dbl", "ret_ba_dbl", "ret_key", "ret_kids", "ret_kind", "ret_arg", "ret_subtree_data",
            "ret_subtree_log_count", "anc", "kind", "name", "mask", "raw", "logp",
            "logq", "ess", "resampled", "final_idx", "status", "tiles_used", "n_launches", "conv_pairs", "device_ms")]
    )


_libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.lgamma.restype = C.c_double
_libm.lgamma.argtypes = [C.c_double]
_tables = {}


def _prior_tables(prior, n):
    """log n!, log n and the FS-CRP root-count term as the host sampler computes them (forest_state.ensure_tables,
    csrc/pcb_hostmod.cpp lf/tlog): the device adds the same doubles."""
    key = (prior._c_const, n)
    t = _tables.get(key)
    if t is None:
        lfact = np.array([_libm.lgamma(float(k) + 1.0) for k in range(n)], dtype=np.float64)
        logn = np.array([0.0] + [float(np.log(k)) for k in range(1, n)], dtype=np.float64)
        rterm = np.array([float(prior._compute_r_term(k)) for k in range(n)], dtype=np.float64)
        t = _tables[key] = (lfact, logn, rterm)
    return t


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RetainedPath(object):
    """The retained path of a conditional sweep (smc/samplers/conditional.py:19-74) as root-level states, without
    building the T trees: per data point the edit the current tree dictates, the as-built state (the tree the
    reference grows by editing in place: what the path's own proposals are scored on) and the thawed state (that tree
    re-read through `Tree.from_dict`: what extends a particle), both kept by the host extension
    (`ForestState.extend_built` / `extend`), plus the as-built tree's score.  A `Tree` is replayed only where the
    picked lineage touches the path (`tree(t)`)."""

    def __init__(self, tree, sigma, kernel):
        from .forest_state import OUTLIER, ensure_tables

        store = kernel.store
        book = store.book
        prior = kernel.tree_dist.prior
        ensure_tables(prior, store.T + 2)
        log_alpha = float(prior.log_alpha)
        labels = tree.labels
        outlier = tree.outlier_node_name
        node_map = {}
        th = ba = ForestState.empty(store)
        self.store, self.sigma, self.grid_size = store, sigma, sigma[0].grid_size
        self.thawed, self.built, self.edits, self.args, self._scores = [], [], [], [], []
        for dp in sigma:
            t = host_tuple(store, dp)
            old = labels[dp.idx]
            if old == outlier:
                kind, arg, edit = OUTLIER, None, (2, 0)
            elif old in node_map:
                kind, arg = ATTACH, node_map[old]
                edit = (0, int(arg))
            else:
                arg = tuple(node_map[c] for c in tree.get_children(old))  # adoption order (conditional.py:55-58)
                kind, edit = NEW, (1, len(arg))
                node_map[old] = th.n_nodes
            th = self._step(th.extend, kind, t, arg)
            ba = self._step(ba.extend_built, kind, t, arg)
            self.thawed.append(th)
            self.built.append(ba)
            self.edits.append(edit)
            self.args.append((kind, arg))
            self._scores.append(ba.score_self(book, log_alpha))
        self._replayed = (0, None)  # (data points replayed, thawed tree)

    def _step(self, fn, kind, t, arg):
        book = self.store.book
        s = fn(book, kind, t, arg)
        if s is None:  # arena full: grow it and record again (ops already recorded are memoised)
            self.store._grow()
            s = fn(book, kind, t, arg)
        return s

    def scores(self):
        """[(log_p, log_p_one)] of the T retained trees; the first read flushes the store (one batched evaluation)."""
        out = []
        for pp, p1, sid in self._scores:
            if sid >= 0:
                lse_sum, last_sum = self.store.value(sid)
                pp, p1 = pp + lse_sum, p1 + last_sum
            out.append((pp, p1))
        return out

    def tree(self, t):
        """The retained particle after `t` data points as the reference's `particle.tree` (a from_dict round trip of
        the as-built tree): the path's edits replayed on a `Tree`, re-read after every edit."""
        done, tree = self._replayed
        if tree is None or done > t:
            done, tree = 0, Tree(self.grid_size, self.store)
        for k in range(done, t):
            dp = self.sigma[k]
            kind, arg = self.args[k]
            # edited in place: `tree` is fresh, or the previous step's thawed copy, or the cached replay that the
            # assignment below supersedes
            if kind == ATTACH:
                tree.add_data_point_to_node(dp, arg)
            elif kind == NEW:
                node = tree.create_root_node(arg)
                tree.add_data_point_to_node(dp, node)
            else:
                tree.add_data_point_to_outliers(dp)
            tree.outliers  # noqa: B018 -- same key-creation side effect as scoring the as-built tree
            tree = tree.thawed_copy()
        self._replayed = (t, tree)
        return tree.copy()


class SweepFallback(Exception):
    """The sweep cannot run on the device (a forest with more than 64 roots, a tree with graph vacancies ...):
    the caller runs the host sampler instead."""


class DeviceSweep(object):
    """One device-resident sweep per call; buffers are reused across calls."""

    def __init__(self, kernel, num_particles, resample_threshold):
        if not getattr(kernel, "use_states", False):
            raise ValueError("the device-resident sweep implements the semi-adapted proposal only")
        if not 0 <= resample_threshold < 1:
            raise ValueError("the device-resident sweep needs a resampling threshold in [0, 1)")
        if not 2 <= num_particles <= 1024:
            raise ValueError("the device-resident sweep handles 2..1024 particles")
        self.kernel = kernel
        self.store = kernel.store
        self.N = int(num_particles)
        self.threshold = float(resample_threshold)
        self.perm = kernel.perm_dist is not None  # weights carry the root-permutation density (kernels/base.py:49-55)
        self.lib = _lib.require_device()
        self.stats = {"sweeps": 0, "fallbacks": 0, "launches": 0, "device_ms": 0.0, "conv_pairs": 0, "tiles": 0,
                      "h2d_bytes": 0, "d2h_bytes": 0}
        self._out = None
        self.last = None  # outputs of the last sweep (tests / diagnostics)
        self.history = None  # set to a list to keep a copy of every sweep's outputs

    # ------------------------------------------------------------------------------------------------ packing
    def _pack_retained(self, path, rmax):
        """Retained-path tables of pcb_sweep_desc: per step the thawed state (what extends the particle), the as-built
        state (what its proposal's candidates are scored on), log_p / log_p_one and the structure hash."""
        T = len(path) - 1
        th, ba = [], []
        for t in range(T):
            holder = path[t + 1]._holder
            s_th = holder.thawed_state()
            s_ba = ForestState.from_tree(holder._built)
            if s_th is None or s_ba is None:
                raise SweepFallback("retained tree is not an SMC-grown tree")
            th.append(s_th)
            ba.append(s_ba)
        try:
            th_i, ba_i, th_d, ba_d, kids = _hm.pack_states(th, ba, rmax)
        except OverflowError as e:
            raise SweepFallback(str(e))
        th_d = np.frombuffer(th_d, dtype=np.float64).reshape(T, DREC).copy()
        for t in range(T):
            holder = path[t + 1]._holder
            th_d[t, 4], th_d[t, 5] = holder.log_p, holder.log_p_one  # the first read scores the whole path: one flush
        # "equal particle" exactly as the host sampler tests it (kernels.get_proposal_distribution)
        keys = np.array([path[t + 1]._holder.vkey[0] for t in range(T)], dtype=np.uint64)
        sub_n = sub_c = None
        if self.perm:
            # root-permutation density (smc/utils.py:18-66): the retained trees' log_pdf and, per root of the thawed
            # state, the data points below it and the log number of compatible orders of its subtree
            from .utils import RootPermutationDistribution as RPD

            sub_n = np.zeros((T, rmax), dtype=np.int32)
            sub_c = np.zeros((T, rmax), dtype=np.float64)
            for t in range(T):
                holder = path[t + 1]._holder
                th_d[t, 6] = RPD.log_pdf(holder._built)
                tree = holder._thawed
                for r, node in enumerate(th[t].roots):
                    sub_c[t, r], sub_n[t, r] = RPD.subtree_terms(tree, node)
        return (np.frombuffer(th_i, dtype=np.int32), np.frombuffer(ba_i, dtype=np.int32), th_d,
                np.frombuffer(ba_d, dtype=np.float64), keys, np.frombuffer(kids, dtype=np.int32), sub_n, sub_c)

    def _pack_states(self, rp, rmax):
        """The same tables from a `RetainedPath` (no trees involved)."""
        T = len(rp.thawed)
        try:
            th_i, ba_i, th_d, ba_d, kids = _hm.pack_states(rp.thawed, rp.built, rmax)
        except OverflowError as e:
            raise SweepFallback(str(e))
        th_d = np.frombuffer(th_d, dtype=np.float64).reshape(T, DREC).copy()
        th_d[:, 4:6] = rp.scores()
        keys = np.array([s.skey for s in rp.thawed], dtype=np.uint64)
        return (np.frombuffer(th_i, dtype=np.int32), np.frombuffer(ba_i, dtype=np.int32), th_d,
                np.frombuffer(ba_d, dtype=np.float64), keys, np.frombuffer(kids, dtype=np.int32), None, None)

    def _outputs(self, T):
        N = self.N
        o = self._out
        if o is None or o["anc"].shape[0] < T:
            o = self._out = {
                "anc": np.zeros((T, N), dtype=np.int32), "kind": np.zeros((T, N), dtype=np.int32),
                "name": np.zeros((T, N), dtype=np.int32), "mask": np.zeros((T, N), dtype=np.uint64),
                "raw": np.zeros((T, N)), "logp": np.zeros((T, N)), "logq": np.zeros((T, N)), "ess": np.zeros(T),
                "resampled": np.zeros(T, dtype=np.int32), "final": np.zeros(1, dtype=np.int32),
                "status": np.zeros(1, dtype=np.int32), "tiles": np.zeros(1, dtype=np.int64),
                "launches": np.zeros(1, dtype=np.int64), "pairs": np.zeros(1, dtype=np.int64), "ms": np.zeros(1)}
        return o

    # ------------------------------------------------------------------------------------------------ the sweep
    def run(self, sigma, seed, path=None, path_edits=None):
        """Sweep over the data points `sigma`; `path` = the retained path ([None, particle_1 .. particle_T]) of a
        conditional sweep and `path_edits` its (kind, argument) per data point, None for an unconditional sweep.
        Returns the picked particle's tree."""
        store, N = self.store, self.N
        T = len(sigma)
        rmax = min(MAX_ROOTS, T + 1)
        prior = self.kernel.tree_dist.prior
        n_tab = max(T + 3, rmax + 3, 8)
        lfact, logn, rterm = _prior_tables(prior, n_tab)
        conditional = path is not None
        if isinstance(path, RetainedPath):
            th_i, ba_i, th_d, ba_d, keys, kids, sub_n, sub_c = self._pack_states(path, rmax)
            path_edits = path.edits
        elif conditional:
            th_i, ba_i, th_d, ba_d, keys, kids, sub_n, sub_c = self._pack_retained(path, rmax)
        store.flush()  # every tile the retained path refers to must exist before the sweep reads it
        tuples = [host_tuple(store, dp) for dp in sigma]
        dp_idx = np.array([t[0] for t in tuples], dtype=np.int32)
        value_id = np.array([t[1] for t in tuples], dtype=np.int32)
        single_id = np.array([t[2] for t in tuples], dtype=np.int32)
        dp_op = np.array([t[3] for t in tuples], dtype=np.float64)
        dp_opn = np.array([t[4] for t in tuples], dtype=np.float64)
        dp_marg = np.array([t[5] for t in tuples], dtype=np.float64)
        need = N * rmax + 2 * N * T + 16
        base = store.book.reserve(need)
        while base < 0:
            store._grow()
            base = store.book.reserve(need)
        arena = store.arena
        o = self._outputs(T)
        log_uniform = float(-np.log(N))
        d = SweepDesc()
        d.N, d.T, d.Rmax, d.conditional = N, T, rmax, 1 if conditional else 0
        d.outliers = 1 if self.kernel.outlier_proposal_prob > 0 else 0
        d.perm_dist = 1 if self.perm else 0
        d.n_tab, d.n_ret_kids = n_tab, (len(kids) if conditional else 0)
        d.log_alpha, d.threshold, d.log_half = float(prior.log_alpha), self.threshold, float(self.kernel.log_half)
        # swarm.log_weights of N particles at the uniform weight: lu - (log(sum of N exp(0)) + lu)  (swarm.py:33-44)
        d.log_uniform, d.lw_uniform = log_uniform, log_uniform - (float(np.log(float(N))) + log_uniform)
        d.prior = arena.prior
        d.seed = int(seed) & ((1 << 64) - 1)
        d.tile_base, d.tile_limit = base, base + need
        keep = [dp_idx, value_id, single_id, dp_op, dp_opn, dp_marg, lfact, logn, rterm]
        (d.dp_idx, d.value_id, d.single_id, d.dp_op, d.dp_opn, d.dp_marg, d.lfact, d.logn, d.rterm) = [_ptr(a) for a in keep]
        h2d = sum(a.nbytes for a in keep)
        if conditional:
            edits = np.asarray(path_edits, dtype=np.int32).reshape(T, 2)
            ret_kind, ret_arg = np.ascontiguousarray(edits[:, 0]), np.ascontiguousarray(edits[:, 1])
            ret = [th_i, ba_i, th_d, ba_d, keys, kids, ret_kind, ret_arg]
            (d.ret_th_int, d.ret_ba_int, d.ret_th_dbl, d.ret_ba_dbl, d.ret_key, d.ret_kids, d.ret_kind,
             d.ret_arg) = [_ptr(a) if a.size else None for a in ret]
            if self.perm:
                ret += [sub_n, sub_c]
                d.ret_subtree_data, d.ret_subtree_log_count = _ptr(sub_n), _ptr(sub_c)
            h2d += sum(a.nbytes for a in ret)
        (d.anc, d.kind, d.name, d.mask, d.raw, d.logp, d.logq, d.ess, d.resampled, d.final_idx, d.status, d.tiles_used,
         d.n_launches, d.conv_pairs, d.device_ms) = [_ptr(o[k]) for k in (
             "anc", "kind", "name", "mask", "raw", "logp", "logq", "ess", "resampled", "final", "status", "tiles",
             "launches", "pairs", "ms")]
        _lib.check(self.lib.pcb_smc_sweep_dev(C.byref(arena.c), C.byref(d), arena._stream()))
        st = self.stats
        st["sweeps"] += 1
        st["launches"] += int(o["launches"][0])
        st["device_ms"] += float(o["ms"][0])
        st["conv_pairs"] += int(o["pairs"][0])
        st["tiles"] += int(o["tiles"][0])
        st["h2d_bytes"] += h2d
        st["d2h_bytes"] += T * N * (3 * 4 + 8 + 3 * 8) + T * 12 + 64
        self.kernel.num_extensions += (N - (1 if conditional else 0)) * T
        status = int(o["status"][0])
        if status:
            why = ", ".join(v for k, v in STATUS.items() if status & k)
            if status & ~1:
                raise _lib.PcbError("device-resident sweep failed: %s" % why)
            raise SweepFallback(why)
        self.last = {k: (v[:T] if v.shape[0] >= T and k not in ("final", "status", "tiles", "launches", "pairs", "ms")
                         else v) for k, v in o.items()}
        if self.history is not None:
            self.history.append({k: np.array(v, copy=True) for k, v in self.last.items()})
        return self._rebuild(sigma, o, path, T)

    # ------------------------------------------------------------------------------------------------ the picked tree
    def _rebuild(self, sigma, o, path, T):
        """Replay the picked particle's edits (the chain of to_dict / from_dict round trips along its genealogy,
        smc/swarm/tree_holder.py:80-82), starting from the latest point where the lineage is the retained path."""
        anc, kind, name, mask = o["anc"], o["kind"], o["name"], o["mask"]
        i = int(o["final"][0])
        edits, tree, t = [], None, T - 1
        retained = (lambda n: path.tree(n)) if isinstance(path, RetainedPath) else (lambda n: path[n].tree)
        while True:
            if path is not None and i == 0:
                tree = retained(t + 1)
                break
            k = int(kind[t, i])
            edits.append((t, k & 7, int(name[t, i]), int(mask[t, i])))
            if k & FLAG_RETAINED_PARENT:
                tree = retained(t)
                break
            if t == 0:
                tree = Tree(sigma[0].grid_size, self.store)
                break
            i = int(anc[t - 1, i])
            t -= 1
        for t, k, nm, mk in reversed(edits):
            dp = sigma[t]
            if k == ATTACH:
                tree.add_data_point_to_node(dp, nm)
            elif k == NEW:
                roots = list(tree.roots)
                # adoption in descending root position: the new node then folds its children in root order
                children = [roots[j] for j in range(len(roots) - 1, -1, -1) if (mk >> j) & 1]
                tree.create_root_node(children=children, data=[dp])
            else:
                tree.add_data_point_to_outliers(dp)
            tree.outliers  # noqa: B018 -- key-creation side effect of scoring a tree (see distributions.prepare)
            tree = tree.thawed_copy()
        return tree
