"""Root-level snapshot of a forest: what an SMC extension needs, and nothing else.

Both SMC extensions only touch the top of the forest -- a data point joins an existing ROOT node, or a new
root adopts some of the current roots (smc/kernels/semi_adapted.py:111-183) -- and every node below a root
is frozen for the rest of the sweep.  So instead of copying a whole `Tree` per candidate (the reference
copies ~20 k trees per PG iteration at 50 clusters), a candidate is evaluated from an immutable
`ForestState`: the ordered roots with their tile ids and counts plus the running prior statistics.  An
extension is O(#roots) tuple work and a couple of recorded tile ops; a real `Tree` is rebuilt (by replaying
the extensions, `materialise`) only for the one particle picked at the end of the sweep.

Two orders matter and are kept apart exactly as in the reference:
  * a particle is SCORED on the tree as it was built: the proposal's base tree (the parent re-read through
    `Tree.from_dict`, or the retained path's own tree) plus the edit;
  * it is EXTENDED later from its own `from_dict` round trip ("thawed"): roots and children in descending
    edge-slot order, touched tiles recomputed.  For a thawed parent the two coincide except for the child
    order of a new node, which is the set iteration order when built and the slot order afterwards.
"""

import numpy as np

from .._hostmod import load as _load_hostmod

ATTACH, NEW, OUTLIER = 0, 1, 2

_hm = _load_hostmod()
_tables = [None, 0]


def ensure_tables(prior, n):
    """Hand the extension log(k) and the FS-CRP root-count term r(k), k <= n, computed with NumPy exactly as the
    Python prior does (tree/distributions.py:39-133), so both sides add the same doubles."""
    if _tables[0] != prior._c_const or _tables[1] < n:
        n = max(n, 64)
        logs = [0.0] + [float(np.log(k)) for k in range(1, n + 2)]
        _hm.set_tables(logs, [float(prior._compute_r_term(k)) for k in range(n + 1)])
        _tables[0], _tables[1] = prior._c_const, n


def host_tuple(store, dp):
    """(idx, value tile, singleton log_p tile, outlier prob, not-outlier prob, outlier marginal) of a data point"""
    try:
        if dp._ts is store:
            return dp._t
    except AttributeError:  # a foreign data point class: not cached
        return (dp.idx, store.value_id[dp.idx], store.single_id[dp.idx], dp.outlier_prob, dp.outlier_prob_not,
                dp.outlier_marginal_prob)
    dp._t = (dp.idx, store.value_id[dp.idx], store.single_id[dp.idx], dp.outlier_prob, dp.outlier_prob_not,
             dp.outlier_marginal_prob)
    dp._ts = store
    return dp._t


class ForestState(object):
    """Constructors of `_pcbhost.ForestState` (csrc/pcb_hostmod.cpp): roots in order with their log_r / log_p tile
    ids, ordered child tiles, data counts, subtree sizes, edge slots and graph indices, plus the running prior
    statistics (crp sum, multiplicity, outlier prior, outlier marginal) and the structure hash."""

    @staticmethod
    def empty(store, grid_size=None):
        return _hm.forest_state((), (), (), (), (), (), (), (), 0, 1, 0, 0.0, 0.0, 0.0, 0.0, 0, 0, None)

    @staticmethod
    def from_tree(tree):
        """Snapshot of a `Tree` (any adjacency); None if its graph has vacancies (not an SMC-grown tree)."""
        g = tree._graph
        if g.free_nodes or g.free_edges:
            return None
        crp, mult, size, oprior, omarg = tree.prior_terms()
        root = tree._node_indices[tree._ROOT_NODE_NAME]
        gis = g.successor_indices(root)
        roots = [g.payload[i] for i in gis]
        return _hm.forest_state(
            roots, [tree._lr[i] for i in gis], [tree._lp[i] for i in gis],
            [[tree._lr[c] for c in g.successor_indices(i)] for i in gis], [len(tree._data[r]) for r in roots],
            [size[i] for i in gis], [g.inn[i][0] for i in gis], gis, g.n_nodes - 1, len(g.payload), len(g.esrc),
            crp, mult, oprior, omarg, len(tree._data[tree._OUTLIER_NODE_NAME]), tree.structure_key()[0],
            tree._last_node_added_to)


def extend(state, store, kind, dp, arg):
    """State of `from_dict(to_dict(state + edit))`; `state` must itself be a thawed state."""
    t = host_tuple(store, dp)
    if kind == NEW:
        arg = tuple(arg)
    s = state.extend(store.book, kind, t, arg)
    if s is None:  # arena full: grow it and record again (ops already recorded are memoised)
        store._grow()
        s = state.extend(store.book, kind, t, arg)
    return s


class StateScore(object):
    """As-built score of `base + edit`: host prior terms now, device data terms through a lazy handle."""

    __slots__ = ("prior_p", "prior_one", "handle", "store", "_val")

    def __init__(self, prior_p, prior_one, handle, store):
        self.prior_p, self.prior_one, self.handle, self.store = prior_p, prior_one, handle, store
        self._val = None

    def resolve(self):
        v = self._val
        if v is None:
            log_p, log_p_one = self.prior_p, self.prior_one
            if self.handle is not None:
                lse_sum, last_sum = self.store.value(self.handle)
                log_p += lse_sum
                log_p_one += last_sum
            v = self._val = (log_p, log_p_one)
        return v


_LOG = {}


def _log(n):
    """np.log of a small integer, memoised (NumPy's scalar log costs ~1 us a call)."""
    v = _LOG.get(n)
    if v is None:
        v = _LOG[n] = float(np.log(n))
    return v


def _priors(prior, n, crp, mult, sizes, oprior, omarg):
    """tree/distributions.py:39-133 from running statistics."""
    base = n * prior.log_alpha + crp
    log_p = base - (n - 1) * _log(n + 1) - mult
    ways = 0
    for z in sizes:
        ways += (z - 1) * _log(z)
    log_p_one = base + (-ways + prior._compute_r_term(len(sizes))) - mult
    extra = oprior + omarg
    return log_p + extra, log_p_one + extra


def score_edit(base, store, kind, dp, arg, tree_dist):
    """(StateScore, as-built roots order, children of the node the point went to) of `base + edit`, where
    `base` is the state of the tree the reference would have copied (thawed parent, or the retained path's
    as-built tree)."""
    t = host_tuple(store, dp)
    log_alpha = tree_dist.prior.log_alpha
    if kind == OUTLIER:
        pp, p1, sid, roots = base.score_outlier(store.book, t, log_alpha)
        return StateScore(pp, p1, None if sid < 0 else sid, store), roots, 0
    if kind == ATTACH:
        (score, nkids), = score_attach_all(base, store, dp, [arg], tree_dist)
        return score, base.roots, nkids
    # NEW: children adopted in the set's iteration order, each new edge at the head of the adjacency list
    arg = tuple(arg)
    rec = base.score_new(store.book, t, arg, log_alpha)
    if rec is None:
        store._grow()
        rec = base.score_new(store.book, t, arg, log_alpha)
    pp, p1, sid, roots, k = rec
    return StateScore(pp, p1, sid, store), roots, k


def score_attach_all(base, store, dp, nodes, tree_dist):
    """Scores of attaching `dp` to every root in `nodes`, the shared prior terms computed once;
    returns [(StateScore, number of children of the node)]."""
    t = host_tuple(store, dp)
    log_alpha = tree_dist.prior.log_alpha
    recs = base.score_attach_all(store.book, t, nodes, log_alpha)
    if recs is None:
        store._grow()
        recs = base.score_attach_all(store.book, t, nodes, log_alpha)
    return [(StateScore(pp, p1, sid, store), nkids) for pp, p1, sid, nkids in recs]
