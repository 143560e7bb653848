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

from ..utils.math import cached_log_factorial as lf

ATTACH, NEW, OUTLIER = 0, 1, 2


class ForestState(object):
    __slots__ = ("roots", "lr", "lp", "kids", "ndata", "size", "slot", "gidx", "n_nodes", "n_slots", "n_edges",
                 "crp", "mult", "oprior", "omarg", "n_out", "skey", "last_added", "index", "store", "grid_size")

    def index_of(self, node):
        ix = self.index
        if ix is None:
            ix = self.index = {r: j for j, r in enumerate(self.roots)}
        return ix[node]

    @staticmethod
    def empty(store, grid_size):
        s = ForestState()
        s.roots = s.lr = s.lp = s.kids = s.ndata = s.size = s.slot = s.gidx = ()
        s.n_nodes, s.n_slots, s.n_edges = 0, 1, 0
        s.crp = s.mult = s.omarg = 0.0
        s.oprior = 0
        s.n_out = 0
        s.skey = 0
        s.last_added = None
        s.index = None
        s.store, s.grid_size = store, grid_size
        return s

    @staticmethod
    def from_tree(tree):
        """Snapshot of a `Tree` (any adjacency); None if its graph has vacancies (not an SMC-grown tree)."""
        g = tree._graph
        if g.free_nodes or g.free_edges:
            return None
        crp, mult, size, oprior, omarg = tree.prior_terms()
        root = tree._node_indices[tree._ROOT_NODE_NAME]
        gis = g.successor_indices(root)
        s = ForestState()
        s.gidx = tuple(gis)
        s.roots = tuple(g.payload[i] for i in gis)
        s.lr = tuple(tree._lr[i] for i in gis)
        s.lp = tuple(tree._lp[i] for i in gis)
        s.kids = tuple(tuple(tree._lr[c] for c in g.successor_indices(i)) for i in gis)
        s.ndata = tuple(len(tree._data[r]) for r in s.roots)
        s.size = tuple(size[i] for i in gis)
        s.slot = tuple(g.inn[i][0] for i in gis)
        s.n_nodes = g.n_nodes - 1
        s.n_slots = len(g.payload)
        s.n_edges = len(g.esrc)
        s.crp, s.mult, s.oprior, s.omarg = crp, mult, oprior, omarg
        s.n_out = len(tree._data[tree._OUTLIER_NODE_NAME])
        s.skey = tree.structure_key()[0]
        s.last_added = tree._last_node_added_to
        s.index = None
        s.store, s.grid_size = tree.store, tree.grid_size
        return s

    # ---------------------------------------------------------------- thawed successor of a thawed state
    def extend(self, kind, dp, arg):
        """State of `from_dict(to_dict(self + edit))`; `self` must itself be a thawed state."""
        st = self.store
        s = ForestState()
        s.store, s.grid_size, s.index = st, self.grid_size, None
        s.n_out, s.omarg, s.oprior = self.n_out, self.omarg, self.oprior
        if kind == OUTLIER:
            for f in ("roots", "lr", "lp", "kids", "ndata", "size", "slot", "gidx", "n_nodes", "n_slots", "n_edges",
                      "crp", "mult"):
                setattr(s, f, getattr(self, f))
            s.skey = self.skey ^ hash((-1, self.n_out, dp.idx, "d"))
            s.n_out = self.n_out + 1
            s.omarg = self.omarg + dp.outlier_marginal_prob
            if dp.outlier_prob != 0:
                s.oprior = self.oprior + dp.outlier_prob
            s.last_added = -1
            return s
        if dp.outlier_prob != 0:
            s.oprior = self.oprior + dp.outlier_prob_not
        v = st.value_id[dp.idx]
        if kind == ATTACH:
            j = self.index_of(arg)
            n = self.ndata[j]
            lp = st.single_id[dp.idx] if self.lp[j] == st.prior_id else st.add(self.lp[j], v)
            lr = st.node(lp, self.kids[j]) if self.kids[j] else lp
            s.roots, s.kids, s.size, s.slot, s.gidx = self.roots, self.kids, self.size, self.slot, self.gidx
            s.lp = self.lp[:j] + (lp,) + self.lp[j + 1:]
            s.lr = self.lr[:j] + (lr,) + self.lr[j + 1:]
            s.ndata = self.ndata[:j] + (n + 1,) + self.ndata[j + 1:]
            s.n_nodes, s.n_slots, s.n_edges, s.mult = self.n_nodes, self.n_slots, self.n_edges, self.mult
            s.crp = self.crp + (lf(n) - lf(n - 1) if n >= 1 else 0.0)
            s.skey = self.skey ^ hash((arg, n, dp.idx, "d"))
            s.last_added = arg
            return s
        # NEW root adopting the roots in `arg` (a frozenset of names)
        take = [j for j, r in enumerate(self.roots) if r in arg]  # descending slot order = thawed child order
        keep = [j for j, r in enumerate(self.roots) if r not in arg]
        name = self.n_nodes
        gi = self.n_slots
        root_gi = 0
        lp = st.single_id[dp.idx]
        kids = tuple(self.lr[j] for j in take)
        lr = st.node(lp, kids) if kids else lp
        s.roots = (name,) + tuple(self.roots[j] for j in keep)
        s.lr = (lr,) + tuple(self.lr[j] for j in keep)
        s.lp = (lp,) + tuple(self.lp[j] for j in keep)
        s.kids = (kids,) + tuple(self.kids[j] for j in keep)
        s.ndata = (1,) + tuple(self.ndata[j] for j in keep)
        s.size = (1 + sum(self.size[j] for j in take),) + tuple(self.size[j] for j in keep)
        s.slot = (self.n_edges,) + tuple(self.slot[j] for j in keep)
        s.gidx = (gi,) + tuple(self.gidx[j] for j in keep)
        s.n_nodes, s.n_slots, s.n_edges = self.n_nodes + 1, self.n_slots + 1, self.n_edges + 1
        R, k = len(self.roots), len(take)
        s.mult = self.mult - lf(R) + lf(R - k + 1) + lf(k)
        s.crp = self.crp
        key = self.skey ^ hash((self.n_edges, root_gi, gi)) ^ hash((name, 0, dp.idx, "d"))
        for j in take:
            e, c = self.slot[j], self.gidx[j]
            key ^= hash((e, root_gi, c)) ^ hash((e, gi, c))
        s.skey = key
        s.last_added = name
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


def score_edit(base, kind, dp, arg, tree_dist):
    """(StateScore, as-built roots order, children of the node the point went to) of `base + edit`, where
    `base` is the state of the tree the reference would have copied (thawed parent, or the retained path's
    as-built tree)."""
    st = base.store
    prior = tree_dist.prior
    oprior, omarg = base.oprior, base.omarg
    if kind == OUTLIER:
        if dp.outlier_prob != 0:
            oprior = oprior + dp.outlier_prob
        omarg = omarg + dp.outlier_marginal_prob
        lp, lp1 = _priors(prior, base.n_nodes, base.crp, base.mult, base.size, oprior, omarg)
        handle = st.score(base.lr) if base.lr else None
        return StateScore(lp, lp1, handle, st), base.roots, 0
    if dp.outlier_prob != 0:
        oprior = oprior + dp.outlier_prob_not
    v = st.value_id[dp.idx]
    if kind == ATTACH:
        j = base.index_of(arg)
        n = base.ndata[j]
        crp = base.crp + (lf(n) - lf(n - 1) if n >= 1 else 0.0)
        lr = st.add(base.lr[j], v)  # log_r += value on the node itself (tree/tree_node.py:47-52)
        lp, lp1 = _priors(prior, base.n_nodes, crp, base.mult, base.size, oprior, omarg)
        handle = st.score(base.lr[:j] + (lr,) + base.lr[j + 1:])
        return StateScore(lp, lp1, handle, st), base.roots, len(base.kids[j])
    # NEW: children adopted in the set's iteration order, each new edge at the head of the adjacency list
    order = [base.index_of(c) for c in arg]
    kids = tuple(base.lr[j] for j in reversed(order))
    lp_t = st.single_id[dp.idx]
    lr = st.node(lp_t, kids) if kids else lp_t
    taken = set(order)
    keep = [j for j in range(len(base.roots)) if j not in taken]
    R, k = len(base.roots), len(order)
    mult = base.mult - lf(R) + lf(R - k + 1) + lf(k)
    sizes = (1 + sum(base.size[j] for j in order),) + tuple(base.size[j] for j in keep)
    lp, lp1 = _priors(prior, base.n_nodes + 1, base.crp, mult, sizes, oprior, omarg)
    handle = st.score((lr,) + tuple(base.lr[j] for j in keep))
    roots = (base.n_nodes,) + tuple(base.roots[j] for j in keep)
    return StateScore(lp, lp1, handle, st), roots, k


def score_attach_all(base, dp, nodes, tree_dist):
    """`score_edit(base, ATTACH, dp, node)` for every node of `nodes` with the shared prior terms computed once;
    returns [(StateScore, number of children of the node)]."""
    st = base.store
    prior = tree_dist.prior
    oprior = base.oprior
    if dp.outlier_prob != 0:
        oprior = oprior + dp.outlier_prob_not
    n = base.n_nodes
    head = n * prior.log_alpha
    ways = 0
    for z in base.size:
        ways += (z - 1) * _log(z)
    tail_p = -(n - 1) * _log(n + 1) - base.mult + (oprior + base.omarg)
    tail_one = (-ways + prior._compute_r_term(len(base.size))) - base.mult + (oprior + base.omarg)
    v = st.value_id[dp.idx]
    lr_all = base.lr
    out = []
    for node in nodes:
        j = base.index_of(node)
        m = base.ndata[j]
        crp = head + (base.crp + (lf(m) - lf(m - 1) if m >= 1 else 0.0))
        lr = st.add(lr_all[j], v)
        handle = st.score(lr_all[:j] + (lr,) + lr_all[j + 1:])
        out.append((StateScore(crp + tail_p, crp + tail_one, handle, st), len(base.kids[j])))
    return out
