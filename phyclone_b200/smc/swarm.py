"""TreeHolder / Particle / ParticleSwarm (reference smc/swarm/*.py).

A holder keeps the as-built tree and a lazy score: reading `log_p` is what finally triggers the
batched device evaluation, so building all candidate holders of an SMC step first and reading
their scores afterwards costs one flush.
"""

import itertools

import numpy as np

from .._hostmod import load as _load_hostmod
from ..utils.math import log_sum_exp
from .forest_state import host_tuple

_hm = _load_hostmod()
_uid = itertools.count()  # holders made by the extension count downwards from -1: the two never collide


class TreeHolder(object):
    __slots__ = ("_tree_dist", "_pending", "_built", "_thawed", "tree_roots", "_tree_nodes", "_labels",
                 "node_last_added_to", "num_children_on_node_that_matters", "outlier_node_name", "log_pdf", "uid",
                 "vkey", "_tstate")

    def __init__(self, tree, tree_dist, perm_dist=None):
        self.uid = next(_uid)  # cache key (object ids can be recycled within an iteration)
        self._tree_dist = tree_dist
        self.outlier_node_name = tree.outlier_node_name
        self.log_pdf = 0.0  # perm_dist is None in production (run.py:416)
        self._pending = tree_dist.prepare(tree)
        self.tree_roots = np.asarray(tree.roots)
        self._tree_nodes = None
        self._built = tree  # callers hand the tree over: it is never mutated afterwards
        self.vkey = tree.structure_key()  # "equal particle" in the reference's sense (dict equality of frozen trees)
        self._thawed = None
        self._tstate = False
        self._labels = None
        self.node_last_added_to = tree.node_last_added_to
        if self.node_last_added_to != tree.outlier_node_name:
            self.num_children_on_node_that_matters = tree.get_number_of_children(self.node_last_added_to)
        else:
            self.num_children_on_node_that_matters = 0

    @property
    def tree_nodes(self):
        if self._tree_nodes is None:
            self._tree_nodes = set(self._built.nodes)
        return self._tree_nodes

    @property
    def labels(self):
        if self._labels is None:
            self._labels = self._built.labels
        return self._labels

    @property
    def log_p(self):
        return self._pending.resolve()[0]

    @property
    def log_p_one(self):
        return self._pending.resolve()[1]

    @property
    def tree(self):
        """What `Tree.from_dict(holder._tree)` gives in the reference (tree_holder.py:80-82): a fresh
        tree whose adjacency follows slot order and whose tiles were all refreshed."""
        if self._thawed is None:
            self._thawed = self._built.thawed_copy()
        return self._thawed.copy()

    def thawed_state(self):
        """Root-level `ForestState` of `self.tree` (None if the tree is not SMC-grown)."""
        if self._tstate is False:
            from .forest_state import ForestState

            if self._thawed is None:
                self._thawed = self._built.thawed_copy()
            self._tstate = ForestState.from_tree(self._thawed)
        return self._tstate


class StateHolder(_load_hostmod().Holder):
    """Holder of `thawed(parent) + one SMC edit`, evaluated from root-level `ForestState`s instead of a copied
    `Tree` (see forest_state.py).  Same read interface as `TreeHolder`.  The object itself -- score, lazy
    `log_p` / `log_p_one` (the first read flushes the tile store), `thawed_state()`, `vkey` -- is the `_pcbhost.Holder`
    C++ type and is created by the factories below; this subclass adds what needs a real `Tree`: it is only
    rebuilt, by replaying the edits, when `.tree` is asked for."""

    __slots__ = ()

    @staticmethod
    def attach_all(store, parent_holder, parent_state, base_state, dp, nodes, tree_dist, shared_roots):
        """One candidate per root in `nodes` (the data point joins that root), scored with shared prior terms."""
        t = host_tuple(store, dp)
        la = tree_dist.prior.log_alpha
        hs = _hm.attach_holders(StateHolder, store, base_state, parent_state, parent_holder, dp, t, nodes, la, shared_roots)
        if hs is None:  # arena full: grow and record again (ops already recorded are memoised)
            store._grow()
            hs = _hm.attach_holders(StateHolder, store, base_state, parent_state, parent_holder, dp, t, nodes, la,
                                    shared_roots)
        return hs

    @staticmethod
    def edit(store, parent_holder, parent_state, base_state, kind, dp, arg, tree_dist):
        """Candidate `base + (NEW root adopting the roots in arg | OUTLIER)`."""
        t = host_tuple(store, dp)
        la = tree_dist.prior.log_alpha
        h = _hm.edit_holder(StateHolder, store, base_state, parent_state, parent_holder, kind, dp, t, arg, la)
        if h is None:
            store._grow()
            h = _hm.edit_holder(StateHolder, store, base_state, parent_state, parent_holder, kind, dp, t, arg, la)
        return h

    @property
    def tree_roots(self):
        r = self._roots_arr
        if r is None:
            r = self._roots_arr = np.asarray(self._roots_tuple)
        return r

    @property
    def labels(self):
        return self.tree.labels

    @property
    def tree_nodes(self):
        return set(self.tree.nodes)

    @property
    def tree(self):
        """Replay: thaw(parent) + edit, then its own `from_dict` round trip -- the chain of to_dict/from_dict
        round trips the reference performs along a particle's genealogy."""
        if self._thawed is None:
            from ..tree.tree import Tree
            from .forest_state import ATTACH, NEW

            # iterative (a genealogy is as long as the data set: 2000 edits at config #5)
            chain, h = [], self
            while isinstance(h, StateHolder) and h._thawed is None:
                chain.append(h)
                h = h._parent_holder
            t = Tree(self._grid_size, self._store) if h is None else h.tree
            for hh in reversed(chain):
                if hh.kind == ATTACH:
                    t.add_data_point_to_node(hh.dp, hh.arg)
                elif hh.kind == NEW:
                    t.create_root_node(children=hh.arg, data=[hh.dp])
                else:
                    t.add_data_point_to_outliers(hh.dp)
                t.outliers  # noqa: B018 -- key-creation side effect of scoring a tree (see distributions.prepare)
                t = t.thawed_copy()
            self._thawed = t
        return self._thawed.copy()


# smc/swarm/particle.py: (log_q, parent_particle, holder) with the incremental weight (kernels/base.py:38-57) evaluated on
# first read, so that every new tree of an SMC step is scored by one batched flush.  Native type: csrc/pcb_hostmod.cpp.
Particle = _hm.Particle


class ParticleSwarm(object):
    """smc/swarm/swarm.py:6-81"""

    __slots__ = ("particles", "_log_norm_const", "_unnormalized_log_weights")

    def __init__(self):
        self.particles = []
        self._log_norm_const = None
        self._unnormalized_log_weights = []

    @property
    def ess(self):
        return 1 / np.square(self.weights).sum()

    @property
    def log_norm_const(self):
        if self._log_norm_const is None:
            self._log_norm_const = log_sum_exp(self.unnormalized_log_weights)
        return self._log_norm_const

    @property
    def log_weights(self):
        return self.unnormalized_log_weights - self.log_norm_const

    @property
    def num_particles(self):
        return len(self.particles)

    @property
    def relative_ess(self):
        return self.ess / self.num_particles

    @property
    def unnormalized_log_weights(self):
        return np.fromiter(self._unnormalized_log_weights, dtype=np.float64, count=len(self._unnormalized_log_weights))

    @property
    def weights(self):
        weights = np.exp(self.log_weights)
        weights /= weights.sum()
        return weights

    @staticmethod
    def of(particles, log_weights):
        """Swarm holding the given particles / unnormalised log weights (lists; taken over, not copied)."""
        sw = ParticleSwarm()
        sw.particles = particles
        sw._unnormalized_log_weights = log_weights
        return sw

    def add_particle(self, log_weight, particle):
        self.particles.append(particle)
        self._unnormalized_log_weights.append(log_weight)
        self._log_norm_const = None

    def to_list(self):
        return zip(self.particles, self.weights)
