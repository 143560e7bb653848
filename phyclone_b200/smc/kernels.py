"""SMC proposal kernels (reference smc/kernels/{base,semi_adapted,fully_adapted,bootstrap}.py).

Same classes, candidate order and probabilities as the reference; construction is split in two so
that a whole SMC step batches on the device: `__init__` only *builds* the candidate trees (host
bookkeeping + recorded tile ops), the normalised proposal is computed on first use, i.e. after the
sampler has built the candidates of every distinct parent (one flush for all of them).
"""

import itertools
from collections import OrderedDict

import numpy as np

from ..tree.tree import Tree
from ..utils.math import cached_log_binomial_coefficient, discrete_rvs, log_binomial_coefficient, log_normalize
from .forest_state import NEW, OUTLIER, ForestState, ensure_tables
from .swarm import Particle, StateHolder, TreeHolder


class _LRU(object):
    """functools.lru_cache(maxsize) semantics for explicit keys (semi_adapted.py:186-237)."""

    def __init__(self, maxsize=1024):
        self.maxsize = maxsize
        self.d = OrderedDict()
        self.hits = self.misses = 0

    def get(self, key, make):
        d = self.d
        if key in d:
            d.move_to_end(key)
            self.hits += 1
            return d[key]
        self.misses += 1
        val = d[key] = make()
        if len(d) > self.maxsize:
            d.popitem(last=False)
        return val

    def clear(self):
        self.d.clear()


class Kernel(object):
    """smc/kernels/base.py:4-67"""

    def __init__(self, tree_dist, rng, store, outlier_proposal_prob=0.0, perm_dist=None):
        self.tree_dist = tree_dist
        self.perm_dist = perm_dist
        self._rng = rng
        self.store = store
        self.outlier_proposal_prob = outlier_proposal_prob
        self.num_extensions = 0
        self._dist_cache = _LRU(1024)
        self._new_tree_cache = _LRU(1024)

    @property
    def rng(self):
        return self._rng

    def fast_draw_address(self):
        """Address of numpy's bitgen_t if this kernel's draws may be done in C (see samplers._propose_all)."""
        return None

    def clear_proposal_dist_caches(self):
        """utils/dev.py:12-17"""
        self._dist_cache.clear()
        self._new_tree_cache.clear()

    def _make_proposal(self, data_point, parent_particle, parent_tree):
        raise NotImplementedError

    def get_proposal_distribution(self, data_point, parent_particle, parent_tree=None):
        # keyed on particle VALUE like the reference's lru_cache (semi_adapted.py:219-237): the first proposal
        # built for a given frozen tree is the one every equal particle gets, child order and all
        key = (data_point.idx, parent_particle._holder.vkey if parent_particle is not None else None,
               self.tree_dist.prior.alpha)
        return self._dist_cache.get(key, lambda: self._make_proposal(data_point, parent_particle, parent_tree))

    def create_particle(self, log_q, parent_particle, tree_holder):
        """base.py:38-57; log_w = log_p - parent.log_p - log_q is finalised lazily by `Particle.log_w`.  The
        root-permutation density (`perm_dist`, base.py:49-55; None in production, run.py:403-417) enters the weights
        of the device-resident sweep only (smc/device_sweep.py): the retained path built here never reads them."""
        return Particle(log_q, parent_particle, tree_holder)

    def propose_particle(self, data_point, parent_particle):
        """base.py:59-67"""
        self.num_extensions += 1
        proposal_dist = self.get_proposal_distribution(data_point, parent_particle)
        tree_holder = proposal_dist.sample()
        log_q = proposal_dist.log_p(tree_holder)
        return self.create_particle(log_q, parent_particle, tree_holder)


class ProposalDistribution(object):
    def __init__(self, data_point, kernel, parent_particle, parent_tree=None):
        self.data_point = data_point
        self.kernel = kernel
        self.tree_dist = kernel.tree_dist
        self.outlier_proposal_prob = kernel.outlier_proposal_prob
        self.parent_particle = parent_particle
        self._rng = kernel.rng
        if parent_particle is not None:
            self.parent_tree = parent_tree if parent_tree is not None else parent_particle.tree
        else:
            self.parent_tree = None

    def _empty_tree(self):
        return (self.parent_particle is None) or (len(self.parent_particle.tree_roots) == 0)

    def _holder(self, tree):
        return TreeHolder(tree, self.tree_dist)

    def _get_existing_node_trees(self):
        """semi_adapted.py:111-126"""
        trees = []
        if self.parent_particle is None:
            return trees
        for node in self.parent_particle.tree_roots:
            tree = self.parent_tree.copy()
            tree.add_data_point_to_node(self.data_point, node)
            trees.append(self._holder(tree))
        return trees

    def _get_outlier_tree(self):
        """semi_adapted.py:128-141"""
        if self.parent_particle is None:
            tree = Tree(self.data_point.grid_size, self.kernel.store)
        else:
            tree = self.parent_tree.copy()
        tree.add_data_point_to_outliers(self.data_point)
        return self._holder(tree)


class SemiAdaptedProposalDistribution(ProposalDistribution):
    """smc/kernels/semi_adapted.py:11-183"""

    def __init__(self, data_point, kernel, parent_particle, parent_tree=None):
        super().__init__(data_point, kernel, parent_particle, parent_tree)
        self.log_half = kernel.log_half
        self.parent_is_empty_tree = False
        trees = self._get_existing_node_trees()
        if self.outlier_proposal_prob > 0:
            trees.append(self._get_outlier_tree())
        if self._empty_tree():
            self.parent_is_empty_tree = True
            if self.parent_particle is None:
                tree = Tree(self.data_point.grid_size, kernel.store)
            else:
                tree = self.parent_tree.copy()
            tree.create_root_node(children=[], data=[self.data_point])
            trees.append(self._holder(tree))
        else:
            self._cached_log_old_num_roots = np.log(len(self.parent_particle.tree_roots) + 1)
        self._curr_trees = trees
        self._index = {h.uid: i for i, h in enumerate(trees)}
        self._log_q = None
        self._q_dist = None
        self.parent_tree = None

    def _finish(self):
        """_set_log_p_dist / _set_q_dist (semi_adapted.py:151-163); first score read of the step."""
        log_q = log_normalize(np.array([x.log_p for x in self._curr_trees]))
        q = np.exp(log_q)
        q_sum = np.sum(q)
        assert abs(1 - q_sum) < 1e-6
        self._log_q = log_q
        self._q_dist = q / q_sum

    def candidate_index(self, tree_holder):
        """Index of a holder among the enumerated candidates (the reference looks it up by value)."""
        i = self._index.get(tree_holder.uid)
        if i is None:
            # a tree built elsewhere (the retained path): find the candidate that put the data point in the same place
            node = tree_holder.labels[self.data_point.idx]
            if node == tree_holder.outlier_node_name:
                i = len(self.parent_particle.tree_roots) if self.parent_particle is not None else 0
            elif self.parent_is_empty_tree:
                i = len(self._curr_trees) - 1
            else:
                i = list(self.parent_particle.tree_roots).index(node)
        return i

    def log_p(self, tree_holder):
        """semi_adapted.py:36-61"""
        if self._log_q is None:
            self._finish()
        if self.parent_is_empty_tree:
            return self._log_q[self.candidate_index(tree_holder)]
        node = tree_holder.node_last_added_to  # == tree_holder.labels[data_point.idx] (semi_adapted.py:44-45)
        # "node in parent.tree_nodes": SMC only ever adds to a ROOT of the parent, or to a brand-new node whose name
        # (= number of nodes) no parent node has
        kind = getattr(tree_holder, "kind", None)
        if (kind is not None and kind != NEW) or (kind is None and (
                node == tree_holder.outlier_node_name or node in self.parent_particle.tree_roots)):
            return self.log_half + self._log_q[self.candidate_index(tree_holder)]
        old_num_roots = len(self.parent_particle.tree_roots)
        num_children = tree_holder.num_children_on_node_that_matters
        return self.log_half - (self._cached_log_old_num_roots + cached_log_binomial_coefficient(old_num_roots,
                                                                                                  num_children))

    def sample(self):
        """semi_adapted.py:63-83"""
        if self._q_dist is None:
            self._finish()
        if self.parent_is_empty_tree:
            return self._propose_existing_node()
        u = self._rng.random()
        if u < 0.5:
            return self._propose_existing_node()
        return self._propose_new_node()

    def _propose_existing_node(self):
        idx = self._rng.multinomial(1, self._q_dist).argmax()
        return self._curr_trees[idx]

    def _propose_new_node(self):
        """semi_adapted.py:165-194"""
        num_roots = len(self.parent_particle.tree_roots)
        num_children = self._rng.integers(0, num_roots + 1)
        if num_children == 0:
            children = []
        else:
            children = self._rng.choice(self.parent_particle.tree_roots, num_children, replace=False)
        children = frozenset(children)
        key = (self.parent_particle._holder.vkey, self.data_point.idx, children)

        def make():
            tree = self.parent_particle.tree
            tree.create_root_node(children=children, data=[self.data_point])
            return self._holder(tree)

        return self.kernel._new_tree_cache.get(key, make)


class StateSemiAdaptedProposalDistribution(object):
    """Semi-adapted proposal (smc/kernels/semi_adapted.py:11-183) evaluated on root-level forest states: same
    candidates, order, probabilities and RNG calls as the tree-copying version above, O(#roots) work each."""

    fast_draws = True

    def __init__(self, data_point, kernel, parent_particle, parent_state, base_state):
        self.data_point = data_point
        self.kernel = kernel
        self.roots_array = (np.asarray(parent_particle.tree_roots, dtype=np.int64) if parent_particle is not None
                            else np.zeros(0, dtype=np.int64))
        self.tree_dist = kernel.tree_dist
        self.parent_particle = parent_particle
        self._rng = kernel.rng
        self.log_half = kernel.log_half
        self._pstate = parent_state
        ph = parent_particle._holder if parent_particle is not None else None
        self._pholder = ph
        td = self.tree_dist
        st = self._store = kernel.store
        trees = []
        if parent_particle is not None and len(parent_particle.tree_roots):
            nodes = parent_particle.tree_roots.tolist()
            shared_roots = np.asarray(base_state.roots)  # attaching does not change the as-built root order
            trees = StateHolder.attach_all(st, ph, parent_state, base_state, data_point, nodes, td, shared_roots)
        if kernel.outlier_proposal_prob > 0:
            trees.append(StateHolder.edit(st, ph, parent_state, base_state, OUTLIER, data_point, None, td))
        self.parent_is_empty_tree = parent_particle is None or len(parent_particle.tree_roots) == 0
        if self.parent_is_empty_tree:
            trees.append(StateHolder.edit(st, ph, parent_state, base_state, NEW, data_point, frozenset(), td))
        else:
            self._cached_log_old_num_roots = np.log(len(parent_particle.tree_roots) + 1)
        self._curr_trees = trees
        self._index = {h.uid: i for i, h in enumerate(trees)}
        self._log_q = None
        self._q_dist = None

    _finish = SemiAdaptedProposalDistribution._finish
    candidate_index = SemiAdaptedProposalDistribution.candidate_index
    log_p = SemiAdaptedProposalDistribution.log_p
    sample = SemiAdaptedProposalDistribution.sample
    _propose_existing_node = SemiAdaptedProposalDistribution._propose_existing_node

    def _propose_new_node(self):
        num_roots = len(self.parent_particle.tree_roots)
        num_children = self._rng.integers(0, num_roots + 1)
        if num_children == 0:
            children = []
        else:
            children = self._rng.choice(self.parent_particle.tree_roots, num_children, replace=False)
        return self.new_node_holder(frozenset(children))

    def new_node_holder(self, children):
        """get_cached_new_tree (semi_adapted.py:186-194): the new node always goes onto the parent re-read through
        from_dict; one holder per (parent value, data point, child set)."""
        key = (self._pholder.vkey, self.data_point.idx, children)
        return self.kernel._new_tree_cache.get(
            key, lambda: StateHolder.edit(self._store, self._pholder, self._pstate, self._pstate, NEW, self.data_point,
                                          children, self.tree_dist))


class SemiAdaptedKernel(Kernel):
    log_half = np.log(0.5)

    use_states = True  # evaluate candidates from root-level ForestStates instead of copied trees

    def fast_draw_address(self):
        if not self.use_states:
            return None
        from .. import _hostlib

        return _hostlib.bitgen_address(self._rng)

    def _make_proposal(self, data_point, parent_particle, parent_tree):
        if self.use_states:
            ensure_tables(self.tree_dist.prior, self.store.T + 2)
            if parent_particle is None:
                empty = ForestState.empty(self.store, data_point.grid_size)
                return StateSemiAdaptedProposalDistribution(data_point, self, None, empty, empty)
            pstate = parent_particle._holder.thawed_state()
            base = pstate if parent_tree is None else ForestState.from_tree(parent_tree)
            if pstate is not None and base is not None:
                return StateSemiAdaptedProposalDistribution(data_point, self, parent_particle, pstate, base)
        return SemiAdaptedProposalDistribution(data_point, self, parent_particle, parent_tree)


class FullyAdaptedProposalDistribution(ProposalDistribution):
    """smc/kernels/fully_adapted.py:11-105"""

    def __init__(self, data_point, kernel, parent_particle, parent_tree=None):
        super().__init__(data_point, kernel, parent_particle, parent_tree)
        trees = self._get_existing_node_trees()
        if self.parent_particle is None:
            tree = Tree(self.data_point.grid_size, kernel.store)
            tree.create_root_node(children=[], data=[self.data_point])
            trees.append(self._holder(tree))
        else:
            roots = self.parent_particle.tree_roots
            for r in range(0, len(roots) + 1):
                for children in itertools.combinations(roots, r):
                    tree = self.parent_tree.copy()
                    tree.create_root_node(children=children, data=[self.data_point])
                    trees.append(self._holder(tree))
        if self.outlier_proposal_prob > 0:
            trees.append(self._get_outlier_tree())
        self._curr_trees = trees
        self._index = {h.uid: i for i, h in enumerate(trees)}
        self._log_q = None
        self.parent_tree = None

    def _finish(self):
        self._log_q = log_normalize(np.array([x.log_p for x in self._curr_trees]))

    def candidate_index(self, tree_holder):
        i = self._index.get(tree_holder.uid)
        if i is None:  # retained path: match on the frozen structure
            want = tree_holder._built.to_dict()
            for j, h in enumerate(self._curr_trees):
                have = h._built.to_dict()
                if have["graph"] == want["graph"] and have["node_data"] == want["node_data"]:
                    return j
            raise KeyError("retained tree is not among the fully-adapted candidates")
        return i

    def log_p(self, tree_holder):
        if self._log_q is None:
            self._finish()
        return self._log_q[self.candidate_index(tree_holder)]

    def sample(self):
        if self._log_q is None:
            self._finish()
        idx = discrete_rvs(np.exp(self._log_q), self._rng)
        return self._curr_trees[idx]


class FullyAdaptedKernel(Kernel):
    def _make_proposal(self, data_point, parent_particle, parent_tree):
        return FullyAdaptedProposalDistribution(data_point, self, parent_particle, parent_tree)


class BootstrapProposalDistribution(ProposalDistribution):
    """smc/kernels/bootstrap.py:8-117"""

    def log_p(self, tree_holder):
        rho = self.outlier_proposal_prob
        node = tree_holder.labels[self.data_point.idx]
        outlier = tree_holder.outlier_node_name
        if self.parent_particle is None:
            return np.log(rho) if node == outlier else np.log(1 - rho)
        if node == outlier:
            return np.log(rho)
        if node in self.parent_tree.nodes:
            return np.log((1 - rho) / 2) - np.log(len(self.parent_particle.tree_roots))
        old_num_roots = len(self.parent_particle.tree_roots)
        log_p = np.log((1 - rho) / 2)
        if old_num_roots > 0:
            log_p -= np.log(old_num_roots + 1) + log_binomial_coefficient(
                old_num_roots, tree_holder.num_children_on_node_that_matters)
        return log_p

    def sample(self):
        rho = self.outlier_proposal_prob
        u = self._rng.random()
        if self.parent_particle is None:
            tree = Tree(self.data_point.grid_size, self.kernel.store)
            if u < (1 - rho):
                node = tree.create_root_node([])
                tree.add_data_point_to_node(self.data_point, node)
            else:
                tree.add_data_point_to_outliers(self.data_point)
        elif len(self.parent_tree.nodes) == 0:
            tree = self._propose_new_node() if u < (1 - rho) else self._propose_outlier()
        elif u < (1 - rho) / 2:
            node = self._rng.choice(list(self.parent_particle.tree_roots))
            tree = self.parent_tree.copy()
            tree.add_data_point_to_node(self.data_point, node)
        elif u < (1 - rho):
            tree = self._propose_new_node()
        else:
            tree = self._propose_outlier()
        return self._holder(tree)

    def _propose_new_node(self):
        num_roots = len(self.parent_particle.tree_roots)
        num_children = self._rng.integers(0, num_roots + 1)
        children = self._rng.choice(self.parent_particle.tree_roots, num_children, replace=False)
        tree = self.parent_tree.copy()
        tree.create_root_node(children=children, data=[self.data_point])
        return tree

    def _propose_outlier(self):
        tree = self.parent_tree.copy()
        tree.add_data_point_to_outliers(self.data_point)
        return tree


class BootstrapKernel(Kernel):
    def get_proposal_distribution(self, data_point, parent_particle, parent_tree=None):
        return BootstrapProposalDistribution(data_point, self, parent_particle, parent_tree)
