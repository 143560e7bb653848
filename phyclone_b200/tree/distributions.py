"""FS-CRP prior and tree joint distribution (reference tree/distributions.py:6-219).

The integer/scalar prior terms stay on the host; the two data terms (sum_s LSE_g and
sum_s last column of the dummy root's log_r) come from the device through a lazy score handle,
so `prepare(tree)` costs no kernel launch and a whole batch of prepared trees resolves in one.
"""

import numpy as np

from ..utils.math import cached_log_factorial


_R_CACHE = {}


class FSCRPDistribution(object):
    """FSCRP prior distribution on trees."""

    __slots__ = ("_alpha", "log_alpha", "_c_const")

    def __init__(self, alpha, c_const=1000):
        self.alpha = alpha
        self.c_const = c_const

    def __eq__(self, other):
        return self.alpha == other.alpha

    def __hash__(self):
        return hash(self.alpha)

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, alpha):
        self._alpha = alpha
        self.log_alpha = np.log(alpha)

    @property
    def c_const(self):
        return self._c_const

    @c_const.setter
    def c_const(self, c_const):
        self._c_const = np.log(c_const)

    def _compute_r_term(self, num_roots):
        """distributions.py:103-133 (a function of the root count only: memoised)"""
        cache = self.__dict__.setdefault("_r_cache", {}) if hasattr(self, "__dict__") else _R_CACHE
        key = (num_roots, self._c_const)
        val = cache.get(key)
        if val is None:
            val = cache[key] = self._r_term_uncached(num_roots)
        return val

    def _r_term_uncached(self, num_roots):
        log_const = self._c_const
        if num_roots == 0:
            z_term, num_roots = 0.0, 1
        else:
            z_term = np.log1p(-np.exp(-(log_const * num_roots))) - np.log1p(-np.exp(-(log_const * 1)))
        return -(z_term + (log_const * (num_roots - 1)))

    def compute_both_log_p_and_log_p_one_priors(self, tree, tree_node_data=None):
        """distributions.py:39-101, from the tree's running statistics (`Tree.prior_terms`)."""
        crp_sum, multiplicity, size, _, _ = tree.prior_terms()
        num_nodes = tree.get_number_of_nodes()
        crp = 0
        crp += num_nodes * self.log_alpha
        crp += crp_sum
        log_p = crp - (num_nodes - 1) * np.log(num_nodes + 1)
        log_p -= multiplicity
        root_idx = tree._graph.successor_indices(tree._node_indices[tree.root_node_name])
        num_ways = 0
        for i in root_idx:
            n = size[i]
            num_ways += (n - 1) * np.log(n)
        log_p_one = crp + (-num_ways + self._compute_r_term(len(root_idx)))
        log_p_one -= multiplicity
        return log_p, log_p_one

    def log_p(self, tree, tree_node_data=None):
        return self.compute_both_log_p_and_log_p_one_priors(tree, tree_node_data)[0]

    def log_p_one(self, tree, tree_node_data=None):
        return self.compute_both_log_p_and_log_p_one_priors(tree, tree_node_data)[1]


class PendingScore(object):
    """Host prior terms + a lazy device handle; `.resolve()` -> (log_p, log_p_one)."""

    __slots__ = ("prior_p", "prior_one", "handle", "store", "outlier_marg", "_val")

    def __init__(self, prior_p, prior_one, handle, store, outlier_marg):
        self.prior_p, self.prior_one, self.handle, self.store = prior_p, prior_one, handle, store
        self.outlier_marg = outlier_marg
        self._val = None

    def resolve(self):
        if self._val is None:
            log_p, log_p_one = self.prior_p, self.prior_one
            if self.handle is not None:
                lse_sum, last_sum = self.store.value(self.handle)
                log_p += lse_sum
                log_p_one += last_sum
            for m in self.outlier_marg:
                log_p += m
                log_p_one += m
            self._val = (log_p, log_p_one)
        return self._val


class TreeJointDistribution(object):
    __slots__ = "prior"

    def __init__(self, prior):
        self.prior = prior

    def __eq__(self, other):
        return self.prior == other.prior

    def __hash__(self):
        return hash(self.prior)

    @staticmethod
    def outlier_prior(tree_node_data, outlier_node_name):
        """distributions.py:208-219"""
        log_p = 0
        for node, node_data in tree_node_data.items():
            for data_point in node_data:
                if data_point.outlier_prob != 0:
                    if node == outlier_node_name:
                        log_p += data_point.outlier_prob
                    else:
                        log_p += data_point.outlier_prob_not
        return log_p

    def prepare(self, tree):
        """Record everything `compute_both_log_p_and_log_p_one` (distributions.py:186-206) needs; no launch."""
        log_p, log_p_one = self.prior.compute_both_log_p_and_log_p_one_priors(tree)
        pt = tree.prior_terms()
        log_p += pt[3]
        log_p_one += pt[3]
        tree.outliers  # noqa: B018 -- the reference touches _data[-1] here, which fixes that key's position
        return PendingScore(log_p, log_p_one, tree.score_handle(), tree.store, (pt[4],) if pt[4] != 0.0 else ())

    def compute_both_log_p_and_log_p_one(self, tree):
        return self.prepare(tree).resolve()

    def log_p(self, tree):
        return self.prepare(tree).resolve()[0]

    def log_p_one(self, tree):
        return self.prepare(tree).resolve()[1]
