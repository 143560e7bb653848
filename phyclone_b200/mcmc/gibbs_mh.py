"""Data-point Gibbs and prune-regraph moves (reference mcmc/gibbs_mh.py:6-129).

Every candidate tree of a move is built first (host bookkeeping + recorded tile ops) and the
candidate scores are read afterwards, so one Gibbs step is one batched device evaluation.
"""

import numpy as np

from ..utils.math import cached_log_factorial as lf
from ..utils.math import exp_normalize, log_normalize


class DataPointSampler(object):
    def __init__(self, tree_dist, rng, outliers=False):
        self.tree_dist = tree_dist
        self.outliers = outliers
        self._rng = rng

    def sample_tree(self, tree):
        tree_labels = tree.labels
        data_idxs = list(tree_labels.keys())
        self._rng.shuffle(data_idxs)
        for data_idx in data_idxs:
            old_node = tree_labels[data_idx]
            if tree.get_data_len(old_node) > 1:
                tree = self._sample_tree(data_idx, tree, old_node)
                tree_labels = tree.labels
        return tree

    def _sample_tree(self, data_idx, tree, old_node):
        """gibbs_mh.py:35-70.  The reference copies the tree once per candidate node; here every candidate is
        scored from the ONE tree by recording the tile ops its two path refreshes would perform (remove from
        `old_node`: refresh old_node..root; add to `new_node`: log_r += value, refresh parent(new_node)..root),
        and only the candidate that is drawn gets built."""
        data_point = tree.data[data_idx]
        assert data_point.idx == data_idx
        pending = self._score_candidates(tree, data_point, old_node)
        log_q = np.array([p.resolve()[1] for p in pending])
        log_q = log_normalize(log_q)
        q = np.exp(log_q)
        q = q / sum(q)
        tree_idx = self._rng.multinomial(1, q).argmax()
        nodes = tree.nodes
        new_tree = tree.copy()
        new_tree.remove_data_point_from_node(data_point, old_node)
        if tree_idx < len(nodes):
            new_tree.add_data_point_to_node(data_point, nodes[tree_idx])
        else:
            new_tree.add_data_point_to_outliers(data_point)
        return new_tree

    def _score_candidates(self, tree, data_point, old_node):
        from ..smc.forest_state import StateScore, _priors

        st = tree.store
        g, lp, lr, idx = tree._graph, tree._lp, tree._lr, tree._node_indices
        root = idx[tree._ROOT_NODE_NAME]
        v = st.value_id[data_point.idx]
        succ_cache = {}

        def succ(i, _get=succ_cache.get, _raw=g.successor_indices):  # the same node is asked for by every candidate below it
            r = _get(i)
            if r is None:
                r = succ_cache[i] = _raw(i)
            return r

        pred = g.predecessor_indices
        crp, mult, size, oprior, omarg = tree.prior_terms()
        n_nodes = tree.get_number_of_nodes()
        roots = succ(root)
        sizes = [size[i] for i in roots]
        prior = self.tree_dist.prior
        over = {}
        if old_node == tree._OUTLIER_NODE_NAME:
            # the point is currently an outlier: removing it touches no tile (tree/tree.py:446-454)
            oi, lp_old = -1, None
            crp_removed = crp
            if data_point.outlier_prob != 0:
                oprior = oprior - data_point.outlier_prob + data_point.outlier_prob_not
            omarg = omarg - data_point.outlier_marginal_prob
        else:
            # ---- removal: new log_p of old_node, refreshed log_r along old_node..root
            oi = idx[old_node]
            lp_old = st.sub(lp[oi], v)
            i, first = oi, True
            while i != root:
                kids = tuple(over.get(c, lr[c]) for c in succ(i))
                base = lp_old if first else lp[i]
                over[i] = st.node(base, kids) if kids else base
                first = False
                i = pred(i)[0]
            n_old = len(tree._data[old_node])
            crp_removed = crp - (lf(n_old - 1) - lf(n_old - 2))
        out = []
        for new_node in tree.nodes:
            ni = idx[new_node]
            m = len(tree._data[new_node]) - (1 if ni == oi else 0)  # points in the node once the point is removed
            over2 = {ni: st.add(over.get(ni, lr[ni]), v)}
            i = pred(ni)[0]
            while i != root:
                kids = tuple(over2.get(c, over.get(c, lr[c])) for c in succ(i))
                over2[i] = st.node(lp_old if i == oi else lp[i], kids)
                i = pred(i)[0]
            kids = tuple(over2.get(c, over.get(c, lr[c])) for c in roots)
            c2 = crp_removed + (lf(m) - lf(m - 1) if m >= 1 else 0.0)
            p, p1 = _priors(prior, n_nodes, c2, mult, sizes, oprior, omarg)
            out.append(StateScore(p, p1, st.score(kids), st))
        if self.outliers:
            kids = tuple(over.get(c, lr[c]) for c in roots)
            op = oprior
            if data_point.outlier_prob != 0:
                op = oprior - data_point.outlier_prob_not + data_point.outlier_prob
            p, p1 = _priors(prior, n_nodes, crp_removed, mult, sizes, op, omarg + data_point.outlier_marginal_prob)
            out.append(StateScore(p, p1, st.score(kids) if kids else None, st))
        return out


class PruneRegraphSampler(object):
    def __init__(self, tree_dist, rng):
        self.tree_dist = tree_dist
        self._rng = rng

    def sample_tree(self, tree):
        if tree.get_number_of_nodes() <= 1:
            return tree
        pruned_tree = tree.copy()
        subtree_root = self._rng.choice(pruned_tree.nodes)
        subtree = pruned_tree.get_subtree(subtree_root)
        pruned_tree.remove_subtree(subtree)
        remaining_nodes = pruned_tree.nodes
        if len(remaining_nodes) == 0:
            return tree
        remaining_nodes.append(None)
        trees = []
        for parent in remaining_nodes:
            new_tree = pruned_tree.copy()
            nc = new_tree.get_number_of_children(new_tree.root_node_name if parent is None else parent)
            new_tree.add_subtree(subtree, parent=parent)
            new_tree.update()
            trees.append((nc, new_tree))
        pending = [self.tree_dist.prepare(x) for _, x in trees]
        log_p = np.array([np.log(n + 1) + p.resolve()[1] for (n, _), p in zip(trees, pending)])
        p, _ = exp_normalize(log_p)
        idx = self._rng.multinomial(1, p).argmax()
        return trees[idx][1]
