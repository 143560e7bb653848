"""Data-point Gibbs and prune-regraph moves (reference mcmc/gibbs_mh.py:6-129).

Every candidate tree of a move is built first (host bookkeeping + recorded tile ops) and the
candidate scores are read afterwards, so one Gibbs step is one batched device evaluation.
"""

import numpy as np

from .._hostmod import load as _load_hostmod
from ..utils.math import exp_normalize, log_normalize

_hm = _load_hostmod()


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
        """Scores of `tree` with `data_point` moved to every node (in `tree.nodes` order) and, with outlier
        modelling on, to the outlier set.  The tree is flattened into per-slot arrays and `_pcbhost.gibbs_scores`
        records the tile ops of the two path refreshes of every candidate (removal from `old_node`, `log_r += value`
        on the new node, parents up to the root) and computes the prior terms from the running statistics."""
        from ..smc.forest_state import StateScore, ensure_tables, host_tuple

        st = tree.store
        g, lp, lr, idx = tree._graph, tree._lp, tree._lr, tree._node_indices
        prior = self.tree_dist.prior
        ensure_tables(prior, st.T + 2)
        n_slots = len(g.payload)
        parent, kids = [-1] * n_slots, [()] * n_slots
        lp_l, lr_l = [-1] * n_slots, [-1] * n_slots
        succ, pred = g.successor_indices, g.predecessor_indices
        for i in g.node_indices():
            kids[i] = succ(i)
            p = pred(i)
            if p:
                parent[i] = p[0]
            lp_l[i], lr_l[i] = lp[i], lr[i]
        root = idx[tree._ROOT_NODE_NAME]
        crp, mult, size, oprior, omarg = tree.prior_terms()
        sizes = [size[i] for i in kids[root]]
        nodes = tree.nodes
        data = tree._data
        cand = [idx[n] for n in nodes]
        cand_n = [len(data[n]) for n in nodes]
        if old_node == tree._OUTLIER_NODE_NAME:
            oi, n_old = -1, 0
        else:
            oi, n_old = idx[old_node], len(data[old_node])
        args = (st.book, parent, kids, lp_l, lr_l, root, cand, cand_n, oi, n_old, tree.get_number_of_nodes(), crp, mult,
                sizes, oprior, omarg, host_tuple(st, data_point), prior.log_alpha, bool(self.outliers))
        recs = _hm.gibbs_scores(*args)
        if recs is None:  # arena full: grow and record again (ops already recorded are memoised)
            st._grow()
            recs = _hm.gibbs_scores(*args)
        return [StateScore(p, p1, None if sid < 0 else sid, st) for p, p1, sid in recs]


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
