"""Data-point Gibbs and prune-regraph moves (reference mcmc/gibbs_mh.py:6-129).

Every candidate tree of a move is built first (host bookkeeping + recorded tile ops) and the
candidate scores are read afterwards, so one Gibbs step is one batched device evaluation.
"""

import numpy as np

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
        data_point = tree.data[data_idx]
        assert data_point.idx == data_idx
        new_trees = []
        for new_node in tree.nodes:
            new_tree = tree.copy()
            new_tree.remove_data_point_from_node(data_point, old_node)
            new_tree.add_data_point_to_node(data_point, new_node)
            new_trees.append(new_tree)
        if self.outliers:
            new_tree = tree.copy()
            new_tree.remove_data_point_from_node(data_point, old_node)
            new_tree.add_data_point_to_outliers(data_point)
            new_trees.append(new_tree)
        pending = [self.tree_dist.prepare(x) for x in new_trees]
        log_q = np.array([p.resolve()[1] for p in pending])
        log_q = log_normalize(log_q)
        q = np.exp(log_q)
        q = q / sum(q)
        tree_idx = self._rng.multinomial(1, q).argmax()
        return new_trees[tree_idx]


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
