"""Data-point Gibbs and prune-regraph moves (reference mcmc/gibbs_mh.py:6-129).

Every candidate tree of a move is built first (host bookkeeping + recorded tile ops) and the
candidate scores are read afterwards, so one Gibbs step is one batched device evaluation.
"""

import os

import numpy as np

from .._hostmod import load as _load_hostmod
from ..utils.math import exp_normalize, log_normalize

_hm = _load_hostmod()


class DataPointSampler(object):
    # moves scored per device round trip (see sample_tree); any value gives the same chain
    LOOKAHEAD = int(os.environ.get("PCB_GIBBS_LOOKAHEAD", "8"))

    def __init__(self, tree_dist, rng, outliers=False):
        self.tree_dist = tree_dist
        self.outliers = outliers
        self._rng = rng
        self._flat = None  # (graph, flattened structure) of the tree the current sweep edits in place
        self._one_shape = False  # True while sample_tree runs: all its trees are copies of one shape
        self.moves = self.stays = self.lookahead_hits = 0

    def sample_tree(self, tree):
        """gibbs_mh.py:17-33: one Gibbs move per data point whose node holds more than one point, in shuffled order.
        The reference copies the tree for every move; here the input is copied ONCE and the moves edit that copy in
        place (nothing else holds it), the label map is updated instead of rebuilt, and data points are looked up in
        a map built once -- at 2000 unclustered mutations the per-move rebuilds were O(T) each."""
        tree_labels = tree.labels
        data_idxs = list(tree_labels.keys())
        self._rng.shuffle(data_idxs)
        self._flat = None
        self._one_shape = True
        by_idx = None
        # A move is one synchronising device round trip, and most moves put the point back where it was.  So the
        # candidates of the next LOOKAHEAD moves are recorded at once, each against the tree its predecessors leave
        # behind IF they stay put (tile ops are lazy and memoised: recording costs no arithmetic, and "remove + add
        # back" yields the same tile ids on the look-ahead copy as later on the real tree); one flush scores them
        # all.  The moves are then drawn one by one exactly as before -- same scores, same RNG calls -- and the first
        # point that really moves throws the rest of the look-ahead away.
        ahead = []  # [(position in data_idxs, pending scores | None, the tree after this point stays put | None)]
        k, n = 0, len(data_idxs)
        while k < n:
            if not ahead:
                t = tree
                for j in range(k, min(k + self.LOOKAHEAD, n)):
                    old = tree_labels[data_idxs[j]]
                    if t.get_data_len(old) <= 1:
                        ahead.append((j, None, None))
                        continue
                    if by_idx is None:
                        t = tree = tree.copy()
                        by_idx = {dp.idx: dp for lst in tree._data.values() for dp in lst}
                    dp = by_idx[data_idxs[j]]
                    pending = self._score_candidates(t, dp, old)
                    stay = None
                    if j + 1 < min(k + self.LOOKAHEAD, n):
                        t = stay = t.copy()
                        t.remove_data_point_from_node(dp, old)
                        if old == t.outlier_node_name:
                            t.add_data_point_to_outliers(dp)
                        else:
                            t.add_data_point_to_node(dp, old)
                    ahead.append((j, pending, stay))
            j, pending, stay = ahead.pop(0)
            assert j == k
            data_idx = data_idxs[k]
            if pending is not None:
                old_node = tree_labels[data_idx]
                # a point that stays put: the look-ahead already built exactly that tree (the next points were
                # scored on it) -- adopt it instead of applying "remove + add back" a second time
                new_node, tree = self._move(by_idx[data_idx], tree, old_node, pending, stay)
                tree_labels[data_idx] = new_node
                self.moves += 1
                if new_node == old_node:
                    self.stays += 1
                    self.lookahead_hits += 1 if ahead else 0
                else:
                    ahead = []
            k += 1
        self._one_shape = False
        self._flat = None
        return tree

    def _sample_tree(self, data_idx, tree, old_node):
        """gibbs_mh.py:35-70 with the reference's signature: returns a new tree, the input is left alone."""
        data_point = tree.data[data_idx]
        assert data_point.idx == data_idx
        self._one_shape = False
        new_tree = tree.copy()
        return self._move(data_point, new_tree, old_node)[1]

    def _move(self, data_point, tree, old_node, pending=None, stay=None):
        """Draw the node `data_point` moves to and apply the move to `tree` IN PLACE (or, if the point stays where it
        is and the caller already holds that tree as `stay`, take that one); returns (node name, tree).
        Every candidate is scored from the one tree by recording the tile ops its two path refreshes would perform
        (remove from `old_node`: refresh old_node..root; add to the new node: log_r += value, refresh its parents up
        to the root); only the candidate that is drawn gets built."""
        if pending is None:
            pending = self._score_candidates(tree, data_point, old_node)
        log_q = np.array([p.resolve()[1] for p in pending])
        log_q = log_normalize(log_q)
        q = np.exp(log_q)
        q = q / sum(q)
        tree_idx = self._rng.multinomial(1, q).argmax()
        nodes = tree.nodes
        new_node = nodes[tree_idx] if tree_idx < len(nodes) else tree.outlier_node_name
        if stay is not None and new_node == old_node:
            return new_node, stay
        tree.remove_data_point_from_node(data_point, old_node)
        if tree_idx < len(nodes):
            tree.add_data_point_to_node(data_point, new_node)
        else:
            tree.add_data_point_to_outliers(data_point)
        return new_node, tree

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
        # the shape of the tree does not change during a Gibbs sweep (points move, nodes stay): flatten it once
        flat = self._flat
        # inside one sweep every tree handed in (the tree being edited and its look-ahead copies) has the same shape
        if flat is None or (flat[0] is not g and not self._one_shape):
            n_slots = len(g.payload)
            parent, kids = [-1] * n_slots, [()] * n_slots
            succ, pred = g.successor_indices, g.predecessor_indices
            live = list(g.node_indices())
            for i in live:
                kids[i] = succ(i)
                p = pred(i)
                if p:
                    parent[i] = p[0]
            root = idx[tree._ROOT_NODE_NAME]
            nodes = tree.nodes
            flat = self._flat = (g, n_slots, parent, kids, live, root, nodes, [idx[n] for n in nodes])
        _, n_slots, parent, kids, live, root, nodes, cand = flat
        lp_l, lr_l = [-1] * n_slots, [-1] * n_slots
        for i in live:
            lp_l[i], lr_l[i] = lp[i], lr[i]
        crp, mult, size, oprior, omarg = tree.prior_terms()
        sizes = [size[i] for i in kids[root]]
        data = tree._data
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
            # the reference refreshes the whole tree here (gibbs_mh.py:108); add_subtree has already refreshed the
            # path from `parent` to the root and every other node holds the tile a refresh would look up again
            # (node tiles are memoised by log_p tile and child tiles), so the full walk is skipped
            trees.append((nc, new_tree))
        pending = [self.tree_dist.prepare(x) for _, x in trees]
        log_p = np.array([np.log(n + 1) + p.resolve()[1] for (n, _), p in zip(trees, pending)])
        p, _ = exp_normalize(log_p)
        idx = self._rng.multinomial(1, p).argmax()
        return trees[idx][1]
