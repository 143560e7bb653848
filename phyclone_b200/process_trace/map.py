"""MAP cellular prevalences / clonal prevalences of a tree (reference process_trace/map.py:5-173).

The reference converts the tree to networkx and runs a pure-Python O(S*G^2) max-plus fold per child with arg-max
back-pointers; here the tree structure is flattened on the host and `pcb_map_trees_dev` walks every
(tree, sample) pair of a batch in one launch, reading the node log_p tiles where they already are (the arena).
Children are folded in ascending edge-slot order: the successor order of the networkx graph the reference builds
from `weighted_edge_list()` (process_trace/utils.py:5-20).
"""

import numpy as np


def _flatten(tree, base):
    """(graph indices, ordered children as global node numbers, post-order as global node numbers) of one tree;
    node numbers start at `base`, the dummy root gets the last post-order position."""
    g = tree._graph
    esrc, edst = g.esrc, g.edst
    kids = {}
    for e in range(len(esrc)):
        if esrc[e] >= 0:
            kids.setdefault(esrc[e], []).append(edst[e])
    root = tree._node_indices[tree._ROOT_NODE_NAME]
    gidx, local = [], {}
    stack = [root]
    while stack:  # numbering: any order covering the root's component
        i = stack.pop()
        local[i] = base + len(gidx)
        gidx.append(i)
        stack.extend(kids.get(i, ()))
    post, stack = [], [(root, False)]
    while stack:
        i, seen = stack.pop()
        if seen:
            post.append(local[i])
        else:
            stack.append((i, True))
            stack.extend((c, False) for c in kids.get(i, ()))
    children = [[local[c] for c in kids.get(i, ())] for i in gidx]
    return gidx, children, post


def map_trees(trees):
    """[(ccf dict, clonal prevalence dict)] for trees sharing one tile store; one device launch for the batch."""
    if not trees:
        return []
    store = trees[0].store
    lp_ids, tree_off, post_order, child_off, child_idx, spans = [], [0], [], [0], [], []
    for tree in trees:
        if tree.store is not store:
            raise ValueError("map_trees: all trees must live in the same tile store")
        gidx, children, post = _flatten(tree, len(lp_ids))
        spans.append((len(lp_ids), gidx, children))
        lp_ids.extend(tree._lp[i] for i in gidx)
        post_order.extend(post)
        tree_off.append(len(post_order))
        for c in children:
            child_idx.extend(c)
            child_off.append(len(child_idx))
    store.flush()
    max_idx = store.arena.map_trees(lp_ids, tree_off, post_order, child_off, child_idx)
    G = store.G
    out = []
    for tree, (base, gidx, children) in zip(trees, spans):
        rev = tree._node_indices_rev
        ccf = max_idx[base:base + len(gidx)] / (G - 1)  # map.py:169-173
        ccfs, prevs = {}, {}
        for k, i in enumerate(gidx):
            name = rev[i]
            if name == tree._ROOT_NODE_NAME:
                continue
            prev = ccf[k].copy()  # map.py:36-45
            for c in children[k]:
                prev -= ccf[c - base]
            ccfs[name], prevs[name] = ccf[k].copy(), prev
        out.append((ccfs, prevs))
    return out


def get_map_node_ccfs_and_clonal_prev_dicts(tree):
    """map.py:5-16"""
    return map_trees([tree])[0]
