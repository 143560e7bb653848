"""Consensus tree of a trace (reference process_trace/consensus.py:8-154): clade frequencies (or weights) over the
sampled trees, majority clades, the containment forest they form, and its conversion into a `Tree` whose MAP
prevalences are then assigned on the device (process_trace/map.py).

Pure set algebra on the host; no tile arithmetic happens here.  The forest is a small insertion-ordered digraph
(`_Forest`): node and edge insertion order decide the node numbering and the child order of the result exactly as the
networkx graph does in the reference, so the outputs can be compared file against file.
"""

from collections import defaultdict


class _Forest(object):
    """Insertion-ordered digraph: {node: [children...]} plus parent counts (what the reference uses of nx.DiGraph)."""

    def __init__(self):
        self.children = {}
        self.n_parents = {}

    def add_node(self, n):
        if n not in self.children:
            self.children[n] = []
            self.n_parents[n] = 0

    def add_edge(self, a, b):
        self.add_node(a)
        self.add_node(b)
        if b not in self.children[a]:
            self.children[a].append(b)
            self.n_parents[b] += 1

    def roots(self):
        return [n for n in self.children if self.n_parents[n] == 0]

    def preorder(self):
        """Depth-first pre-order over all nodes, sources and children in insertion order (nx.dfs_preorder_nodes)."""
        seen, order = set(), []
        for source in self.children:
            if source in seen:
                continue
            stack = [source]
            while stack:
                n = stack.pop()
                if n in seen:
                    continue
                seen.add(n)
                order.append(n)
                stack.extend(reversed([c for c in self.children[n] if c not in seen]))
        return order


def clade_weights(trees, weights=None):
    """consensus.py:85-102: share of trees containing each clade, or the sum of the given tree weights."""
    acc = defaultdict(float)
    for i, tree in enumerate(trees):
        for clade in tree.get_clades():
            acc[clade] += 1 if weights is None else weights[i]
    if weights is None:
        for clade in acc:
            acc[clade] = acc[clade] / len(trees)
    return acc


def smallest_superset(candidates, query):
    """consensus.py:128-154: the smallest strict superset of `query`; two of equal size mean inconsistent clades."""
    best, best_size = None, float("inf")
    for cand in candidates:
        if cand == query or not cand.issuperset(query):
            continue
        if len(cand) == best_size:
            raise Exception("Inconsistent set of clades")
        if len(cand) < best_size:
            best, best_size = cand, len(cand)
    return best


def containment_forest(clades):
    """consensus.py:23-38: every clade hangs under its smallest superset."""
    forest = _Forest()
    for clade in clades:
        parent = smallest_superset(clades, clade)
        if parent is None:
            forest.add_node(clade)
        else:
            forest.add_edge(parent, clade)
    return forest


def own_mutations(forest):
    """consensus.py:41-55, 105-121: each node keeps the data points of its clade that no child clade holds."""
    out = _Forest()

    def visit(clade):
        mine = set(clade)
        for child in forest.children[clade]:
            mine.difference_update(child)
        mine = frozenset(mine)
        out.add_node(mine)
        for child in forest.children[clade]:
            out.add_edge(mine, visit(child))
        return mine

    for root in forest.roots():
        visit(root)
    return out


def get_consensus_tree(trees, data=None, threshold=0.5, weighted=False, log_p_list=None):
    """consensus.py:8-20.  Returns (children: {node id: [child ids]}, idxs: {node id: sorted data idxs}, names)."""
    weights = clade_weights(trees, log_p_list if weighted else None)
    chosen = set([clade for clade, w in weights.items() if w > threshold])
    forest = own_mutations(containment_forest(chosen))
    number = {node: k for k, node in enumerate(forest.preorder())}
    children = {number[n]: [number[c] for c in forest.children[n]] for n in forest.children}
    idxs = {number[n]: sorted(n) for n in forest.children}
    names = None
    if data is not None:
        names = {k: [data[i].name for i in v] for k, v in idxs.items()}
    return children, idxs, names


def tree_from_consensus(data, children, idxs, store):
    """process_trace.py:236-291 (get_tree_from_consensus_graph + from_dict_nx): nodes in numbering order, the parentless
    ones under the dummy root, data points attached in label order, then a full refresh."""
    from ..tree.tree import Tree

    tree = Tree(data[0].grid_size, store)
    by_idx = {x.idx: x for x in data}
    labels = {}
    for node in children:
        for idx in idxs[node]:
            labels[idx] = node
    for x in data:
        if x.idx not in labels:
            labels[x.idx] = tree.outlier_node_name
    has_parent = {c for kids in children.values() for c in kids}
    edges = {node: list(kids) for node, kids in children.items()}
    edges[tree.root_node_name] = [node for node in children if node not in has_parent]
    for node in edges:
        if node != tree.root_node_name:
            tree._add_node(node)
    for parent, kids in edges.items():
        for child in kids:
            tree._graph.add_edge(tree._node_indices[parent], tree._node_indices[child])
    for idx, node in labels.items():
        tree._attach_without_refresh(by_idx[idx], node)
    tree.update()
    return tree
