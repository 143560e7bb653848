"""Slot-vector digraph with the enumeration order of petgraph's `StableGraph`.

The reference keeps its forest in a `rustworkx.PyDiGraph` (tree/tree.py:37); rustworkx does no
arithmetic but its enumeration order decides the child fold order of every node refresh and the
candidate order of every proposal, so the host side of the hot path needs a
container with the same ordering contract:

  * node / edge slots are recycled LIFO; new slots are appended;
  * a new edge goes to the HEAD of its source's out-list and its target's in-list, so successors
    are visited most-recent-edge first; removing an edge keeps the relative order of the others;
  * `nodes()` / `node_indices()` / `edge_list()` run in slot order and skip vacancies.

Designed for cheap copies (a particle filter copies trees far more often than it edits them):
adjacency lists are immutable tuples shared between copies and replaced on edit.
"""

VACANT = object()


class SlotDiGraph:
    __slots__ = ("payload", "out", "inn", "free_nodes", "esrc", "edst", "free_edges", "n_nodes", "n_edges")

    def __init__(self):
        self.payload = []  # per node slot: payload or VACANT
        self.out = []  # per node slot: tuple of edge ids, most recent first
        self.inn = []
        self.free_nodes = []  # LIFO
        self.esrc = []  # per edge slot: source or -1 (vacant)
        self.edst = []
        self.free_edges = []  # LIFO
        self.n_nodes = 0
        self.n_edges = 0

    def copy(self):
        g = SlotDiGraph.__new__(SlotDiGraph)
        g.payload = self.payload.copy()
        g.out = self.out.copy()
        g.inn = self.inn.copy()
        g.free_nodes = self.free_nodes.copy()
        g.esrc = self.esrc.copy()
        g.edst = self.edst.copy()
        g.free_edges = self.free_edges.copy()
        g.n_nodes = self.n_nodes
        g.n_edges = self.n_edges
        return g

    # ------------------------------------------------------------------ nodes
    def add_node(self, obj):
        if self.free_nodes:
            i = self.free_nodes.pop()
            self.payload[i] = obj
            self.out[i] = ()
            self.inn[i] = ()
        else:
            i = len(self.payload)
            self.payload.append(obj)
            self.out.append(())
            self.inn.append(())
        self.n_nodes += 1
        return i

    def remove_node(self, i):
        if i < 0 or i >= len(self.payload) or self.payload[i] is VACANT:
            return
        for e in self.out[i]:  # outgoing first, head first; then incoming
            self._remove_edge(e)
        for e in self.inn[i]:
            self._remove_edge(e)
        self.payload[i] = VACANT
        self.free_nodes.append(i)
        self.n_nodes -= 1

    def remove_nodes_from(self, idxs):
        for i in idxs:
            self.remove_node(i)

    def node_indices(self):
        p = self.payload
        return [i for i in range(len(p)) if p[i] is not VACANT]

    def nodes(self):
        return [x for x in self.payload if x is not VACANT]

    def num_nodes(self):
        return self.n_nodes

    def __getitem__(self, i):
        x = self.payload[i]
        if x is VACANT:
            raise IndexError(i)
        return x

    def __setitem__(self, i, obj):
        if self.payload[i] is VACANT:
            raise IndexError(i)
        self.payload[i] = obj

    # ------------------------------------------------------------------ edges
    def add_edge(self, a, b):
        if self.free_edges:
            e = self.free_edges.pop()
            self.esrc[e] = a
            self.edst[e] = b
        else:
            e = len(self.esrc)
            self.esrc.append(a)
            self.edst.append(b)
        self.out[a] = (e,) + self.out[a]
        self.inn[b] = (e,) + self.inn[b]
        self.n_edges += 1
        return e

    def _remove_edge(self, e):
        a, b = self.esrc[e], self.edst[e]
        self.out[a] = tuple(x for x in self.out[a] if x != e)
        self.inn[b] = tuple(x for x in self.inn[b] if x != e)
        self.esrc[e] = -1
        self.edst[e] = -1
        self.free_edges.append(e)
        self.n_edges -= 1

    def remove_edge(self, a, b):
        edst = self.edst
        for e in self.out[a]:
            if edst[e] == b:
                self._remove_edge(e)
                return
        raise KeyError("no edge %r -> %r" % (a, b))

    def edge_list(self):
        s, d = self.esrc, self.edst
        return [(s[e], d[e]) for e in range(len(s)) if s[e] >= 0]

    def extend_from_edge_list(self, edges):
        for s, t in edges:
            top = s if s > t else t
            while top >= self.n_nodes:
                self.add_node(None)
            self.add_edge(s, t)

    # -------------------------------------------------------------- adjacency
    def successor_indices(self, i):
        edst = self.edst
        res = [edst[e] for e in self.out[i]]
        if len(res) > 1 and len(set(res)) != len(res):  # parallel edges: each successor once, first occurrence
            seen, uniq = set(), []
            for t in res:
                if t not in seen:
                    seen.add(t)
                    uniq.append(t)
            return uniq
        return res

    def predecessor_indices(self, i):
        esrc = self.esrc
        return [esrc[e] for e in self.inn[i]]

    def successors(self, i):
        p = self.payload
        return [p[t] for t in self.successor_indices(i)]

    def out_degree(self, i):
        return len(self.out[i])

    def descendants(self, i):
        """Set of node indices reachable from i (excluding i)."""
        seen = []
        mark = {i}
        stack = [i]
        while stack:
            u = stack.pop()
            for v in self.successor_indices(u):
                if v not in mark:
                    mark.add(v)
                    seen.append(v)
                    stack.append(v)
        return set(seen)

    # ------------------------------------------------------------ whole-graph
    def compose(self, other, node_map):
        """Add `other`'s nodes (slot order) and edges (slot order), then the `node_map` edges
        {this_idx: other_idx}; returns {other_idx: new_idx}."""
        mapping = {}
        for i in other.node_indices():
            mapping[i] = self.add_node(other.payload[i])
        s, d = other.esrc, other.edst
        for e in range(len(s)):
            if s[e] >= 0:
                self.add_edge(mapping[s[e]], mapping[d[e]])
        for this_idx, other_idx in node_map.items():
            self.add_edge(this_idx, mapping[other_idx])
        return mapping

    def remove_node_retain_edges(self, i):
        srcs = [self.esrc[e] for e in self.inn[i]]
        dsts = [self.edst[e] for e in self.out[i]]
        for s in srcs:
            for t in dsts:
                self.add_edge(s, t)
        self.remove_node(i)

    def subgraph(self, idxs):
        """Sub-graph on `idxs`, re-indexed compactly in ascending original index; returns
        (graph, {old_idx: new_idx})."""
        keep = sorted(set(idxs))
        g = SlotDiGraph()
        remap = {}
        for i in keep:
            remap[i] = g.add_node(self.payload[i])
        s, d = self.esrc, self.edst
        for e in range(len(s)):
            if s[e] >= 0 and s[e] in remap and d[e] in remap:
                g.add_edge(remap[s[e]], remap[d[e]])
        return g, remap
