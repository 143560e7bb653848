"""Pure-Python stand-in for the subset of `rustworkx` (>= 0.15.1) PhyClone touches.

TEST INFRASTRUCTURE ONLY.  rustworkx is a third-party Rust wheel that is not
installed in this image and cannot be fetched (no network).  The reference
(`/root/reference/phyclone/tree/tree.py:5`) needs it only as a topology
container, but the *order* in which it enumerates successors / nodes / edges
fixes the child fold order and candidate order of the sampler (SURVEY.md
Appendix C).  This module restates petgraph's `StableGraph` bookkeeping
(published algorithm: slot vectors for nodes and edges, LIFO free lists,
per-node singly linked adjacency lists with head insertion) so that the
unmodified reference can run in this container and generate golden vectors.

Ordering parity against the real wheel is UNPINNED (it cannot be imported
here); every fixture generated through this shim says so.  Nothing under
`phyclone_b200/` imports this file.
"""

import json

from . import visit  # noqa: F401

_END = -1


class _Graph:
    """Slot-vector graph with petgraph `StableGraph` ordering semantics."""

    directed = True

    def __init__(self):
        # node slot: [payload, head_out, head_in]; vacant: [_VACANT, next_free, _END]
        self._nw = []  # node payloads (or _VACANT)
        self._nout = []  # head of outgoing edge list per node
        self._nin = []  # head of incoming edge list per node
        self._nfree_next = []  # free-list link per node slot
        self._nlive = []
        self._free_node = _END
        self._node_count = 0
        # edge slot
        self._esrc = []
        self._edst = []
        self._ew = []
        self._enext_out = []
        self._enext_in = []
        self._elive = []
        self._efree_next = []
        self._free_edge = _END
        self._edge_count = 0

    # ------------------------------------------------------------------ nodes
    def add_node(self, obj):
        if self._free_node != _END:
            idx = self._free_node
            self._free_node = self._nfree_next[idx]
            self._nw[idx] = obj
            self._nlive[idx] = True
            self._nout[idx] = _END
            self._nin[idx] = _END
        else:
            idx = len(self._nw)
            self._nw.append(obj)
            self._nout.append(_END)
            self._nin.append(_END)
            self._nfree_next.append(_END)
            self._nlive.append(True)
        self._node_count += 1
        return idx

    def _check_node(self, idx):
        if idx < 0 or idx >= len(self._nw) or not self._nlive[idx]:
            raise IndexError("No node found for index {}".format(idx))

    def __getitem__(self, idx):
        self._check_node(idx)
        return self._nw[idx]

    def __setitem__(self, idx, obj):
        self._check_node(idx)
        self._nw[idx] = obj

    def __len__(self):
        return self._node_count

    def num_nodes(self):
        return self._node_count

    def num_edges(self):
        return self._edge_count

    def node_indices(self):
        live = self._nlive
        return [i for i in range(len(self._nw)) if live[i]]

    node_indexes = node_indices

    def nodes(self):
        live = self._nlive
        return [self._nw[i] for i in range(len(self._nw)) if live[i]]

    def remove_node(self, idx):
        live = self._nlive
        if idx < 0 or idx >= len(self._nw) or not live[idx]:
            return
        # petgraph: drop outgoing edges first (list head first), then incoming
        while self._nout[idx] != _END:
            self._remove_edge_slot(self._nout[idx])
        while self._nin[idx] != _END:
            self._remove_edge_slot(self._nin[idx])
        live[idx] = False
        self._nw[idx] = None
        self._nfree_next[idx] = self._free_node
        self._free_node = idx
        self._node_count -= 1

    def remove_nodes_from(self, index_list):
        for idx in index_list:
            self.remove_node(idx)

    # ------------------------------------------------------------------ edges
    def add_edge(self, a, b, weight):
        self._check_node(a)
        self._check_node(b)
        if self._free_edge != _END:
            e = self._free_edge
            self._free_edge = self._efree_next[e]
            self._esrc[e] = a
            self._edst[e] = b
            self._ew[e] = weight
            self._elive[e] = True
        else:
            e = len(self._esrc)
            self._esrc.append(a)
            self._edst.append(b)
            self._ew.append(weight)
            self._enext_out.append(_END)
            self._enext_in.append(_END)
            self._elive.append(True)
            self._efree_next.append(_END)
        # head insertion into both adjacency lists
        self._enext_out[e] = self._nout[a]
        self._nout[a] = e
        self._enext_in[e] = self._nin[b]
        self._nin[b] = e
        self._edge_count += 1
        return e

    def _remove_edge_slot(self, e):
        a, b = self._esrc[e], self._edst[e]
        # unlink from a's outgoing list
        cur, prev = self._nout[a], _END
        while cur != e:
            prev, cur = cur, self._enext_out[cur]
        if prev == _END:
            self._nout[a] = self._enext_out[e]
        else:
            self._enext_out[prev] = self._enext_out[e]
        # unlink from b's incoming list
        cur, prev = self._nin[b], _END
        while cur != e:
            prev, cur = cur, self._enext_in[cur]
        if prev == _END:
            self._nin[b] = self._enext_in[e]
        else:
            self._enext_in[prev] = self._enext_in[e]
        self._elive[e] = False
        self._ew[e] = None
        self._efree_next[e] = self._free_edge
        self._free_edge = e
        self._edge_count -= 1

    def _out_edges(self, idx):
        e = self._nout[idx]
        while e != _END:
            yield e
            e = self._enext_out[e]

    def _in_edges(self, idx):
        e = self._nin[idx]
        while e != _END:
            yield e
            e = self._enext_in[e]

    def remove_edge(self, a, b):
        for e in self._out_edges(a):
            if self._edst[e] == b:
                self._remove_edge_slot(e)
                return
        raise NoEdgeBetweenNodes("No edge found between nodes")

    def has_edge(self, a, b):
        return any(self._edst[e] == b for e in self._out_edges(a))

    def edge_list(self):
        return [(self._esrc[e], self._edst[e]) for e in range(len(self._esrc)) if self._elive[e]]

    def weighted_edge_list(self):
        return [(self._esrc[e], self._edst[e], self._ew[e]) for e in range(len(self._esrc)) if self._elive[e]]

    def extend_from_edge_list(self, edge_list):
        for s, t in edge_list:
            top = max(s, t)
            while top >= self._node_count:
                self.add_node(None)
            self.add_edge(s, t, None)

    # -------------------------------------------------------------- adjacency
    def _succ_indices(self, idx):
        self._check_node(idx)
        seen, out = set(), []
        for e in self._out_edges(idx):
            t = self._edst[e]
            if t not in seen:
                seen.add(t)
                out.append(t)
        return out

    def _pred_indices(self, idx):
        self._check_node(idx)
        seen, out = set(), []
        for e in self._in_edges(idx):
            s = self._esrc[e]
            if s not in seen:
                seen.add(s)
                out.append(s)
        return out

    def successors(self, idx):
        return [self._nw[t] for t in self._succ_indices(idx)]

    def predecessors(self, idx):
        return [self._nw[s] for s in self._pred_indices(idx)]

    def successor_indices(self, idx):
        return self._succ_indices(idx)

    def predecessor_indices(self, idx):
        return self._pred_indices(idx)

    def neighbors(self, idx):
        return self._succ_indices(idx)

    def out_degree(self, idx):
        self._check_node(idx)
        return sum(1 for _ in self._out_edges(idx))

    def in_degree(self, idx):
        self._check_node(idx)
        return sum(1 for _ in self._in_edges(idx))

    # ------------------------------------------------------------ whole-graph
    def copy(self):
        new = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            new.__dict__[k] = list(v) if isinstance(v, list) else v
        return new

    def compose(self, other, node_map):
        mapping = {}
        for idx in other.node_indices():
            mapping[idx] = self.add_node(other._nw[idx])
        for e in range(len(other._esrc)):
            if other._elive[e]:
                self.add_edge(mapping[other._esrc[e]], mapping[other._edst[e]], other._ew[e])
        for this_idx, (other_idx, weight) in node_map.items():
            self.add_edge(this_idx, mapping[other_idx], weight)
        return mapping

    def remove_node_retain_edges(self, idx, use_outgoing=False, condition=None):
        self._check_node(idx)
        in_edges = [(self._esrc[e], self._ew[e]) for e in self._in_edges(idx)]
        out_edges = [(self._edst[e], self._ew[e]) for e in self._out_edges(idx)]
        for s, w_in in in_edges:
            for t, w_out in out_edges:
                self.add_edge(s, t, w_out if use_outgoing else w_in)
        self.remove_node(idx)

    def subgraph(self, nodes, preserve_attrs=False):
        keep = sorted(set(nodes))
        new = self.__class__()
        remap = {}
        for idx in keep:
            self._check_node(idx)
            remap[idx] = new.add_node(self._nw[idx])
        for e in range(len(self._esrc)):
            if self._elive[e] and self._esrc[e] in remap and self._edst[e] in remap:
                new.add_edge(remap[self._esrc[e]], remap[self._edst[e]], self._ew[e])
        return new


class PyDiGraph(_Graph):
    def __init__(self, check_cycle=False, multigraph=True, attrs=None, *, node_count_hint=None, edge_count_hint=None):
        super().__init__()
        self.attrs = attrs


class PyGraph(_Graph):
    """Undirected flavour; PhyClone only uses it in an isinstance check."""

    directed = False


class NoEdgeBetweenNodes(Exception):
    pass


# ---------------------------------------------------------------- algorithms
def dfs_search(graph, source, visitor):
    """Iterative DFS with the event order of rustworkx-core `depth_first_search`."""
    if source is None:
        source = graph.node_indices()
    discovered, finished = set(), set()
    time = 0
    for start in source:
        if start in discovered:
            continue
        discovered.add(start)
        visitor.discover_vertex(start, time)
        time += 1
        stack = [(start, iter(_out_targets(graph, start)))]
        while stack:
            u, it = stack[-1]
            advanced = False
            for v, w in it:
                if v not in discovered:
                    visitor.tree_edge((u, v, w))
                    discovered.add(v)
                    visitor.discover_vertex(v, time)
                    time += 1
                    stack.append((v, iter(_out_targets(graph, v))))
                    advanced = True
                    break
                elif v not in finished:
                    visitor.back_edge((u, v, w))
                else:
                    visitor.forward_or_cross_edge((u, v, w))
            if not advanced:
                stack.pop()
                finished.add(u)
                visitor.finish_vertex(u, time)
                time += 1


def _out_targets(graph, idx):
    return [(graph._edst[e], graph._ew[e]) for e in graph._out_edges(idx)]


def descendants(graph, node):
    graph._check_node(node)
    seen = []
    mark = {node}
    stack = [node]
    while stack:
        u = stack.pop()
        for v in graph._succ_indices(u):
            if v not in mark:
                mark.add(v)
                seen.append(v)
                stack.append(v)
    return set(seen)


def ancestors(graph, node):
    graph._check_node(node)
    mark = {node}
    out = []
    stack = [node]
    while stack:
        u = stack.pop()
        for v in graph._pred_indices(u):
            if v not in mark:
                mark.add(v)
                out.append(v)
                stack.append(v)
    return set(out)


def all_simple_paths(graph, origin, to, min_depth=None, cutoff=None):
    graph._check_node(origin)
    graph._check_node(to)
    paths = []
    if origin == to:
        return paths
    path = [origin]
    on_path = {origin}

    def rec(u):
        for v in graph._succ_indices(u):
            if v in on_path:
                continue
            if v == to:
                paths.append(path + [v])
                continue
            path.append(v)
            on_path.add(v)
            rec(v)
            path.pop()
            on_path.discard(v)

    rec(origin)
    return paths


def _forest_signature(graph):
    """Canonical form of a rooted forest given as a directed graph (AHU encoding)."""
    sig = {}
    order = []
    roots = [i for i in graph.node_indices() if graph.in_degree(i) == 0]
    stack = list(roots)
    while stack:
        u = stack.pop()
        order.append(u)
        stack.extend(graph._succ_indices(u))
    if len(order) != graph.num_nodes():
        return None
    for u in reversed(order):
        sig[u] = "(" + "".join(sorted(sig[v] for v in graph._succ_indices(u))) + ")"
    return sorted(sig[r] for r in roots)


def is_isomorphic(first, second, node_matcher=None, edge_matcher=None, id_order=True, call_limit=None):
    if first.num_nodes() != second.num_nodes() or first.num_edges() != second.num_edges():
        return False
    a, b = _forest_signature(first), _forest_signature(second)
    if a is None or b is None:
        raise NotImplementedError("refshim.is_isomorphic only handles rooted forests")
    return a == b


def node_link_json(graph, path=None, graph_attrs=None, node_attrs=None, edge_attrs=None):
    nodes = []
    for idx in graph.node_indices():
        data = node_attrs(graph._nw[idx]) if node_attrs is not None else None
        nodes.append({"id": idx, "data": data})
    links = []
    for e in range(len(graph._esrc)):
        if graph._elive[e]:
            data = edge_attrs(graph._ew[e]) if edge_attrs is not None else None
            links.append({"source": graph._esrc[e], "target": graph._edst[e], "id": e, "data": data})
    doc = {
        "directed": graph.directed,
        "multigraph": True,
        "attrs": graph_attrs(graph.attrs) if graph_attrs is not None else None,
        "nodes": nodes,
        "links": links,
    }
    text = json.dumps(doc)
    if path is not None:
        with open(path, "w") as fh:
            fh.write(text)
        return None
    return text
