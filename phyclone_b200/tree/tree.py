"""Host-side forest with device-resident node tiles.

Drop-in for the reference's `phyclone.tree.Tree` (tree/tree.py:17-523): same public method names,
same candidate / child enumeration order (see `graph.SlotDiGraph`), same `to_dict()` schema
(SURVEY.md App. E).  What differs is where the numbers live: a node's `log_p` / `log_r` is a tile
id in the device arena (`tiles.TileStore`), every refresh is a recorded, memoised op, and nothing is
computed until a score is read -- so the candidate trees of a whole SMC step evaluate in a couple
of batched launches.  Copying a tree copies a few small lists and dicts, never a tile.
"""

from collections import defaultdict
from itertools import chain

from .._hostmod import load as _load_hostmod
from ..utils.math import cached_log_factorial
from .graph import SlotDiGraph

_edge_hash = _load_hostmod().edge_hash  # the structure hash is shared with the extension's ForestState
_data_hash = _load_hostmod().data_hash

NO_TILE = -1


class Tree(object):
    _ROOT_NODE_NAME = "root"
    _OUTLIER_NODE_NAME = -1
    __slots__ = ("grid_size", "store", "_data", "_graph", "_node_indices", "_node_indices_rev", "_last_node_added_to",
                 "_lp", "_lr", "_dirty", "_skey", "_pt", "_gen")

    def __init__(self, grid_size, store):
        self.grid_size = grid_size
        self.store = store
        # node name -> tuple of data points (insertion order).  Like the reference's defaultdict(list), a key
        # springs into existence on first access; the key ORDER feeds `labels` and thus the Gibbs sweep order.
        self._data = defaultdict(tuple)
        self._graph = SlotDiGraph()
        self._node_indices = {}
        self._node_indices_rev = {}
        self._last_node_added_to = None
        self._lp = {}  # graph index -> tile id of log_p
        self._lr = {}  # graph index -> tile id of log_r (NO_TILE until first refresh)
        # graph indices whose adjacency order / log_r a from_dict(to_dict()) round trip would change
        # (see thawed_copy); None = unknown, take the generic round trip
        self._dirty = set()
        self._skey = 0  # running structure hash (see structure_key); None = recompute
        # running integer/scalar statistics behind the FS-CRP prior (see prior_terms); None = recompute
        self._pt = [0.0, 0.0, {}, 0.0, 0.0]  # crp_sum, multiplicity, subtree node counts, outlier prior, outlier marginals
        # arena generation the tile ids belong to: a tree carried across `store.reset()` may still be read structurally
        # (labels, children, sigma), but recording tile ops on its ids would read slots the new generation reuses
        self._gen = getattr(store, "generation", 0)
        self._add_node(self._ROOT_NODE_NAME)

    # ------------------------------------------------------------------ identity
    def get_clades(self):
        """Set of clades (tree/tree.py:64-69, visitors.py:48-88)."""
        out = set()
        g, rev, data = self._graph, self._node_indices_rev, self._data
        root = self._ROOT_NODE_NAME

        def rec(i):
            name = rev[i]
            acc = {dp.idx for dp in data.get(name, ())}
            for c in g.successor_indices(i):
                acc |= rec(c)
            if name != root:
                out.add(frozenset(acc))
            return acc

        rec(self._node_indices[root])
        return frozenset(out)

    def __hash__(self):
        return hash((self.get_clades(), frozenset(self.outliers)))

    def __eq__(self, other):
        return (self.get_clades(), frozenset(self.outliers)) == (other.get_clades(), frozenset(other.outliers))

    def to_newick_string(self):
        g, rev = self._graph, self._node_indices_rev

        def rec(i):
            kids = [rec(c) for c in g.successor_indices(i)]
            name = rev[i]
            return "({}){}".format(",".join(kids), name) if kids else "{}".format(name)

        return rec(self._node_indices[self._ROOT_NODE_NAME]) + ";"

    def structure_key(self):
        """Value identity of the frozen tree, i.e. of what `to_dict()` would hold: edges by slot, data lists
        in order, last node added to.  The reference keys its proposal caches on dict equality of frozen
        trees (smc/swarm/particle.py:44-49, tree_holder.py:44-49) -- and because the first-built proposal
        wins and was folded in ITS parent's child order, those cache hits are part of the result once floors
        occur -- so the host needs the same notion of "equal particle".  Kept as an XOR of 64-bit hashes that
        is updated in O(1) by the two SMC edits; anything else recomputes it."""
        if self._skey is None:
            g = self._graph
            k = 0
            s, d = g.esrc, g.edst
            for e in range(len(s)):
                if s[e] >= 0:
                    k ^= _edge_hash(e, s[e], d[e])
            for node, lst in self._data.items():
                for pos, dp in enumerate(lst):
                    k ^= _data_hash(node, pos, dp.idx)
            self._skey = k
        return (self._skey, self._last_node_added_to, self._graph.n_nodes)

    def prior_terms(self):
        """(crp_sum, multiplicity, {graph idx: nodes in its subtree}, outlier_prior, outlier_marginals): the tree
        statistics FSCRPDistribution / TreeJointDistribution need (tree/distributions.py:54-133, 186-219), kept
        up to date in O(1) by the SMC edits instead of being re-derived from the whole tree per candidate."""
        pt = self._pt
        if pt is None:
            g = self._graph
            out = self._OUTLIER_NODE_NAME
            root = self._ROOT_NODE_NAME
            crp = sum(cached_log_factorial(len(v) - 1) for k, v in self._data.items() if k != out and k != root)
            mult = sum(map(cached_log_factorial, map(g.out_degree, g.node_indices())))
            size = {}

            def rec(i):
                n = 1
                for c in g.successor_indices(i):
                    n += rec(c)
                size[i] = n
                return n

            rec(self._node_indices[root])
            oprior = 0
            for node, lst in self._data.items():
                if node == root:
                    continue
                for dp in lst:
                    if dp.outlier_prob != 0:
                        oprior += dp.outlier_prob if node == out else dp.outlier_prob_not
            omarg = 0.0
            for dp in self._data[out]:
                omarg += dp.outlier_marginal_prob
            pt = self._pt = [crp, mult, size, oprior, omarg]
        return pt

    def _pt_attach(self, node, data_points, old_len):
        pt = self._pt
        if pt is None:
            return
        out = self._OUTLIER_NODE_NAME
        if node != out:
            n = old_len
            for _ in data_points:
                # sum of log((len-1)!) over nodes: a node going from n to n+1 points gains log(n) (n >= 1)
                if n >= 1:
                    pt[0] += cached_log_factorial(n) - cached_log_factorial(n - 1)
                n += 1
        for dp in data_points:
            if dp.outlier_prob != 0:
                pt[3] += dp.outlier_prob if node == out else dp.outlier_prob_not
            if node == out:
                pt[4] += dp.outlier_marginal_prob

    def _pt_detach(self, node, dp, old_len):
        """Inverse of `_pt_attach` for one data point (the Gibbs move removes and re-adds a point thousands of times
        per sweep at config #5; re-deriving the statistics is O(T) each time)."""
        pt = self._pt
        if pt is None:
            return
        out = self._OUTLIER_NODE_NAME
        if node != out:
            if old_len >= 2:  # a node going from n to n-1 points loses log(n-1)
                pt[0] -= cached_log_factorial(old_len - 1) - cached_log_factorial(old_len - 2)
            if dp.outlier_prob != 0:
                pt[3] -= dp.outlier_prob_not
        else:
            if dp.outlier_prob != 0:
                pt[3] -= dp.outlier_prob
            pt[4] -= dp.outlier_marginal_prob

    # ------------------------------------------------------------------ queries
    @staticmethod
    def get_single_node_tree(data, store):
        """tree/tree.py:70-88"""
        tree = Tree(data[0].grid_size, store)
        node = tree.create_root_node([])
        tree._add_list_of_data_points_to_node(data, node)
        tree.update()
        return tree

    @property
    def root_node_name(self):
        return self._ROOT_NODE_NAME

    @property
    def outlier_node_name(self):
        return self._OUTLIER_NODE_NAME

    @property
    def data(self):
        return sorted(chain.from_iterable(self._data.values()), key=lambda x: x.idx)

    @property
    def labels(self):
        root = self._ROOT_NODE_NAME
        return {dp.idx: k for k, lst in self._data.items() if k != root for dp in lst}

    @property
    def multiplicity(self):
        g = self._graph
        return sum(map(cached_log_factorial, map(g.out_degree, g.node_indices())))

    @property
    def nodes(self):
        root = self._ROOT_NODE_NAME
        return [n for n in self._graph.nodes() if n != root]

    @property
    def node_last_added_to(self):
        return self._last_node_added_to

    def get_number_of_nodes(self):
        return self._graph.num_nodes() - 1

    @property
    def node_data(self):
        root = self._ROOT_NODE_NAME
        return {k: list(v) for k, v in self._data.items() if k != root}

    @property
    def outliers(self):
        return list(self._data[self._OUTLIER_NODE_NAME])

    @property
    def roots(self):
        return self._graph.successors(self._node_indices[self._ROOT_NODE_NAME])

    def get_children(self, node):
        return self._graph.successors(self._node_indices[node])

    def get_number_of_children(self, node):
        return len(self._graph.out[self._node_indices[node]])

    def get_descendants(self, source=None):
        source = self._ROOT_NODE_NAME if source is None else source
        g = self._graph
        return [g[c] for c in g.descendants(self._node_indices[source])]

    def get_number_of_descendants(self, source=None):
        source = self._ROOT_NODE_NAME if source is None else source
        return len(self._graph.descendants(self._node_indices[source]))

    def get_parent(self, node):
        if node == self._ROOT_NODE_NAME:
            return None
        g = self._graph
        return g[g.predecessor_indices(self._node_indices[node])[0]]

    def get_data(self, node):
        return list(self._data[node])

    def get_data_len(self, node):
        return len(self._data[node])

    def get_subtree_data_len(self, node):
        return self.get_data_len(node) + sum(self.get_data_len(d) for d in self.get_descendants(node))

    # ------------------------------------------------------------- device values
    def root_children_tiles(self):
        """Ordered log_r tile ids of the dummy root's children: the fold that scores the tree."""
        g, lr = self._graph, self._lr
        return tuple(lr[c] for c in g.successor_indices(self._node_indices[self._ROOT_NODE_NAME]))

    def score_handle(self):
        """Lazy (sum_s LSE_g, sum_s last column) of the dummy root's log_r, or None for an empty forest
        (tree/distributions.py:197-200)."""
        self._live()
        kids = self.root_children_tiles()
        return self.store.score(kids) if kids else None

    def node_tiles(self, node):
        """(log_p, log_r) tile ids of a node."""
        i = self._node_indices[node]
        return self._lp[i], self._lr[i]

    def fetch_log_r(self, nodes=None):
        """Dense host copy {node: log_r[S,G]} (debug / tests); the dummy root is folded on request."""
        nodes = self.nodes if nodes is None else nodes
        ids = [self._lr[self._node_indices[n]] for n in nodes]
        arr = self.store.fetch(ids)
        return dict(zip(nodes, arr))

    # ------------------------------------------------------------------ editing
    def _add_node(self, node):
        i = self._graph.add_node(node)
        self._lp[i] = self.store.prior_id
        self._lr[i] = NO_TILE
        self._node_indices[node] = i
        self._node_indices_rev[i] = node
        return i

    def _live(self):
        if self._gen != getattr(self.store, "generation", self._gen):
            raise RuntimeError("this Tree's tiles belong to arena generation %d, the store is at %d: re-read it with "
                               "Tree.from_dict(tree.to_dict(), store) or keep its tiles with store.rebase(tree)"
                               % (self._gen, self.store.generation))

    def _absorb(self, i, data_points):
        """TreeNode.add_data_point(_list) (tree/tree_node.py:34-52): log_p += value; log_r += value."""
        self._live()
        st = self.store
        lp, lr = self._lp[i], self._lr[i]
        for dp in data_points:
            v = st.value_id[dp.idx]
            lp = st.single_id[dp.idx] if lp == st.prior_id else st.add(lp, v)
            if lr != NO_TILE:
                lr = st.add(lr, v)
        self._lp[i], self._lr[i] = lp, lr

    def _add_list_of_data_points_to_node(self, data, node):
        i = self._node_indices[node]
        if len(data) > 0:
            old = self._data[node]
            self._data[node] = old + tuple(data)
            if self._skey is not None:
                for pos, dp in enumerate(data, len(old)):
                    self._skey ^= _data_hash(node, pos, dp.idx)
            self._pt_attach(node, data, len(old))
            self._absorb(i, data)
        return i

    def create_root_node(self, children=None, data=None):
        """tree/tree.py:287-321"""
        data = [] if data is None else data
        children = [] if children is None else children
        g = self._graph
        node = g.num_nodes() - 1
        self._add_node(node)
        r = self._node_indices[self._ROOT_NODE_NAME]
        i = self._add_list_of_data_points_to_node(data, node)
        k = self._skey
        e = g.add_edge(r, i)
        if k is not None:
            k ^= _edge_hash(e, r, i)
        for child in children:
            ci = self._node_indices[child]
            g.remove_edge(r, ci)
            e = g.add_edge(i, ci)
            if k is not None:
                k ^= _edge_hash(e, r, ci) ^ _edge_hash(e, i, ci)  # the freed slot is the one re-used
        self._skey = k
        pt = self._pt
        if pt is not None:
            nk = len(children)
            size = pt[2] = pt[2].copy()
            size[i] = 1 + sum(size[self._node_indices[c]] for c in children)
            size[r] = size.get(r, 1) + 1
            # out-degrees: the dummy root loses nk children and gains one, the new node has nk
            d_root = g.out_degree(r)
            pt[1] += cached_log_factorial(d_root) - cached_log_factorial(d_root + nk - 1) + cached_log_factorial(nk)
        self._last_node_added_to = node
        if self._dirty is not None:
            self._dirty.add(i)
        self._update_path_to_root(node)
        return node

    def add_data_point_to_node(self, data_point, node):
        """tree/tree.py:231-246"""
        old = self._data[node]
        self._data[node] = old + (data_point,)
        if self._skey is not None:
            self._skey ^= _data_hash(node, len(old), data_point.idx)
        self._pt_attach(node, (data_point,), len(old))
        self._last_node_added_to = node
        if node != self._OUTLIER_NODE_NAME:
            i = self._node_indices[node]
            self._absorb(i, (data_point,))
            if self._dirty is not None:  # log_r += value is not what a rebuild computes, here and above
                g, root = self._graph, self._node_indices[self._ROOT_NODE_NAME]
                while i != root:
                    self._dirty.add(i)
                    i = g.predecessor_indices(i)[0]
            self._update_path_to_root(self.get_parent(node))

    def add_data_point_to_outliers(self, data_point):
        self.add_data_point_to_node(data_point, self._OUTLIER_NODE_NAME)

    def _attach_without_refresh(self, data_point, node):
        """tree/tree.py:236-246 with build_add=True: the point joins the node's data and log_p, no path refresh (the
        caller finishes with `update()`).  Used when a tree is assembled from a finished structure (consensus)."""
        self._data[node] = self._data[node] + (data_point,)
        self._last_node_added_to = node
        self._skey = self._pt = self._dirty = None
        if node != self._OUTLIER_NODE_NAME:
            self._absorb(self._node_indices[node], (data_point,))

    def remove_data_point_from_node(self, data_point, node):
        """tree/tree.py:446-454"""
        lst = list(self._data[node])
        lst.remove(data_point)
        self._data[node] = tuple(lst)
        self._skey = None
        self._pt_detach(node, data_point, len(lst) + 1)
        if node != self._OUTLIER_NODE_NAME:
            i = self._node_indices[node]
            self._lp[i] = self.store.sub(self._lp[i], self.store.value_id[data_point.idx])
            self._dirty = None
            self._update_path_to_root(node)

    def remove_data_point_from_outliers(self, data_point):
        lst = list(self._data[self._OUTLIER_NODE_NAME])
        lst.remove(data_point)
        self._data[self._OUTLIER_NODE_NAME] = tuple(lst)
        self._skey = None
        self._pt = None

    def copy(self):
        new = Tree.__new__(Tree)
        new.grid_size = self.grid_size
        new.store = self.store
        new._data = self._data.copy()  # defaultdict.copy() keeps the factory; tuples are shared
        new._graph = self._graph.copy()
        new._node_indices = self._node_indices.copy()
        new._node_indices_rev = self._node_indices_rev.copy()
        new._last_node_added_to = self._last_node_added_to
        new._lp = self._lp.copy()
        new._lr = self._lr.copy()
        new._dirty = None if self._dirty is None else self._dirty.copy()
        new._skey = self._skey
        pt = self._pt
        new._pt = None if pt is None else pt.copy()  # the size dict is replaced, never mutated, on edit
        new._gen = self._gen
        return new

    def thawed_copy(self):
        """Equivalent of `Tree.from_dict(self.to_dict())` (what the reference does every time a particle's
        tree is read, smc/swarm/tree_holder.py:80-82): edges re-inserted in slot order, so every node lists
        its children by descending edge slot, and every node tile refreshed from its data list.  For trees
        grown by SMC extensions only the touched nodes can differ, and those were recorded in `_dirty`."""
        g = self._graph
        if self._dirty is None or g.free_nodes or g.free_edges:
            return Tree.from_dict(self.to_dict(), self.store)
        new = self.copy()
        if not self._dirty:
            return new
        g = new._graph
        depth = {}
        root = new._node_indices[self._ROOT_NODE_NAME]
        for i in self._dirty:
            d, j = 0, i
            while j != root:
                j = g.predecessor_indices(j)[0]
                d += 1
            depth[i] = d
        st = self.store
        for i in sorted(self._dirty, key=lambda x: -depth[x]):
            out = g.out[i]
            if len(out) > 1:
                g.out[i] = tuple(sorted(out, reverse=True))
            # log_p: prior + values in list order (same op chain as the incremental adds: memo hits)
            lp = st.prior_id
            for dp in new._data[new._node_indices_rev[i]]:
                lp = st.single_id[dp.idx] if lp == st.prior_id else st.add(lp, st.value_id[dp.idx])
            new._lp[i] = lp
            new._update_node(i)
        new._dirty = set()
        return new

    # ------------------------------------------------------------ refresh schedule
    def _update_node(self, i):
        """TreeNode.update_node_from_child_r_vals (tree/tree_node.py:61-71), recorded lazily."""
        g, lr = self._graph, self._lr
        kids = tuple(lr[c] for c in g.successor_indices(i))
        if self._node_indices_rev[i] == self._ROOT_NODE_NAME:
            return  # the dummy root's tile is only ever reduced: see score_handle()
        self._live()
        lr[i] = self.store.node(self._lp[i], kids) if kids else self._lp[i]

    def _update_path_to_root(self, source):
        """tree/tree.py:501-518"""
        g = self._graph
        i = self._node_indices[source]
        root = self._node_indices[self._ROOT_NODE_NAME]
        while i != root:
            self._update_node(i)
            i = g.predecessor_indices(i)[0]

    def update(self):
        """tree/tree.py:486-490: post-order refresh of every node."""
        g = self._graph
        order = []

        def rec(i):
            for c in g.successor_indices(i):
                rec(c)
            order.append(i)

        rec(self._node_indices[self._ROOT_NODE_NAME])
        for i in order:
            self._update_node(i)

    # --------------------------------------------------------------- (de)serialise
    def to_dict(self):
        """tree/tree.py:207-217; node_data lists hold the DataPoint objects."""
        return {
            "graph": self._graph.edge_list(),
            "node_idx": self._node_indices.copy(),
            "node_idx_rev": self._node_indices_rev.copy(),
            "node_data": {k: list(v) for k, v in self._data.items()},
            "grid_size": self.grid_size,
            "node_last_added_to": self._last_node_added_to,
            "log_prior": self.store.prior,
        }

    @classmethod
    def from_dict(cls, tree_dict, store):
        """tree/tree.py:163-205"""
        new = cls.__new__(cls)
        new.grid_size = tree_dict["grid_size"]
        new.store = store
        new._data = defaultdict(tuple)
        new._data.update({k: tuple(v) for k, v in tree_dict["node_data"].items()})
        new._node_indices = dict(tree_dict["node_idx"])
        new._node_indices_rev = dict(tree_dict["node_idx_rev"])
        new._last_node_added_to = tree_dict["node_last_added_to"]
        g = new._graph = SlotDiGraph()
        new._lp, new._lr = {}, {}
        new._dirty = set()
        new._skey = None
        new._pt = None
        new._gen = getattr(store, "generation", 0)  # every tile is re-derived below: the tree belongs to the store's now
        r = g.add_node(cls._ROOT_NODE_NAME)
        new._lp[r], new._lr[r] = store.prior_id, NO_TILE
        if len(tree_dict["graph"]) > 0:
            g.extend_from_edge_list(tree_dict["graph"])
            for node, data_list in tree_dict["node_data"].items():
                if node == cls._OUTLIER_NODE_NAME or node == cls._ROOT_NODE_NAME:
                    continue
                i = new._node_indices[node]
                g[i] = node
                new._lp[i], new._lr[i] = store.prior_id, NO_TILE
                new._absorb(i, data_list)
            rev = new._node_indices_rev
            holes = [i for i in g.node_indices() if i not in rev]
            if holes:
                g.remove_nodes_from(holes)
        new.update()
        return new

    # ------------------------------------------------------------ subtree surgery
    def get_subtree(self, subtree_root):
        """tree/tree.py:400-433"""
        if subtree_root == self._ROOT_NODE_NAME:
            return self.copy()
        new = Tree(self.grid_size, self.store)
        g = self._graph
        top = self._node_indices[subtree_root]
        keep = [top] + list(g.descendants(top))
        sub, remap = g.subgraph(keep)
        new_root = new._node_indices[self._ROOT_NODE_NAME]
        mapping = new._graph.compose(sub, {new_root: remap[top]})
        back = {v: k for k, v in remap.items()}
        for si, ni in mapping.items():
            new._lp[ni] = self._lp[back[si]]
            new._lr[ni] = self._lr[back[si]]
        for ni in new._graph.node_indices():
            node = new._graph[ni]
            new._data[node] = tuple(self._data[node])
            new._node_indices[node] = ni
            new._node_indices_rev[ni] = node
        new._dirty = None
        new._skey = None
        new._pt = None
        new.update()
        return new

    def add_subtree(self, subtree, parent=None):
        """tree/tree.py:251-285"""
        subtree = subtree.copy()
        parent = self._ROOT_NODE_NAME if parent is None else parent
        g = self._graph
        pi = self._node_indices[parent]
        sub_root = subtree._node_indices[subtree._ROOT_NODE_NAME]
        mapping = g.compose(subtree._graph, {pi: sub_root})
        g.remove_node_retain_edges(mapping[sub_root])
        first_label = max(self.nodes + subtree.nodes + [-1])
        for old_idx, new_idx in mapping.items():
            if old_idx == sub_root:
                continue
            node_name = old_name = g[new_idx]
            if node_name in self._data:
                first_label += 1
                node_name = first_label
                g[new_idx] = node_name
            self._lp[new_idx] = subtree._lp[old_idx]
            self._lr[new_idx] = subtree._lr[old_idx]
            self._data[node_name] = subtree._data[old_name]
            self._node_indices[node_name] = new_idx
            self._node_indices_rev[new_idx] = node_name
        self._last_node_added_to = subtree._last_node_added_to
        self._dirty = None
        self._skey = None
        self._pt = None
        self._update_path_to_root(parent)

    def remove_subtree(self, subtree):
        """tree/tree.py:459-484"""
        if subtree == self:
            self.__init__(self.grid_size, self.store)
            return
        assert len(subtree.roots) == 1
        sub_root = subtree.roots[0]
        parent = self.get_parent(sub_root)
        g = self._graph
        top = self._node_indices[sub_root]
        for node_id in subtree._graph.nodes():
            if node_id != self._ROOT_NODE_NAME:
                del self._data[node_id]
                i = self._node_indices.pop(node_id)
                del self._node_indices_rev[i]
        doomed = list(g.descendants(top)) + [top]
        g.remove_nodes_from(doomed)
        for i in doomed:
            self._lp.pop(i, None)
            self._lr.pop(i, None)
        self._dirty = None
        self._skey = None
        self._pt = None
        self._update_path_to_root(parent)

    def relabel_nodes(self):
        """tree/tree.py:435-444 + visitors.py:18-45: rename nodes 0.. in pre-order from the dummy root."""
        out = self._OUTLIER_NODE_NAME
        data = defaultdict(tuple)
        data[out] = tuple(self._data[out])
        idx, rev = {}, {}
        g, old_rev, old_data = self._graph, self._node_indices_rev, self._data
        counter = [0]
        root = self._ROOT_NODE_NAME

        def rec(i):
            name = g[i]
            if name != root:
                old, name = name, counter[0]
                counter[0] += 1
                g[i] = name
                data[name] = old_data[old]
            idx[name] = i
            rev[i] = name
            for c in g.successor_indices(i):
                rec(c)

        rec(self._node_indices[root])
        self._data, self._node_indices, self._node_indices_rev = data, idx, rev
        self._skey = None
        self._pt = None
