"""CPU oracle for a whole PhyClone chain  --  TEST INFRASTRUCTURE ONLY.

A self-contained NumPy restatement of the reference sampler around the hot path:
forest bookkeeping, FS-CRP prior, tree scoring, semi-adapted / bootstrap / fully-adapted
proposals, (conditional) SMC with adaptive multinomial resampling, particle Gibbs, the
data-point Gibbs and prune-regraph moves and the concentration update.  It exists so that
the GPU sampler in `phyclone_b200/` can be checked decision-for-decision (same Generator
state => same candidate order, same draws, same topologies) and so that `bench.py` has a
CPU baseline ("port") that runs on the GPU box, where /root/reference does not exist.

Nothing under `phyclone_b200/` imports this module.

Parity status: PINNED against the unmodified reference run in the build container with
the two content-hash LRU caches bypassed: `tests/golden/make_golden_chain.py`
records reference traces, `tests/test_oracle_chain.py` replays them through this module and
requires identical topologies and alpha / log_p_one to 1e-12.  Both sides use the same graph
container (`oracle/refshim/rustworkx`, a stand-in for the rustworkx wheel, which is not
installable here -- ordering parity with the real wheel is unpinned).

The reference's order-INSENSITIVE tile caches are reproduced only by `use_tile_caches=True`
(timing baseline); the default evaluates every fold in the order given (parity baseline).

`file:line` citations are relative to /root/reference/phyclone/.
"""

import math
import os
import sys
from collections import OrderedDict, defaultdict
from functools import lru_cache
from itertools import chain, combinations, repeat

import numpy as np

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)
import rustworkx as rx  # noqa: E402  (the stand-in container, or the real wheel where it exists)

from . import numerics as onum  # noqa: E402

ROOT = "root"  # tree/tree.py:18
OUT = -1  # tree/tree.py:19


# The reference's log-gamma is Numba's math.lgamma, which lowers to the C library's lgamma; CPython's
# own math.lgamma (a Lanczos series) differs from it in the last bit for small integers, and one ulp in
# a proposal log-probability is enough to derail Generator.multinomial (see numerics.lse_1d).
import ctypes as _ct  # noqa: E402
import ctypes.util as _ctu  # noqa: E402

_libm = _ct.CDLL(_ctu.find_library("m") or "libm.so.6")
_libm.lgamma.restype = _ct.c_double
_libm.lgamma.argtypes = [_ct.c_double]


def _lfact(x):  # utils/math.py:126-141 (lgamma(x + 1))
    return _libm.lgamma(float(x) + 1.0)


def _lbinom(n, x):  # utils/math.py:149-151
    return _lfact(n) - _lfact(x) - _lfact(n - x)


# ------------------------------------------------------------------------------------
# data point                                                          data/base.py:5-55
# ------------------------------------------------------------------------------------
class Datum:
    __slots__ = ("idx", "value", "name", "outlier_prob", "outlier_prob_not", "outlier_marginal_prob")

    def __init__(self, idx, value, name=None, outlier_prob=0, outlier_prob_not=1):
        self.idx = idx
        self.value = np.ascontiguousarray(value, dtype=np.float64)
        self.name = idx if name is None else name
        self.outlier_prob = outlier_prob
        self.outlier_prob_not = outlier_prob_not
        self.outlier_marginal_prob = onum.outlier_marginal(self.value)

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return self.name == other.name

    @property
    def grid_size(self):
        return self.value.shape


# ------------------------------------------------------------------------------------
# tile arithmetic with the reference's (optional) caches             tree/utils.py, utils/utils.py
# ------------------------------------------------------------------------------------
class TileMath:
    """`log_s(children)` either in-order (default) or through content-hash LRUs that mimic
    utils/utils.py:70-149 (keys ignore child order; first-seen order wins)."""

    def __init__(self, use_tile_caches=False):
        self.cached = use_tile_caches
        self.n_conv = 0
        if use_tile_caches:
            from xxhash import xxh3_64_hexdigest as digest

            self._digest = digest
            self._pair = OrderedDict()
            self._list = OrderedDict()

    def _conv(self, a, b):
        self.n_conv += 1
        return onum.conv_pair(a, b)

    def _conv_cached(self, a, b):
        key = frozenset([self._digest(a), self._digest(b)])
        hit = self._pair.get(key)
        if hit is not None:
            self._pair.move_to_end(key)
            return hit
        res = self._conv(a, b)
        self._pair[key] = res
        if len(self._pair) > 1024:
            self._pair.popitem(last=False)
        return res

    def log_s(self, children):
        if not self.cached:
            conv = self._conv
        else:
            key = tuple(sorted(self._digest(c) for c in children))
            hit = self._list.get(key)
            if hit is not None:
                self._list.move_to_end(key)
                return hit
            conv = self._conv_cached
        n = len(children)
        if n == 1:
            acc = children[0]
        else:
            acc = conv(children[0], children[1])
            for j in range(2, n):
                acc = conv(children[j], acc)
        res = np.ascontiguousarray(onum.prefix_lse(acc))
        if self.cached:
            self._list[key] = res
            if len(self._list) > 4096:
                self._list.popitem(last=False)
        return res


# ------------------------------------------------------------------------------------
# forest                                                              tree/tree.py:17-523
# ------------------------------------------------------------------------------------
class Forest:
    """Forest with a dummy root.  Graph payloads are the node names; the per-node tiles
    (log_p, log_r) live in dicts keyed by graph index."""

    def __init__(self, grid_size, tm):
        self.grid_size = grid_size
        self.tm = tm
        self.prior = -np.log(grid_size[1])
        self.members = defaultdict(list)  # node name -> [Datum]  (tree.py _data)
        self.g = rx.PyDiGraph()
        self.gi = {}  # name -> graph index
        self.name_of = {}  # graph index -> name
        self.lp = {}
        self.lr = {}
        self.last_added = None
        self._new_slot(ROOT)

    # -- construction helpers
    def _new_slot(self, name):  # tree.py:492-499
        i = self.g.add_node(name)
        self.lp[i] = np.full(self.grid_size, self.prior, order="C")  # tree_node.py:12-13
        self.lr[i] = np.zeros(self.grid_size, order="C")
        self.gi[name] = i
        self.name_of[i] = name
        return i

    def clone(self):  # tree.py:331-355
        new = Forest.__new__(Forest)
        new.grid_size, new.tm, new.prior = self.grid_size, self.tm, self.prior
        new.members = defaultdict(list)
        new.members.update({k: v.copy() for k, v in self.members.items()})
        new.g = self.g.copy()
        new.gi = self.gi.copy()
        new.name_of = self.name_of.copy()
        new.lp = {i: t.copy() for i, t in self.lp.items()}
        new.lr = {i: t.copy() for i, t in self.lr.items()}
        new.last_added = self.last_added
        return new

    # -- queries
    @property
    def root_i(self):
        return self.gi[ROOT]

    @property
    def roots(self):  # tree.py:157-161
        return list(self.g.successors(self.root_i))

    @property
    def nodes(self):  # tree.py:133-136
        return [n for n in self.g.nodes() if n != ROOT]

    def n_nodes(self):
        return self.g.num_nodes() - 1

    def children(self, node):
        return list(self.g.successors(self.gi[node]))

    def n_children(self, node):
        return len(self.g.successors(self.gi[node]))

    def parent(self, node):  # tree.py:380-385
        if node == ROOT:
            return None
        return self.g.predecessors(self.gi[node])[0]

    def n_descendants(self, node):
        return len(rx.descendants(self.g, self.gi[node]))

    @property
    def node_data(self):  # tree.py:146-153
        d = self.members.copy()
        d.pop(ROOT, None)
        return d

    @property
    def outliers(self):
        return list(self.members[OUT])

    @property
    def labels(self):  # tree.py:118-121
        return {dp.idx: k for k, lst in self.node_data.items() for dp in lst}

    @property
    def data(self):  # tree.py:106-109
        return sorted(chain.from_iterable(self.members.values()), key=lambda d: d.idx)

    @property
    def root_tile(self):  # tree.py:112-116 data_log_likelihood
        return self.lr[self.root_i]

    def multiplicity(self):  # tree.py:123-131
        return sum(_lfact(self.g.out_degree(i)) for i in self.g.node_indices())

    def clades(self):  # tree.py:64-69 + visitors.py:48-88
        out = set()

        def rec(i):
            acc = {dp.idx for dp in self.members[self.name_of[i]]}
            for c in self.g.successor_indices(i):
                acc |= rec(c)
            if self.name_of[i] != ROOT:
                out.add(frozenset(acc))
            return acc

        rec(self.root_i)
        return frozenset(out)

    def key(self):
        return (self.clades(), frozenset(self.outliers))

    def __hash__(self):
        return hash(self.key())

    def __eq__(self, other):
        return self.key() == other.key()

    # -- refresh schedule                                             tree.py:486-523
    def _refresh(self, i):  # tree_node.py:61-71
        kids = [self.lr[c] for c in self.g.successor_indices(i)]
        if not kids:
            np.copyto(self.lr[i], self.lp[i])
        else:
            np.add(self.lp[i], self.tm.log_s(kids), out=self.lr[i])

    def refresh_up(self, node):  # tree.py:501-518
        path = [self.gi[node]]
        while self.name_of[path[-1]] != ROOT:
            path.append(self.g.predecessor_indices(path[-1])[0])
        for i in path:
            self._refresh(i)

    def refresh_all(self):  # tree.py:486-490 (post-order DFS from the dummy root)
        order = []

        def rec(i):
            for c in self.g.successor_indices(i):
                rec(c)
            order.append(i)

        rec(self.root_i)
        for i in order:
            self._refresh(i)

    # -- edits
    def _absorb(self, i, dps):  # tree_node.py:34-52
        for dp in dps:
            self.lp[i] += dp.value
            self.lr[i] += dp.value

    def new_root(self, children=None, data=None):  # tree.py:287-321
        data = [] if data is None else data
        children = [] if children is None else children
        node = self.g.num_nodes() - 1
        i = self._new_slot(node)
        if len(data) > 0:
            self.members[node].extend(data)
            self._absorb(i, data)
        r = self.root_i
        self.g.add_edge(r, i, None)
        for c in children:
            ci = self.gi[c]
            self.g.remove_edge(r, ci)
            self.g.add_edge(i, ci, None)
        self.last_added = node
        self.refresh_up(node)
        return node

    def attach(self, dp, node, refresh=True):  # tree.py:231-246
        self.members[node].append(dp)
        self.last_added = node
        if node != OUT:
            self._absorb(self.gi[node], [dp])
            if refresh:
                self.refresh_up(self.parent(node))

    def detach(self, dp, node):  # tree.py:446-457
        self.members[node].remove(dp)
        if node != OUT:
            self.lp[self.gi[node]] -= dp.value  # tree_node.py:54-59
            self.refresh_up(node)

    @staticmethod
    def single_node(data, tm):  # tree.py:70-88
        t = Forest(data[0].grid_size, tm)
        node = t.new_root([])
        t.members[node].extend(data)
        t._absorb(t.gi[node], data)
        t.refresh_all()
        return t

    # -- (de)serialisation                                            tree.py:163-217
    def freeze(self):
        return {
            "graph": self.g.edge_list(),
            "node_idx": self.gi.copy(),
            "node_idx_rev": self.name_of.copy(),
            "node_data": {k: v.copy() for k, v in self.members.items()},
            "grid_size": self.grid_size,
            "node_last_added_to": self.last_added,
            "log_prior": self.prior,
        }

    @staticmethod
    def thaw(d, tm):
        t = Forest.__new__(Forest)
        t.grid_size, t.tm, t.prior = d["grid_size"], tm, d["log_prior"]
        t.members = defaultdict(list)
        t.members.update({k: v.copy() for k, v in d["node_data"].items()})
        t.gi = d["node_idx"].copy()
        t.name_of = d["node_idx_rev"].copy()
        t.last_added = d["node_last_added_to"]
        t.g = rx.PyDiGraph()
        t.lp, t.lr = {}, {}
        r = t.g.add_node(ROOT)
        t.lp[r] = np.full(t.grid_size, t.prior, order="C")
        t.lr[r] = np.zeros(t.grid_size, order="C")
        if len(d["graph"]) > 0:
            t.g.extend_from_edge_list(d["graph"])
            for node, dps in d["node_data"].items():
                if node == OUT or node == ROOT:
                    continue
                i = t.gi[node]
                t.g[i] = node
                t.lp[i] = np.full(t.grid_size, t.prior, order="C")
                t.lr[i] = np.zeros(t.grid_size, order="C")
                t._absorb(i, dps)
            holes = [i for i in t.g.node_indices() if i not in d["node_idx_rev"]]
            if holes:
                t.g.remove_nodes_from(holes)
        t.refresh_all()
        return t

    # -- subtree surgery                                              tree.py:251-285, 400-484
    def subtree(self, top):
        if top == ROOT:
            return self.clone()
        new = Forest(self.grid_size, self.tm)
        ti = self.gi[top]
        keep = [ti] + list(rx.descendants(self.g, ti))
        sub = self.g.subgraph(keep, preserve_attrs=True)
        sub_lp, sub_lr = {}, {}
        for si, old in zip(sub.node_indices(), sorted(set(keep))):
            sub_lp[si], sub_lr[si] = self.lp[old], self.lr[old]
        sub_top = next(si for si in sub.node_indices() if sub[si] == top)
        mapping = new.g.compose(sub, {new.root_i: (sub_top, None)})
        for si, ni in mapping.items():
            new.lp[ni] = sub_lp[si].copy()
            new.lr[ni] = sub_lr[si].copy()
        for ni in new.g.node_indices():
            node = new.g[ni]
            new.members[node] = list(self.members[node])
            new.gi[node] = ni
            new.name_of[ni] = node
        new.refresh_all()
        return new

    def graft(self, sub, parent=None):
        sub = sub.clone()
        parent = ROOT if parent is None else parent
        pi = self.gi[parent]
        sub_root = sub.root_i
        mapping = self.g.compose(sub.g, {pi: (sub_root, None)})
        self.g.remove_node_retain_edges(mapping[sub_root])
        first = max(self.nodes + sub.nodes + [-1])
        for old, ni in mapping.items():
            if old == sub_root:
                continue
            name = old_name = self.g[ni]
            if name in self.members:
                first += 1
                name = first
                self.g[ni] = name
            self.lp[ni], self.lr[ni] = sub.lp[old], sub.lr[old]
            self.members[name] = sub.members[old_name]
            self.gi[name] = ni
            self.name_of[ni] = name
        self.last_added = sub.last_added
        self.refresh_up(parent)

    def prune(self, sub):
        if sub == self:
            self.__init__(self.grid_size, self.tm)
            return
        top = sub.roots[0]
        parent = self.parent(top)
        ti = self.gi[top]
        for name in sub.g.nodes():
            if name != ROOT:
                del self.members[name]
                i = self.gi.pop(name)
                del self.name_of[i]
        doomed = list(rx.descendants(self.g, ti)) + [ti]
        self.g.remove_nodes_from(doomed)
        for i in doomed:
            self.lp.pop(i, None)
            self.lr.pop(i, None)
        self.refresh_up(parent)

    def renumber(self):  # tree.py:435-444 + visitors.py:18-45 (pre-order from the dummy root)
        members = defaultdict(list)
        members[OUT] = list(self.members[OUT])
        gi, name_of = {}, {}
        counter = [0]

        def rec(i):
            name = self.g[i]
            if name != ROOT:
                old, name = name, counter[0]
                counter[0] += 1
                self.g[i] = name
                members[name] = self.members[old]
            gi[name] = i
            name_of[i] = name
            for c in self.g.successor_indices(i):
                rec(c)

        rec(self.root_i)
        self.members, self.gi, self.name_of = members, gi, name_of


# ------------------------------------------------------------------------------------
# FS-CRP prior + joint                                   tree/distributions.py:6-219
# ------------------------------------------------------------------------------------
class Joint:
    def __init__(self, alpha, c_const=1000, perm_dist=False):
        self.alpha = alpha
        self.log_c = np.log(c_const)
        # True: particle weights carry the root-permutation density (smc/kernels/base.py:49-55), as in the reference's
        # posterior tests; production runs without it (run.py:403-417 builds the kernel with perm_dist=None)
        self.perm_dist = perm_dist

    @property
    def log_alpha(self):
        return np.log(self.alpha)

    def _r_term(self, n_roots):  # distributions.py:103-133
        if n_roots == 0:
            z, n_roots = 0.0, 1
        else:
            num = np.log1p(-np.exp(-(self.log_c * n_roots)))
            den = np.log1p(-np.exp(-(self.log_c * 1)))
            z = num - den
        return -(z + self.log_c * (n_roots - 1))

    def priors(self, t):  # distributions.py:39-101
        nd = t.node_data
        n = t.n_nodes()
        crp = 0
        crp += n * self.log_alpha
        crp += sum(_lfact(len(v) - 1) for k, v in nd.items() if k != OUT)
        mult = t.multiplicity()
        lp = crp - (n - 1) * np.log(n + 1)
        lp -= mult
        roots = t.roots
        ways = 0
        for r in roots:
            sz = t.n_descendants(r) + 1
            ways += (sz - 1) * np.log(sz)
        lp1 = crp + (-ways + self._r_term(len(roots)))
        lp1 -= mult
        return lp, lp1, nd

    @staticmethod
    def outlier_prior(nd):  # distributions.py:208-219
        lp = 0
        for node, dps in nd.items():
            for dp in dps:
                if dp.outlier_prob != 0:
                    lp += dp.outlier_prob if node == OUT else dp.outlier_prob_not
        return lp

    def both(self, t):  # distributions.py:186-206
        lp, lp1, nd = self.priors(t)
        op = self.outlier_prior(nd)
        lp += op
        lp1 += op
        if t.n_children(ROOT) > 0:
            tile = t.root_tile
            for s in range(t.grid_size[0]):
                lp += onum.lse_1d(tile[s, :])
                lp1 += tile[s, -1]
        for dp in t.outliers:
            lp += dp.outlier_marginal_prob
            lp1 += dp.outlier_marginal_prob
        return lp, lp1

    def log_p_one(self, t):
        return self.both(t)[1]


# ------------------------------------------------------------------------------------
# tree holder / particle                      smc/swarm/tree_holder.py, particle.py
# ------------------------------------------------------------------------------------
def _freeze_key(d):
    """Hashable stand-in for `dict == dict` on frozen trees (dict equality ignores key order)."""
    return (
        tuple(d["graph"]),
        frozenset(d["node_idx"].items()),
        frozenset((k, tuple(dp.name for dp in v)) for k, v in d["node_data"].items()),
        tuple(d["grid_size"]),
        d["node_last_added_to"],
    )


def perm_log_pdf(t):
    """RootPermutationDistribution.log_pdf (smc/utils.py:18-66): minus the log number of root permutations."""

    def data_len(i):
        return len(t.members[t.name_of[i]])

    def sub_len(i):
        return data_len(i) + sum(data_len(c) for c in rx.descendants(t.g, i))

    def multinomial(sizes):  # utils/math.py:159-177
        if len(sizes) == 0:
            return 0
        res = _lfact(sum(sizes))
        for x in sizes:
            res -= _lfact(x)
        return res

    def count(i):
        total, sizes = 0, []
        for c in t.g.successor_indices(i):
            total += count(c)
            sizes.append(sub_len(c))
        total += multinomial(sizes)
        if t.name_of[i] != ROOT:
            total += _lfact(data_len(i))
        return total

    n_out = len(t.members[OUT])
    n_all = len(t.data)
    return -(count(t.root_i) + _lbinom(n_all, n_out))


class Holder:
    """Scored, frozen tree (TreeHolder): hash by clades/outliers, equality by frozen dict."""

    __slots__ = ("log_p", "log_p_one", "h", "frozen", "fkey", "tree_roots", "tree_nodes", "labels", "last_added",
                 "n_children_last", "tm", "log_pdf")

    def __init__(self, t, joint):
        self.log_p, self.log_p_one = joint.both(t)
        self.log_pdf = perm_log_pdf(t) if getattr(joint, "perm_dist", False) else 0.0  # tree_holder.py:62-65
        self.tree_roots = np.asarray(t.roots)
        self.tree_nodes = t.nodes
        self.h = hash(t)
        self.frozen = t.freeze()
        self.fkey = _freeze_key(self.frozen)
        self.labels = t.labels
        self.last_added = t.last_added
        self.n_children_last = t.n_children(self.last_added) if self.last_added != OUT else 0
        self.tm = t.tm

    def __hash__(self):
        return self.h

    def __eq__(self, other):
        return self.fkey == other.fkey

    def thaw(self):
        return Forest.thaw(self.frozen, self.tm)


class Particle:
    __slots__ = ("log_w", "parent", "holder", "built")

    def __init__(self, log_w, parent, holder):
        self.log_w, self.parent, self.holder = log_w, parent, holder
        self.built = None

    log_p = property(lambda s: s.holder.log_p)
    log_p_one = property(lambda s: s.holder.log_p_one)
    tree_roots = property(lambda s: s.holder.tree_roots)
    tree_nodes = property(lambda s: s.holder.tree_nodes)

    def __hash__(self):
        return self.holder.h

    def __eq__(self, other):
        return isinstance(other, Particle) and self.holder == other.holder

    def thaw(self):
        return self.holder.thaw()


# ------------------------------------------------------------------------------------
# proposals                                                        smc/kernels/*.py
# ------------------------------------------------------------------------------------
class Kernel:
    """Shared by the three proposal families (smc/kernels/base.py:4-67)."""

    name = "semi-adapted"

    def __init__(self, joint, rng, tm, outlier_proposal_prob=0.0):
        self.joint, self.rng, self.tm = joint, rng, tm
        self.rho = outlier_proposal_prob
        self.n_ext = 0
        self._dist_cache = OrderedDict()
        self._new_cache = OrderedDict()

    def clear(self):  # utils/dev.py:12-17
        self._dist_cache.clear()
        self._new_cache.clear()

    def _lru(self, cache, key, make, size=1024):
        if getattr(self.rng, "unbounded_caches", False):  # uniform-driven mode (oracle/uniform.py): never evict
            size = 1 << 62
        hit = cache.get(key)
        if hit is not None:
            cache.move_to_end(key)
            return hit
        val = make()
        cache[key] = val
        if len(cache) > size:
            cache.popitem(last=False)
        return val

    def proposal(self, dp, parent, parent_tree=None):
        if parent is not None:
            parent.built = parent_tree
        key = (dp, parent, self.rho, self.joint.alpha)
        return self._lru(self._dist_cache, key, lambda: self._make(dp, parent))

    def _make(self, dp, parent):
        pt = None
        if parent is not None:
            pt, parent.built = parent.built, None
            if pt is None:
                pt = parent.thaw()
        return SemiAdapted(self, dp, parent, pt)

    def particle(self, log_q, parent, holder):  # base.py:38-57 (perm_dist is None in production)
        if not getattr(self.joint, "perm_dist", False):
            lw = holder.log_p - log_q if parent is None else holder.log_p - parent.log_p - log_q
        elif parent is None:
            lw = holder.log_p + holder.log_pdf - log_q
        else:
            lw = holder.log_p - parent.log_p + holder.log_pdf - parent.holder.log_pdf - log_q
        return Particle(lw, parent, holder)

    def propose(self, dp, parent):  # base.py:59-67
        self.n_ext += 1
        dist = self.proposal(dp, parent)
        holder = dist.sample()
        return self.particle(dist.log_p(holder), parent, holder)


class _Dist:
    def __init__(self, k, dp, parent, parent_tree):
        self.k, self.dp, self.parent, self.ptree = k, dp, parent, parent_tree

    def _empty(self):
        return self.parent is None or len(self.parent.tree_roots) == 0

    def _existing(self):  # semi_adapted.py:111-126
        out = []
        if self.parent is None:
            return out
        for node in self.parent.tree_roots:
            t = self.ptree.clone()
            t.attach(self.dp, node)
            out.append(Holder(t, self.k.joint))
        return out

    def _outlier(self):  # semi_adapted.py:128-141
        t = Forest(self.dp.grid_size, self.k.tm) if self.parent is None else self.ptree.clone()
        t.attach(self.dp, OUT)
        return Holder(t, self.k.joint)


class SemiAdapted(_Dist):  # smc/kernels/semi_adapted.py:11-183
    def __init__(self, k, dp, parent, parent_tree):
        super().__init__(k, dp, parent, parent_tree)
        self.empty_parent = False
        trees = self._existing()
        if k.rho > 0:
            trees.append(self._outlier())
        if self._empty():
            self.empty_parent = True
            t = Forest(dp.grid_size, k.tm) if parent is None else self.ptree.clone()
            t.new_root(children=[], data=[dp])
            trees.append(Holder(t, k.joint))
        else:
            self.log_r1 = np.log(len(parent.tree_roots) + 1)
        self.log_q, self.q = onum.candidate_q(np.array([h.log_p for h in trees]))
        self.trees = trees
        self.table = dict(zip(trees, self.log_q))
        self.ptree = None

    def log_p(self, holder):  # :36-61
        if self.empty_parent:
            return self.table[holder]
        node = holder.labels[self.dp.idx]
        assert node == holder.last_added
        if node in self.parent.tree_nodes or node == OUT:
            return self.k.log_half + self.table[holder]
        r = len(self.parent.tree_roots)
        return self.k.log_half - (self.log_r1 + _lbinom(r, holder.n_children_last))

    def sample(self):  # :71-83, 143-183
        rng = self.k.rng
        if self.empty_parent or rng.random() < 0.5:
            return self.trees[rng.multinomial(1, self.q).argmax()]
        r = len(self.parent.tree_roots)
        n = rng.integers(0, r + 1)
        kids = [] if n == 0 else rng.choice(self.parent.tree_roots, n, replace=False)
        kids = frozenset(kids)
        key = (self.parent, self.dp, kids)

        def make():
            t = self.parent.thaw()
            adopt = kids
            if getattr(rng, "canonical_adoption", False):  # oracle/uniform.py: descending root position
                adopt = [c for c in reversed(t.roots) if c in kids]
            t.new_root(children=adopt, data=[self.dp])
            return Holder(t, self.k.joint)

        return self.k._lru(self.k._new_cache, key, make)


class SemiAdaptedKernel(Kernel):
    name = "semi-adapted"
    log_half = np.log(0.5)


class FullyAdapted(_Dist):  # smc/kernels/fully_adapted.py:11-105
    def __init__(self, k, dp, parent, parent_tree):
        super().__init__(k, dp, parent, parent_tree)
        trees = self._existing()
        if parent is None:
            t = Forest(dp.grid_size, k.tm)
            t.new_root(children=[], data=[dp])
            trees.append(Holder(t, k.joint))
        else:
            roots = parent.tree_roots
            for r in range(0, len(roots) + 1):
                for kids in combinations(roots, r):
                    t = self.ptree.clone()
                    t.new_root(children=kids, data=[dp])
                    trees.append(Holder(t, k.joint))
        if k.rho > 0:
            trees.append(self._outlier())
        log_q = onum.log_normalize(np.array([h.log_p for h in trees]))
        self.table = dict(zip(trees, log_q))
        self.ptree = None

    def log_p(self, holder):
        return self.table[holder]

    def sample(self):
        p = np.exp(np.array(list(self.table.values())))
        p = p / np.sum(p)
        idx = self.k.rng.multinomial(1, p).argmax()
        return list(self.table.keys())[idx]


class FullyAdaptedKernel(Kernel):
    name = "fully-adapted"

    def _make(self, dp, parent):
        pt = None
        if parent is not None:
            pt, parent.built = parent.built, None
            if pt is None:
                pt = parent.thaw()
        return FullyAdapted(self, dp, parent, pt)


class Bootstrap(_Dist):  # smc/kernels/bootstrap.py:8-130
    def log_p(self, holder):
        rho, dp = self.k.rho, self.dp
        node = holder.labels[dp.idx]
        if self.parent is None:
            return np.log(rho) if node == OUT else np.log(1 - rho)
        if node == OUT:
            return np.log(rho)
        if node in self.ptree.nodes:
            return np.log((1 - rho) / 2) - np.log(len(self.parent.tree_roots))
        r = len(self.parent.tree_roots)
        lp = np.log((1 - rho) / 2)
        if r > 0:
            lp -= np.log(r + 1) + _lbinom(r, holder.n_children_last)
        return lp

    def sample(self):
        rng, rho, dp = self.k.rng, self.k.rho, self.dp
        u = rng.random()
        if self.parent is None:
            t = Forest(dp.grid_size, self.k.tm)
            if u < (1 - rho):
                node = t.new_root([])
                t.attach(dp, node)
            else:
                t.attach(dp, OUT)
        elif len(self.ptree.nodes) == 0:
            t = self._new() if u < (1 - rho) else self._out()
        elif u < (1 - rho) / 2:
            node = rng.choice(list(self.parent.tree_roots))
            t = self.ptree.clone()
            t.attach(dp, node)
        elif u < (1 - rho):
            t = self._new()
        else:
            t = self._out()
        return Holder(t, self.k.joint)

    def _new(self):
        rng = self.k.rng
        r = len(self.parent.tree_roots)
        n = rng.integers(0, r + 1)
        kids = rng.choice(self.parent.tree_roots, n, replace=False)
        t = self.ptree.clone()
        t.new_root(children=kids, data=[self.dp])
        return t

    def _out(self):
        t = self.ptree.clone()
        t.attach(self.dp, OUT)
        return t


class BootstrapKernel(Kernel):
    name = "bootstrap"

    def proposal(self, dp, parent, parent_tree=None):  # not cached (bootstrap.py:118-130)
        pt = parent_tree
        if parent is not None and pt is None:
            pt = parent.thaw()
        return Bootstrap(self, dp, parent, pt)


KERNELS = {"semi-adapted": SemiAdaptedKernel, "fully-adapted": FullyAdaptedKernel, "bootstrap": BootstrapKernel}


# ------------------------------------------------------------------------------------
# swarm + SMC                                    smc/swarm/swarm.py, smc/samplers/*.py
# ------------------------------------------------------------------------------------
class Swarm:
    def __init__(self):
        self.particles, self.raw = [], []

    def add(self, lw, particle):
        self.particles.append(particle)
        self.raw.append(lw)

    def stats(self):
        return onum.swarm_stats(np.fromiter(self.raw, dtype=np.float64, count=len(self.raw)))

    @property
    def log_weights(self):
        return self.stats()[1]

    @property
    def weights(self):
        return self.stats()[2]

    @property
    def relative_ess(self):
        return self.stats()[3] / len(self.particles)


def _last_term(p, last):  # smc/samplers/base.py:52-58
    return p.log_w - p.log_p + p.log_p_one if last else p.log_w


def _at(rng, t, i):
    """Uniform-driven mode (oracle/uniform.py): announce which (SMC step, particle slot) draws next."""
    at = getattr(rng, "at", None)
    if at is not None:
        at(t, i)


def _note(kernel, rec):
    """Decision log for step-level comparisons with the device sweep (tests)."""
    log = getattr(kernel, "sweep_log", None)
    if log is not None:
        log.append(rec)


def smc_unconditional(sigma, kernel, n_particles, threshold):  # standard.py + base.py:32-45
    rng = kernel.rng
    T = len(sigma)
    swarm = Swarm()
    for _ in range(n_particles):
        swarm.add(-np.log(n_particles), None)
    swarm = _maybe_resample(swarm, rng, n_particles, threshold, None, -1, kernel)
    for it in range(T):
        new = Swarm()
        for i, (lw, parent) in enumerate(zip(swarm.log_weights, swarm.particles)):
            _at(rng, it, i)
            p = kernel.propose(sigma[it], parent)
            new.add(lw + _last_term(p, it >= T - 1), p)
        swarm = new
        _note(kernel, ("step", it, list(swarm.raw), [p.log_p for p in swarm.particles],
                       [p.holder.last_added for p in swarm.particles]))
        if it < T - 1:
            swarm = _maybe_resample(swarm, rng, n_particles, threshold, None, it, kernel)
    return swarm


def _maybe_resample(swarm, rng, n_particles, threshold, retained, it=0, kernel=None):
    """standard.py:20-32 / conditional.py:91-109 (`retained` = callable giving the retained particle)."""
    if swarm.relative_ess > threshold:
        return swarm
    new = Swarm()
    lu = -np.log(n_particles)
    from .uniform import RESAMPLE_SLOT

    _at(rng, it, RESAMPLE_SLOT)
    if retained is None:
        counts = rng.multinomial(n_particles, swarm.weights)
    else:
        counts = rng.multinomial(n_particles - 1, swarm.weights)
        new.add(lu, retained())
    _note(kernel, ("resample", it, [int(c) for c in counts]))
    for p, m in zip(swarm.particles, counts):
        for _ in range(m):
            new.add(lu, p)
    return new


def constrained_path(tree, sigma, kernel):  # conditional.py:19-74
    path = [None]
    label = tree.labels
    node_map = {}
    new = Forest(tree.grid_size, kernel.tm)
    prev = None
    for dp in sigma:
        new = new.clone()
        old = label[dp.idx]
        if old == OUT:
            new.attach(dp, OUT)
        elif old in node_map:
            new.attach(dp, node_map[old])
        else:
            kids = [node_map[c] for c in tree.children(old)]
            node = new.new_root(kids)
            node_map[old] = node
            new.attach(dp, node)
        parent = path[-1]
        dist = kernel.proposal(dp, parent, prev)
        holder = Holder(new, kernel.joint)
        path.append(kernel.particle(dist.log_p(holder), parent, holder))
        prev = new
    return path


def smc_conditional(tree, sigma, kernel, n_particles, threshold):  # conditional.py:76-125 + base.py:32-45
    rng = kernel.rng
    T = len(sigma)
    path = constrained_path(tree, sigma, kernel)
    swarm = Swarm()
    lu = -np.log(n_particles)
    swarm.add(lu, path[1])
    for i in range(1, n_particles):
        _at(rng, 0, i)
        swarm.add(lu, kernel.propose(sigma[0], None))
    _note(kernel, ("step", 0, list(swarm.raw), [p.log_p for p in swarm.particles],
                   [p.holder.last_added for p in swarm.particles]))
    it = 1
    # base.py:35 resamples right after the initial swarm with the *next* retained particle
    # (conditional.py:101 reads constrained_path[iteration + 1]); uniform weights never trigger it.
    swarm = _maybe_resample(swarm, rng, n_particles, threshold, lambda: path[it + 1], 0, kernel)
    while it < T:
        new = Swarm()
        last = it >= T - 1
        new.add(swarm.log_weights[0] + _last_term(path[it + 1], last), path[it + 1])
        for i, (lw, parent) in enumerate(zip(swarm.log_weights[1:], swarm.particles[1:]), 1):
            _at(rng, it, i)
            p = kernel.propose(sigma[it], parent)
            new.add(lw + _last_term(p, last), p)
        swarm = new
        _note(kernel, ("step", it, list(swarm.raw), [p.log_p for p in swarm.particles],
                       [p.holder.last_added for p in swarm.particles]))
        if it < T - 1:
            swarm = _maybe_resample(swarm, rng, n_particles, threshold, lambda: path[it + 1], it, kernel)
        it += 1
    return swarm


# ------------------------------------------------------------------------------------
# sigma                                                          smc/utils.py:68-119
# ------------------------------------------------------------------------------------
def _interleave(lists, rng):
    marks = []
    for i, lst in enumerate(lists):
        marks.extend(repeat(i, len(lst)))
    rng.shuffle(marks)
    return [lists[i].pop(0) for i in marks]


def sample_sigma(tree, rng, source=None):
    if source is None:
        sig = [sample_sigma(tree, rng, r) for r in tree.roots]
        sig = _interleave(sig, rng)
        outs = list(tree.outliers)
        rng.shuffle(outs)
        return _interleave([sig, outs], rng)
    sub = [sample_sigma(tree, rng, c) for c in tree.children(source)]
    sig = _interleave(sub, rng)
    own = list(tree.members[source])
    rng.shuffle(own)
    sig.extend(own)
    return sig


# ------------------------------------------------------------------------------------
# MCMC moves                                   mcmc/particle_gibbs.py, gibbs_mh.py, concentration.py
# ------------------------------------------------------------------------------------
def _pick(swarm, rng, kernel=None):  # utils/math.py:19-21 via particle_gibbs.py:42-48
    p = swarm.weights
    p = p / np.sum(p)
    from .uniform import FINAL_SLOT

    _at(rng, 0, FINAL_SLOT)
    idx = rng.multinomial(1, p).argmax()
    _note(kernel, ("pick", int(idx)))
    return swarm.particles[idx].thaw()


def _sweep(rng, run):
    """Uniform-driven mode: one seed from the chain's Generator per sweep (after sigma), keyed uniforms inside."""
    begin = getattr(rng, "begin", None)
    if begin is None:
        return run()
    begin(int(rng.generator.integers(0, 1 << 62)))
    try:
        return run()
    finally:
        rng.end()


def pg_move(tree, kernel, n_particles, threshold):
    sigma = sample_sigma(tree, kernel.rng)
    return _sweep(kernel.rng,
                  lambda: _pick(smc_conditional(tree, sigma, kernel, n_particles, threshold), kernel.rng, kernel))


def subtree_pg_move(tree, kernel, n_particles, threshold):  # mcmc/particle_gibbs.py:51-109
    """ParticleGibbsSubtreeSampler: re-sample the subtree under the parent of a node picked through a random data
    point (outliers travel with the subtree), then re-weight every particle by the full tree it completes."""
    rng, joint = kernel.rng, kernel.joint
    nodes = [label for label in tree.labels.values() if label != OUT]
    child = rng.choice(nodes)
    top = tree.parent(child)
    parent = tree.parent(top)
    sub = tree.subtree(top)
    tree.prune(sub)
    for dp in tree.outliers:
        tree.members[OUT].remove(dp)
        sub.attach(dp, OUT)
    sigma = sample_sigma(sub, rng)

    def run():
        swarm = smc_conditional(sub, sigma, kernel, n_particles, threshold)
        fixed = Swarm()
        for p, w in zip(swarm.particles, swarm.raw):
            part = p.thaw()
            w = w - p.log_p_one
            full = tree.clone()
            full.graft(part, parent=parent)
            for dp in part.outliers:
                full.attach(dp, OUT)
            full.refresh_all()
            q = Particle(p.log_w, p.parent, Holder(full, joint))
            fixed.add(w + q.log_p_one, q)
        return _pick(fixed, rng, kernel)

    return _sweep(rng, run)


def burnin_move(tree, kernel, n_particles, threshold):  # smc/samplers/unconditional.py
    sigma = sample_sigma(tree, kernel.rng)
    return _sweep(kernel.rng,
                  lambda: _pick(smc_unconditional(sigma, kernel, n_particles, threshold), kernel.rng, kernel))


def datapoint_move(tree, joint, rng, outliers):  # gibbs_mh.py:20-70
    label = tree.labels
    idxs = list(label.keys())
    rng.shuffle(idxs)
    for idx in idxs:
        old = label[idx]
        if len(tree.members[old]) > 1:
            dp = tree.data[idx]
            assert dp.idx == idx
            cands = []
            for node in tree.nodes:
                t = tree.clone()
                t.detach(dp, old)
                t.attach(dp, node)
                cands.append(t)
            if outliers:
                t = tree.clone()
                t.detach(dp, old)
                t.attach(dp, OUT)
                cands.append(t)
            log_q = onum.log_normalize(np.array([joint.log_p_one(t) for t in cands]))
            q = np.exp(log_q)
            q = q / sum(q)
            tree = cands[rng.multinomial(1, q).argmax()]
            label = tree.labels
    return tree


def prune_regraft_move(tree, joint, rng):  # gibbs_mh.py:83-129
    if tree.n_nodes() <= 1:
        return tree
    pruned = tree.clone()
    top = rng.choice(pruned.nodes)
    sub = pruned.subtree(top)
    pruned.prune(sub)
    remaining = pruned.nodes
    if len(remaining) == 0:
        return tree
    remaining.append(None)
    cands = []
    for parent in remaining:
        t = pruned.clone()
        nc = t.n_children(ROOT if parent is None else parent)
        t.graft(sub, parent=parent)
        t.refresh_all()
        cands.append((nc, t))
    log_p = np.array([np.log(n + 1) + joint.log_p_one(t) for n, t in cands])
    p, _ = onum.exp_normalize(log_p)
    return cands[rng.multinomial(1, p).argmax()][1]


def alpha_move(joint, tree, rng, a=0.01, b=0.01):  # concentration.py:29-60, run.py:305-315
    from scipy.stats import bernoulli, beta, gamma

    sizes = [len(v) for k, v in tree.node_data.items() if k != OUT]
    k, n = len(sizes), sum(sizes)
    rs = getattr(rng, "generator", rng)  # scipy wants the real Generator (tests may pass a recording proxy)
    if k == 0:
        joint.alpha = gamma.rvs(a, scale=(1 / b), random_state=rs)
        return
    eta = beta.rvs(a=joint.alpha + 1, b=n, random_state=rs)
    shape = a + k - 1
    rate = b - np.log(eta)
    x = shape / (n * rate)
    shape += bernoulli.rvs(x / (1 + x), random_state=rs)
    joint.alpha = max(gamma.rvs(shape, scale=(1 / rate), random_state=rs), 1e-10)


# ------------------------------------------------------------------------------------
# chain driver                                                         run.py:177-380
# ------------------------------------------------------------------------------------
def run_chain(data, rng, num_iters, burnin=1, num_particles=100, proposal="semi-adapted", resample_threshold=0.5,
              outlier_prob=0.0, concentration_value=1.0, concentration_update=True, num_samples_data_point=1,
              num_samples_prune_regraph=1, use_tile_caches=False, on_iteration=None, sweep_log=None,
              subtree_update_prob=0.0, perm_dist=False):
    """Returns {"trace": [...], "extensions": int, "convs": int}; trace entry =
    {"iter", "alpha", "log_p_one", "tree": frozen dict, "clades", "outliers"} (run.py:293-302).
    `sweep_log`: a list that receives the step / resample / pick records of every SMC sweep (see `_note`)."""
    tm = TileMath(use_tile_caches)
    joint = Joint(concentration_value, perm_dist=perm_dist)
    kernel = KERNELS[proposal](joint, rng, tm, outlier_proposal_prob=0.1 if outlier_prob > 0 else 0.0)
    if sweep_log is not None:
        kernel.sweep_log = sweep_log
    tree = Forest.single_node(data, tm)

    def moves(t):
        for _ in range(num_samples_data_point):
            t = datapoint_move(t, joint, rng, outlier_prob > 0)
        for _ in range(num_samples_prune_regraph):
            t = prune_regraft_move(t, joint, rng)
        t.renumber()
        return t

    for _ in range(burnin):
        kernel.clear()
        tree = moves(burnin_move(tree, kernel, num_particles, resample_threshold))

    def entry(i, t):
        return {
            "iter": i,
            "alpha": joint.alpha,
            "log_p_one": joint.log_p_one(t),
            "tree": t.freeze(),
            "clades": t.clades(),
            "outliers": frozenset(d.idx for d in t.outliers),
        }

    trace = [entry(0, tree)]
    for i in range(num_iters):
        kernel.clear()
        if rng.random() < subtree_update_prob:  # run.py:268 (the draw is consumed even when the probability is 0)
            tree = moves(subtree_pg_move(tree, kernel, num_particles, resample_threshold))
        else:
            tree = moves(pg_move(tree, kernel, num_particles, resample_threshold))
        if concentration_update:
            alpha_move(joint, tree, rng)
        trace.append(entry(i, tree))
        if on_iteration is not None:
            on_iteration(i, trace[-1])
    return {"trace": trace, "extensions": kernel.n_ext, "convs": tm.n_conv}
