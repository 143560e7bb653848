"""Drop-in for the reference's `TreeNode` (tree/tree_node.py:8-80) on host buffers.

The sampler of this package never builds these objects (its trees hold tile ids into the device arena, see
tree/tree.py); the class exists for code written against the reference seam.  A node owns two NumPy tiles,
`log_p = log_prior + sum of its data points' grids` and `log_r = log_p + log_S(children)`, plus the set of data-point
indices assigned to it.  Bookkeeping is host-side array arithmetic as in the reference; the one expensive method --
`update_node_from_child_r_vals`, the fold + scan over the children -- is a single fused launch through the C ABI
(`pcb_node_update_host`).  Children are folded in the order given (no order-insensitive LRU).
"""

import numpy as np

from .. import ops


class TreeNode(object):
    __slots__ = ("log_p", "log_r", "node_id", "data_points")

    def __init__(self, grid_size, log_prior, node_id):
        self.node_id = node_id
        self.data_points = set()
        self.log_p = np.full(grid_size, log_prior, order="C")
        self.log_r = np.zeros(grid_size, order="C")

    # ------------------------------------------------------------------ data points
    def _claim(self, indices):
        """Register data-point indices with this node; a point may sit in a node only once (tree_node.py:35-36, 48-49)."""
        fresh = set(indices)
        assert not (fresh & self.data_points)
        self.data_points |= fresh

    def _shift(self, grid, sign):
        np.add(self.log_p, grid, out=self.log_p) if sign > 0 else np.subtract(self.log_p, grid, out=self.log_p)

    def add_data_point(self, data_point):
        """tree_node.py:46-52: both tiles move by the point's grid (log_r stays valid for an unchanged subtree)."""
        self._claim((data_point.idx,))
        self._shift(data_point.value, +1)
        np.add(self.log_r, data_point.value, out=self.log_r)

    def add_data_point_list(self, data_point_list):
        """tree_node.py:34-44, one point after the other (the floating-point sum follows the list order)."""
        self._claim(dp.idx for dp in data_point_list)
        for dp in data_point_list:
            self._shift(dp.value, +1)
            np.add(self.log_r, dp.value, out=self.log_r)

    def remove_data_point(self, data_point):
        """tree_node.py:54-59: only log_p is corrected; the caller refreshes log_r along the path to the root."""
        assert data_point.idx in self.data_points
        self.data_points.remove(data_point.idx)
        self._shift(data_point.value, -1)

    # ------------------------------------------------------------------ likelihood
    def update_node_from_child_r_vals(self, child_log_r_values):
        """tree_node.py:61-71: log_r = log_p for a leaf, else log_p + prefix-LSE of the left fold of the children."""
        n_children = len(child_log_r_values)
        if n_children:
            stacked = np.asarray(child_log_r_values)
            self.log_r[...] = ops.node_update_batch(self.log_p[None], [stacked])[0]
        else:
            self.log_r[...] = self.log_p

    # ------------------------------------------------------------------ value semantics
    def copy(self):
        twin = TreeNode.__new__(type(self))
        twin.node_id = self.node_id if isinstance(self.node_id, str) else int(self.node_id)
        twin.data_points = set(self.data_points)
        twin.log_p, twin.log_r = np.array(self.log_p), np.array(self.log_r)
        return twin

    __copy__ = copy

    def __eq__(self, other):
        return all(np.array_equal(getattr(self, f), getattr(other, f)) for f in ("log_p", "log_r"))

    __hash__ = None

    def to_dict(self):
        return dict(log_p=self.log_p, log_R=self.log_r, node_id=self.node_id)

    def serialize(self):
        return dict(node_id=str(self.node_id), data_points=str(list(self.data_points)))
