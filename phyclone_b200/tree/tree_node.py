"""Drop-in for the reference's `TreeNode` (tree/tree_node.py:8-80) on host buffers.

The sampler of this package never builds these objects (its trees hold tile ids into the device arena, see
tree/tree.py); the class exists for code written against the reference seam: `log_p` / `log_r` are NumPy tiles owned
by the node, the data-point bookkeeping is the reference's, and the one expensive method --
`update_node_from_child_r_vals`, the fold + scan of the children -- is a single fused launch through the C ABI
(`pcb_node_update_host`).  Children are folded in the order given (no order-insensitive LRU).
"""

import numpy as np

from .. import ops


class TreeNode(object):
    __slots__ = ("log_p", "log_r", "node_id", "data_points")

    def __init__(self, grid_size, log_prior, node_id):
        self.log_p = np.full(grid_size, log_prior, order="C")
        self.log_r = np.zeros(grid_size, order="C")
        self.node_id = node_id
        self.data_points = set()

    def __copy__(self):
        new = self.__class__.__new__(self.__class__)
        new.log_p = self.log_p.copy()
        new.log_r = self.log_r.copy()
        new.node_id = str(self.node_id) if isinstance(self.node_id, str) else int(self.node_id)
        new.data_points = self.data_points.copy()
        return new

    def __eq__(self, other):
        return np.array_equal(self.log_p, other.log_p) and np.array_equal(self.log_r, other.log_r)

    def add_data_point_list(self, data_point_list):
        dp_idx_set = {dp.idx for dp in data_point_list}
        assert self.data_points.isdisjoint(dp_idx_set)
        self.data_points.update(dp_idx_set)
        for data_point in data_point_list:
            self.log_p += data_point.value
            self.log_r += data_point.value

    def add_data_point(self, data_point):
        assert data_point.idx not in self.data_points
        self.data_points.add(data_point.idx)
        self.log_p += data_point.value
        self.log_r += data_point.value

    def remove_data_point(self, data_point):
        assert data_point.idx in self.data_points
        self.data_points.discard(data_point.idx)
        self.log_p -= data_point.value

    def update_node_from_child_r_vals(self, child_log_r_values):
        if len(child_log_r_values) == 0:
            np.copyto(self.log_r, self.log_p)
            return
        np.copyto(self.log_r, ops.node_update_batch(self.log_p[None], [np.asarray(child_log_r_values)])[0])

    def copy(self):
        return self.__copy__()

    def to_dict(self):
        return {"log_p": self.log_p, "log_R": self.log_r, "node_id": self.node_id}

    def serialize(self):
        return {"node_id": str(self.node_id), "data_points": str(list(self.data_points))}
