"""Helpers shared by the MAP-assignment tests: read a tests/golden/map_*.npz case."""

import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = -7  # stand-in for the "root" name inside the integer arrays of the fixtures


def case_paths():
    return sorted(glob.glob(os.path.join(HERE, "golden", "map_*.npz")))


def load_case(path):
    z = np.load(path)
    c = {k: z[k] for k in z.files}
    N = len(c["names"])
    c["children"] = [list(map(int, c["child_idx"][c["child_off"][v]:c["child_off"][v + 1]])) for v in range(N)]
    post, stack = [], [(0, False)]  # node 0 is the dummy root
    while stack:
        v, seen = stack.pop()
        if seen:
            post.append(v)
        else:
            stack.append((v, True))
            stack.extend((k, False) for k in c["children"][v])
    c["post_order"] = post
    return c


def tree_dict(c, data_points, prior):
    """The reference `Tree.to_dict()` of the case, rebuilt around the given DataPoint objects."""
    name = lambda k: "root" if k == ROOT else int(k)  # noqa: E731
    node_idx = {name(k): int(v) for k, v in zip(c["node_idx_keys"], c["node_idx_vals"])}
    node_data = {}
    for k, idx in c["node_data"]:
        node_data.setdefault(int(k), []).append(data_points[int(idx)])
    S, G = c["values"].shape[1:]
    return {"graph": [tuple(map(int, e)) for e in c["edges"]], "node_idx": node_idx,
            "node_idx_rev": {v: k for k, v in node_idx.items()}, "node_data": node_data, "grid_size": (S, G),
            "node_last_added_to": None, "log_prior": prior}
