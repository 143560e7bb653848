"""Final data point (reference data/base.py:5-55): idx, value [S,G] (host copy of the grid; the device
copy lives in the TileStore), name, outlier log-probabilities and the outlier marginal."""

import numpy as np

from .. import ops


class DataPoint(object):
    __slots__ = ("idx", "value", "name", "outlier_prob", "outlier_marginal_prob", "outlier_prob_not", "_t", "_ts")

    def __init__(self, idx, value, name=None, outlier_prob=0, outlier_prob_not=1, outlier_marginal_prob=None):
        self.idx = idx
        self.value = value
        self._t = self._ts = None  # host tuple cached by smc/forest_state.py for one tile store
        self.name = idx if name is None else name
        self.outlier_prob = outlier_prob
        self.outlier_prob_not = outlier_prob_not
        if outlier_marginal_prob is None:  # data/base.py:31-37, on the device
            outlier_marginal_prob = float(ops.outlier_marginal(np.asarray(value)))
        self.outlier_marginal_prob = outlier_marginal_prob

    def __reduce__(self):  # the cached host tuple points at a tile store: never pickled
        return (DataPoint, (self.idx, self.value, self.name, self.outlier_prob, self.outlier_prob_not,
                            self.outlier_marginal_prob))

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return self.name == other.name

    @property
    def grid_size(self):
        return self.value.shape

    @property
    def shape(self):
        return self.value.shape


def make_data_points(values, names=None, outlier_probs=None):
    """Build DataPoints for [T,S,G] grids with ONE batched outlier-marginal launch."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    marg = ops.outlier_marginal(values)
    out = []
    for i in range(values.shape[0]):
        op, opn = (0, 1) if outlier_probs is None else outlier_probs[i]
        out.append(DataPoint(i, values[i], None if names is None else names[i], op, opn, float(marg[i])))
    return out
