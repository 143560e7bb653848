"""Trace files: gzip + pickle with the reference's schema (process_trace.py:386-393)

    {chain_num: {"data": [DataPoint], "samples": [str], "trace": [entry], "chain_num": int, "clusters": DataFrame}}
    entry = {"iter", "time", "alpha", "log_p_one", "tree": Tree.to_dict()}

Written so that upstream `phyclone map / consensus / topology-report` can open them -- data points are pickled
under the reference's class path `phyclone.data.base.DataPoint` with its slot layout (data/base.py:5-29) -- and read
so that upstream traces open here: that class path, and the `rustworkx.EdgeList` a real upstream trace holds in
`tree["graph"]`, are mapped to local types without importing either package.
"""

import gzip
import io
import pickle

from ..data.base import DataPoint

_REF_SLOTS = ("idx", "value", "name", "outlier_prob", "outlier_marginal_prob", "outlier_prob_not")


class _RefDataPoint(object):
    """Pickles exactly like an instance of the reference's slotted class."""

    __slots__ = _REF_SLOTS


_RefDataPoint.__module__ = "phyclone.data.base"
_RefDataPoint.__qualname__ = _RefDataPoint.__name__ = "DataPoint"


class _EdgeList(list):
    """Stand-in for `rustworkx.EdgeList` when reading upstream traces: an iterable of (src, dst) pairs."""

    def __setstate__(self, state):
        self[:] = [tuple(e) for e in state]


class _RefModule(object):
    DataPoint = _RefDataPoint


def _reduce_data_point(obj):
    # object.__new__(phyclone.data.base.DataPoint) + slot state: what the reference's own pickles rebuild
    return (object.__new__, (_RefDataPoint,), (None, {k: getattr(obj, k) for k in _REF_SLOTS}))


def _dumps(results):
    import copyreg
    import sys

    # the class global must resolve while dumping; the reference package itself is never imported
    saved = {k: sys.modules.get(k) for k in ("phyclone", "phyclone.data", "phyclone.data.base")}
    try:
        for k in saved:
            sys.modules[k] = _RefModule
        buf = io.BytesIO()
        p = pickle.Pickler(buf, protocol=4)
        p.dispatch_table = copyreg.dispatch_table.copy()
        p.dispatch_table[DataPoint] = _reduce_data_point
        p.dispatch_table[_LoadedDataPoint] = _reduce_data_point
        p.dump(results)
        return buf.getvalue()
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "phyclone.data.base" and name == "DataPoint":
            return _LoadedDataPoint
        if module.split(".")[0] == "rustworkx" and name == "EdgeList":
            return _EdgeList
        return super().find_class(module, name)


class _LoadedDataPoint(DataPoint):
    """DataPoint restored from slot state (no __init__ call, so no device work at load time)."""

    __slots__ = ()

    def __setstate__(self, state):
        slots = state[1] if isinstance(state, tuple) else state
        for k, v in slots.items():
            setattr(self, k, v)
        self._t = self._ts = None


def write_trace(results, out_file):
    with gzip.GzipFile(out_file, mode="wb") as fh:
        fh.write(_dumps(results))


def read_trace(in_file):
    with gzip.GzipFile(in_file, "rb") as fh:
        return _Unpickler(fh).load()
