"""CPU: the drop-in boundary itself.  The C-ABI library must load without a GPU and export exactly what
include/phyclone_b200.h declares; the ctypes table must cover the same set; the product must not reach for the
oracle or any CPU fallback."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "phyclone_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(pcb_[A-Za-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    from phyclone_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else _lib.load()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_the_header():
    from phyclone_b200 import _lib

    names = declared_symbols()
    assert set(_lib.SIGNATURES) == names, (sorted(set(_lib.SIGNATURES) - names), sorted(names - set(_lib.SIGNATURES)))
    lib = _lib.load()
    assert lib.pcb_version() >= 100  # no CUDA call behind this one


def test_layout_struct_mirrors_the_header():
    """pcb_make_layout is pure host code: the ctypes mirror of pcb_layout must agree with what the library fills in."""
    from phyclone_b200 import _lib

    lib = _lib.load()
    for S, G, variant, kernel in [(8, 101, -1, 0), (8, 101, -2, 0), (4, 101, -1, 1), (20, 201, -1, 0), (8, 101, 11304, 1)]:
        lay = _lib.Layout()
        _lib.check(lib.pcb_make_layout(S, G, variant, ctypes.byref(lay)))
        assert (lay.S, lay.G, lay.kernel) == (S, G, kernel)
        assert lay.n_strips == -(-G // lay.strip) and lay.lead == lay.strip
        assert lay.pitch >= lay.lead + (lay.n_strips + 1) * lay.strip
        assert lay.tile_elems == S * lay.pitch and (lay.tile_elems * 8) % 16 == 0  # TMA: 16-byte multiples
        assert (lay.n_strips + 1) // 2 <= lay.lanes_per_row
    # grids beyond 1088: the 17-wide strip with several strip pairs per lane (`reserved` = pairs per lane)
    for G in (1089, 2001, 5000):
        lay = _lib.Layout()
        _lib.check(lib.pcb_make_layout(8, G, -1, ctypes.byref(lay)))
        assert (lay.strip, lay.lanes_per_row, lay.kernel) == (17, 32, 0)
        assert lay.reserved == -(-((lay.n_strips + 1) // 2) // 32) and lay.reserved >= 2
        assert lay.pitch >= lay.lead + (lay.n_strips + 1) * lay.strip and (lay.tile_elems * 8) % 16 == 0
    bad = _lib.Layout()
    assert lib.pcb_make_layout(8, 9000, -1, ctypes.byref(bad)) != 0
    assert b"layout" in lib.pcb_last_error()


def test_host_extension_and_draw_library_load():
    from phyclone_b200 import _hostlib
    from phyclone_b200._hostmod import load

    hm = load()
    for name in ("TileBook", "ForestState", "Holder", "Particle", "forest_state", "attach_holders", "edit_holder",
                 "finish_particles", "particle_log_ws", "gibbs_scores", "log_sum_exp", "edge_hash", "data_hash"):
        assert hasattr(hm, name), name
    assert _hostlib.load().pcb_host_version() >= 1
    x = np.array([-3.0, -1.0, -2.5])
    assert hm.log_sum_exp(x) == pytest.approx(float(np.log(np.exp(x).sum())), rel=1e-15)
    assert hm.log_sum_exp(np.array([-np.inf, -np.inf])) == -np.inf


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phyclone_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "/root/reference" in text:
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present: the failure path cannot be exercised here")
    from phyclone_b200 import ops

    with pytest.raises(RuntimeError):
        ops.np_conv_dims(np.zeros((2, 5)), np.zeros((2, 5)))
    with pytest.raises(RuntimeError):
        ops.map_trees(np.zeros((1, 2, 5)), [0, 1], [0], [0, 0], [])


def test_long_holder_and_particle_chains_are_released_iteratively():
    """Genealogies are as long as the data set; dropping the last reference to a 100 000-long chain of native holders /
    particles must not recurse on the C stack."""
    from phyclone_b200._hostmod import load
    from phyclone_b200.data.base import DataPoint
    from phyclone_b200.smc.forest_state import OUTLIER, ForestState, ensure_tables
    from phyclone_b200.smc.swarm import StateHolder
    from phyclone_b200.tree.distributions import FSCRPDistribution, TreeJointDistribution
    from tests.null_store import NullTileStore

    hm = load()
    vals = np.zeros((2, 2, 8))
    store = NullTileStore(vals)
    tree_dist = TreeJointDistribution(FSCRPDistribution(1.0))
    ensure_tables(tree_dist.prior, 10)
    dp = DataPoint(0, vals[0], outlier_prob=-1.0, outlier_prob_not=-0.1, outlier_marginal_prob=-3.0)
    empty = ForestState.empty(store)
    holder, particle = None, None
    for _ in range(100000):
        holder = StateHolder.edit(store, holder, empty, empty, OUTLIER, dp, None, tree_dist)
        particle = hm.Particle(0.0, particle, holder)
    assert particle.parent_particle is not None and holder.kind == OUTLIER
    del holder, particle
