"""CPU: the C sampling draws (phyclone_b200/csrc/pcb_host.cpp) consume numpy's bit generator exactly like
the Generator methods the reference calls: same values AND same generator state afterwards."""

import ctypes as C

import numpy as np

from phyclone_b200 import _hostlib


def _addr(g):
    return C.c_void_p(_hostlib.bitgen_address(g))


def test_scalar_draws_match_generator_methods():
    lib = _hostlib.load()
    for seed in range(40):
        a, b = np.random.default_rng(seed), np.random.default_rng(seed)
        a.random()  # desynchronise the internal uint32 buffer now and then
        b.random()
        for rep in range(30):
            assert lib.pcb_draw_random(_addr(a)) == b.random()
            high = int(b.integers(1, 40))
            lib.pcb_draw_integers(_addr(a), 40 - 1 + 1) if False else a.integers(1, 40)
            assert lib.pcb_draw_integers(_addr(a), high + 1) == int(b.integers(0, high + 1))
            pop = int(b.integers(1, 20))
            a.integers(1, 20)
            k = int(b.integers(1, pop + 1))
            a.integers(1, pop + 1)
            idx = np.zeros(k, dtype=np.int64)
            lib.pcb_draw_choice(_addr(a), pop, k, idx.ctypes.data_as(C.c_void_p))
            roots = np.arange(100, 100 + pop)
            np.testing.assert_array_equal(roots[idx], b.choice(roots, k, replace=False))
            d = int(b.integers(1, 12))
            a.integers(1, 12)
            p = b.dirichlet(np.ones(d) * 0.3)
            a.dirichlet(np.ones(d) * 0.3)
            for n in (1, 15, 99, 499):
                out = np.zeros(d, dtype=np.int64)
                lib.pcb_draw_multinomial(_addr(a), n, p.ctypes.data_as(C.c_void_p), d, out.ctypes.data_as(C.c_void_p))
                np.testing.assert_array_equal(out, b.multinomial(n, p))
        assert a.random() == b.random()


def test_smc_step_draws_match_python_loop():
    lib = _hostlib.load()
    rs = np.random.default_rng(99)
    for trial in range(60):
        n_dists = int(rs.integers(1, 6))
        qs, roots, empty = [], [], []
        for d in range(n_dists):
            R = int(rs.integers(0, 9))
            emp = R == 0 or rs.random() < 0.15
            roots.append(np.sort(rs.choice(50, R, replace=False)).astype(np.int64))
            nq = (R + (1 if emp else 0)) or 1
            qs.append(rs.dirichlet(np.ones(nq)))
            empty.append(1 if (emp or R == 0) else 0)
        n = int(rs.integers(1, 120))
        dist_of = rs.integers(0, n_dists, size=n).astype(np.int32)
        seed = int(rs.integers(1 << 30))
        # python loop (the reference's calls)
        g = np.random.default_rng(seed)
        ref = []
        for i in range(n):
            d = dist_of[i]
            if empty[d] or g.random() < 0.5:
                ref.append((0, int(g.multinomial(1, qs[d]).argmax()), ()))
            else:
                k = g.integers(0, len(roots[d]) + 1)
                kids = [] if k == 0 else g.choice(roots[d], k, replace=False)
                ref.append((1, int(k), tuple(int(x) for x in kids)))
        # C
        h = np.random.default_rng(seed)
        q_flat = np.concatenate(qs)
        q_off = np.cumsum([0] + [len(q) for q in qs]).astype(np.int64)
        r_flat = np.concatenate(roots + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        r_off = np.cumsum([0] + [len(r) for r in roots]).astype(np.int64)
        emp = np.array(empty, dtype=np.uint8)
        kind = np.zeros(n, dtype=np.int32)
        index = np.zeros(n, dtype=np.int64)
        kids = np.zeros(n * 8 + 8, dtype=np.int64)
        koff = np.zeros(n + 1, dtype=np.int64)
        P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = lib.pcb_smc_draws(_addr(h), n, P(dist_of), P(q_flat), P(q_off), P(r_flat), P(r_off), P(emp), P(kind), P(index),
                               P(kids), P(koff))
        assert rc == 0
        got = [(int(kind[i]), int(index[i]), tuple(int(x) for x in kids[koff[i]:koff[i + 1]])) for i in range(n)]
        assert got == ref
        assert h.random() == g.random()
