"""Exact FS-CRP posterior over forests of a handful of data points, by enumeration (test infrastructure).

Follows the reference's own check (phyclone/tests/exact_posterior.py:11-36 + test_pg_posterior.py:15-158): every
ordered partition of the data into nodes x every rooted forest over those nodes, scored with `log_p_one`, keyed by
clade set; and its data simulator (phyclone/tests/simulate.py:8-33: binomial log-likelihood over a 101-point grid)."""

import itertools

import numpy as np
import scipy.stats as stats

from oracle import numerics as onum
from oracle import sampler as osm


def binomial_datum(idx, n, p, rng, eps=1e-10, grid_size=101):
    """phyclone/tests/simulate.py:8-33"""
    p = np.atleast_1d(p)
    grid = np.linspace(0 + eps, 1 - eps, grid_size)
    rows = []
    for p_i in p:
        x = stats.binom.rvs(n, p_i, random_state=rng)
        rows.append(stats.binom.logpmf(x, n, grid))
    return osm.Datum(idx, np.atleast_2d(rows))


def _partitions(items):
    items = list(items)
    if len(items) == 1:
        yield [items]
        return
    first = items[0]
    for smaller in _partitions(items[1:]):
        for n, subset in enumerate(smaller):
            yield smaller[:n] + [[first] + subset] + smaller[n + 1:]
        yield [[first]] + smaller


def _forests(n):
    """All parent vectors (parent index or -1) over n labelled nodes without cycles."""
    for par in itertools.product(range(-1, n), repeat=n):
        ok = True
        for i in range(n):
            seen, j = set(), i
            while j != -1 and ok:
                if j in seen:
                    ok = False
                seen.add(j)
                j = par[j]
            if not ok:
                break
        if ok:
            yield par


def exact_posterior(data, alpha=1.0):
    tm = osm.TileMath()
    joint = osm.Joint(alpha)
    log_p = {}
    for blocks in _partitions(data):
        n = len(blocks)
        for par in _forests(n):
            t = osm.Forest(data[0].grid_size, tm)
            made = {}

            def build(i):
                if i in made:
                    return made[i]
                kids = [build(c) for c in range(n) if par[c] == i]
                node = t.new_root(children=kids)
                for dp in blocks[i]:
                    t.attach(dp, node)
                made[i] = node
                return node

            for i in range(n):
                if par[i] == -1:
                    build(i)
            key = t.clades()
            if key not in log_p:
                log_p[key] = joint.log_p_one(t)
    keys = list(log_p)
    p, _ = onum.exp_normalize(np.array([log_p[k] for k in keys]))
    return dict(zip(keys, p))


CASES = {
    "two_points_non_informative": lambda rng: [binomial_datum(0, 0, 1.0, rng), binomial_datum(1, 0, 1.0, rng)],
    "three_points_non_informative": lambda rng: [binomial_datum(i, 0, 1.0, rng) for i in range(3)],
    "two_points_single_cluster": lambda rng: [binomial_datum(0, 10, 1.0, rng), binomial_datum(1, 10, 1.0, rng)],
    "two_points_two_clusters": lambda rng: [binomial_datum(0, 10, 1.0, rng), binomial_datum(1, 10, 0.5, rng)],
    "two_points_2d_two_clusters": lambda rng: [binomial_datum(0, 100, [1.0, 1.0], rng), binomial_datum(1, 100, [0.5, 0.7], rng)],
}
