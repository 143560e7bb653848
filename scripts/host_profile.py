"""Profile the HOST side of the sampler on a CPU box (no device): tests/null_store.NullTileStore stands in for the
tile store, so decisions are arbitrary but the bookkeeping cost per extension / candidate is the real one.
usage: python scripts/host_profile.py [--profile]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyclone_b200.data.base import DataPoint  # noqa: E402
from phyclone_b200.run import Chain  # noqa: E402
from tests.null_store import NullTileStore  # noqa: E402


def main():
    T, S, G = 50, 8, 101
    vals = np.zeros((T, S, G))
    data = [DataPoint(i, vals[i], outlier_marginal_prob=-100.0) for i in range(T)]
    chain = Chain(data, np.random.default_rng(1), num_particles=100, store=NullTileStore(vals))
    chain.burnin_step()
    chain.step()

    def go():
        for _ in range(3):
            chain.step()

    t0 = time.time()
    if "--profile" in sys.argv:
        pr = cProfile.Profile()
        pr.runcall(go)
        pstats.Stats(pr).sort_stats("tottime").print_stats(35)
    else:
        go()
    print("host seconds per iteration: %.3f" % ((time.time() - t0) / 3), chain.stats)


if __name__ == "__main__":
    main()
