"""Where the host time of a device-mode chain goes: wall-clock timers around the sampler's stages (config #2).
usage: python scripts/host_split.py [iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from phyclone_b200 import ops  # noqa: E402
from phyclone_b200.data.base import make_data_points  # noqa: E402
from phyclone_b200.run import Chain  # noqa: E402

T = {}


def timed(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def wrap(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T[label] = T.get(label, 0.0) + time.perf_counter() - t0

    setattr(obj, name, wrap)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    bench.set_config(2)
    values = bench.workload_values_gpu(ops)
    chain = Chain(make_data_points(values), bench.chain_rng(0, 1), num_particles=100, sweep="device")
    from phyclone_b200.smc import device_sweep as ds

    timed(chain.dp_sampler, "sample_tree", "gibbs data-point sweep")
    timed(chain.prg_sampler, "sample_tree", "prune-regraph")
    timed(chain.tree_sampler, "sample_tree", "PG sweep (host + device)")
    timed(ds.RetainedPath, "__init__", "  retained path states")
    timed(ds.RetainedPath, "scores", "  retained path scores (flush)")
    timed(ds.DeviceSweep, "_rebuild", "  lineage replay")
    timed(ds.DeviceSweep, "_pack_states", "  pack states")
    timed(chain.store, "flush", "tile-store flushes (inside the above)")
    chain.burnin_step()
    for _ in range(5):
        chain.step()
    T.clear()
    p0 = ops.plan_stats()
    s0 = chain.stats["sweep"]["device_ms"]
    t0 = time.perf_counter()
    for _ in range(iters):
        chain.step()
    wall = time.perf_counter() - t0
    p1 = ops.plan_stats()
    print("ms per iteration: %.2f" % (1e3 * wall / iters))
    for k, v in T.items():
        print("  %-42s %.2f" % (k, 1e3 * v / iters))
    print("  %-42s %.2f" % ("device sweep (CUDA events)", (chain.stats["sweep"]["device_ms"] - s0) / iters))
    print("  %-42s %.2f / %.2f" % ("flush enqueue / wait", 1e3 * (p1[0] - p0[0]) / iters, 1e3 * (p1[1] - p0[1]) / iters))
    print("  moves %d stays %d look-ahead hits %d" % (chain.dp_sampler.moves, chain.dp_sampler.stays,
                                                      chain.dp_sampler.lookahead_hits))


if __name__ == "__main__":
    main()
