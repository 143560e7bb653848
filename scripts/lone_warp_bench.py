"""One small launch of the job kernel (default: 100 jobs x 8 children, the latency layout): what a fold step costs in
the lane-per-unit mapping and in the warp-per-row ("latency") mapping.
usage: python scripts/lone_warp_bench.py [variant] [jobs] [children]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_bench import bench_variant  # noqa: E402

v = int(sys.argv[1]) if len(sys.argv) > 1 else 708
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for c in ([int(sys.argv[3])] if len(sys.argv) > 3 else [2, 4, 8, 16]):
    for lat in (False, True):
        r = bench_variant(8, 101, v, n, c, 2048, 20, latency=lat)
        print("variant %d jobs %d children %d %s: %.1f us per launch, %.2f us per fold step" % (
            v, n, c, "warp-per-row " if lat else "lane-per-unit", r["ms"] * 1e3, r["ms"] * 1e3 / max(c - 1, 1)))
