"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`)
into markdown: per-kernel totals / shares, and the kernels of the longest SMC step of a device-resident sweep.
usage: python scripts/launch_list_summary.py gpurun_out/launches.csv "command that was profiled" > profiles/rN_launch_list.md"""
import collections
import csv
import re
import sys

SETUP = ("fp64_peak_kernel", "vectorized_elementwise_kernel", "likelihood_grid_kernel", "import_tiles_kernel",
         "outlier_marginal_rows_kernel", "cluster_sum_kernel", "reduce_rows_kernel")


def main():
    path = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else "bench.py"
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    order = []
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        order.append((name, v, row.get("Grid Size", "")))
    agg = collections.defaultdict(list)
    for name, v, _ in order:
        agg[name].append(v)
    step = {k: v for k, v in agg.items() if not any(s in k for s in SETUP)}
    tot = sum(sum(v) for v in step.values())
    print("# Launch list of `%s`\n" % what)
    print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv` on one B200; %d launches captured.  Per-launch "
          "times are cold-cache and serialised by ncu: the SHARES are the evidence, not the absolute values.  Set-up "
          "kernels (FP64 peak microbenchmark, likelihood grid, imports, torch fills) are listed but left out of the "
          "shares.\n" % len(order))
    print("| kernel | launches | total us | share of the steps | avg us | median us | max us |\n|---|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        xs = sorted(v)
        share = "%.1f %%" % (100 * sum(v) / tot) if k in step else "(set-up)"
        print("| `%s` | %d | %.0f | %s | %.1f | %.1f | %.1f |" % (k[:64], len(v), sum(v), share, sum(v) / len(v),
                                                                  xs[len(xs) // 2], xs[-1]))
    idx = [i for i, (n, _, _) in enumerate(order) if "weigh_kernel" in n]
    if len(idx) > 2:
        spans = [(sum(v for _, v, _ in order[a + 1:b + 1]), a, b) for a, b in zip(idx[:-1], idx[1:])
                 if all("sweep::" in n or "node_jobs" in n for n, _, _ in order[a + 1:b + 1])]
        if spans:
            t, a, b = max(spans)
            print("\nLongest SMC step of a device-resident sweep in the capture (%.0f us of serialised launches):\n" % t)
            print("| kernel | us | grid |\n|---|---|---|")
            for n, v, g in order[a + 1:b + 1]:
                print("| `%s` | %.1f | %s |" % (n[:64], v, g))
    jobs = sum(sum(v) for k, v in step.items() if "node_jobs" in k)
    ctl = sum(sum(v) for k, v in step.items() if "sweep::" in k)
    print("\n`node_jobs_kernel` (all instantiations) = %.1f %% of the steps' device time; the sweep's control kernels "
          "(build / add_list / decide / weigh) = %.1f %%; tile-store kernels (tile_add, reduce_rows2) = the rest."
          % (100 * jobs / tot, 100 * ctl / tot))


if __name__ == "__main__":
    main()
