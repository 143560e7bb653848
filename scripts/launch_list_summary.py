"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python scripts/launch_list_summary.py gpurun_out/launches.csv [out.md] [title]"""
import collections
import csv
import re
import sys


def main():
    src = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    title = sys.argv[3] if len(sys.argv) > 3 else src
    with open(src) as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    h = rows[0]
    ik, iv, ig, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0, 0])
    n = 0
    for row in rows[1:]:
        if len(row) <= iv:
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", row[ik]).replace("void ", "").replace("pcb::", ""))
        try:
            t = float(row[iv].replace(",", ""))
        except ValueError:
            continue
        scale = 1e-3 if row[iu] == "ns" else 1.0
        g = int(re.sub(r"[^0-9]", "", row[ig].split(",")[0]) or 0)
        a = agg[name]
        a[0] += 1
        a[1] += t * scale
        a[2] += g
        n += 1
    tot = sum(a[1] for a in agg.values())
    print("# ncu launch list of %s\n" % title, file=out)
    print("%d launches, serialised with cold caches: compare shares, not absolutes.\n" % n, file=out)
    print("| kernel | launches | total us | share | avg us | avg grid |\n|---|---|---|---|---|---|", file=out)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.0f | %.1f%% | %.1f | %d |" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2] // a[0]), file=out)
    chain = {k: a for k, a in agg.items() if k in ("node_jobs_kernel", "node_jobs_wide_kernel", "tile_add_kernel", "reduce_rows_kernel", "reduce_rows2_kernel")}
    ct = sum(a[1] for a in chain.values())
    print("\ntotal device time %.1f ms; the chain's own kernels (node_jobs, tile_add, reduce_rows) %.1f ms: %s" % (
        tot / 1e3, ct / 1e3, ", ".join("%s %.1f %%" % (k, 100 * a[1] / ct) for k, a in sorted(chain.items(), key=lambda kv: -kv[1][1]))),
        file=out)


if __name__ == "__main__":
    main()
