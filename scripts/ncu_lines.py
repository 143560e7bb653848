"""Stall samples and executed instructions per CUDA source line of one kernel of an .ncu-rep (needs -lineinfo):
joins `ncu --page source --csv` (SASS) with `nvdisasm -g` line info of the built library.
usage: python scripts/ncu_lines.py report.ncu-rep mangled_kernel_name [top_n]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "phyclone_b200", "lib", "libphyclone_b200.so")


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    tmp = "/tmp/ncu_lines"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kernel + ":"))
    line, off2line = None, {}
    for l in dis[start + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            off2line[int(m.group(1), 16)] = line
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    hdr, data = None, []
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r
            continue
        if hdr is not None and r and r[0].startswith("0x"):
            data.append(r)
    base = int(data[0][0], 16)
    isamp, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    st = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    samp, ex, why = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for r in data:
        ln = off2line.get(int(r[0], 16) - base)
        samp[ln] += int(r[isamp] or 0)
        ex[ln] += int(r[iex] or 0)
        for i, c in st:
            why[ln][c.replace("stall_", "")] += int(r[i] or 0)
    tot = sum(samp.values())
    files = {}
    print("total samples %d, instructions %d" % (tot, sum(ex.values())))
    for ln, n in samp.most_common(top):
        f, l = ln if ln else ("?", 0)
        if f not in files and os.path.exists(f):
            files[f] = open(f).read().split("\n")
        text = files[f][l - 1].strip()[:72] if f in files and l else ""
        print("%5.1f%% %8d exec %s:%d %s | %s" % (100 * n / max(tot, 1), ex[ln], os.path.basename(f)[:14], l, text,
                                                   dict(why[ln].most_common(2))))


if __name__ == "__main__":
    main()
