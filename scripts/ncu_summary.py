"""Summarise an .ncu-rep (first profiled launch): key raw metrics, opcode mix, stall mix.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_blocks", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    print("# ncu summary of %s\n" % rep, file=out)
    for li, r in enumerate(raw[2:4]):
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("## launch %d: %s\n" % (li, name), file=out)
        print("| metric | value | unit |\n|---|---|---|", file=out)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("| %s | %s | %s |" % (k, r[i], units[i]), file=out)
        print(file=out)
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    sec = None
    for r in src:
        if r and r[0] == "Kernel Name":
            if sec is not None:
                break
            sec = []
            continue
        if sec is not None:
            sec.append(r)
    h = sec[0]
    ia, ie = h.index("Source"), h.index("Instructions Executed")
    ops, stalls = collections.Counter(), collections.Counter()
    st_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    for r in sec[1:]:
        try:
            n = int(r[ie])
        except (ValueError, IndexError):
            continue
        t = r[ia].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        ops[op] += n
        for i, c in st_cols:
            try:
                stalls[c] += int(r[i])
            except ValueError:
                pass
    tot = sum(ops.values())
    print("## opcode mix (warp instructions executed, launch 0; total %d)\n" % tot, file=out)
    print("| opcode | executed | share |\n|---|---|---|", file=out)
    for op, n in ops.most_common(14):
        print("| %s | %d | %.1f%% |" % (op, n, 100.0 * n / tot), file=out)
    ts = sum(stalls.values()) or 1
    print("\n## warp stall samples (launch 0)\n\n| reason | share |\n|---|---|", file=out)
    for c, n in stalls.most_common(10):
        print("| %s | %.1f%% |" % (c, 100.0 * n / ts), file=out)


if __name__ == "__main__":
    main()
