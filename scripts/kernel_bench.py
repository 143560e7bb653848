"""Kernel-level throughput of the node-refresh job kernel on large synthetic batches
(SURVEY.md 8d micro-inputs): algorithmic FP64 FMA rate vs the measured FMA-chain peak.

usage: python scripts/kernel_bench.py [--S 8 --G 101 --jobs 32768 --children 2,4,8 --variants 708,1304,...]
"""

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyclone_b200 import engine, ops  # noqa: E402


def bench_variant(S, G, variant, n_jobs, C, n_tiles, reps, score_only=True, latency=False):
    """score_only=True: dummy-root jobs (scores, no tile written, no log_p read).  score_only=False: full node
    refreshes -- every job reads a log_p tile and writes its log_r tile (log + lin + row maxima) to its own slot."""
    ar = engine.DeviceArena(S, G, n_tiles + 1 + (0 if score_only else n_jobs), variant=variant)
    if latency:
        ar.set_latency(True)  # the warp-per-row mapping for launches of <= PCB_ROWS_MAX_JOBS jobs
    rng = np.random.default_rng(0)
    dense = torch.log(torch.rand((n_tiles, S, G), dtype=torch.float64, device=ar.device) * 999.9 + 0.1)
    ar.import_tiles(dense, np.arange(n_tiles))
    kids = rng.integers(0, n_tiles, size=n_jobs * C).astype(np.int32)
    jobs = np.zeros((n_jobs, 8), dtype=np.int32)
    jobs[:, 0] = np.arange(n_jobs) * C
    jobs[:, 1] = C
    if score_only:
        jobs[:, 2] = -1
        jobs[:, 3] = -1
        jobs[:, 4] = engine.JOB_SCAN | engine.JOB_ADD_PRIOR
        jobs[:, 5] = np.arange(n_jobs)
    else:
        jobs[:, 2] = rng.integers(0, n_tiles, size=n_jobs)
        jobs[:, 3] = n_tiles + 1 + np.arange(n_jobs)
        jobs[:, 4] = engine.JOB_SCAN
        jobs[:, 5] = -1
    for _ in range(3):
        ar.run_jobs(jobs, kids, n_slots=n_jobs)
    torch.cuda.synchronize()
    # time the job kernel alone: enqueue through the C ABI between two events
    import ctypes as Ct

    from phyclone_b200 import _lib

    jd = ar._to_dev_i32(jobs, "jobs")
    cd = ar._to_dev_i32(kids, "kids")
    rows_l = torch.empty((n_jobs, S), dtype=torch.float64, device=ar.device)
    rows_t = torch.empty((n_jobs, S), dtype=torch.float64, device=ar.device)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(ar.lib.pcb_node_jobs_dev(Ct.byref(ar.c), jd.data_ptr(), cd.data_ptr(), n_jobs, ar.prior,
                                            rows_l.data_ptr(), rows_t.data_ptr(), ar._stream()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fma = n_jobs * (C - 1) * S * G * (G + 1) / 2
    lay = ar.layout
    return {
        "variant": variant, "L": lay.strip, "P": lay.lanes_per_row, "pitch": lay.pitch, "RB": lay.rows_per_cta,
        "S": S, "G": G, "jobs": n_jobs, "children": C, "ms": ms, "score_only": bool(score_only),
        "tflops": 2 * fma / (ms * 1e-3) / 1e12, "convs_per_s": n_jobs * (C - 1) / (ms * 1e-3),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=8)
    ap.add_argument("--G", type=int, default=101)
    ap.add_argument("--jobs", type=int, default=32768)
    ap.add_argument("--children", default="2,4,8")
    ap.add_argument("--variants", default="-1,708,716,908,1304,1308")
    ap.add_argument("--tiles", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    peak, _ = ops.fp64_peak(8192)
    peak2, ms = ops.fp64_peak(8192)
    print(json.dumps({"fp64_peak_tflops": max(peak, peak2) / 1e12, "ms": ms}))
    for v in [int(x) for x in args.variants.split(",")]:
        for C in [int(x) for x in args.children.split(",")]:
            try:
                r = bench_variant(args.S, args.G, v, args.jobs, C, args.tiles, args.reps)
                r["frac_of_peak"] = r["tflops"] / (max(peak, peak2) / 1e12)
                print(json.dumps(r))
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"variant": v, "children": C, "error": str(e)}))


if __name__ == "__main__":
    main()
