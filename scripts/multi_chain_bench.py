"""Several independent chains on ONE GPU, one host process each (BASELINE.json configs[2] at N=1: "8 chains").
Chains never talk to each other; the GPU time-slices their small launches.  Prints aggregate PG iterations/s.
usage: python scripts/multi_chain_bench.py --chains 8 --steps 5 --warmup 3"""

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, n, steps, warmup, barrier, q, config=2, sweep="host"):
    import numpy as np
    import torch

    import bench
    from phyclone_b200 import ops
    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import Chain

    from phyclone_b200 import workloads

    cfg = bench.set_config(config)
    values = bench.workload_values_gpu(ops)
    data = make_data_points(values, outlier_probs=workloads.outlier_log_probs(cfg))
    rng = np.random.default_rng(1).spawn(max(n, 2))[rank] if n > 1 else np.random.default_rng(1)
    chain = Chain(data, rng, num_particles=cfg["particles"], outlier_prob=cfg["outlier_prob"], sweep=sweep)
    chain.burnin_step()
    for _ in range(warmup):
        chain.step()
    torch.cuda.synchronize()
    barrier.wait(timeout=300)
    t0 = time.perf_counter()
    for _ in range(steps):
        chain.step()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    q.put((rank, t0, t1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", default="1,2,4,8")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mps", action="store_true", help="share the GPU through an MPS daemon started for the run")
    args = ap.parse_args()
    from phyclone_b200.run import MpsServer

    with MpsServer(enable=args.mps) as server:
        print(json.dumps({"mps": server.started}))
        sweep(args)


def run_chains_once(n, steps, warmup, timeout=600, config=2, sweep="host"):
    """n chains, one process each, same GPU; returns the aggregate PG iterations/s over the common timed window."""
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(n)
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, n, steps, warmup, barrier, q, config, sweep), daemon=True) for r in range(n)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=timeout) for _ in range(n)]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
    t0 = min(r[1] for r in res)
    t1 = max(r[2] for r in res)
    return n * steps / (t1 - t0)


def sweep(args):
    for n in [int(x) for x in args.chains.split(",")]:
        rate = run_chains_once(n, args.steps, args.warmup)
        print(json.dumps({"chains_on_one_gpu": n, "steps": args.steps, "aggregate_pg_it_per_s": rate,
                          "per_chain_it_per_s": rate / n, "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
