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


def worker(rank, n, steps, warmup, barrier, q, config=2, sweep="host", device=0, chain=None, n_chains=None):
    """One chain in its own process: chain `chain` of `n_chains` (default: worker `rank` of `n`) on GPU `device`.
    Puts (rank, wall t0, wall t1, CUDA-event ms of the timed steps, stats) on the queue."""
    import numpy as np
    import torch

    import bench
    from phyclone_b200 import ops, workloads
    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import Chain

    chain = rank if chain is None else chain
    n_chains = n if n_chains is None else n_chains
    torch.cuda.set_device(device)
    cfg = bench.set_config(config)
    values = bench.workload_values_gpu(ops)
    data = make_data_points(values, outlier_probs=workloads.outlier_log_probs(cfg))
    rng = np.random.default_rng(1).spawn(max(8, n_chains))[chain]  # bench.chain_rng: the same family of chains
    ch = Chain(data, rng, num_particles=cfg["particles"], outlier_prob=cfg["outlier_prob"], sweep=sweep)
    ch.burnin_step()
    for _ in range(warmup):
        ch.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan0, s0 = ops.plan_stats(), ch.stats
    barrier.wait(timeout=600)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        ch.step()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    plan1, s1 = ops.plan_stats(), ch.stats
    sweep_ms = (s1.get("sweep", {}).get("device_ms", 0.0) - s0.get("sweep", {}).get("device_ms", 0.0)) / steps
    q.put((rank, t0, t1, float(e0.elapsed_time(e1)),
           {"device_sweep_ms": sweep_ms, "flush_wait_ms": 1e3 * (plan1[1] - plan0[1]) / steps}))


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


def run_chains_once(n, steps, warmup, timeout=600, config=2, sweep="host", device=0, chains=None, n_chains=None,
                    before_go=None, details=False):
    """n chains, one process each, same GPU; returns the aggregate PG iterations/s over the common timed window.
    `chains` / `n_chains`: which chains of a larger job these workers are (default 0..n-1 of n).  `before_go` is called
    once every worker is warmed up and waiting (e.g. a cross-rank barrier).  details=True returns
    (rate, max CUDA-event ms over the workers, per-worker split dicts)."""
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(n + (1 if before_go is not None else 0))
    q = ctx.Queue()
    chains = list(range(n)) if chains is None else list(chains)
    procs = [ctx.Process(target=worker, args=(r, n, steps, warmup, barrier, q, config, sweep, device, chains[r], n_chains),
                         daemon=True) for r in range(n)]
    for p in procs:
        p.start()
    try:
        if before_go is not None:
            # the parent joins the barrier last: workers are all parked on it by then
            while barrier.n_waiting < n:
                if not all(p.is_alive() for p in procs):
                    raise RuntimeError("a chain worker died during warm-up")
                time.sleep(0.01)
            before_go()
            barrier.wait(timeout=60)
        res = [q.get(timeout=timeout) for _ in range(n)]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
    t0 = min(r[1] for r in res)
    t1 = max(r[2] for r in res)
    rate = n * steps / (t1 - t0)
    if details:
        return rate, max(r[3] for r in res), [r[4] for r in sorted(res)]
    return rate


def sweep(args):
    for n in [int(x) for x in args.chains.split(",")]:
        rate = run_chains_once(n, args.steps, args.warmup)
        print(json.dumps({"chains_on_one_gpu": n, "steps": args.steps, "aggregate_pg_it_per_s": rate,
                          "per_chain_it_per_s": rate / n, "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
