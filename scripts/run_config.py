"""Run the GPU sampler on one synthetic BASELINE config and print timing + engine statistics.
usage: python scripts/run_config.py [--clusters 50 --samples 8 --grid 101 --particles 100 --iters 3 --depth 200]"""

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=50)
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--grid", type=int, default=101)
    ap.add_argument("--particles", type=int, default=100)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--burnin", type=int, default=1)
    ap.add_argument("--muts", type=int, default=10)
    ap.add_argument("--depth", type=int, default=200)
    ap.add_argument("--outlier-prob", type=float, default=0.0)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--sweep", default="host", choices=["host", "device"])
    args = ap.parse_args()
    import torch

    from phyclone_b200 import ops
    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import run_phyclone_chain
    from phyclone_b200.simulate import simulate_counts

    c = simulate_counts(args.clusters, args.samples, args.muts, args.depth, seed=0)
    t0 = time.time()
    grids = ops.likelihood_grid(c["ref"], c["alt"], c["tumour_content"], c["cn"], c["mu"], c["log_pi"], c["n_geno"],
                                args.grid, "beta-binomial", 400)
    values = ops.cluster_sum(grids, c["cluster_off"])
    if args.outlier_prob > 0:
        probs = [(np.log(args.outlier_prob) * args.muts, np.log1p(-args.outlier_prob) * args.muts)] * args.clusters
    else:
        probs = None
    data = make_data_points(values, outlier_probs=probs)
    t_load = time.time() - t0
    rng = np.random.default_rng(1)
    marks = []

    def on_it(i, e):
        torch.cuda.synchronize()
        marks.append(time.time())

    def go():
        return run_phyclone_chain(args.burnin, True, 1.0, data, float("inf"), args.iters, args.particles, 1, 1,
                                  args.outlier_prob, 10**9, "semi-adapted", 0.5, rng, ["s"], 1, 0, 0.0, on_iteration=on_it,
                                  sweep=args.sweep)

    t0 = time.time()
    if args.profile:
        import cProfile
        import pstats

        pr = cProfile.Profile()
        res = pr.runcall(go)
        pstats.Stats(pr).sort_stats(os.environ.get("PROF_SORT", "cumulative")).print_stats(60)
    else:
        res = go()
    wall = time.time() - t0
    per_it = np.diff(marks) if len(marks) > 1 else np.array([wall])
    out = {"load_s": t_load, "wall_s": wall, "iters": args.iters, "s_per_iter_after_first": float(per_it.mean()),
           "it_per_s": float(1.0 / per_it.mean()), "stats": res["stats"],
           "final_log_p_one": res["trace"][-1]["log_p_one"],
           "n_nodes": len([k for k in res["trace"][-1]["tree"]["node_data"] if k not in ("root", -1)])}
    enq, syn, nfl = ops.plan_stats()
    out["flush_split"] = {"enqueue_s": enq, "sync_s": syn, "flushes": nfl}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
