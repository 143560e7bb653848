#!/usr/bin/env python
"""Benchmark of the PhyClone hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels + host sampler)
    python bench.py --impl reference --steps K --warmup W     # CPU arm: the oracle port of the reference
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one chain per GPU

A "step" is one particle-Gibbs iteration = one trip of the reference's main loop (run.py:261-288):
conditional-SMC sweep over all data points with N particles, data-point Gibbs, prune-regraph, relabel and
concentration update, on the synthetic BASELINE.json configs[1] shape (50 clusters x 8 samples, grid 101,
100 particles, one chain per GPU).  `value` = PG iterations/s summed over all chains (chains are
independent, so N GPUs run N chains: weak scaling, no data-path collective).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = {"clusters": 50, "samples": 8, "grid": 101, "particles": 100, "muts_per_cluster": 10, "depth": 200,
       "proposal": "semi-adapted", "resample_threshold": 0.5, "precision": 400, "density": "beta-binomial"}
WORKLOAD = ("synthetic 50 clusters x 8 samples, grid 101, 100 particles, semi-adapted proposal, 1 chain per GPU "
            "(BASELINE.json configs[1]); data seed 0, chain seed 1")
METRIC = "pg_iterations_per_sec"
UNIT = "PG iterations/s"


# ----------------------------------------------------------------------------- inputs
def synthetic_counts():
    from phyclone_b200.simulate import simulate_counts

    return simulate_counts(CFG["clusters"], CFG["samples"], CFG["muts_per_cluster"], CFG["depth"], seed=0)


def chain_rng(rank, world):
    """run.py:87-139: a single chain uses the main Generator itself, several chains use its spawned children."""
    main = np.random.default_rng(1)
    return main if world == 1 else main.spawn(world)[rank]


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region.  NVML in-process (a few microseconds per
    sample); spawning `nvidia-smi` five times a second measurably slowed the launch-bound step it was watching, so
    that is only the fallback."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []  # [sm_mhz, sm_max_mhz, power_w, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap]
        self._stop = threading.Event()
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._nvml = self._handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = None
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(index).uuid)
            except Exception:  # noqa: BLE001
                pass
            handle = None
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(h)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        handle = h
                        break
            self._handle = handle if handle is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _sample_nvml(self):
        n, h = self._nvml, self._handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        sm_max = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        power = n.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:  # noqa: BLE001
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flag = lambda bit: "Active" if r & bit else "Not Active"  # noqa: E731
        self.rows.append([str(sm), str(sm_max), str(power), flag(0x8), flag(0x40), flag(0x20), flag(0x4)])

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 7:
            self.rows.append(parts)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1 if self._nvml is not None else 1.0)

    def __enter__(self):
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------- CPU arms
def oracle_chain_rate(iters, warmup, chain=0, n_chains=1):
    """PG iterations/s of the CPU port of the reference (oracle/sampler.py, with the reference's tile
    caches on: its real configuration) on the same workload; 1 core (a chain is single-threaded)."""
    from oracle import sampler as osm
    from oracle.synthetic import make_synthetic

    syn = make_synthetic(CFG["clusters"], CFG["samples"], CFG["grid"], CFG["muts_per_cluster"], CFG["depth"], seed=0,
                         density=CFG["density"], precision=CFG["precision"])
    data = [osm.Datum(i, v) for i, v in enumerate(syn["values"])]
    marks = []
    res = osm.run_chain(data, chain_rng(chain, n_chains), warmup + iters, burnin=1, num_particles=CFG["particles"],
                        proposal=CFG["proposal"], resample_threshold=CFG["resample_threshold"], use_tile_caches=True,
                        on_iteration=lambda i, e: marks.append(time.perf_counter()))
    t0 = marks[warmup - 1] if warmup > 0 else None
    if t0 is None:
        raise ValueError("need at least one warm-up iteration for the CPU arm")
    elapsed = marks[-1] - t0
    return iters / elapsed, elapsed, res["extensions"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: at ~10 s per iteration, cap the timed iterations so the run ends within a few minutes
    iters = min(args.steps, 12)
    warm = min(max(args.warmup, 1), 2)
    n_chains = max(args.gpus, 1)  # the GPU arm runs one chain per GPU: same job here
    cores = min(n_chains, os.cpu_count() or 1)
    if n_chains == 1:
        rate, elapsed, ext = oracle_chain_rate(iters, warm)
    else:
        # the reference runs its chains in a ProcessPoolExecutor, one single-threaded process per chain (run.py:111-149)
        from concurrent.futures import ProcessPoolExecutor

        with ProcessPoolExecutor(max_workers=cores) as pool:
            t0 = time.perf_counter()
            futs = [pool.submit(oracle_chain_rate, iters, warm, c, n_chains) for c in range(n_chains)]
            done = [f.result() for f in futs]
            wall = time.perf_counter() - t0
        # whole-job rate over the timed iterations: chains ran side by side (queued if there are fewer cores)
        waves = -(-n_chains // cores)
        elapsed = max(d[1] for d in done) * waves if waves > 1 else max(d[1] for d in done)
        rate = n_chains * iters / elapsed
        elapsed = wall
    sample = ("%d timed PG iterations per chain, %d chain(s) in %d process(es), after 1 burn-in + %d warm-up iterations "
              "(%.1f s wall); a chain is single-threaded in the reference" % (iters, n_chains, cores, warm, elapsed))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(CFG, workload=WORKLOAD, chains=n_chains),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU port of the reference sampler (oracle/sampler.py, tile caches on); /root/reference is a Python "
                "package with an uninstallable Rust dependency and does not exist on the GPU box",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from phyclone_b200 import ops
    from phyclone_b200.data.base import make_data_points
    from phyclone_b200.run import Chain
    from phyclone_b200.tiles import TileStore

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    # one host thread drives each chain: give every rank its own block of cores so that the ranks of a box do not
    # migrate over each other (PCB_NO_PIN=1 turns it off)
    pinned = None
    if world > 1 and not os.environ.get("PCB_NO_PIN") and hasattr(os, "sched_setaffinity"):
        cores = sorted(os.sched_getaffinity(0))
        block = len(cores) // world
        if block >= 1:
            mine = cores[local * block:(local + 1) * block]
            os.sched_setaffinity(0, mine)
            pinned = "%d cores per rank (contiguous blocks of the %d available)" % (block, len(cores))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs (one-off, a2/a3/a4 of the path on the device; not timed)
    c = synthetic_counts()
    grids = ops.likelihood_grid(c["ref"], c["alt"], c["tumour_content"], c["cn"], c["mu"], c["log_pi"], c["n_geno"],
                                CFG["grid"], CFG["density"], CFG["precision"])
    values = ops.cluster_sum(grids, c["cluster_off"])
    data = make_data_points(values)
    fp64_peak, _ = ops.fp64_peak(8192)
    fp64_peak = max(fp64_peak, ops.fp64_peak(8192)[0])

    store = TileStore(values, device=dev, variant=TileStore.LATENCY_LAYOUT)  # what run_phyclone_chain uses
    chain = Chain(data, chain_rng(rank, world), num_particles=CFG["particles"], proposal=CFG["proposal"],
                  resample_threshold=CFG["resample_threshold"], store=store)
    chain.burnin_step()
    for _ in range(args.warmup):
        chain.step()

    stream = torch.cuda.current_stream(dev)

    per_rank_ms = []

    def timed(n_steps, before_step=None, after_step=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(n_steps):
            if before_step:
                before_step()
            chain.step()
            if after_step:
                after_step()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            per_rank_ms[:] = [float(t.item()) / n_steps for t in every]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        else:
            per_rank_ms[:] = [float(ms.item()) / n_steps]
        return float(ms.item())

    # ---- timed region 1: inputs resident in HBM.  Iterations of an MCMC chain differ in cost (the forest changes), so
    # the chain is rewound after this leg: both legs run the SAME iterations and differ only in the host transfers
    snap = chain.snapshot()
    ops.timing_begin(2000 * args.steps + 4096)
    s0 = dict(chain.stats)
    launches0 = ops.launch_count()
    staged0 = store.arena.staged_bytes
    with ClockSampler(local) as clocks:
        ms = timed(args.steps)
    value_per_rank_ms = list(per_rank_ms)
    launches = ops.launch_count() - launches0
    s1 = dict(chain.stats)
    kernel_ms, n_timed = ops.timing_end()
    conv_pairs = s1["conv_pairs"] - s0["conv_pairs"]
    flop = conv_pairs * CFG["samples"] * CFG["grid"] * (CFG["grid"] + 1)  # 2 flop x S*G*(G+1)/2 FMA per pair conv
    n_launch = max(n_timed, 1)
    achieved = flop / max(kernel_ms * 1e-3, 1e-12)
    value = world * args.steps / (ms * 1e-3)

    # ---- timed region 2: end to end through the host-facing API, inputs re-sent from pinned host memory
    chain.restore(snap)
    staged = lambda: store.arena.staged_bytes + 4 * chain.stats["staged_ints"]  # noqa: E731 -- id lists / job tables sent
    d2h0, st0 = chain.stats["d2h_bytes"], staged()
    sink = []
    ms_e2e = timed(args.steps, before_step=lambda: store.upload_values(values),
                   after_step=lambda: sink.append(chain.log_p_one()))
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    h2d = values.nbytes + (staged() - st0) / args.steps
    d2h = (chain.stats["d2h_bytes"] - d2h0) / args.steps

    # ---- the same iterations once more with nothing attached (no kernel events, no transfers): what the event pairs
    # of region 1 cost
    chain.restore(snap)
    ms_plain = timed(args.steps)
    value_plain = world * args.steps / (ms_plain * 1e-3)

    # ---- saturated-batch figure for the same kernel (how far the kernel itself goes when a launch is big)
    sat = None
    if rank == 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            from kernel_bench import bench_variant

            lay = store.arena.layout
            live_variant = (10000 if lay.kernel else 0) + 100 * lay.strip + (0 if lay.kernel else lay.lanes_per_row)
            r = bench_variant(CFG["samples"], CFG["grid"], -1, 32768, 8, 4096, 5)
            r2 = bench_variant(CFG["samples"], CFG["grid"], live_variant, 32768, 8, 4096, 5)
            sat = {"achieved": r["tflops"], "unit": "TFLOP/s", "frac": r["tflops"] * 1e12 / fp64_peak,
                   "launch": "32768 jobs x 8 children",
                   "layout": "throughput layout (L=%d, P=%d): what a saturated batch would run with" % (r["L"], r["P"]),
                   "chain_layout": {"layout": "latency layout (L=%d, P=%d): what the timed chain runs with, its launches "
                                              "being ~20 jobs" % (r2["L"], r2["P"]),
                                    "achieved": r2["tflops"], "frac": r2["tflops"] * 1e12 / fp64_peak}}
        except Exception as e:  # noqa: BLE001
            sat = {"error": str(e)}

    # ---- BASELINE configs[2] on this one GPU: 8 chains in 8 host processes sharing the device through MPS (one chain
    # leaves it ~70 % idle).  Reported beside the headline, never mixed into it; skipped under torchrun.
    mps_chains = None
    if rank == 0 and world == 1 and not args.no_mps_chains:
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            from multi_chain_bench import run_chains_once

            from phyclone_b200.run import MpsServer

            n_chains = max(2, min(8, (os.cpu_count() or 2) // 2))
            with MpsServer() as server:
                if server.started:
                    rate = run_chains_once(n_chains, args.steps, args.warmup, timeout=300)
                    mps_chains = {"chains": n_chains, "value": rate, "unit": UNIT, "per_chain": rate / n_chains,
                                  "how": "one host process per chain, all on GPU 0 through an MPS daemon started for this "
                                         "leg (run.MpsServer / run.run_chains_local); same config, chain seeds spawned"}
                else:
                    mps_chains = {"unavailable": "nvidia-cuda-mps-control did not start"}
        except Exception as e:  # noqa: BLE001
            mps_chains = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, elapsed, _ = oracle_chain_rate(2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "2 PG iterations of the same config after 1 burn-in + 1 warm-up iteration (%.1f s), "
                         "oracle/sampler.py with the reference's tile caches on" % elapsed}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        ext = (s1["extensions"] - s0["extensions"]) * world
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture of this command
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_node_jobs_traffic.json")))
            traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/r1_node_jobs_traffic.json (ncu --set full)"
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(CFG, workload=WORKLOAD, chains=world,
                           l2="no explicit flush: one iteration allocates %d fresh tiles (~%d MB) in the arena, larger than "
                              "the 126 MB L2" % ((s1["adds"] + s1["node_jobs"] - s0["adds"] - s0["node_jobs"]) // args.steps,
                                                 (s1["adds"] + s1["node_jobs"] - s0["adds"] - s0["node_jobs"]) // args.steps
                                                 * store.arena.layout.tile_elems * 16 // (1 << 20))),
            "particle_extensions_per_sec": ext / (ms * 1e-3),
            "value_without_kernel_timing": value_plain,  # same iterations, no CUDA events around the kernel launches
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "per_rank_ms_per_step": [round(x, 2) for x in value_per_rank_ms],
            "cpu_pinning": pinned,
            "roofline": {"bound": "fp64", "kernel": "node_jobs_kernel", "achieved": achieved / 1e12,
                         "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "launches": int(n_timed), "avg_launch_us": 1e3 * kernel_ms / n_launch,
                         "flop_per_launch": flop / n_launch, "kernel_share_of_step": kernel_ms / ms,
                         "peak_source": "pcb_fp64_peak (dependent-DFMA-chain microbenchmark) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 figure (hbm_gbs=%s)" % peaks.get("hbm_gbs"),
                         "saturated_batch": sat},
            "cpu_baseline": cpu,
            "chains_on_one_gpu_mps": mps_chains,
            "clocks": clocks.summary(),
            "per_step": {k: (s1[k] - s0[k]) / args.steps for k in ("flushes", "launches", "node_jobs", "adds", "scores",
                                                                    "conv_pairs", "memo_hits", "extensions")},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mps-chains", action="store_true", help="skip the 8-chains-on-one-GPU (MPS) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
