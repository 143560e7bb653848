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

METRIC = "pg_iterations_per_sec"
UNIT = "PG iterations/s"
CFG = None  # set by main() from --config (phyclone_b200/workloads.py)


def set_config(number):
    global CFG
    from phyclone_b200 import workloads

    CFG = workloads.config_dict(number)
    return CFG


# ----------------------------------------------------------------------------- inputs
def workload_values_cpu():
    """[T,S,G] grids of the configured workload computed on the CPU (NumPy oracle): CPU arms only."""
    from oracle import numerics as onum
    from phyclone_b200 import workloads

    c = workloads.counts(CFG)
    if c is None:
        return workloads.fixture_values(CFG)
    grids = onum.likelihood_grid(c["ref"], c["alt"], c["tumour_content"], c["cn"], c["mu"], c["log_pi"], c["n_geno"],
                                 CFG["grid"], CFG["density"], CFG["precision"])
    return onum.cluster_sum(grids, c["cluster_off"])


def workload_values_gpu(ops):
    """The same grids through the device path (a2/a3 of the hot path); one-off, not timed."""
    from phyclone_b200 import workloads

    c = workloads.counts(CFG)
    if c is None:
        return workloads.fixture_values(CFG)
    grids = ops.likelihood_grid(c["ref"], c["alt"], c["tumour_content"], c["cn"], c["mu"], c["log_pi"], c["n_geno"],
                                CFG["grid"], CFG["density"], CFG["precision"])
    return ops.cluster_sum(grids, c["cluster_off"])


def chain_rng(rank, world):
    """Chain `rank` of the job: child `rank` of `default_rng(1).spawn(max(8, world))` (run.py:87-113 spawns one child
    per chain).  The same children at every N -- also for one chain, where the reference would use the main Generator
    itself -- so that rank r runs the SAME chain whatever the number of GPUs: per-N values are then comparable and
    `--gpus N`, `--chains 8` and the reference arm all draw from one family of eight chains."""
    return np.random.default_rng(1).spawn(max(8, world))[rank]


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
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
REF_SHIM = os.path.join(ROOT, "oracle", "refshim")


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "phyclone"))


def reference_chain_times(values, outlier_probs, cfg, iters, warmup, chain, n_chains, max_seconds):
    """The UNMODIFIED reference (pip-installed into baseline/_ref by scripts/install_reference.sh; its absent Rust
    dependency `rustworkx` is served by oracle/refshim) run through its own public entry point
    `phyclone.run.run_phyclone_chain` (run.py:177-232): 1 burn-in sweep, then warmup + iters PG iterations, caches
    on (its real configuration).  Returns the cumulative `time` field of its trace entries (entry j = after
    iteration j-1; its Timer covers burn-in and main loop, run.py:262) -- fewer entries than asked if the wall cap
    `max_seconds` (the reference's own --max-time) stopped it."""
    import io
    from contextlib import redirect_stdout

    for p in (REF_SHIM, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    from phyclone.data.base import DataPoint as RefDataPoint
    from phyclone.run import run_phyclone_chain as ref_run

    data = []
    for i, v in enumerate(values):
        op, opn = (0, 1) if outlier_probs is None else outlier_probs[i]
        data.append(RefDataPoint(i, v, outlier_prob=op, outlier_prob_not=opn))
    with redirect_stdout(io.StringIO()):
        res = ref_run(1, True, 1.0, data, float(max_seconds), warmup + iters, cfg["particles"], 1, 1, cfg["outlier_prob"],
                      10**9, cfg["proposal"], cfg["resample_threshold"], chain_rng(chain, n_chains), ["s"], 1, chain, 0.0)
    return [e["time"] for e in res["trace"]]


def port_chain_times(values, outlier_probs, cfg, iters, warmup, chain, n_chains, max_seconds):
    """Same contract as reference_chain_times for the CPU port (oracle/sampler.py, the reference's tile caches on)."""
    from oracle import sampler as osm

    data = []
    for i, v in enumerate(values):
        op, opn = (0, 1) if outlier_probs is None else outlier_probs[i]
        data.append(osm.Datum(i, v, outlier_prob=op, outlier_prob_not=opn))
    t_start = time.perf_counter()
    marks = []

    class Stop(Exception):
        pass

    def on_it(i, e):
        marks.append(time.perf_counter() - t_start)
        if marks[-1] >= max_seconds:
            raise Stop()

    try:
        osm.run_chain(data, chain_rng(chain, n_chains), warmup + iters, burnin=1, num_particles=cfg["particles"],
                      proposal=cfg["proposal"], resample_threshold=cfg["resample_threshold"],
                      outlier_prob=cfg["outlier_prob"], use_tile_caches=True, on_iteration=on_it)
    except Stop:
        pass
    # entry 0 = the end of burn-in is not observable through on_iteration: the first warm-up iteration absorbs it
    return [0.0] + marks


def _cpu_chain(kind, cfg, iters, warmup, chain, n_chains, max_seconds):
    global CFG
    CFG = cfg
    from phyclone_b200 import workloads

    values = workload_values_cpu()
    fn = reference_chain_times if kind == "reference" else port_chain_times
    return fn(values, workloads.outlier_log_probs(cfg), cfg, iters, warmup, chain, n_chains, max_seconds)


def cpu_rate(kind, iters, warmup, n_chains=1, max_seconds=420.0):
    """(PG iterations/s whole job, timed iterations per chain actually run, warm-up actually run, wall seconds,
    cores) of the CPU arm `kind` ("reference" | "port") on the configured workload.  Chains run side by side in
    one single-threaded process each, as the reference's ProcessPoolExecutor does (run.py:111-149)."""
    cores = min(n_chains, os.cpu_count() or 1)
    warmup = max(warmup, 1)  # the first iteration carries the Numba JIT / first-touch costs
    t0 = time.perf_counter()
    if n_chains == 1:
        traces = [_cpu_chain(kind, CFG, iters, warmup, 0, 1, max_seconds)]
    else:
        from concurrent.futures import ProcessPoolExecutor

        with ProcessPoolExecutor(max_workers=cores) as pool:
            futs = [pool.submit(_cpu_chain, kind, CFG, iters, warmup, c, n_chains, max_seconds) for c in range(n_chains)]
            traces = [f.result() for f in futs]
    wall = time.perf_counter() - t0
    done = min(len(t) - 1 for t in traces)  # iterations every chain completed
    warm_done = min(warmup, max(done - 1, 0))
    timed = done - warm_done
    if timed < 1:
        return None, 0, warm_done, wall, cores
    # the chains ran side by side (in waves if there are fewer cores than chains): the job's time for the timed
    # iterations is the slowest chain's, times the number of waves
    waves = -(-n_chains // cores)
    elapsed = max(t[warm_done + timed] - t[warm_done] for t in traces) * waves
    if not elapsed > 0:  # the wall cap cut the run inside an iteration: the trace's last stamps coincide
        return None, 0, warm_done, wall, cores
    return n_chains * timed / elapsed, timed, warm_done, wall, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = "reference" if reference_available() and not args.port else "port"
    n_chains = max(args.gpus, 1) if args.chains is None else args.chains  # same job as the GPU arm
    rate, timed, warm, wall, cores = cpu_rate(kind, args.steps, args.warmup, n_chains, args.ref_max_seconds)
    if rate is None:
        print(json.dumps({"impl": "reference", "unavailable": "no PG iteration of %s finished within the %.0f s cap "
                                                               "(%s)" % (CFG["workload"], args.ref_max_seconds, kind)}))
        return
    how = ("unmodified reference (phyclone 0.5.1 pip-installed into baseline/_ref, rustworkx served by oracle/refshim), "
           "phyclone.run.run_phyclone_chain, caches on" if kind == "reference"
           else "CPU port of the reference sampler (oracle/sampler.py, tile caches on)")
    sample = ("%d timed PG iterations per chain (asked: %d) after 1 burn-in sweep + %d warm-up iterations (asked: %d), "
              "%d chain(s) in %d single-threaded process(es), wall %.1f s, wall cap %.0f s; %s"
              % (timed, args.steps, warm, args.warmup, n_chains, cores, wall, args.ref_max_seconds, how))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": timed,
        "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup, "ms_per_step": 1e3 * n_chains / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(CFG, chains=n_chains),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
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
    from phyclone_b200 import workloads

    values = workload_values_gpu(ops)
    data = make_data_points(values, outlier_probs=workloads.outlier_log_probs(CFG))
    fp64_peak, _ = ops.fp64_peak(8192)
    fp64_peak = max(fp64_peak, ops.fp64_peak(8192)[0])

    store = TileStore(values, device=dev, variant=TileStore.LATENCY_LAYOUT)  # what run_phyclone_chain uses
    chain = Chain(data, chain_rng(rank, world), num_particles=CFG["particles"], proposal=CFG["proposal"],
                  resample_threshold=CFG["resample_threshold"], outlier_prob=CFG["outlier_prob"], store=store,
                  sweep=args.sweep)
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
    # every 3rd launch of the job kernel is bracketed by an event pair (odd: a sweep step launches the kernel twice, the
    # sample must alternate between the two); see roofline.timing
    TIMING_STRIDE = 3
    ops.timing_begin(2000 * args.steps + 4096, stride=TIMING_STRIDE)
    s0 = dict(chain.stats)
    plan0 = ops.plan_stats()
    launches0 = ops.launch_count()
    staged0 = store.arena.staged_bytes
    with ClockSampler(local) as clocks:
        ms = timed(args.steps)
    value_per_rank_ms = list(per_rank_ms)
    launches = ops.launch_count() - launches0
    s1 = dict(chain.stats)
    plan1 = ops.plan_stats()
    # where a step's time goes on this rank: device-resident sweep (CUDA-event time), waiting for flushes of the tile
    # store (host blocked on the stream), everything else = host sampler logic
    sweep_ms = (s1.get("sweep", {}).get("device_ms", 0.0) - s0.get("sweep", {}).get("device_ms", 0.0)) / args.steps
    wait_ms = 1e3 * (plan1[1] - plan0[1]) / args.steps
    split = torch.tensor([sweep_ms, wait_ms], dtype=torch.float64, device=dev)
    if world > 1:
        splits = [torch.zeros_like(split) for _ in range(world)]
        dist.all_gather(splits, split)
    else:
        splits = [split]
    splits = [[float(x) for x in t.tolist()] for t in splits]
    n_seen = ops.timing_seen()
    kernel_ms, n_timed = ops.timing_end()
    ops.timing_begin(1, stride=1)  # back to the default stride for whoever times next
    ops.timing_end()
    def all_pairs(st):  # pair convolutions folded by flushes of the tile store + by device-resident sweeps
        return st["conv_pairs"] + sum(st.get(k, {}).get("conv_pairs", 0) for k in ("sweep", "burnin_sweep"))

    conv_pairs = all_pairs(s1) - all_pairs(s0)
    flop = conv_pairs * CFG["samples"] * CFG["grid"] * (CFG["grid"] + 1)  # 2 flop x S*G*(G+1)/2 FMA per pair conv
    n_launch = max(n_timed, 1)
    avg_launch_ms = kernel_ms / n_launch
    kernel_ms = avg_launch_ms * max(n_seen, n_timed)  # all launches of the region at the sampled average duration
    achieved = flop / max(kernel_ms * 1e-3, 1e-12)
    value = world * args.steps / (ms * 1e-3)

    # ---- timed region 2: end to end through the host-facing API, inputs re-sent from pinned host memory
    chain.restore(snap)
    def sweep_bytes(which):  # control tables up / genealogy down of the device-resident sweeps
        st = chain.stats
        return sum(st.get(k, {}).get(which, 0) for k in ("sweep", "burnin_sweep"))

    staged = lambda: (store.arena.staged_bytes + 4 * chain.stats["staged_ints"]  # noqa: E731 -- id lists / job tables sent
                      + sweep_bytes("h2d_bytes"))
    d2h_all = lambda: chain.stats["d2h_bytes"] + sweep_bytes("d2h_bytes")  # noqa: E731
    d2h0, st0 = d2h_all(), staged()
    sink = []
    ms_e2e = timed(args.steps, before_step=lambda: store.upload_values(values),
                   after_step=lambda: sink.append(chain.log_p_one()))
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    h2d = values.nbytes + (staged() - st0) / args.steps
    d2h = (d2h_all() - d2h0) / args.steps

    # ---- the same iterations once more with nothing attached (no kernel events, no transfers): what the event pairs
    # of region 1 cost
    chain.restore(snap)
    ms_plain = timed(args.steps)
    value_plain = world * args.steps / (ms_plain * 1e-3)

    # ---- the same workload with the sweeps on the host (NumPy Generator draws: RNG-identical to the reference)
    other = None
    if rank == 0 and world == 1 and not args.no_other_sweep:
        which = "host" if args.sweep == "device" else "device"
        try:
            store2 = TileStore(values, device=dev, variant=TileStore.LATENCY_LAYOUT)
            chain2 = Chain(data, chain_rng(rank, world), num_particles=CFG["particles"], proposal=CFG["proposal"],
                           resample_threshold=CFG["resample_threshold"], outlier_prob=CFG["outlier_prob"], store=store2,
                           sweep=which)
            chain2.burnin_step()
            for _ in range(args.warmup):
                chain2.step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                chain2.step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            other = {"sweep": which, "value": args.steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / args.steps,
                     "how": "second chain object, same data / seed / iteration window, wall clock around the timed steps"}
            del chain2, store2
        except Exception as e:  # noqa: BLE001
            other = {"sweep": which, "error": "%s: %s" % (type(e).__name__, e)}

    # ---- saturated-batch figure for the same kernel (how far the kernel itself goes when a launch is big)
    sat = None
    if rank == 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            from kernel_bench import bench_variant

            lay = store.arena.layout
            live_variant = (10000 if (lay.kernel & 3) else 0) + 100 * lay.strip + (0 if (lay.kernel & 3) else lay.lanes_per_row)
            r = bench_variant(CFG["samples"], CFG["grid"], -1, 32768, 8, 4096, 5)
            r2 = bench_variant(CFG["samples"], CFG["grid"], live_variant, 32768, 8, 4096, 5)
            r3 = bench_variant(CFG["samples"], CFG["grid"], -1, 16384, 8, 4096, 5, score_only=False)
            sat = {"achieved": r["tflops"], "unit": "TFLOP/s", "frac": r["tflops"] * 1e12 / fp64_peak,
                   "launch": "32768 score-only jobs x 8 children (dummy-root folds: no tile written, no log_p read)",
                   "full_node_jobs": {"launch": "16384 node refreshes x 8 children, each reading a log_p tile and "
                                                "writing its log_r tile (log + linear + row maxima)",
                                      "achieved": r3["tflops"], "frac": r3["tflops"] * 1e12 / fp64_peak},
                   "layout": "throughput layout (L=%d, P=%d): what a saturated batch would run with" % (r["L"], r["P"]),
                   "chain_layout": {"layout": "the chain's tile layout (L=%d, P=%d) in the lane-per-unit mapping on a "
                                              "saturated batch; the chain's own launches (<= 1024 jobs) run the "
                                              "warp-per-row mapping, which trades throughput for latency" % (r2["L"], r2["P"]),
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
                    rate = run_chains_once(n_chains, args.steps, args.warmup, timeout=300, config=CFG["config_number"],
                                           sweep=args.sweep)
                    mps_chains = {"chains": n_chains, "value": rate, "unit": UNIT, "per_chain": rate / n_chains,
                                  "how": "one host process per chain, all on GPU 0 through an MPS daemon started for this "
                                         "leg (run.MpsServer / run.run_chains_local); same config, chain seeds spawned"}
                else:
                    mps_chains = {"unavailable": "nvidia-cuda-mps-control did not start"}
        except Exception as e:  # noqa: BLE001
            mps_chains = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = "reference" if reference_available() else "port"
        rate, timed_it, warm_it, wall, _ = cpu_rate(kind, 2, 1, 1, 120.0)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "%d PG iterations of the same config after 1 burn-in sweep + %d warm-up iteration (%.1f s wall, "
                         "cap 120 s); %s" % (timed_it, warm_it, wall,
                                             "unmodified reference from baseline/_ref, caches on" if kind == "reference"
                                             else "oracle/sampler.py with the reference's tile caches on")}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        ext = (s1["extensions"] - s0["extensions"]) * world
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture of this command
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_node_jobs_traffic.json")))
            traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/r2_node_jobs_traffic.json (ncu --set full)"
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(CFG, chains=world),
            "l2": "no explicit flush: one iteration allocates %d fresh tiles (~%d MB) in the arena, larger than the 126 MB "
                  "L2" % ((s1["adds"] + s1["node_jobs"] - s0["adds"] - s0["node_jobs"]) // args.steps,
                          (s1["adds"] + s1["node_jobs"] - s0["adds"] - s0["node_jobs"]) // args.steps
                          * store.arena.layout.tile_elems * 16 // (1 << 20)),
            "particle_extensions_per_sec": ext / (ms * 1e-3),
            "value_without_kernel_timing": value_plain,  # same iterations, no CUDA events around the kernel launches
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "per_rank_ms_per_step": [round(x, 2) for x in value_per_rank_ms],
            "per_rank_split_ms_per_step": [
                {"device_sweep": round(sw, 2), "flush_wait": round(fw, 2), "host": round(max(t - sw - fw, 0.0), 2)}
                for t, (sw, fw) in zip(value_per_rank_ms, splits)],
            "limiter": "host sampler logic (one Python thread per chain) + serial device latency: see "
                       "per_rank_split_ms_per_step; chains never exchange data",
            "cpu_pinning": pinned,
            "roofline": {"bound": "fp64", "kernel": "node_jobs_rows_kernel / node_jobs_kernel (the job kernel's two mappings)", "achieved": achieved / 1e12,
                         "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "launches": int(max(n_seen, n_timed)), "avg_launch_us": 1e3 * avg_launch_ms,
                         "flop_per_launch": flop / max(n_seen, n_timed, 1), "kernel_share_of_step": kernel_ms / ms,
                         "timing": "CUDA event pairs (library-owned, on the launching stream) around every %d%s launch of the "
                                   "job kernel inside the timed region: %d of %d launches; an event pair drains the stream "
                                   "and keeps the next kernel from starting in its predecessor's tail, so bracketing every "
                                   "launch slowed the chain it measured by 15 %%" % (
                                       TIMING_STRIDE, "rd" if TIMING_STRIDE == 3 else "th", n_timed, max(n_seen, n_timed)),
                         "peak_source": "pcb_fp64_peak (dependent-DFMA-chain microbenchmark) measured in this run; "
                                        "MEASURED_PEAKS.json has no FP64 figure (hbm_gbs=%s)" % peaks.get("hbm_gbs"),
                         "saturated_batch": sat},
            "cpu_baseline": cpu,
            "chains_on_one_gpu_mps": mps_chains,
            "clocks": clocks.summary(),
            "per_step": dict({k: (s1[k] - s0[k]) / args.steps for k in ("flushes", "launches", "node_jobs", "adds", "scores",
                                                                         "conv_pairs", "memo_hits", "extensions")},
                             **({"sweep_" + k: (s1["sweep"][k] - s0["sweep"][k]) / args.steps for k in s1["sweep"]}
                                if "sweep" in s1 else {})),
            "sweep": args.sweep,
            "sweep_note": "device = conditional-SMC sweeps device-resident with uniform-driven draws (csrc/pcb_sweep.cuh; "
                          "pinned decision for decision against oracle/uniform.py); host = NumPy Generator draws, "
                          "RNG-identical to the reference",
            "other_sweep": other,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_chains_job(args):
    """--chains C: BASELINE configs[2] as written -- the C chains of one run (spawned generators, run.py:87-113) on the
    N GPUs of the job, chain c on GPU c % N, one worker process per chain (the reference's ProcessPoolExecutor,
    run.py:111-149); the workers of a GPU share it through an MPS daemon.  The same C chains run whatever N is, so the
    per-N values are the same work.  Timed on the device (CUDA events around each worker's steps), max over all workers
    of all ranks; no data-path collective."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from multi_chain_bench import run_chains_once

    from phyclone_b200.run import MpsServer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    C = args.chains
    mine = [c for c in range(C) if c % world == rank]
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cores = sorted(os.sched_getaffinity(0))
        block = len(cores) // world
        if block >= 1 and not os.environ.get("PCB_NO_PIN"):
            os.sched_setaffinity(0, cores[local * block:(local + 1) * block])  # inherited by this rank's workers
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local] \
        if os.environ.get("CUDA_VISIBLE_DEVICES") else str(local)  # workers and the MPS daemon see this rank's GPU only

    def go():
        if world > 1:
            dist.barrier()

    with MpsServer(enable=len(mine) > 1) as server:
        if server.started:
            # the daemon was started seeing this rank's GPU only and exposes it as ITS device 0: clients index into the
            # server's enumeration, not the machine's
            os.environ["CUDA_VISIBLE_DEVICES"] = "0"
        rate, ev_ms, splits = run_chains_once(len(mine), args.steps, args.warmup, timeout=900, config=CFG["config_number"],
                                              sweep=args.sweep, device=0, chains=mine, n_chains=C, before_go=go,
                                              details=True)
        mps = server.started
    ms = torch.tensor([ev_ms], dtype=torch.float64, device="cuda" if world > 1 else "cpu")
    per_rank = [ev_ms]
    if world > 1:
        every = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(every, ms)
        per_rank = [float(t.item()) for t in every]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    job_ms = float(ms.item())
    if rank == 0:
        line = {
            "metric": METRIC, "value": C * args.steps / (job_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": job_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": dict(CFG, chains=C),
            "sweep": args.sweep, "chains_per_gpu": [len([c for c in range(C) if c % world == r]) for r in range(world)],
            "mps": mps, "per_rank_ms_per_step": [round(x / args.steps, 2) for x in per_rank],
            "rank0_worker_split_ms_per_step": splits,
            "how": "one worker process per chain, chain c on GPU c mod N, CUDA events around each worker's timed steps, max "
                   "over workers and ranks; strong scaling: the same %d chains at every N" % C,
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
    ap.add_argument("--no-other-sweep", action="store_true", help="skip the comparison leg with the other sweep mode")
    ap.add_argument("--sweep", default="device", choices=["host", "device"],
                    help="SMC sweeps on the host (NumPy Generator draws, RNG-identical to the reference) or device-resident "
                         "(uniform-driven draws, csrc/pcb_sweep.cuh)")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 4, 5],
                    help="BASELINE.json configuration number (1-based; 3 = config 2 with --chains 8)")
    ap.add_argument("--chains", type=int, default=None, help="number of chains of the job (default: one per GPU)")
    ap.add_argument("--port", action="store_true", help="reference arm: time the CPU port instead of baseline/_ref")
    ap.add_argument("--ref-max-seconds", type=float, default=420.0,
                    help="reference arm: wall cap handed to the reference's own --max-time")
    args = ap.parse_args()
    set_config(args.config)
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.chains is not None:
        run_chains_job(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
