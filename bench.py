#!/usr/bin/env python
"""bench.py - RQ-KMeans vectors/sec per iteration (512-d, [128,128,256]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows R]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 1M x 512 fp32 synthetic vectors per GPU (weak scaling: the fit is
ONE balanced K-Means over all ranks' rows, sharded row-wise; per iteration only the K x 512 sums,
K counts and the auction's threshold histograms cross NVLink).  A step = one `fit_by_min_loss`
iteration of the reference (balancekmeans/__init__.py:304-362): fused score pass, balanced auction,
centroid update, loss/shift read-back; steps cycle over the three levels (K = 128 on X, 128 and 256 on
the normalised residuals).  value = rows of all ranks x steps / time (CUDA events, max over ranks).

Extra keys: `e2e` = the same metric through HierarchicalRQKMeans.train() on HOST arrays (H2D of X and
D2H of the ids inside the timed region); `roofline` = the auction's streaming pass kernels (one read of the K x N
fp16 score matrix per launch) against the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle
port of the reference (all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLUSTERS = [128, 128, 256]
DIM = 512
METRIC = "RQ-KMeans vectors/sec per iteration (512-d, [128,128,256])"
UNIT = "vectors/s"


# --------------------------------------------------------------------------------------------
def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed region of
    a default run is ~0.1 s, shorter than nvidia-smi's start-up, so the sampler reads NVML directly (the same
    counters nvidia-smi prints) every 2 ms on a host thread; nvidia-smi -lms is the fallback without pynvml.
    Only samples taken between mark_begin() and mark_end() count."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index = index
        self.samples = []          # (time, sm_mhz, reasons bitmask or set)
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.t0 = self.t1 = None
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"

            def loop():
                while not self.stop_flag:
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            bits = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((time.perf_counter(), mhz, bits))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"

            def read():
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for line in self.proc.stdout:
                    f = [x.strip() for x in line.split(",")]
                    try:
                        mhz, self.max_mhz = float(f[0]), float(f[1])
                    except (ValueError, IndexError):
                        continue
                    bits = sum(self.REASONS[n] for n, v in zip(names, f[2:6]) if v.lower().startswith("active"))
                    self.samples.append((time.perf_counter(), mhz, bits))

            self.thread = threading.Thread(target=read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or x[0])]
        mhz = sorted(x[1] for x in inside)
        bits = 0
        for x in inside:
            bits |= x[2]
        reasons = sorted(n for n, b in self.REASONS.items() if bits & b)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(mhz), "source": self.source}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------
def cpu_baseline_port(sample_rows, seed=1234):
    """One fit_by_min_loss iteration of the oracle port per level on a bounded sample, all host threads.
    Same regime as the bench workload (N % K != 0 -> the reference's 1002-round fallback)."""
    import numpy as np
    from oracle import rqk_oracle as O
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((sample_rows, DIM), dtype=np.float32)
    per_level, rounds = [], []
    cur = x
    for lvl, k in enumerate(CLUSTERS):
        c = cur[rng.choice(sample_rows, k, replace=False)].copy()
        t0 = time.perf_counter()
        d = O.pairwise_distance_full(cur, c, 100000)                     # :308
        res = O.auction_lap_half(-d)                                     # :310
        c1 = O.update_centers(cur, res.assignment, c)                    # :314-324
        d2 = O.pairwise_distance_full(cur, c1, 100000)                   # :327
        cnt = np.bincount(np.argmin(d2, axis=1), minlength=k)            # :328-329
        O.overflow_loss(cnt, 1 << 30)
        O.center_shift(c1, c)
        per_level.append(time.perf_counter() - t0)
        rounds.append(res.rounds)
        if lvl < len(CLUSTERS) - 1:
            cur = O.residual_normalised(cur, np.argmin(d2, axis=1), c1, [DIM])
    total = sum(per_level)
    return {"value": sample_rows * len(CLUSTERS) / total, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "sample": f"{sample_rows} x {DIM} fp32 S-iso rows, one fit iteration per level {CLUSTERS}, "
                      f"auction rounds {rounds}, {total:.1f} s CPU",
            "seconds_per_level": [round(t, 2) for t in per_level]}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself is Python + torch CPU
    and does not exist on the GPU box) on the box's host cores, same metric/unit/config."""
    rank = _env_int("RANK", 0)
    if rank != 0:
        return
    rows = args.cpu_rows
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline_port(rows, seed=1234 + i)
        if i >= args.warmup:
            vals.append(base["value"])
        if sum(base["seconds_per_level"]) * (args.warmup + args.steps - i - 1) > 240:
            break
    v = sum(vals) / max(len(vals), 1) if vals else base["value"]
    steps = max(len(vals), 1)
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * rows / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32 distance / fp16 auction (CPU)", "data": "synthetic",
            "config": workload_config(args, args.gpus), "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"{args.rows} x {DIM} fp32 rows per GPU, 3-level RQ-KMeans {CLUSTERS} balanced fit iteration "
                        f"(score pass + auction + centroid update), levels cycled",
            "rows_per_gpu": args.rows, "rows_total": args.rows * world, "dim": DIM, "codebook": CLUSTERS,
            "sharding": f"rows x{world}", "l2": "inputs (2 GB X + 256/512 MB scores per level) exceed the 126 MB L2"}


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1000000, help="rows per GPU")
    ap.add_argument("--cpu-rows", type=int, default=20032, help="rows of the CPU baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from generative_ranking_recommender_b200 import HierarchicalRQKMeans, HierarchicalRQKMeansConfig, engine
    from generative_ranking_recommender_b200.balancekmeans import KMeans

    rank, world, local = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    shard = engine.ShardGroup() if world > 1 else None
    n = args.rows
    n_global = n * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic level inputs, resident in HBM before anything is timed ----
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    x0 = torch.randn((n, DIM), device=dev, generator=g)
    np.random.seed(42)
    torch.manual_seed(42)
    levels = []
    cur = x0
    for lvl, k in enumerate(CLUSTERS):
        km = KMeans(n_clusters=k, device=dev, balanced=True, shard=shard)
        km.cluster_centers = km.initialize(cur)
        levels.append((km, cur))
        if lvl < len(CLUSTERS) - 1:
            km._iterate(cur, n_global)                       # one real iteration to get sensible centroids
            ids = engine.score_pass(cur, km.cluster_centers, argmin=True).argmin
            cur = engine.residual_normalise(cur, ids, km.cluster_centers, [DIM])
    bufs = [None] * len(levels)
    passes, rounds = [], []

    def step(i):
        lvl = i % len(levels)
        km, xl = levels[lvl]
        score, _assign, stats, _shift = km._iterate(xl, n_global, bufs[lvl])
        bufs[lvl] = score.scores_t
        c = score.counts.to(torch.int64)
        if shard is not None:
            shard.all_reduce(c, "sum")
        c.cpu()                                              # loss read-back of the fit loop (:333-341)
        if stats is not None:
            passes.append(stats.passes)
            rounds.append(stats.rounds)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                      # before the warm-up: it is streaming when the timed region starts
    for i in range(args.warmup):
        step(i)
    passes.clear()
    rounds.clear()
    barrier()
    if rank == 0:
        sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record()
    barrier()
    if rank == 0:
        sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n_global * args.steps / (ms / 1e3)

    # launches of our kernels inside the timed region: per step score pass (pad, minmax, split, tc) = 4,
    # auction init (memset + init) 2 + 5 per round (window sampling, HIST + resolve, tie prefix, bid-list replay +
    # resolve, S-scanning BID fallback; a round = 2 passes; rounds are enqueued two at a time, one batch ahead
    # of the host's look at the state) + finalize 1, centroid update 8
    def auction_launches(p):
        rounds_ = (p + 1) // 2
        return 5 * 2 * (-(-rounds_ // 2) + 1)
    gpu_launches = int(sum(4 + 2 + auction_launches(p) + 1 + 8 for p in passes)) if passes else 0

    # ---- roofline of the dominant kernel: the auction's streaming HIST pass at level 0 (K=128) ----
    # It reads the K x N fp16 score matrix exactly once: 2*K bytes per vector.  The bidding half of
    # a round replays the HIST pass's survivor lists (L2-resident) and is reported beside it.
    roofline = None
    if rank == 0:
        km, xl = levels[0]
        sc = engine.score_pass(xl, km.cluster_centers, scores=True, argmin=False)
        sess = engine.AuctionSession(sc.scores_t, n, n)
        sess.init(sc.minmax)
        t_hist, t_bid = [], []
        prev = sess.poll()
        for cyc in range(40):
            sess.do_pass(1)                                  # window sampling (early-exits unless needed), untimed
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            sess.do_pass(2)                                  # streaming HIST kernel (returns at once in a BID cycle)
            e1.record()
            sess.do_pass(4)                                  # bid-list replay (+ the S-scanning fallback, which returns at once)
            e2.record()
            sess.resolve()
            cur = sess.poll()
            if cur.done:
                break
            if cyc >= 4:                                     # past the cold start
                if cur.cold_passes > prev.cold_passes:
                    t_hist.append(e0.elapsed_time(e1))
                else:
                    t_bid.append(e1.elapsed_time(e2))
            prev = cur
        k0 = CLUSTERS[0]
        alg_bytes = 2.0 * k0 * n
        peak, how = measured_peaks()

        def line(name, ts):
            if not ts:
                return None
            t = sum(ts) / len(ts)
            ach = alg_bytes / (t * 1e-3) / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": recorded_traffic(name.split("(")[0].strip()), "algorithmic_bytes_per_launch": alg_bytes,
                    "ms_per_launch": t, "launches_timed": len(ts), "peak_source": how}

        roofline = line("auction_hist_kernel (threshold select, K=128)", t_hist)
        if roofline is not None and t_bid:
            roofline["other_kernel"] = {
                "kernel": "auction_bidlist_kernel (bids replayed from the HIST pass's survivor lists, K=128)",
                "ms_per_launch": sum(t_bid) / len(t_bid), "launches_timed": len(t_bid),
                "note": "not a stream over S: reads ~2 % of it as L2-resident lists plus cost/owner of every job; "
                        "timed together with the early-exiting S-scanning fallback kernel"}

    # ---- encode (the KMeans.predict + residual chain train() emits its ids with), rows resident in HBM ----
    encode = None
    if rank == 0:
        try:
            cs = [km.cluster_centers for km, _ in levels]
            for _ in range(2):
                engine.encode(x0, cs, CLUSTERS, [DIM], mode=0)
            ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ee0.record()
            for _ in range(3):
                engine.encode(x0, cs, CLUSTERS, [DIM], mode=0)
            ee1.record()
            torch.cuda.synchronize()
            t_enc = ee0.elapsed_time(ee1) / 3
            encode = {"value": n / (t_enc / 1e3), "unit": "vectors/s", "rows": n, "ms": t_enc, "gpus": 1,
                      "what": "3-level ids of resident fp32 rows (rqk_encode, mode 0)"}
        except Exception as e:      # never lose the bench line over the side metric
            encode = {"error": repr(e)}

    # ---- end to end through the public API: host array in, ids out ----
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((n, DIM), dtype=torch.float32, pin_memory=True)
        xh.copy_(x0)
        x_np = xh.numpy()
        del levels, bufs, cur, x0
        engine.SCRATCH.clear()
        torch.cuda.empty_cache()
        cfg = HierarchicalRQKMeansConfig(layer_clusters=CLUSTERS, need_clusters=CLUSTERS, embedding_dim=DIM,
                                         group_dims=[DIM], hierarchical_weights=[[1.0]] * 3, iter_limit=20)
        np.random.seed(42)
        torch.manual_seed(42)
        model = HierarchicalRQKMeans(cfg, device=dev, shard=shard)
        barrier()
        t0 = time.perf_counter()
        out = model.train(x_np, resume=False)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        iters = sum(len(s) for s in model.fit_stats)
        ids_bytes = sum(t.numel() * t.element_size() for t in out["cluster_ids"])
        e2e = {"value": n_global * iters / dt, "unit": UNIT, "h2d_bytes_per_step": x_np.nbytes * world / max(iters, 1),
               "d2h_bytes_per_step": ids_bytes * world / max(iters, 1), "seconds": dt, "iterations": iters,
               "iterations_per_level": [len(s) for s in model.fit_stats],
               "api": "HierarchicalRQKMeans.train(np.ndarray) -> cluster_ids (int64, host), iter_limit=20"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline_port(args.cpu_rows)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "3xTF32 distance (fp32 accumulate), fp16 auction, fp32 centroids",
                "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
                "gpu_launches": gpu_launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
                "auction": {"passes_per_step": passes, "reference_rounds_per_step": rounds}, "encode": encode}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
