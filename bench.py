#!/usr/bin/env python
"""Benchmark of the sknnr query-time hot path (kneighbors + predict, k=7) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], "C3"): EuclideanKNNRegressor, 10M query rows x 50k
reference plots x 32 features, k=7, predict(weights="distance") over 8 targets; synthetic
N(0,1) data with the seeds of BASELINE.md section 3.  One "step" = one pass of the hot path
over the whole 10M-row query batch.  N > 1: every rank holds the full reference set and its
own 10M-row shard (weak scaling), results are gathered to rank 0 with NCCL inside the timed
region.

`value`  : device-resident inputs/outputs (queries/s, whole job).
`e2e`    : the same through the host-buffer C-ABI call a user of the estimators makes, with
           pinned host inputs/outputs and the H2D/D2H copies inside the timed region.
`roofline`: the dominant kernel (fused distance + top-k) against the measured FP32 FMA peak.
`cpu_baseline` / `--impl reference`: the arithmetic the reference itself runs
           (scikit-learn's brute KNeighborsRegressor + sknnr's re-ordering glue, restated in
           oracle/), timed on the box's host cores on a bounded sample.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "query-neighbors/sec (kneighbors+predict, k=7)"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-queries", type=int, default=10_000_000)
    ap.add_argument("--n-ref", type=int, default=50_000)
    ap.add_argument("--dim", type=int, default=32)
    ap.add_argument("--n-out", type=int, default=8)
    ap.add_argument("--k", type=int, default=7)
    ap.add_argument("--cpu-sample", type=int, default=1_500_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--kc", type=int, default=0)
    ap.add_argument("--tc-seed-stride", type=int, default=-1)
    ap.add_argument("--tc-streams", type=int, default=0)
    ap.add_argument("--tc-debug", type=int, default=0)
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--host-slots", type=int, default=0)
    return ap.parse_args()


def workload_name(a):
    return (f"C3 EuclideanKNNRegressor {a.n_queries} queries x {a.n_ref} reference plots x "
            f"{a.dim} features, k={a.k}, predict(weights='distance') over {a.n_out} targets")


def make_reference_set(a):
    """Seeds of BASELINE.md section 3: refs seed 0, y seed 1."""
    R = np.random.default_rng(0).standard_normal((a.n_ref, a.dim))
    y = np.random.default_rng(1).standard_normal((a.n_ref, a.n_out))
    mean, scale = R.mean(axis=0), R.std(axis=0, ddof=1)
    return R, y, mean, scale


def query_block(a, block, rows, rank=0):
    """Queries: N(0,1) in 1M-row blocks from SeedSequence(2).spawn (rank-offset per GPU)."""
    ss = np.random.SeedSequence(2).spawn(block + 1 + 1000 * rank)[-1]
    return np.random.default_rng(ss).standard_normal((rows, a.dim))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_leg(a, rows, repeats=1):
    """Time the reference's own arithmetic for the path on host cores: sklearn brute
    kneighbors + sknnr re-ordering + distance-weighted predict (oracle/sknnr_oracle.py)."""
    from oracle import sknnr_oracle as orc

    R, y, mean, scale = make_reference_set(a)
    fit_Z = (R - mean) / scale
    Q = query_block(a, 0, rows)
    from sklearn.neighbors import KNeighborsRegressor

    reg = KNeighborsRegressor(n_neighbors=a.k, algorithm="brute", weights="distance").fit(fit_Z, y)
    best = float("inf")
    # every host thread the process may use, whatever OMP_NUM_THREADS says (torchrun exports
    # OMP_NUM_THREADS=1 to its workers, which would make this a one-core number)
    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=usable_cores()):
        for _ in range(repeats):
            t0 = time.perf_counter()
            Z = orc.affine_project(Q, mean, scale, None)
            dist, idx = reg.kneighbors(Z)
            dist, idx = orc.deterministic_order(dist, idx)
            orc.weighted_average(y, idx, orc.get_weights(dist, "distance"))
            best = min(best, time.perf_counter() - t0)
    return rows / best, best


def usable_cores():
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return max(os.cpu_count() or 1, 1)


def host_threads():
    """(threads the CPU leg runs with, cores of the box)."""
    return usable_cores(), os.cpu_count()


def run_reference(a, rank):
    if rank != 0:
        return
    rows = min(a.cpu_sample, a.n_queries)
    times = []
    for _ in range(a.warmup):
        cpu_reference_leg(a, min(rows, 20_000))
    for _ in range(a.steps):
        _, t = cpu_reference_leg(a, rows)
        times.append(t)
    total = sum(times)
    value = rows * a.steps / total
    import sklearn

    threads, cores = host_threads()
    sample = (f"{rows} of {a.n_queries} query rows per step (brute cost is linear in n_q); "
              f"scikit-learn {sklearn.__version__} KNeighborsRegressor(algorithm='brute') "
              f"EuclideanArgKmin64 + sknnr ordering/predict glue restated in oracle/")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "os_cpu_count": cores,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(a, rank, world, local_rank):
    import torch

    from sknnr_b200 import _lib as L
    from sknnr_b200._engine import KNNIndex

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    L.set_option("timing", 1)
    if a.engine:
        L.set_option("engine", a.engine)
    if a.kc:
        L.set_option("kc", a.kc)
    if a.tc_streams:
        L.set_option("tc_streams", a.tc_streams)
    if a.tc_debug:
        L.set_option("tc_debug", a.tc_debug)
    if a.chunk_rows:
        L.set_option("chunk_rows", a.chunk_rows)
    if a.host_slots:
        L.set_option("host_slots", a.host_slots)
    if a.tc_seed_stride >= 0:
        L.set_option("tc_seed_stride", a.tc_seed_stride)

    R, y, mean, scale = make_reference_set(a)
    index = KNNIndex((R - mean) / scale, mean, scale, None, y, device=local_rank)

    # queries: host pinned (e2e leg) and device resident (value leg), float64 like the
    # reference's inputs
    n_q, d, k, n_out = a.n_queries, a.dim, a.k, a.n_out
    X_host = torch.empty((n_q, d), dtype=torch.float64, pin_memory=True)
    xh = X_host.numpy()
    blk = 1_000_000
    for b, s in enumerate(range(0, n_q, blk)):
        rows = min(blk, n_q - s)
        xh[s:s + rows] = query_block(a, b, rows, rank)
    X_dev = X_host.to(dev, non_blocking=False)
    o_dist = torch.empty((n_q, k), dtype=torch.float64, device=dev)
    o_idx = torch.empty((n_q, k), dtype=torch.int64, device=dev)
    o_pred = torch.empty((n_q, n_out), dtype=torch.float64, device=dev)
    gathered = None
    if world > 1 and rank == 0:
        gathered = [
            [torch.empty_like(t) for _ in range(world)] for t in (o_dist, o_idx, o_pred)
        ]
    stream = torch.cuda.current_stream(dev)

    def step_device():
        index.query_device(X_dev.data_ptr(), False, n_q, d, k, dist_ptr=o_dist.data_ptr(),
                           idx_ptr=o_idx.data_ptr(), pred_ptr=o_pred.data_ptr(),
                           weights="distance", row_offset=rank * n_q, stream=stream.cuda_stream)
        if world > 1:
            import torch.distributed as dist

            for t, g in zip((o_dist, o_idx, o_pred), gathered or (None, None, None)):
                dist.gather(t, g, dst=0)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(a.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # the library keeps per-call stats; accumulate the search-kernel time of the timed steps
    search_ms, launches, fallbacks = 0.0, 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * n_q * a.steps / (ms_total * 1e-3)

    # dominant-kernel timing: one more device-resident step; the library brackets every search
    # kernel launch with CUDA events on the launching stream (option "timing") and the stats
    # query sums them.  (The host-buffer call below overlaps chunks on several streams, so its
    # per-kernel brackets would include waiting for the other streams' kernels.)
    step_device()
    barrier()
    dev_stats = index.stats()
    search_ms_dev = dev_stats["search_ms"]

    # e2e through the host-buffer call (pinned host in/out, copies inside the timed region)
    e2e = None
    stats = {}
    if not a.no_e2e:
        h_dist = torch.empty((n_q, k), dtype=torch.float64, pin_memory=True)
        h_idx = torch.empty((n_q, k), dtype=torch.int64, pin_memory=True)
        h_pred = torch.empty((n_q, n_out), dtype=torch.float64, pin_memory=True)
        import ctypes as C

        def step_host():
            L.check(index._lib.sknnr_kneighbors(
                index._h, C.c_void_p(X_host.data_ptr()), L.F64, n_q, d, rank * n_q, k,
                L.DETERMINISTIC, 10, C.c_void_p(h_dist.data_ptr()), C.c_void_p(h_idx.data_ptr()),
                L.W_DISTANCE, C.c_void_p(h_pred.data_ptr()), None))

        step_host()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(a.steps, 3))
        for _ in range(e2e_steps):
            step_host()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        stats = index.stats()
        search_ms, launches, fallbacks = search_ms_dev, stats["kernel_launches"], stats["n_fallback"]
        e2e = {"value": world * n_q * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(stats["h2d_bytes"]), "d2h_bytes_per_step": int(stats["d2h_bytes"]),
               "steps": e2e_steps, "timer": "host wall clock around the synchronous C-ABI call"}
        # device-path results must equal host-path results bit for bit
        assert torch.equal(h_idx.to(dev), o_idx), "device and host paths disagree"
    else:
        # still need the search-kernel time: run the host path on a slice
        stats = dev_stats
        search_ms = search_ms_dev
        launches, fallbacks = stats["kernel_launches"], stats["n_fallback"]

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (fused distance + top-k) ----
    engine = int(stats.get("engine", 0))
    dpad = (d + 7) // 8 * 8
    flops = 2.0 * d * n_q * a.n_ref                      # algorithmic: 2 * d' * n_q * n_ref
    # search launches of the device-resident step: device-pointer calls use chunks of 2 x chunk_rows
    chunk_dev = 2 * (a.chunk_rows if a.chunk_rows > 0 else (1 << 20))
    n_chunks = max(1, -(-n_q // chunk_dev))
    achieved = flops / (search_ms * 1e-3) / 1e12 if search_ms > 0 else None
    fp32_peak = L.measure_fp32_peak(local_rank)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm = {"algorithmic_bytes_per_step": n_q * (8 * d + k * 16 + 8 * n_out) + 4 * a.n_ref * dpad,
           "peak_gbs": peaks.get("hbm_gbs"), "peak_source": "MEASURED_PEAKS.json" if peaks else "absent"}
    if engine == L.ENGINE_TENSOR:
        # tcgen05 kind::tf32 runs at half the bf16 rate; the launches are timed inside a long,
        # power-capped step, so the denominator is the SUSTAINED cuBLAS bf16 figure / 2
        bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        tf32_peak = bf16 / 2.0
        roofline = {
            "kernel": "search_tc_kernel", "bound": "tensor", "achieved": achieved, "peak": tf32_peak,
            "unit": "TFLOP/s", "frac": (achieved / tf32_peak) if achieved else None,
            "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 = dense TF32, of measured" if peaks
                            else "fallback 1400 bf16 sustained / 2 (B200_PROFILING.md), of fallback"),
            # dram__bytes_read.sum + dram__bytes_write.sum of one 2^20-row launch, ncu --set full
            # (profiles/r01_search_tc_v11.md); scaled to this run's rows per launch
            "traffic": 231.9e6 * (n_q / n_chunks) / float(1 << 20) if d == 32 else None,
            "algorithmic_flops_per_launch": flops / n_chunks,
            "launches_per_step": n_chunks, "kernel_ms_per_step": search_ms,
            "fp32_simt_peak_measured": fp32_peak,
            "note": "algorithmic FLOPs 2*d'*n_q*n_ref over the CUDA-event time of the search kernel "
                    "launches of one device-resident step; the kernel executes (d'+8)/d' of them (|r|^2 "
                    "folded into the MMA) plus a 25 % threshold-seeding pre-pass, and is paced by its "
                    "TMEM epilogue (thread <-> query min tree + hit path), see DESIGN.md section 4",
            "hbm": hbm,
        }
    else:
        roofline = {
            "kernel": "search_simt_kernel" if engine == L.ENGINE_SIMT else f"engine{engine}",
            "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": (achieved / fp32_peak) if achieved else None,
            "peak_source": "measured in this run: register-resident FFMA2 probe on all SMs "
                           "(sknnr_measure_fp32_peak); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
            "traffic": None,
            "algorithmic_flops_per_launch": flops / n_chunks,
            "launches_per_step": n_chunks, "kernel_ms_per_step": search_ms, "hbm": hbm,
        }
    if roofline["hbm"]["peak_gbs"] and search_ms > 0:
        roofline["hbm"]["achieved_gbs_whole_step"] = (
            roofline["hbm"]["algorithmic_bytes_per_step"] / (ms_total / a.steps * 1e-3) / 1e9)

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        rows = min(a.cpu_sample, n_q)
        cpu_reference_leg(a, min(rows, 20_000))
        v, t = cpu_reference_leg(a, rows)
        import sklearn

        threads, cores = host_threads()
        cpu = {"value": v, "unit": UNIT, "cores": threads, "os_cpu_count": cores, "kind": "port",
               "sample": f"{rows} of {n_q} query rows, {t:.2f} s; scikit-learn {sklearn.__version__} "
                         "brute KNeighborsRegressor (EuclideanArgKmin64) + sknnr ordering/predict "
                         "glue restated in oracle/; linear in n_q"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": ("tf32 filter + f64 refine" if engine == L.ENGINE_TENSOR else "f32 filter + f64 refine"),
        "data": "synthetic",
        "config": {"workload": workload_name(a), "n_ref": a.n_ref, "dim": a.dim, "k": k,
                   "queries_per_gpu": n_q, "l2": f"inputs {n_q * d * 8 / 1e9:.2f} GB per step exceed the 126 MB L2 (no flush needed)",
                   "multi_gpu": "queries sharded, reference set replicated, results gathered to rank 0 with NCCL inside the timed region"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches * a.steps), "fallback_rows_per_step": int(fallbacks),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    run_ours(a, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
