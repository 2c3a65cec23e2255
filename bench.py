#!/usr/bin/env python
"""bench.py -- image pairs/sec of exhaustive u8 SIFT-128 putative matching (BF squared-L2 2-NN + ratio test).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json): 10,000 u8 SIFT-128 descriptors per image, exhaustive pairs, ratio 0.8.  At N=1 this is
config 3 (100 images, 4,950 pairs).  For N>1 the pair list is sharded with no collective (pairs are independent);
the collection grows as ~sqrt(N) (142 / 200 / 282 images) so that every GPU keeps ~4,950 pairs: weak scaling.
Every rank holds a replica of the descriptor arena (the north star's pair scheduler).

One "step" = one pass of the hot path over this rank's pair shard.
  value  whole-job pairs/s with descriptors already resident in HBM (device time, CUDA events, max over ranks)
  e2e    same metric through the public collection API (MatcherCudaAllInMemory mirror -> C ABI) from PINNED HOST
         buffers: H2D of all descriptors + kernels + D2H of matches + host coordinate de-dup, every step
  roofline       the fused tcgen05 kernel: algorithmic int8 ops (2*nI*nJ*128 per pair) / its CUDA-event time
  cpu_baseline   the reference's own CPU brute-force code (oracle/_ref, built from /root/reference) on a bounded sample
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS = 10000
RATIO = 0.8
PAIRS_PER_GPU = 4950
SYNTH_CONFIG = 3
METRIC = "image pairs/sec, 10k u8 SIFT/img exhaustive BF-L2+ratio"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_reference_run(nq=ROWS, flann=False):
    """Time the reference's own CPU path (oracle/_ref; the L1 port if it is absent): one (10k-row db, nq-query) unit per
    host thread, nq <= 10000 (a full pair when nq == 10000; cost is linear in nq)."""
    from oracle import oracle
    synth = importlib.import_module("3dreconstruction_b200.synth")
    pkg = importlib.import_module("3dreconstruction_b200")
    cores = os.cpu_count() or 1
    imgs = synth.collection(SYNTH_CONFIG, 4, ROWS)
    rs = float(pkg.square_f32(RATIO))
    use_ref = os.path.exists(oracle.L0_OMP_PATH)
    n = max(1, min(cores, 512))
    combos = [(a, b) for a in range(4) for b in range(4) if a != b]
    dbs = [imgs[combos[k % len(combos)][0]] for k in range(n)]
    qs = [imgs[combos[k % len(combos)][1]][:nq] for k in range(n)]
    t0 = time.time()
    if use_ref:
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        oracle.L0(openmp=True).bench(dbs, qs, rs, flann=flann)
        kind = "reference"
    else:
        oracle.L1().bench_bf(dbs, qs, rs)
        kind = "port"
    dt = time.time() - t0
    pairs_equiv = n * nq / float(ROWS)
    return {"value": pairs_equiv / dt, "unit": "pairs/s", "cores": cores, "kind": kind, "seconds": dt, "pairs_equiv": pairs_equiv,
            "sample": f"{n} units of (10000-row db x {nq} queries) = {pairs_equiv:.2f} pairs of 10000x10000 u8 SIFT-128, "
                      f"{'FLANN kd-tree' if flann else 'brute force'} + ratio test, "
                      f"{'reference code in its OpenMP mode (-DUSE_OPENMP), one unit per thread' if use_ref else 'oracle L1 port, OpenMP over queries'}"}


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return 0
    steps = max(args.steps, 1)
    cal = cpu_reference_run(nq=256)                      # calibration, ~0.5 s
    per_unit_pair_s = cal["seconds"] / (256.0 / ROWS)    # seconds for one full pair per thread
    budget = 120.0 / (steps + max(args.warmup, 0))       # whole run within a few minutes
    nq = int(min(ROWS, max(256, ROWS * budget / per_unit_pair_s)))
    for _ in range(max(args.warmup, 0)):
        cpu_reference_run(nq=nq)
    runs = [cpu_reference_run(nq=nq) for _ in range(steps)]
    total_s = sum(x["seconds"] for x in runs)
    val = sum(x["pairs_equiv"] for x in runs) / total_s
    r = runs[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (the reference accumulates exact integers in float, metric.h:51-82)", "data": "synthetic",
        "config": {"workload": f"bounded sample of BASELINE config 3 per step: {r['sample']}, ratio {RATIO}",
                   "rows_per_image": ROWS, "ratio": RATIO},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local = _dist_env()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    pkg = importlib.import_module("3dreconstruction_b200")
    sharding = importlib.import_module("3dreconstruction_b200.sharding")
    synth = pkg.synth

    n_images = sharding.images_for_pairs_per_gpu(world, PAIRS_PER_GPU)
    descs = synth.collection(SYNTH_CONFIG, n_images, ROWS)
    # pinned host staging, so the e2e leg's H2D is a true async PCIe copy
    pinned = [torch.empty((ROWS, 128), dtype=torch.uint8).pin_memory() for _ in range(n_images)]
    for t, d in zip(pinned, descs):
        t.numpy()[:] = d
    descs = [t.numpy() for t in pinned]
    feats = [synth.features(SYNTH_CONFIG, k, ROWS)[:, :2].copy() for k in range(n_images)]
    rows = [ROWS] * n_images
    all_pairs = pkg.pairs_exhaustive(n_images)
    my_pairs, (b0, b1) = sharding.shard_pairs(all_pairs, rows, rank, world)
    rs = float(pkg.square_f32(RATIO))

    ctx = pkg.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    info = ctx.device_info()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- value: descriptors resident in HBM
    ctx.upload_images(descs, pinned=True)
    for _ in range(max(args.warmup, 3)):
        ctx.match_pairs(my_pairs, rs, collect=False)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    ev0.record(stream)
    knn_ms, knn_launches, launches, n_matches, rescanned = 0.0, 0, 0, 0, 0
    for _ in range(args.steps):
        pm = ctx.match_pairs(my_pairs, rs, collect=False)
        knn_ms += pm.knn_kernel_ms
        knn_launches += pm.knn_kernel_launches
        launches += pm.total_launches
        n_matches = int(pm.offsets[pm.n_pairs])
        rescanned = int(pm.rescanned_queries)
    ev1.record(stream)
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1)
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    total_pairs = sum_over_ranks(float(len(my_pairs)))
    value = total_pairs * args.steps / (ms * 1e-3)
    gpu_launches = int(sum_over_ranks(float(launches)))

    # ---------------- e2e: host buffers -> public collection API -> matches on the host, every step
    host_threads = max(1, (os.cpu_count() or 1) // world)   # the ranks of one box share its host cores
    matcher = pkg.MatcherCudaAllInMemory(RATIO, ctx, host_threads=host_threads)
    for _ in range(1):
        matcher.LoadArrays(descs, feats)
        matcher._ctx.match_collection(my_pairs, rs, host_threads, collect=False)
    barrier()
    t0 = time.time()
    e2e_matches = 0
    # N > 1: rank 0 alone reads the collection over PCIe; the replicas of the other GPUs arrive over NVLink (one NCCL
    # broadcast of the arena), instead of N concurrent PCIe uploads of the same bytes
    stage = torch.empty((n_images * ROWS, 128), dtype=torch.uint8, device=f"cuda:{local}") if world > 1 else None
    host_all = None
    if world > 1 and rank == 0:
        host_all = torch.empty((n_images * ROWS, 128), dtype=torch.uint8).pin_memory()
        for k, d in enumerate(descs):
            host_all[k * ROWS:(k + 1) * ROWS].numpy()[:] = d

    def load_step():
        if world == 1:
            matcher.LoadArrays(descs, feats)                              # H2D of every descriptor array (+ norms kernel)
            return
        if rank == 0:
            stage.copy_(host_all, non_blocking=True)                      # the one H2D of the job
        dist.broadcast(stage, src=0)                                      # NVLink / NVSwitch
        torch.cuda.current_stream().synchronize()
        ctx.upload_images_device([stage.data_ptr() + k * ROWS * 128 for k in range(n_images)], rows)   # D2D into the arena
        ctx.set_features(feats)

    if world > 1:
        load_step()
        ctx.match_collection(my_pairs, rs, host_threads, collect=False)
        barrier()
        t0 = time.time()
    for _ in range(args.steps):
        load_step()
        pm = ctx.match_collection(my_pairs, rs, host_threads, collect=False)   # kernels + D2H + host de-dup (row 13)
        e2e_matches = int(pm.offsets[pm.n_pairs])
    barrier()
    e2e_s = max_over_ranks(time.time() - t0)
    e2e_value = total_pairs * args.steps / e2e_s
    h2d = n_images * ROWS * 128 + len(my_pairs) * 24 + (len(my_pairs) + 1) * 4   # per step; at N > 1 only rank 0 reads the descriptors over PCIe
    d2h = n_matches * 8 + len(my_pairs) * 4 + (len(my_pairs) + 1) * 8 + 8

    # ---------------- roofline of the dominant kernel (this rank)
    peaks, peak_src = _peaks()
    ops_per_pair = 2.0 * ROWS * ROWS * 128
    achieved = ops_per_pair * len(my_pairs) * args.steps / (knn_ms * 1e-3) / 1e12 if knn_ms > 0 else 0.0
    peak = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    probe_ops, _ = ctx.probe_i8_peak(20000)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "knn2_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "tensor", "kernel": "knn2_kernel (tcgen05 kind::i8 GEMM + fused top-2 epilogue)",
        "achieved": achieved, "peak": peak, "unit": "TOP/s (int8, 2 ops per MAC)", "frac": achieved / peak if peak else None,
        "peak_source": f"2 x bf16_tflops_sustained of MEASURED_PEAKS.json ({peak_src}); kind::i8 issues at twice the bf16 rate",
        "frac_of_nominal_4500": achieved / 4500.0, "probe_i8_tops_mma_only_burst": probe_ops / 1e12,
        "frac_of_probe": achieved / (probe_ops / 1e12) if probe_ops else None,
        "algorithmic_ops_per_launch": ops_per_pair * len(my_pairs), "launch_ms": knn_ms / max(knn_launches, 1),
        "share_of_step": knn_ms / ev0.elapsed_time(ev1), "traffic": traffic,
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cal = cpu_reference_run(nq=256)
        nq = int(min(ROWS, max(256, ROWS * 15.0 / (cal["seconds"] / (256.0 / ROWS)))))   # ~15 s of CPU work
        cpu = cpu_reference_run(nq=nq)
        cpu.pop("seconds", None)
        cpu.pop("pairs_equiv", None)
        try:  # the reference's shipped default matcher (FLANN kd-tree, approximate + randomised): speed only
            fl = cpu_reference_run(nq=ROWS, flann=True)
            cpu["flann_kdtree"] = {"value": fl["value"], "unit": "pairs/s", "cores": fl["cores"], "sample": fl["sample"],
                                   "note": "ArrayMatcherKdtreeFlann<uchar, flann::L2<uchar>>, 4 trees / 128 checks (matcher_kdtree_flann.h:47-49,115); results are approximate"}
        except Exception as e:  # FLANN not compiled into oracle/_ref
            cpu["flann_kdtree"] = {"unavailable": str(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 operands, exact int32 accumulate/compare; fp32 only in the ratio test", "data": "synthetic",
            "config": {"workload": f"exhaustive pairs over {n_images} images x {ROWS} u8 SIFT-128 (BASELINE config 3 at 1 GPU: 100 images, "
                                   f"4950 pairs); {int(total_pairs)} pairs total, {len(my_pairs)} on rank 0, ratio {RATIO}",
                       "n_images": n_images, "rows_per_image": ROWS, "pairs_total": int(total_pairs), "ratio": RATIO,
                       "sharding": "replicated descriptor arena, contiguous cost-balanced split of the pair list, no collective",
                       "l2": "per step 0.13 GB of descriptors + 0.79 GB of per-query records stream through HBM, larger than the 126 MB L2; no explicit flush",
                       "matches_per_step_rank0": n_matches,
                       "second_pass": f"{rescanned} of {len(my_pairs) * ROWS} queries of rank 0 matched twice per step (ambiguous after pruning, prune_rho 0.8); their time is inside the timed region, their ops are not credited",
                       "device": info["name"]},
            "clocks": clocks, "gpu_launches": gpu_launches,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "includes": ("H2D of all descriptors from pinned host memory" + (" on rank 0 + NCCL broadcast of the arena over NVLink to the other ranks" if world > 1 else "")) + ", norm kernel, matching kernels, D2H of matches, "
                                f"host coordinate de-dup (IndexedMatchDecorator) on {host_threads} host threads per rank, overlapped with the GPU batches",
                    "matches": e2e_matches},
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
