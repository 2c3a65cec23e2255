"""BASELINE configs[3] / configs[4] at FULL size, strong-scaled over the GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
        tests/tools/config_full.py --config 4          # 1,000 images x 8,000 SIFT  -> 499,500 pairs
        tests/tools/config_full.py --config 5          #   200 images x 40,000 SIFT ->  19,900 pairs

Every rank generates 1/N of the seeded synthetic collection, the descriptor blocks are all-gathered over NVLink into a
staging buffer and copied device-to-device into each GPU's replicated arena (no collective on the data
path -- this is the one-time load).  Each rank then matches its contiguous, cost-balanced shard of the i<j pair list:
once at pair level (kernels only, device-timed) and once at collection level (kernels + D2H + host de-dup row 13).
Checks: (i) a seeded sample of pairs of every rank bit-exact against the CPU oracle (rows 7-13), (ii) the pair-level
and collection-level runs agree on every pair that de-dup-2 leaves untouched, (iii) totals are summed over ranks.
This is a measurement/acceptance script, not a bench line (bench.py keeps configs[2]); logs go to profiles/."""
import argparse, importlib, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dreconstruction_b200")
sharding = importlib.import_module("3dreconstruction_b200.sharding")
from oracle import oracle  # noqa: E402  (checker only)

SHAPES = {4: (1000, 8000), 5: (200, 40000), 3: (100, 10000)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4, choices=sorted(SHAPES))
    ap.add_argument("--images", type=int, default=0)
    ap.add_argument("--check-pairs", type=int, default=3, help="pairs per rank checked against the oracle")
    args = ap.parse_args()
    n_img, rows = SHAPES[args.config]
    if args.images:
        n_img = args.images
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def log(*a):
        if rank == 0:
            print(*a, flush=True)

    # ---------------- generate 1/N of the collection per rank, all-gather over NVLink
    per = (n_img + world - 1) // world
    lo, hi = min(rank * per, n_img), min((rank + 1) * per, n_img)
    t0 = time.time()
    pool = pkg.synth.scene_pool(args.config, rows)
    mine = np.zeros((per, rows, 128), np.uint8)
    fmine = np.zeros((per, rows, 2), np.float32)
    for k in range(lo, hi):
        mine[k - lo] = pkg.synth.image(args.config, k, rows, pool)
        fmine[k - lo] = pkg.synth.features(args.config, k, rows)[:, :2]
    t_gen = time.time() - t0
    barrier(); t0 = time.time()
    stage = torch.empty((world * per, rows, 128), dtype=torch.uint8, device="cuda")
    fstage = torch.empty((world * per, rows, 2), dtype=torch.float32, device="cuda")
    if world > 1:
        dist.all_gather_into_tensor(stage, torch.from_numpy(mine).cuda())
        dist.all_gather_into_tensor(fstage, torch.from_numpy(fmine).cuda())
    else:
        stage.copy_(torch.from_numpy(mine)); fstage.copy_(torch.from_numpy(fmine))
    torch.cuda.synchronize()
    ctx = pkg.Context(local)
    rows_list = [rows] * n_img
    ctx.upload_images_device([stage.data_ptr() + k * rows * 128 for k in range(n_img)], rows_list)
    feats_all = fstage[:n_img].cpu().numpy()
    feats = [feats_all[k] for k in range(n_img)]
    ctx.set_features(feats)
    barrier(); t_load = time.time() - t0
    del stage
    log(f"config {args.config}: {n_img} images x {rows} rows on {world} GPU(s); generated {hi - lo} images/rank in {t_gen:.1f} s; "
        f"all-gather + arena copy + features {t_load:.2f} s; arena {n_img * rows * 128 / 1e9:.2f} GB per GPU")

    pairs = pkg.pairs_exhaustive(n_img)
    my_pairs, (p_lo, p_hi) = sharding.shard_pairs(pairs, rows_list, rank, world)
    rs = float(pkg.square_f32(0.8))
    host_threads = max(1, (os.cpu_count() or 1) // world)
    ops_per_pair = 2.0 * rows * rows * 128

    # ---------------- pair level (rows 7-12): kernels only, device time
    ctx.match_pairs(my_pairs[: max(1, len(my_pairs) // 50)], rs, collect=False)  # warm-up
    barrier(); t0 = time.time()
    pm = ctx.match_pairs(my_pairs, rs, collect=True)
    torch.cuda.synchronize(); t_pair_wall = time.time() - t0
    gpu_ms = reduce(pm.timing["gpu_ms"], dist.ReduceOp.MAX)
    knn_ms = reduce(pm.timing["knn_kernel_ms"], dist.ReduceOp.MAX)
    wall_pair = reduce(t_pair_wall, dist.ReduceOp.MAX)
    n_put = reduce(float(pm.offsets[-1]), dist.ReduceOp.SUM)
    launches = reduce(float(pm.timing["knn_kernel_launches"]), dist.ReduceOp.SUM)
    log(f"pair level : {len(pairs)} pairs, GPU time {gpu_ms:.0f} ms (max over ranks; knn kernel {knn_ms:.0f} ms, {int(launches)} launches) "
        f"-> {len(pairs) / (gpu_ms * 1e-3):,.0f} pairs/s device-timed, {ops_per_pair * len(pairs) / (knn_ms * 1e-3) / 1e12 / world:,.0f} int8 TOP/s per GPU; "
        f"wall incl. D2H {wall_pair:.2f} s; {int(n_put):,} matches after rows 10-12")
    pair_counts = pm.counts.copy()

    # ---------------- collection level (rows 7-13): + host de-dup on the rank's share of the host cores
    barrier(); t0 = time.time()
    cm = ctx.match_collection(my_pairs, rs, host_threads, collect=True)
    t_coll = reduce(time.time() - t0, dist.ReduceOp.MAX)
    n_coll = reduce(float(cm.offsets[-1]), dist.ReduceOp.SUM)
    log(f"collection : {len(pairs)} pairs end to end (kernels + D2H + host de-dup, {host_threads} host threads/rank) {t_coll:.2f} s "
        f"-> {len(pairs) / t_coll:,.0f} pairs/s; {int(n_coll):,} putative matches")

    # ---------------- checks
    ok = bool(np.all(cm.counts <= pair_counts))  # de-dup-2 only removes
    l1 = oracle.L1()
    rng = np.random.default_rng(100 + rank)
    sample = rng.choice(len(my_pairs), min(args.check_pairs, len(my_pairs)), replace=False)
    t0 = time.time()
    descs_cache = {}

    def desc(k):
        if k not in descs_cache:
            descs_cache[k] = pkg.synth.image(args.config, k, rows, pool)
        return descs_cache[k]

    for p in sample:
        i, j = (int(v) for v in my_pairs[p])
        raw = l1.pair_matches(desc(i), desc(j), rs)
        ok &= bool(np.array_equal(pm.pair(p), raw))
        ok &= bool(np.array_equal(cm.pair(p), l1.dedup_xy(raw, feats[i], feats[j])))
    n_bad = reduce(0.0 if ok else 1.0, dist.ReduceOp.SUM)
    log(f"oracle     : {len(sample)} sampled pairs per rank x {world} ranks, rows 7-12 and 7-13 vs CPU oracle in {time.time() - t0:.1f} s: "
        f"{'bit-exact' if n_bad == 0 else 'MISMATCH on %d rank(s)' % int(n_bad)}")
    log("CONFIG FULL OK" if n_bad == 0 else "CONFIG FULL FAILED")
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if n_bad == 0 else 1)


if __name__ == "__main__":
    main()
