"""Where one end-to-end step of bench.py goes (developer probe): host time of the streamed upload calls, of the match call
and of the final wait, with the images in upload-friendly order as bench.py sends them."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dreconstruction_b200")
synth = pkg.synth
rows, n_img = 10000, 100
descs0 = synth.collection(3, n_img, rows)
pd = [torch.empty((rows, 128), dtype=torch.uint8).pin_memory() for _ in range(n_img)]
pf = [torch.empty((rows, 2), dtype=torch.float32).pin_memory() for _ in range(n_img)]
for k in range(n_img):
    pd[k].numpy()[:] = descs0[k]; pf[k].numpy()[:] = synth.features(3, k, rows)[:, :2]
descs = [t.numpy() for t in pd]; feats = [t.numpy() for t in pf]
head = os.environ.get('HEAD')
pairs = pkg.upload_friendly_order(pkg.pairs_exhaustive(n_img), int(head) if head else None)
rs = float(pkg.square_f32(0.8))
ctx = pkg.Context(0)
m = pkg.MatcherCudaAllInMemory(0.8, ctx)
staged = m.StageArrays(descs, feats, order=list(range(n_img)))
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.time(); m.LoadStaged(staged, wait=False); t1 = time.time()
    pm = ctx.match_collection(pairs, rs, collect=False); t2 = time.time()
    ctx.stream_end(); t3 = time.time()
    print(f"[rep {rep}] LoadStaged(wait=False) {1e3*(t1-t0):.2f} ms | match_collection {1e3*(t2-t1):.2f} ms (gpu {pm.gpu_ms:.2f}, knn {pm.knn_kernel_ms:.2f}, {pm.knn_kernel_launches} batches) | "
          f"stream_end {1e3*(t3-t2):.2f} ms | total {1e3*(t3-t0):.2f} ms -> {len(pairs)/(t3-t0):.0f} pairs/s", flush=True)
