#!/usr/bin/env bash
# One visit to an 8-GPU box: scaling lines (weak + strong) at N = 2, 4, 8, the reference arm at N = 8 (thread count check),
# the C++ driver's byte-identity across GPU counts and BASELINE config 4 file -> file on 8 GPUs.
# Usage (via gpurun --gpus 8):  bash scripts/gpu8_round.sh <tag>
set -u
TAG="${1:-dev8}"; shift || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpus.csv 2>&1
nproc > $OUT/${TAG}_nproc.txt
PORT=29500
for CASE in "8 weak" "8 strong" "4 weak" "2 weak"; do
  set -- $CASE; N=$1
  for SC in $2; do
    PORT=$((PORT + 1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N --steps 4 --warmup 3 --scaling $SC 2> $OUT/${TAG}_bench_${SC}_${N}gpu.err | tail -1 > $OUT/${TAG}_bench_${SC}_${N}gpu.json
    python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${SC}_${N}gpu.json"))
    print("N=$N $SC value %.0f e2e %.0f pairs/s checksum %s" % (d["value"], d["e2e"]["value"], d["checksum"]["pairs_hash_sum64"]))
except Exception as e:
    print("N=$N $SC: no line", e)
PY
  done
done
timeout 300 python scripts/driver_multi_gpu_check.py 8 2>&1 | tail -6 | tee $OUT/${TAG}_driver_multi_gpu.log
NIMG=${NIMG:-1000} ROWS=8000 GPUS=8 timeout 1200 python tests/tools/config4_files.py 2>&1 | tail -8 | tee $OUT/${TAG}_config4_files_8gpu.log
timeout 300 python tests/tools/startup_probe.py 2>&1 | tail -10 | tee $OUT/${TAG}_startup_probe.log
