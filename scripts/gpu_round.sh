#!/usr/bin/env bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + one full capture of the dominant kernel.
# Usage (from the repo root, via gpurun):  bash scripts/gpu_round.sh <tag> [tests|bench|ncu|all]
set -u
TAG="${1:-dev}"
WHAT="${2:-all}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/${TAG}_gpu.csv 2>&1
if [[ "$WHAT" == "all" || "$WHAT" == "tests" ]]; then
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
fi
if [[ "$WHAT" == "all" || "$WHAT" == "bench" ]]; then
  timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 | tee $OUT/${TAG}_bench.json
fi
if [[ "$WHAT" == "all" || "$WHAT" == "ncu" ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench_stdout.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2 -s 2 -c 1 -f -o $OUT/${TAG}_knn2 \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full_stdout.log 2>&1
  ls -la $OUT | tail -12
fi
