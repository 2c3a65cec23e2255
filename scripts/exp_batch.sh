#!/usr/bin/env bash
# Developer batch: kernel throughput of every variant build under build/exp/ (libmvgcuda_<name>.so) by image size, one
# GPU-box visit.   Usage: bash scripts/exp_batch.sh <tag> [rows ...]
set -u
TAG="${1:-exp}"; shift || true
ROWS_LIST="${*:-10000}"
OUT=gpurun_out/${TAG}_variants.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/${TAG}_gpu.csv 2>&1
: > $OUT
for lib in build/exp/libmvgcuda_*.so; do
  for rows in $ROWS_LIST; do
    nimg=40; [[ $rows -le 4000 ]] && nimg=100; [[ $rows -ge 30000 ]] && nimg=10
    MVGCUDA_LIB=$lib ROWS=$rows NIMG=$nimg REPS=4 timeout 120 python scripts/exp_perf.py 2>&1 | tail -2 | tee -a $OUT
  done
done
