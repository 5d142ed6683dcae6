#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pipeline bench, small"
timeout 600 python scripts/bench_pipeline.py --size 4000 --cycles 3 --zplanes 2 2>&1 | grep -v Warn | tail -3
echo "== pipeline bench C4, 1 GPU"
timeout 1200 python scripts/bench_pipeline.py > gpurun_out/r2c24_c4_1gpu.json 2> gpurun_out/r2c24_c4_1gpu.err; tail -c 1200 gpurun_out/r2c24_c4_1gpu.json; tail -3 gpurun_out/r2c24_c4_1gpu.err
