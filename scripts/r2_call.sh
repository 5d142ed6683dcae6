#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 scripts/timeline.py 50000 > gpurun_out/r2c29_timeline_4gpu_50000.txt 2>&1
grep "^rank\|^device" gpurun_out/r2c29_timeline_4gpu_50000.txt
