#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scripts/diag_pin6.py 2>&1 | grep "rank\|microaligner_b200:"
echo "== bench 2 GPUs 50000"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 > gpurun_out/r2c16_bench_2gpu_50000.json 2> gpurun_out/r2c16_bench_2gpu_50000.err; grep "microaligner_b200:" gpurun_out/r2c16_bench_2gpu_50000.err | head
python - <<'PY'
import json
for f in ['r2c16_bench_2gpu_50000']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2), d['parity'], d['phases_ms']['max_over_ranks'])
    except Exception as e: print(f,'ERR',e)
PY
