#!/bin/bash
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 > gpurun_out/r2c27_bench_4gpu_50000.json 2> gpurun_out/r2c27_bench_4gpu_50000.err; grep "microaligner_b200:\|Error\|error" gpurun_out/r2c27_bench_4gpu_50000.err | head -5
python - <<'PY'
import json
for f in ['r2c27_bench_4gpu_50000']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2), d['parity'])
        print(' max', d['phases_ms']['max_over_ranks'])
    except Exception as e: print(f,'ERR',e)
PY
