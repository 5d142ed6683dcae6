#!/bin/bash
echo "== bench 8 GPUs 50000"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 > gpurun_out/r2c21_bench_8gpu_50000.json 2> gpurun_out/r2c21_bench_8gpu_50000.err ) 2>&1 | tail -3; grep "microaligner_b200:\|Error\|error" gpurun_out/r2c21_bench_8gpu_50000.err | head
python - <<'PY'
import json
for f in ['r2c21_bench_8gpu_50000']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2), d['parity'], d['e2e']['h2d_bytes_per_step'])
        print(' max', d['phases_ms']['max_over_ranks']); print(' r0 ', d['phases_ms']['rank0'])
        for k,v in d['kernels'].items(): print('   ',k,v['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
