#!/bin/bash
N=${NPROC:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N > gpurun_out/r2c32_bench_${N}gpu_50000.json 2> gpurun_out/r2c32_bench_${N}gpu_50000.err; grep "microaligner_b200:\|Error\|error" gpurun_out/r2c32_bench_${N}gpu_50000.err | head -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c32_bench_${N}gpu_50000.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2), d['parity'])
print(' replicated', d.get('replicated_result',{}).get('ms_per_step'))
print(' max', d['phases_ms']['max_over_ranks'])
PY
