#!/bin/bash
echo "== bench 2 GPUs 50000"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 > gpurun_out/r2c17_bench_2gpu_50000.json 2> gpurun_out/r2c17_bench_2gpu_50000.err; grep "microaligner_b200:\|Error\|error" gpurun_out/r2c17_bench_2gpu_50000.err | head
python - <<'PY'
import json
for f in ['r2c17_bench_2gpu_50000']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2), d['parity'], d['phases_ms']['max_over_ranks'])
    except Exception as e: print(f,'ERR',e)
PY
echo "== multirank numpy API tests (gloo ranks on one GPU)"
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "numpy_api" 2>&1 | tail -3
