#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2c20_pytest.log 2>&1; tail -8 gpurun_out/r2c20_pytest.log
echo "== nmi timing"
python scripts/time_nmi.py > gpurun_out/r2c20_nmi.log 2>&1; tail -4 gpurun_out/r2c20_nmi.log
echo "== bench default (50000, with side configs)"
( time timeout 900 python bench.py > gpurun_out/r2c20_bench_50000.json 2> gpurun_out/r2c20_bench_50000.err ) 2>&1 | tail -3; grep "microaligner_b200:\|Error" gpurun_out/r2c20_bench_50000.err | head -5
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c20_bench_50000.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2), d['parity'], d['clocks'])
print(d['phases_ms']['rank0'])
for k,v in d['kernels'].items(): print('   ',k,v['ms_per_step'])
for k in ('configs1','configs2','cpu_baseline'):
    x=d[k]; print(k, round(x['value'],1), x.get('ms_per_step'), x.get('e2e',{}).get('value'), x.get('e2e',{}).get('ms_per_step'))
PY
