#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c4_pytest.log 2>&1; tail -8 gpurun_out/r2c4_pytest.log
echo "== nmi timing"
python scripts/time_nmi.py > gpurun_out/r2c4_nmi.log 2>&1; tail -4 gpurun_out/r2c4_nmi.log
echo "== bench 20000"
timeout 600 python bench.py --size 20000 > gpurun_out/r2c4_bench_20000.json 2> gpurun_out/r2c4_bench_20000.err; tail -c 200 gpurun_out/r2c4_bench_20000.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c4_bench_20000.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2), d['parity'])
for k,v in d['kernels'].items(): print('   ',k,v['ms_per_step'])
PY
echo "== ncu stages (full)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_tiles|nmi_chunk|zmip|norm_u8|warp_affine|compose_flows|dog_row|pyrdown|polyexp_march|merge_tiles|tile_max" -s 11 -c 13 -f -o gpurun_out/r2c4_prof_stages python scripts/prof_stages.py > gpurun_out/r2c4_ncu_stages.log 2>&1; tail -3 gpurun_out/r2c4_ncu_stages.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fb_" -c 6 -f -o gpurun_out/r2c4_prof_fb python scripts/prof_farneback.py 3000 > gpurun_out/r2c4_ncu_fb.log 2>&1; tail -2 gpurun_out/r2c4_ncu_fb.log
