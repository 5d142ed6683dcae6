#!/bin/bash
# One GPU call that measures every prepared-but-unpromoted variant:
#   gpurun --timeout 900 -- 'bash scripts/ab_all.sh'
# Results land in gpurun_out/ (ab_pipeline.json, ab_nmi.json, *.log, bench_variant_*.json).
set -u
mkdir -p gpurun_out
echo "== kernel variants: parity + per-kernel timing"
timeout 120 python scripts/ab_pipeline.py > gpurun_out/ab_pipeline.log 2>&1; tail -30 gpurun_out/ab_pipeline.log | grep -v '^{'
echo "== NMI histogram variants"
timeout 120 python scripts/ab_nmi.py > gpurun_out/ab_nmi.log 2>&1; tail -8 gpurun_out/ab_nmi.log
echo "== band-local pyramid and sharded host I/O on 2 / 3 gloo ranks (one GPU)"
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -k "local_pyramid or host_io" -rxX > gpurun_out/local_pyramid.log 2>&1; tail -5 gpurun_out/local_pyramid.log
echo "== whole step with the candidate defaults"
for v in "0,0,0" "2,0,0" "2,4,0" "2,4,1"; do
    MA_FB_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "gpurun_out/bench_variant_${v//,/_}.json" 2> "gpurun_out/bench_variant_${v//,/_}.err"
    python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_variant_{v.replace(',', '_')}.json").read().strip().splitlines()[-1])
    print(v, "value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:  # noqa: BLE001
    print(v, "failed:", e)
PY
done
MA_NMI_VARIANT=1 MA_MINMAX_VARIANT=1 MA_FB_VARIANT=2,4,1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_variant_all.json 2> gpurun_out/bench_variant_all.err
tail -c 600 gpurun_out/bench_variant_all.json
