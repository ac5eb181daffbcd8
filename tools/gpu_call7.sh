#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/sweep.py --out gpurun_out/r2_sweep_1gpu.json 2> gpurun_out/r2_sweep_1gpu.err; echo "sweep rc=$?"
HP_SWEEP_REPS_UNIT=4194304 timeout -s KILL 300 python tools/sweep.py --out gpurun_out/r2_sweep_1gpu_long.json 2> gpurun_out/r2_sweep_1gpu.err; echo "sweep rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_sweep_1gpu.json','gpurun_out/r2_sweep_1gpu_long.json'):
    d=json.load(open(f))
    print(f, [(r['crops'], round(r['crops_per_s']/1e6,2)) for r in d['rows'] if r['path']=='tensor' and r['crops']>=4096])
PY
