#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 tools/dbg/_bin/tmem_bench > gpurun_out/c3_tmem_bench.txt 2>&1; cat gpurun_out/c3_tmem_bench.txt
for v in "A=1" "HP_CONV_PIPE=0" "HP_CONV_TANH=accurate" "HP_CONV_PIPE=0 HP_CONV_TANH=accurate" "HP_CONV_V1=1" "HP_CONV_V1=1 HP_CONV_TANH=accurate"; do
  echo "== $v"; env $v timeout -s KILL 180 python tools/conv_check.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['eval_init_worst'], d['eval_peaky_worst'], d['timing'])"
done
timeout -s KILL 300 python -m pytest tests -m gpu -q --timeout 600 -k "dropin or nan_quirk or within_bound" 2>&1 | tail -3
timeout -s KILL 120 python tools/prof_conv.py > gpurun_out/c3_plain.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 2 -c 1 -o gpurun_out/prof_r2b python tools/prof_conv.py > gpurun_out/c3_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c3_ncu.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/c3_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c3_bench.json')); print(d['value'], d['parity_check'], [ (k['kernel'][:14], round(k['launch_ms']*1e3,1), round(k['frac'],3)) for k in d['roofline']['all_kernels']]); print(d.get('e2e')); print(d.get('device_u16_in_decoded_out')); print(d.get('train'))"
