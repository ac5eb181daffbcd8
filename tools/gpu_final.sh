#!/bin/bash
# final 1-GPU pass: the whole GPU suite, smoke(), the bench line, the training-step timeline, the ncu launch list
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/final_bench.err
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['e2e']['value'], d.get('parity_check',{}).get('ok'), [(k['kernel'][:14], round(k['launch_ms']*1e3,1), round(k['frac'],3)) for k in d['roofline']['all_kernels']]); print(d.get('device_u16_in_decoded_out')); print(json.dumps(d.get('train'))[:900])"
timeout -s KILL 200 python tools/dbg/step_trace.py > gpurun_out/final_step_trace.txt 2>&1; grep "step =" gpurun_out/final_step_trace.txt
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_plain_noextras.json 2> gpurun_out/r2_plain_noextras.err && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
