#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 150 python tools/conv_check.py > gpurun_out/s_check.log 2>&1 || { echo "conv_check failed"; tail -5 gpurun_out/s_check.log; exit 1; }
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "decode or depth or resample" 2>&1 | tail -4
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s_bench.json')); print(d['value'], d['device_u16_in_decoded_out']); print(d['e2e_depth_in_decoded_out'])"
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_plain_noextras.json 2> gpurun_out/r2_plain_noextras.err && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout -s KILL 120 python tools/prof_conv.py > gpurun_out/r2_plain_prof.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"tc_conv2_kernel|tc_gemm_kernel" -s 6 -c 3 -o gpurun_out/prof_r2_full python tools/prof_conv.py > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
