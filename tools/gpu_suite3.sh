#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout -s KILL 150 python tools/conv_check.py > gpurun_out/s_check.log 2>&1 || { echo "conv_check failed"; tail -5 gpurun_out/s_check.log; exit 1; }
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s_pytest.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/s_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s_bench.json')); print(d['value'], d['train']); print(d['device_u16_in_decoded_out']); print(d['e2e_depth_in_decoded_out'])"
