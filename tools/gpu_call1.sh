#!/bin/bash
# first GPU call of round 2: v2 conv kernel sanity (guarded against hangs), then the GPU suite, then short bench runs
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout -s KILL 180 python tools/conv_check.py > gpurun_out/c1_check_v2.log 2>&1; V2=$?
echo "v2 check rc=$V2"; tail -3 gpurun_out/c1_check_v2.log
HP_CONV_V1=1 timeout -s KILL 180 python tools/conv_check.py > gpurun_out/c1_check_v1.log 2>&1; echo "v1 check rc=$?"; tail -3 gpurun_out/c1_check_v1.log
if [ $V2 -ne 0 ]; then export HP_CONV_V1=1; echo "v2 failed: suite runs on the v1 conv kernel"; fi
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c1_pytest.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/c1_bench.json
