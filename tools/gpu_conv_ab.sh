#!/bin/bash
# guarded A/B of a conv-kernel change: parity + stage times (a hang costs 150 s), role timeline, short bench
mkdir -p gpurun_out
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['eval_init_worst'], d['eval_peaky_worst'], d['grad9']['conv1.W'], d['timing'])"; }
timeout -s KILL 150 python tools/conv_check.py > gpurun_out/c6_pipe.log 2>&1; RC=$?
echo "== pipelined: rc=$RC"; summ < gpurun_out/c6_pipe.log
if [ $RC -ne 0 ]; then tail -5 gpurun_out/c6_pipe.log; exit 1; fi
echo "== HP_CONV_PIPE=0"; HP_CONV_PIPE=0 timeout -s KILL 120 python tools/conv_check.py 2>&1 | summ
HP_LIB_OVERRIDE=tools/dbg/_bin/libhandposedd_trace.so timeout -s KILL 120 python tools/dbg/conv2_trace.py 2>&1 | tail -12 | tee gpurun_out/conv2_trace.txt
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/c6_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c6_bench.json')); print(d['value'], d.get('parity_check'), [ (k['kernel'][:14], round(k['launch_ms']*1e3,1), round(k['frac'],3)) for k in d['roofline']['all_kernels']])"
