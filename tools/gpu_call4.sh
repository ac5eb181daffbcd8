#!/bin/bash
# guarded: the first step decides whether the pipelined conv kernel is usable; a hang there costs 150 s, not the call
mkdir -p gpurun_out
summ() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['eval_init_worst'], d['eval_peaky_worst'], d['grad9']['conv1.W'], d['timing'])"; }
timeout -s KILL 150 python tools/conv_check.py > gpurun_out/c4_pipe.log 2>&1; PIPE_RC=$?
echo "== pipelined v2: rc=$PIPE_RC"; summ < gpurun_out/c4_pipe.log
if [ $PIPE_RC -ne 0 ]; then export HP_CONV_V1=1; echo "V2 KERNEL UNUSABLE -> HP_CONV_V1=1 for the rest"; tail -5 gpurun_out/c4_pipe.log; fi
if [ $PIPE_RC -eq 0 ]; then
for v in "HP_CONV_PIPE=0" "HP_CONV_V1=1"; do
  echo "== $v"; env $v timeout -s KILL 120 python tools/conv_check.py 2>&1 | summ
done
fi
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 300 -x -k "dropin or nan_quirk or within_bound or resample or depth_upload or minibatch256 or pool_winners" 2>&1 | tail -4
timeout -s KILL 120 python tools/prof_conv.py > gpurun_out/c4_plain.log 2>&1 && \
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:tc_conv -s 2 -c 1 -o gpurun_out/prof_r2b python tools/prof_conv.py > gpurun_out/c4_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c4_ncu.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/c4_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c4_bench.json')); print(d['value'], d['parity_check'], [ (k['kernel'][:14], round(k['launch_ms']*1e3,1), round(k['frac'],3)) for k in d['roofline']['all_kernels']]); print(d.get('e2e')); print(d.get('device_u16_in_decoded_out')); print(d.get('train'))"
