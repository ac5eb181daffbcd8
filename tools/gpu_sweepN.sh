#!/bin/bash
# usage: gpu_sweepN.sh N  -- BASELINE configs[4] sweep on N GPUs of this box
N=$1
mkdir -p gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 298$N tools/sweep.py --out gpurun_out/r2_sweep_${N}gpu.json 2> gpurun_out/r2_sweep_${N}gpu.err; echo "sweep$N rc=$?"; tail -2 gpurun_out/r2_sweep_${N}gpu.err
