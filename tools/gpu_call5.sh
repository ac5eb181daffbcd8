#!/bin/bash
# 1-GPU call: configs[4] sweep with the shipped kernels, then the role timeline of the conv kernel from the trace build
mkdir -p gpurun_out
timeout -s KILL 300 python tools/sweep.py --out gpurun_out/r2_sweep_1gpu.json 2> gpurun_out/r2_sweep_1gpu.err; echo "sweep rc=$?"
HP_LIB_OVERRIDE=tools/dbg/_bin/libhandposedd_trace.so timeout -s KILL 120 python tools/dbg/conv2_trace.py 2>&1 | tee gpurun_out/conv2_trace.txt
