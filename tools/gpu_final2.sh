#!/bin/bash
# after the last source change: the whole GPU suite again, then one ncu --set full capture of the three inference kernels
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
timeout -s KILL 120 python tools/prof_conv.py > gpurun_out/r2_plain_prof.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"tc_conv2_kernel|tc_gemm_kernel" -s 6 -c 3 -o gpurun_out/prof_r2_final python tools/prof_conv.py > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_r2_final.ncu-rep
