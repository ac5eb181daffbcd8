#!/bin/bash
# second GPU call: v2 vs v1 conv kernel with the ftz tanh, failed test re-run, ncu of the v2 conv kernel and the FC GEMMs
mkdir -p gpurun_out
timeout -s KILL 180 python tools/conv_check.py > gpurun_out/c2_check_v2.log 2>&1; echo "v2 check rc=$?"; tail -1 gpurun_out/c2_check_v2.log
HP_CONV_V1=1 timeout -s KILL 180 python tools/conv_check.py > gpurun_out/c2_check_v1.log 2>&1; echo "v1 check rc=$?"; tail -1 gpurun_out/c2_check_v1.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 -k "minibatch256 or tensor or dropin or loss_curves or depth" > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c2_pytest.log
timeout -s KILL 300 python tools/conv_check.py 16384 > gpurun_out/c2_plain.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"tc_conv2_kernel|tc_gemm_kernel" -s 6 -c 3 -o gpurun_out/prof_r2a python tools/conv_check.py 16384 > gpurun_out/c2_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c2_ncu.log
ls -la gpurun_out/*.ncu-rep
