#!/bin/bash
# training-step changes: guarded first step (a hang costs 120 s), gradient / training parity subset, device timeline of the
# replayed step with and without the change under test (A/B switch in $AB, e.g. AB="HP_GEMM_CLUSTER=0")
mkdir -p gpurun_out
timeout -s KILL 120 python tools/dbg/step_trace.py > gpurun_out/ab_trace.txt 2>&1; RC=$?
echo "== first step: rc=$RC"; tail -30 gpurun_out/ab_trace.txt
if [ $RC -ne 0 ]; then exit 1; fi
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "grad or train or minibatch or loss or winners or backward or shadows" 2>&1 | tail -5
if [ -n "$AB" ]; then env $AB timeout -s KILL 200 python tools/dbg/step_trace.py 2>&1 | grep "step ="; fi
TB=2048 timeout -s KILL 200 python tools/dbg/step_trace.py 2>&1 | grep "step ="
