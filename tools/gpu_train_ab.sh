#!/bin/bash
# training-step changes: gradient / training parity subset, then the device timeline of the replayed step
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "grad or train or minibatch or loss or winners or backward" 2>&1 | tail -5
timeout -s KILL 200 python tools/dbg/step_trace.py 2>&1 | tail -30
HP_NO_PRIORITY=1 timeout -s KILL 200 python tools/dbg/step_trace.py 2>&1 | grep "step ="
TB=2048 timeout -s KILL 200 python tools/dbg/step_trace.py 2>&1 | grep "step ="
