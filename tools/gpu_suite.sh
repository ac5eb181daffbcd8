#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_measured.jsonl
timeout -s KILL 150 python tools/conv_check.py > gpurun_out/s_check.log 2>&1 || { echo "conv_check failed"; tail -5 gpurun_out/s_check.log; exit 1; }
tail -1 gpurun_out/s_check.log | cut -c1-400
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
