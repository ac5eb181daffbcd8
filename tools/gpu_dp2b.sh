#!/bin/bash
mkdir -p gpurun_out
export HP_PEER_TIMEOUT_S=20
run_tb() { timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|timeline|rror|Traceback" ; }
echo "== peer TMA"; HP_DP_MODE=peer run_tb 29601
echo "== peer TMA timeline"; HP_STEP_TIMING=1 HP_DP_MODE=peer run_tb 29602
echo "== peer LDG"; HP_PEER_TMA=0 HP_DP_MODE=peer run_tb 29603
echo "== peer LDG timeline"; HP_STEP_TIMING=1 HP_PEER_TMA=0 HP_DP_MODE=peer run_tb 29605
timeout -s KILL 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 500 2>&1 | tail -3
