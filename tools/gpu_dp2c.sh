#!/bin/bash
export HP_PEER_TIMEOUT_S=20
run_tb() { timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|rror|Traceback" ; }
p=29700
for tma in 1 0; do for r in 24 32 48; do p=$((p+1)); echo "== TMA=$tma reserve=$r"; HP_PEER_TMA=$tma HP_DP_RESERVE_SMS=$r HP_DP_MODE=peer run_tb $p; done; done
