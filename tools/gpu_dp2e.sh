#!/bin/bash
# 2 GPUs: data-parallel step time against the number of SMs reserved for the exchange kernels (graph-replayed step)
mkdir -p gpurun_out
export HP_PEER_TIMEOUT_S=20
run_tb() { timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|rror|Traceback" ; }
p=29700
for r in 16 24 32 48 64; do p=$((p+1)); HP_DP_RESERVE_SMS=$r HP_DP_MODE=peer run_tb $p; done
p=$((p+1)); HP_PEER_TMA=1 HP_DP_RESERVE_SMS=32 HP_DP_MODE=peer run_tb $p
