#!/bin/bash
# usage: [HP_PEER_TMA=1] gpu_dpR.sh N r1 r2 ... -- N-GPU data-parallel step time for several exchange CTA counts
N=$1; shift
export HP_PEER_TIMEOUT_S=20
p=29800
for r in "$@"; do p=$((p+1)); HP_DP_RESERVE_SMS=$r timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $p tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|rror|Traceback"; done
