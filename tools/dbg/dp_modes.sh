# usage: bash tools/dbg/dp_modes.sh N   -- training step time at N GPUs, peer-memory exchange vs NCCL baseline
N=${1:-2}
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dbg/dp_train_bench.py 2>&1 | grep "step timeline\|mode\|rror" ; }
for m in peer nccl; do HP_DP_MODE=$m run; HP_STEP_TIMING=1 HP_DP_MODE=$m run; done
