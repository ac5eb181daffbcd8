# usage: bash tools/dbg/dp_matrix.sh N -- peer-memory exchange: training step time at N GPUs vs SMs reserved for the exchange CTAs
N=${1:-2}
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dbg/dp_train_bench.py 2>&1 | grep "step timeline\|mode\|rror" ; }
for r in 4 8 16 24; do echo "== reserve $r"; HP_DP_MODE=peer HP_DP_RESERVE_SMS=$r run | grep mode; done
HP_STEP_TIMING=1 HP_DP_MODE=peer run
