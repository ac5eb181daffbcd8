# what slows the backward kernels that run beside an exchange?  0 = normal, 1 = barriers only, 2 = same traffic aimed at local memory
N=${1:-2}
for m in 0 1 2; do echo "== HP_PEER_DEBUG=$m"; HP_PEER_DEBUG=$m HP_DP_MODE=peer timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dbg/dp_train_bench.py 2>&1 | grep "^mode\|rror"; done
HP_PEER_DEBUG=2 HP_STEP_TIMING=1 HP_DP_MODE=peer timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dbg/dp_train_bench.py 2>&1 | grep "timeline"
