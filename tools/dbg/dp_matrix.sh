run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dbg/dp_train_bench.py 2>&1 | grep "step timeline\|mode" ; }
for r in 0 8 16; do for b in 16 48 128; do echo "== reserve $r blocks $b"; HP_STEP_TIMING=1 HP_DP_MODE=peer HP_DP_RESERVE_SMS=$r HP_PEER_BLOCKS=$b run; done; done
echo "== nccl timing"; HP_STEP_TIMING=1 HP_DP_MODE=nccl run
