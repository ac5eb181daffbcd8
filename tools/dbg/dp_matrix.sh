# usage: bash tools/dbg/dp_matrix.sh N -- peer-memory exchange: step time at N GPUs vs reserved SMs x exchange CTAs
N=${1:-2}
run() { timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dbg/dp_train_bench.py 2>&1 | grep "^mode\|rror" ; }
for r in 8 16; do for b in 4 8 16; do if [ $b -le $r ]; then echo "== reserve $r blocks $b"; HP_DP_MODE=peer HP_DP_RESERVE_SMS=$r HP_PEER_BLOCKS=$b run; fi; done; done
