#!/bin/bash
# 2-GPU call: correctness of both exchange paths, step time peer-TMA vs peer-LDG vs NCCL, then the N=2 bench line
mkdir -p gpurun_out
export HP_PEER_TIMEOUT_S=20
run_tb() { timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|timeline|rror|Traceback" ; }
echo "== peer TMA"; HP_DP_MODE=peer run_tb 29601
echo "== peer TMA timeline"; HP_STEP_TIMING=1 HP_DP_MODE=peer run_tb 29602
echo "== peer LDG"; HP_PEER_TMA=0 HP_DP_MODE=peer run_tb 29603
echo "== nccl"; HP_DP_MODE=nccl run_tb 29604
timeout -s KILL 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 500 2>&1 | tail -5
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/dp2_bench.json 2> gpurun_out/dp2_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/dp2_bench.err; python -c "
import json; d=json.load(open('gpurun_out/dp2_bench.json')); print(d['value'], d['e2e']); print(d.get('e2e_depth_in_decoded_out')); print(json.dumps(d.get('train'))[:600]); print(json.dumps(d.get('train_scaling'), indent=1))"
