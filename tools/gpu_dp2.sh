#!/bin/bash
# 2-GPU call: data-parallel step as a replayed CUDA graph (barrier epochs in device memory) -- correctness, step time with and
# without the graph, the N=2 bench line and the N=2 sweep
mkdir -p gpurun_out
export HP_PEER_TIMEOUT_S=20
timeout -s KILL 400 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 350 2>&1 | tail -5
run_tb() { timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|timeline|rror|Traceback" ; }
echo "== peer, graph"; HP_DP_MODE=peer run_tb 29601
echo "== peer, eager"; HP_NO_GRAPH=1 HP_DP_MODE=peer run_tb 29602
echo "== nccl"; HP_DP_MODE=nccl run_tb 29604
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/dp2_bench.json 2> gpurun_out/dp2_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/dp2_bench.err; python -c "
import json; s=open('gpurun_out/dp2_bench.json').read(); d=json.loads(s[s.index('{\"'):]); print(d['value'], d['e2e']); print(json.dumps(d.get('train_scaling'), indent=1))"
bash tools/gpu_sweepN.sh 2
