#!/bin/bash
# the one 8-GPU call of the round: exchange-kernel A/B at 8 ranks, configs[4] sweeps at 8 and 4 GPUs, the N=8 bench line
mkdir -p gpurun_out
export HP_PEER_TIMEOUT_S=20
run_tb() { timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|timeline|rror|Traceback" ; }
echo "== peer LDG"; HP_DP_MODE=peer run_tb 29801
echo "== peer TMA"; HP_PEER_TMA=1 HP_DP_MODE=peer run_tb 29802
echo "== peer LDG timeline"; HP_STEP_TIMING=1 HP_DP_MODE=peer run_tb 29803
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 tools/sweep.py > gpurun_out/r2_sweep_8gpu.json 2> gpurun_out/r2_sweep_8gpu.err; echo "sweep8 rc=$?"
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29812 tools/sweep.py > gpurun_out/r2_sweep_4gpu.json 2> gpurun_out/r2_sweep_4gpu.err; echo "sweep4 rc=$?"
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29813 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/r2_bench_8gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_8gpu.json')); print(d['value'], d['e2e']['value'], d['e2e_depth_in_decoded_out']['value']); print(json.dumps(d.get('train_scaling')))"
