#!/bin/bash
# usage: gpu_dpN.sh N [bench] -- N-GPU data-parallel step (graph-replayed, peer exchange), configs[4] sweep, optionally the bench line
N=$1
mkdir -p gpurun_out
export HP_PEER_TIMEOUT_S=20
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N tools/dbg/dp_train_bench.py 2>&1 | grep -E "^mode|rror|Traceback"
bash tools/gpu_sweepN.sh $N
if [ "$2" = "bench" ]; then
timeout -s KILL 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/dp${N}_bench.json 2> gpurun_out/dp${N}_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/dp${N}_bench.err; python -c "
import json; s=open('gpurun_out/dp${N}_bench.json').read(); d=json.loads(s[s.index('{\"'):]); print(d['value'], d['e2e']['value'], d.get('e2e_depth_in_decoded_out',{}).get('value')); print(json.dumps(d.get('train_scaling'))[:1500])"
fi
