"""BASELINE.json configs[4]: inference throughput sweep, batch 1 .. 1M synthetic crops, sharded evenly over the ranks
(contiguous slices, replicated weights, no collective -- SURVEY.md 8e), device-resident inputs, with the roofline
fraction of every point.

  python tools/sweep.py > profiles/r2_sweep_1gpu.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep.py --out profiles/r2_sweep_Ngpu.json

Per point: time of one Eval pass over the global batch = max over ranks of the CUDA-event time between barriers
(mean of `reps` back-to-back passes after 3 warm-up passes; the input slice of a rank is reused across passes, so points
whose slice fits the 126 MB L2 -- below ~8k crops per GPU -- run L2-warm, which is what a small batch looks like in
practice).  Roofline: useful FLOPs (26,472,960 per crop) / time / (ranks x measured burst bf16 peak); for the small
end also the weight-stream floor (37.8 MB of fp32 master... the 18.9 MB of 16-bit shadows per rank at HBM speed).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from hand_tracking_samples_b200 import cnn as hp, dp  # noqa: E402

FLOP = 26472960
REPS_UNIT = int(os.environ.get("HP_SWEEP_REPS_UNIT", 1 << 20))   # crops of tensor-path work per timed point and rank


def main():
    rank, world, local = dp.env_rank_world()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
        {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    net = hp.PoseInitializerCNN("", device=local)
    st = torch.cuda.current_stream().cuda_stream
    nmax = 1 << 20
    per_rank_max = (nmax + world - 1) // world
    x = torch.rand((per_rank_max, 4096), device="cuda")
    y = torch.empty((per_rank_max, 2304), device="cuda")
    rows = []

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for prec, name in ((hp.PRECISION_TENSOR, "tensor"), (hp.PRECISION_FP32, "fp32")):
        for k in range(0, 21):
            n = 1 << k
            if name == "fp32" and n > (1 << 17):
                break
            lo, hi = dp.shard_range(n, rank, world)
            m = hi - lo
            # ~35 ms of work per point (bench.py's 20 x 65,536-crop steps): longer back-to-back runs are power-capped on this
            # pool (1-GPU sweep of 63 passes at 65,536 crops: 33.4 M crops/s against 38-39 M in bench.py on the same box)
            reps = max(3, min(200, REPS_UNIT // max((n // world + 1) * (1 if name == "tensor" else 32), 1)))

            def one():
                if m > 0:
                    net.eval_batch_device(x.data_ptr(), m, y.data_ptr(), precision=prec, stream=st)

            for _ in range(3):
                one()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                one()
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            cps = n / (ms * 1e-3)
            tf = cps * FLOP / 1e12
            rows.append({"path": name, "crops": n, "crops_per_gpu": (n + world - 1) // world, "ms": ms, "crops_per_s": cps, "tflops": tf,
                         "frac_of_burst_bf16_peak": tf / (world * pk["bf16_tflops"]) if name == "tensor" else None,
                         "frac_of_fp32_ffma_nominal": tf / (world * 74.0) if name == "fp32" else None,
                         "weight_stream_floor_us": 18.9e6 / (pk["hbm_gbs"] * 1e9) * 1e6 if name == "tensor" else 37.8e6 / (pk["hbm_gbs"] * 1e9) * 1e6})
    if rank == 0:
        out = open(sys.argv[sys.argv.index("--out") + 1], "w") if "--out" in sys.argv else sys.stdout   # NCCL prints its version on stdout
        print(json.dumps({"gpu": torch.cuda.get_device_name(0), "n_gpus": world, "burst_bf16_tflops_per_gpu": pk["bf16_tflops"], "hbm_gbs": pk["hbm_gbs"],
                          "note": "global batch sharded evenly, no collective; time = max over ranks", "reps_unit": REPS_UNIT, "rows": rows}, indent=1), file=out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
