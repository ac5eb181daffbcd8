#!/usr/bin/env python
"""bench.py -- headline benchmark of the handposedd hot path on B200 (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)

Metric: handposedd 64x64 crops/sec (batched inference, BASELINE.json configs[1]: 65,536 synthetic
crops per GPU, tensor-core path), whole-job aggregate; plus, in the same JSON line, the FP32-exact
arm, the training samples/sec arm (configs[2]: minibatch 256 per GPU, data-parallel over NCCL when
N > 1), the roofline of the dominant kernel, the CPU baseline and the end-to-end (host buffers) number.

A "step" is one pass of CNN::Eval over one batch of 65,536 crops per GPU that is already resident in
HBM (`value`), or that starts in pinned host memory and ends in pinned host memory (`e2e`).
One process per GPU; under torchrun the ranks shard the crops with no data-path collective
(weak scaling) and the time is the max over ranks of CUDA-event time between barriers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CROP = 26472960          # 2 x 13,236,480 MAC
FLOP_PER_TRAIN_SAMPLE = 76538880
FC1_FLOP = 2 * 2304 * 2048        # per crop
FC2_FLOP = 2 * 2048 * 2304
CONV_FLOP = 2 * (1440000 + 2359296)
METRIC = "handposedd_64x64_inference_crops_per_sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region.  Uses NVML directly (about a
    millisecond per sample, so even a 20 ms timed region gets samples); falls back to polling
    nvidia-smi.  mark_begin()/mark_end() bracket the timed region; only samples inside count."""

    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.t_begin, self.t_end = None, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((time.perf_counter(), float(sm), float(mx), int(rs)))
                time.sleep(0.001)
        except Exception:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
            while not self.stop_flag:
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.rows.append((time.perf_counter(), float(out[0]), float(out[1]), int(out[2].strip(), 16)))
                except Exception:
                    pass

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def summary(self):
        self.stop_flag = True
        inside = [r for r in self.rows if self.t_begin is not None and self.t_begin <= r[0] <= (self.t_end or 1e30)]
        rows = inside or self.rows
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[3]
        reasons = [name for name, bit in self.REASONS if bits & bit]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": rows[0][2] if rows else None, "reasons": reasons,
                "samples": len(inside), "source": "nvml, sampled inside the timed region" if inside else "outside timed region"}


def cpu_baseline_eval(sample_crops, threads):
    """The reference's own CPU implementation (oracle/_ref, else the oracle port) on a bounded sample."""
    import numpy as np
    from hand_tracking_samples_b200 import synth
    from oracle import oracle as orc
    x = synth.uniform_crops(sample_crops, 1234)
    if orc.have_ref():
        r = orc.Ref(fast=True)
        r.init()
        r.eval(x[:threads], threads=threads)  # warm
        t0 = time.perf_counter()
        r.eval(x, threads=threads)
        dt = time.perf_counter() - t0
        kind = "reference"
    else:
        o = orc.Oracle()
        p = o.init_xavier()
        t0 = time.perf_counter()
        o.eval(p, x)
        dt = time.perf_counter() - t0
        kind, threads = "port", 1
    return sample_crops / dt, kind, threads, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(threads * 8, 64)
    vals = []
    for i in range(args.warmup + args.steps):
        v, kind, used, dt = cpu_baseline_eval(sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    tot = sum(sample for _ in vals) / sum(dt for _, dt in vals)
    line = {"impl": "reference", "metric": METRIC, "value": tot, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(dt for _, dt in vals) / len(vals), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "handposedd.cnnb-shaped Init() weights, CNN::Eval over %d synthetic 64x64 crops per step, "
                                   "one reference net per host thread" % sample},
            "cpu_baseline": {"value": tot, "unit": "crops/s", "cores": used, "kind": kind,
                             "sample": "%d uniform[0,1) crops per step, %d threads, unmodified third_party/cnn.h" % (sample, used)},
            "e2e": {"value": tot, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def emit(line):
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


# fd 1 is pointed at stderr for the whole run so that library chatter (NCCL's version banner, ...) cannot land next
# to the ONE JSON line, which is written to the saved real stdout
sys.stdout.flush()
REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="crops per GPU per step")
    ap.add_argument("--train-batch", type=int, default=256, help="samples per GPU per optimiser step")
    ap.add_argument("--no-extras", action="store_true", help="skip the FP32 arm, training arm, e2e and CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from hand_tracking_samples_b200 import cnn as hp
    from hand_tracking_samples_b200 import dp

    rank, world, local = dp.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    W, K, B = max(args.warmup, 3), args.steps, args.batch
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream

    net = hp.PoseInitializerCNN("", device=local)     # Init() weights: assets/handposedd.cnnb is absent from the mount
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand((B, 4096), device=dev, generator=gen)   # 1 GiB per GPU: larger than the 126 MB L2
    y = torch.empty((B, 2304), device=dev)

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- headline: tensor-core inference, inputs resident in HBM -------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    l0 = net.launch_count()
    net.profile(True)
    # (profiling only records events around the three stages; warm-up intervals are excluded below)
    for _ in range(W):
        net.eval_batch_device(x.data_ptr(), B, y.data_ptr(), precision=hp.PRECISION_TENSOR, stream=stream)
    barrier()
    net.profile(True)
    l0 = net.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(K):
        net.eval_batch_device(x.data_ptr(), B, y.data_ptr(), precision=hp.PRECISION_TENSOR, stream=stream)
    e1.record()
    barrier()
    sampler.mark_end()
    launches = net.launch_count() - l0
    stage_ms, stage_cnt = net.profile_read(3)
    net.profile(False)
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if distributed:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t.item())
    clocks = sampler.summary()
    value = world * B * K / (ms * 1e-3)

    pk = peaks()
    # per-kernel rooflines from the CUDA-event stage timings recorded inside the timed region (hp_profile)
    names = ["tc_conv_kernel (conv1+pool4+tanh, conv2+tanh+pool2)", "tc_gemm_kernel<fc1 + tanh>", "tc_gemm_kernel<fc2 + chunked softmax>"]
    flops = [CONV_FLOP, FC1_FLOP, FC2_FLOP]
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from ncu --set full
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    per_kernel = []
    for i in range(3):
        launch_ms = stage_ms[i] / max(stage_cnt[i], 1)
        crops_per_launch = B * K / max(stage_cnt[i], 1)
        ach = flops[i] * crops_per_launch / max(launch_ms * 1e-3, 1e-12) / 1e12
        per_kernel.append({"kernel": names[i], "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": ach / pk["bf16_tflops_sustained"], "launch_ms": launch_ms, "crops_per_launch": crops_per_launch,
                           "algorithmic_flop_per_crop": flops[i], "share_of_step": stage_ms[i] / max(sum(stage_ms), 1e-9),
                           "traffic": traffic.get(("tc_conv_kernel", "tc_gemm_kernel<0>", "tc_gemm_kernel<1>")[i])})
    dom = max(range(3), key=lambda i: stage_ms[i])
    roofline = dict(per_kernel[dom])
    roofline["peak_source"] = pk["source"] + ", sustained bf16 (kernels are timed inside a long step)"
    roofline["all_kernels"] = per_kernel
    roofline["whole_step"] = {"achieved": FLOP_PER_CROP * value / world / 1e12, "unit": "TFLOP/s",
                              "frac": FLOP_PER_CROP * value / world / 1e12 / pk["bf16_tflops_sustained"]}

    line = {"metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: handposedd batched inference, %d synthetic uniform[0,1) 64x64 crops per GPU "
                                   "per step, Init() weights (assets/handposedd.cnnb absent), tensor-core path" % B,
                       "crops_per_gpu": B, "parallelism": "replicated weights, batch sharded, no collective",
                       "l2_policy": "inputs (1 GiB/GPU) exceed the 126 MB L2"},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline}

    if not args.no_extras:
        # ---- FP32-exact arm (same workload, fewer steps: it is ~FFMA-bound) ----------------------
        k32 = max(1, min(K, 2))
        ms32 = timed(lambda: net.eval_batch_device(x.data_ptr(), B, y.data_ptr(), precision=hp.PRECISION_FP32, stream=stream), k32, 1)
        line["fp32_exact"] = {"value": world * B * k32 / (ms32 * 1e-3), "unit": "crops/s",
                              "tflops": FLOP_PER_CROP * B * k32 / (ms32 * 1e-3) / 1e12, "note": "FFMA path, parity 1e-5"}

        # ---- end to end through the host-buffer C ABI (pinned host -> device -> pinned host) ------
        xh = torch.empty((B, 4096), dtype=torch.float32).pin_memory()
        xh.copy_(x.cpu())
        yh = torch.empty((B, 2304), dtype=torch.float32).pin_memory()
        ke = max(1, min(K, 3))
        net.eval_batch(xh.numpy(), out=yh.numpy(), precision=hp.PRECISION_TENSOR)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            net.eval_batch(xh.numpy(), out=yh.numpy(), precision=hp.PRECISION_TENSOR)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if distributed:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": world * B * ke / float(dt.item()), "unit": "crops/s", "h2d_bytes_per_step": B * 4096 * 4,
                       "d2h_bytes_per_step": B * 2304 * 4, "api": "hp_eval_batch (host buffers, pinned; chunked H2D/compute/D2H overlap)"}
        # the PCIe ceiling this number lives under: a bare pinned H2D copy of the same 1 GiB
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xdst = torch.empty_like(x)
        xdst.copy_(xh, non_blocking=True)
        torch.cuda.synchronize()
        c0.record()
        xdst.copy_(xh, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = B * 4096 * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        line["e2e"]["pcie_h2d_gbs_measured"] = h2d_gbs
        line["e2e"]["frac_of_pcie_bound"] = (line["e2e"]["value"] / world) * 4096 * 4 / 1e9 / h2d_gbs
        del xdst
        # same call chain as handtrack.h:700-702 through the compact entry point: 16-bit depth crops up (8 KB/crop),
        # normalise + Eval + decode on the device, 48 decoded floats down (192 B/crop)
        dh = torch.randint(0, 1200, (B, 4096), dtype=torch.int32).to(torch.int16).pin_memory()
        dech = torch.empty((B, 48), dtype=torch.float32).pin_memory()
        dnp = dh.numpy().view(np.uint16)
        net.eval_depth_batch(dnp, precision=hp.PRECISION_TENSOR, want_y=False, out_dec=dech.numpy())
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            net.eval_depth_batch(dnp, precision=hp.PRECISION_TENSOR, want_y=False, out_dec=dech.numpy())
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if distributed:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        line["e2e_depth_in_decoded_out"] = {"value": world * B * ke / float(dt.item()), "unit": "crops/s", "h2d_bytes_per_step": B * 4096 * 2,
                                            "d2h_bytes_per_step": B * 48 * 4,
                                            "api": "hp_eval_depth_batch (u16 depth in, handtrack.h:700 normalisation + Eval + CNNOutputAnalysis decode on device)"}
        del xh, yh, dh, dech

        # ---- training arm: minibatch 256 per GPU, FP32 path, NCCL all-reduce when N > 1 ----------
        try:
            from hand_tracking_samples_b200 import synth
            TB = args.train_batch
            tx = torch.rand((TB, 4096), device=dev, generator=gen)
            tt = torch.from_numpy(synth.heatmap_labels(TB, 4321 + rank)).to(dev)
            mse = torch.empty(TB, device=dev)
            exchange = "none (1 GPU)"
            if distributed:
                try:
                    dp.init_data_parallel(net, mode="peer")
                    exchange = ("one kernel per gradient bucket (fc2 | fc1 | conv) over NVLink peer memory behind backward: reduce-scatter of the "
                                "9,458,400 fp32 gradient sums + SGD + all-gather of the updated weights (csrc/hp_peer.cu); bf16-shadow refresh "
                                "per bucket on a third stream")
                except Exception as e:
                    dp.init_data_parallel(net, mode="nccl")
                    exchange = "NCCL all-reduce + local SGD (peer-memory path unavailable: %s)" % str(e)[:120]
            kt = max(min(K, 200), 20)
            line["train"] = {}
            for name, prec in (("tensor", hp.PRECISION_TENSOR), ("fp32", hp.PRECISION_FP32)):
                mst = timed(lambda: net.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 0.001 / (TB * world), mse.data_ptr(),
                                                           precision=prec, stream=stream), kt, 3)
                line["train"][name] = {"value": world * TB * kt / (mst * 1e-3), "unit": "samples/s", "batch_per_gpu": TB, "steps": kt,
                                       "ms_per_step": mst / kt, "tflops": FLOP_PER_TRAIN_SAMPLE * TB * kt / (mst * 1e-3) / 1e12,
                                       "final_mse": float(mse.mean().item())}
            line["train"]["exchange"] = exchange
            # compute-dominated point: 2048 samples per GPU per step
            TL = 2048
            txl = torch.rand((TL, 4096), device=dev, generator=gen)
            ttl = tt.repeat(TL // TB, 1).contiguous()
            msel = torch.empty(TL, device=dev)
            ktl = max(kt // 4, 10)
            mst = timed(lambda: net.train_batch_device(txl.data_ptr(), ttl.data_ptr(), TL, 0.001 / (TL * world), msel.data_ptr(),
                                                       precision=hp.PRECISION_TENSOR, stream=stream), ktl, 3)
            line["train"]["tensor_batch2048"] = {"value": world * TL * ktl / (mst * 1e-3), "unit": "samples/s", "batch_per_gpu": TL,
                                                 "ms_per_step": mst / ktl, "tflops": FLOP_PER_TRAIN_SAMPLE * TL * ktl / (mst * 1e-3) / 1e12}
            del txl, ttl, msel
            if distributed:
                # comparison arm: the same step with NCCL all-reduce + local SGD kernel (fp32 wire, then opt-in bf16 wire)
                net_n = hp.PoseInitializerCNN("", device=local)
                dp.init_data_parallel(net_n, mode="nccl")
                for key, bf in (("tensor_nccl", False), ("tensor_nccl_bf16_wire", True)):
                    net_n.dp_set_bf16_gradients(bf)
                    mst = timed(lambda: net_n.train_batch_device(tx.data_ptr(), tt.data_ptr(), TB, 0.001 / (TB * world), mse.data_ptr(),
                                                                 precision=hp.PRECISION_TENSOR, stream=stream), kt, 3)
                    line["train"][key] = {"value": world * TB * kt / (mst * 1e-3), "unit": "samples/s", "ms_per_step": mst / kt,
                                          "note": "baseline exchange: NCCL all-reduce (%s) + local SGD kernel" % ("bf16 wire" if bf else "fp32 wire")}
                dp.shutdown_data_parallel(net_n)
                del net_n
            line["train"]["workload"] = "BASELINE.json configs[2]: forward+backward+SGD, minibatch %d synthetic crops per GPU" % TB
        except Exception as e:  # the training arm must not take the headline down with it
            line["train"] = {"error": str(e)[:200]}

        # ---- BASELINE.json configs[0]: one crop at a time through the drop-in call (the reference's own usage pattern) ----
        try:
            from hand_tracking_samples_b200 import synth as _synth
            x1, t1 = _synth.depthlike_crops(1, 3), _synth.heatmap_labels(1, 4)
            lat = {}
            for name, prec in (("fp32", hp.PRECISION_FP32), ("tensor", hp.PRECISION_TENSOR)):
                for _ in range(20):
                    net.eval_batch(x1, precision=prec)
                t0 = time.perf_counter()
                for _ in range(200):
                    net.eval_batch(x1, precision=prec)
                lat["eval_us_" + name] = (time.perf_counter() - t0) / 200 * 1e6
            line["single_crop"] = dict(lat, note="host-call latency of CNN::Eval on one crop, pageable host buffers, H2D+kernels+D2H+sync")
        except Exception as e:
            line["single_crop"] = {"error": str(e)[:200]}

        # ---- CPU baseline: the reference's own code on this box's host cores (rank 0, bounded sample) ----
        if rank == 0:
            threads = os.cpu_count() or 1
            v1, kind, _, dt1 = cpu_baseline_eval(64, 1)
            vN, kind, used, dtN = cpu_baseline_eval(max(64, 16 * threads), threads)
            line["cpu_baseline"] = {"value": vN, "unit": "crops/s", "cores": used, "kind": kind,
                                    "sample": "%d uniform[0,1) crops, %d threads (one reference net per thread); single-thread: %.1f crops/s on 64 crops"
                                              % (max(64, 16 * threads), used, v1),
                                    "single_thread": v1}
    if rank == 0:
        emit(line)
    if distributed:
        dp.shutdown_data_parallel(net)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
