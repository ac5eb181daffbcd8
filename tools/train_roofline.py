"""Per-kernel roofline table of one tensor-path training step from an ncu launch list.

  python tools/train_roofline.py <launches.csv> <batch>  ->  markdown table

launches.csv: ncu --metrics gpu__time_duration.sum --clock-control none --csv of tools/dbg/train_timeline.py (TB=<batch>).
Each kernel of the LAST complete step is matched, in launch order, with its algorithmic work (SURVEY.md 8d conventions:
useful MACs x 2, compulsory bytes; padding never counts) and the bound it is measured against -- the dense bf16 tensor
peak for the GEMM-shaped kernels whose operands are large enough to be compute-bound, HBM bandwidth for the rest
(peaks from MEASURED_PEAKS.json).  ncu serialises launches and runs them cold, so these are per-kernel figures, not the
step time.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TENSOR = PK["hbm_gbs"], PK["bf16_tflops"]   # burst figures: kernels timed alone


def work(n):
    fc = 2.0 * 2304 * 2048 * n
    W = 2304 * 2048
    return [
        # (substring of the kernel name, label, flops, bytes)
        ("tc_conv_kernel", "conv1+pools+tanh, conv2+tanh+pool (emits p1, winners)", 2.0 * (1440000 + 2359296) * n, n * (16384 + 4608 + 14400 + 3600 + 2304)),
        ("tc_gemm_kernel", "fc1 forward + tanh", fc, W * 2 + n * (4608 + 4096)),
        ("tc_gemm_kernel", "fc2 forward (+ softmax at large batch)", fc, W * 2 + n * (4096 + 9216)),
        ("softmax_kernel|loss_from_y", "softmax/loss/softmax backward", 0, n * (9216 * 4 + 4608)),
        ("colsum", "fc2 bias gradient", 0, n * 9216),
        ("reduce_partials(", "  (second stage)", 0, 0),
        ("transpose_bf16_pair", "h1^T, dlogit^T", 0, n * (2048 + 2304) * 2 * 2),
        ("tc_gemm_kernel", "fc2 dW = h1^T x dlogit", fc, n * (2048 + 2304) * 2 + W * 4),
        ("tc_gemm_kernel", "fc2 dX x tanh'", fc, W * 2 + n * (4608 + 4096 + 8192 + 4096)),
        ("colsum", "fc1 bias gradient", 0, n * 8192),
        ("reduce_partials(", "  (second stage)", 0, 0),
        ("transpose_bf16_pair", "p2^T, da1^T", 0, n * (2304 + 2048) * 2 * 2),
        ("tc_gemm_kernel", "fc1 dW = p2^T x da1", fc, n * (2048 + 2304) * 2 + W * 4),
        ("tc_gemm_kernel", "fc1 dX x tanh'", fc, W * 2 + n * (4096 + 4608 + 9216 + 4608)),
        ("conv2_bwd_operands", "E, E^T, col^T (bf16) + per-crop dB", 0, n * (9216 + 2304 + 14400 + 18432 + 18432 + 73728)),
        ("reduce_partials_warp", "conv2 dB", 0, n * 256),
        ("tc_gemm_kernel", "conv2 dW^T = col^T x E (split-K)", 2.0 * 256 * 64 * 144 * n, n * (73728 + 18432)),
        ("reduce_c2w_t", "  split-K reduction", 0, 72 * 65536),
        ("tc_gemm_kernel", "dL/dcol = E x W2 (bf16 out)", 2.0 * 144 * 256 * 64 * n, n * (18432 + 73728)),
        ("col2im_g1_vec", "col2im + tanh' (conv1 stage)", 0, n * (73728 + 14400 + 14400)),
        ("conv1_wgrad", "conv1 dW, dB (winners only, FFMA)", 2.0 * 16 * 225 * 26 * n, n * (16384 + 14400 + 3600)),
        ("reduce_partials_warp", "  conv1 partial reduction", 0, n * 416 * 4),
    ]


def main():
    path, n = sys.argv[1], int(sys.argv[2])
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5 and r[0].isdigit()]
    names = [r[4] for r in rows]
    dur = [float(r[-1]) / 1e3 for r in rows]
    starts = [i for i, nm in enumerate(names) if "tc_conv_kernel" in nm]
    a, b = starts[-2], starts[-1]
    step = list(zip(names[a:b], dur[a:b]))
    out = ["| kernel | what | us | algorithmic | achieved | bound | fraction of peak |", "|---|---|---|---|---|---|---|"]
    i = 0
    tot = 0.0
    for pat, label, flops, byts in work(n):
        if label.startswith("  "):   # optional follow-up launch: only if it is the very next kernel
            if i >= len(step) or not any(p in step[i][0] for p in pat.split("|")):
                continue
        while i < len(step) and not any(p in step[i][0] for p in pat.split("|")):
            i += 1
        if i >= len(step):
            break
        nm, us = step[i]
        i += 1
        tot += us
        short = nm.split("(")[0].replace("void hp::", "").replace("hp::", "")
        tf = flops / (us * 1e-6) / 1e12 if flops else 0.0
        gb = byts / (us * 1e-6) / 1e9 if byts else 0.0
        # compute-bound if the FLOP/byte ratio exceeds the machine balance
        # (the conv kernel is paced by the tensor pipe's shared-memory operand port whatever its HBM traffic)
        tensor_bound = flops > 0 and byts > 0 and "tc_" in nm and (flops / byts > TENSOR * 1e12 / (HBM * 1e9) or "tc_conv_kernel" in nm)
        if tensor_bound:
            ach, bound, frac = "%.0f TFLOP/s" % tf, "tensor", tf / TENSOR
        elif byts:
            ach, bound, frac = "%.0f GB/s" % gb, "hbm", gb / HBM
        else:
            ach, bound, frac = "-", "-", 0.0
        alg = ("%.2f GFLOP, " % (flops / 1e9) if flops else "") + ("%.1f MB" % (byts / 1e6) if byts else "")
        out.append("| `%s` | %s | %.1f | %s | %s | %s | %s |" % (short, label, us, alg or "-", ach, bound, ("%.3f" % frac) if frac else "-"))
    upd = [(nm, us) for nm, us in step[i:]]
    if upd:
        us = sum(u for _, u in upd)
        byts = 3 * 37833600 + 2 * (18874368 + 2 * 9437184)
        out.append("| update tail (side streams) | 3 x `sgd_kernel` (113.5 MB) + bf16 shadow refresh | %.1f | %.1f MB | %.0f GB/s | hbm | %.3f |" % (us, byts / 1e6, byts / (us * 1e-6) / 1e9, byts / (us * 1e-6) / 1e9 / HBM))
    print("\n".join(out))
    print("\nsum of the main-stream kernels: %.1f us (batch %d); peaks: HBM %.0f GB/s, bf16 %.0f TFLOP/s" % (tot, n, HBM, TENSOR))


if __name__ == "__main__":
    main()
