"""End-to-end (pinned host -> device -> pinned host) crops/s of hp_eval_batch vs the pipeline chunk size (HP_PIPE_CHUNK);
one process per setting because the library reads the variable once."""
import os, subprocess, sys
if len(sys.argv) > 1:
    import time, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from hand_tracking_samples_b200 import cnn as hp
    net = hp.PoseInitializerCNN("")
    B = 65536
    xh = torch.rand((B, 4096)).pin_memory(); yh = torch.empty((B, 2304)).pin_memory()
    net.eval_batch(xh.numpy(), out=yh.numpy(), precision=hp.PRECISION_TENSOR)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        net.eval_batch(xh.numpy(), out=yh.numpy(), precision=hp.PRECISION_TENSOR)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print("HP_PIPE_CHUNK=%s: %.2f ms per 65,536 crops, %.3f M crops/s" % (os.environ.get("HP_PIPE_CHUNK", "default"), dt * 1e3, B / dt / 1e6), flush=True)
else:
    for c in ("1024", "2048", "3072", "4096", "8192"):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, HP_PIPE_CHUNK=c))
