"""In-tree nvcc build of libhandposedd.so (sm_100a only; no JIT cache, no fallback arch)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhandposedd.so")
SOURCES = ["hp_api.cu", "hp_fp32.cu", "hp_tc.cu", "hp_tc_conv.cu", "hp_tc_conv2.cu", "hp_post.cu", "hp_peer.cu", "hp_dataset.cu"]
HEADERS = ["hp_common.cuh", "hp_ptx.cuh", "hp_tc.cuh", "hp_peer.cuh", os.path.join("..", "..", "include", "handposedd.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"] + (["-DHP_CONV_TRACE"] if os.environ.get("HP_CONV_TRACE") else [])
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libhandposedd.so next to this file."""
    if not force and not _stale():
        return LIB
    objs = []
    log = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + log[-1])
        objs.append(obj)
    cmd = [nvcc()] + ARCH_FLAGS + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + log[-1])
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def build_trace():
    """Debug build for tools/dbg/conv2_trace.py: the same objects with hp_tc_conv2.cu recompiled under -DHP_CONV_TRACE, linked
    to tools/dbg/_bin/libhandposedd_trace.so (selected with HP_LIB_OVERRIDE).  Never the product library."""
    build()
    out_dir = os.path.join(os.path.dirname(HERE), "tools", "dbg", "_bin")
    os.makedirs(out_dir, exist_ok=True)
    obj = os.path.join(out_dir, "hp_tc_conv2_trace.o")
    subprocess.check_call([nvcc()] + NVCC_FLAGS + ["-DHP_CONV_TRACE", "-c", os.path.join(CSRC, "hp_tc_conv2.cu"), "-o", obj])
    objs = [obj if s == "hp_tc_conv2.cu" else os.path.join(CSRC, s.replace(".cu", ".o")) for s in SOURCES]
    lib = os.path.join(out_dir, "libhandposedd_trace.so")
    subprocess.check_call([nvcc()] + ARCH_FLAGS + ["-shared", "-o", lib] + objs + ["-cudart", "static", "-ldl", "-lpthread"])
    return lib


ROOT = os.path.dirname(HERE)
DROPIN_BIN = os.path.join(ROOT, "tests", "_bin", "dropin_main")


def build_dropin_test(reference_root="/root/reference"):
    """g++-compile the C++ drop-in checks against include/handposedd/cnn.h (no CUDA needed to build):
    tests/_bin/dropin_main (the PoseInitializerCNN call sequence, always), tests/_bin/ht_dropin (the
    reference's own include/handtrack.h compiled UNMODIFIED against the drop-in header) and tests/_bin/dataset_dropin
    (load_dataset rebuilt on hp_dataset_* beside the reference's own loader) -- the last two only where the
    reference tree is present.  The binaries travel to the GPU box; the reference tree does not."""
    os.makedirs(os.path.dirname(DROPIN_BIN), exist_ok=True)
    common = ["g++", "-std=c++14", "-O2", "-I" + os.path.join(ROOT, "include"), "-L" + HERE, "-lhandposedd",
              "-Wl,-rpath,$ORIGIN/../../hand_tracking_samples_b200"]
    subprocess.check_call(common[:4] + [os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), "-o", DROPIN_BIN] + common[4:])
    out = [DROPIN_BIN]
    if os.path.isdir(reference_root):
        ht = os.path.join(ROOT, "tests", "_bin", "ht_dropin")
        subprocess.check_call(common[:4] + ["-fpermissive", "-Wno-narrowing", "-w", "-I" + reference_root,
                                            os.path.join(ROOT, "tests", "cpp", "handtrack_dropin.cpp"), "-o", ht] + common[4:] + ["-lpthread"])
        out.append(ht)
        # the reference-side load_dataset binding of INTEGRATION.md next to the reference's own loader (host-only program)
        dsb = os.path.join(ROOT, "tests", "_bin", "dataset_dropin")
        subprocess.check_call(common[:4] + ["-fpermissive", "-Wno-narrowing", "-w", "-I" + reference_root,
                                            os.path.join(ROOT, "tests", "cpp", "dataset_dropin.cpp"), "-o", dsb] + common[4:])
        out.append(dsb)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
