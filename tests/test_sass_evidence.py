"""The shipped library really is Blackwell-native code (B200_PROFILING.md, "What proves a Blackwell-native kernel"): the
contraction kernels contain tcgen05 MMAs (UTCHMMA), TMEM loads (LDTM) and TMA / bulk copies (UTMALDG, UBLKCP), no legacy
mma.sync (HMMA) and only sm_100a code.  Needs cuobjdump (CUDA toolkit), no GPU."""
import os
import re
import shutil
import subprocess

import pytest

from hand_tracking_samples_b200 import _build

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


def sass_by_kernel():
    out = subprocess.run([CUOBJDUMP, "-sass", _build.LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and name:
            kernels[name].append(m.group(1))
    return out, kernels


def test_contraction_kernels_use_tcgen05_tmem_and_tma():
    out, kernels = sass_by_kernel()
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs
    gemm = [k for k in kernels if "tc_gemm_kernel" in k]
    conv = [k for k in kernels if "tc_conv_kernel" in k]
    assert len(gemm) >= 8 and len(conv) == 2          # epilogue x tile-width instantiations; inference + training conv kernel
    for k in gemm:
        ops = set(kernels[k])
        assert {"UTCHMMA", "LDTM", "UTMALDG"} <= ops, (k, sorted(ops & {"UTCHMMA", "LDTM", "UTMALDG", "UBLKCP"}))
    for k in conv:
        ops = set(kernels[k])
        assert {"UTCHMMA", "LDTM", "UBLKCP"} <= ops, k   # weights by cp.async.bulk; the A operands are views of smem images
    every = {op for ops in kernels.values() for op in ops}
    assert "HMMA" not in every and "HGMMA" not in every    # no mma.sync / wgmma anywhere in the library


def test_exchange_kernels_are_system_scope_and_present():
    _, kernels = sass_by_kernel()
    peer = [k for k in kernels if "peer_sgd_kernel" in k or "peer_small_kernel" in k]
    assert len(peer) == 14                               # world sizes 2..8, fc-bucket and conv-bucket kernels
    text = subprocess.run([CUOBJDUMP, "-sass", "-fun", peer[0], _build.LIB], capture_output=True, text=True).stdout
    assert ".SYS" in text                                # system-scope loads / stores / fences of the flag barrier
