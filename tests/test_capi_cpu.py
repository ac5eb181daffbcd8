"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/handposedd.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, _have_gpu
from hand_tracking_samples_b200 import capi, cnn


def header_symbols():
    text = open(os.path.join(ROOT, "include", "handposedd.h")).read()
    return re.findall(r"HP_API\s+[\w\s\*]+?\b(hp_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(syms) == sorted(capi.SYMBOLS)
    for s in syms:
        assert getattr(L, s) is not None
    assert b"sm_100a" in L.hp_version()


def test_library_contains_only_sm100a_code():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", capi._build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_unsupported_layer_list_is_rejected():
    L = capi.lib()
    layers = list(cnn.HANDPOSEDD_LAYERS)
    layers[7] = (4, (2304, 0, 0), (0, 0, 0, 0), (1024, 0, 0), None)  # a different LFull
    arr, keep = cnn._descs(layers)
    h = C.c_void_p()
    st = L.hp_create(arr, len(arr), 0, C.byref(h))
    assert st == 2 and not h.value
    assert b"handposedd" in L.hp_last_error()


@pytest.mark.skipif(_have_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(capi.HpError) as e:
        cnn.CNN()
    assert e.value.status == 6  # HP_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_null_arguments_are_errors_not_crashes():
    L = capi.lib()
    assert L.hp_init_xavier(None) == 1
    assert L.hp_eval_batch(None, None, 1, None, 0) == 1
    assert L.hp_destroy(None) == 0
    assert L.hp_launch_count(None) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hand_tracking_samples_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f
