"""Argument validation of the host-only entry points (no GPU needed): every bad call returns a status and a message
instead of crashing, as include/handposedd.h promises."""
import ctypes as C

import pytest

from hand_tracking_samples_b200 import capi


def test_dataset_entry_points_reject_bad_arguments(tmp_path):
    L = capi.lib()
    h = C.c_void_p()
    assert L.hp_dataset_open(None, 17, C.byref(h)) == 1                       # HP_ERR_INVALID
    assert L.hp_dataset_open(b"x", -1, C.byref(h)) == 1
    assert L.hp_dataset_open(str(tmp_path / "missing").encode(), 17, C.byref(h)) == 4   # HP_ERR_IO
    assert b".rs" in L.hp_last_error()
    info = capi.DatasetInfo()
    assert L.hp_dataset_get_info(None, C.byref(info)) == 1
    assert L.hp_dataset_read(None, 0, 1, None, None, None) == 1
    assert L.hp_dataset_eval_depth(None, None, 0, 1, 0.1, 0.7, None, None, 0) == 1
    L.hp_dataset_close(None)                                                   # a no-op, like free(NULL)


def test_peer_entry_points_reject_bad_arguments():
    L = capi.lib()
    buf = (C.c_char * capi.PEER_HANDLE_BYTES)()
    assert L.hp_dp_peer_export(None, buf) == 1
    assert L.hp_dp_peer_init(None, buf, 0, 2) == 1
    v = C.c_int(0)
    assert L.hp_dp_peer_status(None, C.byref(v)) == 1
    assert L.hp_dp_shutdown(None) == 0


def test_json_header_corner_cases(tmp_path):
    """dcamera.dims must be positive (the reference would loop forever on zero-sized frames); nested / escaped strings parse."""
    import json
    import numpy as np
    from hand_tracking_samples_b200.dataset import Dataset
    base = str(tmp_path / "d")
    np.zeros(12, np.uint16).tofile(base + ".rs")
    json.dump({"dcamera": {"dims": [0, 3]}}, open(base + ".json", "w"))
    with pytest.raises(capi.HpError):
        Dataset(base, 1)
    open(base + ".json", "w").write("not json")
    with pytest.raises(capi.HpError):
        Dataset(base, 1)
    open(base + ".json", "w").write('{"camtype": "a\\"b\\\\c", "dcamera": {"dims": [4, 3], "extra": {"deep": [1, [2, 3]]}}, "hasir": false}')
    ds = Dataset(base, 1)
    assert len(ds) == 1 and ds.info.camtype == b'a"b\\c'
