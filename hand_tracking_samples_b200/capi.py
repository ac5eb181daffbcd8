"""ctypes binding of include/handposedd.h (libhandposedd.so, built in-tree by _build.py).

This is the only route from Python into the product: there is no PyTorch or NumPy
implementation of the path behind it, so a missing library or a missing GPU is a hard error.
"""
import ctypes as C
import os

from . import _build

HP_OK = 0
PRECISION_FP32 = 0
PRECISION_TENSOR = 1
N_IN = 4096
N_OUT = 2304
N_PARAMS = 9458400
CNNB_BYTES = 37833600
PEER_HANDLE_BYTES = 256

# every symbol include/handposedd.h declares
SYMBOLS = [
    "hp_create", "hp_create_handposedd", "hp_retain", "hp_destroy", "hp_init_xavier", "hp_load_cnnb",
    "hp_save_cnnb", "hp_get_params_range", "hp_set_params_range", "hp_load_cnnb_file", "hp_save_cnnb_file", "hp_eval_batch", "hp_eval_batch_device",
    "hp_decode_batch", "hp_decode_batch_device", "hp_eval_decode_batch", "hp_eval_depth_batch", "hp_eval_depth_batch_device", "hp_normalize_depth_device", "hp_resample_depth_device", "hp_eval_frames_device",
    "hp_render_labels", "hp_render_labels_device", "hp_train_batch_points",
    "hp_train_batch", "hp_train_batch_device", "hp_grad_batch_device", "hp_get_grads", "hp_device_ptrs",
    "hp_apply_grads_device", "hp_dp_unique_id", "hp_dp_init", "hp_dp_set_bf16_gradients", "hp_dp_peer_export", "hp_dp_peer_init", "hp_dp_peer_status", "hp_dp_shutdown", "hp_launch_count", "hp_debug_step_times", "hp_profile", "hp_profile_read",
    "hp_peek", "hp_last_error", "hp_version",
    "hp_dataset_open", "hp_dataset_get_info", "hp_dataset_read", "hp_dataset_eval_depth", "hp_dataset_close",
]


class LayerDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("in_dims", C.c_int * 3), ("w_dims", C.c_int * 4),
                ("out_dims", C.c_int * 3), ("n_spans", C.c_int), ("spans", C.POINTER(C.c_int))]


class DatasetInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("focal", C.c_float * 2), ("principal", C.c_float * 2),
                ("depth_scale", C.c_float), ("mplane", C.c_float * 4), ("hasir", C.c_int32), ("rgb_dim", C.c_int32 * 2),
                ("feye_dim", C.c_int32 * 2), ("segment_scale", C.c_float), ("camtype", C.c_char * 32), ("n_frames", C.c_int64),
                ("pose_array_size", C.c_int32), ("has_ir_file", C.c_int32), ("has_pose_file", C.c_int32)]


class HpError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("handposedd status %d: %s" % (status, msg))
        self.status = status


_lib = None


def lib():
    """Load (building if stale and nvcc is present) libhandposedd.so.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("HP_LIB_OVERRIDE"):   # debug builds (tools/dbg): load exactly this file
        path = os.environ["HP_LIB_OVERRIDE"]
    elif not os.path.exists(path) or _build._stale():
        if os.path.exists(_build.nvcc()):
            _build.build()
    if not os.path.exists(path):
        raise FileNotFoundError(path + ": the CUDA library is not built (python -m hand_tracking_samples_b200._build)")
    L = C.CDLL(path)
    vp, i64, fp = C.c_void_p, C.c_int64, C.c_float
    L.hp_create.argtypes = [C.POINTER(LayerDesc), C.c_int, C.c_int, C.POINTER(vp)]
    L.hp_create_handposedd.argtypes = [C.c_int, C.POINTER(vp)]
    L.hp_retain.argtypes = [vp]
    L.hp_destroy.argtypes = [vp]
    L.hp_init_xavier.argtypes = [vp]
    L.hp_load_cnnb.argtypes = [vp, vp, C.c_size_t]
    L.hp_save_cnnb.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.hp_get_params_range.argtypes = [vp, i64, i64, vp]
    L.hp_set_params_range.argtypes = [vp, i64, i64, vp]
    L.hp_load_cnnb_file.argtypes = [vp, C.c_char_p]
    L.hp_save_cnnb_file.argtypes = [vp, C.c_char_p]
    L.hp_eval_batch.argtypes = [vp, vp, i64, vp, C.c_int]
    L.hp_eval_batch_device.argtypes = [vp, vp, i64, vp, C.c_int, vp]
    L.hp_decode_batch.argtypes = [vp, vp, i64, vp]
    L.hp_decode_batch_device.argtypes = [vp, vp, i64, vp, vp]
    L.hp_eval_decode_batch.argtypes = [vp, vp, i64, vp, vp, C.c_int]
    L.hp_eval_depth_batch.argtypes = [vp, vp, i64, fp, fp, fp, vp, vp, C.c_int]
    L.hp_eval_depth_batch_device.argtypes = [vp, vp, i64, fp, fp, fp, vp, vp, C.c_int, vp]
    L.hp_resample_depth_device.argtypes = [vp, vp, C.c_int32, C.c_int32, C.POINTER(fp), vp, vp, i64, C.c_uint16, vp, vp]
    L.hp_eval_frames_device.argtypes = [vp, vp, C.c_int32, C.c_int32, C.POINTER(fp), vp, vp, i64, C.c_uint16, fp, fp, fp, vp, vp, C.c_int, vp]
    L.hp_normalize_depth_device.argtypes = [vp, vp, i64, fp, fp, fp, vp, vp]
    L.hp_render_labels.argtypes = [vp, vp, vp, i64, vp]
    L.hp_render_labels_device.argtypes = [vp, vp, vp, i64, vp, vp]
    L.hp_train_batch_points.argtypes = [vp, vp, vp, vp, i64, fp, vp, C.c_int]
    L.hp_train_batch.argtypes = [vp, vp, vp, i64, fp, vp, C.c_int]
    L.hp_train_batch_device.argtypes = [vp, vp, vp, i64, fp, vp, C.c_int, vp]
    L.hp_grad_batch_device.argtypes = [vp, vp, vp, i64, vp, C.c_int, vp]
    L.hp_get_grads.argtypes = [vp, vp]
    L.hp_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.hp_apply_grads_device.argtypes = [vp, fp, vp]
    L.hp_dp_unique_id.argtypes = [vp]
    L.hp_dp_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.hp_dp_set_bf16_gradients.argtypes = [vp, C.c_int]
    L.hp_dp_peer_export.argtypes = [vp, vp]
    L.hp_dp_peer_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.hp_dp_peer_status.argtypes = [vp, C.POINTER(C.c_int)]
    L.hp_dp_shutdown.argtypes = [vp]
    L.hp_launch_count.argtypes = [vp]
    L.hp_launch_count.restype = i64
    L.hp_debug_step_times.argtypes = [vp, vp]
    L.hp_profile.argtypes = [vp, C.c_int]
    L.hp_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(i64)]
    L.hp_peek.argtypes = [vp, C.c_int, i64, vp]
    L.hp_dataset_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.hp_dataset_get_info.argtypes = [vp, C.POINTER(DatasetInfo)]
    L.hp_dataset_read.argtypes = [vp, i64, i64, vp, vp, vp]
    L.hp_dataset_eval_depth.argtypes = [vp, vp, i64, i64, fp, fp, vp, vp, C.c_int]
    L.hp_dataset_close.argtypes = [vp]
    L.hp_dataset_close.restype = None
    L.hp_last_error.restype = C.c_char_p
    L.hp_version.restype = C.c_char_p
    _lib = L
    return L


def check(status):
    if status != HP_OK:
        raise HpError(status, lib().hp_last_error().decode())
