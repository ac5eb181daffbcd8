"""Python mirror of the reference's ``CNN`` class API for the handposedd path.

Same member names and argument meaning as third_party/cnn.h:100-605 of
IntelRealSense/hand_tracking_samples -- ``Eval``, ``Train``, ``Init``, ``loadb``,
``saveb`` -- plus the batched entry points this framework adds.  Every call goes
through the C ABI (include/handposedd.h) into hand-written sm_100a kernels; there is
no NumPy/PyTorch implementation behind it.

``PoseInitializerCNN(filename)`` mirrors include/handtrack.h:103-130: build the
11-layer net, ``Init()``, then ``loadb`` if the file opens (silently keeping the random
weights otherwise).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import N_IN, N_OUT, N_PARAMS, PRECISION_FP32, PRECISION_TENSOR  # noqa: F401

# the layer list of include/handtrack.h:108-118, as (kind, in_dims, w_dims, out_dims, spans)
HANDPOSEDD_LAYERS = [
    (1, (64, 64, 1), (5, 5, 1, 16), (60, 60, 16), None),
    (2, (60 * 60 * 16, 0, 0), (0, 0, 0, 0), (0, 0, 0), None),
    (3, (60, 60, 16), (0, 0, 0, 0), (0, 0, 0), None),
    (3, (30, 30, 16), (0, 0, 0, 0), (0, 0, 0), None),
    (1, (15, 15, 16), (4, 4, 16, 64), (12, 12, 64), None),
    (2, (12 * 12 * 64, 0, 0), (0, 0, 0, 0), (0, 0, 0), None),
    (3, (12, 12, 64), (0, 0, 0, 0), (0, 0, 0), None),
    (4, (6 * 6 * 64, 0, 0), (0, 0, 0, 0), (16 * 16 * 8, 0, 0), None),
    (2, (16 * 16 * 8, 0, 0), (0, 0, 0, 0), (0, 0, 0), None),
    (4, (16 * 16 * 8, 0, 0), (0, 0, 0, 0), (16 * 16 * 8 + 16 * 16, 0, 0), None),
    (5, (0, 0, 0), (0, 0, 0, 0), (0, 0, 0), [256] * 8 + [16] * 16),
]


def _descs(layers):
    arr = (capi.LayerDesc * len(layers))()
    keep = []
    for d, (kind, i, w, o, spans) in zip(arr, layers):
        d.kind = kind
        d.in_dims[:] = i
        d.w_dims[:] = w
        d.out_dims[:] = o
        if spans:
            sp = (C.c_int * len(spans))(*spans)
            keep.append(sp)
            d.n_spans = len(spans)
            d.spans = C.cast(sp, C.POINTER(C.c_int))
    return arr, keep


def _f32(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a.reshape(-1, cols)


class CNN:
    """Device-resident handposedd net.  Copies made with ``copy()`` share one weight store
    (the reference's CNN is a bag of raw layer pointers copied shallowly, handtrack.h:129)."""

    def __init__(self, layers=None, device=0, precision=PRECISION_FP32, _handle=None):
        self.L = capi.lib()
        self.precision = precision
        if _handle is not None:
            self.h = _handle
            return
        h = C.c_void_p()
        arr, keep = _descs(HANDPOSEDD_LAYERS if layers is None else layers)
        capi.check(self.L.hp_create(arr, len(arr), device, C.byref(h)))
        self.h = h

    def copy(self):
        capi.check(self.L.hp_retain(self.h))
        return CNN(precision=self.precision, _handle=self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.hp_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- reference API -------------------------------------------------------------
    def Init(self):
        """CNN::Init, cnn.h:581."""
        capi.check(self.L.hp_init_xavier(self.h))

    def Eval(self, x):
        """CNN::Eval, cnn.h:550: one 4096-float crop -> 2304 floats."""
        return self.eval_batch(np.asarray(x, np.float32).reshape(1, N_IN))[0]

    def Train(self, x, t, alpha=0.01):
        """CNN::Train, cnn.h:558: one SGD step on one sample; returns the MSE."""
        return float(self.train_batch(np.asarray(x, np.float32).reshape(1, N_IN),
                                      np.asarray(t, np.float32).reshape(1, N_OUT), alpha)[0])

    def loadb(self, src):
        """CNN::loadb (cnn.h:590,592): a path (silent no-op if it cannot be opened, like the
        reference), bytes, or a binary stream."""
        if isinstance(src, str):
            st = self.L.hp_load_cnnb_file(self.h, src.encode())
            if st not in (0, 4):
                capi.check(st)
            return
        data = src if isinstance(src, (bytes, bytearray, memoryview)) else src.read(capi.CNNB_BYTES)
        buf = (C.c_char * len(data)).from_buffer_copy(bytes(data))
        capi.check(self.L.hp_load_cnnb(self.h, buf, len(data)))

    def saveb(self, dst=None):
        """CNN::saveb (cnn.h:591,593): to a path, a binary stream, or returned as bytes."""
        if isinstance(dst, str):
            capi.check(self.L.hp_save_cnnb_file(self.h, dst.encode()))
            return None
        buf = (C.c_char * capi.CNNB_BYTES)()
        n = C.c_size_t()
        capi.check(self.L.hp_save_cnnb(self.h, buf, capi.CNNB_BYTES, C.byref(n)))
        data = bytes(buf[:n.value])
        if dst is None:
            return data
        dst.write(data)
        return None

    # ---- batched entry points (new) ------------------------------------------------
    def eval_batch(self, x, precision=None, out=None):
        x = _f32(x, N_IN)
        n = x.shape[0]
        y = out if out is not None else np.empty((n, N_OUT), np.float32)
        capi.check(self.L.hp_eval_batch(self.h, x.ctypes.data, n, y.ctypes.data,
                                        self.precision if precision is None else precision))
        return y

    def decode_batch(self, y):
        """CNNOutputAnalysis numeric core (handtrack.h:218-241): y[n][2304] -> [n][48]."""
        y = _f32(y, N_OUT)
        out = np.empty((y.shape[0], 48), np.float32)
        capi.check(self.L.hp_decode_batch(self.h, y.ctypes.data, y.shape[0], out.ctypes.data))
        return out

    def eval_decode_batch(self, x, precision=None, want_y=True):
        x = _f32(x, N_IN)
        n = x.shape[0]
        y = np.empty((n, N_OUT), np.float32) if want_y else None
        dec = np.empty((n, 48), np.float32)
        capi.check(self.L.hp_eval_decode_batch(self.h, x.ctypes.data, n, y.ctypes.data if want_y else None, dec.ctypes.data,
                                               self.precision if precision is None else precision))
        return (y, dec) if want_y else dec

    def eval_depth_batch(self, depth_u16, depth_scale=0.001, dmin=0.1, dmax=0.7, precision=None, want_y=True, want_decoded=True,
                         out_y=None, out_dec=None):
        """handtrack.h:700-702: 16-bit depth crops -> normalise -> Eval -> (decode)."""
        d = np.ascontiguousarray(depth_u16, np.uint16).reshape(-1, N_IN)
        n = d.shape[0]
        y = out_y if out_y is not None else (np.empty((n, N_OUT), np.float32) if want_y else None)
        dec = out_dec if out_dec is not None else (np.empty((n, 48), np.float32) if want_decoded else None)
        capi.check(self.L.hp_eval_depth_batch(self.h, d.ctypes.data, n, depth_scale, dmin, dmax, y.ctypes.data if y is not None else None,
                                              dec.ctypes.data if dec is not None else None,
                                              self.precision if precision is None else precision))
        return y, dec

    def render_labels(self, points, vals):
        """GatherHandExpectedCNN label vector (handtrack.h:160-173): points[n][8][2], vals[n][16] -> [n][2304]."""
        p = _f32(points, 16)
        v = _f32(vals, 16)
        t = np.empty((p.shape[0], N_OUT), np.float32)
        capi.check(self.L.hp_render_labels(self.h, p.ctypes.data, v.ctypes.data, p.shape[0], t.ctypes.data))
        return t

    def train_batch_points(self, x, points, vals, alpha, precision=None):
        x, p, v = _f32(x, N_IN), _f32(points, 16), _f32(vals, 16)
        mse = np.empty(x.shape[0], np.float32)
        capi.check(self.L.hp_train_batch_points(self.h, x.ctypes.data, p.ctypes.data, v.ctypes.data, x.shape[0], alpha, mse.ctypes.data,
                                                self.precision if precision is None else precision))
        return mse

    def train_batch(self, x, t, alpha, precision=None):
        x = _f32(x, N_IN)
        t = _f32(t, N_OUT)
        n = x.shape[0]
        mse = np.empty(n, np.float32)
        capi.check(self.L.hp_train_batch(self.h, x.ctypes.data, t.ctypes.data, n, alpha, mse.ctypes.data,
                                         self.precision if precision is None else precision))
        return mse

    # device-pointer variants: x/y/t are integer device addresses (e.g. torch tensor .data_ptr())
    def eval_batch_device(self, x_ptr, n, y_ptr, precision=None, stream=0):
        capi.check(self.L.hp_eval_batch_device(self.h, x_ptr, n, y_ptr,
                                               self.precision if precision is None else precision, stream))

    def eval_depth_batch_device(self, depth_ptr, n, y_ptr, dec_ptr=None, depth_scale=0.001, dmin=0.1, dmax=0.7, precision=None, stream=0):
        """handtrack.h:700-702 on device buffers: uint16 depth[n][4096] -> y[n][2304] (+ decoded[n][48])."""
        capi.check(self.L.hp_eval_depth_batch_device(self.h, depth_ptr, n, depth_scale, dmin, dmax, y_ptr, dec_ptr,
                                                     self.precision if precision is None else precision, stream))

    def resample_depth_device(self, frames_ptr, width, height, src_intr, cams_ptr, n, crops_ptr, frame_of_crop_ptr=None, background=4000, stream=0):
        """SampleD (misc_image.h:154-162): full frames -> 64x64 uint16 crops for host-computed destination cameras."""
        intr = (C.c_float * 4)(*[float(v) for v in src_intr])
        capi.check(self.L.hp_resample_depth_device(self.h, frames_ptr, width, height, intr, frame_of_crop_ptr, cams_ptr, n, background, crops_ptr, stream))

    def eval_frames_device(self, frames_ptr, width, height, src_intr, cams_ptr, n, y_ptr, dec_ptr=None, frame_of_crop_ptr=None, background=4000,
                           depth_scale=0.001, dmin=0.1, dmax=0.7, precision=None, stream=0):
        """handtrack.h:698-702 on the device: resample -> normalise -> Eval -> (decode)."""
        intr = (C.c_float * 4)(*[float(v) for v in src_intr])
        capi.check(self.L.hp_eval_frames_device(self.h, frames_ptr, width, height, intr, frame_of_crop_ptr, cams_ptr, n, background, depth_scale, dmin,
                                                dmax, y_ptr, dec_ptr, self.precision if precision is None else precision, stream))

    def train_batch_device(self, x_ptr, t_ptr, n, alpha, mse_ptr=None, precision=None, stream=0):
        capi.check(self.L.hp_train_batch_device(self.h, x_ptr, t_ptr, n, alpha, mse_ptr,
                                                self.precision if precision is None else precision, stream))

    def grad_batch_device(self, x_ptr, t_ptr, n, mse_ptr=None, precision=None, stream=0):
        capi.check(self.L.hp_grad_batch_device(self.h, x_ptr, t_ptr, n, mse_ptr,
                                               self.precision if precision is None else precision, stream))

    def apply_grads_device(self, alpha, stream=0):
        capi.check(self.L.hp_apply_grads_device(self.h, alpha, stream))

    def get_grads(self):
        g = np.empty(N_PARAMS, np.float32)
        capi.check(self.L.hp_get_grads(self.h, g.ctypes.data))
        return g

    def get_params(self):
        return np.frombuffer(self.saveb(), np.float32).copy()

    def set_params(self, p):
        p = np.ascontiguousarray(p, np.float32)
        capi.check(self.L.hp_load_cnnb(self.h, p.ctypes.data, p.nbytes))

    def device_ptrs(self):
        p, g = C.c_void_p(), C.c_void_p()
        capi.check(self.L.hp_device_ptrs(self.h, C.byref(p), C.byref(g)))
        return p.value, g.value

    def peek(self, which, n, length):
        out = np.empty((n, length), np.float32)
        capi.check(self.L.hp_peek(self.h, which, n, out.ctypes.data))
        return out

    def profile(self, enable):
        capi.check(self.L.hp_profile(self.h, int(enable)))

    def profile_read(self, n_stages=4):
        ms = (C.c_double * n_stages)()
        cnt = (C.c_int64 * n_stages)()
        capi.check(self.L.hp_profile_read(self.h, n_stages, ms, cnt))
        return list(ms), list(cnt)

    def launch_count(self):
        return int(self.L.hp_launch_count(self.h))

    # ---- data parallelism ----------------------------------------------------------
    @staticmethod
    def dp_unique_id():
        buf = (C.c_char * 128)()
        capi.check(capi.lib().hp_dp_unique_id(buf))
        return bytes(buf)

    def dp_init(self, unique_id, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        capi.check(self.L.hp_dp_init(self.h, buf, rank, world))

    def dp_set_bf16_gradients(self, enable=True):
        capi.check(self.L.hp_dp_set_bf16_gradients(self.h, int(enable)))

    def dp_peer_export(self):
        buf = (C.c_char * capi.PEER_HANDLE_BYTES)()
        capi.check(self.L.hp_dp_peer_export(self.h, buf))
        return bytes(buf)

    def dp_peer_init(self, all_handles, rank, world):
        assert len(all_handles) == world * capi.PEER_HANDLE_BYTES
        buf = (C.c_char * len(all_handles)).from_buffer_copy(all_handles)
        capi.check(self.L.hp_dp_peer_init(self.h, buf, rank, world))

    def dp_peer_status(self):
        v = C.c_int(0)
        capi.check(self.L.hp_dp_peer_status(self.h, C.byref(v)))
        return v.value

    def dp_shutdown(self):
        capi.check(self.L.hp_dp_shutdown(self.h))


def PoseInitializerCNN(filename="", device=0, precision=PRECISION_FP32):
    """include/handtrack.h:103-130."""
    cnn = CNN(device=device, precision=precision)
    cnn.Init()
    if filename:
        cnn.loadb(filename)
    return cnn
