"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package (hand_tracking_samples_b200) never does.

* ``Oracle``  -> oracle/_build/liboracle.so : plain-C restatement of the reference's
  cnn.h for the handposedd architecture (oracle/handposedd_oracle.c).
* ``Ref``     -> oracle/_ref/libcnnref.so   : the UNMODIFIED reference cnn.h compiled
  in place from /root/reference behind a C ABI (oracle/ref_shim.cpp).  Prebuilt here,
  travels to the GPU box as a binary; /root/reference itself is never read at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
N_PARAMS = 9458400
N_IN = 4096
N_OUT = 2304

# .cnnb float offsets: name -> (offset, count)
LAYOUT = {
    "conv1.W": (0, 400), "conv1.B": (400, 16),
    "conv2.W": (416, 16384), "conv2.B": (16800, 64),
    "fc1.W": (16864, 4718592), "fc1.B": (4735456, 2048),
    "fc2.W": (4737504, 4718592), "fc2.B": (9456096, 2304),
}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (gcc/g++).  `ref` needs /root/reference and is skipped without it."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(os.environ.get("REFERENCE_ROOT", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libcnnref.so"))


def have_postref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libpostref.so"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    """Plain-C restatement.  Parameters live in a flat float32 array in .cnnb order."""

    def __init__(self):
        path = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.orc_ws_create.restype = C.c_void_p
        L.orc_ws_destroy.argtypes = [C.c_void_p]
        L.orc_eval.argtypes = [_f32p, _f32p, C.c_long, _f32p, C.c_void_p]
        L.orc_train_seq.argtypes = [_f32p, _f32p, _f32p, C.c_long, C.c_float, _f32p, C.c_void_p]
        L.orc_grad_sample.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_void_p]
        L.orc_grad_sample.restype = C.c_float
        L.orc_train_minibatch.argtypes = [_f32p, _f32p, _f32p, C.c_long, C.c_float, _f64p, _f32p, C.c_int, C.c_void_p]
        L.orc_peek.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.orc_init_xavier.argtypes = [_f32p]
        L.orc_decode.argtypes = [_f32p, C.c_long, _f32p]
        L.orc_normalize_depth.argtypes = [np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS"), C.c_long, C.c_float, C.c_float,
                                          C.c_float, _f32p]
        L.orc_render_labels.argtypes = [_f32p, _f32p, C.c_long, _f32p]
        L.orc_sample_d.argtypes = [_u16p, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int, C.c_int] + [C.c_float] * 4 + [_f32p, C.c_ushort, _u16p]
        self.L = L
        self.ws = C.c_void_p(L.orc_ws_create())

    def __del__(self):
        try:
            self.L.orc_ws_destroy(self.ws)
        except Exception:
            pass

    def init_xavier(self):
        p = np.zeros(N_PARAMS, np.float32)
        self.L.orc_init_xavier(p)
        return p

    def eval(self, params, x):
        x = _f32(x).reshape(-1, N_IN)
        y = np.empty((x.shape[0], N_OUT), np.float32)
        self.L.orc_eval(_f32(params), x, x.shape[0], y, self.ws)
        return y

    def train_seq(self, params, x, t, alpha):
        """n sequential batch-1 reference Train steps; params updated IN PLACE."""
        assert params.dtype == np.float32 and params.flags.c_contiguous
        x = _f32(x).reshape(-1, N_IN)
        t = _f32(t).reshape(-1, N_OUT)
        mse = np.empty(x.shape[0], np.float32)
        self.L.orc_train_seq(params, x, t, x.shape[0], alpha, mse, self.ws)
        return mse

    def grad_sample(self, params, x, t):
        g = np.empty(N_PARAMS, np.float32)
        mse = self.L.orc_grad_sample(_f32(params), _f32(x).reshape(N_IN), _f32(t).reshape(N_OUT), g, self.ws)
        return g, mse

    def train_minibatch(self, params, x, t, alpha, apply=True):
        """Batched-entry-point semantics: W -= alpha * sum_b g_b.  Returns (grad_sum f64, mse[n])."""
        assert params.dtype == np.float32 and params.flags.c_contiguous
        x = _f32(x).reshape(-1, N_IN)
        t = _f32(t).reshape(-1, N_OUT)
        g = np.empty(N_PARAMS, np.float64)
        mse = np.empty(x.shape[0], np.float32)
        self.L.orc_train_minibatch(params, x, t, x.shape[0], alpha, g, mse, int(apply), self.ws)
        return g, mse

    def decode(self, y):
        """CNNOutputAnalysis numeric core: y[n][2304] -> [n][48]."""
        y = _f32(y).reshape(-1, N_OUT)
        out = np.empty((y.shape[0], 48), np.float32)
        self.L.orc_decode(y, y.shape[0], out)
        return out

    def render_labels(self, points, vals):
        """GatherHandExpectedCNN label vector: points[n][8][2], vals[n][16] -> [n][2304]."""
        p = _f32(points).reshape(-1, 16)
        v = _f32(vals).reshape(-1, 16)
        t = np.empty((p.shape[0], N_OUT), np.float32)
        self.L.orc_render_labels(p, v, p.shape[0], t)
        return t

    def normalize_depth(self, d, depth_scale=0.001, dmin=0.1, dmax=0.7):
        d = np.ascontiguousarray(d, np.uint16)
        out = np.empty(d.shape, np.float32)
        self.L.orc_normalize_depth(d.reshape(-1), d.size, depth_scale, dmin, dmax, out.reshape(-1))
        return out

    def sample_d(self, frame, src_intr, dst_cam, background, dst_dim=(64, 64)):
        """SampleD (misc_image.h:154-162): frame[h][w] u16, src_intr = (fx, fy, px, py),
        dst_cam = (fx, fy, px, py, pos xyz, quat xyzw) -> [dh][dw] u16."""
        frame = np.ascontiguousarray(frame, np.uint16)
        h, w = frame.shape
        cam = _f32(dst_cam).reshape(11)
        out = np.empty((dst_dim[1], dst_dim[0]), np.uint16)
        self.L.orc_sample_d(frame, w, h, *[float(v) for v in src_intr], dst_dim[0], dst_dim[1], *[float(v) for v in cam[:4]],
                            np.ascontiguousarray(cam[4:]), int(background), out)
        return out

    _PEEK = {0: 57600, 1: 57600, 3: 3600, 5: 9216, 6: 2304, 8: 2048, 9: 2304, 10: 2304,
             109: 2304, 107: 2048, 106: 2304, 104: 9216, 103: 3600, 100: 57600}

    def peek(self, which):
        out = np.empty(self._PEEK[which], np.float32)
        n = self.L.orc_peek(self.ws, which, out)
        assert n == out.size
        return out


def eval_mt(params, x, threads=None):
    """Oracle.eval over a thread pool (one Oracle workspace per worker; ctypes releases the GIL): the 1,024-crop
    slice checks of the full-size tests and of bench.py's parity_check."""
    from concurrent.futures import ThreadPoolExecutor
    x = _f32(x).reshape(-1, N_IN)
    params = _f32(params)
    threads = max(1, min(threads or (os.cpu_count() or 1), x.shape[0]))
    bounds = [x.shape[0] * i // threads for i in range(threads + 1)]

    def work(i):
        return Oracle().eval(params, x[bounds[i]:bounds[i + 1]])

    with ThreadPoolExecutor(threads) as ex:
        return np.concatenate(list(ex.map(work, range(threads))))


def grad_minibatch_mt(params, x, t, threads=None):
    """sum_b g_b (float64) and the per-sample MSE of a minibatch at frozen weights -- the quantity
    Oracle.train_minibatch(apply=False) returns -- computed sample-parallel over a thread pool."""
    from concurrent.futures import ThreadPoolExecutor
    x = _f32(x).reshape(-1, N_IN)
    t = _f32(t).reshape(-1, N_OUT)
    params = _f32(params)
    n = x.shape[0]
    threads = max(1, min(threads or (os.cpu_count() or 1), n))
    bounds = [n * i // threads for i in range(threads + 1)]

    def work(i):
        o = Oracle()
        g = np.zeros(N_PARAMS, np.float64)
        mse = np.empty(bounds[i + 1] - bounds[i], np.float32)
        for k, b in enumerate(range(bounds[i], bounds[i + 1])):
            gb, mse[k] = o.grad_sample(params, x[b], t[b])
            g += gb
        return g, mse

    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(work, range(threads)))
    return sum(p[0] for p in parts), np.concatenate([p[1] for p in parts])


class PostRef:
    """The reference's decode / crop-normalisation routines (oracle/_ref/libpostref.so, ref_post_shim.cpp)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libpostref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        L.ref_decode.argtypes = [_f32p, C.c_long, _f32p]
        L.ref_normalize_depth.argtypes = [np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS"), C.c_long, C.c_float, C.c_float,
                                          C.c_float, _f32p]
        L.ref_render_labels.argtypes = [_f32p, _f32p, C.c_long, _f32p]
        if hasattr(L, "ref_sample_d"):
            L.ref_sample_d.argtypes = [_u16p, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int, C.c_int] + [C.c_float] * 4 + [_f32p, C.c_ushort, _u16p]
        self.L = L

    def render_labels(self, points, vals):
        p = _f32(points).reshape(-1, 16)
        v = _f32(vals).reshape(-1, 16)
        t = np.empty((p.shape[0], N_OUT), np.float32)
        self.L.ref_render_labels(p, v, p.shape[0], t)
        return t

    def sample_d(self, frame, src_intr, dst_cam, background, dst_dim=(64, 64)):
        frame = np.ascontiguousarray(frame, np.uint16)
        h, w = frame.shape
        cam = _f32(dst_cam).reshape(11)
        out = np.empty((dst_dim[1], dst_dim[0]), np.uint16)
        self.L.ref_sample_d(frame, w, h, *[float(v) for v in src_intr], dst_dim[0], dst_dim[1], *[float(v) for v in cam[:4]],
                            np.ascontiguousarray(cam[4:]), int(background), out)
        return out

    def decode(self, y):
        y = _f32(y).reshape(-1, N_OUT)
        out = np.empty((y.shape[0], 48), np.float32)
        self.L.ref_decode(y, y.shape[0], out)
        return out

    def normalize_depth(self, d, depth_scale=0.001, dmin=0.1, dmax=0.7):
        d = np.ascontiguousarray(d, np.uint16)
        out = np.empty(d.shape, np.float32)
        self.L.ref_normalize_depth(d.reshape(-1), d.size, depth_scale, dmin, dmax, out.reshape(-1))
        return out


def have_datasetref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libdatasetref.so"))


class DatasetRef:
    """The reference's load_dataset / DepthDataStreamOut (oracle/_ref/libdatasetref.so, ref_dataset_shim.cpp)."""

    def __init__(self):
        L = C.CDLL(os.path.join(HERE, "_ref", "libdatasetref.so"))
        L.ref_dataset_load.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.ref_dataset_load.restype = C.c_long
        L.ref_dataset_frame.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_dataset_frame.restype = None
        L.ref_dataset_save.argtypes = [C.c_char_p, C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L = L

    def load(self, basename, pose_array_size):
        """-> None when the reference throws, else (info[17], depth [n][h][w], ir [n][h][w], poses [n][np][7])."""
        info = np.zeros(17, np.float32)
        n = self.L.ref_dataset_load(str(basename).encode(), pose_array_size, info.ctypes.data)
        if n < 0:
            return None
        w, h = int(info[0]), int(info[1])
        depth = np.zeros((n, h, w), np.uint16)
        ir = np.zeros((n, h, w), np.uint8)
        poses = np.zeros((n, pose_array_size, 7), np.float32)
        for i in range(n):
            self.L.ref_dataset_frame(i, depth[i].ctypes.data, ir[i].ctypes.data, poses[i].ctypes.data)
        return info, depth, ir, poses

    def save(self, basename, cam, depth, ir, poses, segment_scale=0.17):
        """cam = (fx, fy, px, py, depth_scale); writes <basename>.json/.rs/.ir/.pose with the reference's writer."""
        n, h, w = depth.shape
        depth, ir, poses = np.ascontiguousarray(depth, np.uint16), np.ascontiguousarray(ir, np.uint8), np.ascontiguousarray(poses, np.float32)
        rc = self.L.ref_dataset_save(str(basename).encode(), w, h, *[float(v) for v in cam], float(segment_scale), n, poses.shape[1],
                                     depth.ctypes.data, ir.ctypes.data, poses.ctypes.data)
        assert rc == 0


class Ref:
    """The unmodified reference cnn.h (handposedd) behind a C ABI."""

    def __init__(self, fast: bool = False):
        name = "libcnnref_fast.so" if fast else "libcnnref.so"
        if fast:
            flags = open("/proc/cpuinfo").read()
            if " avx2" not in flags or " fma" not in flags:
                name = "libcnnref.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(path)
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_init.argtypes = [C.c_void_p]
        L.ref_load.argtypes = [C.c_void_p, _f32p]
        L.ref_save.argtypes = [C.c_void_p, _f32p]
        L.ref_saveb_file.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_loadb_file.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_set_simd.argtypes = [C.c_int]
        L.ref_eval.argtypes = [C.c_void_p, _f32p, C.c_long, _f32p]
        L.ref_eval_mt.argtypes = [C.c_void_p, _f32p, C.c_long, _f32p, C.c_int]
        L.ref_train_seq.argtypes = [C.c_void_p, _f32p, _f32p, C.c_long, C.c_float, _f32p]
        L.ref_train_mt.argtypes = [C.c_void_p, _f32p, _f32p, C.c_long, C.c_float, C.c_int]
        L.ref_forward_trace.argtypes = [C.c_void_p, _f32p, _f32p, _i32p]
        L.ref_forward_trace.restype = C.c_long
        L.ref_grad_sample.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, C.c_void_p]
        L.ref_grad_sample.restype = C.c_float
        self.L = L
        self.h = C.c_void_p(L.ref_create())

    def __del__(self):
        try:
            self.L.ref_destroy(self.h)
        except Exception:
            pass

    def init(self):
        self.L.ref_init(self.h)

    def load(self, params):
        self.L.ref_load(self.h, _f32(params))

    def save(self):
        p = np.empty(N_PARAMS, np.float32)
        self.L.ref_save(self.h, p)
        return p

    def saveb_file(self, path):
        self.L.ref_saveb_file(self.h, path.encode())

    def loadb_file(self, path):
        self.L.ref_loadb_file(self.h, path.encode())

    def eval(self, x, threads=1):
        x = _f32(x).reshape(-1, N_IN)
        y = np.empty((x.shape[0], N_OUT), np.float32)
        self.L.ref_eval_mt(self.h, x, x.shape[0], y, threads)
        return y

    def train_seq(self, x, t, alpha):
        x = _f32(x).reshape(-1, N_IN)
        t = _f32(t).reshape(-1, N_OUT)
        mse = np.empty(x.shape[0], np.float32)
        self.L.ref_train_seq(self.h, x, t, x.shape[0], alpha, mse)
        return mse

    def train_mt(self, x, t, alpha, threads):
        x = _f32(x).reshape(-1, N_IN)
        t = _f32(t).reshape(-1, N_OUT)
        self.L.ref_train_mt(self.h, x, t, x.shape[0], alpha, threads)

    SIZES = [57600, 57600, 14400, 3600, 9216, 9216, 2304, 2048, 2048, 2304, 2304]

    def forward_trace(self, x):
        out = np.empty(sum(self.SIZES), np.float32)
        sizes = np.zeros(11, np.int32)
        n = self.L.ref_forward_trace(self.h, _f32(x).reshape(N_IN), out, sizes)
        assert n == out.size and list(sizes) == self.SIZES
        return np.split(out, np.cumsum(self.SIZES)[:-1])

    def grad_sample(self, x, t, want_errors=False):
        g = np.empty(N_PARAMS, np.float32)
        errs = np.empty(sum(self.SIZES), np.float32) if want_errors else None
        mse = self.L.ref_grad_sample(self.h, _f32(x).reshape(N_IN), _f32(t).reshape(N_OUT), g,
                                     errs.ctypes.data_as(C.c_void_p) if want_errors else None)
        if want_errors:
            return g, mse, np.split(errs, np.cumsum(self.SIZES)[:-1])
        return g, mse
