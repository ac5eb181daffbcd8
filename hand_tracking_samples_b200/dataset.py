"""Host-side mirror of the reference's dataset reader (include/dataset.h:109-163, load_dataset) over the C ABI
(csrc/hp_dataset.cu): memory-mapped .rs/.ir, parsed .json/.pose, frames copied on request into NumPy (or pinned)
batch buffers.  SURVEY.md 8f row 4."""
import ctypes as C

import numpy as np

from . import capi


class Dataset:
    """`Dataset(basename, pose_array_size)` ~ load_dataset(bname, pose_array_size) without materialising every frame."""

    def __init__(self, basename, pose_array_size=17):
        self.L = capi.lib()
        h = C.c_void_p()
        capi.check(self.L.hp_dataset_open(str(basename).encode(), int(pose_array_size), C.byref(h)))
        self.h = h
        info = capi.DatasetInfo()
        capi.check(self.L.hp_dataset_get_info(self.h, C.byref(info)))
        self.info = info
        self.n_frames, self.width, self.height = int(info.n_frames), int(info.width), int(info.height)
        self.pose_array_size = int(info.pose_array_size)

    def close(self):
        if getattr(self, "h", None):
            self.L.hp_dataset_close(self.h)
            self.h = None

    __del__ = close

    def __len__(self):
        return self.n_frames

    def read(self, first=0, count=None, depth=None, ir=None, poses=None):
        """Frames [first, first+count) -> (depth u16 [count][h][w], ir u8 [count][h][w], poses f32 [count][np][7])."""
        count = self.n_frames - first if count is None else count
        if depth is None:
            depth = np.empty((count, self.height, self.width), np.uint16)
        if ir is None:
            ir = np.empty((count, self.height, self.width), np.uint8)
        if poses is None:
            poses = np.empty((count, self.pose_array_size, 7), np.float32)
        capi.check(self.L.hp_dataset_read(self.h, first, count, depth.ctypes.data, ir.ctypes.data, poses.ctypes.data))
        return depth, ir, poses

    def eval_depth(self, net, first=0, count=None, dmin=0.1, dmax=0.7, precision=capi.PRECISION_FP32, want_y=True, want_decoded=True):
        """64x64-crop datasets: handtrack.h:700-702 (normalise, Eval, decode) for frames [first, first+count)."""
        count = self.n_frames - first if count is None else count
        y = np.empty((count, capi.N_OUT), np.float32) if want_y else None
        dec = np.empty((count, 48), np.float32) if want_decoded else None
        capi.check(self.L.hp_dataset_eval_depth(net.h, self.h, first, count, dmin, dmax, y.ctypes.data if want_y else None,
                                                 dec.ctypes.data if want_decoded else None, precision))
        return y, dec
