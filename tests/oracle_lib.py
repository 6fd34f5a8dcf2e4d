"""ctypes binding of oracle/libpigo_oracle.so (the CPU checker; test infrastructure only).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

DET_DTYPE = np.dtype([("row", "<i4"), ("col", "<i4"), ("scale", "<i4"), ("q", "<f4")])


def build(force: bool = False) -> str:
    so = os.path.join(ORACLE_DIR, "libpigo_oracle.so")
    src = os.path.join(ORACLE_DIR, "pigo_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "libpigo_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, vp = C.POINTER(C.c_uint8), C.c_void_p
        L.oracle_face_create.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
        L.oracle_face_create.restype = C.c_int
        L.oracle_face_destroy.argtypes = [vp]
        L.oracle_face_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_classify_region.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.oracle_classify_region.restype = C.c_float
        L.oracle_classify_rotated_region.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, vp, C.c_int]
        L.oracle_classify_rotated_region.restype = C.c_float
        L.oracle_scale_ladder.argtypes = [C.c_int, C.c_int, C.c_double, vp, C.c_int]
        L.oracle_scale_ladder.restype = C.c_int
        L.oracle_count_windows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.oracle_count_windows.restype = C.c_int64
        L.oracle_run_cascade.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_double, vp, C.c_int64, vp]
        L.oracle_run_cascade.restype = C.c_int64
        L.oracle_run_cascade_batch.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_double, C.c_double, C.c_double, vp, C.c_int64, vp, C.c_int]
        L.oracle_run_cascade_batch.restype = C.c_int64
        L.oracle_cluster.argtypes = [vp, C.c_int64, C.c_double, vp, C.c_int64]
        L.oracle_cluster.restype = C.c_int64
        L.oracle_puploc_create.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
        L.oracle_puploc_create.restype = C.c_int
        L.oracle_puploc_destroy.argtypes = [vp]
        L.oracle_puploc_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint32)]
        L.oracle_puploc_classify.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_double, C.c_int, C.c_int, vp,
                                             C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.oracle_puploc_run_detector.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_int, vp, vp, C.c_int, C.c_int,
                                                 C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                 C.POINTER(C.c_float)]
        L.oracle_puploc_run_detector.restype = C.c_int
        L.oracle_get_landmark_seed.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                               C.POINTER(C.c_float)]
        L.oracle_rgba_to_gray.argtypes = [vp, C.c_int64, vp]
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleFace:
    """Mirror of pigo.Pigo restricted to the arithmetic (core/pigo.go)."""

    def __init__(self, packet: bytes):
        self._h = C.c_void_p()
        if lib().oracle_face_create(packet, len(packet), C.byref(self._h)) != 0:
            raise ValueError("oracle: malformed face cascade")
        d, n = C.c_uint32(), C.c_uint32()
        lib().oracle_face_info(self._h, C.byref(d), C.byref(n))
        self.depth, self.ntrees = d.value, n.value

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.oracle_face_destroy(self._h)
            self._h = None

    def classify_region(self, r, c, s, pixels: np.ndarray, dim: int) -> float:
        return lib().oracle_classify_region(self._h, r, c, s, _ptr(pixels), dim)

    def classify_rotated_region(self, r, c, s, a, rows, cols, pixels, dim) -> float:
        return lib().oracle_classify_rotated_region(self._h, r, c, s, a, rows, cols, _ptr(pixels), dim)

    def run_cascade(self, pixels: np.ndarray, rows, cols, dim, min_size, max_size, shift, scale, angle=0.0,
                    cap=1 << 16, with_hist=False):
        pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
        out = np.zeros(cap, dtype=DET_DTYPE)
        hist = np.zeros(self.ntrees + 1, dtype=np.int64) if with_hist else None
        n = lib().oracle_run_cascade(self._h, _ptr(pixels), rows, cols, dim, min_size, max_size, shift, scale, angle,
                                     _ptr(out), cap, _ptr(hist) if with_hist else None)
        if n > cap:
            return self.run_cascade(pixels, rows, cols, dim, min_size, max_size, shift, scale, angle, int(n), with_hist)
        return (out[:n].copy(), hist) if with_hist else out[:n].copy()

    def run_cascade_batch(self, frames: np.ndarray, rows, cols, dim, min_size, max_size, shift, scale, angle=0.0,
                          cap_per_frame=4096, nthreads=1):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        nf = frames.shape[0]
        stride = frames.strides[0]
        out = np.zeros((nf, cap_per_frame), dtype=DET_DTYPE)
        n_out = np.zeros(nf, dtype=np.int64)
        lib().oracle_run_cascade_batch(self._h, _ptr(frames), nf, stride, rows, cols, dim, min_size, max_size, shift,
                                       scale, angle, _ptr(out), cap_per_frame, _ptr(n_out), nthreads)
        return out, n_out


def scale_ladder(min_size, max_size, scale_factor):
    buf = np.zeros(4096, dtype=np.int32)
    n = lib().oracle_scale_ladder(min_size, max_size, scale_factor, _ptr(buf), 4096)
    return buf[:n].tolist()


def count_windows(rows, cols, min_size, max_size, shift, scale_factor) -> int:
    return int(lib().oracle_count_windows(rows, cols, min_size, max_size, shift, scale_factor))


def cluster(dets: np.ndarray, iou: float, cap=None):
    """Returns (sorted_dets, clusters); like the reference, sorts its input (a copy here)."""
    d = np.ascontiguousarray(dets, dtype=DET_DTYPE).copy()
    cap = cap or max(len(d), 1)
    out = np.zeros(cap, dtype=DET_DTYPE)
    n = lib().oracle_cluster(_ptr(d), len(d), iou, _ptr(out), cap)
    return d, out[:n].copy()


class OraclePuploc:
    def __init__(self, packet: bytes):
        self._h = C.c_void_p()
        if lib().oracle_puploc_create(packet, len(packet), C.byref(self._h)) != 0:
            raise ValueError("oracle: malformed puploc cascade")
        st, sc, tr, dp = C.c_uint32(), C.c_float(), C.c_uint32(), C.c_uint32()
        lib().oracle_puploc_info(self._h, C.byref(st), C.byref(sc), C.byref(tr), C.byref(dp))
        self.stages, self.scales, self.trees, self.depth = st.value, sc.value, tr.value, dp.value

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.oracle_puploc_destroy(self._h)
            self._h = None

    def classify(self, r, c, s, pixels, rows, cols, dim, angle=0.0, flipv=False):
        res = (C.c_float * 3)()
        lib().oracle_puploc_classify(self._h, r, c, s, angle, rows, cols, _ptr(pixels), dim, int(flipv), res)
        return np.float32(res[0]), np.float32(res[1]), np.float32(res[2])

    def run_detector(self, row, col, scale, perturbs, randoms, pixels, rows, cols, dim, angle=0.0, flipv=False):
        rnd = np.ascontiguousarray(randoms, dtype=np.float32)
        assert rnd.size >= 3 * max(perturbs, 0)
        orow, ocol, osc = C.c_int(), C.c_int(), C.c_float()
        rc = lib().oracle_puploc_run_detector(self._h, row, col, scale, perturbs, _ptr(rnd), _ptr(pixels), rows, cols,
                                              dim, angle, int(flipv), C.byref(orow), C.byref(ocol), C.byref(osc))
        if rc != 0:
            raise ValueError("oracle: perturbs out of range (the reference panics for > 63)")
        return orow.value, ocol.value, np.float32(osc.value)


def landmark_seed(lrow, lcol, rrow, rcol):
    r, c, s = C.c_int(), C.c_int(), C.c_float()
    lib().oracle_get_landmark_seed(lrow, lcol, rrow, rcol, C.byref(r), C.byref(c), C.byref(s))
    return r.value, c.value, np.float32(s.value)


def rgba_to_gray(rgba: np.ndarray) -> np.ndarray:
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
    n = rgba.size // 4
    out = np.zeros(n, dtype=np.uint8)
    lib().oracle_rgba_to_gray(_ptr(rgba), n, _ptr(out))
    return out.reshape(rgba.shape[:-1])


def ycbcr_to_nrgba(y, cb, cr, subsample, width, height, min_x=0, min_y=0) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.uint8); cb = np.ascontiguousarray(cb, dtype=np.uint8); cr = np.ascontiguousarray(cr, dtype=np.uint8)
    out = np.zeros((height, width, 4), dtype=np.uint8)
    L = lib()
    L.oracle_ycbcr_to_nrgba.argtypes = [C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_void_p]
    L.oracle_ycbcr_to_nrgba(_ptr(y), _ptr(cb), _ptr(cr), y.shape[1], cb.shape[1], subsample, min_x, min_y, width, height, _ptr(out))
    return out


def make_ycbcr_planes(rgb_like_seed: int, subsample: int, width: int, height: int, min_x: int = 0, min_y: int = 0):
    """Random Y/Cb/Cr planes with the plane geometry image.NewYCbCr gives Rect(min_x, min_y, min_x+width, min_y+height)."""
    xd = {0: 1, 1: 2, 2: 2, 3: 1, 4: 4, 5: 4}[subsample]
    yd = {0: 1, 1: 1, 2: 2, 3: 2, 4: 1, 5: 2}[subsample]
    cw = (min_x + width - 1) // xd - min_x // xd + 1
    ch = (min_y + height - 1) // yd - min_y // yd + 1
    rng = np.random.default_rng(rgb_like_seed)
    y = rng.integers(0, 256, size=(height, width + 3), dtype=np.uint8)        # YStride > width on purpose
    cb = rng.integers(0, 256, size=(ch, cw + 2), dtype=np.uint8)
    cr = rng.integers(0, 256, size=(ch, cw + 2), dtype=np.uint8)
    return y, cb, cr
