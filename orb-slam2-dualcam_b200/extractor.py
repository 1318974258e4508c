"""Host-side mirror of the reference's ORBextractor (include/ORBextractor.h:47-108) above the C-ABI.

    ext = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width=640, height=480)
    keypoints, descriptors = ext(image)                    # ORBextractor::operator()(image, mask, kps, desc)
    ext.GetLevels(), ext.GetScaleFactors(), ...            # same getters as the reference

plus the batched forms the B200 path is built for (both cameras of many dual-frames per call):
    ext.extract_batch(imgs[F][C][H][W])   host arrays in, host arrays out (copies inside)
    ext.extract_device(d_imgs, ...)       torch CUDA tensors in/out, asynchronous on the current torch stream
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import stream_handle as _stream_handle
from .capi import KP_DTYPE, check, lib, ptr


class ORBextractor:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, width=640, height=480,
                 cameras=1, max_frames=1, device=0):
        self._h = None
        L = lib()
        h = C.c_void_p()
        check(L.orbx_create(C.byref(h), device, width, height, cameras, max_frames, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST))
        self._h = h
        self.nfeatures, self.scaleFactor, self.nlevels = nfeatures, scaleFactor, nlevels
        self.iniThFAST, self.minThFAST = iniThFAST, minThFAST
        self.width, self.height, self.cameras, self.max_frames, self.device = width, height, cameras, max_frames, device
        self.kp_capacity = check(L.orbx_max_keypoints(h))
        n = nlevels
        self._scale, self._inv_scale, self._sigma2, self._inv_sigma2 = (np.empty(n, np.float32) for _ in range(4))
        self.mnFeaturesPerLevel = np.empty(n, np.int32)
        self.umax = np.empty(16, np.int32)
        check(L.orbx_get_tables(h, *(a.ctypes.data_as(capi.f32p) for a in (self._scale, self._inv_scale, self._sigma2, self._inv_sigma2)),
                                self.mnFeaturesPerLevel.ctypes.data_as(capi.i32p), self.umax.ctypes.data_as(capi.i32p)))

    def close(self):
        if getattr(self, "_h", None):
            lib().orbx_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- reference getters (include/ORBextractor.h:63-83)
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self.scaleFactor

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv_scale.copy()

    def GetScaleSigmaSquares(self):
        return self._sigma2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._inv_sigma2.copy()

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(lib().orbx_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    # ---- operator()
    def __call__(self, image, mask=None):
        """One image (H x W uint8) -> (keypoints[N] structured cv::KeyPoint records, descriptors[N][32] uint8).
        `mask` is ignored, as in the reference (src/ORBextractor.cc:1043-1045)."""
        image = np.asarray(image)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise AssertionError("image.type() == CV_8UC1")   # the reference asserts (src/ORBextractor.cc:1050)
        if self.cameras != 1:
            raise ValueError("operator() on a single image needs an extractor created with cameras=1")
        k, d, n = self.extract_batch(image[None, None])
        return k[0, 0, :n[0, 0]].copy(), d[0, 0, :n[0, 0]].copy()

    def extract_batch(self, imgs):
        """imgs uint8 [F][C][H][W] (host) -> kps [F][C][cap], desc [F][C][cap][32], counts [F][C]."""
        imgs = np.asarray(imgs)
        assert imgs.dtype == np.uint8 and imgs.ndim == 4 and imgs.shape[1:] == (self.cameras, self.height, self.width), imgs.shape
        if imgs.strides[3] != 1 or imgs.strides[1] != imgs.strides[2] * self.height or imgs.strides[0] != imgs.strides[1] * self.cameras:
            imgs = np.ascontiguousarray(imgs)
        F = imgs.shape[0]
        cap = self.kp_capacity
        kps = np.zeros((F, self.cameras, cap), KP_DTYPE)
        desc = np.zeros((F, self.cameras, cap, 32), np.uint8)
        counts = np.zeros((F, self.cameras), np.int32)
        check(lib().orbx_extract(self._h, ptr(imgs), F, imgs.strides[2], ptr(kps), ptr(desc), ptr(counts), cap))
        return kps, desc, counts

    def extract_device(self, d_imgs, d_kps=None, d_desc=None, d_counts=None, stream=None):
        """torch CUDA uint8 tensor [F][C][H][W] -> (kps uint8[F][C][cap][28], desc uint8[F][C][cap][32], counts int32[F][C]),
        all on the device, enqueued on `stream` (default: torch's current stream); no synchronisation."""
        import torch
        assert d_imgs.is_cuda and d_imgs.dtype == torch.uint8 and d_imgs.dim() == 4 and d_imgs.is_contiguous()
        F = d_imgs.shape[0]
        assert tuple(d_imgs.shape[1:]) == (self.cameras, self.height, self.width)
        cap = self.kp_capacity
        dev = d_imgs.device
        if d_kps is None:
            d_kps = torch.empty((F, self.cameras, cap, 28), dtype=torch.uint8, device=dev)
        if d_desc is None:
            d_desc = torch.empty((F, self.cameras, cap, 32), dtype=torch.uint8, device=dev)
        if d_counts is None:
            d_counts = torch.empty((F, self.cameras), dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        L = lib()
        check(L.orbx_set_stream(self._h, _stream_handle(st)))
        check(L.orbx_extract_device(self._h, d_imgs.data_ptr(), F, self.width, d_kps.data_ptr(), d_desc.data_ptr(), d_counts.data_ptr(), cap))
        return d_kps, d_desc, d_counts

    def synchronize(self):
        check(lib().orbx_synchronize(self._h))

    def launch_count(self):
        return int(lib().orbx_launch_count(self._h))

    def profile(self, enable=True):
        check(lib().orbx_profile(self._h, int(enable)))

    def stage_ms(self):
        """(dict of summed stage milliseconds, calls) since profile(True) / the last stage_ms()"""
        ms = np.zeros(4, np.float64)
        n = C.c_int()
        check(lib().orbx_stage_ms(self._h, ms.ctypes.data_as(capi.f64p), C.byref(n)))
        return dict(zip(("pyramid", "fast", "quadtree", "describe"), ms.tolist())), n.value

    # ---- stage taps (parity tests)
    def debug_level(self, img, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        check(lib().orbx_debug_level(self._h, img, level, ptr(out), out.nbytes))
        return out

    def _debug_list(self, fn, img, level, cap=1 << 18):
        out = np.empty((cap, 3), np.int32)
        n = check(fn(self._h, img, level, ptr(out), cap))
        return out[:min(n, cap)].copy()

    def debug_candidates(self, img, level):
        return self._debug_list(lib().orbx_debug_candidates, img, level)

    def debug_selected(self, img, level):
        return self._debug_list(lib().orbx_debug_selected, img, level)
