"""Image-space post effects on the raster path's image: FXAA and Blooming (reference tina/postp/fxaa.py,
tina/postp/blooming.py) over tina_image_fxaa / tina_image_bloom of libtina_b200."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _stream
from .field import HostField


def _img(image):
    t = image.to_torch() if hasattr(image, 'to_torch') else image
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError('post effects need a contiguous float32 [W, H, 3] CUDA image')
    return t


class FXAA:
    def __init__(self, res):  # fxaa.py:12-26
        self.res = (int(res[0]), int(res[1]))
        self.abs_thresh = HostField(np.float32(0.0625))
        self.rel_thresh = HostField(np.float32(0.063))
        self.factor = HostField(np.float32(1))
        self._lumi = self._copy = None

    def apply(self, image):  # fxaa.py:28-68
        t = _img(image)
        W, H = t.shape[:2]
        if self._lumi is None or self._lumi.device != t.device:
            self._lumi = torch.empty((W, H), dtype=torch.float32, device=t.device)
            self._copy = torch.empty((W, H, 3), dtype=torch.float32, device=t.device)
        _lib.check(_lib.lib().tina_image_fxaa(C.c_void_p(t.data_ptr()), W, H, C.c_void_p(self._lumi.data_ptr()),
                                              C.c_void_p(self._copy.data_ptr()), float(self.abs_thresh[None]),
                                              float(self.rel_thresh[None]), float(self.factor[None]), _stream()))


class Blooming:
    def __init__(self, res):  # blooming.py:6-24
        self.res = (int(res[0]), int(res[1]))
        self.thresh = HostField(np.float32(1))
        self.factor = HostField(np.float32(1))
        self.radius = HostField(np.int32(min(self.res) // 16))
        self.sigma = HostField(np.float32(1))
        self.scale = HostField(np.float32(0.25))
        self._a = self._b = self._gwei = None
        self._gkey = None

    def gaussian_weights(self):
        """blooming.py:26-37 in f32 (the reference accumulates `sum` with atomics; serial order here)."""
        radius, sigma = int(self.radius[None]), np.float32(self.sigma[None])
        g = np.zeros(radius + 1, dtype=np.float32)
        s = np.float32(-1.0)
        for i in range(radius + 1):
            x = (sigma * np.float32(i)) / np.float32(radius) if radius else np.float32(np.nan)
            y = np.exp(-(x * x)).astype(np.float32)
            g[i] = y
            s = np.float32(s + y * np.float32(2))
        return (g / s).astype(np.float32)

    def apply(self, image):  # blooming.py:45-68
        t = _img(image)
        W, H = t.shape[:2]
        key = (int(self.radius[None]), float(self.sigma[None]))
        if self._gwei is None or self._gkey != key or self._gwei.device != t.device:
            self._gwei = torch.as_tensor(self.gaussian_weights()).to(t.device)
            self._gkey = key
            self._a = torch.empty((W // 2, H // 2, 3), dtype=torch.float32, device=t.device)
            self._b = torch.empty_like(self._a)
        _lib.check(_lib.lib().tina_image_bloom(C.c_void_p(t.data_ptr()), W, H, C.c_void_p(self._a.data_ptr()),
                                               C.c_void_p(self._b.data_ptr()), C.c_void_p(self._gwei.data_ptr()), key[0],
                                               float(self.thresh[None]), float(self.scale[None]), float(self.factor[None]),
                                               _stream()))
