"""Image-space post effects on the raster path's image: FXAA, Blooming, SSAO and SSR (reference tina/postp/fxaa.py,
tina/postp/blooming.py, tina/postp/ssao.py, tina/postp/ssr.py) over tina_image_fxaa / tina_image_bloom /
tina_engine_ssao_render / tina_engine_ssr_render of libtina_b200."""
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


class SSAO:
    """Screen-space ambient occlusion (postp/ssao.py), non-TAA mode: a fixed table of hemisphere samples and a small
    tile of per-pixel rotations, drawn once at construction like the reference (:24-36, :51-56); `norm` is the scene's
    world-normal G-buffer.  The tables are plain tensors (`samples`, `rotations`) and may be overwritten."""

    def __init__(self, res, norm, nsamples=64, thresh=0.0, radius=0.2, factor=1.0, noise_size=4, taa=False, seed=None):
        # taa=True (ssao.py:52-56, 80-81): fresh samples per pixel and frame; the reference draws them from Taichi's
        # unspecified ti.random(), here from the Wang hash of tina/random.py seeded with (pixel, frame)
        self.taa = bool(taa)
        self.frame = 0
        self.res = (int(res[0]), int(res[1]))
        self.norm = norm
        self.radius = HostField(np.float32(radius))
        self.thresh = HostField(np.float32(thresh))
        self.factor = HostField(np.float32(factor))
        self.nsamples = int(nsamples)
        self.noise_size = int(noise_size)
        dev = (norm.to_torch() if hasattr(norm, 'to_torch') else norm).device
        from .field import Field
        self.img = Field(torch.zeros(self.res, dtype=torch.float32, device=dev))
        self.seed_samples(seed)

    def seed_samples(self, seed=None):
        """ssao.py:30-36, 51-56 in float32: make_sample() = spherical(lerp(u, .01, 1), v) * lerp(w**1.5, .01, 1)."""
        rng = np.random.default_rng(seed)
        f = np.float32
        u, v, w = (rng.random(self.nsamples).astype(f) for _ in range(3))
        r = f(0.01) * (f(1) - w ** f(1.5)) + f(1.0) * w ** f(1.5)
        h = f(0.01) * (f(1) - u) + f(1.0) * u
        s = np.sqrt(np.maximum(f(0), f(1) - h * h))
        tau = f(2 * np.pi)
        smp = np.stack([s * np.cos(v * tau), s * np.sin(v * tau), h], axis=1) * r[:, None]
        t = tau * rng.random((self.noise_size, self.noise_size)).astype(f)
        dev = self.img.to_torch().device
        self.samples = torch.as_tensor(np.ascontiguousarray(smp, dtype=f), device=dev)
        self.rotations = torch.as_tensor(np.ascontiguousarray(np.stack([np.cos(t), np.sin(t)], axis=2), dtype=f), device=dev)

    def render(self, engine):  # ssao.py:58-96
        n = self.norm.to_torch() if hasattr(self.norm, 'to_torch') else self.norm
        if n.dtype != torch.float32 or not n.is_contiguous() or tuple(n.shape) != (self.res[0], self.res[1], 3):
            raise ValueError('SSAO needs a contiguous float32 [W, H, 3] normal buffer')
        if self.taa:
            _lib.check(_lib.lib().tina_engine_ssao_render_taa(engine._h, C.c_void_p(n.data_ptr()), self.nsamples, float(self.radius[None]),
                                                              float(self.thresh[None]), float(self.factor[None]), self.frame & 0xffffffff,
                                                              C.c_void_p(self.img.to_torch().data_ptr()), _stream()))
            self.frame += 1
            self._keep = (n,)
            return
        smp, rot = self.samples.contiguous(), self.rotations.contiguous()
        _lib.check(_lib.lib().tina_engine_ssao_render(engine._h, C.c_void_p(n.data_ptr()), C.c_void_p(smp.data_ptr()), int(smp.shape[0]),
                                                      C.c_void_p(rot.data_ptr()), int(rot.shape[0]), float(self.radius[None]),
                                                      float(self.thresh[None]), float(self.factor[None]),
                                                      C.c_void_p(self.img.to_torch().data_ptr()), _stream()))
        self._keep = (n, smp, rot)

    def apply(self, out):  # ssao.py:38-49
        t = _img(out)
        if self.taa:
            _lib.check(_lib.lib().tina_image_ssao_apply_taa(C.c_void_p(t.data_ptr()), C.c_void_p(self.img.to_torch().data_ptr()), t.shape[0],
                                                            t.shape[1], _stream()))
            return
        _lib.check(_lib.lib().tina_image_ssao_apply(C.c_void_p(t.data_ptr()), C.c_void_p(self.img.to_torch().data_ptr()), t.shape[0],
                                                    t.shape[1], self.noise_size, _stream()))


class SSR:
    """Screen-space reflections (postp/ssr.py): per pixel with a normal, `nsamples` rays drawn with `material.sample()`
    of the pixel's material and marched through the depth buffer for at most `nsteps` steps; a hit adds the image colour
    there times the sample weight.  norm / coor / mtlid are the scene's G-buffers (scene/raster.py:51-64), mtltab its
    MaterialTable.  Non-TAA: the reference's WangHashRNG(P % blurring) stream (ssr.py:73-76), reproduced exactly; taa:
    the same hash seeded with (pixel, frame) instead of Taichi's unspecified ti.random()."""

    def __init__(self, res, norm, coor, mtlid, mtltab, taa=False):  # ssr.py:7-28
        self.res = (int(res[0]), int(res[1]))
        self.norm, self.coor, self.mtlid, self.mtltab, self.taa = norm, coor, mtlid, mtltab, bool(taa)
        dev = (norm.to_torch() if hasattr(norm, 'to_torch') else norm).device
        from .field import Field
        self.img = Field(torch.zeros(self.res + (4,), dtype=torch.float32, device=dev))
        self.nsamples = HostField(np.int32(32 if not taa else 12))
        self.nsteps = HostField(np.int32(32 if not taa else 64))
        self.stepsize = HostField(np.float32(2))
        self.tolerance = HostField(np.float32(15))
        self.blurring = HostField(np.int32(4))
        self.frame = 0

    def render(self, engine, image):  # ssr.py:44-103
        from .material import sample_struct
        t = _img(image)
        W, H = self.res
        n = self.norm.to_torch() if hasattr(self.norm, 'to_torch') else self.norm
        m = self.mtlid.to_torch() if hasattr(self.mtlid, 'to_torch') else self.mtlid
        c = None
        if self.coor is not None:
            c = self.coor.to_torch() if hasattr(self.coor, 'to_torch') else self.coor
            if tuple(c.shape) != (W, H, 2):  # the reference's dummy (1, 1) field when the scene has no texturing (raster.py:63-64)
                c = None
        if n.dtype != torch.float32 or not n.is_contiguous() or tuple(n.shape) != (W, H, 3):
            raise ValueError('SSR needs a contiguous float32 [W, H, 3] normal buffer')
        if m.dtype != torch.int32 or not m.is_contiguous() or tuple(m.shape) != (W, H):
            raise ValueError('SSR needs a contiguous int32 [W, H] material-id buffer')
        if c is not None and (c.dtype != torch.float32 or not c.is_contiguous()):
            raise ValueError('SSR needs a contiguous float32 [W, H, 2] texcoord buffer')
        mats = list(self.mtltab.materials)
        if not mats:
            raise ValueError('SSR: the material table is empty')
        table = (_lib.TinaSampleMaterial * len(mats))()
        keep = []
        for i, mat in enumerate(mats):
            table[i], k = sample_struct(mat, n.device)
            keep.append(k)
        _lib.check(_lib.lib().tina_engine_ssr_render(
            engine._h, C.c_void_p(n.data_ptr()), C.c_void_p(c.data_ptr()) if c is not None else None, C.c_void_p(m.data_ptr()), table,
            len(mats), C.c_void_p(t.data_ptr()), int(self.nsamples[None]), int(self.nsteps[None]), float(self.stepsize[None]),
            float(self.tolerance[None]), int(self.blurring[None]), int(self.taa), self.frame & 0xffffffff,
            C.c_void_p(self.img.to_torch().data_ptr()), _stream()))
        if self.taa:
            self.frame += 1
        self._keep = (n, m, c, keep, t)

    def apply(self, image):  # ssr.py:30-42
        t = _img(image)
        _lib.check(_lib.lib().tina_image_ssr_apply(C.c_void_p(t.data_ptr()), C.c_void_p(self.img.to_torch().data_ptr()), t.shape[0], t.shape[1],
                                                   int(self.blurring[None]), int(self.taa), _stream()))
