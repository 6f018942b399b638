"""ParticleRaster + particle providers: drop-in for the reference's tina/core/particle.py:4-161,
tina/pars/simple.py, tina/pars/trans.py on top of libtina_b200 (tina_pars_* in include/tina_b200.h).
Particles share the Engine's depth / id buffer with the triangle raster."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _stream
from .field import Field
from .material import material_struct, param_signature
from .mesh import MAX, _device
from .shader import Shader, ShaderGroup, _Sink


def _dev_f32(arr, shape_tail):
    if isinstance(arr, torch.Tensor):
        t = arr
        if t.device.type != 'cuda' or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=_device(), dtype=torch.float32).contiguous()
    else:
        t = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32)).to(_device())
    if tuple(t.shape[1:]) != tuple(shape_tail):
        raise ValueError(f'expected shape [N{"".join(", " + str(d) for d in shape_tail)}], got {tuple(t.shape)}')
    return t


class SimpleParticles:
    """pars/simple.py:5-55."""

    def __init__(self, maxpars=65536, radius=0.02):
        self.maxpars, self.radius = maxpars, radius
        self._verts = self._sizes = self._colors = None
        self._n = 0

    def get_npars(self):
        return min(self._n, self.maxpars)

    def set_particles(self, verts):
        v = _dev_f32(verts, (3,))
        if v.shape[0] > self.maxpars:
            raise ValueError(f'{v.shape[0]} particles exceed maxpars={self.maxpars}')
        self._verts, self._n = v, v.shape[0]
        if self._sizes is None or self._sizes.shape[0] != self._n:  # simple.py:12-14 defaults
            self._sizes = torch.full((self._n,), float(np.float32(self.radius)), dtype=torch.float32, device=v.device)
        if self._colors is None or self._colors.shape[0] != self._n:
            self._colors = torch.ones((self._n, 3), dtype=torch.float32, device=v.device)

    def set_particle_radii(self, sizes):
        self._sizes = _dev_f32(sizes, ())[:self._n].contiguous()

    def set_particle_colors(self, colors):
        self._colors = _dev_f32(colors, (3,))[:self._n].contiguous()

    @property
    def verts(self):
        return Field(self._verts)

    @property
    def sizes(self):
        return Field(self._sizes)

    @property
    def colors(self):
        return Field(self._colors)

    def _source(self):
        return dict(verts=self._verts, sizes=self._sizes, colors=self._colors, n=self.get_npars(), trans=None, scale=1.0)


class ParsTransform:
    """pars/trans.py:5-31."""

    def __init__(self, pars):
        self.pars = pars
        self.trans, self.scale = np.eye(4, dtype=np.float32), 1.0

    def __getattr__(self, attr):
        return getattr(self.__dict__['pars'], attr)

    def set_transform(self, trans, scale):
        self.trans, self.scale = np.asarray(trans, dtype=np.float64).astype(np.float32), float(scale)

    def _source(self):
        s = self.pars._source()
        if s['trans'] is not None:
            raise NotImplementedError('nested ParsTransform is not supported')
        s['trans'], s['scale'] = self.trans, self.scale
        return s


class ParticleRaster:
    def __init__(self, engine, maxpars=MAX, coloring=True, clipping=True, **extra_options):
        self.engine, self.res = engine, engine.res
        self.maxpars, self.coloring, self.clipping = maxpars, bool(coloring), bool(clipping)
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.tina_pars_create(C.byref(h), engine._h, int(maxpars), (1 if coloring else 0) | (2 if clipping else 0)))
        self._h = h
        self._keep = None
        self._occup = None
        self._mat_cache = {}

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().tina_pars_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_object(self, pars):  # particle.py:64-76
        if not hasattr(pars, '_source'):
            raise TypeError(f'{type(pars).__name__} is not a particle provider of this package')
        s = pars._source()
        n = int(s['n'])
        if n > self.maxpars:
            raise ValueError(f'{n} particles exceed maxpars={self.maxpars}')
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None and n else None
        trans = np.ascontiguousarray(s['trans'], dtype=np.float32).ctypes.data_as(C.POINTER(C.c_float)) if s['trans'] is not None else None
        _lib.check(_lib.lib().tina_pars_set(self._h, p(s['verts']), p(s['sizes']), p(s['colors']) if self.coloring else None, n,
                                            trans, float(s['scale']), 1, _stream()))
        self._keep = s

    # direct setters of the reference's raster (particle.py:44-61)
    def set_particles(self, verts):
        self._direct = SimpleParticles(maxpars=self.maxpars, radius=0.1)  # particle.py:24: sizes.fill(0.1)
        self._direct.set_particles(verts)
        self.set_object(self._direct)

    def set_particle_radii(self, sizes):
        self._direct.set_particle_radii(sizes)
        self.set_object(self._direct)

    def set_particle_colors(self, colors):
        self._direct.set_particle_colors(colors)
        self.set_object(self._direct)

    def render_occup(self):  # particle.py:78-127
        _lib.check(_lib.lib().tina_pars_render_occup(self._h, _stream()))

    def render_color(self, shader, fill_bg=None, tonemap=False):  # particle.py:129-161
        shaders = shader.shaders if isinstance(shader, ShaderGroup) else (shader,)
        flags = (_lib.TINA_COLOR_TONEMAP if tonemap else 0) | (_lib.TINA_COLOR_FILL_BG if fill_bg is not None else 0)
        bg = np.ascontiguousarray(np.broadcast_to(np.asarray(fill_bg if fill_bg is not None else 0, dtype=np.float32), (3,)))
        # G-buffer shaders of the group (shader.py:21-109, probe.py:21-23): one launch per eight sinks
        sinks = []
        for s in shaders:
            if isinstance(s, _Sink):
                sinks.append(s)
            elif hasattr(s, '_sinks'):
                sinks.extend(s._sinks())
            elif not isinstance(s, Shader):
                raise NotImplementedError(f'{type(s).__name__} is not supported by the B200 particle render_color')
        for i in range(0, len(sinks), _lib.TINA_MAX_SINKS):
            self._render_sinks(sinks[i:i + _lib.TINA_MAX_SINKS])
        for s in shaders:
            if not isinstance(s, Shader):
                continue
            t = s.img.to_torch() if hasattr(s.img, 'to_torch') else s.img
            key = id(s.material)
            sig = param_signature(s.material)
            hit = self._mat_cache.get(key)
            if hit is None or hit[0] != sig or hit[2] is not s.material:
                hit = (sig, material_struct(s.material, self.engine.device, color_is_one=False), s.material)
                self._mat_cache[key] = hit
            mat, keep = hit[1]
            _lib.check(_lib.lib().tina_pars_render_color(self._h, C.byref(mat), s.lighting.struct_ref(), C.c_void_p(t.data_ptr()),
                                                         flags, bg.ctypes.data_as(C.POINTER(C.c_float)), _stream()))

    def _render_sinks(self, sinks):
        n = len(sinks)
        npix = self.res[0] * self.res[1]
        kinds, outs, ncomps, isint = (C.c_int * n)(), (C.c_void_p * n)(), (C.c_int * n)(), (C.c_int * n)()
        params = np.zeros((n, 3), dtype=np.float32)
        for i, s in enumerate(sinks):
            t = s.img.to_torch() if hasattr(s.img, 'to_torch') else s.img
            if t.dtype not in (torch.float32, torch.int32) or not t.is_contiguous() or not t.is_cuda or t.numel() % npix:
                raise ValueError('G-buffer image must be a contiguous float32 / int32 CUDA tensor [W, H] or [W, H, n]')
            if not 1 <= t.numel() // npix <= 3:
                raise ValueError('G-buffer image must have 1..3 components')
            kinds[i], outs[i], ncomps[i], isint[i] = s.kind, t.data_ptr(), t.numel() // npix, 1 if t.dtype == torch.int32 else 0
            p = s.param()
            if p is not None:
                params[i] = np.asarray(p, dtype=np.float32).reshape(-1)[:3]
        _lib.check(_lib.lib().tina_pars_render_gbuffers(self._h, n, kinds, outs, ncomps, isint, params.ctypes.data_as(C.POINTER(C.c_float)),
                                                        _stream()))

    @property
    def occup(self):
        if self._occup is None:
            self._occup = torch.empty(self.res, dtype=torch.int32, device=self.engine.device)
        _lib.check(_lib.lib().tina_pars_occup(self._h, C.c_void_p(self._occup.data_ptr()), _stream()))
        return Field(self._occup)
