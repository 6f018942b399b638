"""WireframeRaster + MeshToWire: drop-in for the reference's tina/core/wireframe.py:4-95 and tina/mesh/wire.py on
top of libtina_b200 (tina_wire_* in include/tina_b200.h).  Lines share the Engine's depth / id buffer."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _stream
from .mesh import MAX, MeshEditBase, _to_device_f32
from .shader import Shader, ShaderGroup


class MeshToWire(MeshEditBase):
    """mesh/wire.py:5-39: the edges of a polygon mesh as 2-vertex "faces"."""

    def __init__(self, mesh):
        super().__init__(mesh)
        self.src_npolygon = mesh.get_npolygon() if hasattr(mesh, 'get_npolygon') else 3

    def get_npolygon(self):
        return 2

    def get_nfaces(self):
        return self.mesh.get_nfaces() * self.src_npolygon


class WireframeRaster:
    def __init__(self, engine, maxwires=MAX, linewidth=1.5, linecolor=(.9, .6, 0), clipping=False, **extra_options):
        self.engine, self.res = engine, engine.res
        self.maxfaces, self.linewidth, self.clipping = maxwires, linewidth, bool(clipping)
        h = C.c_void_p()
        col = np.ascontiguousarray(linecolor, dtype=np.float32)
        _lib.check(_lib.lib().tina_wire_create(C.byref(h), engine._h, int(maxwires), 2 if clipping else 0,
                                               col.ctypes.data_as(C.POINTER(C.c_float))))
        self._h = h
        self._keep = None
        self._expander = None

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().tina_wire_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def set_linecolor(self, linecolor):
        col = np.ascontiguousarray(linecolor, dtype=np.float32)
        _lib.check(_lib.lib().tina_wire_set_color(self._h, col.ctypes.data_as(C.POINTER(C.c_float))))

    def set_wire_verts(self, verts):  # wireframe.py:40-46
        v = _to_device_f32(verts, self.engine.device, (2, 3))
        _lib.check(_lib.lib().tina_wire_set(self._h, C.c_void_p(v.data_ptr()) if v.shape[0] else None, v.shape[0], 0, _stream()))
        self._keep = v

    def set_object(self, mesh):  # wireframe.py:32-38
        if not isinstance(mesh, MeshToWire):
            raise TypeError('WireframeRaster.set_object takes a MeshToWire(mesh)')
        if mesh.src_npolygon != 3:
            raise NotImplementedError('only triangle meshes can be turned into wires here')
        # expanded [N,3,3] faces of the wrapped mesh through the triangle raster's set_object adapters
        if self._expander is None:
            from .triangle import TriangleRaster
            self._expander = TriangleRaster(self.engine, maxfaces=2**31 - 1, culling=False, clipping=False)
        self._expander._set_source(mesh.mesh._source())
        faces = self._expander.verts.to_torch()
        n = faces.shape[0] * 3
        if n > self.maxfaces:
            raise ValueError(f'{n} wires exceed maxwires={self.maxfaces}')
        _lib.check(_lib.lib().tina_wire_set(self._h, C.c_void_p(faces.data_ptr()) if n else None, n, 3, _stream()))
        self._keep = faces

    def render_occup(self):  # wireframe.py:67-68
        pass

    def render_color(self, shader, **_):  # wireframe.py:70-95
        shaders = shader.shaders if isinstance(shader, ShaderGroup) else (shader,)
        imgs = []
        for s in shaders:
            if isinstance(s, Shader):  # the G-buffer shaders inherit IShader.blend_color = no-op (shader.py:17-18)
                t = s.img.to_torch() if hasattr(s.img, 'to_torch') else s.img
                imgs.append(t)
        arr = (C.c_void_p * max(1, len(imgs)))(*[t.data_ptr() for t in imgs])
        _lib.check(_lib.lib().tina_wire_render_color(self._h, arr, len(imgs), _stream()))
