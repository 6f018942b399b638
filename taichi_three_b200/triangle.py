"""TriangleRaster: drop-in for the reference's tina/core/triangle.py:4-153 on top of
libtina_b200 (hand-written sm_100a kernels, C ABI in include/tina_b200.h)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _stream
from .field import Field, wrap_device
from .material import material_struct
from .mesh import MAX, FaceSource, _to_device_f32
from .shader import Shader, ShaderGroup, _Sink


def _fp(a):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.float32).ctypes.data_as(C.POINTER(C.c_float))


class TriangleRaster:
    def __init__(self, engine, maxfaces=MAX, smoothing=False, texturing=False, culling=True, clipping=True,
                 **extra_options):
        self.engine = engine
        self.res = engine.res
        self.maxfaces = maxfaces
        self.smoothing, self.texturing = bool(smoothing), bool(texturing)
        self.culling, self.clipping = bool(culling), bool(clipping)
        self.flags = ((_lib.TINA_SMOOTHING if smoothing else 0) | (_lib.TINA_TEXTURING if texturing else 0) |
                      (_lib.TINA_CULLING if culling else 0) | (_lib.TINA_CLIPPING if clipping else 0))
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.tina_raster_create(C.byref(h), engine._h, int(maxfaces), self.flags))
        self._h = h
        self._occup_fn, self._color_fn = L.tina_raster_render_occup, L.tina_raster_render_color
        self._dev_index = engine.device.index
        self._shader_cache = {}
        self._bg_cache = (None, None)
        self._keep = []  # tensors the C side currently aliases
        self._occup = None
        self._occup_stale = True
        self._direct = {}  # set_face_* fast path staging

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().tina_raster_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- set_object (triangle.py:72-86) ------------------------------------------------
    def set_object(self, mesh):
        if not hasattr(mesh, '_source'):
            raise TypeError(f'{type(mesh).__name__} does not describe itself as a FaceSource; '
                            'use tina SimpleMesh / MeshModel / MeshGrid / MeshTransform / MeshNoCulling ...')
        self._set_source(mesh._source())

    def _set_source(self, s: FaceSource):
        L = _lib.lib()
        st = _stream()
        n = int(s.nfaces)
        # (the reference never checks maxfaces here and overruns its fields, triangle.py:74-78)
        total = n * (2 if s.double_sided else 1)
        if total > self.maxfaces:
            raise ValueError(f'{total} faces exceed maxfaces={self.maxfaces}: pass maxfaces=... to Scene / TriangleRaster')
        # nested MeshTransform wrappers: a chain of matrices, innermost first (each applied with its own rounding)
        chain = s.trans if isinstance(s.trans, list) else ([s.trans] if s.trans is not None else [])
        nchain = s.trans_normal if isinstance(s.trans_normal, list) else ([s.trans_normal] if s.trans is not None else [])
        ntrans = len(chain)
        trans = _fp(np.stack(chain)) if ntrans else None
        tn = _fp(np.stack(nchain)) if ntrans else None
        keep = []
        if s.kind == 'simple' and s.trans is None and s.mode == 0:
            v, nn, tt = s.verts, s.norms if self.smoothing else None, s.coors if self.texturing else None
            if n and v is None:
                raise ValueError('SimpleMesh has no vertices: call set_face_verts first')
            if n and self.smoothing and nn is None:
                raise ValueError('smoothing=True needs set_face_norms')
            if n and self.texturing and tt is None:
                raise ValueError('texturing=True needs set_face_coors')
            keep = [v, nn, tt]
            ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None and n else None
            _lib.check(L.tina_raster_set_faces(self._h, ptr(v), ptr(nn), ptr(tt), n, 1, st))
        elif s.kind in ('simple', 'indexed'):
            if s.kind == 'simple':  # wrapped SimpleMesh: trivial index buffer -> general gather
                idx = torch.arange(n * 3, dtype=torch.int32, device=s.verts.device).view(n, 3, 1).expand(n, 3, 3).contiguous()
                v, vt, vn, faces = s.verts.view(-1, 3), s.coors, s.norms, idx
                vt = vt.view(-1, 2) if vt is not None else None
                vn = vn.view(-1, 3) if vn is not None else None
            else:
                v, vt, vn, faces = s.v, s.vt, s.vn, s.faces
            if self.smoothing and vn is None:
                raise ValueError('smoothing=True needs normals')
            if self.texturing and vt is None:
                raise ValueError('texturing=True needs texture coordinates')
            keep = [v, vt, vn, faces]
            ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            _lib.check(L.tina_raster_set_faces_indexed(self._h, ptr(v), v.shape[0], ptr(vt) if self.texturing else None,
                                                       ptr(vn) if self.smoothing else None,
                                                       vn.shape[0] if vn is not None else 0, ptr(faces), n, trans, tn, ntrans,
                                                       s.mode, st))
        elif s.kind == 'grid':
            keep = [s.pos]
            _lib.check(L.tina_raster_set_faces_grid(self._h, C.c_void_p(s.pos.data_ptr()), s.nx, s.ny, trans, tn, ntrans, s.mode, st))
        else:
            raise ValueError(s.kind)
        self._keep = keep
        self._occup_stale = True

    # north-star fast path: feed face arrays straight to the raster (zero-copy for CUDA tensors)
    def set_face_verts(self, verts):
        self._direct = {'verts': _to_device_f32(verts, self.engine.device, (3, 3))}
        self._apply_direct()

    def set_face_norms(self, norms):
        self._direct['norms'] = _to_device_f32(norms, self.engine.device, (3, 3))
        self._apply_direct()

    def set_face_coors(self, coors):
        self._direct['coors'] = _to_device_f32(coors, self.engine.device, (3, 2))
        self._apply_direct()

    def _apply_direct(self):
        d = self._direct
        if 'verts' not in d or (self.smoothing and 'norms' not in d) or (self.texturing and 'coors' not in d):
            return
        self._set_source(FaceSource('simple', verts=d['verts'], norms=d.get('norms'), coors=d.get('coors'),
                                    nfaces=d['verts'].shape[0]))

    # ---- render_occup (triangle.py:89-131) ---------------------------------------------
    def render_occup(self):
        rc = self._occup_fn(self._h, _stream(self._dev_index))
        if rc:
            _lib.check(rc)
        self._occup_stale = True

    # ---- render_color (triangle.py:134-153) --------------------------------------------
    def render_color(self, shader, fill_bg=None, tonemap=False, finish=False, accum=None):
        """`shader`: a Shader or a ShaderGroup of Shaders.  Fusion hints used by Scene.render: fill_bg = also write the
        background to the pixels this object does not own (first object), tonemap = ACES on what the pass writes,
        finish = last object of a multi-object frame: pixels of the other objects are read back and tonemapped too,
        accum = (accumulator tensor [W,H,3], count): the frame's TAA accumulation (needs fill_bg or finish)."""
        shaders = shader.shaders if isinstance(shader, ShaderGroup) else (shader,)
        flags = ((_lib.TINA_COLOR_TONEMAP if tonemap else 0) | (_lib.TINA_COLOR_FILL_BG if fill_bg is not None else 0) |
                 (_lib.TINA_COLOR_FINISH if finish and fill_bg is None else 0))
        bg = None
        if fill_bg is not None:
            key, bg = self._bg_cache
            if key is not fill_bg:  # (callers reuse one array per frame loop)
                arr = np.ascontiguousarray(np.broadcast_to(np.asarray(fill_bg, dtype=np.float32), (3,)))
                bg = (arr, arr.ctypes.data_as(C.POINTER(C.c_float)))
                self._bg_cache = (fill_bg, bg)
            bg = bg[1]
        st = _stream(self._dev_index)
        # every G-buffer sink of the group (pre / post shaders, ProbeShader parts) goes into ONE launch
        sinks = [p for s in shaders for p in (s._sinks() if hasattr(s, '_sinks') else ((s,) if isinstance(s, _Sink) else ()))]
        for i in range(0, len(sinks), _lib.TINA_MAX_SINKS):
            self._render_sinks(sinks[i:i + _lib.TINA_MAX_SINKS], st)
        for s in shaders:
            if isinstance(s, _Sink) or hasattr(s, '_sinks'):
                continue
            rec = self._shader_cache.get(id(s))
            img = s.img
            if rec is None or rec[0] is not s or rec[1] is not img:
                if not isinstance(s, Shader):
                    raise NotImplementedError(f'{type(s).__name__} is not supported by the B200 render_color yet')
                t = img.to_torch() if hasattr(img, 'to_torch') else img
                if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != self.res[0] * self.res[1] * 3 or not t.is_cuda:
                    raise ValueError('shader image must be a contiguous float32 [W, H, 3] CUDA tensor')
                rec = (s, img, t, C.c_void_p(t.data_ptr()))
                self._shader_cache[id(s)] = rec
            mat, keep = self._material_struct(s.material)
            if accum is not None:
                rc = _lib.lib().tina_raster_render_color_accumulate(self._h, mat, s.lighting.struct_ref(), rec[3], flags, bg,
                                                                    C.c_void_p(accum[0].data_ptr()), int(accum[1]), st)
            else:
                rc = self._color_fn(self._h, mat, s.lighting.struct_ref(), rec[3], flags, bg, st)
            if rc:
                _lib.check(rc)
            self._mat_keep = keep

    def render_color_range(self, shader, first_pixel, npixels, face_base=0, fill_bg=None):
        """Shade pixels [first_pixel, first_pixel + npixels) of the current face arrays, ids offset by
        face_base (sort-last: one screen strip per rank after the key composite)."""
        if not isinstance(shader, Shader):
            raise NotImplementedError('render_color_range takes a single Shader')
        t = shader.img.to_torch() if hasattr(shader.img, 'to_torch') else shader.img
        mat, keep = self._material_struct(shader.material)
        flags = _lib.TINA_COLOR_FILL_BG if fill_bg is not None else 0
        bg = _fp(np.broadcast_to(np.asarray(fill_bg if fill_bg is not None else 0, dtype=np.float32), (3,)))
        _lib.check(_lib.lib().tina_raster_render_color_range(self._h, mat, shader.lighting.struct_ref(), C.c_void_p(t.data_ptr()),
                                                             flags, bg, int(first_pixel), int(npixels), int(face_base),
                                                             _stream(self._dev_index)))
        self._mat_keep = keep

    def render_color_composite(self, shader, first_pixel, npixels, face_base=0, fill_bg=None):
        """render_color_range whose keys are the MIN over the key buffers of all ranks mapped by
        Engine.open_peer_keys(), read over NVLink inside the shading kernel (sort-last composite fused with shading).
        The caller orders it after every rank's render_occup (e.g. dist.barrier() on the same stream)."""
        if not isinstance(shader, Shader):
            raise NotImplementedError('render_color_composite takes a single Shader')
        t = shader.img.to_torch() if hasattr(shader.img, 'to_torch') else shader.img
        mat, keep = self._material_struct(shader.material)
        flags = _lib.TINA_COLOR_FILL_BG if fill_bg is not None else 0
        bg = _fp(np.broadcast_to(np.asarray(fill_bg if fill_bg is not None else 0, dtype=np.float32), (3,)))
        _lib.check(_lib.lib().tina_raster_render_color_composite(self._h, mat, shader.lighting.struct_ref(), C.c_void_p(t.data_ptr()),
                                                                 flags, bg, int(first_pixel), int(npixels), int(face_base),
                                                                 _stream(self._dev_index)))
        self._mat_keep = keep

    def _render_sinks(self, sinks, st):
        """G-buffer shaders (shader.py:21-109, probe.py:21-23) for the current object: one launch for all of them."""
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
        _lib.check(_lib.lib().tina_raster_render_gbuffers(self._h, n, kinds, outs, ncomps, isint, _fp(params), st))

    def _material_struct(self, material):
        """Flattening + folding is pure host work: cache it per material object, keyed by the
        current values of its runtime Param nodes (matr/nodes.py:52-76)."""
        from .material import param_signature
        cache = self.__dict__.setdefault('_mat_cache', {})
        hit = cache.get(id(material))
        if hit is not None and hit[3] is material and not hit[4]:
            return hit[1], hit[2]  # no runtime Param in the graph: nothing can have changed
        sig = param_signature(material)
        if hit is None or hit[0] != sig or hit[3] is not material:
            mat, keep = material_struct(material, self.engine.device)
            hit = (sig, C.byref(mat), (keep, mat), material, len(sig) > 0)
            cache[id(material)] = hit
        return hit[1], hit[2]

    # ---- public state ------------------------------------------------------------------
    @property
    def occup(self):
        """int32[W, H] face id per pixel, -1 = none (triangle.py:16), for the last render_occup."""
        if self._occup is None:
            self._occup = torch.empty(self.res, dtype=torch.int32, device=self.engine.device)
        if self._occup_stale:
            _lib.check(_lib.lib().tina_raster_occup(self._h, C.c_void_p(self._occup.data_ptr()), _stream()))
            self._occup_stale = False
        return Field(self._occup)

    def _setup_cache(self, which):
        """triangle.py:25-29: bcn / can / boo / coo [maxfaces, 2], wsc [maxfaces, 3] -- the per-face setup the reference
        stores during render_occup (:127-131, faces that pass cull + clip only; other rows keep earlier contents).  The
        kernels here recompute it on the fly, so the fields are filled on demand for the current object and camera."""
        c = self.__dict__.get('_cache_fields')
        if c is None:
            dev = self.engine.device
            c = {k: torch.zeros((self.maxfaces, 3 if k == 'wsc' else 2), dtype=torch.float32, device=dev)
                 for k in ('bcn', 'can', 'boo', 'coo', 'wsc')}
            self._cache_fields = c
        if self.nfaces:
            _lib.check(_lib.lib().tina_raster_setup_cache(self._h, *(C.c_void_p(c[k].data_ptr()) for k in ('bcn', 'can', 'boo', 'coo', 'wsc')),
                                                          _stream()))
        return Field(c[which])

    bcn = property(lambda self: self._setup_cache('bcn'))
    can = property(lambda self: self._setup_cache('can'))
    boo = property(lambda self: self._setup_cache('boo'))
    coo = property(lambda self: self._setup_cache('coo'))
    wsc = property(lambda self: self._setup_cache('wsc'))

    def _buffers(self):
        # the indexed path (MeshGrid / MeshModel) keeps no expanded copies until somebody reads them
        _lib.check(_lib.lib().tina_raster_materialize(self._h, _stream()))
        v, n, t, nf = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(_lib.lib().tina_raster_buffers(self._h, C.byref(v), C.byref(n), C.byref(t), C.byref(nf)))
        return v.value, n.value, t.value, nf.value

    @property
    def nfaces(self):
        nf = C.c_int64()
        _lib.check(_lib.lib().tina_raster_buffers(self._h, None, None, None, C.byref(nf)))
        return nf.value

    @property
    def verts(self):
        v, _, _, nf = self._buffers()
        return Field(wrap_device(v, (nf, 3, 3), torch.float32, self.engine.device, owner=self))

    @property
    def norms(self):
        _, n, _, nf = self._buffers()
        return Field(wrap_device(n, (nf, 3, 3), torch.float32, self.engine.device, owner=self))

    @property
    def coors(self):
        _, _, t, nf = self._buffers()
        return Field(wrap_device(t, (nf, 3, 2), torch.float32, self.engine.device, owner=self))

    def kernel_times(self):
        """ms of the last launch of each kernel (needs set_tuning(profile=1)); -1 = not recorded."""
        out = (C.c_float * 5)()
        _lib.check(_lib.lib().tina_raster_kernel_times(self._h, out))
        return dict(zip(('raster_faces', 'frame_prologue', 'unused2', 'large_path', 'render_color'), list(out)))

    def set_tuning(self, tiny_max=None, force_tiles=None, collect_stats=None, profile=None, tighten=None,
                   precheck=None, scan_max=None, generic_vm=None, balance=None, pdl=None, indexed=None, adaptive=None,
                   fast_shading=None, lean_kernels=None, force_general=None, grid_tiles=None, grid_quads=None, overlap_vertex=None, persist_k4=None):
        """Strategy knobs (every setting produces identical ids/depth bits; only fast_shading changes colour, by
        < 1e-4): tiny_max = most candidate pixels a
        face may have to be rasterised per thread in the setup kernel (more -> tile path);
        force_tiles = every face through the tile path; tighten = skip bbox pixels whose sample
        provably fails (0 = walk the full reference bbox); precheck = read the key before the
        atomicMin; scan_max = largest queue the tile path handles without binning;
        generic_vm = always interpret the material program; balance = warp-shared candidate walk in
        the setup kernel (0 never, 1 auto per warp, 2 always); pdl = programmatic dependent launch;
        indexed = per-unique-vertex stage for MeshGrid / MeshModel (takes effect at the next set_object);
        adaptive = stop launching the tile-path kernel after 8 consecutive calls that queued nothing;
        fast_shading = FMA / SFU arithmetic downstream of the (always exact) barycentric weights in render_color;
        lean_kernels = specialised shading kernels for constant-parameter Diffuse / Classic on untextured rasters;
        force_general = indexed sources: every face takes the rasteriser's general path (no per-vertex integer bounds);
        grid_tiles = plain square MeshGrid: persistent warps over cp.async-staged row chunks (default 0: slower);
        grid_quads = plain square MeshGrid: one quad (two faces) per thread (default 1; 0 = one face per thread);
        overlap_vertex = when no set_object happened since the previous render_occup, the vertex stage of this one starts
        while the previous render_color is still in its last wave (default 1; it writes the other of two record sets);
        persist_k4 = render_color passes without frame glue / composite run as a smaller grid walking the 256-pixel chunks
        with a grid stride, persist_k4 / 4 chunks per CTA (default 15 = 3.75; 0 = one CTA per chunk)."""
        L = _lib.lib()
        if tiny_max is not None:
            _lib.check(L.tina_raster_set_tuning(self._h, 0, int(tiny_max)))
        if force_tiles is not None:
            _lib.check(L.tina_raster_set_tuning(self._h, 2, int(force_tiles)))
        if collect_stats is not None:
            _lib.check(L.tina_raster_set_tuning(self._h, 3, int(collect_stats)))
        if profile is not None:
            _lib.check(L.tina_raster_set_tuning(self._h, 4, int(profile)))
        for which, v in ((5, tighten), (6, precheck), (7, scan_max), (8, generic_vm), (9, balance), (10, pdl), (11, indexed), (12, adaptive),
                         (13, fast_shading), (14, lean_kernels), (15, force_general), (16, grid_tiles), (17, grid_quads), (18, overlap_vertex), (19, persist_k4)):
            if v is not None:
                _lib.check(L.tina_raster_set_tuning(self._h, which, int(v)))

    def stats(self):
        out = (C.c_int64 * 6)()
        _lib.check(_lib.lib().tina_raster_stats(self._h, out))
        return dict(zip(('culled', 'clipped', 'survivors', 'unused', 'queued', 'tile_entries'), list(out)))
