"""Mesh providers accepted by TriangleRaster.set_object (reference tina/mesh/*.py).

In the reference the mesh protocol (`pre_compute / get_nfaces / get_face_verts / norms /
coors`) is inlined into the set_object kernel at JIT time.  Here every provider describes
itself as a FaceSource (a base array layout + at most one transform + winding / normal mode
bits) which TriangleRaster hands to ONE gather kernel of libtina_b200 (K0).
"""
import numpy as np
import torch

from .field import Field

MAX = 2**20  # common.py:7


class FaceSource:
    def __init__(self, kind, **kw):
        self.kind = kind  # 'simple' | 'indexed' | 'grid'
        self.trans = None  # 4x4 float32 (mesh/trans.py)
        self.trans_normal = None  # 3x3 float32
        self.double_sided = False  # MeshNoCulling
        self.flip = False  # MeshFlipCulling
        self.negate = False  # MeshFlipNormal
        self.__dict__.update(kw)

    @property
    def mode(self):
        return (1 if self.double_sided else 0) | (2 if self.flip else 0) | (4 if self.negate else 0)


def _to_device_f32(arr, device, shape_tail):
    """numpy / torch -> contiguous float32 CUDA tensor; CUDA float32 contiguous tensors are aliased (zero-copy)."""
    if isinstance(arr, torch.Tensor):
        t = arr
        if t.device.type != 'cuda' or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=device, dtype=torch.float32).contiguous()
    else:
        t = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32)).to(device)
    if tuple(t.shape[1:]) != tuple(shape_tail):
        raise ValueError(f'expected an array of shape [N, {", ".join(map(str, shape_tail))}], got {tuple(t.shape)}')
    return t


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('taichi_three_b200 meshes live in GPU memory: a CUDA device is required (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


class SimpleMesh:
    """mesh/simple.py:5-115.  `set_face_*` take numpy arrays or (zero-copy) CUDA tensors."""

    def __init__(self, maxfaces=MAX, npolygon=3):
        if npolygon != 3:
            raise NotImplementedError('only triangle meshes (npolygon=3) are on the B200 raster path')
        self.maxfaces = maxfaces
        self.npolygon = npolygon
        self._verts = self._norms = self._coors = None
        self._n = 0

    def get_npolygon(self):
        return self.npolygon

    def get_nfaces(self):
        return min(self._n, self.maxfaces)

    @property
    def nfaces(self):
        return self.get_nfaces()

    def _check(self, t):
        if t.shape[0] > self.maxfaces:
            # the reference truncates silently (simple.py:63); overflowing is almost always a bug
            raise ValueError(f'{t.shape[0]} faces exceed maxfaces={self.maxfaces}: pass SimpleMesh(maxfaces=...)')
        return t

    def set_face_verts(self, verts):  # simple.py:53-67
        self._verts = self._check(_to_device_f32(verts, _device(), (3, 3)))
        self._n = self._verts.shape[0]

    def set_face_norms(self, norms):  # simple.py:69-83
        self._norms = self._check(_to_device_f32(norms, _device(), (3, 3)))

    def set_face_coors(self, coors):  # simple.py:85-98
        if not isinstance(coors, torch.Tensor):
            coors = np.asarray(coors)[:, :, :2]
        self._coors = self._check(_to_device_f32(coors, _device(), (3, 2)))

    @property
    def verts(self):
        return Field(self._verts)

    @property
    def norms(self):
        return Field(self._norms)

    @property
    def coors(self):
        return Field(self._coors)

    def _source(self):
        return FaceSource('simple', verts=self._verts, norms=self._norms, coors=self._coors, nfaces=self.get_nfaces())


def primitive_sphere(lons=32, lats=24, rad=1):
    """mesh/prim.py:21-48 as arrays: float32 [nfaces, 3 corners, 3 (vertex, normal, texcoord), 3].  A lat / lon grid of
    (lats + 1) x (lons + 1) points; cell (lat, lon) gives the triangles (p00, p01, p11) unless lat == 0 and (p11, p10, p00)
    unless lat == lats - 1, in that order.  float64 arithmetic, rounded to float32 at the end, like the reference."""
    la, lo = np.meshgrid(np.arange(lats + 1, dtype=np.float64) / lats, np.arange(lons + 1, dtype=np.float64) / lons, indexing='ij')
    a, o = (la - 0.5) * np.pi, (lo * 2 - 1) * np.pi
    nrm = np.stack([np.cos(a) * np.cos(o), np.cos(a) * np.sin(o), np.sin(a)], axis=-1)
    pts = np.stack([nrm * rad, nrm, np.stack([lo, la, np.zeros_like(la)], axis=-1)], axis=-2)  # [lat, lon, 3 attrs, 3]
    return _quad_cells(pts, lats, lons, skip_first_row_upper=True, skip_last_row_lower=True)


def _quad_cells(pts, lats, lons, skip_first_row_upper=False, skip_last_row_lower=False):
    p00, p01, p11, p10 = pts[:-1, :-1], pts[:-1, 1:], pts[1:, 1:], pts[1:, :-1]
    upper = np.stack([p00, p01, p11], axis=2)  # [lats, lons, 3 corners, 3 attrs, 3]
    lower = np.stack([p11, p10, p00], axis=2)
    both = np.stack([upper, lower], axis=2).reshape(lats * lons * 2, 3, 3, 3)
    keep = np.ones((lats, lons, 2), dtype=bool)
    if skip_first_row_upper:
        keep[0, :, 0] = False
    if skip_last_row_lower:
        keep[lats - 1, :, 1] = False
    return np.ascontiguousarray(both[keep.reshape(-1)], dtype=np.float32)


def primitive_cylinder(lons=32, lats=4, rad=1, hei=2):
    """mesh/prim.py:50-87: the side as (lats x lons) quad cells, then the bottom fan (p[0, lon + 1], p[0, lon], centre) with
    normal (0, 0, -1) and the top fan (centre, p[lats, lon], p[lats, lon + 1]) with normal (0, 0, 1); the caps carry their
    normal as the texture coordinate too (sic)."""
    la, lo = np.meshgrid(np.arange(lats + 1, dtype=np.float64) / lats, np.arange(lons + 1, dtype=np.float64) / lons, indexing='ij')
    o = (lo * 2 - 1) * np.pi
    x, y, z = np.cos(o), np.sin(o), la - 0.5
    pts = np.stack([np.stack([x * rad, y * rad, z * hei], -1), np.stack([x, y, np.zeros_like(x)], -1),
                    np.stack([lo, la, np.zeros_like(la)], -1)], axis=-2)
    side = _quad_cells(pts, lats, lons)

    def fan(row, zc, nz, centre_first):
        n = np.array([0.0, 0.0, nz])
        ring = pts[row, :, 0]  # [lons + 1, 3] positions
        centre = np.broadcast_to(np.array([0.0, 0.0, zc]), (lons, 3))
        corners = [centre, ring[:-1], ring[1:]] if centre_first else [ring[1:], ring[:-1], centre]
        tri = np.stack(corners, axis=1)  # [lons, 3 corners, 3]
        attr = np.broadcast_to(n, tri.shape)
        return np.stack([tri, attr, attr], axis=2)
    return np.ascontiguousarray(np.concatenate([side, fan(0, -hei / 2, -1.0, False), fan(lats, hei / 2, 1.0, True)]), dtype=np.float32)


class PrimitiveMesh(SimpleMesh):
    """mesh/prim.py:5-101: a SimpleMesh filled from `faces` [N, 3 corners, 3 (vertex, normal, texcoord), 3]."""

    def __init__(self, faces):
        faces = np.asarray(faces, dtype=np.float32)
        if faces.ndim != 4 or faces.shape[1:] != (3, 3, 3):
            raise ValueError(f'PrimitiveMesh takes [N, 3, 3, 3] faces (vertex, normal, texcoord per corner), got {faces.shape}')
        super().__init__(maxfaces=len(faces), npolygon=3)
        self.set_face_verts(np.ascontiguousarray(faces[:, :, 0]))
        self.set_face_norms(np.ascontiguousarray(faces[:, :, 1]))
        self.set_face_coors(np.ascontiguousarray(faces[:, :, 2]))

    @classmethod
    def sphere(cls, lons=32, lats=24, rad=1):
        return cls(primitive_sphere(lons, lats, rad))

    @classmethod
    def cylinder(cls, lons=32, lats=4, rad=1, hei=2):
        return cls(primitive_cylinder(lons, lats, rad, hei))

    @classmethod
    def asset(cls, name):
        """prim.py:89-101 (quirk kept: the per-corner triples are built as (vertex, texcoord + [0], normal), so the constructor
        reads the texture coordinates as normals and the normals as texture coordinates)."""
        from .assimp import readobj
        obj = readobj('assets/' + name + '.obj', quadok=True)
        f = obj['f']
        if f.shape[1] != 3:
            raise NotImplementedError('only triangle assets are on the B200 raster path')
        verts, coors, norms = obj['v'][f[:, :, 0]], obj['vt'][f[:, :, 1]], obj['vn'][f[:, :, 2]]
        coors3 = np.concatenate([coors, np.zeros(coors.shape[:2] + (1,), coors.dtype)], axis=-1)
        return cls(np.stack([verts, coors3, norms], axis=2))


class ConnectiveMesh:
    """mesh/conn.py:5-103: vertices, normals and texture coordinates per VERTEX, faces as vertex-index triples."""

    def __init__(self, maxfaces=MAX, maxverts=MAX, npolygon=3):
        if npolygon != 3:
            raise NotImplementedError('only triangle meshes (npolygon=3) are on the B200 raster path')
        self.maxfaces, self.maxverts, self.npolygon = maxfaces, maxverts, npolygon
        self._v = np.zeros((0, 3), np.float32)
        self._vn = self._vt = None
        self._f = np.zeros((0, 3), np.int32)
        self._model = None

    def get_npolygon(self):
        return self.npolygon

    def get_nfaces(self):
        return len(self._f)

    def set_vertices(self, verts):  # conn.py:71-76
        self._v, self._model = np.ascontiguousarray(np.asarray(verts, np.float32)[:self.maxverts, :3]), None

    def set_vert_norms(self, norms):  # conn.py:78-83
        self._vn, self._model = np.ascontiguousarray(np.asarray(norms, np.float32)[:self.maxverts, :3]), None

    def set_vert_coors(self, coors):  # conn.py:85-90
        self._vt, self._model = np.ascontiguousarray(np.asarray(coors, np.float32)[:self.maxverts, :2]), None

    def set_faces(self, faces):  # conn.py:92-97
        self._f, self._model = np.ascontiguousarray(np.asarray(faces, np.int32)[:self.maxfaces, :3]), None

    def _source(self):
        if self._model is None:
            n = len(self._v)
            vn = self._vn if self._vn is not None else np.zeros((n, 3), np.float32)
            vt = self._vt if self._vt is not None else np.zeros((n, 2), np.float32)
            if len(vn) < n or len(vt) < n:
                raise ValueError('ConnectiveMesh: normals / texture coordinates must cover every vertex')
            self._model = MeshModel({'v': self._v, 'vn': vn, 'vt': vt, 'f': np.repeat(self._f[:, :, None], 3, axis=2)})
        return self._model._source()


class MeshModel:
    """mesh/model.py:5-73: indexed mesh from an OBJ dict {'v','vt','vn','f'} or a path."""

    def __init__(self, obj, *args, **kwargs):
        if isinstance(obj, str):
            from .assimp import readobj
            obj = readobj(obj, *args, **kwargs)
        dev = _device()
        faces = np.asarray(obj['f'])
        if faces.ndim == 2:  # model.py:23-24: same index for v / vt / vn
            faces = np.stack([faces, faces, faces], axis=2)
        v = np.asarray(obj['v'], dtype=np.float32)
        vt = np.asarray(obj.get('vt', np.zeros((1, 2))), dtype=np.float32)[:, :2]
        vn = np.asarray(obj.get('vn', np.zeros((1, 3))), dtype=np.float32)
        # The kernels index `faces` as int32 [N, 3 corners, 3 = v / vt / vn] without bounds checks: validate here.
        # (The reference fails in from_numpy on a shape mismatch, model.py:25-31; readobj yields [N,3,1] for `f 1 2 3`
        # and [N,3,2] for `f 1/1 2/2`: missing vt / vn columns are filled with index 0 like obj.py:73 does.)
        if faces.ndim != 3 or faces.shape[1] != 3 or not 1 <= faces.shape[2] <= 3:
            raise ValueError(f"MeshModel: 'f' must be [N,3] or [N,3,k<=3] vertex / texcoord / normal indices, got shape {faces.shape}")
        if faces.shape[2] < 3:
            faces = np.concatenate([faces, np.zeros(faces.shape[:2] + (3 - faces.shape[2],), faces.dtype)], axis=2)
        faces = faces.astype(np.int64)
        for col, name in ((1, 'vt'), (2, 'vn')):  # attribute absent from the dict: one placeholder element, index 0
            if name not in obj:
                faces[:, :, col] = 0
        for col, (name, arr) in enumerate((('v', v), ('vt', vt), ('vn', vn))):
            if len(faces) and (faces[:, :, col].min() < 0 or faces[:, :, col].max() >= max(len(arr), 1)):
                raise ValueError(f"MeshModel: '{name}' index out of range [0, {len(arr)}) in 'f' "
                                 f'(min {faces[:, :, col].min()}, max {faces[:, :, col].max()}; relative OBJ indices are not supported)')
        self.faces = torch.as_tensor(np.ascontiguousarray(faces.astype(np.int32))).to(dev)
        self.verts = torch.as_tensor(np.ascontiguousarray(v)).to(dev)
        self.coors = torch.as_tensor(np.ascontiguousarray(vt)).to(dev)
        self.norms = torch.as_tensor(np.ascontiguousarray(vn)).to(dev)
        self.maxfaces, self.maxverts = len(faces), len(v)
        self.maxcoors, self.maxnorms = len(vt), len(vn)
        self._host = {'f': faces, 'v': v}

    def get_npolygon(self):
        return 3

    def get_max_vert_nindex(self):
        return self.maxverts

    def get_nfaces(self):
        return self.maxfaces

    def _source(self):
        return FaceSource('indexed', v=self.verts, vt=self.coors, vn=self.norms, faces=self.faces, nfaces=self.maxfaces)


class MeshGrid:
    """mesh/grid.py:5-70: (nx, ny) vertex grid, two triangles per cell; normals are recomputed
    from `pos` every frame by the set_object kernel (grid.py:26-35)."""

    def __init__(self, res, as_quad=False):
        if as_quad:
            raise NotImplementedError('as_quad grids feed the wireframe raster, which is not on this path')
        if isinstance(res, int):
            res = res, res
        self.res = (int(res[0]), int(res[1]))
        nx, ny = self.res
        dev = _device()
        # grid.py:17-21 (f32 arithmetic: I / (res - 1), u * 2 - 1)
        u = (np.arange(nx, dtype=np.float32) / np.float32(nx - 1))[:, None] * np.ones((1, ny), np.float32)
        v = np.ones((nx, 1), np.float32) * (np.arange(ny, dtype=np.float32) / np.float32(ny - 1))[None, :]
        pos = np.stack([u * np.float32(2) - np.float32(1), v * np.float32(2) - np.float32(1), np.zeros_like(u)], axis=2)
        self._pos = torch.as_tensor(np.ascontiguousarray(pos, dtype=np.float32)).to(dev)
        self._tex = torch.as_tensor(np.ascontiguousarray(np.stack([u, v], axis=2), dtype=np.float32)).to(dev)
        self.as_quad = False

    @property
    def pos(self):
        """float32[nx, ny, 3] device field; edit through `.to_torch()` (in place) or `.from_numpy()`."""
        return Field(self._pos)

    @property
    def tex(self):
        return Field(self._tex)

    def get_npolygon(self):
        return 3

    def get_nfaces(self):
        return 2 * (self.res[0] - 1) * (self.res[1] - 1)

    def _source(self):
        return FaceSource('grid', pos=self._pos, nx=self.res[0], ny=self.res[1], nfaces=self.get_nfaces())


class MeshEditBase:
    def __init__(self, mesh):
        self.mesh = mesh

    def __getattr__(self, attr):
        return getattr(self.__dict__['mesh'], attr)

    def _source(self):
        return self.mesh._source()


class MeshTransform(MeshEditBase):
    """mesh/trans.py:5-40."""

    def __init__(self, mesh, trans=None):
        super().__init__(mesh)
        self.trans = np.eye(4, dtype=np.float32)
        self.trans_normal = np.eye(3, dtype=np.float32)
        if trans is not None:
            self.set_transform(trans)

    def set_transform(self, trans):
        trans = np.asarray(trans, dtype=np.float64)
        self.trans = trans.astype(np.float32)
        # trans.py:23-26: inverse transpose, upper-left 3x3 kept by the 3x3 field
        self.trans_normal = np.transpose(np.linalg.inv(trans))[:3, :3].astype(np.float32)

    def _source(self):
        s = self.mesh._source()
        if s.trans is None:
            s.trans, s.trans_normal = self.trans, self.trans_normal
        else:  # nested MeshTransform: keep the chain (innermost first); the reference rounds after every wrapper
            inner_t = s.trans if isinstance(s.trans, list) else [s.trans]
            inner_n = s.trans_normal if isinstance(s.trans_normal, list) else [s.trans_normal]
            if len(inner_t) >= 4:
                raise NotImplementedError('more than four nested MeshTransform wrappers')
            s.trans, s.trans_normal = inner_t + [self.trans], inner_n + [self.trans_normal]
        return s


class MeshFlipCulling(MeshEditBase):
    """mesh/cull.py:5-28: reversed winding."""

    def _source(self):
        s = self.mesh._source()
        # (around a MeshNoCulling the even copies end up reversed and the odd, normal-negated copies in the original
        # order -- cull.py:17-27 over :40-57 -- which is what mode bits 1 | 2 select in corner_ids)
        s.flip = not s.flip
        return s


class MeshNoCulling(MeshEditBase):
    """mesh/cull.py:31-57: every face twice, odd copies reversed with negated normals."""

    def get_nfaces(self):
        return self.mesh.get_nfaces() * 2

    def _source(self):
        s = self.mesh._source()
        if s.double_sided:
            raise NotImplementedError('nested MeshNoCulling is not supported')
        s.double_sided = True
        return s


class MeshFlipNormal(MeshEditBase):
    """mesh/cull.py:60-66."""

    def _source(self):
        s = self.mesh._source()
        s.negate = not s.negate
        return s


def _face_normals(v, f):
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    n = np.cross(b - a, c - a).astype(np.float32)
    ln = np.sqrt((n * n).sum(axis=1, keepdims=True)).astype(np.float32)
    return (np.float32(1) / ln) * n


class MeshFlatNormal(MeshEditBase):
    """mesh/norm.py:5-14: per-face normal for all three corners (MeshModel sources)."""

    def _source(self):
        s = self.mesh._source()
        if s.kind != 'indexed':
            raise NotImplementedError('MeshFlatNormal needs an indexed mesh')
        host = _model_host(self.mesh, 'MeshFlatNormal')
        fv = host['f'][:, :, 0]
        nrm = _face_normals(host['v'], fv)  # (from the mesh's CURRENT vertices: norm.py:5-14 computes per frame)
        faces = host['f'].copy()
        faces[:, :, 2] = np.arange(len(faces))[:, None]
        s.vn = torch.as_tensor(np.ascontiguousarray(nrm)).to(s.v.device)
        s.faces = torch.as_tensor(np.ascontiguousarray(faces.astype(np.int32))).to(s.v.device)
        return s


def _model_host(mesh, who):
    """Index buffer + the CURRENT vertex positions of a MeshModel as numpy (the normal adapters recompute from what
    mesh.verts holds now, like the reference's pre_compute, mesh/norm.py:28-55).  Wrapped meshes are not supported:
    their transform / flip would have to be applied before the normals are computed."""
    if not isinstance(mesh, MeshModel):
        raise NotImplementedError(f'{who} takes a MeshModel (got {type(mesh).__name__}); apply MeshTransform / culling wrappers outside it')
    return {'f': mesh._host['f'], 'v': mesh.verts.detach().cpu().numpy()}


class MeshSmoothNormal(MeshEditBase):
    """mesh/norm.py:17-55: area-unweighted average of adjacent face normals per vertex.
    (The reference accumulates with atomics in arbitrary order; this sums in face order.)"""

    def __init__(self, mesh, cached=True):
        super().__init__(mesh)
        self.cached = cached
        self._norm = None

    def update_normal(self):
        host = _model_host(self.mesh, 'MeshSmoothNormal')
        fv = host['f'][:, :, 0]
        fn = _face_normals(host['v'], fv)
        acc = np.zeros((len(host['v']), 3), dtype=np.float32)
        for k in range(3):
            np.add.at(acc, fv[:, k], fn)
        ln = np.sqrt((acc * acc).sum(axis=1, keepdims=True)).astype(np.float32)
        with np.errstate(divide='ignore', invalid='ignore'):
            self._norm = ((np.float32(1) / ln) * acc).astype(np.float32)

    def _source(self):
        s = self.mesh._source()
        if s.kind != 'indexed':
            raise NotImplementedError('MeshSmoothNormal needs an indexed mesh')
        if self._norm is None or not self.cached:
            self.update_normal()
        faces = self.mesh._host['f'].copy()
        faces[:, :, 2] = faces[:, :, 0]
        s.vn = torch.as_tensor(np.ascontiguousarray(self._norm)).to(s.v.device)
        s.faces = torch.as_tensor(np.ascontiguousarray(faces.astype(np.int32))).to(s.v.device)
        return s
