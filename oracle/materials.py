"""The oracle's OWN front-end for materials and lights -- TEST INFRASTRUCTURE ONLY.

An independent restatement of how the reference composes a material node graph into the three
functions `Lighting.shade_color` calls (tina/core/lighting.py:84-98):

    brdf(nrm, idir, odir), ambient(), emission()            tina/matr/material.py:5-57

flattened into the postfix program `oracle/tina_oracle.c` interprets (opcodes = the C ABI of
include/tina_b200.h).  Nothing here imports the product package: node graphs are walked by duck typing
on the reference's own class names (`MixMaterial`, `ScaleMaterial`, `AddMaterial`, `Lambert`, `Phong`,
`CookTorrance`, `Emission`, `Const`, `Param`, `Input`, `Texture`, `FresnelFactor`), so the same walker
accepts the product's node objects and the reference's (under oracle/ref_shim); and `stock_*` builds the
default materials without any node objects at all (bench.py --impl reference).
"""
import ctypes as C

import numpy as np

MAX_LIGHTS, MAX_INSTR, MAX_TEX = 16, 96, 4
(OP_CONST, OP_INPUT, OP_TEXTURE, OP_FRESNEL, OP_LAMBERT, OP_PHONG, OP_COOK, OP_MIX, OP_MUL, OP_ADD) = range(10)
OP_BCAST, OP_CHESS = 12, 13  # (10 / 11 are the product's register ops, unused by the unfolded programs here)
INPUTS = {'pos': 0, 'color': 1, 'normal': 2, 'texcoord': 3}  # matr/nodes.py:79-96


class LightingPOD(C.Structure):  # include/tina_b200.h: TinaLighting
    _fields_ = [('dirs', (C.c_float * 4) * MAX_LIGHTS), ('colors', (C.c_float * 4) * MAX_LIGHTS), ('ambient', C.c_float * 4),
                ('nlights', C.c_int32), ('pad', C.c_int32 * 3)]


class Instr(C.Structure):  # TinaInstr
    _fields_ = [('op', C.c_int32), ('arg', C.c_int32), ('c', C.c_float * 3)]


class MaterialPOD(C.Structure):  # TinaMaterial
    _fields_ = [('n_brdf', C.c_int32), ('n_ambient', C.c_int32), ('n_emission', C.c_int32), ('ntex', C.c_int32),
                ('tex', C.c_void_p * MAX_TEX), ('tex_w', C.c_int32 * MAX_TEX), ('tex_h', C.c_int32 * MAX_TEX),
                ('tex_c', C.c_int32 * MAX_TEX), ('n_prologue', C.c_int32), ('prologue_form', C.c_int32), ('pad_', C.c_int32 * 2),
                ('code', Instr * MAX_INSTR)]


# ---- lights (core/lighting.py:33-69) ------------------------------------------------------------
def lighting_pod(lights, ambient):
    """lights: [(xyzw, rgb), ...] with xyzw as the reference stores it after set_light (:55-66: a direction
    normalised in f64 with w = 0, or a position with w = 1); ambient: rgb."""
    L = LightingPOD()
    L.nlights = len(lights)
    for i, (d, c) in enumerate(lights):
        d32, c32 = np.asarray(d, dtype=np.float32), np.asarray(c, dtype=np.float32)
        for k in range(4):
            L.dirs[i][k] = float(d32[k])
        for k in range(3):
            L.colors[i][k] = float(c32[k])
    a32 = np.asarray(ambient, dtype=np.float32)
    for k in range(3):
        L.ambient[k] = float(a32[k])
    return L


def directional(dir, color=(1, 1, 1)):
    """lighting.py:60-62: directions are normalised on the host in f64."""
    d = np.asarray(dir, dtype=np.float64)
    return np.append(d / np.linalg.norm(d), 0.0), color


def lighting_of(lighting):
    """A Lighting object (the product's, or the reference's under ref_shim) -> LightingPOD, read by attribute."""
    def arr(x):
        return np.asarray(x.to_numpy() if hasattr(x, 'to_numpy') else x)
    nl = lighting.nlights
    n = int(np.asarray(nl[None] if hasattr(nl, '__getitem__') else nl))
    dirs, cols = arr(lighting.light_dirs), arr(lighting.light_colors)
    return lighting_pod([(dirs[i], cols[i]) for i in range(n)], arr(lighting.ambient_color).reshape(-1)[:3])


def default_lighting():
    """scene/raster.py:90-93 as used by every parity test: one directional light (1,2,3) x 0.9, ambient 0.1."""
    return lighting_pod([directional([1, 2, 3], [0.9, 0.9, 0.9])], [0.1, 0.1, 0.1])


# ---- materials ----------------------------------------------------------------------------------
class _Prog:
    def __init__(self):
        self.code, self.textures = [], []

    def const(self, v):
        a = np.asarray(v, dtype=np.float64).reshape(-1)
        a = np.repeat(a, 3) if a.size == 1 else (np.append(a, 0.0) if a.size == 2 else a)
        assert a.size == 3, v
        self.code.append((OP_CONST, 0, tuple(float(x) for x in a)))

    def op(self, op, arg=0):
        self.code.append((op, int(arg), (0.0, 0.0, 0.0)))

    def tex(self, image):
        for i, t in enumerate(self.textures):
            if t is image:
                return i
        assert len(self.textures) < MAX_TEX
        self.textures.append(image)
        return len(self.textures) - 1


def _kind(node):
    return type(node).__name__


def _param(node, key):
    return node.params[key] if hasattr(node, 'params') else node.param(key)


def _value(P, node):
    """parameter nodes: matr/nodes.py:42-111, matr/material.py:69-83"""
    k = _kind(node)
    if k == 'Param':
        P.const(node[None])
    elif k == 'Const':
        P.const(node.value)
    elif k == 'Input':
        P.op(OP_INPUT, INPUTS[node.name])
    elif k == 'Texture':
        _value(P, _param(node, 'texcoord'))
        img = node.image if hasattr(node, 'image') else node.texture
        P.op(OP_TEXTURE, P.tex(img))
    elif k == 'FresnelFactor':  # pushes metallic, albedo, specular; pops them in reverse
        for key in ('metallic', 'albedo', 'specular'):
            _value(P, _param(node, key))
        P.op(OP_FRESNEL)
    elif k == 'ChessboardTexture':  # nodes.py:114-126: lerp((texcoord // size).sum() % 2, color0, color1)
        _value(P, _param(node, 'texcoord'))
        _value(P, _param(node, 'size'))
        P.op(OP_CHESS)
        _value(P, _param(node, 'color0'))
        _value(P, _param(node, 'color1'))
        P.op(OP_MIX)
    elif k == 'LerpTexture':  # nodes.py:129-136: lerp(uv.x, x0, x1) + lerp(uv.y, x0, x1)
        for comp in (0, 1):
            _value(P, _param(node, 'texcoord'))
            P.op(OP_BCAST, comp)
            _value(P, _param(node, 'x0'))
            _value(P, _param(node, 'x1'))
            P.op(OP_MIX)
        P.op(OP_ADD)
    else:
        raise NotImplementedError(k)


# what each leaf contributes to (brdf, ambient, emission): material.py:387-396, 445-457, 243-362, 659-676
_LEAF = {'Lambert': (OP_LAMBERT, 1.0, 0.0), 'Phong': (OP_PHONG, 1.0, 0.0), 'CookTorrance': (OP_COOK, 1.0, 0.0),
         'Emission': (None, 0.0, 1.0)}


def _material(P, m, what):
    k = _kind(m)
    if k == 'MixMaterial':      # material.py:96-118: (1 - fac) * mat1.X + fac * mat2.X
        _value(P, _param(m, 'factor'))
        _material(P, m.mat1, what)
        _material(P, m.mat2, what)
        P.op(OP_MIX)
    elif k == 'ScaleMaterial':  # :157-176: fac * mat.X
        _value(P, _param(m, 'factor'))
        _material(P, m.mat, what)
        P.op(OP_MUL)
    elif k == 'AddMaterial':    # :204-220: mat1.X + mat2.X
        _material(P, m.mat1, what)
        _material(P, m.mat2, what)
        P.op(OP_ADD)
    elif k in _LEAF:
        op, amb, emi = _LEAF[k]
        if what != 'brdf':
            P.const(amb if what == 'ambient' else emi)
        elif op is None:
            P.const(0.0)
        else:
            if k == 'Phong':
                _value(P, _param(m, 'shineness'))
            elif k == 'CookTorrance':
                _value(P, _param(m, 'roughness'))
                _value(P, _param(m, 'fresnel'))
            P.op(op)
    else:
        raise NotImplementedError(k)


def material_pod_of(material):
    """node graph -> (MaterialPOD, texture pointer table, keep-alive list)"""
    P = _Prog()
    parts = []
    for what in ('brdf', 'ambient', 'emission'):
        P.code = []
        _material(P, material, what)
        parts.append(P.code)
    return _pod(parts, P.textures)


def _pod(parts, textures=()):
    m = MaterialPOD()
    m.n_brdf, m.n_ambient, m.n_emission, m.ntex, m.n_prologue = len(parts[0]), len(parts[1]), len(parts[2]), len(textures), 0
    code = parts[0] + parts[1] + parts[2]
    assert len(code) <= MAX_INSTR
    for i, (op, arg, c) in enumerate(code):
        m.code[i].op, m.code[i].arg = op, arg
        m.code[i].c[0], m.code[i].c[1], m.code[i].c[2] = c
    arrays = []
    for t in textures:
        a = np.ascontiguousarray(t, dtype=np.float32)
        if a.ndim == 2:
            a = a[:, :, None]
        arrays.append(np.ascontiguousarray(a[:, :, :3] if a.shape[2] == 4 else a))
    ptrs = (C.c_void_p * max(1, len(arrays)))()
    for i, a in enumerate(arrays):
        ptrs[i] = a.ctypes.data
        m.tex_w[i], m.tex_h[i], m.tex_c[i] = a.shape
    return m, ptrs, arrays


def _c(v):
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    a = np.repeat(a, 3) if a.size == 1 else a
    return (OP_CONST, 0, tuple(float(x) for x in a))


_O = lambda op: (op, 0, (0.0, 0.0, 0.0))  # noqa: E731
_COLOR = (OP_INPUT, INPUTS['color'], (0.0, 0.0, 0.0))


def stock_diffuse(color=None):
    """tina.Diffuse(color='color') = Lambert() * color (material.py:691-693), without node objects."""
    col = _COLOR if color is None else _c(color)
    return _pod([[col, _O(OP_LAMBERT), _O(OP_MUL)], [col, _c(1.0), _O(OP_MUL)], [col, _c(0.0), _O(OP_MUL)]])


def stock_classic(color=None, shineness=32, specular=0.4):
    """tina.Classic = MixMaterial(Lambert() * color, Phong(shineness), specular) (material.py:684-688)."""
    col = _COLOR if color is None else _c(color)

    def part(diff, spec):
        return [_c(specular), col] + diff + [_O(OP_MUL)] + spec + [_O(OP_MIX)]
    return _pod([part([_O(OP_LAMBERT)], [_c(shineness), _O(OP_PHONG)]), part([_c(1.0)], [_c(1.0)]), part([_c(0.0)], [_c(0.0)])])


# ---- material.sample() trees for SSR (postp/ssr.py:78-80; matr/material.py sample methods) ------------------------
(SN_LAMBERT, SN_PHONG, SN_COOK, SN_EMISSION, SN_MIX, SN_SCALE, SN_ADD) = range(7)
SAMPLE_MAX_NODES, SAMPLE_MAX_INSTR = 16, 48


class SampleNode(C.Structure):  # include/tina_b200.h: TinaSampleNode
    _fields_ = [('kind', C.c_int32), ('a', C.c_int32), ('b', C.c_int32), ('p0', C.c_int32), ('n0', C.c_int32),
                ('p1', C.c_int32), ('n1', C.c_int32), ('pad_', C.c_int32)]


class SampleMaterialPOD(C.Structure):  # TinaSampleMaterial
    _fields_ = [('nnodes', C.c_int32), ('ncode', C.c_int32), ('ntex', C.c_int32), ('pad_', C.c_int32),
                ('tex', C.c_void_p * MAX_TEX), ('tex_w', C.c_int32 * MAX_TEX), ('tex_h', C.c_int32 * MAX_TEX),
                ('tex_c', C.c_int32 * MAX_TEX), ('nodes', SampleNode * SAMPLE_MAX_NODES), ('code', Instr * SAMPLE_MAX_INSTR)]


def _is_scalar(node):
    """Does the reference evaluate this parameter node to a scalar (Vavg leaves it alone, common.py:36-40)?"""
    k = _kind(node)
    if k == 'Const':
        return np.ndim(node.value) == 0 or np.size(node.value) == 1
    if k == 'Param':
        return np.ndim(node[None]) == 0
    if k == 'Texture':
        img = node.image if hasattr(node, 'image') else node.texture
        shape = img.shape if hasattr(img, 'shape') else ()
        return len(shape) == 2 or (len(shape) == 3 and shape[2] == 1)
    if k == 'FresnelFactor':
        return all(_is_scalar(_param(node, key)) for key in ('metallic', 'albedo', 'specular'))
    if k == 'ChessboardTexture':
        return all(_is_scalar(_param(node, key)) for key in ('color0', 'color1'))
    if k == 'LerpTexture':
        return all(_is_scalar(_param(node, key)) for key in ('x0', 'x1'))
    return False  # Input: vectors


def sample_pod_of(material):
    """node graph -> (SampleMaterialPOD, host texture arrays): the tree material.sample() descends (node 0 = root)."""
    P = _Prog()
    nodes = []

    def value(node):
        start = len(P.code)
        _value(P, node)
        return start, len(P.code) - start

    def walk(m):
        k = _kind(m)
        idx = len(nodes)
        rec = dict(kind=0, a=0, b=0, p0=0, n0=0, p1=0, n1=0, pad_=0)
        nodes.append(rec)
        if k == 'MixMaterial':
            f = _param(m, 'factor')
            rec['kind'] = SN_MIX
            rec['p0'], rec['n0'] = value(f)
            rec['pad_'] = 1 if _is_scalar(f) else 0
            rec['a'], rec['b'] = walk(m.mat1), walk(m.mat2)
        elif k == 'ScaleMaterial':
            rec['kind'] = SN_SCALE
            rec['p0'], rec['n0'] = value(_param(m, 'factor'))
            rec['a'] = walk(m.mat)
        elif k == 'AddMaterial':
            rec['kind'] = SN_ADD
            rec['a'], rec['b'] = walk(m.mat1), walk(m.mat2)
        elif k == 'Lambert':
            rec['kind'] = SN_LAMBERT
        elif k == 'Phong':
            rec['kind'] = SN_PHONG
            rec['p0'], rec['n0'] = value(_param(m, 'shineness'))
        elif k == 'CookTorrance':
            rec['kind'] = SN_COOK
            rec['p0'], rec['n0'] = value(_param(m, 'roughness'))
            rec['p1'], rec['n1'] = value(_param(m, 'fresnel'))
        elif k == 'Emission':
            rec['kind'] = SN_EMISSION
        else:
            raise NotImplementedError(k)
        return idx

    walk(material)
    assert len(nodes) <= SAMPLE_MAX_NODES and len(P.code) <= SAMPLE_MAX_INSTR
    pod = SampleMaterialPOD()
    pod.nnodes, pod.ncode, pod.ntex = len(nodes), len(P.code), len(P.textures)
    for i, rec in enumerate(nodes):
        for key, v in rec.items():
            setattr(pod.nodes[i], key, v)
    for i, (op, arg, c) in enumerate(P.code):
        pod.code[i].op, pod.code[i].arg = op, arg
        pod.code[i].c[0], pod.code[i].c[1], pod.code[i].c[2] = c
    arrays = []
    for t in P.textures:
        a = np.ascontiguousarray(t.to_numpy() if hasattr(t, 'to_numpy') else t, dtype=np.float32)
        if a.ndim == 2:
            a = a[:, :, None]
        arrays.append(np.ascontiguousarray(a[:, :, :3] if a.shape[2] == 4 else a))
    for i, a in enumerate(arrays):
        pod.tex[i] = a.ctypes.data
        pod.tex_w[i], pod.tex_h[i], pod.tex_c[i] = a.shape
    return pod, arrays
