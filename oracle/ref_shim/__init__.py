"""ref_shim -- TEST INFRASTRUCTURE ONLY.

An f32 NumPy emulation of the tiny part of the Taichi 0.7 runtime that the reference's
triangle-raster path touches, so that the reference's OWN, UNMODIFIED Python sources
(/root/reference/tina/...: core/engine.py, core/triangle.py, core/shader.py, core/lighting.py,
matr/*.py, mesh/*.py, scene/raster.py, postp/tonemap.py, common.py, advans.py, assimp/*.py,
util/matrix.py) can be executed here, serially, to produce golden vectors
(tests/golden/make_golden.py).  `taichi` itself is not installable in this image.

Semantics emulated (Taichi defaults: default_fp = f32, default_ip = i32):
  * every @ti.kernel / @ti.func body runs as plain Python over np.float32 / np.int32 values,
    one IEEE rounding per op, no contraction; Python-scope constant arithmetic (e.g.
    `1 / ti.pi`, Const parameters) stays in f64 until it meets a runtime value, as in Taichi;
  * parallel for-loops run serially in index order (= the deterministic outcome the CPU
    oracle restates: lowest face id first);
  * int(f32) follows x86 cvttss2si (out of range / NaN -> INT_MIN), like Taichi's CPU backend;
  * reading a dense field out of bounds yields 0 (the reference's bilerp reads one texel past
    the end with weight 0, common.py:140-149);
  * @ti.func tuple returns become lists (Taichi's expr_init_func), which mesh/trans.py relies on.
"""
import builtins
import importlib
import itertools
import math
import os
import sys
import types

import numpy as np

F32, I32 = np.float32, np.int32


def _f2i(x):
    x = np.asarray(x)
    if x.dtype.kind in 'iub':
        return x.astype(I32)
    with np.errstate(all='ignore'):
        ok = (x >= -2147483648.0) & (x < 2147483648.0)
        return np.where(ok, np.trunc(np.where(ok, x, 0)), -2147483648.0).astype(np.int64).astype(I32)


def _is_int(a):
    return np.asarray(a).dtype.kind in 'iub'


class Matrix:
    """ti.Matrix / ti.Vector: n (x m) entries in a numpy array (float32 or int32)."""
    __array_priority__ = 1000

    def __init__(self, arr, dt=None, _raw=False):
        if _raw:
            self.a = arr
            return
        if isinstance(arr, Matrix):
            arr = arr.a
        if isinstance(arr, (list, tuple)):
            arr = [x.a if isinstance(x, Matrix) else x for x in arr]
            if len(arr) and isinstance(arr[0], (list, tuple)):
                arr = [[y.a if isinstance(y, Matrix) else y for y in x] for x in arr]
        a = np.array(arr)
        if dt is not None:
            a = a.astype(F32 if _dtype(dt) == F32 else I32)
        elif a.dtype.kind == 'f':
            a = a.astype(F32)
        elif a.dtype.kind in 'iu':
            a = a.astype(I32)
        self.a = a

    # ---- structure ----
    @property
    def n(self):
        return self.a.shape[0]

    @property
    def m(self):
        return self.a.shape[1] if self.a.ndim > 1 else 1

    @property
    def entries(self):
        return [_scalar(x) for x in self.a.reshape(-1)]

    def __len__(self):
        return self.a.shape[0]

    def __iter__(self):
        for i in range(self.a.shape[0]):
            yield _scalar(self.a[i]) if self.a.ndim == 1 else Matrix(self.a[i], _raw=True)

    def __getitem__(self, idx):
        if isinstance(idx, Matrix):
            idx = tuple(int(i) for i in idx.a)
        r = self.a[idx]
        return _scalar(r) if np.ndim(r) == 0 else Matrix(r, _raw=True)

    def __setitem__(self, idx, v):
        if isinstance(idx, Matrix):
            idx = tuple(int(i) for i in idx.a)
        self.a[idx] = v.a if isinstance(v, Matrix) else v

    def __call__(self, *idx):
        return self[idx if len(idx) > 1 else idx[0]]

    def _set_component(self, i, v):
        # `v = field[i]` in kernel scope is a COPY in Taichi (postp/ssao.py:84-88 then assigns v.x, v.y): a matrix
        # handed out by Field.__getitem__ stays a live view only for subscript stores (field[None][2, 2] = -1 in
        # Python scope, core/engine.py:24); component assignment detaches it first
        if getattr(self, '_fview', False):
            self.a = self.a.copy()
            self._fview = False
        self.__setitem__(i, v)

    x = property(lambda s: s[0], lambda s, v: s._set_component(0, v))
    y = property(lambda s: s[1], lambda s, v: s._set_component(1, v))
    z = property(lambda s: s[2], lambda s, v: s._set_component(2, v))
    w = property(lambda s: s[3], lambda s, v: s._set_component(3, v))

    def __repr__(self):
        return f'Matrix({self.a.tolist()})'

    def __bool__(self):
        return bool(np.all(self.a))

    # ---- elementwise arithmetic ----
    def _bin(self, other, op, rev=False):
        b = other.a if isinstance(other, Matrix) else other
        a = self.a
        if isinstance(b, (np.floating,)) and not isinstance(b, np.float32):
            b = builtins.float(b)  # strong f64 numpy scalars never appear in Taichi code: treat as literals
        with np.errstate(all='ignore'):
            r = op(b, a) if rev else op(a, b)
        if r.dtype == np.float64:
            r = r.astype(F32)
        elif r.dtype.kind in 'iu' and r.dtype != I32:
            r = r.astype(I32)
        return Matrix(r, _raw=True)

    def __add__(s, o): return s._bin(o, np.add)
    def __radd__(s, o): return s._bin(o, np.add, True)
    def __sub__(s, o): return s._bin(o, np.subtract)
    def __rsub__(s, o): return s._bin(o, np.subtract, True)
    def __mul__(s, o): return s._bin(o, np.multiply)
    def __rmul__(s, o): return s._bin(o, np.multiply, True)

    @staticmethod
    def _tdiv(a, b):
        if _is_int(a) and _is_int(b):  # Taichi: int / int is a true division in default_fp
            a, b = np.asarray(a).astype(F32), np.asarray(b).astype(F32)
        return np.true_divide(a, b)

    def __truediv__(s, o): return s._bin(o, Matrix._tdiv)
    def __rtruediv__(s, o): return s._bin(o, Matrix._tdiv, True)
    @staticmethod
    def _fdiv(a, b):
        # Taichi lowers float `a // b` to floor(a / b) on the ROUNDED f32 quotient (numpy's floor_divide floors the exact one:
        # 1.0f // 0.2f is 5 in Taichi, 4 in numpy); integers divide as in Python
        if _is_int(a) and _is_int(b):
            return np.floor_divide(a, b)
        return np.floor(np.true_divide(np.asarray(a, dtype=F32), np.asarray(b, dtype=F32)))

    @staticmethod
    def _fmod(a, b):
        if _is_int(a) and _is_int(b):
            return np.mod(a, b)
        a, b = np.asarray(a, dtype=F32), np.asarray(b, dtype=F32)
        return a - b * Matrix._fdiv(a, b)  # Taichi: a % b = a - b * (a // b)

    def __floordiv__(s, o): return s._bin(o, Matrix._fdiv)
    def __rfloordiv__(s, o): return s._bin(o, Matrix._fdiv, True)
    def __mod__(s, o): return s._bin(o, Matrix._fmod)
    def __pow__(s, o): return s._bin(o, np.power)
    def __rpow__(s, o): return s._bin(o, np.power, True)
    def __neg__(s): return Matrix(-s.a, _raw=True)
    def __pos__(s): return s
    def __abs__(s): return Matrix(np.abs(s.a), _raw=True)
    def __lt__(s, o): return s._bin(o, np.less)
    def __le__(s, o): return s._bin(o, np.less_equal)
    def __gt__(s, o): return s._bin(o, np.greater)
    def __ge__(s, o): return s._bin(o, np.greater_equal)
    def __eq__(s, o): return s._bin(o, np.equal)
    def __ne__(s, o): return s._bin(o, np.not_equal)
    __hash__ = None

    def __matmul__(s, o):
        # taichi Matrix.__matmul__: acc = a(i,0)*b(0,j); acc = acc + a(i,k)*b(k,j) for k = 1..
        b = o.a if isinstance(o, Matrix) else np.asarray(o)
        vec = b.ndim == 1
        b2 = b[:, None] if vec else b
        out = np.zeros((s.a.shape[0], b2.shape[1]), dtype=np.result_type(s.a.dtype, b2.dtype))
        for i in range(s.a.shape[0]):
            for j in range(b2.shape[1]):
                acc = s.a[i, 0] * b2[0, j]
                for k in range(1, b2.shape[0]):
                    acc = acc + s.a[i, k] * b2[k, j]
                out[i, j] = acc
        return Matrix(out[:, 0] if vec else out, _raw=True)

    # ---- reductions / vector ops (taichi/lang/matrix.py) ----
    def sum(s):
        r = s.a.reshape(-1)
        acc = r[0]
        for v in r[1:]:
            acc = acc + v
        return _scalar(acc)

    def all(s): return bool(np.all(s.a))
    def any(s): return bool(np.any(s.a))
    def dot(s, o): return (s * o).sum()
    def norm_sqr(s): return (s ** 2).sum()

    def norm(s, eps=0):
        return np.sqrt(s.norm_sqr() + eps) if eps else np.sqrt(s.norm_sqr())

    def normalized(s, eps=0):
        invlen = F32(1) / (s.norm() + eps) if eps else F32(1) / s.norm()
        return invlen * s

    def cross(s, o):
        a, b = s, o
        if len(a) == 2:
            return a[0] * b[1] - a[1] * b[0]
        return Matrix([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    def transpose(s): return Matrix(s.a.T.copy(), _raw=True)
    def to_numpy(s): return s.a.copy()

    # ---- constructors ----
    @staticmethod
    def identity(dt, n): return Matrix(np.eye(n), dt)
    @staticmethod
    def zero(dt, n, m=None): return Matrix(np.zeros((n, m) if m else n), dt)
    @staticmethod
    def unit(n, i, dt=None): return Matrix(np.eye(n)[i], dt or builtins.int)
    @staticmethod
    def cols(cols): return Matrix(np.stack([c.a for c in cols], axis=1), _raw=True)
    @staticmethod
    def rows(rows): return Matrix(np.stack([c.a for c in rows], axis=0), _raw=True)

    @staticmethod
    def field(n, m, dtype, shape=None, **kw):
        return Field(dtype, shape, (n, m))


def Vector(arr, dt=None, **kw):
    return Matrix(arr, dt)


Vector.field = lambda n, dtype, shape=None, **kw: Field(dtype, shape, (n,))
Vector.unit = Matrix.unit
Vector.zero = Matrix.zero


class U32:
    """ti.u32 scalar (tina/random.py: Wang hash): arithmetic modulo 2^32 with ints, f32 arithmetic with floats
    (Taichi promotes u32 (op) f32 to f32: the integer is first rounded to f32)."""
    __array_priority__ = 2000

    def __init__(self, v):
        self.v = builtins.int(v) & 0xffffffff

    @staticmethod
    def _int(o):
        if isinstance(o, U32):
            return o.v
        if isinstance(o, (builtins.int, np.integer, IntRef)) and not isinstance(o, builtins.bool):
            return builtins.int(o) & 0xffffffff
        return None

    def _bin(self, o, fi, ff, rev=False):
        i = U32._int(o)
        if i is not None:
            return U32(fi(i, self.v) if rev else fi(self.v, i))
        a, b = F32(self.v), F32(o)
        return F32(ff(b, a) if rev else ff(a, b))

    def __xor__(self, o): return self._bin(o, lambda a, b: a ^ b, None)
    def __rxor__(self, o): return self._bin(o, lambda a, b: a ^ b, None, True)
    def __and__(self, o): return self._bin(o, lambda a, b: a & b, None)
    def __rand__(self, o): return self._bin(o, lambda a, b: a & b, None, True)
    def __or__(self, o): return self._bin(o, lambda a, b: a | b, None)
    def __ror__(self, o): return self._bin(o, lambda a, b: a | b, None, True)
    def __lshift__(self, o): return U32(self.v << builtins.int(o))
    def __rshift__(self, o): return U32(self.v >> builtins.int(o))
    def __add__(self, o): return self._bin(o, lambda a, b: a + b, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, lambda a, b: a - b)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, lambda a, b: a * b, True)
    def __mod__(self, o): return self._bin(o, lambda a, b: a % b, None)
    def __truediv__(self, o): return F32(F32(self.v) / F32(builtins.int(o) if U32._int(o) is not None else o))
    def __eq__(self, o): return self.v == (U32._int(o) if U32._int(o) is not None else o)
    def __ne__(self, o): return not self.__eq__(o)
    def __lt__(self, o): return self.v < (U32._int(o) if U32._int(o) is not None else o)
    def __hash__(self): return hash(self.v)
    def __int__(self): return self.v
    def __index__(self): return self.v
    def __float__(self): return builtins.float(self.v)
    def __repr__(self): return f'U32({self.v})'


def _scalar(x):
    x = np.asarray(x)
    if x.dtype == np.float64:
        return F32(x)
    if x.dtype.kind in 'iu' and x.dtype != I32:
        return I32(x)
    return x[()]


class TiInt(builtins.int):
    """A runtime i32 value: Taichi's `/` on ints is a true division in default_fp (f32), not Python's f64
    (e.g. `i / siz` in wireframe.py:66, `1 / count` in accumator.py:18)."""

    def __truediv__(self, o):
        return F32(builtins.int(self)) / (F32(o) if isinstance(o, (builtins.int, np.integer)) else o)

    def __rtruediv__(self, o):
        return (F32(o) if isinstance(o, (builtins.int, np.integer)) else o) / F32(builtins.int(self))


class IntRef(TiInt):
    """Value of an int scalar field element that remembers where it lives (lvalue for ti.atomic_*)."""
    def __new__(cls, value, field, idx):
        o = builtins.int.__new__(cls, value)
        o.field, o.idx = field, idx
        return o


def _dtype(dt):
    if dt in (builtins.float, F32, 'f32') or getattr(dt, '_shim_kind', None) == 'f':
        return F32
    if dt in (builtins.int, I32, 'i32') or getattr(dt, '_shim_kind', None) == 'i':
        return I32
    if dt in (np.float64, 'f64'):
        return np.float64
    raise TypeError(dt)


def _shape(shape):
    if shape is None:
        return None
    if isinstance(shape, Matrix):
        return tuple(int(v) for v in shape.a)
    if isinstance(shape, (builtins.int, np.integer)):
        return (int(shape),)
    return tuple(int(v) for v in shape)


class Field:
    def __init__(self, dtype, shape, elem=()):
        self.dt = _dtype(dtype)
        self.elem = tuple(elem)
        self.shape = _shape(shape)
        self.data = np.zeros(self.shape + self.elem, dtype=self.dt)

    @property
    def n(self):
        return self.elem[0] if self.elem else 1

    def _idx(self, idx):
        if idx is None:
            return ()
        if isinstance(idx, Matrix):
            return tuple(int(i) for i in idx.a)
        if isinstance(idx, tuple):
            out = []
            for i in idx:
                out.extend(self._idx(i)) if isinstance(i, Matrix) else out.append(int(i))
            return tuple(out)
        return (int(idx),)

    def __getitem__(self, idx):
        idx = self._idx(idx)
        if any(i < 0 or i >= s for i, s in zip(idx, self.shape)):
            z = np.zeros(self.elem, dtype=self.dt)  # out-of-bounds read of a dense field
            return Matrix(z, _raw=True) if self.elem else z[()]
        if self.elem:
            m = Matrix(self.data[idx], _raw=True)  # live view: field[None][2, 2] = -1 works
            m._fview = True
            return m
        v = self.data[idx]
        return IntRef(v, self, idx) if self.dt == I32 else v

    def __setitem__(self, idx, v):
        idx = self._idx(idx)
        if isinstance(v, Matrix):
            v = v.a
        v = np.asarray(v)
        if self.elem and v.shape != self.elem and v.ndim == len(self.elem):
            v = v[tuple(slice(0, e) for e in self.elem)]  # 4x4 list into a 3x3 field (mesh/trans.py:24-26)
        if self.dt == I32 and v.dtype.kind == 'f':
            v = _f2i(v)
        self.data[idx] = v

    def __iter__(self):  # struct-for
        for idx in itertools.product(*[range(s) for s in self.shape]):
            yield idx[0] if len(idx) == 1 else idx

    def fill(self, v):
        self.data[...] = v.a if isinstance(v, Matrix) else v

    def from_numpy(self, arr):
        self.data[...] = np.asarray(arr).reshape(self.data.shape)

    def to_numpy(self):
        return self.data.copy()

    def copy_from(self, other):
        self.data[...] = other.data


# ---- cast classes that replace the builtins inside the reference's modules ----------------------
class _CastMeta(type):
    def __instancecheck__(cls, obj):
        return isinstance(obj, cls._builtin)

    def __call__(cls, x=0, *a):
        return cls._cast(x)


def _to_float(x):
    if isinstance(x, Matrix):
        return Matrix(x.a.astype(F32), _raw=True)
    if isinstance(x, (np.generic, IntRef)):
        return F32(x)
    return builtins.float(x)


def _to_int(x):
    if isinstance(x, Matrix):
        return Matrix(_f2i(x.a), _raw=True)
    if isinstance(x, np.floating):
        return TiInt(_f2i(x))
    return builtins.int(x)


class shim_float(metaclass=_CastMeta):
    _builtin, _cast, _shim_kind = builtins.float, staticmethod(_to_float), 'f'
    dtype = np.dtype(np.float64)  # so that host code's np.array(x, dtype=float) keeps working


class shim_int(metaclass=_CastMeta):
    _builtin, _cast, _shim_kind = builtins.int, staticmethod(_to_int), 'i'
    dtype = np.dtype(np.int64)


def _elementwise2(fn_scalar, np_fn):
    def f(*args):
        acc = args[0]
        for b in args[1:]:
            if isinstance(acc, Matrix) or isinstance(b, Matrix):
                A = acc if isinstance(acc, Matrix) else Matrix(np.asarray(acc))
                acc = A._bin(b, np_fn)
            else:
                acc = fn_scalar(acc, b)
        return acc
    return f


# llvm minnum / maxnum for floats: the non-NaN operand wins
shim_min = _elementwise2(lambda a, b: b if (b < a or a != a) else a, np.fmin)
shim_max = _elementwise2(lambda a, b: b if (b > a or a != a) else a, np.fmax)


def shim_abs(x):
    return abs(x)


def _func(f):
    """@ti.func / @ti.pyfunc: tuple returns become lists of fresh values (expr_init_func)."""
    import functools
    import inspect
    if inspect.isgeneratorfunction(f):
        return f

    @functools.wraps(f)
    def wrapped(*a, **k):
        r = f(*a, **k)
        if isinstance(r, tuple):
            return [Matrix(x.a.copy(), _raw=True) if isinstance(x, Matrix) else x for x in r]
        return r
    return wrapped


def _unary(np_fn):
    def f(x):
        with np.errstate(all='ignore'):
            if isinstance(x, Matrix):
                return Matrix(np_fn(x.a), _raw=True)
            if isinstance(x, (builtins.int, builtins.float)) and not isinstance(x, IntRef):
                return builtins.float(np_fn(x))
            return _scalar(np_fn(F32(x)))
    return f


def make_taichi():
    ti = types.ModuleType('taichi')
    ti.__path__ = []
    ti._tinahacked = 1  # tina/hacker.py:4: its monkey patches target real Taichi internals; skipped
    ti.smart = lambda x: x  # tina/hacker.py:57: generator-for, plain iteration here
    ti.Matrix, ti.Vector = Matrix, Vector
    ti.field = lambda dtype, shape=None, **kw: Field(dtype, shape)
    ti.kernel = lambda f: f
    ti.func = _func
    ti.pyfunc = _func
    ti.data_oriented = lambda c: c
    ti.template = lambda: None
    ti.ext_arr = lambda: None
    ti.static = lambda x, *xs: [x] + list(xs) if xs else x  # tina/hacker.py:5-6
    ti.static_assert = lambda *a, **k: None
    ti.inside_kernel = lambda: True
    ti.pi, ti.tau = math.pi, math.tau
    ti.f32, ti.i32, ti.f64 = F32, I32, np.float64
    ti.cpu, ti.gpu, ti.cuda, ti.opengl = 'cpu', 'gpu', 'cuda', 'opengl'
    ti.init = lambda *a, **k: None
    ti.GUI = type('GUI', (), {'__init__': lambda self, *a, **k: None, 'show': lambda self, *a, **k: None})

    def materialize_callback(f):  # fields exist immediately: run now
        f()
        return f
    ti.materialize_callback = materialize_callback
    ti.grouped = lambda it: (Matrix(np.array(i if isinstance(i, tuple) else (i,), dtype=I32), _raw=True) for i in it)
    ti.ndrange = lambda *rs: itertools.product(*[range(int(r[0]), int(r[1])) if isinstance(r, tuple) else range(int(r)) for r in rs])
    ti.sqrt, ti.floor, ti.ceil = _unary(np.sqrt), _unary(np.floor), _unary(np.ceil)
    ti.sin, ti.cos, ti.exp, ti.log, ti.tan = _unary(np.sin), _unary(np.cos), _unary(np.exp), _unary(np.log), _unary(np.tan)
    ti.min, ti.max, ti.abs = shim_min, shim_max, shim_abs
    ti.u32 = 'u32'

    def cast(x, dt):
        if dt == 'u32':  # tina/random.py:29,48: scalars become U32; an integer vector keeps its (non-negative) i32 entries,
            return x if isinstance(x, Matrix) else U32(x)  # which meet U32 operands through the reflected operators
        return _to_float(x) if _dtype(dt) == F32 else _to_int(x)
    ti.cast = cast
    ti.expr_init = lambda x: x

    def atomic_min(ref, v):
        old = builtins.int(ref)
        ref.field.data[ref.idx] = min(old, builtins.int(v))
        return old
    ti.atomic_min = atomic_min
    ti.random = lambda *a: F32(np.random.rand())

    def imread(path, channels=0):
        from PIL import Image
        img = np.array(Image.open(path))
        return img.swapaxes(0, 1)[:, ::-1]
    ti.imread = imread
    lang = types.ModuleType('taichi.lang')
    lang.__path__ = []
    ops = types.ModuleType('taichi.lang.common_ops')
    ops.TaichiOperations = type('TaichiOperations', (), {})
    ti.lang, lang.common_ops = lang, ops
    return ti, {'taichi': ti, 'taichi.lang': lang, 'taichi.lang.common_ops': ops}


def make_transformations():
    t = types.ModuleType('transformations')
    for name in ('quaternion_matrix', 'quaternion_multiply', 'quaternion_from_matrix', 'euler_matrix'):
        setattr(t, name, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError('transformations is not installed')))
    return t


NEEDED = ['common', 'advans', 'util.matrix', 'matr.nodes', 'matr.material', 'core.engine', 'core.lighting',
          'core.shader', 'core.triangle', 'mesh.base', 'mesh.simple', 'mesh.model', 'mesh.grid', 'mesh.trans',
          'mesh.cull', 'mesh.norm', 'postp.tonemap', 'postp.fxaa', 'postp.blooming', 'postp.ssao', 'random', 'postp.ssr', 'assimp.obj', 'assimp.gltf', 'core.particle', 'core.wireframe', 'mesh.wire', 'pars.base',
          'pars.simple', 'pars.trans', 'scene.raster']


def load_tina(ref_root='/root/reference'):
    """Import the reference's own modules (unmodified) on top of the emulated runtime.
    -> the `tina` package with the flat namespace the reference code expects (tina.Engine, ...)."""
    for k in [k for k in sys.modules if k == 'tina' or k.startswith('tina.') or k == 'taichi' or k.startswith('taichi.')]:
        del sys.modules[k]
    ti, mods = make_taichi()
    sys.modules.update(mods)
    sys.modules['transformations'] = make_transformations()
    tina = types.ModuleType('tina')
    tina.__path__ = [os.path.join(ref_root, 'tina')]
    tina.lazyguard = False  # tina/lazimp.py: sub-package __init__ files import nothing
    tina.ti = ti
    sys.modules['tina'] = tina
    overrides = {'float': shim_float, 'int': shim_int, 'min': shim_min, 'max': shim_max, 'print': lambda *a, **k: None}

    # the names must be in the module globals BEFORE its body runs (decorators, defaults): pre-seed
    class Finder:
        pass
    import importlib.abc
    import importlib.util

    class Loader(importlib.abc.Loader):
        def __init__(self, path):
            self.path = path

        def create_module(self, spec):
            return None

        def exec_module(self, module):
            if not module.__name__.startswith(('tina.util', 'tina.assimp')):  # pure-numpy host modules
                module.__dict__.update(overrides)
            src = open(self.path).read()
            exec(compile(src, self.path, 'exec'), module.__dict__)

    class MetaFinder(importlib.abc.MetaPathFinder):
        def find_spec(self, fullname, path, target=None):
            if not fullname.startswith('tina.'):
                return None
            rel = fullname.split('.')[1:]
            base = os.path.join(ref_root, 'tina', *rel)
            if os.path.isdir(base):
                return importlib.util.spec_from_file_location(fullname, os.path.join(base, '__init__.py'),
                                                              loader=Loader(os.path.join(base, '__init__.py')),
                                                              submodule_search_locations=[base])
            if os.path.exists(base + '.py'):
                return importlib.util.spec_from_file_location(fullname, base + '.py', loader=Loader(base + '.py'))
            return None

    finder = MetaFinder()
    sys.meta_path.insert(0, finder)
    try:
        for name in NEEDED:
            mod = importlib.import_module('tina.' + name)
            for k, v in mod.__dict__.items():
                if not k.startswith('_') and k not in overrides and (callable(v) or isinstance(v, (builtins.int, builtins.float))):
                    if getattr(v, '__module__', mod.__name__) == mod.__name__ or k in ('V', 'MAX'):
                        setattr(tina, k, v)
    finally:
        sys.meta_path.remove(finder)
    return tina
