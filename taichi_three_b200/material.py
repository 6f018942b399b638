"""Material node graph (reference tina/matr/nodes.py, tina/matr/material.py).

The reference binds these nodes at Taichi-JIT time; here the same Python-facing classes are
*flattened on the host* into the postfix program of include/tina_b200.h (TinaMaterial)
which the deferred shading kernel interprets.  Nodes outside the rasteriser's subset
(path-tracing `sample*`, IBL, Glass/Mirror/volumes, LambdaNode, procedural textures) raise
NotImplementedError instead of silently rendering something else.
"""
import numpy as np

from . import _lib


class Node:
    arguments = []
    defaults = []

    def __init__(self, **kwargs):
        # nodes.py:9-33: defaults, literal -> Const, str -> Texture (image path) or Input
        self.params = {}
        for dfl, key in zip(self.defaults, self.arguments):
            if key in kwargs:
                value = kwargs.pop(key)
            else:
                if dfl is None:
                    raise ValueError(f'`{key}` must specified for `{type(self)}`')
                value = dfl
            self.params[key] = as_node(value)
        for key in kwargs:
            raise TypeError(f"{type(self).__name__}() got an unexpected keyword argument '{key}', "
                            f"supported keywords are: {self.arguments}")

    def param(self, key):
        return self.params[key]


def as_node(value):
    if isinstance(value, Node):
        return value
    if isinstance(value, (int, float, np.integer, np.floating)):
        return Const(value)
    if isinstance(value, (list, tuple, np.ndarray)):
        return Const(np.asarray(value, dtype=np.float64))
    if isinstance(value, str):
        if any(value.endswith(x) for x in ('.png', '.jpg', '.bmp')):
            return Texture(value)
        return Input(value)
    raise TypeError(f'cannot use {value!r} as a material parameter')


class Const(Node):
    def __init__(self, value):  # nodes.py:42-49
        self.value = value


class Param(Node):
    """0-d runtime parameter (nodes.py:52-76); its current value is read at every render."""

    def __init__(self, dtype=float, dim=None, initial=0):
        self._value = np.zeros(dim if dim is not None else (), dtype=np.float64)
        self._value[...] = initial
        self.initial = initial
        self.value = self

    def __getitem__(self, idx):
        return self._value.copy() if self._value.ndim else float(self._value)

    def __setitem__(self, idx, v):
        self._value[...] = v

    def make_slider(self, gui, title, min=0, max=1, step=0.01):
        self.slider = gui.slider(title, min, max, step)
        self.slider.value = self.initial

        @gui.post_show
        def post_show(gui):
            self._value[...] = self.slider.value


class Input(Node):
    NAMES = {'pos': 0, 'color': 1, 'normal': 2, 'texcoord': 3}

    def __init__(self, name):  # nodes.py:79-96
        if name not in self.NAMES:
            raise ValueError(f'unknown shader input {name!r} (have {list(self.NAMES)})')
        self.name = name


class Texture(Node):
    arguments = ['texcoord']
    defaults = ['texcoord']

    def __init__(self, path, **kwargs):  # nodes.py:99-111 + advans.py:8-28 texture_as_field
        if isinstance(path, str):
            from PIL import Image
            # ti.imread convention: [x][y] with y up
            img = np.array(Image.open(path))
            img = np.swapaxes(img, 0, 1)[:, ::-1]
        else:
            img = np.array(path)
        if img.dtype == np.uint8:
            img = np.float32(img / 255)
        img = np.ascontiguousarray(img, dtype=np.float32)
        if img.ndim == 2:
            img = img[:, :, None]
        if img.ndim != 3 or img.shape[2] not in (1, 3, 4):
            raise ValueError(f'unsupported texture shape {img.shape}')
        if img.shape[2] == 4:  # colour math in the reference is vec3; keep RGB
            img = np.ascontiguousarray(img[:, :, :3])
        self.image = img
        self._device = None
        super().__init__(**kwargs)

    def device_tensor(self, device):
        import torch
        if self._device is None or self._device.device != device:
            t = torch.as_tensor(self.image).to(device)
            if t.dim() == 3 and t.shape[2] == 3:  # RGB -> RGBA texels: the sampler reads one 128-bit word per texel
                t = torch.cat([t, torch.zeros_like(t[..., :1])], dim=2)
            self._device = t.contiguous()
        return self._device


class ChessboardTexture(Node):  # nodes.py:114-126: lerp((texcoord // size).sum() % 2, color0, color1)
    arguments = ['texcoord', 'size', 'color0', 'color1']
    defaults = ['texcoord', 0.1, 0.4, 0.9]


class LerpTexture(Node):  # nodes.py:129-136: lerp(uv.x, x0, x1) + lerp(uv.y, x0, x1) (sic: y0 / y1 are unused there)
    arguments = ['x0', 'x1', 'y0', 'y1', 'texcoord']
    defaults = [0.0, 1.0, 0.0, 0.0, 'texcoord']


class FresnelFactor(Node):  # material.py:69-83
    arguments = ['metallic', 'albedo', 'specular']
    defaults = [0.0, 1.0, 0.5]


class IMaterial(Node):  # material.py:5-57 (brdf / ambient / emission only)
    def __add__(self, other):
        return AddMaterial(self, other)

    def mix(self, other, factor):
        return MixMaterial(self, other, factor)

    def __mul__(self, factor):
        return ScaleMaterial(self, factor)

    __rmul__ = __mul__


class MixMaterial(IMaterial):  # material.py:91-118
    arguments = ['factor']
    defaults = [0.5]

    def __init__(self, mat1, mat2, factor):
        super().__init__(factor=factor)
        self.mat1, self.mat2 = mat1, mat2


class ScaleMaterial(IMaterial):  # material.py:152-176
    arguments = ['factor']
    defaults = [1.0]

    def __init__(self, mat, factor):
        super().__init__(factor=factor)
        self.mat = mat


class AddMaterial(IMaterial):  # material.py:199-220
    def __init__(self, mat1, mat2):
        super().__init__()
        self.mat1, self.mat2 = mat1, mat2


class Lambert(IMaterial):  # material.py:387-396
    pass


class Phong(IMaterial):  # material.py:445-457
    arguments = ['shineness']
    defaults = [32.0]


class CookTorrance(IMaterial):  # material.py:243-362
    arguments = ['roughness', 'fresnel']
    defaults = [0.4, 1.0]


class Emission(IMaterial):  # material.py:659-676
    pass


def Classic(color='color', shineness=32, specular=0.4):  # material.py:684-688
    return MixMaterial(Lambert() * color, Phong(shineness=shineness), specular)


def Diffuse(color='color'):  # material.py:691-693
    return Lambert() * color


def Lamp(color='color'):  # material.py:696-698
    return Emission() * color


def PBR(basecolor='color', metallic=0.0, roughness=0.4, specular=0.5):  # material.py:701-706
    mat_diff = Lambert() * basecolor
    f0 = FresnelFactor(metallic=metallic, albedo=basecolor, specular=specular)
    mat_spec = CookTorrance(roughness=roughness, fresnel=f0)
    return MixMaterial(mat_diff, mat_spec, f0)


# ---------------------------------------------------------------------------------------
# flattening
# ---------------------------------------------------------------------------------------
class _Program:
    def __init__(self):
        self.code = []
        self.textures = []

    def emit(self, op, arg=0, c=(0.0, 0.0, 0.0)):
        self.code.append((op, int(arg), tuple(float(x) for x in c)))

    def const(self, v):
        a = np.asarray(v, dtype=np.float64).reshape(-1)
        if a.size == 1:
            a = np.repeat(a, 3)
        if a.size == 2:
            a = np.append(a, 0.0)
        if a.size != 3:
            raise ValueError(f'material constants must be scalars or 3-vectors, got {v!r}')
        self.emit(_lib.OP_CONST, 0, a)

    def tex_index(self, node):
        for i, t in enumerate(self.textures):
            if t is node:
                return i
        if len(self.textures) >= _lib.TINA_MAX_TEX:
            raise NotImplementedError(f'at most {_lib.TINA_MAX_TEX} textures per material')
        self.textures.append(node)
        return len(self.textures) - 1


def _emit_value(P, node):
    """Parameter nodes (evaluate to a value)."""
    if isinstance(node, Param):
        P.const(node[None])
    elif isinstance(node, Const):
        P.const(node.value)
    elif isinstance(node, Input):
        P.emit(_lib.OP_INPUT, Input.NAMES[node.name])
    elif isinstance(node, Texture):
        _emit_value(P, node.param('texcoord'))
        P.emit(_lib.OP_TEXTURE, P.tex_index(node))
    elif isinstance(node, FresnelFactor):
        for key in ('metallic', 'albedo', 'specular'):
            _emit_value(P, node.param(key))
        P.emit(_lib.OP_FRESNEL)
    elif isinstance(node, ChessboardTexture):  # lerp(fac, color0, color1) = (1 - fac) * color0 + fac * color1 = OP_MIX
        _emit_value(P, node.param('texcoord'))
        _emit_value(P, node.param('size'))
        P.emit(_lib.OP_CHESS)
        _emit_value(P, node.param('color0'))
        _emit_value(P, node.param('color1'))
        P.emit(_lib.OP_MIX)
    elif isinstance(node, LerpTexture):
        for comp in (0, 1):
            _emit_value(P, node.param('texcoord'))
            P.emit(_lib.OP_BCAST, comp)
            _emit_value(P, node.param('x0'))
            _emit_value(P, node.param('x1'))
            P.emit(_lib.OP_MIX)
        P.emit(_lib.OP_ADD)
    else:
        raise NotImplementedError(f'material parameter node {type(node).__name__} is not supported by the B200 rasteriser')


def _emit_material(P, m, what):
    """what: 'brdf' | 'ambient' | 'emission' (the three methods Lighting.shade_color calls, lighting.py:84-98)."""
    if isinstance(m, MixMaterial):
        _emit_value(P, m.param('factor'))
        _emit_material(P, m.mat1, what)
        _emit_material(P, m.mat2, what)
        P.emit(_lib.OP_MIX)
    elif isinstance(m, ScaleMaterial):
        _emit_value(P, m.param('factor'))
        _emit_material(P, m.mat, what)
        P.emit(_lib.OP_MUL)
    elif isinstance(m, AddMaterial):
        _emit_material(P, m.mat1, what)
        _emit_material(P, m.mat2, what)
        P.emit(_lib.OP_ADD)
    elif isinstance(m, Lambert):
        if what == 'brdf':
            P.emit(_lib.OP_LAMBERT)
        else:
            P.const(1.0 if what == 'ambient' else 0.0)
    elif isinstance(m, Phong):
        if what == 'brdf':
            _emit_value(P, m.param('shineness'))
            P.emit(_lib.OP_PHONG)
        else:
            P.const(1.0 if what == 'ambient' else 0.0)
    elif isinstance(m, CookTorrance):
        if what == 'brdf':
            _emit_value(P, m.param('roughness'))
            _emit_value(P, m.param('fresnel'))
            P.emit(_lib.OP_COOK)
        else:
            P.const(1.0 if what == 'ambient' else 0.0)
    elif isinstance(m, Emission):
        P.const(1.0 if what == 'emission' else 0.0)
    else:
        raise NotImplementedError(f'material {type(m).__name__} is not supported by the B200 rasteriser')


def flatten_material(material):
    """-> (code_brdf, code_ambient, code_emission, [Texture nodes]); each code is a list of (op, arg, c3)."""
    P = _Program()
    out = []
    for what in ('brdf', 'ambient', 'emission'):
        P.code = []
        _emit_material(P, material, what)
        out.append(P.code)
    total = sum(len(c) for c in out)
    if total > _lib.TINA_MAX_INSTR:
        raise NotImplementedError(f'material program too long ({total} > {_lib.TINA_MAX_INSTR})')
    return out[0], out[1], out[2], P.textures


_F = np.float32
_INV_PI = _F(0.3183098861837907)  # Python folds `1 / ti.pi` in f64 (material.py:393), stored as f32


def fold_program(code, color_is_one=True):
    """Constant-fold a postfix program on the host with numpy f32 arithmetic in exactly the
    op order of the device VM (IEEE, no contraction => same bits as evaluating per pixel).
    `Input('color')` is the constant (1, 1, 1) on the triangle path (triangle.py:48); particles carry
    a per-particle colour (particle.py:159), so there it stays a runtime input."""
    stack = []  # items: ('c', f32[3]) constant | ('d', [instr, ...]) dynamic sub-program

    def code_of(item):
        return [(_lib.OP_CONST, 0, tuple(float(x) for x in item[1]))] if item[0] == 'c' else item[1]

    def push_dyn(operands, instr):
        seq = []
        for o in operands:
            seq += code_of(o)
        stack.append(('d', seq + [instr]))

    for op, arg, c in code:
        ins = (op, arg, c)
        if op == _lib.OP_CONST:
            stack.append(('c', np.asarray(c, dtype=_F)))
        elif op == _lib.OP_INPUT:
            if arg == 1 and color_is_one:
                stack.append(('c', np.ones(3, dtype=_F)))
            else:
                stack.append(('d', [ins]))
        elif op == _lib.OP_LAMBERT:
            stack.append(('c', np.full(3, _INV_PI, dtype=_F)))
        elif op == _lib.OP_TEXTURE:
            push_dyn([stack.pop()], ins)
        elif op == _lib.OP_BCAST:
            v = stack.pop()
            if v[0] == 'c':
                stack.append(('c', np.full(3, v[1][arg], dtype=_F)))
            else:
                push_dyn([v], ins)
        elif op == _lib.OP_CHESS:
            size, uv = stack.pop(), stack.pop()
            push_dyn([uv, size], ins)
        elif op == _lib.OP_PHONG:
            push_dyn([stack.pop()], ins)
        elif op == _lib.OP_COOK:
            f0, ro = stack.pop(), stack.pop()
            push_dyn([ro, f0], ins)
        elif op == _lib.OP_FRESNEL:
            sp, al, me = stack.pop(), stack.pop(), stack.pop()
            if sp[0] == al[0] == me[0] == 'c':
                m, a, s_ = me[1], al[1], sp[1]
                stack.append(('c', m * a + (_F(1) - m) * _F(0.16) * (s_ * s_)))
            else:
                push_dyn([me, al, sp], ins)
        elif op == _lib.OP_MIX:
            b, a, f = stack.pop(), stack.pop(), stack.pop()
            if b[0] == a[0] == f[0] == 'c':
                stack.append(('c', (_F(1) - f[1]) * a[1] + f[1] * b[1]))
            else:
                push_dyn([f, a, b], ins)
        elif op == _lib.OP_MUL:
            w, f = stack.pop(), stack.pop()
            if w[0] == f[0] == 'c':
                stack.append(('c', f[1] * w[1]))
            elif w[0] == 'c' and np.all(w[1] == _F(1)):
                stack.append(f)  # x * 1 == x bit for bit (PBR's `basecolor * color` with color = 1)
            elif f[0] == 'c' and np.all(f[1] == _F(1)):
                stack.append(w)
            else:
                push_dyn([f, w], ins)
        elif op == _lib.OP_ADD:
            b, a = stack.pop(), stack.pop()
            if b[0] == a[0] == 'c':
                stack.append(('c', a[1] + b[1]))
            else:
                push_dyn([a, b], ins)
        else:
            raise ValueError(op)
    assert len(stack) == 1, 'malformed material program'
    return code_of(stack[0])


# ---- hoisting: light-independent sub-expressions are evaluated once per pixel -------------------
_ARITY = {}


def _arity(op):
    return {_lib.OP_CONST: 0, _lib.OP_INPUT: 0, _lib.OP_LAMBERT: 0, _lib.OP_TEXTURE: 1, _lib.OP_PHONG: 1,
            _lib.OP_COOK: 2, _lib.OP_FRESNEL: 3, _lib.OP_MIX: 3, _lib.OP_MUL: 2, _lib.OP_ADD: 2, _lib.OP_REG: 0,
            _lib.OP_BCAST: 1, _lib.OP_CHESS: 2}[op]


def _to_tree(code):
    """postfix -> nested tuples (op, arg, c, children...) with children in push order"""
    st = []
    for op, arg, c in code:
        n = _arity(op)
        kids = tuple(st[len(st) - n:]) if n else ()
        del st[len(st) - n:]
        st.append((op, arg, tuple(c)) + kids)
    assert len(st) == 1
    return st[0]


def _to_code(tree, out):
    for kid in tree[3:]:
        _to_code(kid, out)
    out.append((tree[0], tree[1], tree[2]))
    return out


def _light_dependent(tree):
    return tree[0] in (_lib.OP_PHONG, _lib.OP_COOK) or any(_light_dependent(k) for k in tree[3:])


def hoist_programs(brdf, amb, emi):
    """(folded postfix programs) -> (brdf, ambient, emission, prologue).
    Every maximal light-independent, non-constant sub-expression -- and every texture sample inside one
    -- is given a register (identical sub-expressions share it), computed by the prologue program once
    per pixel, and replaced by OP_REG in the three programs.  Same ops on the same values => same bits;
    what changes is that a texture feeding both the diffuse colour and the Fresnel factor of tina.PBR is
    sampled once instead of once per use and per light, and that stock materials with textured / per-pixel
    parameters keep the [X, X, ..., COOK|PHONG, MIX] shapes the specialised shading kernels recognise."""
    regs = {}      # structural key -> register index
    prologue = []  # postfix code
    # sub-expressions that occur more than once (tina.PBR: the Fresnel factor of the sampled base colour feeds the
    # specular colour, the diffuse weight and the ambient term) are evaluated once, into a register of their own
    trees = [_to_tree(code) for code in (brdf, amb, emi)]
    seen_count = {}

    def count(tree):
        if tree[0] not in (_lib.OP_CONST, _lib.OP_REG, _lib.OP_INPUT):
            seen_count[tree] = seen_count.get(tree, 0) + 1
        for k in tree[3:]:
            count(k)
    for t in trees:
        count(t)

    def reg_of(tree):
        """emit `tree` (light independent, not constant) into the prologue; -> OP_REG leaf"""
        # inner texture samples get registers of their own so that they are shared
        kids = tuple(hoist_li(k) for k in tree[3:])
        tree = tree[:3] + kids
        if tree in regs:
            return (_lib.OP_REG, regs[tree], (0.0, 0.0, 0.0))
        if len(regs) >= _lib.TINA_MAX_REGS:
            return tree  # out of registers: leave it inline
        r = len(regs)
        regs[tree] = r
        _to_code(tree, prologue)
        prologue.append((_lib.OP_STORE, r, (0.0, 0.0, 0.0)))
        return (_lib.OP_REG, r, (0.0, 0.0, 0.0))

    def hoist_li(tree):
        """inside a light-independent expression: texture samples and repeated sub-expressions get a register"""
        if tree[0] == _lib.OP_TEXTURE or seen_count.get(tree, 0) >= 2:
            return reg_of(tree)
        return tree[:3] + tuple(hoist_li(k) for k in tree[3:])

    def hoist(tree):
        if tree[0] in (_lib.OP_CONST, _lib.OP_REG):
            return tree
        if not _light_dependent(tree):
            return reg_of(tree)
        return tree[:3] + tuple(hoist(k) for k in tree[3:])

    out = [_to_code(hoist(t), []) for t in trees]
    return out[0], out[1], out[2], prologue


def _walk(node, seen):
    if id(node) in seen or not isinstance(node, Node):
        return
    seen[id(node)] = node
    for child in getattr(node, 'params', {}).values():
        _walk(child, seen)
    for attr in ('mat', 'mat1', 'mat2'):
        if hasattr(node, attr):
            _walk(getattr(node, attr), seen)


def param_signature(material):
    """Values of every runtime Param in the graph (cheap after the first call: the node list is cached)."""
    params = material.__dict__.get('_tina_params')
    if params is None:
        seen = {}
        _walk(material, seen)
        params = [n for n in seen.values() if isinstance(n, Param)]
        material.__dict__['_tina_params'] = params
    return tuple(p._value.tobytes() for p in params)


def prologue_three_address(pro):
    """Postfix prologue -> three-address form (include/tina_b200.h, TINA_OP3): [(op, arg, c), ...] slots.
    Leaves (CONST / INPUT / REG) become source codes of the operation that consumes them, results go to the register
    a following STORE names or to a temporary (8..15).  Returns None when the program does not fit (then the postfix
    form is interpreted)."""
    out = []
    stack = []     # operand descriptors: ('r', idx) | ('i', k) | ('c', (x, y, z))
    free_tmp = list(range(_lib.TINA_VM_VALUES - 1, _lib.TINA_MAX_REGS - 1, -1))
    last_op_slot = None  # index in `out` of the header that produced the value on top of the stack (if a temporary)

    def release(d):
        if d[0] == 'r' and d[1] >= _lib.TINA_MAX_REGS:
            free_tmp.append(d[1])

    def src_code(d):
        return d[1] if d[0] == 'r' else 16 + d[1] if d[0] == 'i' else 255

    for op, arg, c in pro:
        if op in (_lib.OP_CONST, _lib.OP_INPUT, _lib.OP_REG, _lib.OP_LAMBERT):
            stack.append(('c', tuple(c)) if op == _lib.OP_CONST else ('i', arg) if op == _lib.OP_INPUT else
                         ('r', arg) if op == _lib.OP_REG else ('c', (_INV_PI,) * 3))
            last_op_slot = None
        elif op == _lib.OP_STORE:
            top = stack.pop()
            if last_op_slot is not None and top[0] == 'r' and top[1] >= _lib.TINA_MAX_REGS:
                o, a, cc = out[last_op_slot]
                out[last_op_slot] = (o, (a & ~0xff) | arg, cc)  # retarget the producing operation
                release(top)
            else:
                out.append((_lib.OP3 | _lib.OP_REG, arg | src_code(top) << 8, (0.0, 0.0, 0.0)))
                if top[0] == 'c':
                    out.append((_lib.OP_CONST, 0, top[1]))
            last_op_slot = None
        elif op in (_lib.OP_TEXTURE, _lib.OP_FRESNEL, _lib.OP_MIX, _lib.OP_MUL, _lib.OP_ADD):
            n = _arity(op)
            srcs = stack[len(stack) - n:]
            del stack[len(stack) - n:]
            for d in srcs:
                release(d)
            if not free_tmp:
                return None
            dst = free_tmp.pop()
            a = dst
            for k, d in enumerate(srcs):
                a |= src_code(d) << (8 + 8 * k)
            last_op_slot = len(out)
            out.append((_lib.OP3 | op, a, (float(arg), 0.0, 0.0)))
            out += [(_lib.OP_CONST, 0, d[1]) for d in srcs if d[0] == 'c']
            stack.append(('r', dst))
        else:
            return None  # light-dependent ops never reach the prologue
    return out if not stack else None


# prologue of tina.PBR(basecolor=Texture(...)) with constant metallic / roughness / specular after folding, CSE and
# hoisting: (op, arg) per slot, None = any constant.  The device runs this shape as straight-line code.
_PBR_TEX_PROLOGUE = [(_lib.OP_INPUT, 3), (_lib.OP_TEXTURE, None), (_lib.OP_STORE, 0),
                     (_lib.OP_CONST, None), (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_FRESNEL, None), (_lib.OP_STORE, 1),
                     (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MUL, None), (_lib.OP_STORE, 2),
                     (_lib.OP_REG, 1), (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MIX, None), (_lib.OP_STORE, 3),
                     (_lib.OP_REG, 1), (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MUL, None), (_lib.OP_CONST, None),
                     (_lib.OP_MIX, None), (_lib.OP_STORE, 4)]


# tina.Classic(color=Texture(...)): r0 = texel, r1 = r0 * c4, r2 = mix(c7, r0, c9), r3 = mix(c12, r0 * c14, c16)
_CLASSIC_TEX_PROLOGUE = [(_lib.OP_INPUT, 3), (_lib.OP_TEXTURE, None), (_lib.OP_STORE, 0),
                         (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MUL, None), (_lib.OP_STORE, 1),
                         (_lib.OP_CONST, None), (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MIX, None), (_lib.OP_STORE, 2),
                         (_lib.OP_CONST, None), (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MUL, None), (_lib.OP_CONST, None),
                         (_lib.OP_MIX, None), (_lib.OP_STORE, 3)]
# tina.Diffuse(color=Texture(...)): r0 = texel, r1 = r0 * c4, r2 = r0 * c8
_DIFFUSE_TEX_PROLOGUE = [(_lib.OP_INPUT, 3), (_lib.OP_TEXTURE, None), (_lib.OP_STORE, 0),
                         (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MUL, None), (_lib.OP_STORE, 1),
                         (_lib.OP_REG, 0), (_lib.OP_CONST, None), (_lib.OP_MUL, None), (_lib.OP_STORE, 2)]
_PROLOGUE_SHAPES = {1: _PBR_TEX_PROLOGUE, 3: _CLASSIC_TEX_PROLOGUE, 4: _DIFFUSE_TEX_PROLOGUE}


def prologue_form(pro):
    """TinaMaterial.prologue_form of a postfix prologue: 1 / 3 / 4 = the shape of tina.PBR / Classic / Diffuse with a
    textured colour (straight-line device code), else 0 (include/tina_b200.h)."""
    for form, shape in _PROLOGUE_SHAPES.items():
        if len(pro) == len(shape) and all(op == top and (targ is None or arg == targ) for (op, arg, _), (top, targ) in zip(pro, shape)):
            return form
    return 0


def compile_material(material, fold=True, color_is_one=True):
    """flatten -> constant-fold -> hoist.  -> (brdf, ambient, emission, prologue, textures)"""
    brdf, amb, emi, textures = flatten_material(material)
    pro = []
    if fold:
        brdf, amb, emi = (fold_program(c, color_is_one) for c in (brdf, amb, emi))
        brdf, amb, emi, pro = hoist_programs(brdf, amb, emi)
    if len(brdf) + len(amb) + len(emi) + len(pro) > _lib.TINA_MAX_INSTR:
        raise NotImplementedError('material program too long')
    return brdf, amb, emi, pro, textures


def material_struct(material, device, fold=True, color_is_one=True):
    """Build the TinaMaterial POD; returns (struct, keepalive list of device tensors)."""
    brdf, amb, emi, pro, textures = compile_material(material, fold, color_is_one)
    m = _lib.TinaMaterial()
    m.n_brdf, m.n_ambient, m.n_emission, m.ntex = len(brdf), len(amb), len(emi), len(textures)
    m.prologue_form = prologue_form(pro)
    if pro and m.prologue_form == 0:  # any other prologue: three-address form, a third of the interpreter steps
        pro3 = prologue_three_address(pro)
        if pro3 is not None and len(brdf) + len(amb) + len(emi) + len(pro3) <= _lib.TINA_MAX_INSTR:
            pro, m.prologue_form = pro3, 2
    m.n_prologue = len(pro)
    keep = []
    for i, t in enumerate(textures):
        d = t.device_tensor(device)
        keep.append(d)
        m.tex[i] = d.data_ptr()
        m.tex_w[i], m.tex_h[i], m.tex_c[i] = d.shape[0], d.shape[1], d.shape[2]
    for i, (op, arg, c) in enumerate(brdf + amb + emi + pro):
        m.code[i].op, m.code[i].arg = op, arg
        m.code[i].c[0], m.code[i].c[1], m.code[i].c[2] = c
    return m, keep


# ---- material.sample() trees for SSR (postp/ssr.py:78-80) ---------------------------------------------------------
def _value_is_scalar(node):
    """Does the reference evaluate this parameter node to a scalar?  (MixMaterial.sample averages a vector factor with
    Vavg and takes a scalar one as it is, common.py:36-40.)"""
    if isinstance(node, Param):
        return np.ndim(node[None]) == 0
    if isinstance(node, Const):
        return np.ndim(node.value) == 0 or np.size(node.value) == 1
    if isinstance(node, Texture):
        return node.image.shape[2] == 1
    if isinstance(node, FresnelFactor):
        return all(_value_is_scalar(node.param(k)) for k in ('metallic', 'albedo', 'specular'))
    if isinstance(node, ChessboardTexture):
        return all(_value_is_scalar(node.param(k)) for k in ('color0', 'color1'))
    if isinstance(node, LerpTexture):
        return all(_value_is_scalar(node.param(k)) for k in ('x0', 'x1'))
    return False


def sample_struct(material, device):
    """Material node graph -> (TinaSampleMaterial, keep-alive): the tree `material.sample(idir, nrm, sign, rng)` of
    matr/material.py descends -- MixMaterial (:123-138), ScaleMaterial (:180-184), AddMaterial (:227-238) over the
    Lambert (:398-405), Phong (:459-472), CookTorrance (:364-384) and Emission (:679-681) leaves -- in pre-order
    (node 0 = root), every node parameter as a postfix value program."""
    P = _Program()
    nodes = []

    def value(node):
        start = len(P.code)
        _emit_value(P, node)
        return start, len(P.code) - start

    def walk(m):
        idx = len(nodes)
        rec = dict(kind=0, a=0, b=0, p0=0, n0=0, p1=0, n1=0, pad_=0)
        nodes.append(rec)
        if isinstance(m, MixMaterial):
            f = m.param('factor')
            rec['kind'] = _lib.SNODE_MIX
            rec['p0'], rec['n0'] = value(f)
            rec['pad_'] = 1 if _value_is_scalar(f) else 0
            rec['a'], rec['b'] = walk(m.mat1), walk(m.mat2)
        elif isinstance(m, ScaleMaterial):
            rec['kind'] = _lib.SNODE_SCALE
            rec['p0'], rec['n0'] = value(m.param('factor'))
            rec['a'] = walk(m.mat)
        elif isinstance(m, AddMaterial):
            rec['kind'] = _lib.SNODE_ADD
            rec['a'], rec['b'] = walk(m.mat1), walk(m.mat2)
        elif isinstance(m, Lambert):
            rec['kind'] = _lib.SNODE_LAMBERT
        elif isinstance(m, Phong):
            rec['kind'] = _lib.SNODE_PHONG
            rec['p0'], rec['n0'] = value(m.param('shineness'))
        elif isinstance(m, CookTorrance):
            rec['kind'] = _lib.SNODE_COOK
            rec['p0'], rec['n0'] = value(m.param('roughness'))
            rec['p1'], rec['n1'] = value(m.param('fresnel'))
        elif isinstance(m, Emission):
            rec['kind'] = _lib.SNODE_EMISSION
        else:
            raise NotImplementedError(f'material {type(m).__name__} has no sample() on the B200 raster path')
        return idx

    walk(material)
    if len(nodes) > _lib.TINA_SAMPLE_MAX_NODES or len(P.code) > _lib.TINA_SAMPLE_MAX_INSTR:
        raise NotImplementedError(f'material too large for SSR ({len(nodes)} nodes, {len(P.code)} parameter instructions)')
    out = _lib.TinaSampleMaterial()
    out.nnodes, out.ncode, out.ntex = len(nodes), len(P.code), len(P.textures)
    for i, rec in enumerate(nodes):
        for key, v in rec.items():
            setattr(out.nodes[i], key, v)
    for i, (op, arg, c) in enumerate(P.code):
        out.code[i].op, out.code[i].arg = op, arg
        out.code[i].c[0], out.code[i].c[1], out.code[i].c[2] = c
    keep = []
    for i, t in enumerate(P.textures):
        dt = t.device_tensor(device)
        keep.append(dt)
        out.tex[i] = dt.data_ptr()
        out.tex_w[i], out.tex_h[i], out.tex_c[i] = dt.shape
    return out, keep
