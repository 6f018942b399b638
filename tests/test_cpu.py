"""CPU-only tests (`-m "not gpu"`): the oracle against known answers, the host-side logic of the
Python mirror, and that the C-ABI library loads and exports every symbol include/tina_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- oracle known answers (SURVEY.md §8c probes, measured with an independent NumPy restatement) ----
def test_oracle_monkey_probe(O, tina):
    v, _, _ = O.indexed(scenes.load_monkey())
    assert v.shape == (968, 3, 3)
    view, proj = scenes.default_camera()
    W2V = (proj @ view).astype(np.float32)
    assert np.allclose(W2V, [[1.7320508, 0, 0, 0], [0, 1.7320508, 0, 0], [0, 0, -1.0002, 2.90059], [0, 0, -1, 3]], atol=1e-6)
    occup, depth, tie, st = O.render_occup(v, W2V, 512, 512)
    assert st['rasterised'] == 614 and st['culled'] == 354
    assert (occup >= 0).sum() == 64082 and tie.sum() == 0
    d = depth[occup >= 0]
    assert 1.0239e9 < d.min() < 1.0241e9 and 1.0431e9 < d.max() < 1.0433e9
    assert ((depth == 2**30) == (occup == -1)).all()


def test_oracle_meshgrid_probe(O):
    pos, _ = O.grid_positions(64, 64)
    view, proj = scenes.default_camera()
    W2V = (proj @ view).astype(np.float32)
    occup, depth, tie, st = O.render_occup(O.grid_faces(pos), W2V, 512, 512)
    assert (occup >= 0).sum() == 87616 and tie.sum() == 160


def test_oracle_serial_semantics(O):
    """lowest face id wins exact ties; a later pass only wins with strictly smaller depth."""
    tri = np.array([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]], np.float32)
    two = np.concatenate([tri, tri])
    view, proj = scenes.default_camera()
    W2V = (proj @ view).astype(np.float32)
    occup, depth, tie, _ = O.render_occup(two, W2V, 64, 64)
    assert set(np.unique(occup)) == {-1, 0} and tie[occup == 0].all()
    occup2, depth2, _, _ = O.render_occup(tri, W2V, 64, 64, depth=depth)
    assert (occup2 == -1).all() and np.array_equal(depth2, depth)


def test_oracle_parallel_matches_serial_away_from_ties(O):
    tri = scenes.soup(20000, 160, 120, s=0.03, seed=4)
    view, proj = scenes.default_camera(160 / 120)
    W2V = (proj @ view).astype(np.float32)
    o1, d1, tie, _ = O.render_occup(tri, W2V, 160, 120)
    o2, d2, _, _ = O.render_occup(tri, W2V, 160, 120, parallel=True)
    assert np.array_equal(d1, d2)
    assert np.array_equal(o1[tie == 0], o2[tie == 0])


# ---- host logic ---------------------------------------------------------------------------------------
def test_material_flattening(tina):
    from taichi_three_b200 import _lib as L
    brdf, amb, emi, tex = tina.flatten_material(tina.Diffuse())
    assert [i[0] for i in brdf] == [L.OP_INPUT, L.OP_LAMBERT, L.OP_MUL] and brdf[0][1] == 1
    brdf, amb, emi, tex = tina.flatten_material(tina.Classic())
    assert [i[0] for i in brdf] == [L.OP_CONST, L.OP_INPUT, L.OP_LAMBERT, L.OP_MUL, L.OP_CONST, L.OP_PHONG, L.OP_MIX]
    assert brdf[0][2] == (0.4, 0.4, 0.4) and brdf[4][2] == (32.0, 32.0, 32.0)
    img = np.zeros((4, 4, 3), np.uint8)
    brdf, amb, emi, tex = tina.flatten_material(tina.PBR(basecolor=tina.Texture(img), metallic=0.0, roughness=0.5))
    assert len(tex) == 1 and [i[0] for i in brdf].count(L.OP_TEXTURE) == 3 and brdf[-1][0] == L.OP_MIX
    with pytest.raises(TypeError):
        tina.Phong(shininess=3)
    with pytest.raises(ValueError):
        tina.Input('nonsense')
    m = tina.Lambert() * [1, 0, 0] + tina.Emission() * 0.5
    brdf, amb, emi, _ = tina.flatten_material(m)
    assert emi[-1][0] == L.OP_ADD


def test_loaders(tina):
    obj = scenes.load_monkey()
    assert obj['f'].shape == (968, 3, 3) and obj['v'].shape == (507, 3)
    g = scenes.load_cornell()
    prims = [p for n in g.nodes for p in n.primitives]
    assert [len(p.obj['f']) for p in prims] == [12, 12, 10]
    assert g.images[0].shape == (512, 512, 3)
    mat = g._material(prims[2].material)
    assert isinstance(mat, tina.MixMaterial)
    # node TRS: +90 degrees about X
    assert np.allclose(g.nodes[1].trans[:3, :3] @ [0, 0, 1], [0, -1, 0], atol=1e-6)


def test_writeobj_and_pfmwrite_round_trip(tina, tmp_path):
    """assimp/obj.py:103-125, assimp/pfm.py:4-12: what writeobj saves, readobj loads back; the PFM header and payload."""
    import io
    obj = scenes.load_monkey()
    path = tmp_path / 'm.obj'
    tina.writeobj(str(path), obj)
    back = tina.readobj(str(path))
    assert np.array_equal(back['f'], obj['f']) and np.array_equal(back['v'], obj['v']) and np.array_equal(back['vn'], obj['vn'])
    buf = io.StringIO()
    tina.writeobj(buf, {'v': obj['v'][:4], 'f': np.array([[0, 1, 2], [0, 2, 3]])})
    text = buf.getvalue().splitlines()
    assert text[0].startswith('# OBJ file saved by tina.writeobj') and text[-1] == 'f 1/1/1 3/3/3 4/4/4'
    img = np.random.default_rng(0).random((5, 4, 3)).astype(np.float32) * 3
    tina.pfmwrite(str(tmp_path / 'i.pfm'), img)
    raw = open(tmp_path / 'i.pfm', 'rb').read()
    head, w_h, scale, payload = raw.split(b'\n', 3)
    assert head == b'PF' and w_h == b'5 4' and float(scale) == -float(img.max())
    assert np.allclose(np.frombuffer(payload, np.float32).reshape(4, 5, 3) * img.max(), img.swapaxes(0, 1), atol=1e-6)


def test_camera_matrices(tina):
    p = tina.perspective(60, 16 / 9)
    assert np.isclose(p[1, 1], 1 / np.tan(np.radians(30))) and np.isclose(p[0, 0], p[1, 1] * 9 / 16) and p[3, 2] == -1
    v = tina.lookat()
    assert np.allclose(v @ [0, 0, 3, 1], [0, 0, 0, 1])
    view, proj = tina.orbit_camera(center=(0, 2, 0), radius=6, theta=0.2, phi=0.5)
    eye = np.linalg.inv(view)[:3, 3]
    assert np.isclose(np.linalg.norm(eye - [0, 2, 0]), 6)

    class G:
        res = (640, 360)
    c = tina.Control(G())
    assert np.allclose(c.get_camera()[1], tina.perspective(60, 640 / 360))
    R = tina.RotationStep(np.eye(4), 0.1, 0.0, 0.0)
    assert np.allclose(R[:3, :3] @ R[:3, :3].T, np.eye(3), atol=1e-12)


def test_lighting_struct(tina):
    L = tina.Lighting()
    L.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
    L.add_light(pos=[1, 1, 1])
    s = L.struct()
    assert s.nlights == 2 and s.dirs[0][3] == 0 and s.dirs[1][3] == 1
    assert np.isclose(np.linalg.norm(list(s.dirs[0])[:3]), 1, atol=1e-6)


# ---- the C ABI ----------------------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol():
    from taichi_three_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'tina_b200.h')).read()
    declared = set(re.findall(r'\b(tina_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 20
    so = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), f'{name} declared in include/tina_b200.h but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().tina_version() == 100
    # struct layouts agree with the header
    assert ctypes.sizeof(_lib.TinaInstr) == 20 and ctypes.sizeof(_lib.TinaLighting) == 16 * 16 * 2 + 16 + 16


def test_no_cpu_fallback(tina):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        tina.Engine(64)
    with pytest.raises(RuntimeError):
        tina.MeshGrid(8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'taichi_three_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            assert 'oracle' not in open(os.path.join(pkg, fn)).read(), fn
    csrc = os.path.join(pkg, 'csrc')
    for fn in os.listdir(csrc):
        if fn.endswith(('.cu', '.cuh')):
            assert 'oracle' not in open(os.path.join(csrc, fn)).read().replace('the serial CPU restatement', ''), fn


def test_material_compile_shapes(tina):
    """fold + hoist keeps the stock materials in the shapes the specialised shading kernels recognise,
    with texture samples shared through prologue registers."""
    from taichi_three_b200 import _lib as L
    from taichi_three_b200.material import compile_material
    img = np.zeros((4, 4, 3), np.uint8)
    ops = lambda code: [i[0] for i in code]
    b, a, e, p, _ = compile_material(tina.Diffuse())
    assert ops(b) == [L.OP_CONST] and ops(a) == [L.OP_CONST] and not p
    b, a, e, p, _ = compile_material(tina.Classic())
    assert ops(b) == [L.OP_CONST] * 3 + [L.OP_PHONG, L.OP_MIX] and not p
    b, a, e, p, tex = compile_material(tina.PBR(basecolor=tina.Texture(img), metallic=0.0, roughness=0.5))
    assert ops(b) == [L.OP_REG, L.OP_REG, L.OP_CONST, L.OP_REG, L.OP_COOK, L.OP_MIX]
    assert ops(p).count(L.OP_TEXTURE) == 1 and ops(a) == [L.OP_REG] and len(tex) == 1
    b, a, e, p, _ = compile_material(tina.Classic(color=tina.Texture(img)))
    assert ops(b) == [L.OP_CONST, L.OP_REG, L.OP_CONST, L.OP_PHONG, L.OP_MIX] and ops(p).count(L.OP_TEXTURE) == 1


def test_three_address_prologue_equals_postfix(tina):
    """material.prologue_three_address (TinaMaterial.prologue_form 2): evaluating the three-address program gives the
    registers of the postfix prologue bit for bit (numpy f32 evaluators of both forms; the texture is a stub)."""
    from taichi_three_b200 import _lib, material as M
    F = np.float32
    img = np.random.default_rng(0).random((5, 4, 3)).astype(F)

    def tex(uv):
        return (uv * F(0.37) + F(0.11)).astype(F)

    def fres(me, al, sp):
        return me * al + (F(1) - me) * F(0.16) * (sp * sp)

    def mix(f, a, b):
        return (F(1) - f) * a + f * b

    def run_postfix(pro, inputs):
        st, regs = [], {}
        for op, arg, c in pro:
            if op == _lib.OP_CONST:
                st.append(np.asarray(c, F))
            elif op == _lib.OP_INPUT:
                st.append(inputs[arg])
            elif op == _lib.OP_REG:
                st.append(regs[arg])
            elif op == _lib.OP_STORE:
                regs[arg] = st.pop()
            elif op == _lib.OP_TEXTURE:
                st.append(tex(st.pop()))
            elif op == _lib.OP_FRESNEL:
                sp, al, me = st.pop(), st.pop(), st.pop()
                st.append(fres(me, al, sp))
            elif op == _lib.OP_MIX:
                b, a, f = st.pop(), st.pop(), st.pop()
                st.append(mix(f, a, b))
            elif op == _lib.OP_MUL:
                w, f = st.pop(), st.pop()
                st.append(f * w)
            elif op == _lib.OP_ADD:
                b, a = st.pop(), st.pop()
                st.append(a + b)
            else:
                raise AssertionError(op)
        assert not st
        return regs

    def run_three(pro3, inputs):
        vals, pc = {}, 0
        while pc < len(pro3):
            op, a, c = pro3[pc]
            pc += 1
            assert op & _lib.OP3
            op &= 0xff
            ns = 1 if op in (_lib.OP_TEXTURE, _lib.OP_REG) else 2 if op in (_lib.OP_MUL, _lib.OP_ADD) else 3
            s = []
            for k in range(ns):
                code = (a >> (8 + 8 * k)) & 0xff
                if code < 16:
                    s.append(vals[code])
                elif code < 20:
                    s.append(inputs[code - 16])
                else:
                    assert pro3[pc][0] == _lib.OP_CONST
                    s.append(np.asarray(pro3[pc][2], F))
                    pc += 1
            r = (tex(s[0]) if op == _lib.OP_TEXTURE else fres(*s) if op == _lib.OP_FRESNEL else mix(*s) if op == _lib.OP_MIX else
                 s[0] * s[1] if op == _lib.OP_MUL else s[0] + s[1] if op == _lib.OP_ADD else s[0])
            vals[a & 0xff] = r
        return vals

    inputs = [np.asarray(v, F) for v in ([0.1, 0.2, 0.3], [1, 1, 1], [0, 0.6, 0.8], [0.25, 0.75, 0])]
    mats = [tina.PBR(basecolor=tina.Texture(img), metallic=0.3, roughness=0.4), tina.Classic(color=tina.Texture(img)),
            tina.Diffuse(color=tina.Texture(img)),
            tina.Lambert() * tina.Texture(img) + tina.Phong(shineness=16) * tina.Texture(img) + tina.Emission() * 0.1,
            tina.PBR(basecolor=tina.Texture(img), metallic=tina.Texture(img), roughness=0.4)]
    forms = []
    for mat in mats:
        b, a, e, pro, _ = M.compile_material(mat)
        assert pro
        forms.append(M.prologue_form(pro))
        pro3 = M.prologue_three_address(pro)
        assert pro3 is not None and sum(1 for o, _, _ in pro3 if o & _lib.OP3) < len(pro) / 2
        want, got = run_postfix(pro, inputs), run_three(pro3, inputs)
        for r, v in want.items():
            assert np.array_equal(got[r], v), (mat, r)
    assert forms[:4] == [1, 3, 4, 0]


# ---- the oracle's own front-end (round 2) ---------------------------------------------------------------
def test_oracle_workload_recipes_equal_the_bench_recipes(O, tina):
    """oracle/workloads.py restates the camera and the C1 / C2 / C3 inputs without the product (bench.py --impl
    reference); they must be the inputs the product-side bench renders."""
    from oracle import workloads as R
    for aspect in (1.0, 16 / 9):
        v0, p0 = scenes.default_camera(aspect)
        v1, p1 = R.default_camera(aspect)
        assert np.array_equal(v0, v1) and np.array_equal(p0, p1)
    assert np.array_equal(R.wave_grid_pos(48), scenes.wave_grid_pos(48))
    assert np.array_equal(R.soup(5000, 320, 200, s=0.01, seed=7), scenes.soup(5000, 320, 200, s=0.01, seed=7))
    v, _, _ = O.indexed(scenes.load_monkey())
    assert np.array_equal(R.monkey_faces(os.path.join(ROOT, 'tests', 'assets', 'monkey.obj')), v)


def test_oracle_material_front_end_is_independent_and_agrees(O, tina):
    """oracle/materials.py flattens node graphs with its own walker (no product code): same programs as the product's
    flattener on the stock materials and on composed graphs, and the node-free stock builders equal both."""
    from oracle import materials as OM
    from taichi_three_b200.material import flatten_material
    src = open(os.path.join(ROOT, 'oracle', 'materials.py')).read() + open(os.path.join(ROOT, 'oracle', 'oracle.py')).read()
    assert 'import taichi_three_b200' not in src and 'from taichi_three_b200' not in src
    img = np.random.default_rng(0).random((4, 4, 3)).astype(np.float32)
    graphs = [tina.Diffuse(), tina.Classic(), tina.PBR(), tina.Lamp(), tina.Diffuse(color=[.2, .4, .6]),
              tina.PBR(basecolor=tina.Texture(img), metallic=0.3, roughness=0.2),
              (tina.Classic(shineness=8) * 0.5 + tina.Lamp(color=[1, 0, 0])).mix(tina.PBR(metallic=1.0), tina.Texture(img))]

    def code(m, n):
        return [(m.code[i].op, m.code[i].arg, tuple(m.code[i].c)) for i in range(n)]
    for g in graphs:
        brdf, amb, emi, tex = flatten_material(g)
        m, _, arrays = OM.material_pod_of(g)
        assert (m.n_brdf, m.n_ambient, m.n_emission, m.ntex) == (len(brdf), len(amb), len(emi), len(tex))
        want = [(op, arg, tuple(np.float32(x) for x in c)) for op, arg, c in brdf + amb + emi]
        assert code(m, len(want)) == want
    for pod, g in ((OM.stock_diffuse(), tina.Diffuse()), (OM.stock_classic(), tina.Classic())):
        m, n = pod[0], pod[0].n_brdf + pod[0].n_ambient + pod[0].n_emission
        assert code(m, n) == code(OM.material_pod_of(g)[0], n)
    # lights: the default light of the parity tests
    L = tina.Lighting()
    L.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
    L.set_ambient_light([0.1, 0.1, 0.1])
    a, b = OM.lighting_of(L), OM.default_lighting()
    assert bytes(a) == bytes(b) == bytes(L.struct())


def test_sample_trees_of_product_and_oracle_agree(O, tina):
    """material.sample() trees for SSR: the product's flattener (material.sample_struct) and the oracle's own walker
    (oracle/materials.sample_pod_of) produce the same nodes and parameter programs; a scalar mix factor is marked."""
    import torch
    from oracle import materials as OM
    from taichi_three_b200.material import sample_struct
    img = np.random.default_rng(1).random((4, 4, 3)).astype(np.float32)
    graphs = [tina.Diffuse(), tina.Classic(), tina.PBR(), tina.Lamp(), tina.PBR(basecolor=tina.Texture(img), metallic=0.3, roughness=0.2),
              (tina.Classic(shineness=8) * 0.5 + tina.Lamp(color=[1, 0, 0])).mix(tina.PBR(metallic=1.0), [0.2, 0.4, 0.6])]
    for g in graphs:
        a, _ = sample_struct(g, torch.device('cpu'))
        b, _ = OM.sample_pod_of(g)
        assert (a.nnodes, a.ncode, a.ntex) == (b.nnodes, b.ncode, b.ntex)
        for i in range(a.nnodes):
            assert all(getattr(a.nodes[i], k) == getattr(b.nodes[i], k) for k in ('kind', 'a', 'b', 'p0', 'n0', 'p1', 'n1', 'pad_'))
        for i in range(a.ncode):
            assert (a.code[i].op, a.code[i].arg, tuple(a.code[i].c)) == (b.code[i].op, b.code[i].arg, tuple(b.code[i].c))
    a, _ = sample_struct(tina.Classic(), torch.device('cpu'))
    assert a.nodes[0].kind == 4 and a.nodes[0].pad_ == 1  # Classic: MixMaterial with the scalar factor 0.4
    a, _ = sample_struct(graphs[-1], torch.device('cpu'))
    assert a.nodes[0].pad_ == 0  # a vector factor is averaged (Vavg)
    with pytest.raises(NotImplementedError):
        sample_struct(object(), torch.device('cpu'))


def test_primitive_mesh_generators_match_reference_golden():
    """mesh/prim.py: PrimitiveMesh.sphere / .cylinder face lists (vertex, normal, texcoord per corner) equal the reference's
    own generator bit for bit (golden: make_golden.py::case_prims)."""
    from taichi_three_b200.mesh import primitive_sphere, primitive_cylinder
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'particles_prims.npz'))
    assert np.array_equal(primitive_sphere(8, 6, 1), g['sphere_8_6']) and np.array_equal(primitive_sphere(5, 3, 0.7), g['sphere_5_3'])
    assert np.array_equal(primitive_sphere(), g['sphere_default']) and g['sphere_default'].shape == (1472, 3, 3, 3)
    assert np.array_equal(primitive_cylinder(8, 2, 1, 2), g['cylinder_8_2']) and np.array_equal(primitive_cylinder(5, 3, 0.6, 1.5), g['cylinder_5_3'])
    assert np.array_equal(primitive_cylinder(), g['cylinder_default']) and g['cylinder_default'].shape == (320, 3, 3, 3)


def test_bench_reference_arm_is_product_free_and_uses_every_core():
    """`bench.py --impl reference` under torchrun's OMP_NUM_THREADS=1: every host core, the same config object as the
    GPU arm, and no product module (nor libtina_b200.so) in the process."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'c1', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['product_modules_loaded'] == []
    assert line['cpu_baseline']['cores'] == len(os.sched_getaffinity(0))
    sys.path.insert(0, ROOT)
    import bench
    assert line['config'] == bench.config_for('c1', 1)
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['value'] > 0
