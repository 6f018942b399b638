"""Pins the CPU oracle (oracle/tina_oracle.c) -- and the host-side mesh providers / loaders -- against
golden vectors produced by the REFERENCE'S OWN SOURCES executed under oracle/ref_shim
(tests/golden/make_golden.py; taichi itself is not installable here).

  face ids (occup) and integer depth: bit-exact
  colour: <= 2e-6 abs (pre-tonemap and final); the residue is Python-scope f64 constant folding in the
          reference's material graph (e.g. `(1 - metallic) * 0.16 * specular**2` over Const nodes) and
          libm-vs-numpy pow, far inside the 1e-4 colour budget
"""
import glob
import os

import numpy as np
import pytest

import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz'))
               if not os.path.basename(p).startswith(('particles', 'matgraphs')))
COLOR_TOL = 2e-6


def _lighting(tina, g):
    L = tina.Lighting()
    L.nlights[None] = len(g['light_dirs'])
    L.light_dirs[:len(g['light_dirs'])] = g['light_dirs']
    L.light_colors[:len(g['light_colors'])] = g['light_colors']
    L.ambient_color[None] = g['ambient']
    return L


MATERIAL_NAMES = ('PBR', 'Classic', 'Diffuse', 'Lamp', 'Lambert', 'Phong', 'Emission', 'CookTorrance', 'Texture', 'FresnelFactor',
                  'MixMaterial', 'ScaleMaterial', 'AddMaterial', 'ChessboardTexture', 'LerpTexture')


def _material(tina, g, k, key='material'):
    ns = {n: getattr(tina, n) for n in MATERIAL_NAMES}
    for i in range(4):
        if f'tex{i}' in g:
            ns[f'tex{i}'] = g[f'tex{i}']
    return eval(str(g[f'{key}{k}']), ns)


def test_goldens_exist():
    assert len(CASES) >= 9


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_sources(tina, O, case):
    g = np.load(os.path.join(GOLDEN, case + '.npz'))
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    lighting = _lighting(tina, g)
    depth = O.clear_depth(W, H)
    image = np.empty((W, H, 3), np.float32)
    image[...] = g['bgcolor']
    with np.errstate(all='ignore'):
        for k in range(int(g['nobjects'])):
            verts = g[f'verts{k}']
            norms = g[f'norms{k}'] if f'norms{k}' in g else None
            coors = g[f'coors{k}'] if f'coors{k}' in g else None
            occup, depth, tie, _ = O.render_occup(verts, g['W2V'], W, H, flags, g['bias'], depth)
            assert np.array_equal(occup, g[f'occup{k}']), f'occup of object {k}: {(occup != g[f"occup{k}"]).sum()} px differ'
            O.render_color(verts, norms, coors, occup, g['W2V'], g['V2W'], W, H, flags, _material(tina, g, k), lighting, image, g['bias'])
    assert np.array_equal(depth, g['depth'])
    ok = np.isfinite(g['image_pre_tonemap'])
    assert np.array_equal(np.isfinite(image), ok)
    assert np.abs(image[ok] - g['image_pre_tonemap'][ok]).max() <= COLOR_TOL
    final = O.tonemap(image)
    ok = np.isfinite(g['image'])
    assert np.abs(final[ok] - g['image'][ok]).max() <= COLOR_TOL
    assert (depth < 2**30).sum() > 100


def test_mesh_providers_match_reference_set_object(tina, O):
    """The arrays the reference's set_object kernel wrote into raster.verts / norms / coors
    (MeshModel, MeshGrid.pre_compute, MeshTransform, MeshNoCulling, MeshFlipCulling, MeshFlipNormal,
    glTF node transforms) equal the oracle-side providers bit for bit."""
    # MeshModel
    g = np.load(os.path.join(GOLDEN, 'monkey_flat_diffuse.npz'))
    v, vn, vt = O.indexed(scenes.load_monkey())
    assert np.array_equal(v, g['verts0'])
    # MeshNoCulling(MeshGrid) with per-frame normals
    g = np.load(os.path.join(GOLDEN, 'grid_wave_nocull_smooth_classic.npz'))
    pos = np.load(os.path.join(GOLDEN, 'grid_wave_pos.npy'))
    p0, _ = O.grid_positions(14, 14)
    assert np.array_equal(p0[..., :2], pos[..., :2])  # MeshGrid init (grid.py:17-21)
    fv, fn, _ = O.no_culling(O.grid_faces(pos), O.grid_faces(O.grid_normals(pos)))
    assert np.array_equal(fv, g['verts0']) and np.array_equal(fn, g['norms0'])
    # MeshFlipNormal(MeshFlipCulling(MeshTransform(MeshModel)))
    g = np.load(os.path.join(GOLDEN, 'monkey_transform_flip_lights_addmaterial.npz'))
    trans = np.load(os.path.join(GOLDEN, 'monkey_trans.npy'))
    assert np.array_equal(trans, tina.translate([0.2, -0.1, 0.3]) @ tina.eularXYZ([0.3, 0.8, -0.2]) @ tina.scale([0.9, 1.1, 0.8]))
    v, vn = O.transform(v, vn, trans)
    assert np.array_equal(v[:, ::-1], g['verts0']) and np.array_equal(-vn[:, ::-1], g['norms0'])
    assert np.array_equal(vt[:, ::-1], g['coors0'])
    # glTF: loader + node TRS + MeshTransform(MeshModel)
    g = np.load(os.path.join(GOLDEN, 'cornell_pbr_textured.npz'))
    gltf = scenes.load_cornell()
    objs = scenes.cornell_oracle_objects(gltf)
    assert len(objs) == int(g['nobjects']) == 3
    for k, (v, vn, vt, mat) in enumerate(objs):
        assert np.array_equal(v, g[f'verts{k}']) and np.array_equal(vn, g[f'norms{k}']) and np.array_equal(vt, g[f'coors{k}'])
    assert np.array_equal(gltf.images[0], g['tex0'])
    # camera: engine.set_camera (engine.py:72-76) of the same view/proj
    view, proj = tina.orbit_camera(center=(0, 2, 0), radius=6.0, theta=0.2, phi=0.7)
    assert np.array_equal((proj @ view).astype(np.float32), g['W2V'])
    assert np.array_equal(np.linalg.inv(proj @ view).astype(np.float32), g['V2W'])


def test_camera_helpers_match_reference(tina):
    g = np.load(os.path.join(GOLDEN, 'monkey_flat_diffuse.npz'))
    W2V = tina.perspective(60, 1.0) @ tina.lookat()
    assert np.array_equal(W2V.astype(np.float32), g['W2V'])
    g = np.load(os.path.join(GOLDEN, 'grid_wave_nocull_smooth_classic.npz'))
    W2V = tina.perspective(60, 80 / 60) @ tina.lookat(back=(1.0, 1.5, 2.5))
    assert np.array_equal(W2V.astype(np.float32), g['W2V'])
    # default lights (scene/raster.py:90-93)
    L = tina.Lighting()
    L.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
    assert np.array_equal(L.light_dirs[0], g['light_dirs'][0]) and np.array_equal(L.light_colors[0], g['light_colors'][0])


SINKS = {'const': (0, (7, 7, 7)), 'position': (1, None), 'depth': (2, None), 'normal': (3, None), 'viewnormal': (4, None),
         'texcoord': (5, None), 'color': (6, None), 'chessboard': (7, (8, 0, 0)), 'viewdir': (8, None), 'simple': (9, None)}


def test_oracle_gbuffer_sinks_match_reference_shadergroup(tina, O):
    """core/shader.py:21-109 through ShaderGroup (shader.py:138-148): the oracle's sinks against the
    reference's own sources under the shim."""
    g = np.load(os.path.join(GOLDEN, 'gbuffer_shadergroup.npz'))
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    occup, depth, _, _ = O.render_occup(g['verts0'], g['W2V'], W, H, flags, g['bias'])
    assert np.array_equal(occup, g['occup0']) and np.array_equal(depth, g['depth'])
    for name, (kind, param) in SINKS.items():
        ref = g['sink_' + name].astype(np.float32).reshape(W, H, -1)
        out = np.zeros_like(ref)
        O.render_gbuffer(kind, g['verts0'], g['norms0'], g['coors0'], occup, depth, g['W2V'], g['V2W'], W, H, flags, out,
                         param or (0, 0, 0), g['bias'])
        assert np.abs(out - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), name
        assert np.abs(ref).max() > 0, name


def test_accumulator_formula_matches_reference_source():
    """util/accumator.py under the shim == acc * (1 - 1/count) + src * (1/count) in f32 (what k_accumulate does)."""
    if not os.path.isdir('/root/reference/tina'):
        pytest.skip('reference tree not present (GPU box)')
    from oracle import ref_shim
    ref = ref_shim.load_tina('/root/reference')
    import importlib
    acc_mod = importlib.import_module('tina.util.accumator') if False else None
    import sys
    sys.meta_path  # noqa
    rng = np.random.default_rng(0)
    # the shim only preloads the raster path; load the accumulator module the same way
    import types
    src_code = open('/root/reference/tina/util/accumator.py').read().replace('from ..common import *', '')
    ns = dict(vars(sys.modules['tina.common']))
    exec(compile(src_code, 'accumator.py', 'exec'), ns)
    A = ns['Accumator']((5, 4))
    mine = np.zeros((5, 4, 3), np.float32)
    img = sys.modules['taichi'].Vector.field(3, float, (5, 4))
    for k in range(1, 5):
        frame = rng.random((5, 4, 3)).astype(np.float32)
        img.from_numpy(frame)
        A.update(img)
        inv = np.float32(1) / np.float32(k)
        mine = mine * (np.float32(1) - inv) + frame * inv
        assert np.array_equal(A.img.to_numpy(), mine)


def _pars_scene(tina, g):
    """(kind, arrays, material) per object of tests/golden/particles_and_mesh.npz"""
    return [('pars', (g['pverts0'], g['psizes0'], g['pcolors0']), tina.Classic()),
            ('pars', (g['pverts1'], g['psizes1'], g['pcolors1']), tina.Diffuse()),
            ('mesh', (g['verts2'],), tina.Diffuse(color=[0.3, 0.5, 0.9]))]


def test_oracle_particles_match_reference_sources(tina, O):
    """core/particle.py + pars/{simple,trans}.py under the shim, interleaved with a triangle mesh on one depth
    buffer: ids + depth bit-exact after every object, colours <= 2e-6."""
    g = np.load(os.path.join(GOLDEN, 'particles_and_mesh.npz'))
    W, H = (int(v) for v in g['res'])
    lighting = _lighting(tina, g)
    # ParsTransform (pars/trans.py:22-31): positions through mapply_pos, radii scaled
    v1, _ = O.transform((g['pos'][:12] * np.float32(0.5)).reshape(-1, 1, 3), None, g['trans'])
    assert np.array_equal(v1.reshape(-1, 3), g['pverts1'])
    assert np.array_equal(np.float32(1.7) * np.full(12, np.float32(0.05)), g['psizes1'])
    assert np.array_equal(g['pverts0'], g['pos']) and np.array_equal(g['psizes0'], g['rad'])
    depth = O.clear_depth(W, H)
    image = np.zeros((W, H, 3), np.float32)
    for k, (kind, arr, mat) in enumerate(_pars_scene(tina, g)):
        if kind == 'pars':
            occup, depth = O.pars_occup(arr[0], arr[1], g['W2V'], g['V2W'], W, H, True, g['bias'], depth)
            O.pars_color(arr[0], arr[1], arr[2], occup, g['W2V'], g['V2W'], W, H, mat, lighting, image, g['bias'])
        else:
            occup, depth, _, _ = O.render_occup(arr[0], g['W2V'], W, H, O.CULLING | O.CLIPPING, g['bias'], depth)
            O.render_color(arr[0], None, None, occup, g['W2V'], g['V2W'], W, H, O.CULLING | O.CLIPPING, mat, lighting, image, g['bias'])
        assert np.array_equal(occup, g[f'occup{k}']), k
        assert np.array_equal(depth, g[f'depth_after{k}']), k
        assert np.abs(image - g[f'image_after{k}']).max() <= COLOR_TOL, k
    assert (g['occup0'] >= 0).sum() > 100 and (g['occup2'] >= 0).sum() > 100


def test_oracle_wireframe_matches_reference_sources(tina, O):
    """core/wireframe.py + mesh/wire.py under the shim over a solid mesh: depth bit-exact, image identical."""
    g = np.load(os.path.join(GOLDEN, 'particles_wireframe_over_mesh.npz'))
    W, H = (int(v) for v in g['res'])
    lighting = _lighting(tina, g)
    flags = O.CULLING | O.CLIPPING
    occup, depth, _, _ = O.render_occup(g['verts0'], g['W2V'], W, H, flags, g['bias'])
    image = np.zeros((W, H, 3), np.float32)
    O.render_color(g['verts0'], None, None, occup, g['W2V'], g['V2W'], W, H, flags, tina.Diffuse(color=[0.2, 0.3, 0.4]), lighting, image, g['bias'])
    assert np.array_equal(depth, g['depth_after0']) and np.abs(image - g['image_after0']).max() <= COLOR_TOL
    v, _, _ = O.indexed(scenes.load_monkey())
    wires = O.mesh_to_wires(v)
    assert np.array_equal(wires, g['wires1'])  # MeshToWire
    depth, image = O.wire_render(wires, g['W2V'], W, H, depth, image, bias=g['bias'])
    assert np.array_equal(depth, g['depth_after1'])
    assert np.abs(image - g['image_after1']).max() <= COLOR_TOL
    assert (g['depth_after1'] != g['depth_after0']).sum() > 500


def test_oracle_postfx_match_reference_sources(tina, O):
    """postp/fxaa.py, postp/blooming.py under the shim: oracle restatements identical, Gaussian weights too."""
    g = np.load(os.path.join(GOLDEN, 'particles_postfx.npz'))
    assert np.array_equal(O.fxaa(g['input']), g['fxaa'])
    W, H = g['input'].shape[:2]
    gw = tina.Blooming((W, H)).gaussian_weights()
    # (x**2 of a runtime scalar: powf in the shim's numpy vs x*x here; Taichi's own lowering is unpinned) -> 1 ulp
    assert np.abs(gw - g['gwei']).max() <= 2e-8
    assert np.array_equal(O.bloom(g['input'], g['gwei']), g['bloom'])
    assert np.abs(O.bloom(g['input'], gw) - g['bloom']).max() <= 1e-6


def test_oracle_ssr_matches_reference_sources(tina, O):
    """postp/ssr.py run from the reference's own sources under the shim (golden: depth, normal / texcoord / material-id
    G-buffers, three material graphs incl. a textured PBR and an Add / Scale / Mix / Emission composite): the oracle's
    material.sample() trees, Wang-hash stream and ray march reproduce the SSR field, SSR.apply is bit-exact."""
    g = np.load(os.path.join(GOLDEN, 'particles_ssr.npz'))
    mats = [_material(tina, g, i, 'spec') for i in range(int(g['nspecs']))]
    img4 = O.ssr_render(g['depth'], g['normals'], g['coors'], g['mtlid'], mats, g['image_before'], g['W2V'], g['V2W'],
                        nsamples=int(g['nsamples']), nsteps=int(g['nsteps']), stepsize=float(g['stepsize']),
                        tolerance=float(g['tolerance']), blurring=int(g['blurring']))
    assert (g['ssr'][..., 3] > 0).sum() > 300
    assert np.abs(img4 - g['ssr']).max() <= 2e-6
    assert np.array_equal(O.ssr_apply(g['image_before'], g['ssr'], int(g['blurring'])), g['image_after'])


def test_oracle_ssr_default_parameters_match_reference_sources(tina, O):
    """The same with SSR's default parameters (32 samples x 32 steps, ssr.py:20-28) in a scene without texturing (texcoord 0)."""
    g = np.load(os.path.join(GOLDEN, 'particles_ssr_defaults.npz'))
    assert (int(g['nsamples']), int(g['nsteps']), int(g['blurring'])) == (32, 32, 4)
    mats = [_material(tina, g, i, 'spec') for i in range(int(g['nspecs']))]
    img4 = O.ssr_render(g['depth'], g['normals'], None, g['mtlid'], mats, g['image_before'], g['W2V'], g['V2W'],
                        nsamples=32, nsteps=32, stepsize=float(g['stepsize']), tolerance=float(g['tolerance']), blurring=4)
    assert (g['ssr'][..., 3] > 0).sum() > 200
    assert np.abs(img4 - g['ssr']).max() <= 2e-6
    assert np.array_equal(O.ssr_apply(g['image_before'], g['ssr'], 4), g['image_after'])


def test_oracle_ssao_matches_reference_sources(tina, O):
    """postp/ssao.py run from the reference's own sources under the shim (golden: depth, normal G-buffer, the sample /
    rotation tables it drew, the AO field, the image before and after apply) against the oracle restatement."""
    g = np.load(os.path.join(GOLDEN, 'particles_ssao.npz'))
    ao = O.ssao_render(g['depth'], g['normals'], g['W2V'], g['V2W'], g['samples'], g['rotations'])
    assert g['ao'].max() > 0.3 and (g['ao'] > 0).sum() > 100
    assert np.array_equal(ao, g['ao'])
    assert np.array_equal(O.ssao_apply(g['image_before'], g['ao']), g['image_after'])


def test_oracle_material_front_end_on_random_graphs(tina, O):
    """36 random material graphs shaded by the reference's own matr/ + lighting sources (make_golden.py::case_matgraphs):
    the oracle -- with its own flattener, oracle/materials.py -- reproduces every image to 2e-6."""
    g = np.load(os.path.join(GOLDEN, 'matgraphs_random.npz'))
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    lighting = _lighting(tina, g)
    assert int(g['nspecs']) >= 30 and len(g['light_dirs']) == 2
    occup, depth, _, _ = O.render_occup(g['verts0'], g['W2V'], W, H, flags, g['bias'])
    assert np.array_equal(occup, g['occup0']) and np.array_equal(depth, g['depth'])
    kinds = set()
    for i in range(int(g['nspecs'])):
        image = np.zeros((W, H, 3), np.float32)
        O.render_color(g['verts0'], g['norms0'], g['coors0'], occup, g['W2V'], g['V2W'], W, H, flags, _material(tina, g, i, 'spec'),
                       lighting, image, g['bias'])
        ref = g[f'image{i}']
        err = (np.abs(image - ref) / np.maximum(1.0, np.abs(ref))).max()  # (sharp Cook-Torrance lobes reach ~30: 1 ulp there is 2e-6)
        assert err <= COLOR_TOL, (i, str(g[f'spec{i}']), err)
        kinds |= {n for n in MATERIAL_NAMES if n in str(g[f'spec{i}'])}
    assert kinds >= {'MixMaterial', 'ScaleMaterial', 'AddMaterial', 'CookTorrance', 'Phong', 'Texture', 'FresnelFactor', 'PBR', 'Classic'}


def test_oracle_procedural_textures_match_reference_sources(tina, O):
    """ChessboardTexture / LerpTexture (matr/nodes.py:114-136) as colours, mix factors and roughness, shaded by the reference's
    own sources (make_golden.py::case_proc_textures): the oracle reproduces every image to 2e-6."""
    g = np.load(os.path.join(GOLDEN, 'matgraphs_proc.npz'))
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    lighting = _lighting(tina, g)
    occup, depth, _, _ = O.render_occup(g['verts0'], g['W2V'], W, H, flags, g['bias'])
    assert np.array_equal(occup, g['occup0']) and np.array_equal(depth, g['depth'])
    for i in range(int(g['nspecs'])):
        image = np.zeros((W, H, 3), np.float32)
        O.render_color(g['verts0'], g['norms0'], g['coors0'], occup, g['W2V'], g['V2W'], W, H, flags, _material(tina, g, i, 'spec'),
                       lighting, image, g['bias'])
        ref = g[f'image{i}']
        err = (np.abs(image - ref) / np.maximum(1.0, np.abs(ref))).max()
        assert err <= COLOR_TOL, (i, str(g[f'spec{i}']), err)


def test_oracle_setup_cache_matches_reference(O):
    """TriangleRaster.bcn / can / boo / coo / wsc as the reference's render_occup stored them (triangle.py:127-131)."""
    g = np.load(os.path.join(GOLDEN, 'setup_cache_monkey.npz'))
    W, H = (int(v) for v in g['res'])
    out, _ = O.face_setup(g['verts0'], g['W2V'], W, H, int(g['flags']))
    ok = out[:, 0] != 0
    ref = np.concatenate([g['bcn'], g['can'], g['boo'], g['coo'], g['wsc']], axis=1)
    assert 300 < ok.sum() < len(ok)
    assert np.array_equal(out[ok, 1:12], ref[ok])
    assert not ref[~ok].any()  # rejected faces: never written
