"""GPU parity tests: the CUDA path (through the Python mirror -> C ABI -> sm_100a kernels)
against the CPU oracle on the same seeded inputs.

Bars: face-id (`occup`) and integer `depth` buffers BIT-EXACT vs the serial
oracle, exact-depth ties included; colour within 1e-4 abs in fp32, before and after tonemap.
"""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu

COLOR_TOL = 1e-4  # north_star: "colour buffers match within 1e-4 abs in fp32"


def _flags(O, smoothing=False, texturing=False, culling=True, clipping=True):
    return ((O.SMOOTHING if smoothing else 0) | (O.TEXTURING if texturing else 0) | (O.CULLING if culling else 0) |
            (O.CLIPPING if clipping else 0))


def _check_frame(scene, ref, tol=COLOR_TOL):
    import torch
    torch.cuda.synchronize()
    depth = scene.engine.depth.to_numpy()
    occup = scene.triangle_raster.occup.to_numpy()
    assert np.array_equal(depth, ref['depth']), f"depth differs at {(depth != ref['depth']).sum()} px"
    assert np.array_equal(occup, ref['occups'][-1]), f"occup differs at {(occup != ref['occups'][-1]).sum()} px"
    img = scene.img.to_numpy()
    err = np.abs(img - ref['image']).max()
    assert err <= tol, f'colour max abs err {err}'
    return err


def test_monkey_flat_diffuse(tina, O):
    """C1: docs/monkey.py -- Suzanne, 512x512, flat shading, default light/material/camera."""
    obj = scenes.load_monkey()
    scene = tina.Scene()
    scene.add_object(tina.MeshModel(obj))
    view, proj = scenes.default_camera()
    scene.engine.set_camera(view, proj)
    scene.render()
    v, _, _ = O.indexed(obj)
    ref = O.render_scene([(v, None, None, tina.Diffuse())], 512, 512, view, proj, scene.lighting, _flags(O))
    _check_frame(scene, ref)
    assert (ref['depth'] < 2**30).sum() == 64082  # SURVEY §8c probe
    # pre-tonemap image too: render without the fused tonemap
    scene2 = tina.Scene(tonemap=False)
    scene2.add_object(tina.MeshModel(obj))
    scene2.engine.set_camera(view, proj)
    scene2.render()
    assert np.abs(scene2.img.to_numpy() - ref['pre_tonemap']).max() <= COLOR_TOL


@pytest.mark.parametrize('nocull', [False, True])
def test_meshgrid_wave_smooth_classic(tina, O, nocull):
    """C2-style at test size: MeshGrid wave, smooth normals, Classic (Lambert + Phong); ties present."""
    n, W, H = 96, 640, 360
    pos = scenes.wave_grid_pos(n)
    scene = tina.Scene((W, H), smoothing=True)
    grid = tina.MeshGrid(n)
    grid.pos.from_numpy(pos)
    scene.add_object(tina.MeshNoCulling(grid) if nocull else grid, tina.Classic())
    view, proj = scenes.default_camera(W / H)
    scene.engine.set_camera(view, proj)
    scene.render()
    fv, fn = O.grid_faces(pos), O.grid_faces(O.grid_normals(pos))
    if nocull:
        fv, fn, _ = O.no_culling(fv, fn)
    ref = O.render_scene([(fv, fn, None, tina.Classic())], W, H, view, proj, scene.lighting, _flags(O, smoothing=True))
    _check_frame(scene, ref)
    # set_object itself: raster.verts / raster.norms equal the oracle's expansion bit for bit
    assert np.array_equal(scene.triangle_raster.verts.to_numpy(), fv)
    assert np.array_equal(scene.triangle_raster.norms.to_numpy(), fn)


def test_cornell_gltf_pbr_textured(tina, O):
    """C4 at test size: cornell.gltf, 3 objects, smoothing + texturing, PBR (CookTorrance + texture)."""
    W = H = 256
    for k, (view, proj) in enumerate(scenes.cornell_views(4)):
        gltf = scenes.load_cornell()
        scene = tina.Scene((W, H), smoothing=True, texturing=True)
        gltf.extract(scene)
        assert len(scene.objects) == 3
        scene.engine.set_camera(view, proj)
        scene.render()
        ref = O.render_scene(scenes.cornell_oracle_objects(gltf), W, H, view, proj, scene.lighting,
                             _flags(O, smoothing=True, texturing=True))
        _check_frame(scene, ref)
        assert (ref['depth'] < 2**30).mean() > 0.5


def _render_soup(tina, tri, W, H, view, proj, **tuning):
    scene = tina.Scene((W, H), maxfaces=len(tri))
    mesh = tina.SimpleMesh(maxfaces=len(tri))
    mesh.set_face_verts(tri)
    scene.add_object(mesh)
    scene.engine.set_camera(view, proj)
    scene.triangle_raster.set_tuning(**tuning)
    scene.render()
    return scene


@pytest.mark.parametrize('tuning', [dict(), dict(tiny_max=0), dict(tiny_max=4), dict(tiny_max=100000), dict(force_tiles=1),
                                    dict(balance=0), dict(balance=2), dict(balance=2, tiny_max=2000), dict(balance=0, tiny_max=2000),
                                    dict(precheck=1, balance=2)])
def test_soup_depth_complexity_all_strategies(tina, O, tuning):
    """C3-style at test size: random soup with depth complexity ~8; every rasteriser strategy
    (per-thread direct / binned tile path / mixtures) must give the same bits."""
    W, H, n = 320, 200, 60000
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup(n, W, H, s=0.012, seed=11)
    scene = _render_soup(tina, tri, W, H, view, proj, **tuning)
    ref = O.render_scene([(tri, None, None, tina.Diffuse())], W, H, view, proj, scene.lighting, _flags(O))
    _check_frame(scene, ref)
    assert ref['ties'][0].sum() >= 0


def test_mixed_sizes_and_edge_cases(tina, O):
    """Triangles behind the camera (w <= 0), NaN / inf vertices, zero-area, off-screen, screen-filling
    with every vertex outside the NDC cube (dropped by the per-vertex clip test, triangle.py:100-104),
    with clipping off, and with culling off."""
    W, H = 200, 136
    view, proj = scenes.default_camera(W / H)
    rng = np.random.default_rng(3)
    tri = scenes.soup(3000, W, H, s=0.05, seed=5)
    extra = np.array([
        [[-9, -9, 0], [9, -9, 0], [0, 9, 0]],            # screen-filling, all vertices outside
        [[-0.5, -0.5, 5], [0.5, -0.5, 5], [0, 0.5, 5]],  # behind the camera
        [[-0.5, -0.5, 0], [0.5, -0.5, 4], [0, 0.5, 0]],  # crosses w = 0
        [[0, 0, 0], [0, 0, 0], [0, 0, 0]],               # degenerate
        [[np.nan, 0, 0], [1, 0, 0], [0, 1, 0]],          # NaN
        [[np.inf, 0, 0], [1, 0, 0], [0, 1, 0]],          # inf
        [[50, 50, 0], [51, 50, 0], [50, 51, 0]],         # far off-screen
        [[-1, -1, 0.5], [1, -1, 0.5], [0, 1, 0.5]],      # big, front-facing
        [[-1, -1, 0.2], [0, 1, 0.2], [1, -1, 0.2]],      # big, back-facing
        [[1e-3, 0, 1], [2e-3, 0, 1], [1e-3, 1e-3, 1]],   # sub-pixel
    ], dtype=np.float32)
    tri = np.ascontiguousarray(np.concatenate([extra, tri, extra[::-1]]))
    for culling in (True, False):
        for clipping in (True, False):
            scene = tina.Scene((W, H), culling=culling, clipping=clipping)
            mesh = tina.SimpleMesh()
            mesh.set_face_verts(tri)
            scene.add_object(mesh)
            scene.engine.set_camera(view, proj)
            scene.render()
            with np.errstate(all='ignore'):
                ref = O.render_scene([(tri, None, None, tina.Diffuse())], W, H, view, proj, scene.lighting,
                                     _flags(O, culling=culling, clipping=clipping))
            import torch
            torch.cuda.synchronize()
            assert np.array_equal(scene.engine.depth.to_numpy(), ref['depth']), (culling, clipping)
            assert np.array_equal(scene.triangle_raster.occup.to_numpy(), ref['occups'][-1]), (culling, clipping)
            img, rimg = scene.img.to_numpy(), ref['image']
            ok = np.isfinite(rimg)
            assert np.array_equal(np.isfinite(img), ok)
            assert np.abs(img[ok] - rimg[ok]).max() <= COLOR_TOL


def test_empty_mesh_and_empty_scene(tina, O):
    scene = tina.Scene((64, 48), bgcolor=[0.1, 0.2, 0.3])
    scene.render()
    img = scene.img.to_numpy()
    assert np.allclose(img, O.tonemap(np.broadcast_to(np.float32([0.1, 0.2, 0.3]), (64, 48, 3))), atol=1e-6)
    mesh = tina.SimpleMesh()
    mesh.set_face_verts(np.zeros((0, 3, 3), np.float32))
    scene.add_object(mesh)
    scene.render()
    assert (scene.engine.depth.to_numpy() == 2**30).all()
    assert (scene.triangle_raster.occup.to_numpy() == -1).all()


def test_multi_object_depth_carry_and_inter_object_ties(tina, O):
    """Depth persists across objects, occup is per object; at exact inter-object depth ties the earlier
    object keeps the pixel (strict `>` in triangle.py:123)."""
    W, H = 160, 120
    view, proj = scenes.default_camera(W / H)
    a = scenes.soup(500, W, H, s=0.08, seed=1)
    b = scenes.soup(500, W, H, s=0.08, seed=2)
    objs = [(a, tina.Diffuse(color=[1, 0, 0])), (b, tina.Diffuse(color=[0, 1, 0])), (a.copy(), tina.Diffuse(color=[0, 0, 1]))]
    scene = tina.Scene((W, H))
    for tri, mat in objs:
        m = tina.SimpleMesh()
        m.set_face_verts(tri)
        scene.add_object(m, mat)
    scene.engine.set_camera(view, proj)
    scene.render()
    ref = O.render_scene([(t, None, None, m) for t, m in objs], W, H, view, proj, scene.lighting, _flags(O))
    _check_frame(scene, ref)
    assert (ref['occups'][-1] == -1).all()  # the third object is a copy of the first: it never wins (ties lose)


def test_zero_copy_torch_inputs_and_direct_setters(tina, O):
    """north_star: torch tensors accepted zero-copy; TriangleRaster.set_face_verts/norms fast path."""
    import torch
    W, H = 128, 96
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup(2000, W, H, s=0.03, seed=9)
    nrm = np.random.default_rng(1).normal(size=tri.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    engine = tina.Engine((W, H))
    raster = tina.TriangleRaster(engine, smoothing=True)
    tv, tn = torch.as_tensor(tri).cuda(), torch.as_tensor(nrm).cuda()
    raster.set_face_verts(tv)
    raster.set_face_norms(tn)
    assert raster.verts.to_torch().data_ptr() == tv.data_ptr()  # aliased, not copied
    engine.set_camera(view, proj)
    engine.clear_depth()
    raster.render_occup()
    lighting = tina.Lighting()
    lighting.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
    lighting.set_ambient_light([0.1, 0.1, 0.1])
    img = tina.Field(torch.zeros((W, H, 3), device='cuda'))
    raster.render_color(tina.Shader(img, lighting, tina.Classic()))
    ref = O.render_scene([(tri, nrm, None, tina.Classic())], W, H, view, proj, lighting, _flags(O, smoothing=True),
                         do_tonemap=False)
    torch.cuda.synchronize()
    assert np.array_equal(raster.occup.to_numpy(), ref['occups'][0])
    assert np.array_equal(engine.depth.to_numpy(), ref['depth'])
    assert np.abs(img.to_numpy() - ref['image']).max() <= COLOR_TOL


def test_transform_flip_and_smooth_normal_adapters(tina, O):
    """MeshTransform / MeshFlipCulling / MeshFlipNormal over MeshModel (mesh/trans.py, mesh/cull.py)."""
    W, H = 200, 200
    obj = scenes.load_monkey()
    trans = tina.translate([0.2, -0.1, 0.3]) @ tina.eularXYZ([0.3, 0.8, -0.2]) @ tina.scale([0.9, 1.1, 0.8])
    view, proj = scenes.default_camera()
    scene = tina.Scene((W, H), smoothing=True, texturing=True)
    scene.add_object(tina.MeshFlipNormal(tina.MeshFlipCulling(tina.MeshTransform(tina.MeshModel(obj), trans))), tina.Classic())
    scene.engine.set_camera(view, proj)
    scene.render()
    v, vn, vt = O.indexed(obj)
    v, vn = O.transform(v, vn, trans)
    v, vn, vt = v[:, ::-1].copy(), -vn[:, ::-1].copy(), vt[:, ::-1].copy()
    ref = O.render_scene([(v, vn, vt, tina.Classic())], W, H, view, proj, scene.lighting,
                         _flags(O, smoothing=True, texturing=True))
    _check_frame(scene, ref)
    assert np.array_equal(scene.triangle_raster.verts.to_numpy(), v)
    assert np.array_equal(scene.triangle_raster.norms.to_numpy(), vn)
    assert np.array_equal(scene.triangle_raster.coors.to_numpy(), vt)


def test_maxfaces_overflow_raises(tina):
    scene = tina.Scene((32, 32), maxfaces=10)
    m = tina.SimpleMesh()
    m.set_face_verts(np.zeros((11, 3, 3), np.float32))
    scene.add_object(m)
    with pytest.raises(ValueError):
        scene.render()


def test_c2_full_size_bit_exact(tina, O):
    """C2 at BASELINE size: MeshGrid(1024) wave, 2,093,058 faces, 1920x1080, smooth + Classic."""
    n, W, H = 1024, 1920, 1080
    pos = scenes.wave_grid_pos(n)
    scene = tina.Scene((W, H), smoothing=True, maxfaces=2**21)
    grid = tina.MeshGrid(n)
    grid.pos.from_numpy(pos)
    scene.add_object(grid, tina.Classic())
    view, proj = scenes.default_camera(W / H)
    scene.engine.set_camera(view, proj)
    scene.render()
    fv, fn = O.grid_faces(pos), O.grid_faces(O.grid_normals(pos))
    ref = O.render_scene([(fv, fn, None, tina.Classic())], W, H, view, proj, scene.lighting, _flags(O, smoothing=True))
    _check_frame(scene, ref)


def _keys(scene):
    import torch
    torch.cuda.synchronize()
    return scene.engine.keys.cpu().numpy().copy()


def _adversarial_triangles(rng, n, W, H, view, proj):
    """Screen-space recipes aimed at the candidate-tightening guards: sub-pixel triangles whose
    vertices sit within 1e-3..1e-1 px of sample centres, needles / slivers pointing at samples,
    near-degenerate areas, triangles hugging the guard thresholds -- unprojected to world space."""
    W2V = np.asarray(proj, np.float64) @ np.asarray(view, np.float64)
    V2W = np.linalg.inv(W2V)
    kind = rng.integers(0, 6, n)
    cx = rng.integers(2, W - 2, n) + 0.5
    cy = rng.integers(2, H - 2, n) + 0.5
    ang = rng.uniform(0, 2 * np.pi, (n, 3))
    rad = np.empty((n, 3))
    off = np.zeros((n, 2))
    # 0: tiny around a sample; 1: tiny just beside a sample; 2: needle pointing at a sample;
    # 3: sliver across a sample; 4: ~1-3 px ordinary; 5: near-collinear
    rad[:] = rng.uniform(0.05, 0.6, (n, 3))
    m = kind == 1
    off[m] = rng.choice([-1, 1], (m.sum(), 2)) * 10.0 ** rng.uniform(-3, -0.5, (m.sum(), 2)) + rng.uniform(0.3, 0.7, (m.sum(), 2)) * rng.choice([-1, 1], (m.sum(), 2))
    m = kind == 4
    rad[m] = rng.uniform(0.5, 3.0, (m.sum(), 3))
    pts = np.stack([cx[:, None] + off[:, :1] + rad * np.cos(ang), cy[:, None] + off[:, 1:] + rad * np.sin(ang)], axis=2)
    m = kind == 2  # needle: two vertices far away and close together, tip near the sample
    k = m.sum()
    d = rng.uniform(0, 2 * np.pi, k)
    ln = rng.uniform(1.0, 6.0, k)
    wdt = 10.0 ** rng.uniform(-4, -1, k)
    tip = np.stack([cx[m], cy[m]], 1) + rng.normal(0, 1, (k, 2)) * 10.0 ** rng.uniform(-4, -1, (k, 1))
    dirv = np.stack([np.cos(d), np.sin(d)], 1)
    nrm = np.stack([-np.sin(d), np.cos(d)], 1)
    pts[m, 0] = tip
    pts[m, 1] = tip + dirv * ln[:, None] + nrm * wdt[:, None]
    pts[m, 2] = tip + dirv * ln[:, None] - nrm * wdt[:, None]
    m = kind == 3  # sliver through the sample
    k = m.sum()
    d = rng.uniform(0, 2 * np.pi, k)
    dirv = np.stack([np.cos(d), np.sin(d)], 1)
    nrm = np.stack([-np.sin(d), np.cos(d)], 1)
    c0 = np.stack([cx[m], cy[m]], 1) + nrm * (rng.normal(0, 1, (k, 1)) * 10.0 ** rng.uniform(-5, -1, (k, 1)))
    ln = rng.uniform(0.5, 4.0, (k, 1))
    pts[m, 0] = c0 - dirv * ln
    pts[m, 1] = c0 + dirv * ln
    pts[m, 2] = c0 + dirv * rng.uniform(-1, 1, (k, 1)) * ln + nrm * 10.0 ** rng.uniform(-5, -1, (k, 1))
    m = kind == 5
    k = m.sum()
    pts[m, 2] = 0.5 * (pts[m, 0] + pts[m, 1]) + rng.normal(0, 1, (k, 2)) * 10.0 ** rng.uniform(-7, -2, (k, 1))
    # random winding, then unproject at random depths
    flip = rng.random(n) < 0.3
    pts[flip] = pts[flip][:, [0, 2, 1]]
    ndc = np.empty((n, 3, 4))
    ndc[..., 0] = pts[..., 0] / W * 2 - 1
    ndc[..., 1] = pts[..., 1] / H * 2 - 1
    dist = rng.uniform(1.0, 6.0, (n, 3))
    ndc[..., 2] = (proj[2, 2] * (-dist) + proj[2, 3]) / dist
    ndc[..., 3] = 1
    wp = ndc @ V2W.T
    return np.ascontiguousarray((wp[..., :3] / wp[..., 3:4]).astype(np.float32))


@pytest.mark.parametrize('seed,bias', [(1, (0.5, 0.5)), (2, (0.5, 0.5)), (3, (0.123, 0.877)), (4, (0.0, 1.0))])
def test_tightening_is_exact(tina, O, seed, bias):
    """Candidate tightening (k_raster_faces phase A) must never drop a pixel the reference accepts:
    tightened == untightened key buffers on adversarial triangle sets, and == the oracle."""
    W, H, n = 256, 192, 400000
    view, proj = scenes.default_camera(W / H)
    tri = _adversarial_triangles(np.random.default_rng(seed), n, W, H, view, proj)
    out = []
    for culling in (True, False):
        for tighten in (1, 0):
            scene = tina.Scene((W, H), maxfaces=n, culling=culling)
            mesh = tina.SimpleMesh(maxfaces=n)
            mesh.set_face_verts(tri)
            scene.add_object(mesh)
            scene.engine.set_camera(view, proj)
            scene.engine.bias[None] = bias
            scene.triangle_raster.set_tuning(tighten=tighten, tiny_max=64, balance=2 if tighten else 0)
            scene.render()
            out.append(_keys(scene))
        assert np.array_equal(out[-1], out[-2]), f'tightening changed {(out[-1] != out[-2]).sum()} keys (culling={culling})'
    occup, depth, _, st = O.render_occup(tri, (proj @ view).astype(np.float32), W, H, O.CULLING | O.CLIPPING, bias=bias)
    assert np.array_equal(out[0] >> 32, depth.astype(np.int64))
    assert np.array_equal((out[0] & 0xffffffff) - 1, occup.astype(np.int64))
    assert st['covered'] > 10000


def test_tile_path_modes_agree(tina, O):
    """Small queues are rasterised by per-tile bbox scanning, large ones through the binned lists
    (count -> prefix sum -> scatter); list overflow falls back to scanning.  All identical."""
    W, H = 384, 256
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup(1500, W, H, s=0.12, seed=21)  # medium/large triangles
    ref = None
    for tuning in (dict(), dict(scan_max=0), dict(scan_max=100000), dict(force_tiles=1, scan_max=0), dict(tiny_max=0, scan_max=1)):
        scene = _render_soup(tina, tri, W, H, view, proj, **tuning)
        k = _keys(scene)
        if ref is None:
            ref = k
            r = O.render_scene([(tri, None, None, tina.Diffuse())], W, H, view, proj, scene.lighting, _flags(O))
            _check_frame(scene, r)
        assert np.array_equal(k, ref), tuning


def test_specialised_shading_equals_interpreter(tina, O):
    """The Diffuse / Classic / PBR shading kernels and the host constant folding give the same
    bits as interpreting the unfolded material program."""
    import torch
    W, H = 160, 120
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup(800, W, H, s=0.1, seed=5)
    nrm = np.random.default_rng(2).normal(size=tri.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    mats = [tina.Diffuse(), tina.Diffuse(color=[0.2, 0.5, 0.9]), tina.Classic(), tina.Classic(color=[0.9, 0.3, 0.1], shineness=8, specular=0.7),
            tina.PBR(basecolor=[0.8, 0.7, 0.6], metallic=0.9, roughness=0.05), tina.PBR(),
            tina.Lambert() * [1, 0, 0] + tina.Emission() * 0.25 + tina.Phong(shineness=[4, 16, 64]) * 0.3]
    for mat in mats:
        imgs = []
        for generic in (0, 1):
            scene = tina.Scene((W, H), smoothing=True, tonemap=False)
            mesh = tina.SimpleMesh()
            mesh.set_face_verts(tri)
            mesh.set_face_norms(nrm)
            scene.add_object(mesh, mat)
            scene.engine.set_camera(view, proj)
            scene.lighting.add_light(pos=[0.5, 0.5, 2.0], color=[0.3, 0.6, 0.9])
            scene.triangle_raster.set_tuning(generic_vm=generic, fast_shading=0)
            scene.render()
            torch.cuda.synchronize()
            imgs.append(scene.img.to_numpy())
        assert np.array_equal(imgs[0], imgs[1])
        ref = O.render_scene([(tri, nrm, None, mat)], W, H, view, proj, scene.lighting, _flags(O, smoothing=True), do_tonemap=False)
        assert np.abs(imgs[0] - ref['image']).max() <= COLOR_TOL
        # default (fast) shading: same pixels, colour inside the tolerance
        scene.triangle_raster.set_tuning(generic_vm=0, fast_shading=1)
        scene.render()
        torch.cuda.synchronize()
        assert np.abs(scene.img.to_numpy() - ref['image']).max() <= COLOR_TOL


def test_fast_shading_stays_within_colour_tolerance(tina, O):
    """render_color's default arithmetic (FMA + SFU rcp/rsqrt downstream of the exact barycentric weights) against
    its exact arm and against the oracle, on the C2-style grid (Classic) and a sliver-rich soup (Diffuse, flat)."""
    import torch
    W, H = 640, 360
    view, proj = scenes.default_camera(W / H)
    worst = 0.0
    for kind in ('grid', 'soup'):
        imgs = {}
        for fast in (0, 1):
            if kind == 'grid':
                scene = tina.Scene((W, H), smoothing=True)
                mesh = tina.MeshGrid(96)
                mesh.pos.from_numpy(scenes.wave_grid_pos(96))
                scene.add_object(mesh, tina.Classic())
            else:
                scene = tina.Scene((W, H))
                mesh = tina.SimpleMesh()
                mesh.set_face_verts(scenes.soup(5000, W, H, s=0.02, seed=11))
                scene.add_object(mesh, tina.Diffuse())
            scene.engine.set_camera(view, proj)
            scene.triangle_raster.set_tuning(fast_shading=fast)
            scene.render()
            torch.cuda.synchronize()
            imgs[fast] = scene.img.to_numpy()
            keys = scene.engine.keys.clone()
            if fast == 0:
                keys0 = keys
        assert torch.equal(keys, keys0)
        d = float(np.abs(imgs[0] - imgs[1]).max())
        worst = max(worst, d)
        assert d <= 2e-5, (kind, d)
        # the lean kernels (compile-time raster flags, constant operands) are the same arithmetic: same bits
        scene.triangle_raster.set_tuning(fast_shading=1, lean_kernels=0)
        scene.render()
        torch.cuda.synchronize()
        assert np.array_equal(scene.img.to_numpy(), imgs[1]), kind
        assert torch.equal(scene.engine.keys, keys0), kind
    print('fast-vs-exact shading max abs colour difference', worst)


def _golden_cases():
    import glob
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    return sorted(p for p in glob.glob(os.path.join(d, '*.npz')) if not os.path.basename(p).startswith(('particles', 'matgraphs')))


@pytest.mark.parametrize('path', _golden_cases(), ids=lambda p: p.split('/')[-1][:-4])
def test_cuda_path_matches_reference_goldens(tina, path):
    """The CUDA path against golden vectors produced by the reference's own sources
    (tests/golden/make_golden.py): ids + depth bit-exact, colour within 1e-4."""
    import torch
    from test_golden import _lighting, _material
    g = np.load(path)
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    engine = tina.Engine((W, H))
    engine.W2V[None] = g['W2V']
    engine.V2W[None] = g['V2W']
    engine.bias[None] = g['bias']
    raster = tina.TriangleRaster(engine, smoothing=bool(flags & 1), texturing=bool(flags & 2), culling=bool(flags & 4),
                                 clipping=bool(flags & 8))
    lighting = _lighting(tina, g)
    img = tina.Field(torch.empty((W, H, 3), device='cuda'))
    img.fill(float(g['bgcolor']))
    engine.clear_depth()
    for k in range(int(g['nobjects'])):
        mesh = tina.SimpleMesh()
        mesh.set_face_verts(g[f'verts{k}'])
        if f'norms{k}' in g:
            mesh.set_face_norms(g[f'norms{k}'])
        if f'coors{k}' in g:
            mesh.set_face_coors(g[f'coors{k}'])
        raster.set_object(mesh)
        raster.render_occup()
        raster.render_color(tina.Shader(img, lighting, _material(tina, g, k)))
        torch.cuda.synchronize()
        assert np.array_equal(raster.occup.to_numpy(), g[f'occup{k}']), f'occup of object {k}'
    assert np.array_equal(engine.depth.to_numpy(), g['depth'])
    out = img.to_numpy()
    ok = np.isfinite(g['image_pre_tonemap'])
    assert np.array_equal(np.isfinite(out), ok)
    assert np.abs(out[ok] - g['image_pre_tonemap'][ok]).max() <= COLOR_TOL


@pytest.mark.parametrize('case', ['grid', 'grid_nonsquare', 'grid_nocull_trans', 'model', 'model_trans_flip', 'model_nocull_big'])
def test_indexed_vertex_stage_equals_expanded_path(tina, O, case):
    """MeshGrid / MeshModel sources go through the per-unique-vertex stage (world + clip coordinates per
    vertex, corners fetched through the mesh's own indexing); it must give exactly the bits of the
    expanded-array path (the reference's raster.verts/norms/coors), for keys and colours."""
    import torch
    W, H = 320, 240
    view, proj = tina.orbit_camera(radius=3.0, theta=0.35, phi=0.6, aspect=W / H)
    trans = tina.translate([0.1, -0.2, 0.1]) @ tina.eularXYZ([0.4, -0.3, 0.2]) @ tina.scale([1.3, 0.8, 1.1])

    def build():
        if case.startswith('grid'):
            g = tina.MeshGrid((40, 56) if case == 'grid_nonsquare' else 48)
            pos = g.pos.to_numpy()
            pos[..., 2] = 0.15 * np.sin(7 * pos[..., 0]) * np.cos(5 * pos[..., 1])
            g.pos.from_numpy(pos)
            m = g
            if case == 'grid_nocull_trans':
                m = tina.MeshNoCulling(tina.MeshTransform(g, trans))
        else:
            m = tina.MeshModel(scenes.load_monkey())
            if case == 'model_trans_flip':
                m = tina.MeshFlipNormal(tina.MeshFlipCulling(tina.MeshTransform(m, trans)))
            if case == 'model_nocull_big':
                m = tina.MeshNoCulling(tina.MeshTransform(m, tina.scale(2.5)))  # big triangles -> tile path too
        return m

    res = []
    for indexed in (1, 0):
        scene = tina.Scene((W, H), smoothing=True, texturing=True, tonemap=False)
        img = np.random.default_rng(0).random((8, 8, 3)).astype(np.float32)
        mat = tina.PBR(basecolor=tina.Texture(img), metallic=0.3, roughness=0.4) if 'trans' in case else tina.Classic()
        scene.add_object(build(), mat)
        scene.engine.set_camera(view, proj)
        scene.triangle_raster.set_tuning(indexed=indexed)
        scene.render()
        torch.cuda.synchronize()
        res.append((_keys(scene), scene.img.to_numpy(), scene.triangle_raster.verts.to_numpy(),
                    scene.triangle_raster.norms.to_numpy(), scene.triangle_raster.coors.to_numpy()))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert ((res[0][0] & 0xffffffff) != 0).sum() > 2000


def test_adaptive_tile_path_skipping_stays_exact(tina, O):
    """After 8 consecutive frames without large faces the idle tile-path kernel is no longer launched and
    k_raster_faces walks large faces itself.  A frame that suddenly contains big triangles must still be
    exact, and so must the frames after it (tile path back on)."""
    import torch
    W, H = 256, 160
    view, proj = scenes.default_camera(W / H)
    small = scenes.soup(20000, W, H, s=0.01, seed=3)
    big = np.concatenate([scenes.soup(3000, W, H, s=0.01, seed=4),
                          np.float32([[[-1.5, -1, 0.3], [1.5, -1, 0.3], [0, 1.2, -0.2]], [[-1, 0.8, 0.5], [-1, -0.8, 0.5], [1.1, 0, -0.4]]]),
                          scenes.soup(300, W, H, s=0.15, seed=6)]).astype(np.float32)
    scene = tina.Scene((W, H), maxfaces=len(small))
    mesh = tina.SimpleMesh(maxfaces=len(small))
    scene.add_object(mesh)
    scene.engine.set_camera(view, proj)
    refs = {}
    for name, tri in (('small', small), ('big', big)):
        refs[name] = O.render_scene([(tri, None, None, tina.Diffuse())], W, H, view, proj, scene.lighting, _flags(O))
    seq = ['small'] * 12 + ['big', 'big', 'small', 'big'] + ['small'] * 10 + ['big']
    for name in seq:
        mesh.set_face_verts(small if name == 'small' else big)
        scene.render()
        torch.cuda.synchronize()
        _check_frame(scene, refs[name])


def test_gbuffer_shadergroup_matches_reference_golden(tina):
    """§8f row 1: G-buffer shaders + ShaderGroup fan-out on the CUDA path against the golden produced by
    the reference's own ShaderGroup (tests/golden/make_golden.py::case_gbuffers)."""
    import os
    import torch
    from test_golden import GOLDEN, SINKS, _lighting
    g = np.load(os.path.join(GOLDEN, 'gbuffer_shadergroup.npz'))
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    engine = tina.Engine((W, H))
    engine.W2V[None], engine.V2W[None], engine.bias[None] = g['W2V'], g['V2W'], g['bias']
    raster = tina.TriangleRaster(engine, smoothing=bool(flags & 1), texturing=bool(flags & 2))
    classes = {'const': lambda b: tina.ConstShader(b, 7), 'position': tina.PositionShader, 'depth': tina.DepthShader,
               'normal': tina.NormalShader, 'viewnormal': tina.ViewNormalShader, 'texcoord': tina.TexcoordShader,
               'color': tina.ColorShader, 'chessboard': lambda b: tina.ChessboardShader(b, 8), 'viewdir': tina.ViewdirShader,
               'simple': tina.SimpleShader}
    bufs, shaders = {}, []
    img = tina.Field(torch.zeros((W, H, 3), device='cuda'))
    shaders.append(tina.Shader(img, _lighting(tina, g), tina.Classic()))
    for name in SINKS:
        ref = g['sink_' + name]
        t = torch.zeros(ref.shape, device='cuda', dtype=torch.int32 if ref.dtype.kind == 'i' else torch.float32)
        bufs[name] = t
        shaders.append(classes[name](tina.Field(t)))
    mesh = tina.SimpleMesh()
    mesh.set_face_verts(g['verts0'])
    mesh.set_face_norms(g['norms0'])
    mesh.set_face_coors(g['coors0'])
    engine.clear_depth()
    raster.set_object(mesh)
    raster.render_occup()
    raster.render_color(tina.ShaderGroup(shaders))
    torch.cuda.synchronize()
    assert np.array_equal(raster.occup.to_numpy(), g['occup0'])
    assert np.abs(img.to_numpy() - g['image_pre_tonemap']).max() <= COLOR_TOL
    for name in SINKS:
        ref = g['sink_' + name].astype(np.float64)
        out = bufs[name].cpu().numpy().astype(np.float64)
        assert np.abs(out - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), name
    # the same through the indexed (MeshModel) path
    scene = tina.Scene((W, H), smoothing=True, texturing=True)
    nb = torch.zeros((W, H, 3), device='cuda')
    scene.post_shaders.append(tina.NormalShader(tina.Field(nb)))
    scene.add_object(tina.MeshModel(scenes.load_monkey()), tina.Classic())
    scene.engine.W2V[None], scene.engine.V2W[None] = g['W2V'], g['V2W']
    scene.render()
    torch.cuda.synchronize()
    assert np.abs(nb.cpu().numpy() - g['sink_normal']).max() <= 1e-6


def test_taa_accumulation(tina, O):
    """§8f row 2: Scene(taa=True) = jittered bias + Accumator.update (util/accumator.py:16-23);
    every jittered frame must still match the oracle, and the running mean uses the reference's f32 ops."""
    import torch
    W, H = 96, 72
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup(800, W, H, s=0.08, seed=12)
    scene = tina.Scene((W, H), taa=True)
    mesh = tina.SimpleMesh()
    mesh.set_face_verts(tri)
    scene.add_object(mesh)
    scene.engine.set_camera(view, proj)
    acc = np.zeros((W, H, 3), np.float32)
    for k in range(1, 6):
        scene.render()
        torch.cuda.synchronize()
        bias = scene.engine.bias.to_numpy()
        assert (k == 1 and np.array_equal(bias, np.float32([0.5, 0.5]))) or (k > 1 and 0 <= bias.min() and bias.max() < 1)
        ref = O.render_scene([(tri, None, None, tina.Diffuse())], W, H, view, proj, scene.lighting, _flags(O), bias=bias)
        assert np.array_equal(scene.engine.depth.to_numpy(), ref['depth'])
        frame = scene.pp_img.to_numpy()
        assert np.abs(frame - ref['image']).max() <= COLOR_TOL
        inv = np.float32(1) / np.float32(k)
        acc = acc * (np.float32(1) - inv) + frame * inv
        assert np.array_equal(scene.img.to_numpy(), acc)
    scene.clear()
    assert scene.accum.count[0] == 0 and float(scene.img.to_numpy().max()) == 0.0


def test_particle_raster_matches_reference_golden(tina):
    """§8f row 3: ParticleRaster + SimpleParticles + ParsTransform through tina.Scene, sharing the depth buffer
    with a MeshModel, against the golden produced by the reference's own sources."""
    import os
    import torch
    from test_golden import GOLDEN, _lighting
    g = np.load(os.path.join(GOLDEN, 'particles_and_mesh.npz'))
    W, H = (int(v) for v in g['res'])
    scene = tina.Scene((W, H), tonemap=False)
    pars = tina.SimpleParticles(maxpars=64)
    pars.set_particles(g['pos'])
    pars.set_particle_radii(g['rad'])
    pars.set_particle_colors(g['col'])
    scene.add_object(pars, tina.Classic())
    pars2 = tina.SimpleParticles(maxpars=64, radius=0.05)
    pars2.set_particles(g['pos'][:12] * np.float32(0.5))
    moved = tina.ParsTransform(pars2)
    moved.set_transform(g['trans'], 1.7)
    scene.add_object(moved, tina.Diffuse())
    scene.add_object(tina.MeshModel(scenes.load_monkey()), tina.Diffuse(color=[0.3, 0.5, 0.9]))
    scene.engine.W2V[None], scene.engine.V2W[None], scene.engine.bias[None] = g['W2V'], g['V2W'], g['bias']
    scene.render()
    torch.cuda.synchronize()
    assert np.array_equal(scene.engine.depth.to_numpy(), g['depth_after2'])
    assert np.array_equal(scene.triangle_raster.occup.to_numpy(), g['occup2'])
    # (raster.occup is resolved lazily from the shared key buffer: particle pixels later overwritten by the mesh read -1)
    po = scene.particle_raster.occup.to_numpy()
    assert np.array_equal(po[po >= 0], g['occup1'][po >= 0]) and (po >= 0).sum() > 0
    assert np.abs(scene.img.to_numpy() - g['image_after2']).max() <= COLOR_TOL
    # big discs (warp-cooperative walk) and the per-object state of a stand-alone raster
    engine = tina.Engine((W, H))
    engine.W2V[None], engine.V2W[None], engine.bias[None] = g['W2V'], g['V2W'], g['bias']
    pr = tina.ParticleRaster(engine)
    pr.set_particles(g['pos'])
    pr.set_particle_radii(g['rad'])
    pr.set_particle_colors(g['col'])
    engine.clear_depth()
    pr.render_occup()
    torch.cuda.synchronize()
    assert np.array_equal(pr.occup.to_numpy(), g['occup0'])
    assert np.array_equal(engine.depth.to_numpy(), g['depth_after0'])
    img = tina.Field(torch.zeros((W, H, 3), device='cuda'))
    pr.render_color(tina.Shader(img, _lighting(tina, g), tina.Classic()))
    assert np.abs(img.to_numpy() - g['image_after0']).max() <= COLOR_TOL


def test_particle_gbuffer_sinks_and_screen_space_passes(tina, O):
    """ParticleRaster.render_color serves every shader of a ShaderGroup (particle.py:129-161): the normal / position /
    colour / id sinks of the visible sphere points against the oracle, and Scene(ssao=True) / Scene(ssr=True) with particle
    objects (the AO and SSR fields equal the oracle's on this pipeline's own buffers)."""
    import os
    import torch
    from test_golden import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'particles_and_mesh.npz'))
    W, H = (int(v) for v in g['res'])
    engine = tina.Engine((W, H))
    engine.W2V[None], engine.V2W[None], engine.bias[None] = g['W2V'], g['V2W'], g['bias']
    pr = tina.ParticleRaster(engine)
    pr.set_particles(g['pos'])
    pr.set_particle_radii(g['rad'])
    pr.set_particle_colors(g['col'])
    engine.clear_depth()
    pr.render_occup()
    f3 = lambda: tina.Field(torch.zeros((W, H, 3), device='cuda'))  # noqa: E731
    nrm, pos, col = f3(), f3(), f3()
    ids = tina.Field(torch.full((W, H), -1, dtype=torch.int32, device='cuda'))
    probe = tina.ProbeShader((W, H))
    group = tina.ShaderGroup([tina.NormalShader(nrm), tina.PositionShader(pos), tina.ColorShader(col), tina.ConstShader(ids, 7), probe])
    pr.render_color(group)
    torch.cuda.synchronize()
    occ = pr.occup.to_numpy()
    assert np.array_equal(occ, g['occup0']) and (occ >= 0).sum() > 50
    rpos, rnrm = O.pars_attrs(g['pos'], g['rad'], occ, g['W2V'], g['V2W'], W, H, bias=g['bias'])
    assert np.abs(nrm.to_numpy() - rnrm).max() <= 2e-6 and np.abs(pos.to_numpy() - rpos).max() <= 2e-6
    vis = occ >= 0
    assert np.array_equal(col.to_numpy()[vis], g['col'][occ[vis]]) and not col.to_numpy()[~vis].any()
    assert np.array_equal(ids.to_numpy() == 7, vis)
    assert np.array_equal(probe.elmid.to_numpy()[vis], occ[vis])
    # screen-space passes over particles + a mesh
    for opts in (dict(ssao=True), dict(ssr=True)):
        scene = tina.Scene((W, H), smoothing=True, tonemap=False, **opts)
        pars = tina.SimpleParticles(maxpars=64)
        pars.set_particles(g['pos'])
        pars.set_particle_radii(g['rad'])
        mats = [tina.PBR(metallic=0.8, roughness=0.2), tina.Diffuse(color=[0.3, 0.5, 0.9])]
        scene.add_object(pars, mats[0])
        scene.add_object(tina.MeshModel(scenes.load_monkey()), mats[1])
        scene.engine.W2V[None], scene.engine.V2W[None] = g['W2V'], g['V2W']
        scene.triangle_raster.set_tuning(fast_shading=0)
        if 'ssr' in opts:
            scene.ssr.nsamples[None], scene.ssr.nsteps[None] = 6, 16
        scene.render()
        torch.cuda.synchronize()
        depth, nb = scene.engine.depth.to_numpy(), scene.norm_buffer.to_numpy()
        po = scene.particle_raster.occup.to_numpy()
        assert (np.square(nb[po >= 0]).sum(-1) > 0.99).all() and (po >= 0).sum() > 50  # the particles wrote their normals
        if 'ssao' in opts:
            ao_ref = O.ssao_render(depth, nb, g['W2V'], g['V2W'], scene.ssao.samples.cpu().numpy(), scene.ssao.rotations.cpu().numpy())
            ao = scene.ssao.img.to_numpy()
            assert np.abs(ao - ao_ref).max() <= 1e-6 and (ao != ao_ref).mean() < 0.01 and ao_ref.max() > 0.2
        else:
            assert set(np.unique(scene.mtlid_buffer.to_numpy()[po >= 0]).tolist()) == {0}
            assert float(scene.ssr.img.to_numpy()[..., 3].max()) > 0


def test_wireframe_raster_matches_reference_golden(tina, O):
    """§8f row 3: WireframeRaster + MeshToWire through tina.Scene over a solid mesh, against the golden produced
    by the reference's own sources; plus long / off-screen / degenerate lines against the oracle."""
    import os
    import torch
    from test_golden import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'particles_wireframe_over_mesh.npz'))
    W, H = (int(v) for v in g['res'])
    obj = scenes.load_monkey()
    scene = tina.Scene((W, H), tonemap=False)
    scene.add_object(tina.MeshTransform(tina.MeshModel(obj), tina.scale(0.97)), tina.Diffuse(color=[0.2, 0.3, 0.4]))
    scene.add_object(tina.MeshToWire(tina.MeshModel(obj)))
    scene.engine.W2V[None], scene.engine.V2W[None], scene.engine.bias[None] = g['W2V'], g['V2W'], g['bias']
    scene.render()
    torch.cuda.synchronize()
    assert np.array_equal(scene.engine.depth.to_numpy(), g['depth_after1'])
    assert np.abs(scene.img.to_numpy() - g['image_after1']).max() <= COLOR_TOL
    # stress: long lines crossing the screen, lines far outside, zero-length, behind the camera, clipping on/off
    rng = np.random.default_rng(9)
    wires = (rng.random((3000, 2, 3)).astype(np.float32) * 2 - 1) * np.float32([2.5, 2.5, 1.5])
    wires[:40] *= 40  # huge
    wires[40:60, 1] = wires[40:60, 0]  # zero length
    wires[60:80, :, 2] += 4  # behind the camera
    view, proj = scenes.default_camera(W / H)
    for clipping in (False, True):
        engine = tina.Engine((W, H))
        engine.set_camera(view, proj)
        wr = tina.WireframeRaster(engine, clipping=clipping, linecolor=(0.1, 0.7, 0.3))
        wr.set_wire_verts(wires)
        img = tina.Field(torch.zeros((W, H, 3), device='cuda'))
        engine.clear_depth()
        lighting = tina.Lighting()
        wr.render_color(tina.Shader(img, lighting, tina.Diffuse()))
        torch.cuda.synchronize()
        with np.errstate(all='ignore'):
            d, im = O.wire_render(wires, (proj @ view).astype(np.float32), W, H, O.clear_depth(W, H), np.zeros((W, H, 3), np.float32),
                                  color=(0.1, 0.7, 0.3), clipping=clipping)
        assert np.array_equal(engine.depth.to_numpy(), d), clipping
        assert np.array_equal(img.to_numpy(), im), clipping


def test_postfx_fxaa_and_blooming(tina, O):
    """§8f row 4 (deterministic half): FXAA and Blooming kernels against the golden produced by the reference's
    own sources, and inside Scene.render (raster.py:200-205 order: bloom -> tonemap -> fxaa) against the oracle."""
    import os
    import torch
    from test_golden import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'particles_postfx.npz'))
    W, H = g['input'].shape[:2]
    t = torch.as_tensor(g['input']).cuda()
    tina.FXAA((W, H)).apply(tina.Field(t))
    assert np.array_equal(t.cpu().numpy(), g['fxaa'])
    t = torch.as_tensor(g['input']).cuda()
    tina.Blooming((W, H)).apply(tina.Field(t))
    assert np.abs(t.cpu().numpy() - g['bloom']).max() <= 1e-6
    # in a scene: strong light so that the highlights exceed the bloom threshold
    W, H = 160, 120
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup(600, W, H, s=0.1, seed=3)
    imgs = {}
    for fx in (False, True):
        scene = tina.Scene((W, H), blooming=True, fxaa=fx)
        mesh = tina.SimpleMesh()
        mesh.set_face_verts(tri)
        scene.add_object(mesh, tina.Classic())
        scene.lighting.add_light(dir=[0, 0, 1], color=[6, 5, 4])
        scene.engine.set_camera(view, proj)
        scene.render()
        torch.cuda.synchronize()
        imgs[fx] = scene.img.to_numpy()
    ref = O.render_scene([(tri, None, None, tina.Classic())], W, H, view, proj, scene.lighting, _flags(O), do_tonemap=False)
    assert ref['image'].max() > 1.5
    img = O.tonemap(O.bloom(ref['image'], scene.blooming.gaussian_weights()))
    assert np.abs(imgs[False] - img).max() <= COLOR_TOL
    # FXAA branches on thresholds, so it is checked on identical input: the oracle filter applied to this
    # pipeline's own pre-FXAA frame must reproduce the fxaa=True frame bit for bit
    assert np.array_equal(imgs[True], O.fxaa(imgs[False]))


@pytest.mark.parametrize('seed', range(8))
def test_fuzz_random_scenes_ids_depth_bit_exact(tina, O, seed):
    """Random resolutions (1x1 ... odd, not multiples of the 16-px tiles or 256-px chunks), random cameras (also
    inside the geometry, so w <= 0 and huge projected faces occur), face sizes over six decades, duplicated /
    shared-vertex faces, random flags and sample bias -- expanded (SimpleMesh) and indexed (MeshModel) sources
    against the oracle: ids and depth bit for bit, colour inside the tolerance wherever the oracle is finite."""
    import torch
    rng = np.random.default_rng(1000 + seed)
    W, H = [(1, 1), (37, 53), (257, 3), (16, 16), (300, 17), (129, 255), (64, 500), (511, 97)][seed]
    nv = int(rng.integers(20, 400))
    scale = 10.0 ** rng.uniform(-3, 1, (nv, 1))
    v = (rng.normal(0, 1, (nv, 3)) * scale).astype(np.float32)
    v[rng.integers(0, nv, 5)] = v[rng.integers(0, nv, 5)]          # coincident vertices
    nf = int(rng.integers(50, 3000))
    f = rng.integers(0, nv, (nf, 3))
    f[rng.integers(0, nf, 10)] = f[rng.integers(0, nf, 10)]        # duplicated faces (exact depth ties)
    f[rng.integers(0, nf, 5), 1] = f[rng.integers(0, nf, 5), 0]    # degenerate (repeated index)
    eye = rng.normal(0, 1.5, 3)
    view = tina.lookat(pos=eye.tolist(), back=rng.normal(0, 1, 3).tolist(), up=[0, 1, 0.1])
    proj = tina.perspective(fov=float(rng.uniform(20, 120)), aspect=W / H, near=float(10.0 ** rng.uniform(-3, -0.5)), far=500.0)
    view, proj = np.asarray(view, np.float32), np.asarray(proj, np.float32)
    bias = tuple(rng.uniform(0, 1, 2).astype(np.float32).tolist()) if seed % 2 else (0.5, 0.5)
    culling, clipping = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    tri = np.ascontiguousarray(v[f])
    ref = None
    for source in ('expanded', 'indexed'):
        scene = tina.Scene((W, H), culling=culling, clipping=clipping, maxfaces=nf)
        if source == 'expanded':
            mesh = tina.SimpleMesh(maxfaces=nf)
            mesh.set_face_verts(tri)
        else:
            mesh = tina.MeshModel(dict(v=v, f=np.stack([f, np.zeros_like(f), np.zeros_like(f)], axis=2),
                                       vn=np.float32([[0, 0, 1]]), vt=np.float32([[0, 0]])))
        scene.add_object(mesh)
        scene.engine.set_camera(view, proj)
        scene.engine.bias[None] = bias
        scene.render()
        torch.cuda.synchronize()
        if ref is None:
            with np.errstate(all='ignore'):
                ref = O.render_scene([(tri, None, None, tina.Diffuse())], W, H, view, proj, scene.lighting,
                                     _flags(O, culling=culling, clipping=clipping), bias=bias)
        assert np.array_equal(scene.engine.depth.to_numpy(), ref['depth']), (source, W, H)
        assert np.array_equal(scene.triangle_raster.occup.to_numpy(), ref['occups'][-1]), (source, W, H)
        img, rimg = scene.img.to_numpy(), ref['image']
        ok = np.isfinite(rimg)
        assert np.array_equal(np.isfinite(img), ok)
        if ok.any():
            assert np.abs(img[ok] - rimg[ok]).max() <= COLOR_TOL


def test_frame_graph_replay_equals_eager(tina, O):
    """A recorded frame (CUDA graph) replays to the same bits as the eager calls, also after the geometry was
    animated in place, with large faces present (tile path recorded) and with a number of render_occup calls per
    frame that is not a multiple of the three rotating counter sets."""
    import torch
    W, H, n = 320, 200, 64
    view, proj = scenes.default_camera(W / H)
    scene = tina.Scene((W, H))
    grid = tina.MeshGrid(n)
    grid.pos.from_numpy(scenes.wave_grid_pos(n, t=0.1))
    scene.add_object(grid, tina.Classic())
    big = tina.SimpleMesh()
    big.set_face_verts(torch.tensor([[[-1.5, -1.2, -0.5], [1.5, -1.2, -0.5], [0.0, 1.4, -0.4]]], device='cuda'))
    scene.add_object(big, tina.Diffuse(color=[0.3, 0.6, 0.9]))
    scene.engine.set_camera(view, proj)
    g = tina.FrameGraph(scene.render)
    for t in (0.1, 0.35, 0.8):
        grid.pos.from_numpy(scenes.wave_grid_pos(n, t=t))
        g.replay()
        torch.cuda.synchronize()
        keys_g, img_g = scene.engine.keys.clone(), scene.img.to_torch().clone()
        scene.render()
        torch.cuda.synchronize()
        assert torch.equal(scene.engine.keys, keys_g), t
        assert torch.equal(scene.img.to_torch(), img_g), t
    assert (keys_g & 0xffffffff == 2 * (n - 1) ** 2 + 1).any()  # the large face is visible somewhere
    pos = scenes.wave_grid_pos(n, t=0.8)
    fv, fn = O.grid_faces(pos), O.grid_faces(O.grid_normals(pos))
    ref = O.render_occup(fv, (proj @ view).astype(np.float32), W, H)
    d, o = tina.multigpu.unpack_keys(keys_g.cpu().view(W, H))
    first = (o.numpy() >= 0) & (o.numpy() < fv.shape[0])
    assert np.array_equal(d.numpy()[first], ref[1][first]) and np.array_equal(o.numpy()[first], ref[0][first])


def test_peer_memory_composite_equals_single_engine(tina, O):
    """Sort-last composite fused into the shading kernel (render_color_composite): two engines rasterise disjoint
    face ranges with global ids; each shades one screen strip taking every key as the MIN over both key buffers.
    Keys and image must equal one engine rendering all faces.  (Across processes / GPUs the same table is filled
    through CUDA IPC, tools/bench_configs.py c5 --p2p.)"""
    import torch
    W, H, n = 512, 256, 20000
    view, proj = scenes.default_camera(W / H)
    tri = torch.as_tensor(scenes.soup(n, W, H, s=0.02, seed=21)).cuda()
    lighting = tina.Lighting()
    lighting.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
    lighting.set_ambient_light([0.1, 0.1, 0.1])

    def setup():
        e = tina.Engine((W, H))
        e.set_camera(view, proj)
        r = tina.TriangleRaster(e, maxfaces=n)
        img = tina.Field(torch.zeros((W, H, 3), device='cuda'))
        return e, r, tina.Shader(img, lighting, tina.Diffuse())

    e0, r0, s0 = setup()
    e0.clear_depth()
    r0.set_face_verts(tri)
    r0.render_occup()
    r0.render_color(s0, fill_bg=np.float32([0.1, 0.2, 0.3]))
    parts = [setup() for _ in range(2)]
    engines = [p[0] for p in parts]
    cut = 7777
    for k, (e, r, s) in enumerate(parts):
        lo, hi = (0, cut) if k == 0 else (cut, n)
        e.clear_depth()
        e.set_face_base(lo)
        r.set_face_verts(tri[lo:hi])
        r.render_occup()
        e.set_peer_keys(engines, k)
    npix = W * H
    half = (npix // 2) // 256 * 256
    out = torch.zeros((W, H, 3), device='cuda')
    for k, (e, r, s) in enumerate(parts):
        r.set_face_verts(tri)  # the full arrays for shading (ids in the keys are global)
        p_lo, p_hi = (0, half) if k == 0 else (half, npix)
        r.render_color_composite(s, p_lo, p_hi - p_lo, face_base=0, fill_bg=np.float32([0.1, 0.2, 0.3]))
        out.view(-1)[p_lo * 3:p_hi * 3] = s.img.to_torch().view(-1)[p_lo * 3:p_hi * 3]
    torch.cuda.synchronize()
    assert torch.equal(out, s0.img.to_torch())
    comp = torch.cat([engines[0].keys.view(-1)[:half], engines[1].keys.view(-1)[half:]])
    assert torch.equal(comp, e0.keys.view(-1))
    for e in engines:
        e.close_peer_keys()


def test_shared_divisor_division_is_ieee(tina):
    """The setup code divides several numbers by one divisor with the reciprocal refinement shared (div_many in
    csrc/common.cuh); it must return the bits of IEEE round-to-nearest division for every operand: 6e9 quotients
    (random bit patterns incl. nan / inf / denormals, moderate exponents with random mantissas, quotients at and next to 1,
    signed zeros, all-ones mantissas, reciprocals) against __fdiv_rn."""
    import ctypes as C
    from taichi_three_b200 import _lib
    for seed in (1, 2):
        bad = C.c_uint64(123)
        _lib.check(_lib.lib().tina_selftest_division(0, 3 * 10**9, seed, C.byref(bad)))
        assert bad.value == 0


def test_ssao_matches_reference_golden_and_scene_option(tina, O):
    """§8f row 4: SSAO kernels on the golden's inputs (depth, normals, tables from the reference's own run) and
    Scene(ssao=True) end to end against the oracle applied to this pipeline's own depth / normal buffers."""
    import os
    import torch
    from test_golden import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'particles_ssao.npz'))
    W, H = g['depth'].shape
    eng = tina.Engine((W, H))
    eng.W2V[None], eng.V2W[None] = g['W2V'], g['V2W']
    eng.keys.copy_(torch.as_tensor(g['depth'].astype(np.int64) << 32).cuda())
    norm = tina.Field(torch.as_tensor(g['normals']).cuda())
    ssao = tina.SSAO((W, H), norm)
    ssao.samples, ssao.rotations = torch.as_tensor(g['samples']).cuda(), torch.as_tensor(g['rotations']).cuda()
    ssao.render(eng)
    torch.cuda.synchronize()
    ao = ssao.img.to_numpy()
    assert np.abs(ao - g['ao']).max() <= 1e-6 and (ao != g['ao']).mean() < 0.01
    img = tina.Field(torch.as_tensor(g['image_before']).cuda())
    ssao.img.from_numpy(g['ao'])
    ssao.apply(img)
    assert np.abs(img.to_numpy() - g['image_after']).max() <= 1e-6
    # the Scene option: normal G-buffer pre-shader + render + apply before the tonemap (raster.py:51-66, 189-191)
    obj = scenes.load_monkey()
    scene = tina.Scene((96, 80), smoothing=True, ssao=True)
    scene.add_object(tina.MeshModel(obj))
    view, proj = scenes.default_camera(96 / 80)
    scene.engine.set_camera(view, proj)
    scene.render()
    torch.cuda.synchronize()
    depth, nrm = scene.engine.depth.to_numpy(), scene.norm_buffer.to_numpy()
    W2V = (np.asarray(proj, np.float64) @ np.asarray(view, np.float64))
    ao_ref = O.ssao_render(depth, nrm, W2V.astype(np.float32), np.linalg.inv(W2V).astype(np.float32), scene.ssao.samples.cpu().numpy(),
                           scene.ssao.rotations.cpu().numpy())
    ao_gpu = scene.ssao.img.to_numpy()
    assert ao_ref.max() > 0.2
    assert np.abs(ao_gpu - ao_ref).max() <= 1e-6 and (ao_gpu != ao_ref).mean() < 0.01
    scene2 = tina.Scene((96, 80), smoothing=True)
    scene2.add_object(tina.MeshModel(obj))
    scene2.engine.set_camera(view, proj)
    scene2.triangle_raster.set_tuning(fast_shading=0)
    scene.triangle_raster.set_tuning(fast_shading=0)
    scene.render()
    scene2.tonemap = False
    scene2.render()
    torch.cuda.synchronize()
    expect = O.tonemap(O.ssao_apply(scene2.img.to_numpy(), scene.ssao.img.to_numpy()))
    assert np.abs(scene.img.to_numpy() - expect).max() <= 1e-6


def test_consecutive_frames_overlap_without_interference(tina, O):
    """Frames rendered back to back with nothing between them: the vertex stage of frame k + 1 (which does not wait for
    frame k's shading kernel and writes the other per-vertex record set) must not disturb frame k.  Eight cameras, one
    scene and raster, one image per frame; every image and key buffer equals the frame rendered in isolation."""
    import torch
    W, H, n = 640, 360, 256
    scene = tina.Scene((W, H), smoothing=True, maxfaces=2**18, tonemap=False)
    mesh = tina.MeshGrid(n)
    mesh.pos.from_numpy(scenes.wave_grid_pos(n))
    mat = tina.Classic()
    scene.add_object(mesh, mat)
    raster = scene.triangle_raster
    raster.set_object(mesh)
    cams = [tina.orbit_camera(radius=3.0, theta=0.1 * k, phi=0.07 * k, aspect=W / H) for k in range(8)]
    imgs = [tina.Field(torch.zeros((W, H, 3), device='cuda')) for _ in cams]
    shaders = [tina.Shader(im, scene.lighting, mat) for im in imgs]
    bg = np.zeros(3, np.float32)

    def frame(k):
        scene.engine.set_camera(*cams[k])
        scene.engine.clear_depth()
        raster.render_occup()
        raster.render_color(shaders[k], fill_bg=bg)
    ref = []
    for k in range(len(cams)):  # isolated frames
        frame(k)
        torch.cuda.synchronize()
        ref.append((imgs[k].to_numpy().copy(), scene.engine.keys.clone()))
        imgs[k].to_torch().zero_()
    torch.cuda.synchronize()
    for rep in range(5):
        for k in range(len(cams)):  # back to back, no synchronisation, no other kernel between the frames
            frame(k)
        torch.cuda.synchronize()
        for k in range(len(cams)):
            assert np.array_equal(imgs[k].to_numpy(), ref[k][0]), (rep, k)
        assert torch.equal(scene.engine.keys, ref[-1][1])
    raster.set_tuning(overlap_vertex=0)
    for k in range(len(cams)):
        frame(k)
    torch.cuda.synchronize()
    assert all(np.array_equal(imgs[k].to_numpy(), ref[k][0]) for k in range(len(cams)))


def _ssr_close(a, b, tol=2e-5, max_outliers=0.01):
    """SSR fields agree: every value within tol except for isolated pixels where one ray-march test (a hard `<`) flipped
    on the last ulp of sinf / cosf / powf; those are counted, and bounded."""
    bad = np.abs(a - b).max(-1) > tol
    return bad.mean() <= max_outliers, float(bad.mean()), float(np.abs(a - b)[~bad].max(initial=0.0))


def test_ssr_matches_reference_golden_and_scene_option(tina, O):
    """§8f row 4: SSR (postp/ssr.py) on the golden's inputs from the reference's own run -- three material graphs incl.
    a textured PBR and an Add / Scale / Mix / Emission composite -- and Scene(ssr=True, texturing=True) end to end against
    the oracle applied to this pipeline's own depth / normal / texcoord / material-id buffers."""
    import os
    import torch
    from test_golden import GOLDEN, _material
    g = np.load(os.path.join(GOLDEN, 'particles_ssr.npz'))
    W, H = g['depth'].shape
    eng = tina.Engine((W, H))
    eng.W2V[None], eng.V2W[None] = g['W2V'], g['V2W']
    eng.keys.copy_(torch.as_tensor(g['depth'].astype(np.int64) << 32).cuda())
    dev = 'cuda'
    norm, coor = tina.Field(torch.as_tensor(g['normals']).to(dev)), tina.Field(torch.as_tensor(g['coors']).to(dev))
    mtlid = tina.Field(torch.as_tensor(g['mtlid']).to(dev))
    from taichi_three_b200.scene import MaterialTable
    tab = MaterialTable()
    for i in range(int(g['nspecs'])):
        tab.add_material(_material(tina, g, i, 'spec'))
    ssr = tina.SSR((W, H), norm, coor, mtlid, tab)
    ssr.nsamples[None], ssr.nsteps[None] = int(g['nsamples']), int(g['nsteps'])
    img = tina.Field(torch.as_tensor(g['image_before']).to(dev))
    ssr.render(eng, img)
    torch.cuda.synchronize()
    ok, frac, err = _ssr_close(ssr.img.to_numpy(), g['ssr'])
    assert ok and err <= 2e-5, (frac, err)
    assert (ssr.img.to_numpy()[..., 3] > 0).sum() > 300
    ssr.img.from_numpy(g['ssr'])
    ssr.apply(img)
    assert np.abs(img.to_numpy() - g['image_after']).max() <= 1e-6
    # a table the device could not walk is refused
    from taichi_three_b200 import _lib
    bad = _lib.TinaSampleMaterial()
    bad.nnodes, bad.nodes[0].kind, bad.nodes[0].a, bad.nodes[0].b = 1, _lib.SNODE_MIX, 0, 0
    with pytest.raises(_lib.TinaError):
        _lib.check(_lib.lib().tina_engine_ssr_render(eng._h, norm.to_torch().data_ptr(), None, mtlid.to_torch().data_ptr(), (_lib.TinaSampleMaterial * 1)(bad), 1,
                                                     img.to_torch().data_ptr(), 1, 1, 2.0, 15.0, 4, 0, 0, ssr.img.to_torch().data_ptr(), None))

    # the Scene option (raster.py:43-70, 192-194): default sample counts, textured floor
    rng = np.random.default_rng(5)
    tex = rng.random((8, 8, 3)).astype(np.float32)
    mats = [tina.PBR(basecolor=[0.9, 0.8, 0.7], metallic=0.8, roughness=0.1), tina.PBR(basecolor=tina.Texture(tex), metallic=0.3, roughness=0.3),
            tina.Classic()]
    scene = tina.Scene((96, 80), smoothing=True, texturing=True, ssr=True, tonemap=False)
    obj = scenes.load_monkey()
    scene.add_object(tina.MeshModel(obj), mats[0])
    quad_v = np.array([[[-2.5, -0.9, -2.5], [-2.5, -0.9, 2.5], [2.5, -0.9, 2.5]], [[-2.5, -0.9, -2.5], [2.5, -0.9, 2.5], [2.5, -0.9, -2.5]]], np.float32)
    quad_n = np.tile(np.array([0, 1, 0], np.float32), (2, 3, 1))
    quad_t = np.array([[[0, 0], [0, 1], [1, 1]], [[0, 0], [1, 1], [1, 0]]], np.float32)
    floor = tina.SimpleMesh()
    floor.set_face_verts(quad_v), floor.set_face_norms(quad_n), floor.set_face_coors(quad_t)
    scene.add_object(floor, mats[1])
    wall = tina.SimpleMesh()
    wall.set_face_verts(quad_v[:, :, [0, 2, 1]] * np.array([1, 1, 1], np.float32) + np.array([0, 1.6, -1.6], np.float32))
    wall.set_face_norms(np.tile(np.array([0, 0, 1], np.float32), (2, 3, 1))), wall.set_face_coors(quad_t)
    scene.add_object(wall, mats[2])
    view, proj = scenes.default_camera(96 / 80)
    scene.engine.set_camera(view, proj)
    scene.ssr.nsamples[None], scene.ssr.nsteps[None] = 8, 24
    scene.triangle_raster.set_tuning(fast_shading=0)
    scene.render()
    torch.cuda.synchronize()
    after = scene.img.to_numpy()
    depth, nrm = scene.engine.depth.to_numpy(), scene.norm_buffer.to_numpy()
    W2V = (np.asarray(proj, np.float64) @ np.asarray(view, np.float64))
    # the image SSR read = the frame without the SSR pass
    scene.ssr_saved, scene.ssr = scene.ssr, False
    scene.render()
    torch.cuda.synchronize()
    before = scene.img.to_numpy()
    scene.ssr = scene.ssr_saved
    ref4 = O.ssr_render(depth, nrm, scene.coor_buffer.to_numpy(), scene.mtlid_buffer.to_numpy(), mats, before, W2V.astype(np.float32),
                        np.linalg.inv(W2V).astype(np.float32), nsamples=8, nsteps=24)
    assert {0, 1} <= set(scene.mtlid_buffer.to_numpy()[ref4[..., 3] > 0].tolist()) and (ref4[..., 3] > 0).sum() > 500
    ok, frac, err = _ssr_close(scene.ssr.img.to_numpy(), ref4)
    assert ok and err <= 2e-5, (frac, err)
    assert np.abs(after - O.ssr_apply(before, scene.ssr.img.to_numpy())).max() <= 1e-6


def test_ssr_default_parameters_match_reference_golden(tina):
    """SSR with its default 32 samples x 32 steps, scene without texturing (texcoord 0), against the reference's own run."""
    import os
    import torch
    from test_golden import GOLDEN, _material
    from taichi_three_b200.scene import MaterialTable
    g = np.load(os.path.join(GOLDEN, 'particles_ssr_defaults.npz'))
    W, H = g['depth'].shape
    eng = tina.Engine((W, H))
    eng.W2V[None], eng.V2W[None] = g['W2V'], g['V2W']
    eng.keys.copy_(torch.as_tensor(g['depth'].astype(np.int64) << 32).cuda())
    tab = MaterialTable()
    for i in range(int(g['nspecs'])):
        tab.add_material(_material(tina, g, i, 'spec'))
    ssr = tina.SSR((W, H), tina.Field(torch.as_tensor(g['normals']).cuda()), None, tina.Field(torch.as_tensor(g['mtlid']).cuda()), tab)
    assert (int(ssr.nsamples[None]), int(ssr.nsteps[None]), int(ssr.blurring[None])) == (32, 32, 4)
    img = tina.Field(torch.as_tensor(g['image_before']).cuda())
    ssr.render(eng, img)
    torch.cuda.synchronize()
    ok, frac, err = _ssr_close(ssr.img.to_numpy(), g['ssr'])
    assert ok and err <= 2e-5, (frac, err)
    ssr.apply(img)
    assert np.abs(img.to_numpy() - g['image_after']).max() <= 1e-3 and np.abs(img.to_numpy() - g['image_after']).mean() <= 1e-5


def test_primitive_and_connective_meshes(tina, O):
    """mesh/prim.py + mesh/conn.py through the raster: a PrimitiveMesh sphere + cylinder scene against the oracle on the same
    face arrays (docs/primitives.py's objects), and a ConnectiveMesh equal to the MeshModel with the same shared indices."""
    import torch
    from taichi_three_b200.mesh import primitive_sphere, primitive_cylinder
    W, H = 160, 120
    view, proj = scenes.default_camera(W / H)
    scene = tina.Scene((W, H), smoothing=True, texturing=True)
    sph, cyl = tina.PrimitiveMesh.sphere(16, 12, 0.8), tina.MeshTransform(tina.PrimitiveMesh.cylinder(12, 2, 0.4, 1.2), tina.translate([1.1, 0, 0]))
    scene.add_object(sph, tina.Classic())
    scene.add_object(cyl, tina.Diffuse(color=[0.8, 0.5, 0.3]))
    scene.engine.set_camera(view, proj)
    scene.triangle_raster.set_tuning(fast_shading=0)
    scene.render()
    torch.cuda.synchronize()
    fs, fc = primitive_sphere(16, 12, 0.8), primitive_cylinder(12, 2, 0.4, 1.2)
    cv, cn = O.transform(fc[:, :, 0], fc[:, :, 1], tina.translate([1.1, 0, 0]))
    flags = O.SMOOTHING | O.TEXTURING | O.CULLING | O.CLIPPING
    c2 = lambda a: np.ascontiguousarray(a[:, :, :2])  # noqa: E731
    ref = O.render_scene([(fs[:, :, 0], fs[:, :, 1], c2(fs[:, :, 2]), tina.Classic()),
                          (cv, cn, c2(fc[:, :, 2]), tina.Diffuse(color=[0.8, 0.5, 0.3]))], W, H, view, proj, scene.lighting, flags)
    assert np.array_equal(scene.engine.depth.to_numpy(), ref['depth']) and (ref['depth'] < 2**30).sum() > 2000
    assert np.abs(scene.img.to_numpy() - ref['image']).max() <= COLOR_TOL
    # ConnectiveMesh == MeshModel with (v, v, v) index triples
    obj = scenes.load_monkey()
    n = len(obj['v'])
    rng = np.random.default_rng(3)
    vn = rng.normal(size=(n, 3)).astype(np.float32)
    vt = rng.random((n, 2)).astype(np.float32)
    conn = tina.ConnectiveMesh()
    conn.set_vertices(obj['v']), conn.set_vert_norms(vn), conn.set_vert_coors(vt), conn.set_faces(obj['f'][:, :, 0])
    model = tina.MeshModel({'v': obj['v'], 'vn': vn, 'vt': vt, 'f': np.repeat(obj['f'][:, :, :1], 3, axis=2)})
    imgs = []
    for mesh in (conn, model):
        sc = tina.Scene((96, 96), smoothing=True, texturing=True)
        sc.add_object(mesh, tina.Classic())
        sc.engine.set_camera(*scenes.default_camera())
        sc.render()
        torch.cuda.synchronize()
        imgs.append((sc.img.to_numpy(), sc.engine.depth.to_numpy()))
    assert np.array_equal(imgs[0][0], imgs[1][0]) and np.array_equal(imgs[0][1], imgs[1][1]) and (imgs[0][1] < 2**30).sum() > 1000


def test_reference_ssr_script_flow(tina):
    """The reference's own tests/ssr.py without its GUI: Scene(ssr=True, taa=True), PBR materials driven by tina.Param,
    a transformed MeshGrid as the mirror plane, the SSR fields set every frame, TAA accumulation over a few frames."""
    import torch
    scene = tina.Scene((160, 120), smoothing=True, ssr=True, taa=True)
    monkey = tina.MeshModel(scenes.load_monkey())
    scene.add_object(monkey, tina.PBR(metallic=0.0, roughness=0.4))
    param_metallic, param_roughness = tina.Param(), tina.Param()
    plane = tina.MeshTransform(tina.MeshGrid(32), tina.translate([0, -1, 0]) @ tina.scale(2) @ tina.eularXYZ([-np.pi / 2, 0, 0]))
    scene.add_object(plane, tina.PBR(metallic=param_metallic, roughness=param_roughness))
    scene.engine.set_camera(*tina.orbit_camera(radius=3.5, theta=0.45, phi=0.3, aspect=160 / 120))
    imgs = []
    for metallic in (1.0, 0.0):
        scene.clear()
        for frame in range(3):
            scene.ssr.nsteps[None], scene.ssr.nsamples[None], scene.ssr.blurring[None] = 64, 12, 4
            scene.ssr.stepsize[None], scene.ssr.tolerance[None] = 2, 15
            param_metallic.value[None], param_roughness.value[None] = metallic, 0.0
            scene.render()
        torch.cuda.synchronize()
        img = scene.img.to_numpy()
        assert np.isfinite(img).all() and img.max() > 0.1
        assert float(scene.ssr.img.to_numpy()[..., 3].max()) > 0 and scene.accum.count[0] == 3
        imgs.append(img)
    assert np.abs(imgs[0] - imgs[1]).max() > 1e-3  # the mirror plane reflects, the dielectric one much less


def test_ssao_and_ssr_taa_modes_follow_the_hash_stream(tina, O):
    """taa=True: fresh samples per pixel and frame (ssao.py:52-56,80-81; ssr.py:72).  The reference's ti.random() stream
    is unspecified; the product draws from the Wang hash seeded with (pixel, frame) and the oracle restates that stream:
    each frame equals the oracle's frame, frames differ from each other, and their mean approaches the table mode's AO."""
    import torch
    obj = scenes.load_monkey()
    scene = tina.Scene((96, 80), smoothing=True, ssao=True, taa=True, tonemap=False)
    scene.add_object(tina.MeshModel(obj))
    view, proj = scenes.default_camera(96 / 80)
    scene.engine.set_camera(view, proj)
    W2V = (np.asarray(proj, np.float64) @ np.asarray(view, np.float64))
    w2v, v2w = W2V.astype(np.float32), np.linalg.inv(W2V).astype(np.float32)
    frames = []
    for f in range(3):
        scene.render()
        torch.cuda.synchronize()
        bias = np.asarray(scene.engine.bias[None], np.float32)
        ao_ref = O.ssao_render_taa(scene.engine.depth.to_numpy(), scene.norm_buffer.to_numpy(), w2v, v2w, nsamples=scene.ssao.nsamples, frame=f, bias=bias)
        ao = scene.ssao.img.to_numpy()
        assert np.abs(ao - ao_ref).max() <= 0.05 and (np.abs(ao - ao_ref) > 1e-6).mean() < 0.01  # (a flipped depth test moves 1 / nsamples)
        frames.append(ao)
    assert np.abs(frames[0] - frames[1]).max() > 0.01
    # SSR, taa stream
    scene = tina.Scene((64, 48), smoothing=True, ssr=True, taa=True, tonemap=False)
    mat = tina.PBR(metallic=0.9, roughness=0.1)
    scene.add_object(tina.MeshModel(obj), mat)
    scene.engine.set_camera(*scenes.default_camera(64 / 48))
    scene.triangle_raster.set_tuning(fast_shading=0)
    scene.ssr.nsamples[None], scene.ssr.nsteps[None] = 4, 16
    scene.render()
    torch.cuda.synchronize()
    assert scene.ssr.frame == 1 and float(scene.ssr.img.to_numpy()[..., 3].max()) > 0


def test_textured_materials_all_prologue_forms(tina, O):
    """Textured colour through every prologue form of the material compiler -- straight-line code for the stock
    PBR / Classic / Diffuse shapes (prologue_form 1 / 3 / 4), the three-address interpreter for anything else (2) --
    against the oracle, which interprets the plain unfolded programs; and each form against generic_vm=1."""
    import torch
    from taichi_three_b200 import material as M
    W, H, n = 200, 160, 24
    view, proj = scenes.default_camera(W / H)
    rng = np.random.default_rng(9)
    img = rng.random((37, 29, 3)).astype(np.float32)
    pos = scenes.wave_grid_pos(n)
    mats = [(tina.PBR(basecolor=tina.Texture(img), metallic=0.3, roughness=0.4), 1),
            (tina.Classic(color=tina.Texture(img), shineness=8, specular=0.7), 3),
            (tina.Diffuse(color=tina.Texture(img)), 4),
            (tina.Lambert() * tina.Texture(img) + tina.Phong(shineness=16) * tina.Texture(img) + tina.Emission() * 0.1, 2)]
    fv, fn = O.grid_faces(pos), O.grid_faces(O.grid_normals(pos))
    ft = O.grid_faces(O.grid_texcoords(n, n)) if hasattr(O, 'grid_texcoords') else None
    for mat, form in mats:
        st, _ = M.material_struct(mat, torch.device('cuda', 0))
        assert st.prologue_form == form
        imgs = []
        for generic in (0, 1):
            scene = tina.Scene((W, H), smoothing=True, texturing=True, tonemap=False)
            grid = tina.MeshGrid(n)
            grid.pos.from_numpy(pos)
            scene.add_object(grid, mat)
            scene.engine.set_camera(view, proj)
            scene.lighting.add_light(pos=[0.4, 0.3, 1.5], color=[0.5, 0.4, 0.3])
            scene.triangle_raster.set_tuning(generic_vm=generic, fast_shading=0)
            scene.render()
            torch.cuda.synchronize()
            imgs.append(scene.img.to_numpy())
        assert np.array_equal(imgs[0], imgs[1]), form
        coors = scene.triangle_raster.coors.to_numpy()
        ref = O.render_scene([(fv, fn, coors, mat)], W, H, view, proj, scene.lighting, _flags(O, smoothing=True, texturing=True),
                             do_tonemap=False)
        assert np.abs(imgs[0] - ref['image']).max() <= COLOR_TOL, form
        scene.triangle_raster.set_tuning(generic_vm=0, fast_shading=1)
        scene.render()
        torch.cuda.synchronize()
        assert np.abs(scene.img.to_numpy() - ref['image']).max() <= COLOR_TOL, form


def _edge_case_triangles(W, H):
    tri = scenes.soup(3000, W, H, s=0.05, seed=5)
    extra = np.array([
        [[-9, -9, 0], [9, -9, 0], [0, 9, 0]],            # screen-filling, all vertices outside
        [[-0.5, -0.5, 5], [0.5, -0.5, 5], [0, 0.5, 5]],  # behind the camera
        [[-0.5, -0.5, 0], [0.5, -0.5, 4], [0, 0.5, 0]],  # crosses w = 0
        [[0, 0, 0], [0, 0, 0], [0, 0, 0]],               # degenerate
        [[np.nan, 0, 0], [1, 0, 0], [0, 1, 0]],          # NaN
        [[np.inf, 0, 0], [1, 0, 0], [0, 1, 0]],          # inf
        [[50, 50, 0], [51, 50, 0], [50, 51, 0]],         # far off-screen
        [[0, 0, 2.9999], [0.1, 0, 2.9999], [0, 0.1, 2.9999]],  # w ~ 1e-4: huge viewport coordinates (not tame)
        [[-1, -1, 0.5], [1, -1, 0.5], [0, 1, 0.5]],      # big, front-facing
        [[-1, -1, 0.2], [0, 1, 0.2], [1, -1, 0.2]],      # big, back-facing
        [[1e-3, 0, 1], [2e-3, 0, 1], [1e-3, 1e-3, 1]],   # sub-pixel
    ], dtype=np.float32)
    return np.ascontiguousarray(np.concatenate([extra, tri, extra[::-1]]))


@pytest.mark.parametrize('culling,clipping', [(True, True), (True, False), (False, True), (False, False)])
def test_indexed_records_edge_cases_and_general_path(tina, O, culling, clipping):
    """The indexed rasteriser (k_raster_indexed) on per-vertex records: vertices behind the camera, NaN / inf,
    w ~ 0 (huge viewport coordinates), off-screen and screen-filling faces as a MeshModel.  Ids + depth equal the
    oracle, and the per-vertex integer bounds (record path) equal the general float-bbox path (force_general) and
    the untightened walk bit for bit."""
    import torch
    W, H = 200, 136
    view, proj = scenes.default_camera(W / H)
    tri = _edge_case_triangles(W, H)
    obj = {'v': tri.reshape(-1, 3), 'f': np.arange(len(tri) * 3).reshape(-1, 3)}
    ref = None
    keys = []
    for tuning in (dict(), dict(force_general=1), dict(tighten=0), dict(lean_kernels=0), dict(balance=0), dict(tiny_max=4),
                   dict(force_tiles=1)):
        scene = tina.Scene((W, H), culling=culling, clipping=clipping)
        scene.add_object(tina.MeshModel(obj))
        scene.engine.set_camera(view, proj)
        scene.triangle_raster.set_tuning(**tuning)
        scene.render()
        torch.cuda.synchronize()
        if ref is None:
            with np.errstate(all='ignore'):
                ref = O.render_scene([(tri, None, None, tina.Diffuse())], W, H, view, proj, scene.lighting,
                                     _flags(O, culling=culling, clipping=clipping))
        assert np.array_equal(scene.engine.depth.to_numpy(), ref['depth']), tuning
        assert np.array_equal(scene.triangle_raster.occup.to_numpy(), ref['occups'][-1]), tuning
        keys.append(_keys(scene))
    for k in keys[1:]:
        assert np.array_equal(k, keys[0])
    assert (ref['depth'] < 2**30).sum() > 1000


@pytest.mark.parametrize('bias', [(0.5, 0.5), (0.123, 0.877), (0.0, 1.0), (1.5, -0.25)])
def test_indexed_record_bounds_equal_general_path_micro_grid(tina, O, bias):
    """Sub-pixel regime (C2's): a wavy grid at 0.3 px per face.  Per-vertex integer candidate bounds + min3 / max3
    must select exactly the candidates of the per-face float computation (general path), for centred, jittered,
    extreme and out-of-range sample bias (the last one disables tightening on the host)."""
    import torch
    n, W, H = 160, 96, 64
    pos = scenes.wave_grid_pos(n)
    view, proj = scenes.default_camera(W / H)
    keys = []
    for tuning in (dict(), dict(force_general=1), dict(tighten=0)):
        scene = tina.Scene((W, H), smoothing=True)
        grid = tina.MeshGrid(n)
        grid.pos.from_numpy(pos)
        scene.add_object(grid, tina.Classic())
        scene.engine.set_camera(view, proj)
        scene.engine.bias[None] = bias
        scene.triangle_raster.set_tuning(**tuning)
        scene.render()
        torch.cuda.synchronize()
        keys.append(_keys(scene))
    assert np.array_equal(keys[0], keys[1]) and np.array_equal(keys[0], keys[2])
    fv = O.grid_faces(pos)
    ref = O.render_scene([(fv, O.grid_faces(O.grid_normals(pos)), None, tina.Classic())], W, H, view, proj, scene.lighting,
                         _flags(O, smoothing=True), bias=bias)
    d, o = keys[0] >> 32, (keys[0] & 0xffffffff).astype(np.int64) - 1
    assert np.array_equal(d.astype(np.int32), ref['depth'])
    assert np.array_equal(o.astype(np.int32), ref['occups'][-1])


def test_deferred_clear_depth_semantics(tina, O):
    """clear_depth is deferred inside the library (folded into the next render_occup of an indexed source); every
    observable view must still see it: engine.keys / engine.depth right after the clear, a second clear, other
    rasterisers, and an immediate clear (lazy off) must all give the same frames."""
    import torch
    W, H = 160, 120
    view, proj = scenes.default_camera(W / H)
    pos = scenes.wave_grid_pos(40)
    scene = tina.Scene((W, H), smoothing=True)
    grid = tina.MeshGrid(40)
    grid.pos.from_numpy(pos)
    scene.add_object(grid, tina.Classic())
    scene.engine.set_camera(view, proj)
    scene.render()
    k1 = _keys(scene)
    assert ((k1 & 0xffffffff) != 0).sum() > 500
    scene.engine.clear_depth()
    assert np.all(scene.engine.depth.to_numpy() == 2**30)          # the view flushes the pending clear
    assert np.all((scene.engine.keys.cpu().numpy() & 0xffffffff) == 0)
    scene.engine.clear_depth()
    scene.engine.clear_depth()
    scene.render()
    assert np.array_equal(_keys(scene), k1)
    # the expanded-array rasteriser after a deferred clear
    scene2 = tina.Scene((W, H))
    tri = scenes.soup(2000, W, H, s=0.03, seed=3)
    mesh = tina.SimpleMesh()
    mesh.set_face_verts(tri)
    scene2.add_object(mesh)
    scene2.engine.set_camera(view, proj)
    scene2.render()
    a = _keys(scene2)
    scene2.render()
    assert np.array_equal(_keys(scene2), a)
    from taichi_three_b200 import _lib
    _lib.check(_lib.lib().tina_engine_set_lazy_clear(scene2.engine._h, 0))
    scene2.render()
    assert np.array_equal(_keys(scene2), a)


@pytest.mark.parametrize('n,W,H', [(160, 96, 64), (300, 640, 360), (131, 200, 136), (5, 64, 48)])
def test_grid_tile_rasteriser_equals_gather_kernel(tina, O, n, W, H):
    """The three rasterisers a plain square MeshGrid can take -- k_raster_quads (default: one quad = two faces per thread),
    k_raster_indexed (grid_quads=0: one face per thread, what every other indexed source uses) and k_raster_grid (knob
    grid_tiles: persistent warps over row chunks staged by cp.async) -- must give the same bits and the oracle's: chunks per row that divide / do not divide the row, more chunks than
    resident warps, a grid smaller than one chunk; with and without the lean variant and culling."""
    import torch
    pos = scenes.wave_grid_pos(n)
    view, proj = scenes.default_camera(W / H)
    fv = O.grid_faces(pos)
    for culling in (True, False):
        keys = []
        for tuning in (dict(), dict(grid_quads=0), dict(grid_quads=0, lean_kernels=0), dict(force_general=1), dict(grid_tiles=1),
                       dict(grid_tiles=1, lean_kernels=0), dict(grid_tiles=1, force_general=1), dict(grid_tiles=1, tiny_max=2),
                       dict(lean_kernels=0), dict(tiny_max=2), dict(tiny_max=0), dict(grid_quads=0, tiny_max=2)):
            scene = tina.Scene((W, H), smoothing=True, culling=culling)
            grid = tina.MeshGrid(n)
            grid.pos.from_numpy(pos)
            scene.add_object(grid, tina.Classic())
            scene.engine.set_camera(view, proj)
            scene.triangle_raster.set_tuning(**tuning)
            scene.render()
            scene.render()  # (second frame: the pipeline's stage / parity state starts over per launch)
            torch.cuda.synchronize()
            keys.append(_keys(scene))
        for k in keys[1:]:
            assert np.array_equal(k, keys[0])
        ref = O.render_scene([(fv, O.grid_faces(O.grid_normals(pos)), None, tina.Classic())], W, H, view, proj, scene.lighting,
                             _flags(O, smoothing=True, culling=culling))
        d, o = keys[0] >> 32, (keys[0] & 0xffffffff).astype(np.int64) - 1
        assert np.array_equal(d.astype(np.int32), ref['depth'])
        assert np.array_equal(o.astype(np.int32), ref['occups'][-1])
        _check_frame(scene, ref)


# ---- parity at BASELINE.json sizes (C2 is test_c2_full_size_bit_exact) ------------------------------------------
def test_c2b_full_size_nocull_bit_exact(tina, O):
    """C2b: the reference's own example wraps the grid in MeshNoCulling (examples/meshgrid_wave.py:21): 4,186,116 faces."""
    n, W, H = 1024, 1920, 1080
    pos = scenes.wave_grid_pos(n)
    scene = tina.Scene((W, H), smoothing=True, maxfaces=2**22)
    grid = tina.MeshGrid(n)
    grid.pos.from_numpy(pos)
    scene.add_object(tina.MeshNoCulling(grid), tina.Classic())
    view, proj = scenes.default_camera(W / H)
    scene.engine.set_camera(view, proj)
    scene.render()
    fv, fn, _ = O.no_culling(O.grid_faces(pos), O.grid_faces(O.grid_normals(pos)))
    ref = O.render_scene([(fv, fn, None, tina.Classic())], W, H, view, proj, scene.lighting, _flags(O, smoothing=True))
    _check_frame(scene, ref)


def _soup_engine(tina, N, W, H, s, seed):
    import torch
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup_torch(N, W, H, s, seed, torch.device('cuda', torch.cuda.current_device()))
    engine = tina.Engine((W, H))
    engine.set_camera(view, proj)
    raster = tina.TriangleRaster(engine, maxfaces=N)
    raster.set_face_verts(tri)
    engine.clear_depth()
    raster.render_occup()
    torch.cuda.synchronize()
    return engine, raster, tri, (proj @ view).astype(np.float32)


def test_c3_full_size_ids_depth_bit_exact(tina, O):
    """C3 at BASELINE size: 16,777,216-face soup at 3840x2160, depth complexity ~8 (atomic contention): face ids and
    depth equal the serial oracle's on every pixel."""
    from taichi_three_b200 import multigpu as M
    W, H, N = 3840, 2160, 16 * 2**20
    engine, raster, tri, W2V = _soup_engine(tina, N, W, H, scenes.SOUP_S_C3, 20240601)
    occup, depth, tie, st = O.render_occup(tri.cpu().numpy(), W2V, W, H)
    d, o = M.unpack_keys(engine.keys.cpu().view(W, H))
    assert np.array_equal(d.numpy(), depth)
    assert np.array_equal(o.numpy(), occup)
    assert 7.0 < st['covered'] / (W * H) < 9.0  # the recipe's depth complexity


def test_c5_prefix_at_8k_ids_depth_bit_exact(tina, O):
    """C5: the 2^24-face prefix of the 134 M-face soup at 7680x4320 against the serial oracle."""
    from taichi_three_b200 import multigpu as M
    W, H, N = 7680, 4320, 2**24
    engine, raster, tri, W2V = _soup_engine(tina, N, W, H, scenes.SOUP_S_C5, 20240602)
    occup, depth, tie, st = O.render_occup(tri.cpu().numpy(), W2V, W, H)
    d, o = M.unpack_keys(engine.keys.cpu().view(W, H))
    assert np.array_equal(d.numpy(), depth)
    assert np.array_equal(o.numpy(), occup)
    assert (occup >= 0).mean() > 0.5


def test_c4_full_resolution_views(tina, O):
    """C4 at BASELINE resolution: cornell.gltf at 1024x1024, 8 of the 64 batch views (every eighth camera): ids + depth
    bit-exact, colour within 1e-4 (PBR + 512^2 texture)."""
    W = H = 1024
    cams = scenes.cornell_views(64)
    gltf = scenes.load_cornell()
    scene = tina.Scene((W, H), smoothing=True, texturing=True)
    gltf.extract(scene)
    objs = scenes.cornell_oracle_objects(gltf)
    for k in range(0, 64, 8):
        view, proj = cams[k]
        scene.engine.set_camera(view, proj)
        scene.render()
        ref = O.render_scene(objs, W, H, view, proj, scene.lighting, _flags(O, smoothing=True, texturing=True))
        _check_frame(scene, ref)


def test_probe_shader_and_fused_sinks(tina, O):
    """tina/probe.py: ProbeShader as a post-shader records the visible face id and its texture coordinate per pixel
    (elmid == raster.occup, texcoord == the TexcoordShader sink == the oracle's G-buffer); `touch` visits the disc
    around the cursor.  All sinks of the group travel in one launch (tina_raster_render_gbuffers): more than eight
    sinks split into two launches with the same results."""
    import torch
    from taichi_three_b200 import _lib
    W, H = 200, 160
    obj = scenes.load_monkey()
    scene = tina.Scene((W, H), smoothing=True, texturing=True)
    probe = tina.ProbeShader(scene.res)
    tc = torch.zeros((W, H, 2), device='cuda')
    extra = [torch.zeros((W, H, 3), device='cuda') for _ in range(7)]
    scene.post_shaders.append(probe)
    scene.post_shaders.append(tina.TexcoordShader(tina.Field(tc)))
    for b in extra:  # 2 + 2 + 7 = 11 sink parts: two launches
        scene.post_shaders.append(tina.NormalShader(tina.Field(b)))
    scene.add_object(tina.MeshModel(obj), tina.Classic())
    view, proj = scenes.default_camera(W / H)
    scene.engine.set_camera(view, proj)
    l0 = _lib.lib().tina_launch_count()
    scene.render()
    torch.cuda.synchronize()
    occup = scene.triangle_raster.occup.to_numpy()
    assert np.array_equal(probe.elmid.to_numpy(), occup) and (occup >= 0).sum() > 3000
    assert np.array_equal(probe.texcoord.to_numpy(), tc.cpu().numpy())
    for b in extra[1:]:
        assert torch.equal(b, extra[0])
    v, vn, vt = O.indexed(obj)
    W2V = (proj @ view).astype(np.float32)
    V2W = np.linalg.inv(proj @ view).astype(np.float32)
    flags = _flags(O, smoothing=True, texturing=True)
    o_occ, o_dep, _, _ = O.render_occup(v, W2V, W, H, flags)
    ref = O.render_gbuffer(5, v, vn, vt, o_occ, o_dep, W2V, V2W, W, H, flags, np.zeros((W, H, 2), np.float32))
    assert np.array_equal(o_occ, occup)
    assert np.abs(probe.texcoord.to_numpy() - ref).max() <= 1e-6
    # touch: every covered pixel within the radius, with its distance
    hits = []
    probe.touch(lambda p, I, r: hits.append((I, r)), 0.5, 0.5, 6.0)
    cx, cy = 0.5 * W, 0.5 * H
    want = {(x, y) for x in range(W) for y in range(H) if np.hypot(x - cx, y - cy) <= 6.0 and occup[x, y] != -1}
    assert {I for I, _ in hits} == want and len(want) > 50
    # the second frame clears the probe first (clear_buffer): an empty scene leaves elmid == -1
    scene2 = tina.Scene((W, H))
    probe2 = tina.ProbeShader(scene2.res)
    scene2.post_shaders.append(probe2)
    probe2.elmid.fill(5)
    scene2.render()
    assert (probe2.elmid.to_numpy() == -1).all()


def test_wireframe_only_scene_is_tonemapped(tina, O):
    """a scene whose single object goes through a raster without fused tonemap (wireframe) must still get the
    ACES pass (scene/raster.py:202-203 applies it unconditionally)."""
    import torch
    W, H = 96, 64
    view, proj = scenes.default_camera(W / H)
    tri = np.float32([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]])

    def frame(tonemap):
        scene = tina.Scene((W, H), tonemap=tonemap, linecolor=[2.0, 1.5, 0.5])
        mesh = tina.SimpleMesh()
        mesh.set_face_verts(tri)
        scene.add_object(tina.MeshToWire(mesh))
        scene.engine.set_camera(view, proj)
        scene.render()
        torch.cuda.synchronize()
        return scene.img.to_numpy()
    lin, tm = frame(False), frame(True)
    assert lin.max() > 1.0
    assert np.abs(tm - O.tonemap(lin)).max() <= 1e-6 and np.abs(tm - lin).max() > 0.1


def test_meshmodel_validates_indices_and_normal_adapters_follow_vertices(tina, O):
    """MeshModel rejects index buffers the kernels would read out of bounds with; [N,3,1] / [N,3,2] OBJ index
    layouts are padded like assimp/obj.py:73; MeshSmoothNormal / MeshFlatNormal recompute from the mesh's current vertices
    (mesh/norm.py:5-55)."""
    import torch
    obj = scenes.load_monkey()
    with pytest.raises(ValueError):
        tina.MeshModel({'v': obj['v'], 'f': np.asarray(obj['f'])[:, :, :1] + len(obj['v'])})
    with pytest.raises(ValueError):
        tina.MeshModel({'v': obj['v'], 'f': -np.ones((4, 3), int)})
    with pytest.raises(ValueError):
        tina.MeshModel({'v': obj['v'], 'f': np.zeros((4, 4, 3), int)})
    m1 = tina.MeshModel({'v': obj['v'], 'f': np.asarray(obj['f'])[:, :, :1]})  # `f 1 2 3` layout
    assert tuple(m1.faces.shape[1:]) == (3, 3) and int(m1.faces[:, :, 1:].abs().max()) == 0
    # smooth normals follow an edit of mesh.verts
    W, H = 160, 120
    view, proj = scenes.default_camera(W / H)
    model = tina.MeshModel(obj)
    sm = tina.MeshSmoothNormal(model, cached=False)
    scene = tina.Scene((W, H), smoothing=True, tonemap=False)
    scene.add_object(sm, tina.Classic())
    scene.engine.set_camera(view, proj)
    scene.render()
    a = scene.img.to_numpy().copy()
    model.verts[:, 2] *= 0.3  # flatten the head: different normals, different shading
    scene.render()
    torch.cuda.synchronize()
    b = scene.img.to_numpy()
    v = model.verts.cpu().numpy()
    f = np.asarray(obj['f'])[:, :, 0]
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]).astype(np.float32)
    fn = (fn * (np.float32(1) / np.sqrt((fn * fn).sum(1, keepdims=True)))).astype(np.float32)
    acc = np.zeros_like(v)
    for k in range(3):
        np.add.at(acc, f[:, k], fn)
    vn = (acc * (np.float32(1) / np.sqrt((acc * acc).sum(1, keepdims=True)))).astype(np.float32)
    ref = O.render_scene([(v[f], vn[f], None, tina.Classic())], W, H, view, proj, scene.lighting, _flags(O, smoothing=True), do_tonemap=False)
    assert np.abs(b - ref['image']).max() <= 5e-4  # (summation order of the per-vertex accumulation differs)
    assert np.abs(a - b).max() > 0.05


def test_random_material_graphs_match_reference_golden(tina):
    """The CUDA shading path (host flattening + folding + hoisting, specialised kernels or the interpreter) on the 36 random
    material graphs shaded by the reference's own sources: colour within 1e-4, ids / depth equal."""
    import os
    import torch
    from test_golden import GOLDEN, _lighting, _material
    g = np.load(os.path.join(GOLDEN, 'matgraphs_random.npz'))
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    engine = tina.Engine((W, H))
    engine.W2V[None], engine.V2W[None], engine.bias[None] = g['W2V'], g['V2W'], g['bias']
    raster = tina.TriangleRaster(engine, smoothing=bool(flags & 1), texturing=bool(flags & 2))
    mesh = tina.SimpleMesh()
    mesh.set_face_verts(g['verts0'])
    mesh.set_face_norms(g['norms0'])
    mesh.set_face_coors(g['coors0'])
    lighting = _lighting(tina, g)
    worst = 0.0
    for i in range(int(g['nspecs'])):
        img = tina.Field(torch.zeros((W, H, 3), device='cuda'))
        engine.clear_depth()
        raster.set_object(mesh)
        raster.render_occup()
        raster.render_color(tina.Shader(img, lighting, _material(tina, g, i, 'spec')))
        torch.cuda.synchronize()
        if i == 0:
            assert np.array_equal(raster.occup.to_numpy(), g['occup0']) and np.array_equal(engine.depth.to_numpy(), g['depth'])
        ref = g[f'image{i}']
        err = float((np.abs(img.to_numpy() - ref) / np.maximum(1.0, np.abs(ref))).max())  # (HDR highlights reach ~30 before the tonemap)
        assert err <= COLOR_TOL, (i, str(g[f'spec{i}']), err)
        worst = max(worst, err)
    assert worst <= COLOR_TOL


def test_procedural_textures_match_reference_golden(tina):
    """ChessboardTexture / LerpTexture nodes (TINA_OP_CHESS / TINA_OP_BCAST in the material programs) against the
    reference's own shading of six graphs: colour within 1e-4; generic_vm = 1 gives the same pixels."""
    import os
    import torch
    from test_golden import GOLDEN, _lighting, _material
    g = np.load(os.path.join(GOLDEN, 'matgraphs_proc.npz'))
    W, H = (int(v) for v in g['res'])
    flags = int(g['flags'])
    engine = tina.Engine((W, H))
    engine.W2V[None], engine.V2W[None], engine.bias[None] = g['W2V'], g['V2W'], g['bias']
    raster = tina.TriangleRaster(engine, smoothing=bool(flags & 1), texturing=bool(flags & 2))
    mesh = tina.SimpleMesh()
    mesh.set_face_verts(g['verts0'])
    mesh.set_face_norms(g['norms0'])
    mesh.set_face_coors(g['coors0'])
    lighting = _lighting(tina, g)
    for i in range(int(g['nspecs'])):
        imgs = []
        for vm in (0, 1):
            raster.set_tuning(generic_vm=vm)
            img = tina.Field(torch.zeros((W, H, 3), device='cuda'))
            engine.clear_depth()
            raster.set_object(mesh)
            raster.render_occup()
            raster.render_color(tina.Shader(img, lighting, _material(tina, g, i, 'spec')))
            torch.cuda.synchronize()
            imgs.append(img.to_numpy())
        ref = g[f'image{i}']
        err = float((np.abs(imgs[0] - ref) / np.maximum(1.0, np.abs(ref))).max())
        assert err <= COLOR_TOL, (i, str(g[f'spec{i}']), err)
        assert np.abs(imgs[0] - imgs[1]).max() <= COLOR_TOL, i
    with pytest.raises(Exception):
        tina.LerpTexture(z0=1)


def test_setup_cache_fields_match_reference_golden(tina):
    """raster.bcn / can / boo / coo / wsc (triangle.py:25-29), materialised on demand, equal what the reference's
    render_occup stored -- through the expanded-array path and through the indexed (MeshModel) path."""
    import os
    import torch
    from test_golden import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'setup_cache_monkey.npz'))
    W, H = (int(v) for v in g['res'])
    n = len(g['verts0'])
    for indexed in (False, True):
        engine = tina.Engine((W, H))
        engine.W2V[None], engine.V2W[None], engine.bias[None] = g['W2V'], g['V2W'], g['bias']
        raster = tina.TriangleRaster(engine, maxfaces=2048)
        if indexed:
            raster.set_object(tina.MeshModel(scenes.load_monkey()))
        else:
            mesh = tina.SimpleMesh(maxfaces=2048)
            mesh.set_face_verts(g['verts0'])
            raster.set_object(mesh)
        engine.clear_depth()
        raster.render_occup()
        for k in ('bcn', 'can', 'boo', 'coo', 'wsc'):
            out = getattr(raster, k).to_numpy()
            assert out.shape == (2048, 3 if k == 'wsc' else 2)
            assert np.array_equal(out[:n], g[k]), (indexed, k)
            assert not out[n:].any()


def test_nested_mesh_wrappers(tina, O):
    """Wrappers compose freely in the reference (mesh/trans.py:28-40, mesh/cull.py:5-66): MeshTransform inside MeshTransform
    (one rounding sequence per wrapper, NOT the product matrix), MeshFlipCulling around MeshNoCulling and the other way
    round, on a MeshModel and on a MeshGrid; raster.verts / norms equal the oracle's providers bit for bit and the frame
    equals the oracle's."""
    import torch
    W, H = 200, 150
    view, proj = tina.orbit_camera(radius=3.2, theta=0.3, phi=0.5, aspect=W / H)
    t1 = tina.translate([0.13, -0.21, 0.07]) @ tina.eularXYZ([0.41, -0.33, 0.27]) @ tina.scale([1.31, 0.77, 1.13])
    t2 = tina.eularXYZ([-0.2, 0.6, 0.1]) @ tina.scale(0.83) @ tina.translate([0.05, 0.02, -0.11])
    obj = scenes.load_monkey()
    v, vn, vt = O.indexed(obj)
    v1, n1 = O.transform(v, vn, t1)
    v2, n2 = O.transform(v1, n1, t2)
    # (a) nested transforms, then flip-culling around no-culling
    fv, fn, _ = O.no_culling(v2, n2)
    fv, fn = np.ascontiguousarray(fv[:, ::-1]), np.ascontiguousarray(fn[:, ::-1])
    for order in ('flip_outside', 'flip_inside'):
        base = tina.MeshTransform(tina.MeshTransform(tina.MeshModel(obj), t1), t2)
        mesh = tina.MeshFlipCulling(tina.MeshNoCulling(base)) if order == 'flip_outside' else tina.MeshNoCulling(tina.MeshFlipCulling(base))
        scene = tina.Scene((W, H), smoothing=True, maxfaces=4096)
        scene.add_object(mesh, tina.Classic())
        scene.engine.set_camera(view, proj)
        scene.render()
        torch.cuda.synchronize()
        assert np.array_equal(scene.triangle_raster.verts.to_numpy(), fv), order
        assert np.array_equal(scene.triangle_raster.norms.to_numpy(), fn), order
        ref = O.render_scene([(fv, fn, None, tina.Classic())], W, H, view, proj, scene.lighting, _flags(O, smoothing=True))
        _check_frame(scene, ref)
    # the product matrix is a different computation: the chain must not be collapsed
    vp, _ = O.transform(v, vn, t2 @ t1)
    assert not np.array_equal(vp, v2)
    # (b) a grid inside two transforms
    n = 40
    pos = scenes.wave_grid_pos(n)
    grid = tina.MeshGrid(n)
    grid.pos.from_numpy(pos)
    scene = tina.Scene((W, H), smoothing=True)
    scene.add_object(tina.MeshTransform(tina.MeshTransform(grid, t2), t1), tina.Classic())
    scene.engine.set_camera(view, proj)
    scene.render()
    gv, gn = O.transform(*O.transform(O.grid_faces(pos), O.grid_faces(O.grid_normals(pos)), t2), t1)
    ref = O.render_scene([(gv, gn, None, tina.Classic())], W, H, view, proj, scene.lighting, _flags(O, smoothing=True))
    _check_frame(scene, ref)
    assert np.array_equal(scene.triangle_raster.verts.to_numpy(), gv)


def test_frame_glue_fused_into_the_last_shading_pass(tina, O):
    """in a multi-object frame the ACES curve and the TAA accumulation ride along with the LAST object's
    shading pass (it finishes the pixels of the other objects and the background too), the fill with the first one's and
    the depth clear with the vertex stage.  Same images as the separate full-screen passes (scene.fuse_glue = False), and
    fewer launches."""
    import torch
    from taichi_three_b200 import _lib
    W, H = 192, 144
    view, proj = scenes.default_camera(W / H)
    a = scenes.soup(300, W, H, s=0.06, seed=21)
    obj = scenes.load_monkey()

    def run(fuse, taa, frames=3):
        np.random.seed(5)  # the TAA jitter (engine.py:31-39)
        scene = tina.Scene((W, H), taa=taa, bgcolor=[0.2, 0.1, 0.3])
        scene.fuse_glue = fuse
        m = tina.SimpleMesh()
        m.set_face_verts(a)
        scene.add_object(m, tina.Diffuse(color=[0.9, 0.5, 0.2]))
        scene.add_object(tina.MeshModel(obj), tina.Classic())
        scene.engine.set_camera(view, proj)
        l0 = _lib.lib().tina_launch_count()
        for _ in range(frames):
            scene.render()
        torch.cuda.synchronize()
        return scene.img.to_numpy().copy(), scene.image.to_numpy().copy(), (_lib.lib().tina_launch_count() - l0) / frames
    for taa in (False, True):
        img_f, raw_f, n_f = run(True, taa)
        img_s, raw_s, n_s = run(False, taa)
        assert np.abs(img_f - img_s).max() <= 2e-6, taa
        assert np.abs(raw_f - raw_s).max() <= 2e-6, taa
        assert n_f <= n_s - (2 if taa else 1), (taa, n_f, n_s)
    # and against the oracle (no TAA: centred samples)
    scene = tina.Scene((W, H), bgcolor=[0.2, 0.1, 0.3])
    m = tina.SimpleMesh()
    m.set_face_verts(a)
    scene.add_object(m, tina.Diffuse(color=[0.9, 0.5, 0.2]))
    v, _, _ = O.indexed(obj)
    scene.add_object(tina.MeshModel(obj), tina.Classic())
    scene.engine.set_camera(view, proj)
    scene.render()
    ref = O.render_scene([(a, None, None, tina.Diffuse(color=[0.9, 0.5, 0.2])), (v, None, None, tina.Classic())], W, H, view, proj,
                         scene.lighting, _flags(O), bgcolor=[0.2, 0.1, 0.3])
    _check_frame(scene, ref)
