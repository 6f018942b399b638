#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN SOURCES (/root/reference/tina, unmodified)
under the f32 NumPy emulation of the Taichi runtime in oracle/ref_shim (taichi itself cannot be
installed here).  Run in the build container only:  python tests/golden/make_golden.py

Each file holds, per object, the face arrays the reference's set_object produced
(raster.verts / norms / coors), the camera (W2V, V2W, bias), lights, a material spec string for
this repo's material classes, and the reference's outputs: per-object occup, final depth,
pre-tonemap image, final image (Scene.img).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
REF = '/root/reference'

from oracle import ref_shim  # noqa: E402

tina = ref_shim.load_tina(REF)


def render_and_dump(name, scene, specs, textures=()):
    """Replays Scene.render (scene/raster.py:168-207) step by step to capture per-object state."""
    eng = scene.engine
    scene.image.fill(scene.bgcolor)
    eng.clear_depth()
    out = {'res': np.array(scene.res.entries if hasattr(scene.res, 'entries') else scene.res, dtype=np.int32),
           'W2V': eng.W2V.to_numpy().astype(np.float32), 'V2W': eng.V2W.to_numpy().astype(np.float32),
           'bias': eng.bias.to_numpy().astype(np.float32), 'bgcolor': np.float32(scene.bgcolor),
           'nobjects': np.int32(len(scene.objects))}
    L = scene.lighting
    nl = int(L.nlights[None])
    out['light_dirs'] = L.light_dirs.to_numpy()[:nl].astype(np.float32)
    out['light_colors'] = L.light_colors.to_numpy()[:nl].astype(np.float32)
    out['ambient'] = L.ambient_color.to_numpy().astype(np.float32)
    for k, (obj, oinfo) in enumerate(scene.objects.items()):
        r = oinfo.raster
        r.set_object(obj)
        r.render_occup()
        r.render_color(scene.shaders[oinfo.material])
        n = int(r.nfaces[None])
        out[f'verts{k}'] = r.verts.to_numpy()[:n].astype(np.float32)
        if r.smoothing:
            out[f'norms{k}'] = r.norms.to_numpy()[:n].astype(np.float32)
        if r.texturing:
            out[f'coors{k}'] = r.coors.to_numpy()[:n].astype(np.float32)
        out[f'occup{k}'] = r.occup.to_numpy().astype(np.int32)
        out[f'material{k}'] = np.array(specs[k])
        out['flags'] = np.int32((1 if r.smoothing else 0) | (2 if r.texturing else 0) | (4 if r.culling else 0) |
                                (8 if r.clipping else 0))
    out['depth'] = eng.depth.to_numpy().astype(np.int32)
    out['image_pre_tonemap'] = scene.image.to_numpy().astype(np.float32)
    if scene.tonemap:
        scene.tonemap.apply(scene.image)
    out['image'] = scene.img.to_numpy().astype(np.float32)
    for i, t in enumerate(textures):
        out[f'tex{i}'] = np.asarray(t)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    cov = int((out['depth'] < 2**30).sum())
    print(f'{name}: {out["res"].tolist()} objects={len(scene.objects)} covered={cov} -> {os.path.getsize(path)} B')


def camera(scene, aspect, back=(0, 0, 3), pos=(0, 0, 0), fov=60):
    scene.engine.set_camera(tina.lookat(pos=pos, back=back), tina.perspective(fov, aspect))


def case_monkey():
    scene = tina.Scene((96, 96))
    scene.add_object(tina.MeshModel(os.path.join(REF, 'assets/monkey.obj')))
    camera(scene, 1.0)
    render_and_dump('monkey_flat_diffuse', scene, ['Diffuse()'])


def case_grid():
    scene = tina.Scene((80, 60), smoothing=True)
    mesh = tina.MeshGrid(14)
    pos = mesh.pos.to_numpy()
    xy = pos[..., :2].astype(np.float64)
    pos[..., 2] = (0.1 * np.sin(10 * np.sqrt((xy**2).sum(-1)) - 2 * np.pi * 0.25)).astype(np.float32)
    mesh.pos.from_numpy(pos)
    scene.add_object(tina.MeshNoCulling(mesh), tina.Classic())
    camera(scene, 80 / 60, back=(1.0, 1.5, 2.5))
    render_and_dump('grid_wave_nocull_smooth_classic', scene, ['Classic()'])
    np.save(os.path.join(HERE, 'grid_wave_pos.npy'), pos)


def case_cornell():
    import taichi_three_b200 as mine
    scene = tina.Scene((64, 64), smoothing=True, texturing=True)
    scene.load_gltf(os.path.join(REF, 'assets/cornell.gltf'))
    view, proj = mine.orbit_camera(center=(0, 2, 0), radius=6.0, theta=0.2, phi=0.7)
    scene.engine.set_camera(view, proj)
    g = tina.readgltf(os.path.join(REF, 'assets/cornell.gltf'))
    specs = ['PBR(basecolor=[0.8, 0.8, 0.8], metallic=M0, roughness=R0)'] * 2 + ['PBR(basecolor=Texture(tex0), metallic=M1, roughness=R1)']
    # exact factors straight from the file
    import json
    root = json.load(open(os.path.join(REF, 'assets/cornell.gltf')))
    m0, m1 = (m['pbrMetallicRoughness'] for m in root['materials'][:2])
    specs = [f"PBR(basecolor={m0['baseColorFactor'][:3]!r}, metallic={m0['metallicFactor']!r}, roughness={m0['roughnessFactor']!r})"] * 2
    specs.append(f"PBR(basecolor=Texture(tex0), metallic={m1['metallicFactor']!r}, roughness={m1['roughnessFactor']!r})")
    from PIL import Image
    import base64
    import io
    uri = root['images'][0].get('uri')
    if uri is None:
        bv = root['bufferViews'][root['images'][0]['bufferView']]
        buf = base64.b64decode(root['buffers'][bv['buffer']]['uri'].split('base64,')[1])
        data = buf[bv['byteOffset']:bv['byteOffset'] + bv['byteLength']]
    else:
        data = base64.b64decode(uri.split('base64,')[1])
    tex = np.swapaxes(np.array(Image.open(io.BytesIO(data))), 0, 1)
    render_and_dump('cornell_pbr_textured', scene, specs, textures=[tex])


EDGE = np.array([
    [[-9, -9, 0], [9, -9, 0], [0, 9, 0]],            # screen-filling, every vertex outside the NDC cube
    [[-0.5, -0.5, 5], [0.5, -0.5, 5], [0, 0.5, 5]],  # behind the camera
    [[-0.5, -0.5, 0], [0.5, -0.5, 4], [0, 0.5, 0]],  # crosses w = 0
    [[0, 0, 0], [0, 0, 0], [0, 0, 0]],               # degenerate
    [[np.nan, 0, 0], [1, 0, 0], [0, 1, 0]],          # NaN
    [[50, 50, 0], [51, 50, 0], [50, 51, 0]],         # far off-screen
    [[-1, -1, 0.5], [1, -1, 0.5], [0, 1, 0.5]],      # big, front-facing
    [[-1, -1, 0.2], [0, 1, 0.2], [1, -1, 0.2]],      # big, back-facing
    [[-0.8, -0.8, 0.8], [0.2, -0.8, 0.8], [-0.8, 0.2, 0.8]],  # overlaps, nearer
    [[1e-3, 0, 1], [2e-3, 0, 1], [1e-3, 1e-3, 1]],   # sub-pixel
    [[-1.2, 0.3, 0.1], [1.2, 0.3, 0.1], [0.0, 0.35, 0.1]],    # sliver across the screen
], dtype=np.float32)


def case_edges():
    import scenes
    soup = scenes.soup(60, 40, 28, s=0.1, seed=5)
    tri = np.ascontiguousarray(np.concatenate([EDGE, soup, EDGE[::-1]]))
    for culling in (True, False):
        for clipping in (True, False):
            scene = tina.Scene((40, 28), culling=culling, clipping=clipping, maxfaces=len(tri))
            mesh = tina.SimpleMesh(maxfaces=len(tri))
            with np.errstate(all='ignore'):
                mesh.set_face_verts(tri)
                scene.add_object(mesh)
                camera(scene, 40 / 28)
                render_and_dump(f'edges_cull{int(culling)}_clip{int(clipping)}', scene, ['Diffuse()'])


def case_multi_object():
    import scenes
    a = scenes.soup(40, 48, 36, s=0.12, seed=1)
    b = scenes.soup(40, 48, 36, s=0.12, seed=2)
    scene = tina.Scene((48, 36), maxfaces=64, bgcolor=0.25)
    specs = ['Diffuse(color=[1.0, 0.2, 0.1])', 'Classic(color=[0.1, 0.9, 0.2], shineness=8, specular=0.6)', 'Diffuse(color=[0.1, 0.2, 1.0])']
    mats = [tina.Diffuse(color=[1.0, 0.2, 0.1]), tina.Classic(color=[0.1, 0.9, 0.2], shineness=8, specular=0.6),
            tina.Diffuse(color=[0.1, 0.2, 1.0])]
    for tri, mat in zip((a, b, a.copy()), mats):
        m = tina.SimpleMesh(maxfaces=64)
        m.set_face_verts(tri)
        scene.add_object(m, mat)
    camera(scene, 48 / 36)
    scene.engine.bias[None] = [0.3, 0.8]  # engine.py:31-39 jittered sample position
    render_and_dump('multi_object_ties_bias', scene, specs)


def case_lights_materials():
    scene = tina.Scene((56, 56), smoothing=True, texturing=True)
    obj = tina.readobj(os.path.join(REF, 'assets/monkey.obj'))
    trans = tina.translate([0.2, -0.1, 0.3]) @ tina.eularXYZ([0.3, 0.8, -0.2]) @ tina.scale([0.9, 1.1, 0.8])
    mesh = tina.MeshFlipNormal(tina.MeshFlipCulling(tina.MeshTransform(tina.MeshModel(obj), trans)))
    mat = tina.Lambert() * [0.8, 0.5, 0.3] + tina.Emission() * 0.05 + tina.Phong(shineness=12) * 0.3
    scene.add_object(mesh, mat)
    scene.lighting.add_light(pos=[0.5, 0.8, 2.0], color=[0.3, 0.6, 0.9])
    scene.lighting.add_light(dir=[-1, 0.2, 0.5], color=[0.4, 0.1, 0.1])
    camera(scene, 1.0, back=(0.5, 0.3, 2.8))
    render_and_dump('monkey_transform_flip_lights_addmaterial', scene,
                    ['Lambert() * [0.8, 0.5, 0.3] + Emission() * 0.05 + Phong(shineness=12) * 0.3'])
    np.save(os.path.join(HERE, 'monkey_trans.npy'), trans)


def case_gbuffers():
    """ShaderGroup fan-out (shader.py:138-148, scene/raster.py:101-107): every G-buffer shader of
    core/shader.py:21-109 as a post-shader of one smooth, textured object."""
    ti = tina.ti
    scene = tina.Scene((52, 44), smoothing=True, texturing=True)
    res = scene.res
    bufs = {
        'const': (ti.field(int, res), lambda b: tina.ConstShader(b, 7)),
        'position': (ti.Vector.field(3, float, res), tina.PositionShader),
        'depth': (ti.field(float, res), tina.DepthShader),
        'normal': (ti.Vector.field(3, float, res), tina.NormalShader),
        'viewnormal': (ti.Vector.field(3, float, res), tina.ViewNormalShader),
        'texcoord': (ti.Vector.field(2, float, res), tina.TexcoordShader),
        'color': (ti.Vector.field(3, float, res), tina.ColorShader),
        'chessboard': (ti.field(float, res), lambda b: tina.ChessboardShader(b, 8)),
        'viewdir': (ti.Vector.field(3, float, res), tina.ViewdirShader),
        'simple': (ti.field(float, res), tina.SimpleShader),
    }
    for name, (buf, make) in bufs.items():
        scene.post_shaders.append(make(buf))
    obj = tina.readobj(os.path.join(REF, 'assets/monkey.obj'))
    scene.add_object(tina.MeshModel(obj), tina.Classic())
    camera(scene, 52 / 44, back=(0.8, 0.4, 2.6))
    for s in scene.post_shaders:
        s.clear_buffer()
    render_and_dump('gbuffer_shadergroup', scene, ['Classic()'])
    path = os.path.join(HERE, 'gbuffer_shadergroup.npz')
    d = dict(np.load(path))
    for name, (buf, _) in bufs.items():
        d['sink_' + name] = buf.to_numpy()
    np.savez_compressed(path, **d)


def case_particles():
    """ParticleRaster (core/particle.py) sharing depth with a triangle mesh: SimpleParticles with radii and
    colours (Classic material, Input('color') = particle colour), a transformed copy, then monkey.obj."""
    rng = np.random.default_rng(5)
    scene = tina.Scene((60, 48))
    n = 40
    pos = (rng.random((n, 3)) * 2 - 1).astype(np.float32) * np.float32([1.0, 0.8, 0.6])
    rad = (rng.random(n) * 0.12 + 0.03).astype(np.float32)
    col = (rng.random((n, 3)) * 0.8 + 0.2).astype(np.float32)
    pars = tina.SimpleParticles(maxpars=64)
    pars.set_particles(pos)
    pars.set_particle_radii(rad)
    pars.set_particle_colors(col)
    scene.add_object(pars, tina.Classic())
    pars2 = tina.SimpleParticles(maxpars=64, radius=0.05)
    pars2.set_particles(pos[:12] * np.float32(0.5))
    moved = tina.ParsTransform(pars2)
    trans = tina.translate([0.3, 0.2, 0.8]) @ tina.eularXYZ([0.2, 0.5, 0.1])
    moved.set_transform(trans, 1.7)
    scene.add_object(moved, tina.Diffuse())
    scene.add_object(tina.MeshModel(os.path.join(REF, 'assets/monkey.obj')), tina.Diffuse(color=[0.3, 0.5, 0.9]))
    camera(scene, 60 / 48, back=(0.4, 0.3, 3.0))
    # replay Scene.render by hand (render_and_dump is triangle specific)
    eng = scene.engine
    scene.image.fill(scene.bgcolor)
    eng.clear_depth()
    out = {'res': np.array(scene.res.entries, dtype=np.int32), 'W2V': eng.W2V.to_numpy().astype(np.float32),
           'V2W': eng.V2W.to_numpy().astype(np.float32), 'bias': eng.bias.to_numpy().astype(np.float32),
           'pos': pos, 'rad': rad, 'col': col, 'trans': trans}
    L = scene.lighting
    nl = int(L.nlights[None])
    out['light_dirs'] = L.light_dirs.to_numpy()[:nl].astype(np.float32)
    out['light_colors'] = L.light_colors.to_numpy()[:nl].astype(np.float32)
    out['ambient'] = L.ambient_color.to_numpy().astype(np.float32)
    for k, (obj, oinfo) in enumerate(scene.objects.items()):
        r = oinfo.raster
        r.set_object(obj)
        r.render_occup()
        r.render_color(scene.shaders[oinfo.material])
        out[f'occup{k}'] = r.occup.to_numpy().astype(np.int32)
        if hasattr(r, 'npars'):
            m = int(r.npars[None])
            out[f'pverts{k}'] = r.verts.to_numpy()[:m].astype(np.float32)
            out[f'psizes{k}'] = r.sizes.to_numpy()[:m].astype(np.float32)
            out[f'pcolors{k}'] = r.colors.to_numpy()[:m].astype(np.float32)
        else:
            out[f'verts{k}'] = r.verts.to_numpy()[:int(r.nfaces[None])].astype(np.float32)
        out[f'depth_after{k}'] = eng.depth.to_numpy().astype(np.int32)
        out[f'image_after{k}'] = scene.image.to_numpy().astype(np.float32)
    path = os.path.join(HERE, 'particles_and_mesh.npz')
    np.savez_compressed(path, **out)
    print('particles_and_mesh:', [int((out[f'occup{k}'] >= 0).sum()) for k in range(3)], os.path.getsize(path), 'B')


def case_wireframe():
    """WireframeRaster + MeshToWire (core/wireframe.py, mesh/wire.py) over a solid mesh: depth-tested lines."""
    scene = tina.Scene((64, 56), tonemap=False)
    obj = tina.readobj(os.path.join(REF, 'assets/monkey.obj'))
    solid = tina.MeshTransform(tina.MeshModel(obj), tina.scale(0.97))
    scene.add_object(solid, tina.Diffuse(color=[0.2, 0.3, 0.4]))
    wire = tina.MeshToWire(tina.MeshModel(obj))
    scene.add_object(wire)
    camera(scene, 64 / 56, back=(0.6, 0.5, 2.7))
    eng = scene.engine
    scene.image.fill(scene.bgcolor)
    eng.clear_depth()
    out = {'res': np.array(scene.res.entries, dtype=np.int32), 'W2V': eng.W2V.to_numpy().astype(np.float32),
           'V2W': eng.V2W.to_numpy().astype(np.float32), 'bias': eng.bias.to_numpy().astype(np.float32)}
    L = scene.lighting
    nl = int(L.nlights[None])
    out['light_dirs'] = L.light_dirs.to_numpy()[:nl].astype(np.float32)
    out['light_colors'] = L.light_colors.to_numpy()[:nl].astype(np.float32)
    out['ambient'] = L.ambient_color.to_numpy().astype(np.float32)
    for k, (o, oinfo) in enumerate(scene.objects.items()):
        r = oinfo.raster
        r.set_object(o)
        r.render_occup()
        r.render_color(scene.shaders[oinfo.material])
        if k == 0:
            out['verts0'] = r.verts.to_numpy()[:int(r.nfaces[None])].astype(np.float32)
        else:
            out['wires1'] = r.verts.to_numpy()[:int(r.nwires[None])].astype(np.float32)
        out[f'depth_after{k}'] = eng.depth.to_numpy().astype(np.int32)
        out[f'image_after{k}'] = scene.image.to_numpy().astype(np.float32)
    path = os.path.join(HERE, 'particles_wireframe_over_mesh.npz')
    np.savez_compressed(path, **out)
    changed = int((out['depth_after1'] != out['depth_after0']).sum())
    print('wireframe:', out['wires1'].shape, 'wire pixels', changed, os.path.getsize(path), 'B')


def case_postfx():
    """postp/fxaa.py and postp/blooming.py on a synthetic HDR image (particles_ prefix = not a raster scene)."""
    ti = tina.ti
    rng = np.random.default_rng(3)
    W, H = 72, 50
    img = (rng.random((W, H, 3)) ** 3 * 3).astype(np.float32)
    img[20:40, 10:30] += np.float32(2.5)  # a bright block: blooming threshold 1, strong FXAA edges
    fld = ti.Vector.field(3, float, (W, H))
    fld.from_numpy(img)
    fx = tina.FXAA((W, H))
    fx.apply(fld)
    out = {'input': img, 'fxaa': fld.to_numpy().astype(np.float32)}
    fld.from_numpy(img)
    bl = tina.Blooming((W, H))
    bl.apply(fld)
    out['bloom'] = fld.to_numpy().astype(np.float32)
    out['gwei'] = bl.gwei.to_numpy()[:int(bl.radius[None]) + 1].astype(np.float32)
    path = os.path.join(HERE, 'particles_postfx.npz')
    np.savez_compressed(path, **out)
    print('postfx: fxaa changed', int((np.abs(out['fxaa'] - img).max(-1) > 0).sum()), 'bloom mean', float((out['bloom'] - img).mean()))


def case_ssao():
    """postp/ssao.py inside Scene(ssao=True) (scene/raster.py:51-66,189-191): the normal G-buffer, the (random, but
    fixed at construction) sample and rotation tables, the AO field and the image before / after SSAO.apply."""
    np.random.seed(11)
    scene = tina.Scene((48, 40), smoothing=True, ssao=True, tonemap=False)
    scene.add_object(tina.MeshModel(os.path.join(REF, 'assets/monkey.obj')))
    camera(scene, 48 / 40)
    eng = scene.engine
    scene.image.fill(scene.bgcolor)
    eng.clear_depth()
    for sh in scene.pre_shaders + scene.post_shaders:
        sh.clear_buffer()
    for obj, oinfo in scene.objects.items():
        oinfo.raster.set_object(obj)
        oinfo.raster.render_occup()
        oinfo.raster.render_color(scene.shaders[oinfo.material])
    out = {'W2V': eng.W2V.to_numpy().astype(np.float32), 'V2W': eng.V2W.to_numpy().astype(np.float32),
           'depth': eng.depth.to_numpy().astype(np.int32), 'normals': scene.norm_buffer.to_numpy().astype(np.float32),
           'samples': scene.ssao.samples.to_numpy().astype(np.float32), 'rotations': scene.ssao.rotations.to_numpy().astype(np.float32),
           'image_before': scene.image.to_numpy().astype(np.float32)}
    scene.ssao.render(eng)
    out['ao'] = scene.ssao.img.to_numpy().astype(np.float32)
    scene.ssao.apply(scene.image)
    out['image_after'] = scene.image.to_numpy().astype(np.float32)
    path = os.path.join(HERE, 'particles_ssao.npz')  # (particles_ prefix = not a plain raster scene, see test_golden.CASES)
    np.savez_compressed(path, **out)
    print('particles_ssao: ao mean %.4f max %.4f, %d B' % (out['ao'].mean(), out['ao'].max(), os.path.getsize(path)))


def case_ssr():
    """postp/ssr.py inside Scene(ssr=True, texturing=True) (scene/raster.py:43-64,192-194): three objects with different
    material graphs (a textured PBR, Classic, an Add / Scale / Mix / Emission composite with a vector mix factor), the
    normal / texcoord / material-id G-buffers, the SSR field (non-TAA: WangHashRNG(P % blurring), ssr.py:73-76) and the
    image before / after SSR.apply.  The material table is filled by hand after add_object: the reference does that in a
    materialize callback, which real Taichi runs at the first kernel launch and the shim at construction.
    nsamples 6 / nsteps 16 instead of 32 / 32 keep the serial Python run short (the fields are public, ssr.py:8-12)."""
    rng = np.random.default_rng(20241018)
    tex0 = rng.random((6, 5, 3)).astype(np.float32)
    ns = {n: getattr(tina, n) for n in ('PBR', 'Classic', 'Diffuse', 'Lamp', 'Lambert', 'Phong', 'Emission', 'CookTorrance', 'Texture',
                                        'FresnelFactor', 'MixMaterial', 'ScaleMaterial', 'AddMaterial')}
    ns.update(tex0=tex0)
    specs = ['PBR(basecolor=Texture(tex0), metallic=0.7, roughness=0.15)',
             'Classic(color=[0.9, 0.85, 0.6], shineness=24, specular=0.5)',
             'MixMaterial(AddMaterial(ScaleMaterial(Lambert(), [0.2, 0.5, 0.8]), ScaleMaterial(Emission(), 0.3)), '
             'CookTorrance(roughness=0.3, fresnel=[0.9, 0.6, 0.4]), [0.3, 0.5, 0.7])']
    W, H = 48, 40
    scene = tina.Scene((W, H), smoothing=True, texturing=True, ssr=True, tonemap=False)
    mats = [eval(sp, dict(ns)) for sp in specs]
    scene.add_object(tina.MeshModel(os.path.join(REF, 'assets/monkey.obj')), mats[0])
    floor = tina.MeshTransform(tina.MeshModel(os.path.join(REF, 'assets/plane.obj')), tina.translate([0, -0.9, 0]) @ tina.scale(2.5))
    scene.add_object(floor, mats[1])
    ball = tina.MeshTransform(tina.MeshModel(os.path.join(REF, 'assets/sphere.obj')), tina.translate([1.3, -0.3, 0.2]) @ tina.scale(0.55))
    scene.add_object(ball, mats[2])
    camera(scene, W / H, back=(0.3, 0.8, 3.0))
    scene.mtltab.clear_materials()
    for m in scene.materials:
        scene.mtltab.add_material(m)
    scene.ssr.nsamples[None], scene.ssr.nsteps[None] = 6, 16
    eng = scene.engine
    scene.image.fill(scene.bgcolor)
    eng.clear_depth()
    for sh in scene.pre_shaders + scene.post_shaders:
        sh.clear_buffer()
    for obj, oinfo in scene.objects.items():
        oinfo.raster.set_object(obj)
        oinfo.raster.render_occup()
        oinfo.raster.render_color(scene.shaders[oinfo.material])
    out = {'W2V': eng.W2V.to_numpy().astype(np.float32), 'V2W': eng.V2W.to_numpy().astype(np.float32),
           'depth': eng.depth.to_numpy().astype(np.int32), 'normals': scene.norm_buffer.to_numpy().astype(np.float32),
           'coors': scene.coor_buffer.to_numpy().astype(np.float32), 'mtlid': scene.mtlid_buffer.to_numpy().astype(np.int32),
           'image_before': scene.image.to_numpy().astype(np.float32), 'tex0': tex0, 'nspecs': np.int32(len(specs)),
           'nsamples': np.int32(6), 'nsteps': np.int32(16), 'stepsize': np.float32(scene.ssr.stepsize[None]),
           'tolerance': np.float32(scene.ssr.tolerance[None]), 'blurring': np.int32(scene.ssr.blurring[None])}
    for i, sp in enumerate(specs):
        out[f'spec{i}'] = np.array(sp)
    scene.ssr.render(eng, scene.image)
    out['ssr'] = scene.ssr.img.to_numpy().astype(np.float32)
    scene.ssr.apply(scene.image)
    out['image_after'] = scene.image.to_numpy().astype(np.float32)
    path = os.path.join(HERE, 'particles_ssr.npz')  # (particles_ prefix = not a plain raster scene, see test_golden.CASES)
    np.savez_compressed(path, **out)
    hit = out['ssr'][..., 3] > 0
    print('particles_ssr: %d px with a normal, %d with hits, mean alpha %.4f, materials hit %s, %d B' % (
        int((np.square(out['normals']).sum(-1) >= 1e-6).sum()), int(hit.sum()), float(out['ssr'][..., 3].mean()),
        sorted(set(out['mtlid'][hit].tolist())), os.path.getsize(path)))


def case_ssr_defaults():
    """postp/ssr.py with its DEFAULT parameters (32 samples x 32 steps, stepsize 2, tolerance 15, blurring 4; ssr.py:20-28)
    in a Scene without texturing (the dummy (1, 1) texcoord field of scene/raster.py:63-64): Diffuse floor, metallic PBR
    monkey, at 32 x 24 so that the serial Python run stays around two minutes."""
    ns = {n: getattr(tina, n) for n in ('PBR', 'Classic', 'Diffuse', 'Lamp', 'Lambert', 'Phong', 'Emission', 'CookTorrance', 'Texture',
                                        'FresnelFactor', 'MixMaterial', 'ScaleMaterial', 'AddMaterial')}
    specs = ['PBR(basecolor=[0.9, 0.7, 0.5], metallic=0.9, roughness=0.1)', 'Diffuse(color=[0.4, 0.6, 0.8])']
    W, H = 32, 24
    scene = tina.Scene((W, H), smoothing=True, ssr=True, tonemap=False)
    mats = [eval(sp, dict(ns)) for sp in specs]
    scene.add_object(tina.MeshModel(os.path.join(REF, 'assets/monkey.obj')), mats[0])
    floor = tina.MeshTransform(tina.MeshModel(os.path.join(REF, 'assets/plane.obj')), tina.translate([0, -0.9, 0]) @ tina.scale(2.5))
    scene.add_object(floor, mats[1])
    camera(scene, W / H, back=(0.2, 0.9, 3.0))
    scene.mtltab.clear_materials()
    for m in scene.materials:
        scene.mtltab.add_material(m)
    eng = scene.engine
    scene.image.fill(scene.bgcolor)
    eng.clear_depth()
    for sh in scene.pre_shaders + scene.post_shaders:
        sh.clear_buffer()
    for obj, oinfo in scene.objects.items():
        oinfo.raster.set_object(obj)
        oinfo.raster.render_occup()
        oinfo.raster.render_color(scene.shaders[oinfo.material])
    out = {'W2V': eng.W2V.to_numpy().astype(np.float32), 'V2W': eng.V2W.to_numpy().astype(np.float32),
           'depth': eng.depth.to_numpy().astype(np.int32), 'normals': scene.norm_buffer.to_numpy().astype(np.float32),
           'mtlid': scene.mtlid_buffer.to_numpy().astype(np.int32), 'image_before': scene.image.to_numpy().astype(np.float32),
           'nspecs': np.int32(len(specs)), 'nsamples': np.int32(scene.ssr.nsamples[None]), 'nsteps': np.int32(scene.ssr.nsteps[None]),
           'stepsize': np.float32(scene.ssr.stepsize[None]), 'tolerance': np.float32(scene.ssr.tolerance[None]),
           'blurring': np.int32(scene.ssr.blurring[None])}
    for i, sp in enumerate(specs):
        out[f'spec{i}'] = np.array(sp)
    scene.ssr.render(eng, scene.image)
    out['ssr'] = scene.ssr.img.to_numpy().astype(np.float32)
    scene.ssr.apply(scene.image)
    out['image_after'] = scene.image.to_numpy().astype(np.float32)
    path = os.path.join(HERE, 'particles_ssr_defaults.npz')
    np.savez_compressed(path, **out)
    print('particles_ssr_defaults: %d px with hits, mean alpha %.4f, %d B' % (int((out['ssr'][..., 3] > 0).sum()),
                                                                              float(out['ssr'][..., 3].mean()), os.path.getsize(path)))


def case_prims():
    """mesh/prim.py:21-87: the face lists PrimitiveMesh.sphere / .cylinder build (vertex, normal, texcoord per corner),
    captured from the reference's own generator before they reach its Taichi fields."""
    import importlib
    prim = importlib.import_module('tina.mesh.prim')

    class Capture:
        def __new__(cls, faces):
            return np.array(faces, dtype=np.float32)
    out = {}
    for name, args in (('sphere_8_6', (8, 6, 1)), ('sphere_5_3', (5, 3, 0.7)), ('sphere_default', ())):
        out[name] = prim.PrimitiveMesh.sphere.__func__(Capture, *args)
    for name, args in (('cylinder_8_2', (8, 2, 1, 2)), ('cylinder_5_3', (5, 3, 0.6, 1.5)), ('cylinder_default', ())):
        out[name] = prim.PrimitiveMesh.cylinder.__func__(Capture, *args)
    path = os.path.join(HERE, 'particles_prims.npz')  # (particles_ prefix = not a rendered raster scene, see test_golden.CASES)
    np.savez_compressed(path, **out)
    print('particles_prims:', {k: v.shape for k, v in out.items()}, os.path.getsize(path), 'B')


def case_micro():
    """The C2 regime at golden size: sub-pixel faces.  (a) MeshGrid(56) wave on a 40x30 screen (~0.3 px per face,
    smooth normals, Classic) -- most faces cover no sample, many samples lie within 1e-2 px of an edge;
    (b) adversarial micro-triangles (vertices within 1e-3..1e-1 px of sample centres, needles, slivers, near-collinear;
    tests/test_gpu_parity.py::_adversarial_triangles) with a jittered sample bias, culling on and off."""
    import scenes
    from test_gpu_parity import _adversarial_triangles
    scene = tina.Scene((40, 30), smoothing=True, maxfaces=2 * 55 * 55)
    mesh = tina.MeshGrid(56)
    pos = scenes.wave_grid_pos(56, t=0.4)
    mesh.pos.from_numpy(pos)
    scene.add_object(mesh, tina.Classic())
    camera(scene, 40 / 30)
    np.save(os.path.join(HERE, 'micro_grid_pos.npy'), pos)
    render_and_dump('micro_grid_wave_smooth_classic', scene, ['Classic()'])
    W, H = 32, 24
    import taichi_three_b200 as mine
    view, proj = np.asarray(mine.lookat(back=(0, 0, 3)), np.float32), np.asarray(mine.perspective(60, W / H), np.float32)
    tri = _adversarial_triangles(np.random.default_rng(77), 1500, W, H, view, proj)
    for culling in (True, False):
        scene = tina.Scene((W, H), culling=culling, maxfaces=len(tri))
        m = tina.SimpleMesh(maxfaces=len(tri))
        m.set_face_verts(tri)
        scene.add_object(m)
        camera(scene, W / H)
        scene.engine.bias[None] = [0.37, 0.81]
        render_and_dump(f'micro_adversarial_cull{int(culling)}', scene, ['Diffuse()'])


def case_setup_cache():
    """The public per-face setup cache of the reference's TriangleRaster (triangle.py:25-29, written at :127-131):
    bcn / can / boo / coo / wsc after render_occup, culling + clipping on (rows of rejected faces stay zero)."""
    scene = tina.Scene((72, 64))
    scene.add_object(tina.MeshModel(os.path.join(REF, 'assets/monkey.obj')))
    camera(scene, 72 / 64, back=(0.6, 0.2, 2.7))
    render_and_dump('setup_cache_monkey', scene, ['Diffuse()'])
    r = scene.triangle_raster
    n = int(r.nfaces[None])
    path = os.path.join(HERE, 'setup_cache_monkey.npz')
    d = dict(np.load(path))
    for k in ('bcn', 'can', 'boo', 'coo', 'wsc'):
        d[k] = getattr(r, k).to_numpy()[:n].astype(np.float32)
    np.savez_compressed(path, **d)


def random_material_spec(rng, depth=0):
    """A random material graph as an expression over the node classes both code bases share."""
    def color():
        c = rng.integers(0, 4)
        if c == 0:
            return 'Texture(tex0)'
        if c == 1:
            return "'color'"
        return repr([round(float(x), 3) for x in rng.uniform(0.05, 1.0, 3)])

    def factor():
        c = rng.integers(0, 5)
        if c == 0:
            return 'Texture(tex1)'
        if c == 1:
            return f'FresnelFactor(metallic={round(float(rng.uniform(0, 1)), 3)}, albedo={color()}, specular={round(float(rng.uniform(0.2, 0.8)), 3)})'
        return repr(round(float(rng.uniform(0.05, 0.95)), 3))
    leafs = ['Lambert()', 'Emission()',
             lambda: f'Phong(shineness={int(rng.integers(2, 48))})',
             lambda: f'CookTorrance(roughness={round(float(rng.uniform(0.25, 0.9)), 3)}, fresnel={factor()})',
             lambda: f'Diffuse(color={color()})',
             lambda: f'Classic(color={color()}, shineness={int(rng.integers(2, 48))}, specular={round(float(rng.uniform(0.1, 0.7)), 3)})',
             lambda: f'PBR(basecolor={color()}, metallic={round(float(rng.uniform(0, 1)), 3)}, roughness={round(float(rng.uniform(0.3, 0.9)), 3)})',
             lambda: f'Lamp(color={color()})']
    if depth >= 2 or rng.random() < 0.25:
        leaf = leafs[rng.integers(0, len(leafs))]
        return leaf if isinstance(leaf, str) else leaf()
    c = rng.integers(0, 3)
    a, b = random_material_spec(rng, depth + 1), random_material_spec(rng, depth + 1)
    if c == 0:
        return f'MixMaterial({a}, {b}, {factor()})'
    if c == 1:
        return f'ScaleMaterial({a}, {color() if rng.random() < 0.5 else factor()})'
    return f'AddMaterial({a}, {b})'


def case_matgraphs():
    """36 random material graphs (Mix / Scale / Add over Lambert, Phong, CookTorrance, Emission, the Diffuse / Classic /
    PBR / Lamp constructors, colour and scalar textures, Fresnel factors as mix factors) shaded by the reference's own
    matr/material.py + matr/nodes.py + core/lighting.py on one small smooth, textured object under a directional and a
    point light.  Pins the oracle's independent material front-end (oracle/materials.py) and, through it, the CUDA path."""
    rng = np.random.default_rng(20241017)
    tex0 = rng.random((6, 5, 3)).astype(np.float32)
    tex1 = rng.random((4, 7)).astype(np.float32)[:, :, None] * np.ones((1, 1, 3), np.float32)  # a scalar texture, as RGB
    ns = {n: getattr(tina, n) for n in ('PBR', 'Classic', 'Diffuse', 'Lamp', 'Lambert', 'Phong', 'Emission', 'CookTorrance', 'Texture',
                                        'FresnelFactor', 'MixMaterial', 'ScaleMaterial', 'AddMaterial')}
    ns.update(tex0=tex0, tex1=tex1)
    specs = [random_material_spec(rng) for _ in range(36)]
    out = None
    for i, spec in enumerate(specs):
        scene = tina.Scene((36, 30), smoothing=True, texturing=True, tonemap=False)
        mesh = tina.MeshGrid(9)
        pos = mesh.pos.to_numpy()
        xy = pos[..., :2].astype(np.float64)
        pos[..., 2] = (0.25 * np.sin(4 * xy[..., 0]) * np.cos(3 * xy[..., 1])).astype(np.float32)
        mesh.pos.from_numpy(pos)
        scene.add_object(mesh, eval(spec, dict(ns)))
        scene.lighting.add_light(pos=[0.4, 0.6, 1.5], color=[0.5, 0.7, 0.9])
        camera(scene, 36 / 30, back=(0.4, 0.7, 2.2))
        render_and_dump('matgraphs_tmp', scene, [spec], textures=[tex0, tex1])
        d = dict(np.load(os.path.join(HERE, 'matgraphs_tmp.npz')))
        if out is None:
            out = {k: d[k] for k in ('res', 'W2V', 'V2W', 'bias', 'bgcolor', 'light_dirs', 'light_colors', 'ambient', 'verts0', 'norms0',
                                     'coors0', 'occup0', 'depth', 'flags', 'tex0', 'tex1')}
            out['nspecs'] = np.int32(len(specs))
        else:
            assert np.array_equal(d['occup0'], out['occup0'])
        out[f'spec{i}'] = np.array(spec)
        out[f'image{i}'] = d['image_pre_tonemap']
    os.remove(os.path.join(HERE, 'matgraphs_tmp.npz'))
    np.savez_compressed(os.path.join(HERE, 'matgraphs_random.npz'), **out)
    print('matgraphs_random:', len(specs), 'graphs;', os.path.getsize(os.path.join(HERE, 'matgraphs_random.npz')), 'B')


def case_proc_textures():
    """The procedural parameter nodes ChessboardTexture and LerpTexture (matr/nodes.py:114-136; tests/probe.py,
    examples/meshgrid_cloth.py use them) as colours, mix factors and roughness of the stock materials, shaded by the
    reference's own sources on the object / lights / camera of case_matgraphs.  Same file layout as matgraphs_random.npz."""
    ns = {n: getattr(tina, n) for n in ('PBR', 'Classic', 'Diffuse', 'Lamp', 'Lambert', 'Phong', 'Emission', 'CookTorrance', 'Texture',
                                        'FresnelFactor', 'MixMaterial', 'ScaleMaterial', 'AddMaterial', 'ChessboardTexture', 'LerpTexture')}
    specs = ['Diffuse(color=LerpTexture(x0=[1.0, 1.0, 1.0], x1=[0.0, 0.0, 1.0]))',
             'Classic(color=ChessboardTexture(size=0.2, color0=[0.9, 0.2, 0.1], color1=0.8))',
             'Diffuse(color=ChessboardTexture())',
             'PBR(basecolor=ChessboardTexture(size=[0.25, 0.125], color0=[0.2, 0.3, 0.9], color1=[0.9, 0.8, 0.2]), metallic=0.3, roughness=LerpTexture(x0=0.1, x1=0.3))',
             'MixMaterial(ScaleMaterial(Lambert(), [0.8, 0.4, 0.2]), Phong(shineness=12), ChessboardTexture(size=0.3, color0=0.1, color1=0.7))',
             'AddMaterial(ScaleMaterial(Lambert(), LerpTexture(x0=[0.1, 0.2, 0.3], x1=[0.4, 0.3, 0.2])), ScaleMaterial(Emission(), ChessboardTexture(size=0.15, color0=0.0, color1=0.2)))']
    out = None
    for i, spec in enumerate(specs):
        scene = tina.Scene((36, 30), smoothing=True, texturing=True, tonemap=False)
        mesh = tina.MeshGrid(9)
        pos = mesh.pos.to_numpy()
        xy = pos[..., :2].astype(np.float64)
        pos[..., 2] = (0.25 * np.sin(4 * xy[..., 0]) * np.cos(3 * xy[..., 1])).astype(np.float32)
        mesh.pos.from_numpy(pos)
        scene.add_object(mesh, eval(spec, dict(ns)))
        scene.lighting.add_light(pos=[0.4, 0.6, 1.5], color=[0.5, 0.7, 0.9])
        camera(scene, 36 / 30, back=(0.4, 0.7, 2.2))
        render_and_dump('matgraphs_tmp', scene, [spec])
        d = dict(np.load(os.path.join(HERE, 'matgraphs_tmp.npz')))
        if out is None:
            out = {k: d[k] for k in ('res', 'W2V', 'V2W', 'bias', 'bgcolor', 'light_dirs', 'light_colors', 'ambient', 'verts0', 'norms0',
                                     'coors0', 'occup0', 'depth', 'flags')}
            out['nspecs'] = np.int32(len(specs))
        out[f'spec{i}'] = np.array(spec)
        out[f'image{i}'] = d['image_pre_tonemap']
    os.remove(os.path.join(HERE, 'matgraphs_tmp.npz'))
    np.savez_compressed(os.path.join(HERE, 'matgraphs_proc.npz'), **out)
    print('matgraphs_proc:', len(specs), 'graphs;', os.path.getsize(os.path.join(HERE, 'matgraphs_proc.npz')), 'B')


if __name__ == '__main__':
    np.seterr(all='ignore')
    if len(sys.argv) > 1:  # python make_golden.py micro ...  -> only these cases
        for name in sys.argv[1:]:
            globals()['case_' + name]()
        sys.exit(0)
    case_micro()
    case_ssao()
    case_ssr()
    case_ssr_defaults()
    case_prims()
    case_monkey()
    case_grid()
    case_cornell()
    case_edges()
    case_multi_object()
    case_lights_materials()
    case_gbuffers()
    case_particles()
    case_wireframe()
    case_postfx()
    case_setup_cache()
    case_matgraphs()
    case_proc_textures()
