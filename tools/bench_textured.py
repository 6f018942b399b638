#!/usr/bin/env python
"""K4 time of textured materials (prologue interpreter forms): MeshGrid(256) filling a 1024^2 screen, smooth normals,
texturing on, a 512^2 texture.  python tools/bench_textured.py  -> us per render_color for Classic / PBR / Lambert*tex."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import scenes, taichi_three_b200 as tina
from taichi_three_b200 import material as M
W = H = 1024
tex = np.random.default_rng(0).random((512, 512, 3)).astype(np.float32)
mats = {'Classic(color=tex)': tina.Classic(color=tina.Texture(tex)), 'PBR(basecolor=tex)': tina.PBR(basecolor=tina.Texture(tex)),
        'Lambert*tex + Phong*tex': tina.Lambert() * tina.Texture(tex) + tina.Phong(shineness=16) * tina.Texture(tex)}
for name, mat in mats.items():
    scene = tina.Scene((W, H), smoothing=True, texturing=True, tonemap=False)
    mesh = tina.MeshGrid(256)
    pos = scenes.wave_grid_pos(256)
    pos[..., :2] *= 1.8
    mesh.pos.from_numpy(pos)
    scene.add_object(mesh, mat)
    scene.engine.set_camera(*scenes.default_camera(1.0))
    raster = scene.triangle_raster
    for _ in range(3):
        scene.render()
    raster.set_tuning(profile=1)
    ts = []
    for _ in range(20):
        scene.render(); torch.cuda.synchronize()
        ts.append(raster.kernel_times()['render_color'] * 1e3)
    b, a, e, p, t = M.compile_material(mat)
    print(f'{name:28s} prologue {len(p):2d} slots  K4 median {np.median(ts):6.1f} us  covered {(scene.engine.depth.to_numpy() < 2**30).mean():.2f}', flush=True)
