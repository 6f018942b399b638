#!/usr/bin/env python
"""Time of the SSR pass (default 32 samples x 32 steps) and of SSAO on a 1080p frame: monkey over a mirror floor."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import scenes, taichi_three_b200 as tina
W, H = 1920, 1080
for opts in (dict(ssr=True), dict(ssao=True), dict()):
    scene = tina.Scene((W, H), smoothing=True, **opts)
    scene.add_object(tina.MeshModel(scenes.load_monkey()), tina.PBR(metallic=0.2, roughness=0.3))
    floor = tina.MeshTransform(tina.MeshGrid(64), tina.translate([0, -1, 0]) @ tina.scale(3) @ tina.eularXYZ([-np.pi / 2, 0, 0]))
    scene.add_object(floor, tina.PBR(metallic=1.0, roughness=0.05))
    scene.engine.set_camera(*tina.orbit_camera(radius=3.5, theta=0.45, phi=0.3, aspect=W / H))
    for _ in range(3): scene.render()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): scene.render()
    b.record(); torch.cuda.synchronize()
    extra = ''
    if 'ssr' in opts:
        im = scene.ssr.img.to_numpy()
        extra = f'  pixels with a normal {(np.square(scene.norm_buffer.to_numpy()).sum(-1) > 1e-6).sum()}, with hits {(im[..., 3] > 0).sum()}'
    print(f'{opts or "plain"}: Scene.render {a.elapsed_time(b) / 10:.3f} ms / frame{extra}')
