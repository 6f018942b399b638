#!/usr/bin/env python
"""A/B a tuning knob on the C2 step: python tools/ab_knob.py pdl 0 1  (prints ms/step per value)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import scenes, taichi_three_b200 as tina
knob, values = sys.argv[1], [int(v) for v in sys.argv[2:]]
W, H, n = 1920, 1080, 1024
scene = tina.Scene((W, H), smoothing=True, maxfaces=2**21, tonemap=False)
mesh = tina.MeshGrid(n); mesh.pos.from_numpy(scenes.wave_grid_pos(n))
mat = tina.Classic(); scene.add_object(mesh, mat)
scene.engine.set_camera(*scenes.default_camera(W / H))
raster, shader = scene.triangle_raster, scene.shaders[id(mat)]
raster.set_object(mesh)
flush = torch.empty(64 * 2**20, device='cuda')
bg = np.zeros(3, np.float32)
def step():
    scene.engine.clear_depth(); raster.render_occup(); raster.render_color(shader, fill_bg=bg)
for rep in range(2):
    for v in values:
        raster.set_tuning(**{knob: v})
        for _ in range(10): step()
        ts = []
        for _ in range(200):
            flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record()
            ts.append((a, b))
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in ts]) * 1e3
        raster.set_tuning(profile=1); step(); torch.cuda.synchronize()
        kt = {k: round(x * 1e3, 1) for k, x in raster.kernel_times().items() if x > 0}
        raster.set_tuning(profile=0)
        ts_ = np.sort(t)[10:-10]
        print(f'{knob}={v}: mean {t.mean():.2f} trimmed {ts_.mean():.2f} median {np.median(t):.1f} us  min {t.min():.1f} us  kernels(us) {kt}', flush=True)
