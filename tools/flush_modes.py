#!/usr/bin/env python
"""C2 step time under different L2 states at the start of the step: python tools/flush_modes.py
  none   back to back;  write  256 MiB fill (L2 left full of DIRTY lines);  write+read  the fill followed by a 256 MiB read
  sweep (L2 left full of CLEAN foreign lines: cold like ncu's cache-control, no write-back owed by the step)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import scenes, taichi_three_b200 as tina
W, H, n = 1920, 1080, 1024
scene = tina.Scene((W, H), smoothing=True, maxfaces=2**21, tonemap=False)
mesh = tina.MeshGrid(n); mesh.pos.from_numpy(scenes.wave_grid_pos(n))
mat = tina.Classic(); scene.add_object(mesh, mat)
scene.engine.set_camera(*scenes.default_camera(W / H))
raster, shader = scene.triangle_raster, scene.shaders[id(mat)]
raster.set_object(mesh)
flush = torch.empty(64 * 2**20, device='cuda')
sweep = torch.ones(64 * 2**20, device='cuda')
sink = torch.zeros(1, device='cuda')
bg = np.zeros(3, np.float32)
def step():
    scene.engine.clear_depth(); raster.render_occup(); raster.render_color(shader, fill_bg=bg)
def prep(mode):
    if mode in ('write', 'write+read'): flush.fill_(1.0)
    if mode in ('read', 'write+read'): sweep.amax()
for rep in range(2):
    for mode in ('none', 'write', 'read', 'write+read'):
        for _ in range(10): prep(mode); step()
        ts = []
        for _ in range(200):
            prep(mode)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record()
            ts.append((a, b))
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in ts]) * 1e3
        print(f'{mode:11s}: mean {t.mean():.2f} median {np.median(t):.2f} min {t.min():.2f} us', flush=True)
