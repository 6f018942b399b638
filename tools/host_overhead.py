#!/usr/bin/env python
"""CPU time per API call on a tiny scene (launch-bound regime): which knob costs host time?"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import scenes, taichi_three_b200 as tina
scene = tina.Scene((512, 512), tonemap=False)
mesh = tina.MeshModel(scenes.load_monkey()); scene.add_object(mesh)
scene.engine.set_camera(*scenes.default_camera())
raster, shader = scene.triangle_raster, scene.shaders[id(scene.default_material)]
bg = np.zeros(3, np.float32)
def run(label, n=2000, **knobs):
    raster.set_tuning(**knobs)
    raster.set_object(mesh)
    for _ in range(50):
        scene.engine.clear_depth(); raster.render_occup(); raster.render_color(shader, fill_bg=bg)
    torch.cuda.synchronize()
    tc = to = tk = 0.0
    t0 = time.perf_counter()
    for _ in range(n):
        a = time.perf_counter(); scene.engine.clear_depth()
        b = time.perf_counter(); raster.render_occup()
        c = time.perf_counter(); raster.render_color(shader, fill_bg=bg)
        d = time.perf_counter()
        tc += b - a; to += c - b; tk += d - c
    cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(f'{label:28s} cpu/step {cpu/n*1e6:6.1f} us (clear {tc/n*1e6:5.1f} occup {to/n*1e6:5.1f} color {tk/n*1e6:5.1f})  wall/step {tot/n*1e6:6.1f} us')
run('default')
run('pdl=0', pdl=0)
run('pdl=0 adaptive=0', pdl=0, adaptive=0)
run('pdl=0 adaptive=0 indexed=0', pdl=0, adaptive=0, indexed=0)
run('pdl=1 adaptive=1 indexed=0', pdl=1, adaptive=1, indexed=0)
run('default again', pdl=1, adaptive=1, indexed=1)
