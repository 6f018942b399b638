#!/usr/bin/env python
"""C2 step time when consecutive steps work on DIFFERENT scenes (inputs larger than L2, no flush kernel between steps):
python tools/rotate_scenes.py [nscenes ...]   -- whole-region CUDA events / K, one stream, PDL across frames."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import scenes, taichi_three_b200 as tina
W, H, n = 1920, 1080, 1024
bg = np.zeros(3, np.float32)
def make(k):
    scene = tina.Scene((W, H), smoothing=True, maxfaces=2**21, tonemap=False)
    mesh = tina.MeshGrid(n); mesh.pos.from_numpy(scenes.wave_grid_pos(n, t=0.25 + 0.01 * k))
    mat = tina.Classic(); scene.add_object(mesh, mat)
    scene.engine.set_camera(*scenes.default_camera(W / H))
    raster, shader = scene.triangle_raster, scene.shaders[id(mat)]
    raster.set_object(mesh)
    if os.environ.get('TINA_PERSIST') is not None: raster.set_tuning(persist_k4=int(os.environ['TINA_PERSIST']))
    if os.environ.get('TINA_OVERLAP') is not None: raster.set_tuning(overlap_vertex=int(os.environ['TINA_OVERLAP']))
    def step():
        scene.engine.clear_depth(); raster.render_occup(); raster.render_color(shader, fill_bg=bg)
    return scene, step
counts = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 6, 8]
pool = [make(k) for k in range(max(counts))]
K = 240
for rep in range(2):
    for ns in counts:
        steps = [p[1] for p in pool[:ns]]
        for i in range(24): steps[i % ns]()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        a.record()
        t0 = time.perf_counter()
        for i in range(K): steps[i % ns]()
        host = (time.perf_counter() - t0) / K * 1e6
        b.record(); torch.cuda.synchronize()
        print(f'{ns} scenes in rotation ({ns * 100} MB touched per cycle): {a.elapsed_time(b) / K * 1e3:.2f} us / step (host enqueue {host:.1f} us / step)', flush=True)
