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
shader = tina.Shader(tina.Field(torch.zeros((W, H, 3), device='cuda')), scene.lighting, mat)
flush = torch.empty(64 * 2**20, device='cuda')
bg = np.zeros(3, np.float32)
scene.engine.clear_depth(); raster.render_occup(); torch.cuda.synchronize()
keys = scene.engine.keys
cov = ((keys & 0xffffffff) != 0)
xs = cov.any(1).nonzero().flatten()
x0, x1 = int(xs.min()), int(xs.max()) + 1
p0 = (x0 * H) // 256 * 256; p1 = min(W * H, ((x1 * H) + 255) // 256 * 256)
print('covered px', int(cov.sum()), 'columns', x0, x1, 'range chunks', (p1 - p0) // 256)
def timeit(fn, flushit=True, reps=100):
    ts = []
    for _ in range(reps):
        if flushit: flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ts]) * 1e3
    return float(np.median(t)), float(t.min())
print('full render_color (flags, fill)   ', timeit(lambda: raster.render_color(shader, fill_bg=bg)))
print('full render_color no flush        ', timeit(lambda: raster.render_color(shader, fill_bg=bg), False))
print('range over covered columns        ', timeit(lambda: raster.render_color_range(shader, p0, p1 - p0, fill_bg=bg)))
print('range over covered cols no flush  ', timeit(lambda: raster.render_color_range(shader, p0, p1 - p0, fill_bg=bg), False))
print('range whole screen (no flags)     ', timeit(lambda: raster.render_color_range(shader, 0, W * H, fill_bg=bg)))
print('empty kernel launch (fill image)  ', timeit(lambda: shader.img.to_torch().zero_()))
raster.set_tuning(fast_shading=0)
print('exact: full render_color          ', timeit(lambda: raster.render_color(shader, fill_bg=bg)))
