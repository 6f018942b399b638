"""Where K4's time goes on C2: whole frame vs the column range the mesh covers vs the background ranges."""
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
raster, shader = scene.triangle_raster, scene.shaders[id(mat)].shaders[0]
raster.set_object(mesh)
flush = torch.empty(64 * 2**20, device='cuda')
bg = np.zeros(3, np.float32)
scene.engine.clear_depth(); raster.render_occup(); torch.cuda.synchronize()
occ = raster.occup.to_torch()
cols = (occ >= 0).any(dim=1).nonzero().flatten()
c0, c1 = int(cols.min()), int(cols.max()) + 1
print('covered pixels', int((occ >= 0).sum()), 'columns', c0, c1)
chunk = 256
lo = (c0 * H) // chunk * chunk; hi = -(-(c1 * H) // chunk) * chunk
def timed(fn, reps=60):
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        scene.engine.clear_depth(); raster.render_occup()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return np.mean(ts[10:]), np.min(ts)
print('whole, fill      ', timed(lambda: raster.render_color(shader, fill_bg=bg)))
print('whole, nofill    ', timed(lambda: raster.render_color(shader)))
print('covered cols fill', timed(lambda: raster.render_color_range(shader, lo, hi - lo, fill_bg=bg)))
print('covered cols nofl', timed(lambda: raster.render_color_range(shader, lo, hi - lo)))
print('left bg fill     ', timed(lambda: raster.render_color_range(shader, 0, lo, fill_bg=bg)))
print('left bg nofill   ', timed(lambda: raster.render_color_range(shader, 0, lo)))
img = scene.img.to_torch()
def z():
    img.zero_()
print('torch zero_ image', timed(z))
