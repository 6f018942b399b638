#!/usr/bin/env python
"""Where a sort-last C5 frame goes on N GPUs (torchrun): raster / fence / composite + shade + image store / fence, per rank.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/c5_breakdown.py [root_share ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, torch.distributed as dist
import scenes, taichi_three_b200 as tina
from taichi_three_b200 import multigpu as M
import bench
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
dev = torch.device('cuda', torch.cuda.current_device())
dist.init_process_group('nccl', device_id=dev)
W, H, N = 7680, 4320, 128 * 2**20
view, proj = scenes.default_camera(W / H)
tri = scenes.soup_torch(N, W, H, bench.SOUP_S_C5, 20240602, dev)
engine = tina.Engine((W, H)); engine.set_camera(view, proj)
raster = tina.TriangleRaster(engine, maxfaces=N)
L = tina.Lighting(); L.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9]); L.set_ambient_light([0.1, 0.1, 0.1])
shared = M.SharedImage((W, H))
shader = tina.Shader(tina.Field(shared.tensor), L, tina.Diffuse())
engine.open_peer_keys()
lo, hi = M.face_range(N, rank, world)
npix = W * H
tok = torch.zeros(1, dtype=torch.int32, device=dev)
flush = torch.empty(64 * 2**20, device=dev)
def frame(share, ev=None):
    p_lo, p_hi, _ = M.sort_last_strip(npix, rank, world, root_share=share)
    rec = (lambda i: ev[i].record()) if ev else (lambda i: None)
    rec(0)
    engine.clear_depth(); engine.set_face_base(lo); raster.set_face_verts(tri[lo:hi]); raster.render_occup(); raster.set_face_verts(tri)
    rec(1)
    dist.all_reduce(tok)
    rec(2)
    raster.render_color_composite(shader, p_lo, p_hi - p_lo, face_base=0, fill_bg=0.0)
    rec(3)
    dist.all_reduce(tok)
    rec(4)
for share in [float(a) for a in sys.argv[1:]] or [0.0, 1.0]:
    for _ in range(3): frame(share); flush.fill_(1.0)
    acc = np.zeros(5)
    K = 6
    for _ in range(K):
        flush.fill_(1.0)
        dist.barrier(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        frame(share, ev); torch.cuda.synchronize()
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(4)] + [ev[0].elapsed_time(ev[4])])
    t = torch.tensor(acc / K, device=dev); allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        print(f'root_share {share}: per rank [clear+raster, fence, composite+shade+store, fence, total] ms')
        for r, a in enumerate(allt): print('  rank', r, np.round(a.cpu().numpy(), 3))
dist.barrier(); engine.close_peer_keys(); shared.close(); dist.destroy_process_group()
