#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs other than the headline (bench.py = C2):

    python tools/bench_configs.py c1|c3|c4|c5 [--check] [--tiny-max N ...]
    torchrun --nproc-per-node G tools/bench_configs.py c4|c5      (view-parallel / sort-last)

Prints one JSON line per measurement (rank 0).  Timing: CUDA events, L2 flushed between iterations,
max over ranks.  --check compares ids/depth with the CPU oracle (c1, c3) or with the 1-GPU result (c5).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def timed(fn, iters, flush, world):
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('config', choices=['c1', 'c3', 'c4', 'c5'])
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--tiny-max', type=int, nargs='*', default=[None])
    ap.add_argument('--faces', type=int, default=None)
    ap.add_argument('--views', type=int, default=64)
    ap.add_argument('--graph', action='store_true', help='c1/c4: record the step into a CUDA graph and time replays')
    ap.add_argument('--p2p', action='store_true', help='c5: composite over NVLink peer memory inside the shading kernel')
    ap.add_argument('--root', action='store_true', help='c5: assemble the image on rank 0 only (strips stored over NVLink by the shading kernels)')
    ap.add_argument('--partitioned', action='store_true', help='c5: partitioned attributes + all-reduce composite')
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    import scenes
    import taichi_three_b200 as tina
    from taichi_three_b200 import multigpu as M

    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dev = torch.device('cuda', torch.cuda.current_device())
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    flush = torch.empty(256 * 2**20 // 4, dtype=torch.float32, device=dev)
    say = (lambda **kw: print(json.dumps(kw), flush=True)) if rank == 0 else (lambda **kw: None)

    if args.config == 'c1':
        W = H = 512
        obj = scenes.load_monkey()
        scene = tina.Scene((W, H), tonemap=False)
        mesh = tina.MeshModel(obj)
        scene.add_object(mesh)
        view, proj = scenes.default_camera()
        scene.engine.set_camera(view, proj)
        raster, shader = scene.triangle_raster, scene.shaders[id(scene.default_material)]
        raster.set_object(mesh)

        def step():
            scene.engine.clear_depth()
            raster.render_occup()
            raster.render_color(shader, fill_bg=np.zeros(3, np.float32))
        for tm in args.tiny_max:
            if tm is not None:
                raster.set_tuning(tiny_max=tm)
            step()
            run = tina.FrameGraph(step).replay if args.graph else step
            med, mn = timed(run, args.iters, flush, world)
            say(config='c1', graph=bool(args.graph), faces=968, res=[W, H], tiny_max=tm, ms=med, ms_min=mn, frames_per_s=1e3 / med, mtris_per_s=968 / med / 1e3)

    elif args.config == 'c3':
        W, H = 3840, 2160
        N = args.faces or 16 * 2**20
        view, proj = scenes.default_camera(W / H)
        tri = scenes.soup_torch(N, W, H, scenes.SOUP_S_C3, 20240601, dev)
        engine = tina.Engine((W, H))
        engine.set_camera(view, proj)
        raster = tina.TriangleRaster(engine, maxfaces=N)
        raster.set_face_verts(tri)
        lighting = tina.Lighting()
        lighting.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
        lighting.set_ambient_light([0.1, 0.1, 0.1])
        img = tina.Field(torch.zeros((W, H, 3), device=dev))
        shader = tina.Shader(img, lighting, tina.Diffuse())

        def step():
            engine.clear_depth()
            raster.render_occup()
            raster.render_color(shader, fill_bg=np.zeros(3, np.float32))
        ref = None
        for tm in args.tiny_max:
            for pre in (1, 0):
                if tm is not None:
                    raster.set_tuning(tiny_max=tm)
                raster.set_tuning(precheck=pre, profile=1)
                step()
                torch.cuda.synchronize()
                kt = raster.kernel_times()
                raster.set_tuning(profile=0)
                med, mn = timed(step, args.iters, flush, world)
                keys = engine.keys.clone()
                if ref is None:
                    ref = keys
                same = bool(torch.equal(keys, ref))
                alg = N * 36 + W * H * 20
                say(config='c3', faces=N, res=[W, H], tiny_max=tm, precheck=pre, ms=med, ms_min=mn, mtris_per_s=N / med / 1e3,
                    frames_per_s=1e3 / med, alg_gbs=alg / med / 1e6, kernel_ms=kt, same_bits_as_first=same,
                    depth_complexity=None)
        if args.check:
            from oracle import oracle as O
            t0 = time.time()
            occup, depth, tie, st = O.render_occup(tri.cpu().numpy(), (proj @ view).astype(np.float32), W, H)
            d, o = M.unpack_keys(ref.cpu().view(W, H))
            say(config='c3', check='oracle', depth_equal=bool(np.array_equal(d.numpy(), depth)),
                occup_equal=bool(np.array_equal(o.numpy(), occup)), covered_per_face=st['covered'] / N,
                depth_complexity=st['covered'] / (W * H), tie_pixels=int(tie.sum()), oracle_s=time.time() - t0)

    elif args.config == 'c4':
        W = H = 1024
        gltf = scenes.load_cornell()
        scene = tina.Scene((W, H), smoothing=True, texturing=True)
        gltf.extract(scene)
        cams = scenes.cornell_views(args.views)
        mine = M.view_partition(len(cams), rank, world)

        def step():
            for k in mine:
                scene.engine.set_camera(*cams[k])
                scene.render()
        step()
        run = tina.FrameGraph(step).replay if args.graph else step
        med, mn = timed(run, args.iters, flush, world)
        say(config='c4', graph=bool(args.graph), views=len(cams), gpus=world, res=[W, H], ms=med, ms_min=mn, views_per_s=len(cams) / med * 1e3,
            ms_per_view_per_gpu=med / max(1, len(mine)))

    elif args.config == 'c5':
        W, H = 7680, 4320
        N = args.faces or 128 * 2**20
        view, proj = scenes.default_camera(W / H)
        lo, hi = M.face_range(N, rank, world)
        # every rank generates the same global soup (replicated attributes: 4.8 GB at 128 M faces)
        tri = scenes.soup_torch(N, W, H, scenes.SOUP_S_C5, 20240602, dev)
        torch.cuda.empty_cache()
        engine = tina.Engine((W, H))
        engine.set_camera(view, proj)
        raster = tina.TriangleRaster(engine, maxfaces=N)
        lighting = tina.Lighting()
        lighting.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
        lighting.set_ambient_light([0.1, 0.1, 0.1])
        shared = M.SharedImage((W, H)) if (args.root and world > 1) else None
        img = tina.Field(shared.tensor if shared is not None else torch.zeros((W, H, 3), device=dev))
        shader = tina.Shader(img, lighting, tina.Diffuse())

        if args.p2p and world > 1:
            engine.open_peer_keys()

        def step():
            if args.partitioned:
                M.render_sort_last(engine, raster, tri[lo:hi], None, None, shader)
            else:
                M.render_sort_last_replicated(engine, raster, tri, None, None, shader, composite='p2p' if args.p2p else 'nccl',
                                              gather='root' if shared is not None else 'all')
        step()
        med, mn = timed(step, args.iters, flush, world)
        out = dict(config='c5', gather='root' if shared is not None else 'all', composite='p2p' if args.p2p else ('allreduce' if args.partitioned else 'reduce_scatter'), faces=N, gpus=world, res=[W, H], ms=med, ms_min=mn, mtris_per_s=N / med / 1e3, frames_per_s=1e3 / med)
        if args.check:
            # checksum of the composited keys and image: identical for every G
            k = engine.keys
            out['keys_checksum'] = int((k ^ (k >> 29)).sum().item())
            out['covered'] = int(((k & 0xffffffff) != 0).sum().item())
            if world == 1 and N <= 2**24:
                # C5 against the CPU oracle on a 2^24-face prefix of the same soup (same generator, same seed)
                from oracle import oracle as O
                t0 = time.time()
                occup, depth, tie, st = O.render_occup(tri.cpu().numpy(), (proj @ view).astype(np.float32), W, H)
                d, o = M.unpack_keys(engine.keys.cpu().view(W, H))
                out['oracle'] = dict(depth_equal=bool(np.array_equal(d.numpy(), depth)), occup_equal=bool(np.array_equal(o.numpy(), occup)),
                                     covered_per_face=st['covered'] / N, tie_pixels=int(tie.sum()), oracle_s=time.time() - t0)
            if world > 1:
                dist.barrier()
            out['image_sum'] = float(img.to_torch().double().sum().item())  # (root gather: every rank reads the root's memory)
        say(**out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
