#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 triangle-raster hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c2b|c1|c3|c4|c5] [--no-cpu] [--no-extra]

Default workload C2 (BASELINE.json configs[1]): one "step" = one frame of the hot path on one GPU,
    Engine.clear_depth -> TriangleRaster.render_occup -> TriangleRaster.render_color
(reference tina/core/engine.py:68-70, tina/core/triangle.py:89-153) on MeshGrid(1024) wave, 2,093,058 faces,
1920x1080, smooth normals, Classic material (Lambert + reflect-vector Phong).  Prints ONE JSON line (rank 0).

  value        Mtris/s with the meshes resident in HBM: K steps timed as one region (CUDA events on the launching stream,
               max over ranks) over 8 different scenes in rotation -- inputs larger than L2, nothing between steps;
               ms_per_step_flushed = round 1's protocol (one scene, 256 MiB flush write between steps, per-step events)
  e2e          the same metric through the public Python API with HOST buffers: pinned H2D of the frame's vertex
               grid, set_object, the step, pinned D2H of the image (3 frames in flight)
  roofline     dominant kernel (k_raster_quads): algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (port of the reference algorithm; the reference itself needs a Taichi 0.7 runtime
               that cannot be installed here) on this box's host cores
  extra        C2b (the grid wrapped in MeshNoCulling like the reference's own wave example) and the two partitioned
               configs of BASELINE.json at this N: C4 (cornell.gltf, 64 views at 1024^2,
               view-partitioned, each rank's views replayed as one CUDA graph; strong scaling) and C5 (134 M-face
               soup at 7680x4320, sort-last by face range + key composite over NVLink; strong scaling), with the
               key / image checksums that must not depend on N
N > 1 (torchrun): C2 = every rank renders its own frames (view-partitioned, no data-path collective) -> weak
scaling; --workload c4 / c5 print the partitioned configs as the headline line instead (strong scaling).
--impl reference times the CPU restatement (oracle/) with every host core and imports nothing of the product.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

K1_NCU = 'r2_v6_k_raster_quads_ncu.txt'


def ncu_traffic(kernel_file=K1_NCU):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` summary
    under profiles/ (tools/ncu_summary.py output).  None if the file is missing."""
    p = os.path.join(ROOT, 'profiles', kernel_file)
    try:
        tot, seen = 0.0, 0
        for line in open(p):
            f = line.split()
            if len(f) >= 3 and f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and seen < 2:
                scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[f[2]]
                tot += float(f[1]) * scale
                seen += 1
        return tot if seen == 2 else None
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


SOUP_S_C3, SOUP_S_C5 = 0.001374, 0.000487  # tests/scenes.py: calibrated with the oracle (covered samples per face)
WORKLOADS = {
    'c2': dict(kind='grid', n=1024, W=1920, H=1080, smoothing=True, material='classic', scaling='weak',
               desc='C2 MeshGrid(1024) wave t=0.25, 2093058 faces, 1920x1080, smooth normals, Classic (Lambert+Phong)'),
    'c2b': dict(kind='grid', n=1024, W=1920, H=1080, smoothing=True, material='classic', scaling='weak', nocull=True,
                desc='C2b MeshNoCulling(MeshGrid(1024)) wave t=0.25 (examples/meshgrid_wave.py:21), 4186116 faces, 1920x1080, '
                     'smooth normals, Classic (Lambert+Phong)'),
    'c1': dict(kind='monkey', W=512, H=512, smoothing=False, material='diffuse', scaling='weak',
               desc='C1 monkey.obj 968 faces, 512x512, flat, Diffuse'),
    'c3': dict(kind='soup', nfaces=16 * 2**20, W=3840, H=2160, smoothing=False, material='diffuse', s=0.00105, scaling='weak',
               desc='C3 random soup 16777216 faces, 3840x2160, depth complexity ~8, flat, Diffuse'),
    'c4': dict(kind='cornell', views=64, W=1024, H=1024, scaling='strong',
               desc='C4 cornell.gltf (3 objects, 34 faces, PBR + 512^2 texture), 64 views at 1024x1024, smoothing + texturing'),
    'c5': dict(kind='sortlast', nfaces=128 * 2**20, W=7680, H=4320, scaling='strong',
               desc='C5 random soup 134217728 faces, 7680x4320, depth complexity ~8, flat, Diffuse, sort-last by face range'),
}


ROTATED = ('c2', 'c2b')  # workloads whose headline is timed over ROTATE_SCENES scenes in rotation (inputs > L2)
ROTATE_SCENES = 8


def config_for(wl, world):
    """The `config` object of the JSON line: a pure function of (workload, N), identical in both arms."""
    w = WORKLOADS[wl]
    c = {'workload': w['desc'], 'res': [w['W'], w['H']],
         'l2': 'flushed between steps (256 MiB write, outside the per-step events)'}
    if wl in ROTATED:
        c['l2'] = (f'inputs larger than L2: consecutive steps render {ROTATE_SCENES} different scenes in rotation (wave phase t = 0.25 + '
                   '0.01 k, own vertex / normal / record / key / image buffers, ~100 MB each = ~6x the 126 MB L2 per cycle), no kernel '
                   'between steps; K steps timed as one region.  ms_per_step_flushed = the round-1 protocol (one scene, 256 MiB '
                   'write between steps, per-step events) beside it')
    if wl == 'c4':
        c.update(step='64 x (set_camera + Scene.render of 3 objects)', views=w['views'],
                 parallelism=f'views k = rank mod {world}, each rank replays its views as one CUDA graph')
    elif wl == 'c5':
        c.update(step='clear_depth + render_occup(own face range) + key composite + render_color(own strip) + image to rank 0',
                 faces=w['nfaces'], parallelism=f'sort-last: {world} contiguous face ranges, composite over NVLink peer memory')
    else:
        c.update(step='clear_depth + render_occup + render_color(fill_bg)',
                 parallelism=f'view-partitioned x{world} (independent frames, no data-path collective)' if world > 1 else 'single GPU')
    return c


def metric_for(wl):
    if wl == 'c4':
        return 'views/s (Scene.render, 64-view batch at 1024x1024)', 'views/s'
    if wl in ('c2', 'c2b'):
        return 'Mtris/s (render_occup+render_color, 1080p)', 'Mtris/s'
    return 'Mtris/s (render_occup+render_color)', 'Mtris/s'


def alg_bytes(wl, nfaces):
    """B_alg = N*(36 + 36*smoothing + 24*texturing) + W*H*20 per frame;
    the rasteriser alone: N*36 (positions read once) + W*H*8 (occup + depth written once)."""
    w = WORKLOADS[wl]
    px = w['W'] * w['H']
    frame = nfaces * (36 + (36 if w.get('smoothing') else 0)) + px * 20
    k1 = nfaces * 36 + px * 8
    return frame, k1


class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '20'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'], samples=0)
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(',') for r in open(self.f.name).read().strip().splitlines() if r.count(',') >= 6]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.strip().lower() == 'active':
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# -------------------------------------------------------------------------------------------------
# CPU side: product-free (oracle/ only).  Used by --impl reference and by the cpu_baseline leg.
# -------------------------------------------------------------------------------------------------
def cpu_inputs(wl):
    from oracle import oracle as O
    from oracle import workloads as R
    w = WORKLOADS[wl]
    view, proj = R.default_camera(w['W'] / w['H'])
    if w['kind'] == 'grid':
        pos = R.wave_grid_pos(w['n'])
        verts, norms = O.grid_faces(pos), O.grid_faces(O.grid_normals(pos))
        if w.get('nocull'):
            verts, norms, _ = O.no_culling(verts, norms)
    elif w['kind'] == 'monkey':
        verts, norms = R.monkey_faces(os.path.join(ROOT, 'tests', 'assets', 'monkey.obj')), None
    elif w['kind'] == 'soup':
        verts, norms = R.soup(w['nfaces'], w['W'], w['H'], s=w['s']), None
    else:
        raise SystemExit(f'--impl reference / cpu_baseline is provided for c1, c2, c2b, c3 (not {wl})')
    return dict(verts=verts, norms=norms, view=view, proj=proj, nfaces=len(verts))


def cpu_frame_time(wl, inp, seconds=10.0, min_frames=2, max_frames=50):
    """Time clear + render_occup + render_color of the CPU oracle (parallel mode = the reference's own parallel
    structure: faces across threads with an atomic min, pixels across threads) on every host core.
    -> (median seconds per frame, frames, threads)."""
    from oracle import oracle as O
    from oracle import materials as OM
    w = WORKLOADS[wl]
    W, H = w['W'], w['H']
    threads = O.set_num_threads()  # all cores, whatever OMP_NUM_THREADS says (torchrun exports 1)
    flags = O.CULLING | O.CLIPPING | (O.SMOOTHING if w['smoothing'] else 0)
    material = OM.stock_classic() if w['material'] == 'classic' else OM.stock_diffuse()
    lighting = OM.default_lighting()
    W2V64 = inp['proj'] @ inp['view']
    W2V, V2W = W2V64.astype(np.float32), np.linalg.inv(W2V64).astype(np.float32)
    image = np.zeros((W, H, 3), np.float32)
    times = []
    t_end = time.perf_counter() + seconds
    while len(times) < min_frames or (time.perf_counter() < t_end and len(times) < max_frames):
        t0 = time.perf_counter()
        image[...] = 0.0
        occup, depth, _, _ = O.render_occup(inp['verts'], W2V, W, H, flags, parallel=True)  # includes the depth clear
        O.render_color(inp['verts'], inp['norms'], None, occup, W2V, V2W, W, H, flags, material, lighting, image, parallel=True)
        times.append(time.perf_counter() - t0)
    return float(np.median(times)), len(times), threads


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference (pure Python + Taichi 0.7
    JIT) cannot run here or on the GPU box (no taichi wheel, no network), so this times the oracle port of its
    algorithm with all host threads, one full frame per step.  Nothing of the product is imported."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    wl = args.workload
    world = int(os.environ.get('WORLD_SIZE', '1'))
    inp = cpu_inputs(wl)
    nfaces = inp['nfaces']
    sec, frames, threads = cpu_frame_time(wl, inp, seconds=1e9, min_frames=args.warmup + args.steps,
                                          max_frames=args.warmup + args.steps)
    assert threads > 1 or (os.cpu_count() or 1) == 1, 'the CPU arm must use every host core'
    value = nfaces / sec / 1e6
    metric, unit = metric_for(wl)
    line = {
        'impl': 'reference', 'metric': metric, 'value': value, 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': WORKLOADS[wl]['scaling'],
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'frames_per_s': 1.0 / sec,
        'config': config_for(wl, world),
        'cpu_baseline': {'value': value, 'unit': unit, 'cores': threads, 'kind': 'port',
                         'sample': f'{frames} full frames (clear+render_occup+render_color) on rank 0, median'},
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'product_modules_loaded': sorted(m for m in sys.modules if m.split('.')[0] in ('taichi_three_b200', 'tina')),
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# GPU side
# -------------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(index):
    """Pin this rank to the CPUs next to its GPU before any pinned host buffer is allocated (NUMA-local staging
    buffers: eight ranks otherwise share one socket's memory controllers).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [i * 64 + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def make_env():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    ncpus = None if os.environ.get('TINA_BENCH_NOBIND') else bind_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    flush = torch.empty(256 * 2**20 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    return dict(rank=rank, local_rank=local_rank, world=world, dev=dev, barrier=barrier, allmax=allmax, flush=flush,
                ncpus=ncpus)


def timed_steps(env, step, K, warmup, sweep=None):
    """W warm-up steps, then K steps with per-step CUDA events, L2 flushed between steps; -> (ms per step = sum of the
    per-step times / K, max over ranks; per-step array of this rank; wall seconds of the timed region).
    sweep: a second > L2 buffer that is READ after the flush write, so that the step starts on an L2 full of clean
    foreign lines instead of dirty ones (informational figure; the headline uses the plain write flush)."""
    import torch
    fill = env['flush']

    class _F:
        @staticmethod
        def fill_(v):
            fill.fill_(v)
            if sweep is not None:
                sweep.amax()
    flush = _F
    for _ in range(max(3, warmup)):
        step()
        flush.fill_(1.0)
    env['barrier']()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    env['barrier']()
    wall0 = time.perf_counter()
    for a, b in evs:
        flush.fill_(1.0)
        a.record()
        step()
        b.record()
    env['barrier']()
    wall = time.perf_counter() - wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    return env['allmax'](float(step_ms.sum())) / K, step_ms, wall


def c4_setup(env):
    import scenes
    import taichi_three_b200 as tina
    from taichi_three_b200 import multigpu as M
    w = WORKLOADS['c4']
    gltf = scenes.load_cornell()
    scene = tina.Scene((w['W'], w['H']), smoothing=True, texturing=True)
    gltf.extract(scene)
    cams = scenes.cornell_views(w['views'])
    mine = M.view_partition(len(cams), env['rank'], env['world'])

    def step():
        for k in mine:
            scene.engine.set_camera(*cams[k])
            scene.render()
    step()
    graph = tina.FrameGraph(step)
    return scene, graph, mine


def c4_measure(env, K, warmup):
    """C4: the fixed 64-view batch, views dealt to ranks, each rank's views replayed as one CUDA graph."""
    import torch
    w = WORKLOADS['c4']
    scene, graph, mine = c4_setup(env)
    ms, step_ms, _ = timed_steps(env, graph.replay, K, warmup)
    torch.cuda.synchronize()
    # checksum of the last view this rank rendered, summed over ranks: must not depend on N for N | 64 ... it does
    # depend on WHICH views are last, so take the batch's view 63 (rank (63 mod N) renders it last)
    img_sum = float(scene.img.to_torch().double().sum().item()) if (w['views'] - 1) in mine else 0.0
    img_sum = env['allmax'](img_sum)
    return dict(views_per_s=w['views'] / (ms * 1e-3), ms_per_batch=ms, views=w['views'], views_per_rank=len(mine),
                ms_per_view_per_gpu=ms / max(1, len(mine)), last_view_image_sum=img_sum,
                step_ms_min_median_max=[float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())])


def c5_measure(env, K, warmup, nfaces=None):
    """C5: sort-last.  Rank r rasterises faces [r N/G, (r+1) N/G) with global ids into its full-resolution key buffer;
    rank r's shading kernel takes the keys of screen strip r as the MIN over all ranks' buffers read over NVLink
    (CUDA IPC peer memory) and stores its strip into the image that lives on rank 0."""
    import torch
    import scenes
    import taichi_three_b200 as tina
    from taichi_three_b200 import multigpu as M
    w = WORKLOADS['c5']
    W, H = w['W'], w['H']
    N = nfaces or w['nfaces']
    world, dev = env['world'], env['dev']
    view, proj = scenes.default_camera(W / H)
    tri = scenes.soup_torch(N, W, H, SOUP_S_C5, 20240602, dev)  # replicated attributes (4.8 GB), same stream on every rank
    torch.cuda.empty_cache()
    engine = tina.Engine((W, H))
    engine.set_camera(view, proj)
    raster = tina.TriangleRaster(engine, maxfaces=N)
    lighting = tina.Lighting()
    lighting.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
    lighting.set_ambient_light([0.1, 0.1, 0.1])
    shared = M.SharedImage((W, H)) if world > 1 else None
    img = tina.Field(shared.tensor if shared is not None else torch.zeros((W, H, 3), device=dev))
    shader = tina.Shader(img, lighting, tina.Diffuse())
    if world > 1:
        engine.open_peer_keys()

    def step():
        M.render_sort_last_replicated(engine, raster, tri, None, None, shader, composite='p2p', gather='root')
    ms, step_ms, _ = timed_steps(env, step, K, warmup)
    env['barrier']()
    k = engine.keys
    out = dict(ms_per_frame=ms, mtris_per_s=N / ms / 1e3, frames_per_s=1e3 / ms, faces=N,
               step_ms_min_median_max=[float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())],
               composite='keys MIN over peer key buffers inside k_render_color (NVLink loads), image strips stored to rank 0' if world > 1 else 'single GPU',
               # every rank reads its strip of the other ranks' keys; every rank but the root stores its image strip to the root
               nvlink_bytes_per_frame=0)
    # checksums that must not depend on the number of GPUs: the composited image (complete on rank 0) and, per strip
    # owner, the composited keys of its strip
    npix = W * H
    lo, hi, share = M.sort_last_strip(npix, env['rank'], world) if world > 1 and npix % (256 * world) == 0 else (0, npix, 1.0)
    if world > 1:
        r_lo, r_hi, _ = M.sort_last_strip(npix, 0, world)
        out['root_strip_share'] = (r_hi - r_lo) / npix
        # every rank reads its strip of the other ranks' keys; every rank but the root stores its image strip to the root
        out['nvlink_bytes_per_frame'] = (world - 1) * npix * 8 + (npix - (r_hi - r_lo)) * 12
    ks = k.view(-1)[lo:hi]
    ksum = torch.stack([(ks ^ (ks >> 29)).sum(), ((ks & 0xffffffff) != 0).sum()])
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ksum, op=dist.ReduceOp.SUM)
    out['keys_checksum'] = int(ksum[0].item())
    out['covered_pixels'] = int(ksum[1].item())
    out['image_sum'] = env['allmax'](float(img.to_torch().double().sum().item()) if env['rank'] == 0 else -1e300)
    env['barrier']()
    if shared is not None:
        engine.close_peer_keys()
        shared.close()
    del tri, raster, engine
    torch.cuda.empty_cache()
    return out


def c2b_measure(env, K, warmup):
    """C2b: the reference's own wave example wraps the grid in MeshNoCulling (examples/meshgrid_wave.py:21): twice the
    faces, the generic indexed rasteriser (k_raster_indexed) instead of the quad kernel.  Same step, same timing."""
    import scenes
    import taichi_three_b200 as tina
    w = WORKLOADS['c2b']
    W, H = w['W'], w['H']
    nfaces = 4 * (w['n'] - 1) ** 2
    sc = tina.Scene((W, H), smoothing=True, maxfaces=nfaces, tonemap=False)
    grid = tina.MeshGrid(w['n'])
    grid.pos.from_numpy(scenes.wave_grid_pos(w['n'], t=0.25 + 0.01 * env['rank']))
    ms, mt = tina.MeshNoCulling(grid), tina.Classic()
    sc.add_object(ms, mt)
    sc.engine.set_camera(*tina.orbit_camera(aspect=W / H))
    raster, shader = sc.triangle_raster, sc.shaders[id(mt)]
    raster.set_object(ms)
    bg = np.zeros(3, np.float32)

    def step():
        sc.engine.clear_depth()
        raster.render_occup()
        raster.render_color(shader, fill_bg=bg)
    ms_step, step_ms, _ = timed_steps(env, step, K, warmup)
    frame_bytes, _ = alg_bytes('c2b', nfaces)
    return dict(ms_per_step=ms_step, mtris_per_s=env['world'] * nfaces / ms_step / 1e3, faces=nfaces,
                frame_roofline_frac=frame_bytes / (ms_step * 1e-3) / 1e9 / load_peaks()[0],
                step_ms_min_median_max=[float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())])


def run_ours(args):
    import torch
    import torch.distributed as dist
    import scenes
    import taichi_three_b200 as tina
    from taichi_three_b200 import _lib as _tl

    env = make_env()
    rank, world, dev, barrier = env['rank'], env['world'], env['dev'], env['barrier']
    wl = args.workload
    w = WORKLOADS[wl]
    W, H = w['W'], w['H']
    K = args.steps
    metric, unit = metric_for(wl)
    peak, peak_src = load_peaks()
    launches0 = int(_tl.lib().tina_launch_count())

    if wl in ('c4', 'c5'):
        sampler = ClockSampler(env['local_rank']) if rank == 0 else None
        if wl == 'c4':
            Kc = min(K, 50)
            res = c4_measure(env, Kc, args.warmup)
            value, ms = res['views_per_s'], res['ms_per_batch']
            frame_bytes = 34 * (36 + 36 + 24) * w['views'] + W * H * 20 * w['views']
        else:
            Kc = min(K, 20)
            res = c5_measure(env, Kc, args.warmup)
            value, ms = res['mtris_per_s'], res['ms_per_frame']
            frame_bytes = w['nfaces'] * 36 + W * H * 20
        clocks = sampler.stop() if sampler is not None else None
        if rank == 0:
            line = {'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': Kc, 'warmup': max(3, args.warmup),
                    'ms_per_step': ms, 'higher_is_better': True, 'scaling': w['scaling'], 'vs_baseline': None, 'dtype': 'f32',
                    'data': 'synthetic', 'config': config_for(wl, world), 'frame_alg_bytes': frame_bytes,
                    'frame_hbm_gbs': frame_bytes / (ms * 1e-3) / 1e9, 'frame_roofline_frac': frame_bytes / (ms * 1e-3) / 1e9 / (peak * world),
                    'detail': res, 'roofline': None, 'cpu_baseline': None, 'e2e': None,
                    'gpu_launches': (int(_tl.lib().tina_launch_count()) - launches0) * world, 'clocks': clocks}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- C2 / C1 / C3: one frame per step on every rank ----
    if w['kind'] == 'grid':
        inputs = dict(pos=scenes.wave_grid_pos(w['n'], t=0.25 + 0.01 * rank), nfaces=(4 if w.get('nocull') else 2) * (w['n'] - 1) ** 2)
    elif w['kind'] == 'monkey':
        obj = scenes.load_monkey()
        inputs = dict(obj=obj, nfaces=len(obj['f']))
    else:
        inputs = dict(tri=scenes.soup(w['nfaces'], W, H, s=w['s'], seed=20240601 + rank), nfaces=w['nfaces'])
    view, proj = tina.orbit_camera(aspect=W / H)
    nfaces = inputs['nfaces']

    def make_scene():
        sc = tina.Scene((W, H), smoothing=w['smoothing'], maxfaces=max(nfaces, 2**20), tonemap=False)
        mt = tina.Classic() if w['material'] == 'classic' else tina.Diffuse()
        if w['kind'] == 'grid':
            ms = tina.MeshGrid(w['n'])
            ms.pos.from_numpy(inputs['pos'])
            if w.get('nocull'):
                ms = tina.MeshNoCulling(ms)
        elif w['kind'] == 'monkey':
            ms = tina.MeshModel(inputs['obj'])
        else:
            ms = tina.SimpleMesh(maxfaces=nfaces)
            ms.set_face_verts(inputs['tri'])
        sc.add_object(ms, mt)
        sc.engine.set_camera(view, proj)
        return sc, ms, sc.shaders[id(mt)]

    scene, mesh, shader = make_scene()
    engine, raster = scene.engine, scene.triangle_raster
    bg = np.zeros(3, np.float32)
    raster.set_object(mesh)  # the mesh is now resident in HBM

    def step():
        engine.clear_depth()
        raster.render_occup()
        raster.render_color(shader, fill_bg=bg)

    flush = env['flush']
    sampler = ClockSampler(env['local_rank']) if rank == 0 else None
    l0 = int(_tl.lib().tina_launch_count())
    flushed_ms, step_ms, wall = timed_steps(env, step, K, args.warmup)
    # launches inside the K timed steps: (count after - count before the warm-up) scaled to the timed share
    per_step_launches = (int(_tl.lib().tina_launch_count()) - l0) / (K + max(3, args.warmup))
    ms_per_step = flushed_ms
    if wl in ROTATED:
        # ---- headline: K steps as ONE timed region over ROTATE_SCENES different scenes (inputs larger than L2, nothing
        # between the steps): the steady state of rendering an animated mesh -- every frame's vertices, normals, keys and
        # image are other memory than the previous frames', and the launches of consecutive frames follow each other on
        # the stream as they do in Scene.render loops ----
        def make_step(k):
            ik = dict(inputs, pos=scenes.wave_grid_pos(w['n'], t=0.25 + 0.01 * (rank + k)))
            sc = tina.Scene((W, H), smoothing=w['smoothing'], maxfaces=max(nfaces, 2**20), tonemap=False)
            mt = tina.Classic() if w['material'] == 'classic' else tina.Diffuse()
            ms = tina.MeshGrid(w['n'])
            ms.pos.from_numpy(ik['pos'])
            if w.get('nocull'):
                ms = tina.MeshNoCulling(ms)
            sc.add_object(ms, mt)
            sc.engine.set_camera(view, proj)
            rs, sh = sc.triangle_raster, sc.shaders[id(mt)]
            rs.set_object(ms)

            def st():
                sc.engine.clear_depth()
                rs.render_occup()
                rs.render_color(sh, fill_bg=bg)
            return sc, st
        pool = [make_step(k) for k in range(ROTATE_SCENES)]
        # (every scene is warmed at least 10 times: its rasteriser stops launching the idle tile-path kernel after 8 frames
        # that queued no large face, as it has in any render loop)
        for i in range(max(10, args.warmup) * ROTATE_SCENES):
            pool[i % ROTATE_SCENES][1]()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l1 = int(_tl.lib().tina_launch_count())
        a.record()
        for i in range(K):
            pool[i % ROTATE_SCENES][1]()
        b.record()
        barrier()
        per_step_launches = (int(_tl.lib().tina_launch_count()) - l1) / K
        ms_per_step = env['allmax'](a.elapsed_time(b)) / K
        del pool
        torch.cuda.empty_cache()
    # keep the GPU under the same load until the clock sampler has a few samples
    clocks = None
    if sampler is not None:
        t_end = time.perf_counter() + max(0.0, 0.6 - wall)
        while time.perf_counter() < t_end:
            flush.fill_(1.0)
            step()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks['sampled'] = 'timed region + same workload looped to >= 0.6 s'
    value = world * nfaces / (ms_per_step * 1e-3) / 1e6

    # ---- the same steps on a cold but CLEAN L2 (flush write followed by a 256 MiB read sweep), for information:
    # after the write flush the step also pays the write-back of the flush buffer's dirty lines it evicts ----
    sweep = torch.ones_like(flush)
    clean_ms, _, _ = timed_steps(env, step, min(K, 100), 3, sweep=sweep)
    del sweep

    # ---- back-to-back (no flush) for information ----
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        step()
    b.record()
    torch.cuda.synchronize()
    b2b_ms = a.elapsed_time(b) / K

    # ---- per-kernel CUDA events (dominant kernel for the roofline), same steps with L2 flush ----
    raster.set_tuning(profile=1)
    kt = {}
    for _ in range(min(K, 50)):
        flush.fill_(1.0)
        step()
        torch.cuda.synchronize()
        for name, ms in raster.kernel_times().items():
            kt.setdefault(name, []).append(ms)
    raster.set_tuning(profile=0)
    kmean = {k: float(np.mean(v)) for k, v in kt.items() if np.mean(v) >= 0}

    # ---- e2e: host buffers, pinned H2D + set_object + step + pinned D2H, public API ----
    # Every step uploads that step's vertex data from pinned host memory and reads that step's image back into pinned
    # host memory.  Up to three frames are in flight on three CUDA streams (three Scene instances), so the PCIe copies
    # of one frame overlap the kernels and the opposite-direction copies of the others; `serial` = one frame at a time.
    def make_lane():
        sc, ms, sh = make_scene()
        if w['kind'] == 'grid':
            src, dst = torch.as_tensor(inputs['pos']).pin_memory(), (ms.mesh if w.get('nocull') else ms).pos.to_torch()
        elif w['kind'] == 'soup':
            src, dst = torch.as_tensor(inputs['tri']).pin_memory(), ms.verts.to_torch()
        else:
            src, dst = torch.as_tensor(inputs['obj']['v']).pin_memory(), ms.verts
        img_d = sc.image.to_torch()
        return dict(scene=sc, mesh=ms, raster=sc.triangle_raster, shader=sh, src=src, dst=dst, img_d=img_d,
                    img_h=torch.empty(img_d.shape, dtype=torch.float32).pin_memory(), stream=torch.cuda.Stream(device=dev),
                    done=torch.cuda.Event())

    lanes = [make_lane() for _ in range(3)]
    h2d, d2h = lanes[0]['src'].numel() * 4, lanes[0]['img_h'].numel() * 4

    def e2e_frame(L):
        with torch.cuda.stream(L['stream']):
            L['dst'].copy_(L['src'], non_blocking=True)
            L['raster'].set_object(L['mesh'])
            L['scene'].engine.clear_depth()
            L['raster'].render_occup()
            L['raster'].render_color(L['shader'], fill_bg=bg)
            L['img_h'].copy_(L['img_d'], non_blocking=True)
            L['done'].record()

    def e2e_run(n, inflight):
        barrier()
        t0 = time.perf_counter()
        for i in range(n):
            L = lanes[i % inflight]
            L['done'].synchronize()  # the host now holds this lane's previous image; its buffers are free again
            e2e_frame(L)
        for L in lanes:
            L['done'].synchronize()
        barrier()
        return env['allmax']((time.perf_counter() - t0) / n)

    Ke = min(K, 100)
    e2e_run(6, 3)
    e2e_s = {n: e2e_run(Ke, n) for n in (1, 2, 3)}
    best = min(e2e_s, key=e2e_s.get)
    e2e_value = world * nfaces / e2e_s[best] / 1e6
    checksum = float(lanes[0]['img_h'].double().sum().item())
    assert abs(checksum - float(lanes[1]['img_h'].double().sum().item())) < 1e-6 * max(1.0, abs(checksum))
    # the PCIe ceiling of this box for the image read-back alone (every rank at once): the floor of e2e
    L0 = lanes[0]
    barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        L0['img_h'].copy_(L0['img_d'], non_blocking=True)
    torch.cuda.synchronize()
    d2h_s = env['allmax']((time.perf_counter() - t0) / 20)

    extra = None
    if wl == 'c2' and not args.no_extra:
        del lanes
        torch.cuda.empty_cache()
        extra = {}
        try:
            extra['c2b'] = c2b_measure(env, 50, 5)
        except Exception as ex:
            extra['c2b'] = {'error': repr(ex)[:300]}
        torch.cuda.empty_cache()
        try:
            extra['c4'] = c4_measure(env, 10, 3)
        except Exception as ex:  # the headline line must survive a failure of the side measurements
            extra['c4'] = {'error': repr(ex)[:300]}
        try:
            extra['c5'] = c5_measure(env, 5, 3)
        except Exception as ex:
            extra['c5'] = {'error': repr(ex)[:300]}

    my_launches = per_step_launches * K
    if rank == 0:
        frame_bytes, k1_bytes = alg_bytes(wl, nfaces)
        k1_ms = kmean.get('raster_faces', float('nan'))
        achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            sec, frames, threads = cpu_frame_time(wl, cpu_inputs(wl), seconds=10.0)
            cpu = {'value': nfaces / sec / 1e6, 'unit': 'Mtris/s', 'cores': threads, 'kind': 'port',
                   'sample': f'{frames} full frames of the same workload (clear+render_occup+render_color), median, '
                             f'{sec * 1e3:.1f} ms/frame'}
        line = {
            'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': K, 'warmup': max(3, args.warmup),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': w['scaling'], 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config_for(wl, world),
            'faces_per_step_per_gpu': nfaces,
            'frames_per_s': world / (ms_per_step * 1e-3),
            'frame_alg_bytes': frame_bytes,
            'frame_hbm_gbs': frame_bytes / (ms_per_step * 1e-3) / 1e9,
            'frame_roofline_frac': frame_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
            'ms_per_step_flushed': flushed_ms,
            'frame_roofline_frac_flushed': frame_bytes / (flushed_ms * 1e-3) / 1e9 / peak,
            'ms_per_step_back_to_back_no_flush': b2b_ms,
            'ms_per_step_clean_cold_l2': clean_ms,
            'flushed_step_ms_min_median_max': [float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())],
            'kernel_ms': kmean,
            'kernel_ms_mode': 'one CUDA-event pair per kernel, programmatic dependent launch off (the events would break the '
                              'launch pairing), adaptive tile-path skipping as in the timed steps; their sum exceeds ms_per_step by the '
                              'launch gaps PDL hides',
            'roofline': {'bound': 'hbm', 'kernel': ('k_raster_indexed' if w.get('nocull') else 'k_raster_quads') if w['kind'] == 'grid' else ('k_raster_indexed' if w['kind'] == 'monkey' else 'k_raster_faces'), 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': ncu_traffic() if wl == 'c2' else None,
                         'traffic_source': f'profiles/{K1_NCU} (ncu --set full, cold cache, dram read+write per launch)',
                         'note': 'latency / occupancy-bound kernel: ncu smsp__issue_active 67 %, long-scoreboard 2.8 cycles per issue, DRAM 16 % of peak (same file)',
                         'alg_bytes': k1_bytes, 'peak_source': peak_src},
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': unit, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_s[best] * 1e3, 'steps': Ke, 'image_checksum': checksum, 'frames_in_flight': best,
                    'ms_per_step_by_frames_in_flight': {str(n): e2e_s[n] * 1e3 for n in e2e_s},
                    'd2h_only_ms': d2h_s * 1e3, 'd2h_only_gbs': d2h / d2h_s / 1e9,
                    'note': 'floor = the f32 image over PCIe (d2h_only_ms, every rank copying at once)',
                    'cpus_bound_per_rank': env['ncpus']},
            # counted by the library (tina_launch_count) on rank 0, x ranks: k_frame_prologue (vertex records + the deferred
            # key clear), k_raster_quads, [k_large_path unless the adaptive tile path is skipping it], k_render_color per step
            'gpu_launches': int(round(my_launches)) * world,
            'clocks': clocks,
            'extra': extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extra', action='store_true', help='c2: skip the C4 / C5 side measurements')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)  # product-free: builds and loads oracle/ only
        return
    import __graft_entry__ as g
    g.build()
    run_ours(args)


if __name__ == '__main__':
    main()
