#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 triangle-raster hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c1]

One "step" = one frame of the hot path on one GPU:
    Engine.clear_depth -> TriangleRaster.render_occup -> TriangleRaster.render_color
(reference tina/core/engine.py:68-70, tina/core/triangle.py:89-153) on BASELINE.json configs[1]
("C2": MeshGrid(1024) wave, 2,093,058 faces, 1920x1080, smooth normals, Classic material =
Lambert + reflect-vector Phong).  Prints ONE JSON line (rank 0).

  value        Mtris/s with the expanded face arrays already resident in HBM, per-step CUDA
               events on the launching stream, L2 flushed between steps, max over ranks
  e2e          same metric through the public Python API with HOST buffers: pinned H2D of the
               frame's vertex grid, set_object, the step, pinned D2H of the image
  roofline     dominant kernel (k_raster_faces): algorithmic bytes / CUDA-event duration
               against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the CPU oracle (port of the reference algorithm; the reference itself needs a
               Taichi 0.7 runtime that cannot be installed here) on this box's host cores
N > 1 (torchrun): every rank renders its own frames of the same workload (view-partitioned,
no data-path collective) -> weak scaling; value = all ranks' triangles / max-over-ranks time.
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


def ncu_traffic(kernel_file='r1_v10_k_raster_faces_ncu.txt'):
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


WORKLOADS = {
    # name: (grid n or None, faces, W, H, smoothing, description)
    'c2': dict(kind='grid', n=1024, W=1920, H=1080, smoothing=True, material='classic',
               desc='C2 MeshGrid(1024) wave t=0.25, 2093058 faces, 1920x1080, smooth normals, Classic (Lambert+Phong)'),
    'c1': dict(kind='monkey', W=512, H=512, smoothing=False, material='diffuse',
               desc='C1 monkey.obj 968 faces, 512x512, flat, Diffuse'),
    'c3': dict(kind='soup', nfaces=16 * 2**20, W=3840, H=2160, smoothing=False, material='diffuse', s=0.00105,
               desc='C3 random soup 16777216 faces, 3840x2160, depth complexity ~8, flat, Diffuse'),
}


def make_inputs(wl, rank=0):
    """-> dict(verts, norms, pos(optional grid), view, proj) as numpy (host) inputs."""
    import scenes
    import taichi_three_b200 as tina
    w = WORKLOADS[wl]
    view, proj = tina.orbit_camera(aspect=w['W'] / w['H'])
    out = dict(view=view, proj=proj)
    if w['kind'] == 'grid':
        out['pos'] = scenes.wave_grid_pos(w['n'], t=0.25 + 0.01 * rank)
        out['nfaces'] = 2 * (w['n'] - 1) ** 2
    elif w['kind'] == 'monkey':
        out['obj'] = scenes.load_monkey()
        out['nfaces'] = len(out['obj']['f'])
    elif w['kind'] == 'soup':
        out['tri'] = scenes.soup(w['nfaces'], w['W'], w['H'], s=w['s'], seed=20240601 + rank)
        out['nfaces'] = w['nfaces']
    return out


def alg_bytes(wl, nfaces):
    """B_alg = N*(36 + 36*smoothing + 24*texturing) + W*H*20 per frame;
    k_raster_faces alone: N*36 (positions read once) + W*H*8 (occup + depth written once)."""
    w = WORKLOADS[wl]
    px = w['W'] * w['H']
    frame = nfaces * (36 + (36 if w['smoothing'] else 0)) + px * 20
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
def cpu_frame_time(wl, inputs, seconds=10.0, min_frames=2, max_frames=50):
    """Time clear + render_occup + render_color of the CPU oracle (parallel mode = the reference's
    own parallel structure) on the host cores.  -> (median seconds per frame, frames, threads)."""
    from oracle import oracle as O
    import taichi_three_b200 as tina
    w = WORKLOADS[wl]
    W, H = w['W'], w['H']
    flags = O.CULLING | O.CLIPPING | (O.SMOOTHING if w['smoothing'] else 0)
    if w['kind'] == 'grid':
        verts, norms = O.grid_faces(inputs['pos']), O.grid_faces(O.grid_normals(inputs['pos']))
    elif w['kind'] == 'monkey':
        verts, norms, _ = O.indexed(inputs['obj'])
        norms = None
    else:
        verts, norms = inputs['tri'], None
    material = tina.Classic() if w['material'] == 'classic' else tina.Diffuse()
    lighting = tina.Lighting()
    lighting.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
    lighting.set_ambient_light([0.1, 0.1, 0.1])
    W2V64 = inputs['proj'] @ inputs['view']
    W2V, V2W = W2V64.astype(np.float32), np.linalg.inv(W2V64).astype(np.float32)
    image = np.zeros((W, H, 3), np.float32)
    times = []
    t_end = time.perf_counter() + seconds
    while len(times) < min_frames or (time.perf_counter() < t_end and len(times) < max_frames):
        t0 = time.perf_counter()
        image[...] = 0.0
        occup, depth, _, _ = O.render_occup(verts, W2V, W, H, flags, parallel=True)  # includes the depth clear
        O.render_color(verts, norms, None, occup, W2V, V2W, W, H, flags, material, lighting, image, parallel=True)
        times.append(time.perf_counter() - t0)
    return float(np.median(times)), len(times), O.num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference (pure Python +
    Taichi 0.7 JIT) cannot run here or on the GPU box (no taichi wheel, no network), so this times
    the oracle port of its algorithm with all host threads, one full frame per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = args.workload
    inputs = make_inputs(wl)
    nfaces = inputs['nfaces']
    # warmup + steps frames, bounded to a few minutes
    sec, frames, threads = cpu_frame_time(wl, inputs, seconds=1e9, min_frames=args.warmup + args.steps,
                                          max_frames=args.warmup + args.steps)
    value = nfaces / sec / 1e6
    line = {
        'impl': 'reference', 'metric': 'Mtris/s (render_occup+render_color, 1080p)' if wl == 'c2' else 'Mtris/s',
        'value': value, 'unit': 'Mtris/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'frames_per_s': 1.0 / sec,
        'config': {'workload': WORKLOADS[wl]['desc'], 'l2': 'n/a (CPU)'},
        'cpu_baseline': {'value': value, 'unit': 'Mtris/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{frames} full frames (clear+render_occup+render_color), median'},
        'e2e': {'value': value, 'unit': 'Mtris/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import taichi_three_b200 as tina

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wl = args.workload
    w = WORKLOADS[wl]
    W, H = w['W'], w['H']
    inputs = make_inputs(wl, rank)
    nfaces = inputs['nfaces']

    scene = tina.Scene((W, H), smoothing=w['smoothing'], maxfaces=max(nfaces, 2**20), tonemap=False)
    material = tina.Classic() if w['material'] == 'classic' else tina.Diffuse()
    if w['kind'] == 'grid':
        mesh = tina.MeshGrid(w['n'])
        mesh.pos.from_numpy(inputs['pos'])
    elif w['kind'] == 'monkey':
        mesh = tina.MeshModel(inputs['obj'])
    else:
        mesh = tina.SimpleMesh(maxfaces=nfaces)
        mesh.set_face_verts(inputs['tri'])
    scene.add_object(mesh, material)
    scene.engine.set_camera(inputs['view'], inputs['proj'])
    engine, raster = scene.engine, scene.triangle_raster
    shader = scene.shaders[id(material)]
    bg = np.zeros(3, np.float32)

    raster.set_object(mesh)  # expanded face arrays now resident in HBM

    def step():
        engine.clear_depth()
        raster.render_occup()
        raster.render_color(shader, fill_bg=bg)

    flush = torch.empty(256 * 2**20 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def flush_l2():
        flush.fill_(1.0)

    for _ in range(max(3, args.warmup)):
        step()
        flush_l2()
    barrier()

    # ---- timed region: K steps, per-step CUDA events, L2 flushed between steps ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    K = args.steps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    from taichi_three_b200 import _lib as _tl
    launches0 = int(_tl.lib().tina_launch_count())
    wall0 = time.perf_counter()
    for a, b in evs:
        flush_l2()
        a.record()
        step()
        b.record()
    my_launches = int(_tl.lib().tina_launch_count()) - launches0  # this library's kernels inside the timed steps
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    my_ms = float(step_ms.sum())
    # keep the GPU under the same load until the clock sampler has a few samples
    clocks = None
    if sampler is not None:
        t_end = time.perf_counter() + max(0.0, 0.6 - wall)
        while time.perf_counter() < t_end:
            flush_l2()
            step()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        clocks['sampled'] = 'timed region + same workload looped to >= 0.6 s'
    tot = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms = float(tot.item())
    ms_per_step = total_ms / K
    value = world * nfaces / (ms_per_step * 1e-3) / 1e6

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
        flush_l2()
        step()
        torch.cuda.synchronize()
        for name, ms in raster.kernel_times().items():
            kt.setdefault(name, []).append(ms)
    raster.set_tuning(profile=0)
    kmean = {k: float(np.mean(v)) for k, v in kt.items() if np.mean(v) >= 0}

    # ---- e2e: host buffers, pinned H2D + set_object + step + pinned D2H, public API ----
    # Every step uploads that step's vertex data from pinned host memory and reads that step's image back
    # into pinned host memory.  Two frames are in flight on two CUDA streams (two Scene instances), so the
    # PCIe copies of one frame overlap the kernels of the other; `serial` = one frame at a time.
    def make_lane():
        sc = tina.Scene((W, H), smoothing=w['smoothing'], maxfaces=max(nfaces, 2**20), tonemap=False)
        mt = tina.Classic() if w['material'] == 'classic' else tina.Diffuse()
        if w['kind'] == 'grid':
            ms = tina.MeshGrid(w['n'])
            src = torch.as_tensor(inputs['pos']).pin_memory()
            dst = ms.pos.to_torch()
        elif w['kind'] == 'soup':
            ms = tina.SimpleMesh(maxfaces=nfaces)
            ms.set_face_verts(inputs['tri'])
            src = torch.as_tensor(inputs['tri']).pin_memory()
            dst = ms.verts.to_torch()
        else:
            ms = tina.MeshModel(inputs['obj'])
            src = torch.as_tensor(inputs['obj']['v']).pin_memory()
            dst = ms.verts
        sc.add_object(ms, mt)
        sc.engine.set_camera(inputs['view'], inputs['proj'])
        img_d = sc.image.to_torch()
        return dict(scene=sc, mesh=ms, raster=sc.triangle_raster, shader=sc.shaders[id(mt)], src=src, dst=dst, img_d=img_d,
                    img_h=torch.empty(img_d.shape, dtype=torch.float32).pin_memory(), stream=torch.cuda.Stream(device=dev),
                    done=torch.cuda.Event())

    lanes = [make_lane(), make_lane()]
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
        t = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    Ke = min(K, 100)
    e2e_run(4, 2)
    e2e_serial_s = e2e_run(Ke, 1)
    e2e_pipe_s = e2e_run(Ke, 2)
    e2e_s = min(e2e_serial_s, e2e_pipe_s)
    e2e_value = world * nfaces / e2e_s / 1e6
    checksum = float(lanes[0]['img_h'].double().sum().item())
    assert abs(checksum - float(lanes[1]['img_h'].double().sum().item())) < 1e-6 * max(1.0, abs(checksum))

    if rank == 0:
        peak, peak_src = load_peaks()
        frame_bytes, k1_bytes = alg_bytes(wl, nfaces)
        k1_ms = kmean.get('raster_faces', float('nan'))
        achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            sec, frames, threads = cpu_frame_time(wl, inputs, seconds=10.0)
            cpu = {'value': nfaces / sec / 1e6, 'unit': 'Mtris/s', 'cores': threads, 'kind': 'port',
                   'sample': f'{frames} full frames of the same workload (clear+render_occup+render_color), median, '
                             f'{sec * 1e3:.1f} ms/frame'}
        line = {
            'metric': 'Mtris/s (render_occup+render_color, 1080p)' if wl == 'c2' else 'Mtris/s (render_occup+render_color)',
            'value': value, 'unit': 'Mtris/s', 'n_gpus': world, 'steps': K, 'warmup': max(3, args.warmup),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': w['desc'], 'faces_per_step_per_gpu': nfaces, 'res': [W, H],
                       'step': 'clear_depth + render_occup + render_color(fill_bg)',
                       'l2': 'flushed between steps (256 MiB write, outside the per-step events)',
                       'parallelism': f'view-partitioned x{world}' if world > 1 else 'single GPU'},
            'frames_per_s': world / (ms_per_step * 1e-3),
            'frame_alg_bytes': frame_bytes,
            'frame_hbm_gbs': frame_bytes / (ms_per_step * 1e-3) / 1e9,
            'frame_roofline_frac': frame_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
            'ms_per_step_back_to_back_no_flush': b2b_ms,
            'step_ms_min_median_max': [float(step_ms.min()), float(np.median(step_ms)), float(step_ms.max())],
            'kernel_ms': kmean,
            'roofline': {'bound': 'hbm', 'kernel': 'k_raster_faces', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': ncu_traffic() if wl == 'c2' else None,
                         'traffic_source': 'profiles/r1_v10_k_raster_faces_ncu.txt (ncu --set full, dram read+write per launch)',
                         'note': 'issue-bound kernel: ncu smsp__issue_active 78 %, DRAM 6 % of peak (same file)',
                         'alg_bytes': k1_bytes, 'peak_source': peak_src},
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': 'Mtris/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_s * 1e3, 'steps': Ke, 'image_checksum': checksum,
                    'frames_in_flight': 2 if e2e_pipe_s <= e2e_serial_s else 1,
                    'ms_per_step_serial': e2e_serial_s * 1e3, 'ms_per_step_2_in_flight': e2e_pipe_s * 1e3},
            # counted by the library (tina_launch_count) on rank 0, x ranks: k_clear_keys, [k_vtx_clip], k_raster_faces,
            # [k_large_path unless the adaptive tile path is skipping it], k_render_color per step
            'gpu_launches': my_launches * world,
            'clocks': clocks,
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
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
