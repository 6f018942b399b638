"""Product-free input recipes for bench.py's reference arm -- TEST INFRASTRUCTURE ONLY.

`bench.py --impl reference` must time the CPU restatement of the reference with none of the product's
code on the path, so the camera and the C2 / C3 / C1 inputs are restated here from the reference
(tina/util/matrix.py:8-12,48-69, tina/util/control.py:10-15,102-113, examples/meshgrid_wave.py:8-16,
SURVEY.md 8d).  tests/test_cpu.py checks these against the recipes the product-side bench uses."""
import os

import numpy as np


def frustum(left, right, bottom, top, near, far):
    """tina/util/matrix.py:48-58"""
    m = np.eye(4)
    m[0, 0] = 2 * near / (right - left)
    m[1, 1] = 2 * near / (top - bottom)
    m[0, 2] = (right + left) / (right - left)
    m[1, 2] = (top + bottom) / (top - bottom)
    m[2, 2] = -(far + near) / (far - near)
    m[2, 3] = -2 * far * near / (far - near)
    m[3, 2] = -1
    m[3, 3] = 0
    return m


def perspective(fov=60, aspect=1.0, near=0.05, far=500):
    """tina/util/matrix.py:66-69"""
    f = np.tan(np.radians(fov) / 2)
    ax, ay = f * aspect, f
    return frustum(-near * ax, near * ax, -near * ay, near * ay, near, far)


def default_camera(aspect=1.0, radius=3.0, fov=60):
    """Control defaults (control.py:10-15: center 0, radius 3, R = I, fov 60) -> (view, proj) as get_camera (:102-113)."""
    cam = np.eye(4)
    cam[2, 3] = radius  # affine(I, center + R @ (0, 0, radius))
    return np.linalg.inv(cam), perspective(fov, aspect)


def grid_positions(n):
    """mesh/grid.py:17-21 in f32"""
    u = (np.arange(n, dtype=np.float32) / np.float32(n - 1))[:, None] * np.ones((1, n), np.float32)
    v = np.ones((n, 1), np.float32) * (np.arange(n, dtype=np.float32) / np.float32(n - 1))[None, :]
    return np.ascontiguousarray(np.stack([u * np.float32(2) - np.float32(1), v * np.float32(2) - np.float32(1), np.zeros_like(u)], axis=2))


def wave_grid_pos(n, t=0.25):
    """examples/meshgrid_wave.py:8-16: z = 0.1 sin(10 |xy| - tau t), f64 -> f32"""
    pos = grid_positions(n)
    xy = pos[..., :2].astype(np.float64)
    pos[..., 2] = (0.1 * np.sin(10 * np.sqrt((xy**2).sum(-1)) - 2 * np.pi * t)).astype(np.float32)
    return pos


def soup(n, W, H, s, seed=20240601):
    """SURVEY 8d C3 recipe (numpy PCG64 stream): n random front-facing triangles, uniform screen coverage."""
    rng = np.random.Generator(np.random.PCG64(seed))
    view, proj = default_camera(W / H)
    W2V = proj @ view
    V2W = np.linalg.inv(W2V)
    x, y, d = rng.uniform(-0.98, 0.98, n), rng.uniform(-0.98, 0.98, n), rng.uniform(2.0, 4.0, n)
    zc = proj[2, 2] * (-d) + proj[2, 3]
    ndc = np.stack([x, y, zc / d, np.ones(n)], axis=1)
    c = ndc @ V2W.T
    c = c[:, :3] / c[:, 3:4]
    e1 = rng.normal(0, 1, (n, 3)) * (s * d)[:, None]
    e2 = rng.normal(0, 1, (n, 3)) * (s * d)[:, None]
    tri = np.stack([c, c + e1, c + e2], axis=1).astype(np.float32)
    h = np.concatenate([tri.astype(np.float64), np.ones((n, 3, 1))], axis=2) @ W2V.T
    p = h[..., :2] / h[..., 3:4]
    facing = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    flip = facing <= 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return np.ascontiguousarray(tri)


def monkey_faces(path):
    """assets/monkey.obj -> [N,3,3] face positions (tina/assimp/obj.py:4-14 _tri_append, positions only)."""
    v, f = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == 'v':
            v.append([float(x) for x in t[1:4]])
        elif t[0] == 'f':
            idx = [int(w.split('/')[0]) - 1 for w in t[1:]]
            if len(idx) == 4:  # obj.py:7-9: (0,1,2), (2,3,0)
                f += [[idx[0], idx[1], idx[2]], [idx[2], idx[3], idx[0]]]
            else:              # :5-6 triangles, :10-12 fans
                f += [[idx[0], idx[k], idx[k + 1]] for k in range(1, len(idx) - 1)]
    return np.ascontiguousarray(np.asarray(v, np.float32)[np.asarray(f)])
