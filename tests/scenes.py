"""Seeded scene recipes shared by the parity tests, smoke() and bench.py.
Each recipe returns plain numpy inputs; `to_tina()` builds the product-side objects and
`to_oracle()` the expanded face arrays the CPU oracle consumes."""
import os

import numpy as np

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets')


def default_camera(aspect=1.0, **kw):
    import taichi_three_b200 as tina
    return tina.orbit_camera(aspect=aspect, **kw)


def wave_grid_pos(n, t=0.25):
    """examples/meshgrid_wave.py:8-16: z = 0.1 sin(10 |xy| - tau t); f64 -> f32 input generation."""
    from oracle import oracle as O
    pos, _ = O.grid_positions(n, n)
    xy = pos[..., :2].astype(np.float64)
    pos[..., 2] = (0.1 * np.sin(10 * np.sqrt((xy**2).sum(-1)) - 2 * np.pi * t)).astype(np.float32)
    return pos


def soup(n, W, H, s, seed=20240601, view=None, proj=None):
    """SURVEY §8d C3 recipe: n random front-facing triangles with uniform screen coverage."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if view is None:
        view, proj = default_camera(W / H)
    W2V = np.asarray(proj, np.float64) @ np.asarray(view, np.float64)
    V2W = np.linalg.inv(W2V)
    x = rng.uniform(-0.98, 0.98, n)
    y = rng.uniform(-0.98, 0.98, n)
    d = rng.uniform(2.0, 4.0, n)
    # view distance d -> ndc z through the projection (camera looks down -z in view space)
    zc = proj[2, 2] * (-d) + proj[2, 3]
    wc = -(-d)
    ndc = np.stack([x, y, zc / wc, np.ones(n)], axis=1)
    c = ndc @ V2W.T
    c = c[:, :3] / c[:, 3:4]
    e1 = rng.normal(0, 1, (n, 3)) * (s * d)[:, None]
    e2 = rng.normal(0, 1, (n, 3)) * (s * d)[:, None]
    tri = np.stack([c, c + e1, c + e2], axis=1).astype(np.float32)
    # make every face front-facing in screen space
    h = np.concatenate([tri.astype(np.float64), np.ones((n, 3, 1))], axis=2) @ W2V.T
    p = h[..., :2] / h[..., 3:4]
    facing = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    flip = facing <= 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return np.ascontiguousarray(tri)


# sizes calibrated once with the CPU oracle on a 2^19-face subsample (mean covered samples per face):
#   C3 3840x2160 target 8 px/N = 3.95 -> s = 0.001374 ; C5 7680x4320 target 1.98 -> s = 0.000487
SOUP_S_C3 = 0.001374
SOUP_S_C5 = 0.000487


def soup_torch(n, W, H, s, seed, device, view=None, proj=None, chunk=1 << 22):
    """The same recipe as soup() generated on the GPU with torch's generator (a different random
    stream than numpy's PCG64; used for the sizes that do not fit host memory comfortably)."""
    import torch
    if view is None:
        view, proj = default_camera(W / H)
    W2V = np.asarray(proj, np.float64) @ np.asarray(view, np.float64)
    V2Wt = torch.as_tensor(np.linalg.inv(W2V).T.copy(), device=device)
    W2Vt = torch.as_tensor(W2V.T.copy(), device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, 3, 3), dtype=torch.float32, device=device)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        u = torch.rand((m, 3), generator=g, device=device, dtype=torch.float64)
        x, y, d = u[:, 0] * 1.96 - 0.98, u[:, 1] * 1.96 - 0.98, u[:, 2] * 2.0 + 2.0
        zc = proj[2, 2] * (-d) + proj[2, 3]
        ndc = torch.stack([x, y, zc / d, torch.ones_like(x)], dim=1)
        c = ndc @ V2Wt
        c = c[:, :3] / c[:, 3:4]
        e = torch.randn((m, 2, 3), generator=g, device=device, dtype=torch.float64) * (s * d)[:, None, None]
        tri = torch.stack([c, c + e[:, 0], c + e[:, 1]], dim=1).to(torch.float32)
        h = torch.cat([tri.to(torch.float64), torch.ones((m, 3, 1), device=device, dtype=torch.float64)], dim=2) @ W2Vt
        p = h[..., :2] / h[..., 3:4]
        facing = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
        flip = facing <= 0
        tri[flip] = tri[flip][:, [0, 2, 1]]
        out[lo:lo + m] = tri
    return out


def load_monkey():
    import taichi_three_b200 as tina
    return tina.readobj(os.path.join(ASSETS, 'monkey.obj'))


def load_cornell():
    import taichi_three_b200 as tina
    return tina.readgltf(os.path.join(ASSETS, 'cornell.gltf'))


def cornell_views(n=64, theta=0.2):
    """SURVEY §8d C4 cameras: center (0,2,0), radius 6, fov 60, orbit in phi."""
    import taichi_three_b200 as tina
    return [tina.orbit_camera(center=(0, 2, 0), radius=6.0, theta=theta, phi=2 * np.pi * k / n) for k in range(n)]


def cornell_oracle_objects(gltf):
    """[(verts, norms, coors, material)] for the oracle, mirroring GltfScene.extract."""
    from oracle import oracle as O
    objs = []
    for node in gltf.nodes:
        for prim in node.primitives:
            v, vn, vt = O.indexed(prim.obj)
            v, vn = O.transform(v, vn, node.trans)
            objs.append((v, vn, vt, gltf._material(prim.material)))
    return objs
