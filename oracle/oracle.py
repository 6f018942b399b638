"""ctypes wrapper of the CPU oracle (oracle/tina_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package taichi_three_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtina_oracle.so')

SMOOTHING, TEXTURING, CULLING, CLIPPING = 1, 2, 4, 8

_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, 'tina_oracle.c'), os.path.join(os.path.dirname(_HERE), 'include', 'tina_b200.h')]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(p) for p in srcs):
        subprocess.run(['make', '-C', _HERE, 'libtina_oracle.so'], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n=None):
    """Use n host threads (default: every core this process may run on) regardless of OMP_NUM_THREADS."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def clear_depth(W, H):
    return np.full((W, H), 2**30, dtype=np.int32)


def render_occup(verts, W2V, W, H, flags=CULLING | CLIPPING, bias=(0.5, 0.5), depth=None, parallel=False):
    """-> (occup[W,H] i32, depth[W,H] i32 (updated copy), tie[W,H] u8, stats dict).  Serial = deterministic."""
    verts = _f(verts).reshape(-1, 9)
    W2V, bias = _f(W2V).reshape(16), _f(bias)
    depth = clear_depth(W, H) if depth is None else np.ascontiguousarray(depth, dtype=np.int32).copy()
    occup = np.empty((W, H), dtype=np.int32)
    if parallel:
        lib().orc_render_occup_parallel(_p(verts), C.c_int64(len(verts)), _p(W2V), _p(bias), W, H, C.c_uint32(flags),
                                        _p(depth), _p(occup))
        return occup, depth, None, {}
    tie = np.zeros((W, H), dtype=np.uint8)
    stats = np.zeros(5, dtype=np.int64)
    lib().orc_render_occup(_p(verts), C.c_int64(len(verts)), _p(W2V), _p(bias), W, H, C.c_uint32(flags), _p(depth),
                           _p(occup), _p(tie), _p(stats))
    keys = ('culled', 'clipped', 'rasterised', 'candidates', 'covered')
    return occup, depth, tie, dict(zip(keys, stats.tolist()))


def face_setup(verts, W2V, W, H, flags=CULLING | CLIPPING):
    verts = _f(verts).reshape(-1, 9)
    out = np.zeros((len(verts), 16), dtype=np.float32)
    bbox = np.zeros((len(verts), 4), dtype=np.int32)
    lib().orc_face_setup(_p(verts), C.c_int64(len(verts)), _p(_f(W2V).reshape(16)), W, H, C.c_uint32(flags), _p(out), _p(bbox))
    return out, bbox


def _material_pod(material):
    """the plain (unfolded, unhoisted) material program + host texture pointers, from the oracle's OWN
    front-end (oracle/materials.py): a node graph (walked by duck typing), or an already built pod"""
    from . import materials as M
    if isinstance(material, tuple) and len(material) == 3 and isinstance(material[0], M.MaterialPOD):
        return material
    return M.material_pod_of(material)


def _lighting_pod(lighting):
    from . import materials as M
    if isinstance(lighting, M.LightingPOD):
        return lighting
    if lighting is None:
        return M.default_lighting()
    return M.lighting_of(lighting)


def pars_occup(verts, sizes, W2V, V2W, W, H, clipping=True, bias=(0.5, 0.5), depth=None):
    """core/particle.py:78-127 -> (occup, depth)"""
    verts, sizes = _f(verts).reshape(-1, 3), _f(sizes).reshape(-1)
    depth = clear_depth(W, H) if depth is None else np.ascontiguousarray(depth, dtype=np.int32).copy()
    occup = np.empty((W, H), dtype=np.int32)
    lib().orc_pars_occup(_p(verts), _p(sizes), C.c_int64(len(verts)), _p(_f(W2V).reshape(16)), _p(_f(V2W).reshape(16)),
                         _p(_f(bias)), W, H, C.c_uint32(2 if clipping else 0), _p(depth), _p(occup))
    return occup, depth


def pars_color(verts, sizes, colors, occup, W2V, V2W, W, H, material, lighting, image, bias=(0.5, 0.5)):
    """core/particle.py:129-161; shades pixels with occup != -1 into image (in place)."""
    m, texptrs, keep = _material_pod(material)
    L = _lighting_pod(lighting)
    verts, sizes = _f(verts).reshape(-1, 3), _f(sizes).reshape(-1)
    colors = _f(colors).reshape(-1, 3) if colors is not None else None
    lib().orc_pars_color(_p(verts), _p(sizes), _p(colors), _p(np.ascontiguousarray(occup, dtype=np.int32)),
                         _p(_f(W2V).reshape(16)), _p(_f(V2W).reshape(16)), _p(_f(bias)), W, H, C.byref(m), texptrs, C.byref(L),
                         _p(image))
    return image


def pars_attrs(verts, sizes, occup, W2V, V2W, W, H, bias=(0.5, 0.5)):
    """core/particle.py:129-158: (pos, normal) [W, H, 3] of the visible sphere points (zero where occup == -1)."""
    verts, sizes = _f(verts).reshape(-1, 3), _f(sizes).reshape(-1)
    pos, nrm = np.zeros((W, H, 3), np.float32), np.zeros((W, H, 3), np.float32)
    lib().orc_pars_attrs(_p(verts), _p(sizes), _p(np.ascontiguousarray(occup, dtype=np.int32)), _p(_f(W2V).reshape(16)),
                         _p(_f(V2W).reshape(16)), _p(_f(bias)), W, H, _p(pos), _p(nrm))
    return pos, nrm


def render_color(verts, norms, coors, occup, W2V, V2W, W, H, flags, material, lighting, image, bias=(0.5, 0.5),
                 parallel=True):
    """Shades pixels with occup != -1 into `image` ([W,H,3] f32, modified in place and returned).
    `material`: a material node graph (the product's or the reference's classes, walked by oracle/materials.py) or a
    pod built by oracle.materials.stock_*; `lighting`: a Lighting object, a LightingPOD, or None = the default light."""
    m, texptrs, tex_arrays = _material_pod(material)
    L = _lighting_pod(lighting)
    verts = _f(verts).reshape(-1, 9)
    norms = _f(norms).reshape(-1, 9) if norms is not None else None
    coors = _f(coors).reshape(-1, 6) if coors is not None else None
    occup = np.ascontiguousarray(occup, dtype=np.int32)
    assert image.dtype == np.float32 and image.flags['C_CONTIGUOUS'] and image.shape == (W, H, 3)
    lib().orc_render_color(_p(verts), _p(norms), _p(coors), _p(occup), _p(_f(W2V).reshape(16)), _p(_f(V2W).reshape(16)),
                           _p(_f(bias)), W, H, C.c_uint32(flags), C.byref(m), texptrs, C.byref(L), _p(image),
                           1 if parallel else 0)
    return image


def render_gbuffer(kind, verts, norms, coors, occup, depth, W2V, V2W, W, H, flags, out, param=(0, 0, 0), bias=(0.5, 0.5)):
    """core/shader.py:21-109 sinks; `out` [W,H,n] f32 is modified in place where occup != -1."""
    verts = _f(verts).reshape(-1, 9)
    norms = _f(norms).reshape(-1, 9) if norms is not None else None
    coors = _f(coors).reshape(-1, 6) if coors is not None else None
    assert out.dtype == np.float32 and out.flags['C_CONTIGUOUS']
    ncomp = out.size // (W * H)
    lib().orc_render_gbuffer(_p(verts), _p(norms), _p(coors), _p(np.ascontiguousarray(occup, dtype=np.int32)),
                             _p(np.ascontiguousarray(depth, dtype=np.int32)), _p(_f(W2V).reshape(16)), _p(_f(V2W).reshape(16)),
                             _p(_f(bias)), W, H, C.c_uint32(flags), int(kind), _p(_f(np.resize(np.asarray(param, np.float32), 3))),
                             _p(out), ncomp)
    return out


def wire_render(verts, W2V, W, H, depth, image, color=(.9, .6, 0), clipping=False, bias=(0.5, 0.5)):
    """core/wireframe.py:70-95 on wires [N,2,3]; -> (depth copy updated, image modified in place)"""
    verts = _f(verts).reshape(-1, 6)
    depth = np.ascontiguousarray(depth, dtype=np.int32).copy()
    lib().orc_wire_render(_p(verts), C.c_int64(len(verts)), _p(_f(W2V).reshape(16)), _p(_f(bias)), W, H,
                          C.c_uint32(2 if clipping else 0), _p(_f(color)), _p(depth), _p(image))
    return depth, image


def mesh_to_wires(face_verts):
    """mesh/wire.py:19-27 on expanded faces [N,3,3] -> [3N,2,3]"""
    fv = _f(face_verts)
    idx = np.arange(len(fv) * 3)
    return np.ascontiguousarray(np.stack([fv[idx // 3, idx % 3], fv[idx // 3, (idx + 1) % 3]], axis=1))


def fxaa(image, abs_thresh=0.0625, rel_thresh=0.063, factor=1.0):
    out = np.ascontiguousarray(image, dtype=np.float32).copy()
    lib().orc_fxaa(_p(out), out.shape[0], out.shape[1], C.c_float(abs_thresh), C.c_float(rel_thresh), C.c_float(factor))
    return out


def bloom(image, gwei, thresh=1.0, scale=0.25, factor=1.0):
    out = np.ascontiguousarray(image, dtype=np.float32).copy()
    g = _f(gwei)
    lib().orc_bloom(_p(out), out.shape[0], out.shape[1], _p(g), len(g) - 1, C.c_float(thresh), C.c_float(scale), C.c_float(factor))
    return out


def ssao_render(depth, normals, W2V, V2W, samples, rotations, radius=0.2, thresh=0.0, factor=1.0, bias=(0.5, 0.5)):
    depth = np.ascontiguousarray(depth, dtype=np.int32)
    W, H = depth.shape
    nrm, smp, rot = _f(normals), _f(samples), _f(rotations)
    w2v, v2w, b = _f(W2V), _f(V2W), _f(bias)
    ao = np.zeros((W, H), dtype=np.float32)
    lib().orc_ssao_render(depth.ctypes.data_as(C.c_void_p), _p(nrm), _p(w2v), _p(v2w), _p(b), W, H, _p(smp), smp.shape[0], _p(rot),
                          rot.shape[0], C.c_float(radius), C.c_float(thresh), C.c_float(factor), _p(ao))
    return ao


def ssao_apply(image, ao, noise_size=4):
    out = np.ascontiguousarray(image, dtype=np.float32).copy()
    a = _f(ao)
    lib().orc_ssao_apply(_p(out), _p(a), out.shape[0], out.shape[1], int(noise_size))
    return out


def ssao_render_taa(depth, normals, W2V, V2W, nsamples=64, radius=0.2, thresh=0.0, factor=1.0, frame=0, bias=(0.5, 0.5)):
    """postp/ssao.py with taa=True: per-pixel, per-frame samples from the hash stream (include/tina_b200.h)."""
    depth = np.ascontiguousarray(depth, dtype=np.int32)
    W, H = depth.shape
    nrm, w2v, v2w, b = _f(normals), _f(W2V), _f(V2W), _f(bias)
    ao = np.zeros((W, H), dtype=np.float32)
    lib().orc_ssao_render_taa(depth.ctypes.data_as(C.c_void_p), _p(nrm), _p(w2v), _p(v2w), _p(b), W, H, int(nsamples),
                              C.c_float(radius), C.c_float(thresh), C.c_float(factor), C.c_uint32(frame), _p(ao))
    return ao


def ssao_apply_taa(image, ao):
    out = np.ascontiguousarray(image, dtype=np.float32).copy()
    a = _f(ao)
    lib().orc_ssao_apply_taa(_p(out), _p(a), out.shape[0], out.shape[1])
    return out


def ssr_render(depth, normals, coors, mtlid, materials, image, W2V, V2W, nsamples=32, nsteps=32, stepsize=2.0, tolerance=15.0,
               blurring=4, taa=False, frame=0, bias=(0.5, 0.5)):
    """postp/ssr.py:44-103.  materials: list of material node graphs (the scene's material table, scene/raster.py:43-49),
    flattened by the oracle's own front-end (oracle/materials.py).  -> img4 [W, H, 4]."""
    from . import materials as OM
    depth = np.ascontiguousarray(depth, dtype=np.int32)
    W, H = depth.shape
    pods = [OM.sample_pod_of(m) for m in materials]
    table = (OM.SampleMaterialPOD * max(1, len(pods)))()
    texhost = (C.c_void_p * (OM.MAX_TEX * max(1, len(pods))))()
    for i, (pod, arrays) in enumerate(pods):
        table[i] = pod
        for t, a in enumerate(arrays):
            texhost[i * OM.MAX_TEX + t] = a.ctypes.data
    nrm, img, w2v, v2w, b = _f(normals), _f(image), _f(W2V), _f(V2W), _f(bias)
    co = _f(coors) if coors is not None else None
    mid = np.ascontiguousarray(mtlid, dtype=np.int32)
    out = np.zeros((W, H, 4), dtype=np.float32)
    lib().orc_ssr_render(depth.ctypes.data_as(C.c_void_p), _p(nrm), _p(co), mid.ctypes.data_as(C.c_void_p), table, texhost, len(pods),
                         _p(img), _p(w2v), _p(v2w), _p(b), W, H, int(nsamples), int(nsteps), C.c_float(stepsize), C.c_float(tolerance),
                         int(blurring), int(bool(taa)), C.c_uint32(frame), _p(out))
    return out


def ssr_apply(image, img4, blurring=4, taa=False):
    out = np.ascontiguousarray(image, dtype=np.float32).copy()
    a = _f(img4)
    lib().orc_ssr_apply(_p(out), _p(a), out.shape[0], out.shape[1], int(blurring), int(bool(taa)))
    return out


def tonemap(image):
    out = np.ascontiguousarray(image, dtype=np.float32).copy()
    lib().orc_tonemap(_p(out), C.c_int64(out.size))
    return out


# ---- mesh providers (set_object side) ---------------------------------------------------
def grid_normals(pos):
    pos = _f(pos)
    nx, ny = pos.shape[:2]
    nrm = np.empty_like(pos)
    lib().orc_grid_normals(_p(pos), nx, ny, _p(nrm))
    return nrm


def grid_faces(prop):
    prop = _f(prop)
    nx, ny, dim = prop.shape
    out = np.empty((2 * (nx - 1) * (ny - 1), 3, dim), dtype=np.float32)
    lib().orc_grid_faces(_p(prop), nx, ny, dim, _p(out))
    return out


def grid_positions(nx, ny):
    """mesh/grid.py:17-21 in f32."""
    u = (np.arange(nx, dtype=np.float32) / np.float32(nx - 1))[:, None] * np.ones((1, ny), np.float32)
    v = np.ones((nx, 1), np.float32) * (np.arange(ny, dtype=np.float32) / np.float32(ny - 1))[None, :]
    pos = np.stack([u * np.float32(2) - np.float32(1), v * np.float32(2) - np.float32(1), np.zeros_like(u)], axis=2)
    return np.ascontiguousarray(pos, dtype=np.float32), np.ascontiguousarray(np.stack([u, v], axis=2), dtype=np.float32)


def transform(verts, norms, trans):
    """mesh/trans.py:23-40 on [N,3,3] arrays."""
    trans = np.asarray(trans, dtype=np.float64)
    t32 = _f(trans).reshape(16)
    tn = _f(np.transpose(np.linalg.inv(trans))[:3, :3]).reshape(9)
    verts = _f(verts).copy()
    lib().orc_transform_verts(_p(verts), C.c_int64(verts.size // 3), _p(t32))
    if norms is not None:
        norms = _f(norms).copy()
        lib().orc_transform_norms(_p(norms), C.c_int64(norms.size // 3), _p(tn))
    return verts, norms


def no_culling(verts, norms=None, coors=None):
    """mesh/cull.py:31-57: [N,...] -> [2N,...], odd copies reversed, normals negated."""
    def dup(a, negate=False):
        if a is None:
            return None
        out = np.repeat(a, 2, axis=0)
        out[1::2] = out[1::2][:, ::-1]
        if negate:
            out[1::2] = -out[1::2]
        return np.ascontiguousarray(out)
    return dup(verts), dup(norms, True), dup(coors)


def indexed(obj):
    """mesh/model.py:56-73: dict {'v','vt','vn','f'} -> ([N,3,3] verts, norms, [N,3,2] coors)."""
    f = np.asarray(obj['f'])
    if f.ndim == 2:
        f = np.stack([f, f, f], axis=2)
    f = f.astype(np.int64)
    v = _f(obj['v'])[f[:, :, 0]]
    vt = _f(obj['vt'])[:, :2][f[:, :, 1]] if 'vt' in obj else None
    vn = _f(obj['vn'])[f[:, :, 2]] if 'vn' in obj else None
    return np.ascontiguousarray(v), (np.ascontiguousarray(vn) if vn is not None else None), \
        (np.ascontiguousarray(vt) if vt is not None else None)


def render_scene(objects, W, H, view, proj, lighting, flags, bgcolor=0.0, do_tonemap=True, bias=(0.5, 0.5)):
    """scene/raster.py:168-207 for a list of (verts, norms, coors, material).
    -> dict(image, pre_tonemap, depth, occups=[...], ties=[...])"""
    W2V64 = np.asarray(proj, dtype=np.float64) @ np.asarray(view, dtype=np.float64)
    W2V, V2W = W2V64.astype(np.float32), np.linalg.inv(W2V64).astype(np.float32)
    image = np.empty((W, H, 3), dtype=np.float32)
    image[...] = np.broadcast_to(np.asarray(bgcolor, dtype=np.float32), (3,))
    depth = clear_depth(W, H)
    occups, ties = [], []
    for verts, norms, coors, material in objects:
        occup, depth, tie, _ = render_occup(verts, W2V, W, H, flags, bias, depth)
        render_color(verts, norms, coors, occup, W2V, V2W, W, H, flags, material, lighting, image, bias)
        occups.append(occup)
        ties.append(tie)
    pre = image.copy()
    if do_tonemap:
        image = tonemap(image)
    return dict(image=image, pre_tonemap=pre, depth=depth, occups=occups, ties=ties, W2V=W2V, V2W=V2W)
