"""Multi-GPU sharding of the raster path.  One process per GPU (torchrun); the
reference has no distributed code, so this is host-side orchestration over the same kernels:

  * view-parallel: independent views / frames are dealt round-robin to ranks, no collective;
  * sort-last: contiguous face ranges per rank, global face ids via Engine.face_base, one elementwise
    MIN all-reduce of the packed int64 (depth, face-id) keys (NCCL over NVLink on GPUs, gloo in the CPU
    tests), after which every rank shades exactly the pixels whose winning face it owns.  `min` on the
    packed key is associative and commutative, so the result is bit-identical to a single GPU.
"""
import numpy as np
import torch
import torch.distributed as dist

CLEAR_KEY = (2**30) << 32  # engine.py:68-70 depth clear, no winner


def view_partition(nviews, rank, world):
    """Indices of the views rank `rank` renders (k = rank mod world)."""
    return list(range(rank, nviews, world))


def face_range(nfaces, rank, world):
    """Contiguous [lo, hi) face range of `rank`; ranges tile [0, nfaces) in rank order."""
    base, rem = divmod(nfaces, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_keys(depth, occup, face_base=0):
    """int32 depth + int32 occup (-1 = none) -> int64 keys, as the kernels store them."""
    depth = torch.as_tensor(depth).to(torch.int64)
    ids = torch.as_tensor(occup).to(torch.int64) + 1 + face_base
    ids = torch.where(torch.as_tensor(occup) < 0, torch.zeros_like(ids), ids)
    return (depth << 32) | ids


def unpack_keys(keys, face_base=0, nfaces=None):
    """-> (depth int32, occup int32) for faces [face_base, face_base + nfaces)."""
    keys = torch.as_tensor(keys)
    depth = (keys >> 32).to(torch.int32)
    ids = keys & 0xffffffff
    f = ids - 1 - face_base
    ok = (ids != 0) & (f >= 0)
    if nfaces is not None:
        ok &= f < nfaces
    return depth, torch.where(ok, f, torch.full_like(f, -1)).to(torch.int32)


def composite_min(keys, group=None):
    """Sort-last merge: in-place elementwise MIN of the packed keys across ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


def strip_range(npix, rank, world, root=None, root_share=1.0):
    """Contiguous pixel strip [lo, hi) of `rank` (x-major pixel indices, multiples of 256).  root / root_share: the
    rank that also receives everybody's image strips takes `root_share` of an equal share (0 = it shades nothing), so
    that the keys it reads and the image strips it receives do not pile up on its one NVLink port."""
    nblk = (npix + 255) // 256
    if root is None or root_share == 1.0 or world < 2:
        per = (nblk + world - 1) // world
        return min(rank * per * 256, npix), min((rank + 1) * per * 256, npix)
    w = [float(root_share) if r == root else 1.0 for r in range(world)]
    tot = sum(w)
    edge = lambda r: int(round(sum(w[:r]) / tot * nblk))  # noqa: E731
    lo, hi = (0 if rank == 0 else edge(rank)), (nblk if rank == world - 1 else edge(rank + 1))
    return min(lo * 256, npix), min(hi * 256, npix)


def sort_last_strip(npix, rank, world, composite='p2p', gather='root', root_share=None):
    """The strip render_sort_last_replicated gives `rank` (same defaults): -> (lo, hi, root_share used)."""
    if root_share is None:
        import os
        env = os.environ.get('TINA_ROOT_SHARE')  # experiments
        root_share = float(env) if env is not None else (0.0 if (world >= 4 and composite == 'p2p' and gather == 'root') else 1.0)
    if not (composite == 'p2p' and gather == 'root'):
        root_share = 1.0  # (the collectives need equal strips)
    lo, hi = strip_range(npix, rank, world, root=0, root_share=root_share)
    return lo, hi, root_share


class SharedImage:
    """A float32 [W, H, 3] frame image that lives on rank `root` and is mapped into every rank of the node
    (CUDA IPC): the ranks' shading kernels store their screen strips straight into it over NVLink, so a sort-last
    frame needs no gather.  `.tensor` is the image (the root's memory on every rank).  Collective constructor."""

    def __init__(self, res, group=None, root=0):
        import ctypes as C
        from . import _lib
        from .field import wrap_device
        self.group, self.root = group, root
        self.rank = dist.get_rank(group)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._ptr = C.c_void_p()
        nbytes = int(res[0]) * int(res[1]) * 3 * 4
        handle = (C.c_uint8 * 64)()
        if self.rank == root:
            _lib.check(_lib.lib().tina_shared_alloc(self.device.index, nbytes, C.byref(self._ptr), handle))
        t = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        dist.broadcast(t, src=dist.get_global_rank(group, root) if group is not None else root, group=group)
        if self.rank != root:
            buf = (C.c_uint8 * 64)(*t.cpu().tolist())
            _lib.check(_lib.lib().tina_shared_open(self.device.index, buf, C.byref(self._ptr)))
        self.tensor = wrap_device(self._ptr.value, (int(res[0]), int(res[1]), 3), torch.float32, self.device, owner=self)

    def close(self):
        from . import _lib
        if getattr(self, '_ptr', None) and self._ptr.value:
            dist.barrier(group=self.group)  # nobody is still writing into the root's buffer
            if self.rank == self.root:
                _lib.lib().tina_shared_free(self.device.index, self._ptr)
            else:
                _lib.lib().tina_shared_close(self.device.index, self._ptr)
            self._ptr = None


def render_sort_last_replicated(engine, raster, verts, norms, coors, shader, bgcolor=0.0, group=None, composite='nccl', gather='all',
                                root_share=None):
    """Sort-last frame with REPLICATED face attributes (every rank holds all N faces, C5: 4.8 GB):
    rank r rasterises faces face_range(N, r, G) with global ids, the keys are MIN-reduce-scattered so
    rank r ends up with the final keys of screen strip r (1/G of the all-reduce traffic), shades that
    strip from the full attribute arrays, and the image strips are all-gathered.  Bit-identical to one
    GPU.  Needs W*H divisible by 256*G (else falls back to the all-reduce composite).

    composite='p2p' (after engine.open_peer_keys(group), ranks of one node): no reduce-scatter at all -- after a
    barrier the strip's shading kernel takes each pixel's key as the MIN over all ranks' key buffers, read over NVLink
    peer memory inside the kernel (TriangleRaster.render_color_composite).  Same bits.

    gather='root' (shader.img wraps a SharedImage tensor): the image is assembled on the root rank only, by the
    shading kernels' own stores over NVLink; no all-gather.  root_share (p2p + root only): the share of an equal strip
    rank 0 takes (default 0 from 4 ranks on: the root's NVLink port already receives every other rank's image strip --
    348 MB at 8K on 8 GPUs -- and would read 7/8 of its own strip's keys through it as well)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    N = verts.shape[0]
    lo, hi = face_range(N, rank, world)
    npix = engine.res[0] * engine.res[1]
    engine.clear_depth()
    engine.set_face_base(lo)
    raster.set_face_verts(verts[lo:hi])
    if raster.smoothing:
        raster.set_face_norms(norms[lo:hi])
    if raster.texturing:
        raster.set_face_coors(coors[lo:hi])
    raster.render_occup()
    keys = engine.keys.view(-1)
    img = shader.img.to_torch() if hasattr(shader.img, 'to_torch') else shader.img
    # the full arrays for shading (ids in the keys are global)
    raster.set_face_verts(verts)
    if raster.smoothing:
        raster.set_face_norms(norms)
    if raster.texturing:
        raster.set_face_coors(coors)
    if world > 1 and npix % (256 * world) == 0:
        p_lo, p_hi, _ = sort_last_strip(npix, rank, world, composite, gather, root_share)
        if composite == 'p2p':
            # a one-word all-reduce on the same stream: it completes on this rank only after every rank has
            # enqueued it behind its own render_occup (dist.barrier() would also block the host)
            tok = torch.zeros(1, dtype=torch.int32, device=keys.device)
            dist.all_reduce(tok, group=group)
            raster.render_color_composite(shader, p_lo, p_hi - p_lo, face_base=0, fill_bg=bgcolor)
        else:
            strip = torch.empty(p_hi - p_lo, dtype=torch.int64, device=keys.device)
            dist.reduce_scatter_tensor(strip, keys, op=dist.ReduceOp.MIN, group=group)
            keys[p_lo:p_hi] = strip
            raster.render_color_range(shader, p_lo, p_hi - p_lo, face_base=0, fill_bg=bgcolor)
        if gather == 'root':
            # shader.img is a SharedImage: the strip has been stored into the root's memory by the shading kernel.
            # One-word all-reduce = end-of-frame fence: the root's image is complete, and no rank clears its keys
            # while a peer still reads them.
            tok = torch.zeros(1, dtype=torch.int32, device=keys.device)
            dist.all_reduce(tok, group=group)
        else:
            flat = img.view(-1)
            dist.all_gather_into_tensor(flat, flat[p_lo * 3:p_hi * 3].clone(), group=group)
    else:
        composite_min(keys, group)
        raster.render_color_range(shader, 0, npix, face_base=0, fill_bg=bgcolor)
    return img


def render_sort_last(engine, raster, verts, norms, coors, shader, bgcolor=0.0, group=None):
    """One sort-last frame.  `verts/norms/coors` are this rank's CUDA slices of the global face arrays
    (range = face_range(N, rank, world)); returns the composited [W, H, 3] image on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    counts = torch.tensor([verts.shape[0]], dtype=torch.int64, device=verts.device)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    if world > 1:
        dist.all_gather(allc, counts, group=group)
    else:
        allc = [counts]
    lo = int(sum(int(c.item()) for c in allc[:rank]))
    engine.clear_depth()
    engine.set_face_base(lo)
    raster.set_face_verts(verts)
    if raster.smoothing:
        raster.set_face_norms(norms)
    if raster.texturing:
        raster.set_face_coors(coors)
    raster.render_occup()
    composite_min(engine.keys, group)
    img = shader.img.to_torch() if hasattr(shader.img, 'to_torch') else shader.img
    img.zero_()
    raster.render_color(shader)  # shades pixels whose winner is in [lo, lo + n)
    if world > 1:
        dist.all_reduce(img, op=dist.ReduceOp.SUM, group=group)  # exactly one non-zero contributor per pixel
    empty = (engine.keys & 0xffffffff) == 0
    bg = torch.as_tensor(np.broadcast_to(np.asarray(bgcolor, dtype=np.float32), (3,)).copy(), device=img.device)
    img[empty] = bg
    return img
