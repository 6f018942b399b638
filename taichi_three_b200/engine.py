"""Engine: per-view state shared by rasterisers (reference tina/core/engine.py:5-76).

The depth buffer lives in the high words of the int64 visibility keys owned by
libtina_b200 (see include/tina_b200.h); `engine.depth` is a zero-copy strided view.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .field import Field, HostField, wrap_device


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream(device_index=None):
    """torch's current CUDA stream as a void* (the fast private getter when torch has it)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device() if device_index is None else device_index))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    def __init__(self, res_x=512, res_y=None, device=None):
        if res_y is None:
            res_y = res_x
        res = (res_x, res_y) if isinstance(res_x, (int, np.integer)) else tuple(res_x)
        self.res = (int(res[0]), int(res[1]))
        self.maxdepth = 2**30  # engine.py:12
        if not torch.cuda.is_available():
            raise RuntimeError('taichi_three_b200.Engine needs a CUDA device (no CPU fallback)')
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.tina_engine_create(C.byref(h), self.device.index, self.res[0], self.res[1]))
        self._h = h
        self._clear = L.tina_engine_clear_depth
        self._flush = L.tina_engine_flush
        kp = C.c_void_p()
        _lib.check(L.tina_engine_keys(self._h, C.byref(kp)))
        self._keys = wrap_device(kp.value, self.res, torch.int64, self.device, owner=self)
        # engine.py:21-26
        m = np.eye(4, dtype=np.float32)
        m[2, 2] = -1
        self.W2V = HostField(m.copy(), self._push_camera)
        self.V2W = HostField(m.copy(), self._push_camera)
        self.bias = HostField(np.array([0.5, 0.5], dtype=np.float32), self._push_camera)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                _lib.lib().tina_engine_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _push_camera(self):
        L = _lib.lib()
        _, w2v = _lib.f32_array(self.W2V.to_numpy(), 16)
        _, v2w = _lib.f32_array(self.V2W.to_numpy(), 16)
        _lib.check(L.tina_engine_set_camera(self._h, w2v, v2w))
        b = self.bias.to_numpy()
        _lib.check(L.tina_engine_set_bias(self._h, float(b[0]), float(b[1])))

    @property
    def keys(self):
        """int64[W, H]  key = depth << 32 | (global face id + 1).  (clear_depth is deferred inside the library until
        something touches the keys; handing out the tensor is such a touch.)"""
        rc = self._flush(self._h, _stream(self.device.index))
        if rc:
            _lib.check(rc)
        return self._keys

    @property
    def depth(self):
        """int32[W, H] view of the depth words (engine.py:11)."""
        return Field(self.keys.view(torch.int32).view(self.res[0], self.res[1], 2)[..., 1])

    def set_camera(self, view, proj):
        """engine.py:72-76: W2V = proj @ view (float64), V2W = inv(W2V), both stored as f32."""
        W2V = np.asarray(proj, dtype=np.float64) @ np.asarray(view, dtype=np.float64)
        V2W = np.linalg.inv(W2V)
        self.W2V._v[...] = W2V.astype(np.float32)
        self.V2W._v[...] = V2W.astype(np.float32)
        self._push_camera()

    def clear_depth(self):
        """engine.py:68-70."""
        rc = self._clear(self._h, _stream(self.device.index))
        if rc:
            _lib.check(rc)

    def set_face_base(self, base):
        """Offset of the next render_occup's face ids (sort-last multi-GPU: global ids)."""
        _lib.check(_lib.lib().tina_engine_set_face_base(self._h, int(base)))

    def open_peer_keys(self, group=None):
        """Map the key buffers of every rank of `group` (one node) into this engine through CUDA IPC, for
        TriangleRaster.render_color_composite (sort-last composite over NVLink peer memory).  Collective."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        mine = (C.c_uint8 * 64)()
        _lib.check(_lib.lib().tina_engine_ipc_export(self._h, mine))
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.keys.device)
        allh = torch.empty(world * 64, dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(allh, t, group=group)
        buf = (C.c_uint8 * (world * 64))(*allh.cpu().tolist())
        _lib.check(_lib.lib().tina_engine_ipc_open_peers(self._h, buf, world, rank))
        self._peers = world

    def set_peer_keys(self, engines, rank):
        """Same table from engines of this process (e.g. several engines on one GPU): engines[rank] is self."""
        ptrs = (C.c_void_p * len(engines))(*[e.keys.data_ptr() for e in engines])
        _lib.check(_lib.lib().tina_engine_set_peer_keys(self._h, ptrs, len(engines), int(rank)))
        self._peers = len(engines)
        self._peer_refs = list(engines)

    def close_peer_keys(self):
        _lib.check(_lib.lib().tina_engine_ipc_close_peers(self._h))
        self._peers = 0

    @property
    def face_base(self):
        v = C.c_uint32()
        _lib.check(_lib.lib().tina_engine_get_face_base(self._h, C.byref(v)))
        return int(v.value)

    def randomize_bias(self, center=False):
        """engine.py:31-39 (TAA jitter)."""
        self.bias[None] = [0.5, 0.5] if center else np.random.rand(2).astype(np.float32)

    # host-side numpy versions of the @ti.func helpers (engine.py:52-65), f32 like the device code
    def to_viewspace(self, p):
        M = self.W2V.to_numpy()
        p = np.asarray(p, dtype=np.float32)
        r = (M[:3, :3] @ p + M[:3, 3]).astype(np.float32)
        return r / np.float32(M[3, :3] @ p + M[3, 3])

    def from_viewspace(self, p):
        M = self.V2W.to_numpy()
        p = np.asarray(p, dtype=np.float32)
        r = (M[:3, :3] @ p + M[:3, 3]).astype(np.float32)
        return r / np.float32(M[3, :3] @ p + M[3, 3])

    def to_viewport(self, p):
        return (np.asarray(p, dtype=np.float32)[:2] * 0.5 + 0.5) * np.asarray(self.res, dtype=np.float32)

    def from_viewport(self, p):
        return np.asarray(p, dtype=np.float32) / np.asarray(self.res, dtype=np.float32) * 2 - 1
