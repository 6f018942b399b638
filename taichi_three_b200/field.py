"""Small stand-ins for the Taichi fields the reference exposes on its objects
(`engine.depth`, `raster.occup`, `scene.img`, `engine.W2V` ...), so scripts that call
`.to_numpy()`, `.from_numpy()`, `.fill()` or `field[None] = x` keep working.
Device fields wrap a CUDA torch.Tensor (zero-copy `to_torch()`); host fields wrap numpy.
"""
import numpy as np
import torch


class RawCudaBuffer:
    """`__cuda_array_interface__` view of memory owned by libtina_b200 (torch.as_tensor wraps it zero-copy)."""

    def __init__(self, ptr, shape, typestr, owner=None):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}
        self.owner = owner


def wrap_device(ptr, shape, dtype, device, owner=None):
    typestr = {torch.float32: '<f4', torch.int32: '<i4', torch.int64: '<i8'}[dtype]
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return torch.empty(shape, dtype=dtype, device=device)
    t = torch.as_tensor(RawCudaBuffer(ptr, shape, typestr, owner), device=device)
    t._tina_owner = owner
    return t


class Field:
    """Device field: x-major like `ti.field(dtype, (W, H))` (reference triangle.py:16)."""

    def __init__(self, tensor):
        self._t = tensor

    @property
    def shape(self):
        t = self._t
        return tuple(t.shape[:2]) if t.dim() >= 2 else tuple(t.shape)

    @property
    def dtype(self):
        return self._t.dtype

    def to_torch(self):
        return self._t

    def to_numpy(self):
        return self._t.detach().cpu().numpy()

    def from_numpy(self, arr):
        self._t.copy_(torch.as_tensor(np.ascontiguousarray(arr)).to(self._t.dtype).reshape(self._t.shape))

    def from_torch(self, t):
        self._t.copy_(t.reshape(self._t.shape))

    def fill(self, value):
        if isinstance(value, (list, tuple, np.ndarray)):
            self._t.copy_(torch.as_tensor(np.asarray(value, dtype=np.float32), device=self._t.device).expand_as(self._t))
        else:
            self._t.fill_(value)

    def copy_from(self, other):
        self._t.copy_(other.to_torch() if hasattr(other, 'to_torch') else torch.as_tensor(other))

    def __getitem__(self, idx):
        v = self._t[idx]
        return v.item() if v.dim() == 0 else v.cpu().numpy()

    def __setitem__(self, idx, value):
        self._t[idx] = torch.as_tensor(np.asarray(value), device=self._t.device).to(self._t.dtype)

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a.astype(dtype) if dtype is not None else a


class HostField:
    """0-d field kept on the host (`engine.W2V[None]`, `engine.bias[None]`, `lighting.nlights[None]`)."""

    def __init__(self, value, on_change=None):
        self._v = np.array(value)
        self._on_change = on_change

    def __getitem__(self, idx):
        if idx is None or idx == ():
            return self._v.copy() if self._v.ndim else self._v.item()
        return self._v[idx]

    def __setitem__(self, idx, value):
        if idx is None or idx == ():
            self._v[...] = np.asarray(value, dtype=self._v.dtype).reshape(self._v.shape)
        else:
            self._v[idx] = value
        if self._on_change:
            self._on_change()

    def to_numpy(self):
        return self._v.copy()

    def from_numpy(self, arr):
        self[None] = arr
