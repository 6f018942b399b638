"""Lighting: up to 16 directional / point lights + ambient (reference tina/core/lighting.py:25-98).
State is host-side; it travels to the shading kernel as a TinaLighting kernel parameter."""
import numpy as np

from . import _lib
from .field import HostField


class Lighting:
    def __init__(self, maxlights=16):
        if maxlights > _lib.TINA_MAX_LIGHTS:
            raise ValueError(f'maxlights > {_lib.TINA_MAX_LIGHTS} is not supported')
        self.maxlights = maxlights
        # lighting.py:33-38 defaults
        self.light_dirs = np.tile(np.array([0, 0, 1, 0], dtype=np.float32), (maxlights, 1))
        self.light_colors = np.ones((maxlights, 3), dtype=np.float32)
        self.ambient_color = HostField(np.zeros(3, dtype=np.float32))
        self.nlights = HostField(np.array(0, dtype=np.int32))

    def set_lights(self, light_dirs):  # lighting.py:40-44
        self.nlights[None] = len(light_dirs)
        for i, (dir, color) in enumerate(light_dirs):
            self.light_dirs[i] = dir
            self.light_colors[i] = color

    def clear_lights(self):
        self.nlights[None] = 0

    def add_light(self, dir=(0, 0, 1), pos=None, color=(1, 1, 1)):  # lighting.py:49-53
        i = int(self.nlights[None])
        if i >= self.maxlights:
            raise ValueError(f'too many lights (max {self.maxlights})')
        self.nlights[None] = i + 1
        self.set_light(i, dir, pos, color)
        return i

    def set_light(self, i, dir=(0, 0, 1), pos=None, color=(1, 1, 1)):  # lighting.py:55-66
        if pos is not None:
            d, w = np.array(pos, dtype=np.float64), 1
        else:
            d = np.array(dir, dtype=np.float64)
            d, w = d / np.linalg.norm(d), 0
        self.light_dirs[i] = np.append(d, w)
        self.light_colors[i] = np.array(color, dtype=np.float64)

    def set_ambient_light(self, color):  # lighting.py:68-69
        self.ambient_color[None] = np.array(color, dtype=np.float32)

    def struct_ref(self):
        """ctypes reference to the current TinaLighting POD (what the C ABI takes)."""
        L = self.struct()
        ref = self.__dict__.get('_cache_ref')
        if ref is None or ref[0] is not L:
            import ctypes
            ref = (L, ctypes.byref(L))
            self._cache_ref = ref
        return ref[1]

    def struct(self):
        """TinaLighting POD, rebuilt only when the light state changed."""
        key = (self.light_dirs.tobytes(), self.light_colors.tobytes(), self.ambient_color.to_numpy().tobytes(),
               int(self.nlights[None]))
        if getattr(self, '_cache_key', None) == key:
            return self._cache_struct
        L = self._build_struct()
        self._cache_key, self._cache_struct = key, L
        return L

    def _build_struct(self):
        L = _lib.TinaLighting()
        n = int(self.nlights[None])
        L.nlights = n
        amb = self.ambient_color.to_numpy()
        for k in range(3):
            L.ambient[k] = float(amb[k])
        for i in range(n):
            for k in range(4):
                L.dirs[i][k] = float(self.light_dirs[i, k])
            for k in range(3):
                L.colors[i][k] = float(self.light_colors[i, k])
        return L
