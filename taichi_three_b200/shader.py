"""Pixel sinks (reference tina/core/shader.py).  `Shader` = material + lighting -> img
(shader.py:112-135); `ShaderGroup` fans one render_color out to several sinks (:138-148)."""


class IShader:
    def __init__(self, img):
        self.img = img

    def clear_buffer(self):
        self.img.fill(0)


class _Sink(IShader):
    """G-buffer sinks (core/shader.py:21-109): write one attribute of the visible surface into `img`
    (a Field / CUDA tensor [W, H] or [W, H, n], float32 or int32)."""
    kind = None

    def param(self):
        return None


class ConstShader(_Sink):  # shader.py:21-28
    kind = 0

    def __init__(self, img, value):
        super().__init__(img)
        self.value = value

    def param(self):
        import numpy as np
        v = np.asarray(self.value, dtype=np.float32).reshape(-1)
        return np.resize(v, 3) if v.size < 3 else v[:3]


class PositionShader(_Sink):  # shader.py:31-34
    kind = 1


class DepthShader(_Sink):  # shader.py:37-40
    kind = 2


class NormalShader(_Sink):  # shader.py:43-46
    kind = 3


class ViewNormalShader(_Sink):  # shader.py:49-56
    kind = 4


class TexcoordShader(_Sink):  # shader.py:59-62
    kind = 5


class ColorShader(_Sink):  # shader.py:65-68
    kind = 6


class ChessboardShader(_Sink):  # shader.py:71-79
    kind = 7

    def __init__(self, img, size=8):
        super().__init__(img)
        self.size = size

    def param(self):
        import numpy as np
        return np.float32([self.size, 0, 0])


class ViewdirShader(_Sink):  # shader.py:96-101
    kind = 8


class SimpleShader(_Sink):  # shader.py:104-109
    kind = 9


class _ElmidSink(_Sink):  # probe.py:21-22
    kind = 10


class ProbeShader:
    """tina/probe.py:5-29: per-pixel id of the visible face (`elmid`, -1 = none) and its interpolated texture
    coordinate (`texcoord`), for picking / painting (`scene.post_shaders.append(probe)` before the objects are added).
    Face ids are those of the object rendered last at that pixel, like the reference's (its `f`)."""

    def __init__(self, res):
        import torch
        from .field import Field
        res = (res, res) if isinstance(res, int) else tuple(res)
        self.res = (int(res[0]), int(res[1]))
        dev = torch.device('cuda', torch.cuda.current_device())
        self.elmid = Field(torch.full(self.res, -1, dtype=torch.int32, device=dev))
        self.texcoord = Field(torch.zeros(self.res + (2,), dtype=torch.float32, device=dev))
        self._parts = (_ElmidSink(self.elmid), TexcoordShader(self.texcoord))

    def clear_buffer(self):  # probe.py:11-15
        self.elmid.fill(-1)
        self.texcoord.fill(0)

    def _sinks(self):
        return self._parts

    def touch(self, callback, mx, my, rad):
        """probe.py:25-35: call `callback(probe, (x, y), r)` for every pixel within `rad` of the cursor (mx, my in
        [0, 1]) that shows a face.  The reference's callback is a Taichi function; here it is a Python callable run
        on the host over the (small) disc."""
        import numpy as np
        p = np.float32([mx, my]) * np.float32(self.res)
        bot = np.floor(p - np.float32(rad)).astype(int)
        top = np.ceil(p + np.float32(rad)).astype(int)
        x0, y0 = max(int(bot[0]), 0), max(int(bot[1]), 0)
        x1, y1 = min(int(top[0]), self.res[0] - 1), min(int(top[1]), self.res[1] - 1)
        if x1 < x0 or y1 < y0:
            return
        ids = self.elmid.to_torch()[x0:x1 + 1, y0:y1 + 1].cpu().numpy()
        for x in range(x0, x1 + 1):
            for y in range(y0, y1 + 1):
                r = float(np.sqrt(np.float32((x - p[0]) ** 2 + (y - p[1]) ** 2)))
                if r > rad or ids[x - x0, y - y0] == -1:
                    continue
                callback(self, (x, y), r)


class Shader(IShader):
    def __init__(self, img, lighting, material):
        super().__init__(img)
        self.lighting = lighting
        self.material = material


class ShaderGroup(IShader):
    def __init__(self, shaders=()):
        self.shaders = list(shaders)

    def clear_buffer(self):
        for s in self.shaders:
            s.clear_buffer()
