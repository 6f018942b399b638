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
