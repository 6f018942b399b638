"""Pixel sinks (reference tina/core/shader.py).  `Shader` = material + lighting -> img
(shader.py:112-135); `ShaderGroup` fans one render_color out to several sinks (:138-148)."""


class IShader:
    def __init__(self, img):
        self.img = img

    def clear_buffer(self):
        self.img.fill(0)


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
