"""Scene: frame orchestration for the raster pipeline (reference tina/scene/raster.py:5-258),
restricted to the triangle-raster path: objects -> set_object / render_occup / render_color,
default light + ambient, ACES tonemap, and the image-space passes ssao / ssr / blooming / fxaa / taa in the reference's
order.  Options that enable other subsystems (ibl) raise."""
import numpy as np
import torch

from . import _lib
from .engine import Engine, _stream
from .field import Field
from .lighting import Lighting
from .material import Diffuse
from .shader import Shader, ShaderGroup
from .triangle import TriangleRaster

import ctypes as C


class _ObjInfo:
    def __init__(self, material, raster):
        self.material, self.raster = material, raster


class Accumator:
    """TAA accumulation buffer (reference tina/util/accumator.py:5-23)."""

    def __init__(self, res, device):
        self.img = Field(torch.zeros((res[0], res[1], 3), dtype=torch.float32, device=device))
        self.count = [0]

    def clear(self):
        self.count[0] = 0
        self.img.to_torch().zero_()

    def update(self, src):
        self.count[0] += 1
        a, s = self.img.to_torch(), src.to_torch()
        _lib.check(_lib.lib().tina_image_accumulate(C.c_void_p(a.data_ptr()), C.c_void_p(s.data_ptr()), a.numel(),
                                                    self.count[0], _stream()))


class MaterialTable:
    """The scene's materials by id (matr/material.py:643-656); SSR looks a pixel's material up by its mtlid."""

    def __init__(self):
        self.materials = []

    def clear_materials(self):
        self.materials.clear()

    def add_material(self, matr):
        self.materials.append(matr)


class Scene:
    UNSUPPORTED = ('ibl',)

    def __init__(self, res_x=512, res_y=None, **options):
        self.engine = Engine(res_x, res_y)
        self.res = self.engine.res
        self.options = options
        for key in self.UNSUPPORTED:
            if options.get(key, False):
                raise NotImplementedError(f'Scene option {key}=True is outside the B200 triangle-raster path')
        self.taa = options.get('taa', False)
        self.tonemap = options.get('tonemap', True)
        self.bgcolor = options.get('bgcolor', 0)
        self.lighting = Lighting()
        self.image = Field(torch.zeros((self.res[0], self.res[1], 3), dtype=torch.float32, device=self.engine.device))
        self.default_material = Diffuse()
        self.post_shaders = []
        self.pre_shaders = []
        self.materials = []
        self.shaders = {}
        self.objects = {}
        self.pp_img = self.image
        self.blooming = options.get('blooming', False)
        self.fxaa = options.get('fxaa', False)
        if self.blooming:  # raster.py:74-75
            from .postp import Blooming
            self.blooming = Blooming(self.res)
        if self.fxaa:  # raster.py:82-83
            from .postp import FXAA
            self.fxaa = FXAA(self.res)
        self.ssao = options.get('ssao', False)
        self.ssr = options.get('ssr', False)
        dev = self.engine.device
        if self.ssr:  # raster.py:43-49: the materials by id, for SSR's material.sample()
            self.mtltab = MaterialTable()
        if self.ssao or self.ssr:  # raster.py:51-54: world-normal G-buffer as a pre-shader
            from .shader import NormalShader
            self.norm_buffer = Field(torch.zeros((self.res[0], self.res[1], 3), dtype=torch.float32, device=dev))
            self.norm_shader = NormalShader(self.norm_buffer)
            self.pre_shaders.append(self.norm_shader)
        if self.ssr:  # raster.py:56-64: material ids (a ConstShader per material) and, with texturing, texture coordinates
            from .shader import TexcoordShader
            self.mtlid_buffer = Field(torch.zeros(self.res, dtype=torch.int32, device=dev))
            self.coor_buffer = None
            if 'texturing' in options:
                self.coor_buffer = Field(torch.zeros((self.res[0], self.res[1], 2), dtype=torch.float32, device=dev))
                self.coor_shader = TexcoordShader(self.coor_buffer)
                self.pre_shaders.append(self.coor_shader)
        if self.ssao:  # raster.py:65-66
            from .postp import SSAO
            self.ssao = SSAO(self.res, self.norm_buffer, taa=self.taa)
        if self.ssr:  # raster.py:68-70
            from .postp import SSR
            self.ssr = SSR(self.res, self.norm_buffer, self.coor_buffer, self.mtlid_buffer, self.mtltab, taa=self.taa)
        if self.taa:
            self.accum = Accumator(self.res, self.engine.device)
        # raster.py:90-93
        self.lighting.add_light(dir=[1, 2, 3], color=[0.9, 0.9, 0.9])
        self.lighting.set_ambient_light([0.1, 0.1, 0.1])

    def _ensure_material_shader(self, material):  # raster.py:95-110
        if any(material is m for m in self.materials):
            return
        shader = Shader(self.image, self.lighting, material)
        base_shaders = [shader]
        if self.ssr:  # raster.py:102-105 (the table itself is filled by a materialize callback there, :45-49)
            from .shader import ConstShader
            base_shaders.append(ConstShader(self.mtlid_buffer, len(self.materials)))
            self.mtltab.add_material(material)
        self.materials.append(material)
        self.shaders[id(material)] = ShaderGroup(self.pre_shaders + base_shaders + self.post_shaders)

    def add_object(self, object, material=None, raster=None):  # raster.py:112-146
        assert id(object) not in self.objects
        if material is None:
            material = self.default_material
        if raster is None:
            if hasattr(object, 'get_nfaces') and hasattr(object, 'get_npolygon') and object.get_npolygon() == 2:
                if not hasattr(self, 'wireframe_raster'):  # raster.py:124-128
                    from .wireframe import WireframeRaster
                    opts = {k: v for k, v in self.options.items() if k in ('maxwires', 'linewidth', 'linecolor', 'clipping')}
                    self.wireframe_raster = WireframeRaster(self.engine, **opts)
                raster = self.wireframe_raster
            elif hasattr(object, 'get_nfaces'):
                if not hasattr(self, 'triangle_raster'):
                    opts = {k: v for k, v in self.options.items() if k in ('maxfaces', 'smoothing', 'texturing', 'culling', 'clipping')}
                    self.triangle_raster = TriangleRaster(self.engine, **opts)
                raster = self.triangle_raster
            elif hasattr(object, 'get_npars'):  # raster.py:133-136
                if not hasattr(self, 'particle_raster'):
                    from .particle import ParticleRaster
                    opts = {k: v for k, v in self.options.items() if k in ('maxpars', 'coloring', 'clipping')}
                    self.particle_raster = ParticleRaster(self.engine, **opts)
                raster = self.particle_raster
            elif hasattr(object, 'sample_volume'):
                raise NotImplementedError('the volume rasteriser is outside the B200 raster path')
            else:
                raise ValueError(f'cannot determine raster type of object: {object}')
        self._ensure_material_shader(material)
        self.objects[id(object)] = (object, _ObjInfo(material, raster))

    def init_control(self, gui, center=None, theta=None, phi=None, radius=None, fov=60, is_ortho=False, blendish=True):
        from .control import Control
        self.control = Control(gui, fov=fov, is_ortho=is_ortho, blendish=blendish)
        if center is not None:
            self.control.center[:] = center
        self.control.init_rot(theta, phi)
        if radius is not None:
            self.control.radius = radius

    def render(self):
        """raster.py:168-207.  image.fill(bg) is fused into the first object's shading pass and, for
        single-object scenes, so is the ACES tonemap; results are identical to the separate passes."""
        if self.taa:  # raster.py:173-174: jittered sample position, centred on the first frame
            self.engine.randomize_bias(self.accum.count[0] == 0)
        self.engine.clear_depth()
        for s in self.pre_shaders + self.post_shaders:
            s.clear_buffer()
        items = list(self.objects.values())
        bg = np.broadcast_to(np.asarray(self.bgcolor, dtype=np.float32), (3,))
        if not items:
            self.image.fill(bg)
        def fuses(raster):  # rasterisers whose render_color takes the fusion hints
            return isinstance(raster, TriangleRaster)
        # Frame glue folded into the shading passes (results identical to the separate passes):
        #  * image.fill(bg) rides along with the FIRST object's pass (every rasteriser that takes fill_bg);
        #  * the ACES curve and the TAA accumulation ride along with the LAST object's pass when that is a triangle pass
        #    and nothing sits between shading and tonemap (no SSAO / SSR / blooming): it finishes the pixels it does not own too.
        post_free = not self.blooming and not self.ssao and not self.ssr and getattr(self, 'fuse_glue', True)  # (fuse_glue = False: separate passes)
        last_fuses = bool(items) and fuses(items[-1][1].raster) and post_free
        fuse_tm = bool(self.tonemap) and (last_fuses or (len(items) == 1 and post_free and hasattr(items[0][1].raster, 'set_particles')))
        fuse_acc = bool(self.taa) and last_fuses and not self.fxaa
        if fuse_acc:
            self.accum.count[0] += 1
        for i, (obj, info) in enumerate(items):
            shader = self.shaders[id(info.material)]
            info.raster.set_object(obj)
            info.raster.render_occup()
            last = i == len(items) - 1
            if fuses(info.raster):
                info.raster.render_color(shader, fill_bg=bg if i == 0 else None, tonemap=fuse_tm and last, finish=last and (fuse_tm or fuse_acc),
                                         accum=(self.accum.img.to_torch(), self.accum.count[0]) if (fuse_acc and last) else None)
            elif hasattr(info.raster, 'set_particles'):
                info.raster.render_color(shader, fill_bg=bg if i == 0 else None, tonemap=fuse_tm)
            else:
                if i == 0:
                    self.image.fill(bg)
                info.raster.render_color(shader)
        if self.ssao:  # raster.py:189-191
            self.ssao.render(self.engine)
            self.ssao.apply(self.image)
        if self.ssr:  # raster.py:192-194
            self.ssr.render(self.engine, self.image)
            self.ssr.apply(self.image)
        if self.blooming:  # raster.py:200-201
            self.blooming.apply(self.image)
        if self.tonemap and not fuse_tm:
            t = self.image.to_torch()
            _lib.check(_lib.lib().tina_image_tonemap(C.c_void_p(t.data_ptr()), t.numel(), _stream()))
        if self.fxaa:  # raster.py:204-205
            self.fxaa.apply(self.image)
        if self.taa and not fuse_acc:  # raster.py:206-207
            self.accum.update(self.pp_img)

    @property
    def img(self):
        return self.accum.img if self.taa else self.pp_img

    def input(self, gui):  # raster.py:216-228
        if not hasattr(self, 'control'):
            from .control import Control
            self.control = Control(gui)
        changed = self.control.apply_camera(self.engine)
        if changed:
            self.clear()
        return changed

    def clear(self):  # raster.py:230-232
        if self.taa:
            self.accum.clear()

    def load_gltf(self, path):  # raster.py:234-241
        from .assimp import readgltf
        return readgltf(path).extract(self)
