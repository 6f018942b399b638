"""taichi_three_b200: B200-native (sm_100a) triangle-raster hot path behind the Python API of
taichi-dev/taichi_three ("Tina").  Flat namespace like the reference's lazy `tina.X` lookup
(tina/lazimp.py:69-143).  No Taichi, no Triton, no CPU fallback: kernels live in
csrc/libtina_b200.so (C ABI: include/tina_b200.h)."""
__version__ = (0, 1, 1)

from .matrix import (identity, affine, lookat, ortho, frustum, orthogonal, perspective, scale, translate, quaternion,
                     eularXYZ, euler_matrix, orbit_camera)
from .material import (Node, Const, Param, Input, Texture, ChessboardTexture, LerpTexture, FresnelFactor, IMaterial, MixMaterial, ScaleMaterial,
                       AddMaterial, Lambert, Phong, CookTorrance, Emission, Classic, Diffuse, Lamp, PBR,
                       flatten_material)
from .assimp import readobj, readgltf, writeobj, pfmwrite, objverts, objnorms, objcoors, objorient, objautoscale
from .lighting import Lighting
from .shader import (IShader, Shader, ShaderGroup, ConstShader, PositionShader, DepthShader, NormalShader,
                     ViewNormalShader, TexcoordShader, ColorShader, ChessboardShader, ViewdirShader, SimpleShader, ProbeShader)
from .mesh import (MAX, SimpleMesh, PrimitiveMesh, ConnectiveMesh, MeshModel, MeshGrid, MeshTransform, MeshFlipCulling, MeshNoCulling,
                   MeshFlipNormal, MeshFlatNormal, MeshSmoothNormal, MeshEditBase)
from .engine import Engine
from .triangle import TriangleRaster
from .particle import ParticleRaster, SimpleParticles, ParsTransform
from .wireframe import WireframeRaster, MeshToWire
from .postp import FXAA, Blooming, SSAO, SSR
from .scene import Scene
from .control import Control, RotationStep
from .field import Field
from .graph import FrameGraph
from . import multigpu
