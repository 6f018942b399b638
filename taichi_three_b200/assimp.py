"""Host-side asset loaders feeding the raster path: Wavefront OBJ and glTF 2.0 (embedded
buffers), behaviourally matching the reference's tina/assimp/obj.py:17-102 and
tina/assimp/gltf.py:6-247 including their quirks: polygons are fanned
the reference's way, missing OBJ indices become 0, accessor byteOffset / byteStride and node
hierarchies are ignored, images are stored [x][y] without a vertical flip.
"""
import base64
import io
import json

import numpy as np


def _fan(poly):
    """obj.py:4-14: quad -> (0,1,2),(2,3,0); n-gon -> fan around vertex 0."""
    if len(poly) == 3:
        return [poly]
    if len(poly) == 4:
        return [[poly[0], poly[1], poly[2]], [poly[2], poly[3], poly[0]]]
    assert len(poly) > 4, len(poly)
    return [[poly[0], poly[k], poly[k + 1]] for k in range(1, len(poly) - 1)]


def readobj(path, orient='xyz', scale=None, simple=False, usemtl=True, quadok=False):
    if callable(getattr(path, 'read', None)):
        lines = path.readlines()
    else:
        with open(path, 'rb') as fh:
            lines = fh.readlines()
    attrs = {b'v': [], b'vt': [], b'vn': []}
    faces, usemtls, mtllib = [], [], None
    for raw in lines:
        parts = raw.strip().split(maxsplit=1)
        if len(parts) != 2:
            continue
        tag, rest = parts
        if tag in attrs:
            try:
                attrs[tag].append([float(t) for t in rest.split()])
            except ValueError:
                pass
    for raw in lines:
        parts = raw.strip().split(maxsplit=1)
        if len(parts) != 2:
            continue
        tag, rest = parts
        fields = rest.split()
        if tag == b'mtllib':
            mtllib = fields[0]
        elif tag == b'usemtl':
            usemtls.append([len(faces), fields[0]])
        elif tag == b'f':
            poly = [[int(t) - 1 if t else 0 for t in field.split(b'/')] for field in fields]
            if quadok:
                faces.append(poly)
            else:
                faces.extend(_fan(poly))

    def arr(rows, width):
        return np.array(rows, dtype=np.float32) if rows else np.zeros((1, width), dtype=np.float32)

    obj = {'v': arr(attrs[b'v'], 3), 'vt': arr(attrs[b'vt'], 2), 'vn': arr(attrs[b'vn'], 3),
           'f': np.array(faces, dtype=np.int32) if faces else np.zeros((1, 3, 3), dtype=np.int32)}
    if usemtl:
        obj['usemtl'], obj['mtllib'] = usemtls, mtllib
    if orient is not None:
        objorient(obj, orient)
    if scale is not None:
        if scale == 'auto':
            objautoscale(obj)
        else:
            obj['v'] *= scale
    if simple:
        return obj['v'], obj['f'][:, :, 0]
    return obj


def writeobj(path, obj):
    """assimp/obj.py:103-125: `v` / `vt` / `vn` lines with Python's shortest repr of every coordinate, faces 1-based as
    v/vt/vn triples ([N, P, 3] index arrays) or the vertex index three times ([N, P]).  `path`: a file name or a writable
    text stream."""
    import numpy as np
    lines = ['# OBJ file saved by tina.writeobj', '# https://github.com/taichi-dev/taichi_three']
    for tag in ('v', 'vt', 'vn'):
        if tag in obj:
            lines += [tag + ' ' + ' '.join(str(x) for x in row) for row in np.asarray(obj[tag])]
    if 'f' in obj:
        faces = np.asarray(obj['f'])
        if faces.ndim >= 3:
            lines += ['f ' + ' '.join('/'.join(str(i + 1) for i in corner) for corner in face) for face in faces]
        else:
            lines += ['f ' + ' '.join('/'.join([str(i + 1)] * 3) for i in face) for face in faces]
    text = '\n'.join(lines) + '\n'
    if callable(getattr(path, 'write', None)):
        path.write(text)
    else:
        with open(path, 'w') as fh:
            fh.write(text)


def pfmwrite(path, im):
    """assimp/pfm.py:4-12: a [W, H(, 3)] float image (x-major like every field here) as a PFM file, rows swapped to y-major,
    values divided by the largest magnitude, which travels as the (negative = little-endian) scale of the header."""
    import sys
    import numpy as np
    im = np.asarray(im.to_numpy() if hasattr(im, 'to_numpy') else im).swapaxes(0, 1)
    scale = max(1e-10, float(-im.min()), float(im.max()))
    h, w = im.shape[:2]
    with open(path, 'wb') as fh:
        fh.write(b'PF\n' if im.ndim >= 3 else b'Pf\n')
        fh.write(f'{w} {h}\n'.encode())
        fh.write(f'{scale if sys.byteorder == "big" else -scale}\n'.encode())
        fh.write((im / scale).astype(np.float32).tobytes())


def objautoscale(obj):  # obj.py:179-181
    obj['v'] -= np.average(obj['v'], axis=0)
    obj['v'] /= np.max(np.abs(obj['v']))


def objorient(obj, orient):  # obj.py:184-203
    flip = orient.startswith('-')
    if flip:
        orient = orient[1:]
    perm = ['xyz'.index(o.lower()) for o in orient]
    neg = [o.isupper() for o in orient]
    if perm != [0, 1, 2]:
        obj['v'] = np.ascontiguousarray(obj['v'][:, perm])
        obj['vn'] = np.ascontiguousarray(obj['vn'][:, perm])
    for i, n in enumerate(neg):
        if n:
            obj['v'][:, i] = -obj['v'][:, i]
            obj['vn'][:, i] = -obj['vn'][:, i]
    if flip:
        obj['f'] = np.ascontiguousarray(obj['f'][:, ::-1, :])


def objverts(obj):
    return obj['v'][obj['f'][:, :, 0]]


def objnorms(obj):
    return obj['vn'][obj['f'][:, :, 2]]


def objcoors(obj):
    return obj['vt'][obj['f'][:, :, 1]]


# ---------------------------------------------------------------------------------------
# glTF
# ---------------------------------------------------------------------------------------
_COMPONENT = {0x1400: 'b', 0x1401: 'B', 0x1402: 'h', 0x1403: 'H', 0x1404: 'i', 0x1405: 'I', 0x1406: 'f', 0x140A: 'd'}
_VECTOR = {'SCALAR': '', 'VEC2': '2', 'VEC3': '3', 'VEC4': '4'}


class GltfPrimitive:
    def __init__(self, obj, material):
        self.obj, self.material = obj, material


class GltfNode:
    def __init__(self, name, trans, primitives):
        self.name, self.trans, self.primitives = name, trans, primitives


class GltfScene:
    """Parsed glTF scene; `extract(scene)` adds one MeshTransform(MeshModel) + PBR material per primitive
    (gltf.py:55-62,79-92,118-147)."""

    def __init__(self, name, nodes, images):
        self.name, self.nodes, self.images = name, nodes, images

    def extract(self, scene):
        from . import mesh as M
        for node in self.nodes:
            for prim in node.primitives:
                m = M.MeshTransform(M.MeshModel(prim.obj), node.trans)
                scene.add_object(m, self._material(prim.material))
        return scene

    def _material(self, pbr):
        from . import material as mt
        if pbr is None:
            return None
        if pbr is False:  # glTF material without pbrMetallicRoughness (gltf.py:119-120)
            return mt.Lambert()
        kwargs = {}
        for key, value in pbr.items():
            if key == 'baseColorFactor':
                kwargs['basecolor'] = value[:3]
            elif key == 'baseColorTexture':
                kwargs['basecolor'] = mt.Texture(self.images[value['index']])
            elif key == 'metallicFactor':
                kwargs['metallic'] = value
            elif key == 'metallicTexture':
                kwargs['metallic'] = mt.Texture(self.images[value['index']])
            elif key == 'roughnessFactor':
                kwargs['roughness'] = value
            elif key == 'roughnessTexture':
                kwargs['roughness'] = mt.Texture(self.images[value['index']])
            elif key == 'metallicRoughnessTexture':
                img = self.images[value['index']]
                kwargs['metallic'] = mt.Texture(img[..., 2])
                kwargs['roughness'] = mt.Texture(img[..., 1])
        return mt.PBR(**kwargs)


def readgltf(path):
    from . import matrix as mx
    if isinstance(path, str):
        with open(path, 'rb') as fh:
            root = json.load(fh)
    elif isinstance(path, dict):
        root = path
    else:
        root = json.load(path)

    def load_uri(uri):
        if uri.startswith('data:'):
            return base64.b64decode(uri[uri.index('base64,') + len('base64,'):].encode('ascii'))
        with open(uri, 'rb') as fh:
            return fh.read()

    buffers = [load_uri(b['uri']) for b in root['buffers']]

    def view_bytes(i):
        bv = root['bufferViews'][i]
        off = bv['byteOffset']  # sic: required by the reference (gltf.py:167)
        return buffers[bv['buffer']][off:off + bv['byteLength']]

    def accessor(i):
        acc = root['accessors'][i]
        dtype = _VECTOR[acc['type']] + _COMPONENT[acc['componentType']]
        return np.frombuffer(view_bytes(acc['bufferView']), dtype=dtype, count=acc['count'])

    images = []
    for image in root.get('images', []):
        data = view_bytes(image['bufferView']) if 'bufferView' in image else load_uri(image['uri'])
        from PIL import Image
        with io.BytesIO(data) as fh:
            im = np.array(Image.open(fh))
        images.append(np.swapaxes(im, 0, 1))  # gltf.py:199-200: [x][y], no flip

    materials = []
    for material in root.get('materials', []):
        materials.append(material['pbrMetallicRoughness'] if 'pbrMetallicRoughness' in material else False)

    scene = root['scenes'][root.get('scene', 0)]
    nodes = []
    for node_id in scene['nodes']:
        node = root['nodes'][node_id]
        trans = mx.identity()  # gltf.py:79-92: T @ R @ S
        if node.get('scale') is not None:
            trans = mx.scale(node['scale']) @ trans
        if node.get('rotation') is not None:
            trans = mx.quaternion(node['rotation']) @ trans
        if node.get('translation') is not None:
            trans = mx.translate(node['translation']) @ trans
        prims = []
        if 'mesh' in node:
            for primitive in root['meshes'][node['mesh']]['primitives']:
                obj = {}
                if 'indices' in primitive:
                    idx = accessor(primitive['indices'])
                    obj['f'] = idx.reshape(len(idx) // 3, 3)
                for name, acc_id in primitive['attributes'].items():
                    a = accessor(acc_id)
                    if name == 'POSITION':
                        obj['v'] = a
                    elif name == 'NORMAL':
                        obj['vn'] = a
                    elif name.startswith('TEXCOORD') and 'vt' not in obj:
                        obj['vt'] = a
                mat = materials[primitive['material']] if 'material' in primitive else None
                prims.append(GltfPrimitive(obj, mat))
        nodes.append(GltfNode(node.get('name', 'Untitled'), trans, prims))
    return GltfScene(scene.get('name', 'Untitled'), nodes, images)
