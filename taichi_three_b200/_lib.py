"""ctypes binding of the C ABI in include/tina_b200.h (libtina_b200.so).

There is NO fallback: if the CUDA library is missing or a call fails this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TINA_B200_LIB') or os.path.join(_HERE, 'csrc', 'libtina_b200.so')  # env: kernel-variant experiments

TINA_SMOOTHING, TINA_TEXTURING, TINA_CULLING, TINA_CLIPPING = 1, 2, 4, 8
TINA_COLOR_TONEMAP, TINA_COLOR_FILL_BG, TINA_COLOR_FINISH = 1, 2, 4
TINA_MAX_LIGHTS, TINA_MAX_INSTR, TINA_MAX_TEX = 16, 96, 4

(OP_CONST, OP_INPUT, OP_TEXTURE, OP_FRESNEL, OP_LAMBERT, OP_PHONG, OP_COOK, OP_MIX, OP_MUL,
 OP_ADD, OP_REG, OP_STORE, OP_BCAST, OP_CHESS) = range(14)
TINA_MAX_REGS = 8
OP3 = 0x100  # three-address prologue form (include/tina_b200.h)
TINA_VM_VALUES = 16
(SINK_CONST, SINK_POSITION, SINK_DEPTH, SINK_NORMAL, SINK_VIEWNORMAL, SINK_TEXCOORD, SINK_COLOR, SINK_CHESSBOARD,
 SINK_VIEWDIR, SINK_SIMPLE, SINK_ELMID) = range(11)
TINA_MAX_SINKS = 8


class TinaLighting(C.Structure):
    _fields_ = [('dirs', (C.c_float * 4) * TINA_MAX_LIGHTS),
                ('colors', (C.c_float * 4) * TINA_MAX_LIGHTS),
                ('ambient', C.c_float * 4),
                ('nlights', C.c_int32),
                ('pad', C.c_int32 * 3)]


class TinaInstr(C.Structure):
    _fields_ = [('op', C.c_int32), ('arg', C.c_int32), ('c', C.c_float * 3)]


class TinaMaterial(C.Structure):
    _fields_ = [('n_brdf', C.c_int32), ('n_ambient', C.c_int32), ('n_emission', C.c_int32), ('ntex', C.c_int32),
                ('tex', C.c_void_p * TINA_MAX_TEX),
                ('tex_w', C.c_int32 * TINA_MAX_TEX), ('tex_h', C.c_int32 * TINA_MAX_TEX),
                ('tex_c', C.c_int32 * TINA_MAX_TEX),
                ('n_prologue', C.c_int32), ('prologue_form', C.c_int32), ('pad_', C.c_int32 * 2),
                ('code', TinaInstr * TINA_MAX_INSTR)]


(SNODE_LAMBERT, SNODE_PHONG, SNODE_COOK, SNODE_EMISSION, SNODE_MIX, SNODE_SCALE, SNODE_ADD) = range(7)
TINA_SAMPLE_MAX_NODES, TINA_SAMPLE_MAX_INSTR = 16, 48


class TinaSampleNode(C.Structure):
    _fields_ = [('kind', C.c_int32), ('a', C.c_int32), ('b', C.c_int32), ('p0', C.c_int32), ('n0', C.c_int32),
                ('p1', C.c_int32), ('n1', C.c_int32), ('pad_', C.c_int32)]


class TinaSampleMaterial(C.Structure):
    _fields_ = [('nnodes', C.c_int32), ('ncode', C.c_int32), ('ntex', C.c_int32), ('pad_', C.c_int32),
                ('tex', C.c_void_p * TINA_MAX_TEX),
                ('tex_w', C.c_int32 * TINA_MAX_TEX), ('tex_h', C.c_int32 * TINA_MAX_TEX), ('tex_c', C.c_int32 * TINA_MAX_TEX),
                ('nodes', TinaSampleNode * TINA_SAMPLE_MAX_NODES), ('code', TinaInstr * TINA_SAMPLE_MAX_INSTR)]


# name -> (restype, argtypes); every symbol include/tina_b200.h declares
_vp, _i, _i64, _u32, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_float
_fp = C.POINTER(C.c_float)
SIGNATURES = {
    'tina_last_error': (C.c_char_p, []),
    'tina_version': (_i, []),
    'tina_launch_count': (C.c_uint64, []),
    'tina_selftest_division': (_i, [_i, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    'tina_engine_create': (_i, [C.POINTER(_vp), _i, _i, _i]),
    'tina_engine_destroy': (_i, [_vp]),
    'tina_engine_set_camera': (_i, [_vp, _fp, _fp]),
    'tina_engine_set_bias': (_i, [_vp, _f, _f]),
    'tina_engine_clear_depth': (_i, [_vp, _vp]),
    'tina_engine_flush': (_i, [_vp, _vp]),
    'tina_engine_set_lazy_clear': (_i, [_vp, _i]),
    'tina_engine_keys': (_i, [_vp, C.POINTER(_vp)]),
    'tina_engine_depth': (_i, [_vp, _vp, _vp]),
    'tina_engine_set_face_base': (_i, [_vp, _u32]),
    'tina_engine_get_face_base': (_i, [_vp, C.POINTER(_u32)]),
    'tina_engine_ipc_export': (_i, [_vp, _vp]),
    'tina_engine_ipc_open_peers': (_i, [_vp, _vp, _i, _i]),
    'tina_engine_ipc_close_peers': (_i, [_vp]),
    'tina_engine_set_peer_keys': (_i, [_vp, C.POINTER(_vp), _i, _i]),
    'tina_raster_create': (_i, [C.POINTER(_vp), _vp, _i64, _u32]),
    'tina_raster_destroy': (_i, [_vp]),
    'tina_raster_set_faces': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp]),
    'tina_raster_set_faces_indexed': (_i, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _fp, _fp, _i, _u32, _vp]),
    'tina_raster_materialize': (_i, [_vp, _vp]),
    'tina_raster_set_faces_grid': (_i, [_vp, _vp, _i, _i, _fp, _fp, _i, _u32, _vp]),
    'tina_raster_render_occup': (_i, [_vp, _vp]),
    'tina_raster_render_color': (_i, [_vp, C.POINTER(TinaMaterial), C.POINTER(TinaLighting), _vp, _u32, _fp, _vp]),
    'tina_raster_render_color_accumulate': (_i, [_vp, C.POINTER(TinaMaterial), C.POINTER(TinaLighting), _vp, _u32, _fp, _vp, _i, _vp]),
    'tina_raster_render_color_range': (_i, [_vp, C.POINTER(TinaMaterial), C.POINTER(TinaLighting), _vp, _u32, _fp, _i64, _i64,
                                             _u32, _vp]),
    'tina_raster_render_color_composite': (_i, [_vp, C.POINTER(TinaMaterial), C.POINTER(TinaLighting), _vp, _u32, _fp, _i64,
                                                 _i64, _u32, _vp]),
    'tina_shared_alloc': (_i, [_i, _i64, C.POINTER(_vp), _vp]),
    'tina_shared_free': (_i, [_i, _vp]),
    'tina_shared_open': (_i, [_i, _vp, C.POINTER(_vp)]),
    'tina_shared_close': (_i, [_i, _vp]),
    'tina_raster_render_gbuffer': (_i, [_vp, _i, _vp, _i, _i, _fp, _vp]),
    'tina_raster_render_gbuffers': (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _fp, _vp]),
    'tina_raster_occup': (_i, [_vp, _vp, _vp]),
    'tina_raster_setup_cache': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tina_raster_buffers': (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    'tina_raster_set_tuning': (_i, [_vp, _i, _i]),
    'tina_raster_stats': (_i, [_vp, C.POINTER(_i64)]),
    'tina_raster_kernel_times': (_i, [_vp, _fp]),
    'tina_pars_create': (_i, [C.POINTER(_vp), _vp, _i64, _u32]),
    'tina_pars_destroy': (_i, [_vp]),
    'tina_pars_set': (_i, [_vp, _vp, _vp, _vp, _i64, _fp, _f, _i, _vp]),
    'tina_pars_render_occup': (_i, [_vp, _vp]),
    'tina_pars_render_color': (_i, [_vp, C.POINTER(TinaMaterial), C.POINTER(TinaLighting), _vp, _u32, _fp, _vp]),
    'tina_pars_occup': (_i, [_vp, _vp, _vp]),
    'tina_pars_render_gbuffers': (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _fp, _vp]),
    'tina_wire_create': (_i, [C.POINTER(_vp), _vp, _i64, _u32, _fp]),
    'tina_wire_destroy': (_i, [_vp]),
    'tina_wire_set_color': (_i, [_vp, _fp]),
    'tina_wire_set': (_i, [_vp, _vp, _i64, _i, _vp]),
    'tina_wire_render_color': (_i, [_vp, C.POINTER(_vp), _i, _vp]),
    'tina_image_fill': (_i, [_vp, _i64, _fp, _vp]),
    'tina_image_tonemap': (_i, [_vp, _i64, _vp]),
    'tina_engine_ssao_render': (_i, [_vp, _vp, _vp, _i, _vp, _i, _f, _f, _f, _vp, _vp]),
    'tina_image_ssao_apply': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'tina_engine_ssao_render_taa': (_i, [_vp, _vp, _i, _f, _f, _f, _u32, _vp, _vp]),
    'tina_image_ssao_apply_taa': (_i, [_vp, _vp, _i, _i, _vp]),
    'tina_engine_ssr_render': (_i, [_vp, _vp, _vp, _vp, C.POINTER(TinaSampleMaterial), _i, _vp, _i, _i, _f, _f, _i, _i, _u32, _vp, _vp]),
    'tina_image_ssr_apply': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    'tina_image_fxaa': (_i, [_vp, _i, _i, _vp, _vp, _f, _f, _f, _vp]),
    'tina_image_bloom': (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _f, _f, _f, _vp]),
    'tina_image_accumulate': (_i, [_vp, _vp, _i64, _i, _vp]),
}

_lib = None


class TinaError(RuntimeError):
    pass


def lib():
    """Load libtina_b200.so (built in-tree by __graft_entry__.build()); raise if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a). taichi_three_b200 has no CPU fallback.')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise TinaError(f'libtina_b200 error {rc}: {lib().tina_last_error().decode()}')


def f32_array(values, n):
    import numpy as np
    a = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(-1))
    assert a.size == n, (a.size, n)
    return a, a.ctypes.data_as(_fp)
