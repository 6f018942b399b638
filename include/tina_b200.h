/*
 * tina_b200.h -- C ABI of the B200-native triangle-raster hot path.
 *
 * The reference (taichi-dev/taichi_three, "Tina") has NO FFI boundary: its
 * rasteriser is a set of Taichi-JIT'd Python classes.  This header is the
 * boundary a maintainer would bind (ctypes / pybind / cgo) to replace
 *
 *     tina/core/engine.py:5-76      Engine          (depth, W2V/V2W, bias)
 *     tina/core/triangle.py:4-153   TriangleRaster  (set_object / render_occup / render_color)
 *     tina/core/shader.py:112-148   Shader / ShaderGroup
 *     tina/core/lighting.py:25-98   Lighting
 *     tina/matr/material.py         Lambert / Phong / CookTorrance / Mix / Scale / Add / Emission
 *     tina/postp/tonemap.py:9-12    ToneMapping.apply
 *
 * Conventions
 *   - extern "C", no exceptions; every call returns 0 on success, <0 on error,
 *     tina_last_error() gives a thread-local message.
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.
 *   - Every launch takes a cudaStream_t (as void*) and is asynchronous.
 *   - Image-like buffers are x-major like the reference's fields
 *     (`ti.field(int, (W, H))`, triangle.py:16): element (x, y) lives at x*H + y.
 *   - Handles are not thread-safe; distinct handles are independent.
 *
 * Visibility key (replaces Engine.depth + TriangleRaster.occup, engine.py:11,
 * triangle.py:16): one signed 64-bit word per pixel,
 *     key = ((int64)depth << 32) | (uint32)(face_base + f + 1)
 * merged with a signed atomicMin.  Smaller depth wins; at equal depth the lower
 * global face id wins, which is exactly the serial outcome of triangle.py:123-125
 * (strict `>`), including "a later object only wins with strictly smaller depth".
 * Low word 0 = "no face".  face_base is the number of faces rasterised since the
 * last tina_clear_depth, so no per-object reset pass is needed.
 */
#ifndef TINA_B200_H
#define TINA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TinaEngine TinaEngine;
typedef struct TinaRaster TinaRaster;

/* raster option bits: TriangleRaster.__init__ kwargs (triangle.py:6-7) */
#define TINA_SMOOTHING 1u
#define TINA_TEXTURING 2u
#define TINA_CULLING   4u
#define TINA_CLIPPING  8u

/* tina_render_color flags */
#define TINA_COLOR_TONEMAP 1u /* fuse aces_tonemap (advans.py:32-35) into the store            */
#define TINA_COLOR_FILL_BG 2u /* also write `bg` to pixels this object does not own (raster.py:176) */
#define TINA_COLOR_FINISH 4u  /* last shading pass of a multi-object frame: pixels this object does not own are read back and
                               * finished too (tonemapped if TINA_COLOR_TONEMAP is set, accumulated if an accumulator is given) */

#define TINA_MAX_LIGHTS 16 /* lighting.py:26 */
#define TINA_MAX_INSTR 96
#define TINA_MAX_TEX 4
#define TINA_MAX_PEERS 16        /* ranks of one node whose key buffers can be composited over peer memory */
#define TINA_IPC_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */

/* Lighting (lighting.py:25-98).  dirs[i] = (x,y,z,w): w=0 directional (already
 * normalised by the host, lighting.py:60-62), w=1 point light. */
typedef struct {
    float dirs[TINA_MAX_LIGHTS][4];
    float colors[TINA_MAX_LIGHTS][4];
    float ambient[4];
    int32_t nlights;
    int32_t pad[3];
} TinaLighting;

/* Material program: the reference's compile-time node graph (matr/nodes.py,
 * matr/material.py) flattened by the host into postfix code over a vec3 stack.
 * Scalars are broadcast to vec3 (identical arithmetic per component). */
enum {
    TINA_OP_CONST = 0,   /* push c                                                     nodes.py:42-49  */
    TINA_OP_INPUT = 1,   /* push input[arg]: 0 pos, 1 color, 2 normal, 3 texcoord      nodes.py:79-96  */
    TINA_OP_TEXTURE = 2, /* pop uv; push bilerp(tex[arg], uv*(shape-1))                nodes.py:99-111 */
    TINA_OP_FRESNEL = 3, /* pop specular, albedo, metallic; push f0                    material.py:69-83 */
    TINA_OP_LAMBERT = 4, /* push 1/pi                                                  material.py:392 */
    TINA_OP_PHONG = 5,   /* pop shineness; push VoR**m*(m+2)/2                         material.py:450-454 */
    TINA_OP_COOK = 6,    /* pop fresnel, roughness; push F*G*D                         material.py:323-362 */
    TINA_OP_MIX = 7,     /* pop b, a, fac; push (1-fac)*a + fac*b                      material.py:96-118 */
    TINA_OP_MUL = 8,     /* pop wei, fac; push fac*wei                                 material.py:157-176 */
    TINA_OP_ADD = 9,     /* pop b, a; push a+b                                         material.py:204-220 */
    TINA_OP_REG = 10,    /* push register[arg] (written by the per-pixel prologue program)               */
    TINA_OP_STORE = 11,  /* pop -> register[arg]                                                         */
    TINA_OP_BCAST = 12,  /* pop v; push (v[arg], v[arg], v[arg])   (uv.x / uv.y of LerpTexture)   nodes.py:129-136 */
    TINA_OP_CHESS = 13,  /* pop size, uv; push ((uv // size).sum() % 2) broadcast (ChessboardTexture) nodes.py:114-126 */
};
#define TINA_MAX_REGS 8
/* Three-address prologue (TinaMaterial.prologue_form = 2), produced by the host compiler from the postfix prologue: one
 * header slot per operation, op = TINA_OP3 | <postfix op>, arg = dst | s0 << 8 | s1 << 16 | s2 << 24 with source codes
 * 0..15 = value register (0..7 are the registers the other programs read with TINA_OP_REG, 8..15 temporaries),
 * 16..19 = TINA_OP_INPUT 0..3, 255 = the constant in the next slot's c[] (slots are consumed in source order).
 * TINA_OP_TEXTURE keeps its texture slot in c[0] of the header; (TINA_OP3 | TINA_OP_REG) copies s0 to dst.
 * Same operations on the same values as the postfix form, a third to a quarter of the interpreter steps. */
#define TINA_OP3 0x100
#define TINA_VM_VALUES 16

typedef struct {
    int32_t op;
    int32_t arg;
    float c[3];
} TinaInstr;

typedef struct {
    /* three programs laid out back to back in `code`:
     * [0, n_brdf) brdf(n, l, v); then n_ambient of ambient(); then n_emission of emission() */
    int32_t n_brdf, n_ambient, n_emission, ntex;
    const float *tex[TINA_MAX_TEX]; /* device, [w][h][c] f32, x-major (advans.py:8-28) */
    int32_t tex_w[TINA_MAX_TEX], tex_h[TINA_MAX_TEX], tex_c[TINA_MAX_TEX];
    /* optional 4th program after emission: run ONCE per pixel before lighting; it evaluates the
     * light-independent, non-constant sub-expressions the host hoisted out of the other three
     * (texture samples, Fresnel factors, ...) into registers they then read with TINA_OP_REG */
    int32_t n_prologue;     /* TinaInstr slots of the prologue program (after brdf, ambient, emission in code[]) */
    int32_t prologue_form;  /* 0: interpret it; 1: it is the 24-slot prologue of tina.PBR with a textured base colour
                             * (TEXTURE -> r0, Fresnel -> r1, diffuse -> r2, ambient -> r3, emission -> r4): run as straight-line code */
    int32_t pad_[2];        /* (prologue_form 2: three-address form, see TINA_OP3; 3 / 4: the 19- / 11-slot prologues of
                             * tina.Classic / tina.Diffuse with a textured colour, straight-line code as well) */
    TinaInstr code[TINA_MAX_INSTR];
} TinaMaterial;

const char *tina_last_error(void);
int tina_version(void);
/* number of kernels this library has launched in this process (all engines, all streams); for benchmarks */
uint64_t tina_launch_count(void);
/* self-test of the shared-divisor IEEE division used by the setup code: computes `nquotients` quotients on random
 * and structured operands both ways and returns the number that differ from __fdiv_rn (must be 0) */
int tina_selftest_division(int device, uint64_t nquotients, uint64_t seed, uint64_t *mismatch_host);

/* ---- Engine (core/engine.py) ------------------------------------------------ */
/* engine.py:6-28: owns the per-pixel key buffer (depth + winner), W2V/V2W (init
 * diag(1,1,-1,1)), bias (.5,.5). */
int tina_engine_create(TinaEngine **out, int device, int W, int H);
int tina_engine_destroy(TinaEngine *e);
/* engine.py:72-76 (host already did proj@view and inv in f64, cast to f32, row-major) */
int tina_engine_set_camera(TinaEngine *e, const float *W2V_host, const float *V2W_host);
/* engine.py:31-39 */
int tina_engine_set_bias(TinaEngine *e, float bx, float by);
/* engine.py:68-70: depth := 2**30, winner := none, face_base := 0.  The device-side clear is deferred to the next
 * library call that touches the keys (render_occup of a MeshGrid / MeshModel folds it into its vertex-stage launch);
 * host-side state changes at once.  Nothing is deferred under stream capture. */
int tina_engine_clear_depth(TinaEngine *e, void *stream);
/* run a deferred clear now (call before reading / writing the memory behind tina_engine_keys directly) */
int tina_engine_flush(TinaEngine *e, void *stream);
/* on = 0: tina_engine_clear_depth clears immediately (default 1 = deferred) */
int tina_engine_set_lazy_clear(TinaEngine *e, int on);
/* int64 keys[W*H]; the high words are Engine.depth viewed with stride 2 (see tina_engine_flush) */
int tina_engine_keys(TinaEngine *e, int64_t **keys);
/* materialise Engine.depth as a dense int32[W*H] */
int tina_engine_depth(TinaEngine *e, int32_t *depth, void *stream);
/* sort-last support: every later face id is offset by `base` (global ids across ranks) */
int tina_engine_set_face_base(TinaEngine *e, uint32_t base);
int tina_engine_get_face_base(TinaEngine *e, uint32_t *base_host);

/* ---- TriangleRaster (core/triangle.py) --------------------------------------- */
int tina_raster_create(TinaRaster **out, TinaEngine *e, int64_t maxfaces, uint32_t flags);
int tina_raster_destroy(TinaRaster *r);
/* set_object for SimpleMesh-like sources (triangle.py:72-86, mesh/simple.py:33-47):
 * verts [N,3,3], norms [N,3,3] (iff SMOOTHING), coors [N,3,2] (iff TEXTURING).
 * borrow=1 aliases the caller's buffers (zero-copy) until the next set_faces*. */
int tina_raster_set_faces(TinaRaster *r, const float *verts, const float *norms, const float *coors,
                          int64_t nfaces, int borrow, void *stream);
/* set_object for MeshModel (+MeshTransform, +MeshNoCulling/FlipCulling/FlipNormal):
 * mesh/model.py:56-73, mesh/trans.py:28-40, mesh/cull.py:6-57.
 * v [nverts,3], vt [*,2], vn [nnorms,3]; faces [N,3,3] int32 = [corner][v, vt, vn].
 * trans_host / trans_normal_host: ntrans 4x4 / 3x3 matrices of nested MeshTransform wrappers, innermost first, each
 * applied with its own rounding like the reference's call chain (ntrans = 0 or NULL: none; at most 4).
 * mode bits: 1 = double sided (MeshNoCulling), 2 = flip winding (MeshFlipCulling),
 *            4 = negate normals (MeshFlipNormal).
 * Indexed sources are NOT expanded: a per-unique-vertex stage (world position / normal here, clip
 * coordinates in render_occup) feeds the kernels through the mesh's own index buffer, which must
 * stay valid until the next set_faces*.  tina_raster_materialize writes the expanded copies. */
int tina_raster_set_faces_indexed(TinaRaster *r, const float *v, int64_t nverts, const float *vt, const float *vn,
                                  int64_t nnorms, const int32_t *faces, int64_t nfaces, const float *trans_host,
                                  const float *trans_normal_host, int ntrans, uint32_t mode, void *stream);
/* set_object for MeshGrid (mesh/grid.py:26-58): pos [nx,ny,3]; recomputes the
 * per-vertex normals like MeshGrid.pre_compute, texcoords (i/(nx-1), j/(ny-1)). */
int tina_raster_set_faces_grid(TinaRaster *r, const float *pos, int nx, int ny, const float *trans_host,
                               const float *trans_normal_host, int ntrans, uint32_t mode, void *stream);
/* triangle.py:89-131 */
int tina_raster_render_occup(TinaRaster *r, void *stream);
/* triangle.py:134-153 + shader.py:119-131 + lighting.py:84-98; image [W,H,3] f32 */
int tina_raster_render_color(TinaRaster *r, const TinaMaterial *mat_host, const TinaLighting *light_host,
                             float *image, uint32_t flags, const float *bg_host, void *stream);
/* the same with the frame's TAA accumulation (util/accumator.py:16-23: acc = acc * (1 - 1/count) + image * (1/count)) fused
 * into the pass: needs TINA_COLOR_FILL_BG (single-object frame) or TINA_COLOR_FINISH (last object), i.e. a pass that
 * visits every pixel; acc [W,H,3] f32, count >= 1 */
int tina_raster_render_color_accumulate(TinaRaster *r, const TinaMaterial *mat_host, const TinaLighting *light_host, float *image,
                                        uint32_t flags, const float *bg_host, float *acc, int count, void *stream);
/* render_color restricted to pixels [first_pixel, first_pixel + npixels) (first_pixel a multiple of 256) of
 * the CURRENT face arrays with an explicit id offset: after a sort-last key composite every rank shades
 * one screen strip from the replicated attributes (face_base = 0). */
int tina_raster_render_color_range(TinaRaster *r, const TinaMaterial *mat_host, const TinaLighting *light_host,
                                   float *image, uint32_t flags, const float *bg_host, int64_t first_pixel, int64_t npixels,
                                   uint32_t face_base, void *stream);
/* Sort-last over peer memory (no counterpart in the reference): every rank of one node rasterises its
 * face range into its own engine; tina_engine_ipc_export gives a CUDA IPC handle of the engine's key buffer,
 * tina_engine_ipc_open_peers maps the buffers of all `world` ranks (handles_host = world x 64 bytes in rank order;
 * entry `rank` is ignored).  tina_raster_render_color_composite then shades pixels [first_pixel, first_pixel +
 * npixels) like render_color_range, but takes each pixel's key as the MIN over all ranks' buffers, read over
 * NVLink inside the shading kernel, and stores the composited key in the local buffer.  The caller orders it
 * after every rank's render_occup (any collective on the same stream, e.g. a barrier) and keeps the next
 * clear_depth behind every rank's composite (the all-gather of the image strips does that). */
int tina_engine_ipc_export(TinaEngine *e, uint8_t *handle64_host);
int tina_engine_ipc_open_peers(TinaEngine *e, const uint8_t *handles_host, int world, int rank);
int tina_engine_ipc_close_peers(TinaEngine *e);
/* the same table from plain device pointers (engines of one process: several engines on one GPU, or several GPUs
 * with peer access enabled); keys_host[rank] is ignored, the pointers stay owned by the caller */
int tina_engine_set_peer_keys(TinaEngine *e, int64_t *const *keys_host, int world, int rank);
int tina_raster_render_color_composite(TinaRaster *r, const TinaMaterial *mat_host, const TinaLighting *light_host,
                                       float *image, uint32_t flags, const float *bg_host, int64_t first_pixel,
                                       int64_t npixels, uint32_t face_base, void *stream);
/* A device buffer one rank allocates and the other ranks of the node map (CUDA IPC): the frame image the ranks'
 * shading kernels store their strips into directly (sort-last with the image assembled on one rank, no gather). */
int tina_shared_alloc(int device, int64_t bytes, void **ptr, uint8_t *handle64_host);
int tina_shared_free(int device, void *ptr);
int tina_shared_open(int device, const uint8_t *handle64_host, void **ptr);
int tina_shared_close(int device, void *ptr);
/* G-buffer sinks of core/shader.py:21-109 for the current object (ShaderGroup fan-out, shader.py:138-148):
 * writes `ncomp` (1..3) float32 (or int32 if out_is_int) values per pixel where the object is visible,
 * out[(x*H + y)*ncomp + k].  param_host: ConstShader value (3 floats) / ChessboardShader size (1 float). */
enum {
    TINA_SINK_CONST = 0,      /* shader.py:21-28  */
    TINA_SINK_POSITION = 1,   /* shader.py:31-34  */
    TINA_SINK_DEPTH = 2,      /* shader.py:37-40  */
    TINA_SINK_NORMAL = 3,     /* shader.py:43-46  */
    TINA_SINK_VIEWNORMAL = 4, /* shader.py:49-56  */
    TINA_SINK_TEXCOORD = 5,   /* shader.py:59-62  */
    TINA_SINK_COLOR = 6,      /* shader.py:65-68  */
    TINA_SINK_CHESSBOARD = 7, /* shader.py:71-79  */
    TINA_SINK_VIEWDIR = 8,    /* shader.py:96-101 */
    TINA_SINK_SIMPLE = 9,     /* shader.py:104-109 */
    TINA_SINK_ELMID = 10,     /* probe.py:21-22: the visible face's id (ProbeShader.elmid) */
};
int tina_raster_render_gbuffer(TinaRaster *r, int kind, void *out, int ncomp, int out_is_int, const float *param_host,
                               void *stream);
/* the same for up to 8 sinks in ONE launch (a ShaderGroup's pre / post shaders, scene/raster.py:101-107): the face is
 * gathered and the weights recomputed once per pixel.  Host arrays of nsinks entries; params_host: [nsinks][3] or NULL */
int tina_raster_render_gbuffers(TinaRaster *r, int nsinks, const int *kinds_host, void *const *outs_host,
                                const int *ncomps_host, const int *is_int_host, const float *params_host, void *stream);
/* materialise TriangleRaster.occup as int32[W*H] (-1 = none) for the last render_occup */
int tina_raster_occup(TinaRaster *r, int32_t *occup, void *stream);
/* write the expanded [N,3,3] / [N,3,2] copies of the current object (the reference's raster.verts /
 * norms / coors fields, triangle.py:18-22) if the indexed path skipped them */
int tina_raster_materialize(TinaRaster *r, void *stream);
/* the reference's public per-face setup cache (triangle.py:25-29,127-131) of the current object under the current camera,
 * materialised on demand: bcn, can, boo, coo [nfaces][2], wsc [nfaces][3] (device); only faces that pass cull + clip are
 * written, like the reference */
int tina_raster_setup_cache(TinaRaster *r, float *bcn, float *can, float *boo, float *coo, float *wsc, void *stream);
/* device views of the current object's expanded attribute buffers (NULL until materialised) */
int tina_raster_buffers(TinaRaster *r, const float **verts, const float **norms, const float **coors,
                        int64_t *nfaces);
/* strategy knobs (every setting yields identical bits); which:
 * 0 = most candidate pixels a face may have to be rasterised per thread in the setup kernel,
 * 2 = force every face through the tile path, 3 = collect stats, 4 = record CUDA events around
 * every kernel (tina_raster_kernel_times), 5 = candidate tightening on/off, 6 = read the key
 * before the atomicMin on/off, 7 = largest queue the tile path handles without binning,
 * 8 = always interpret the material program (no specialised shading kernels),
 * 9 = warp-shared candidate walk in the setup kernel: 0 never, 1 decide per warp, 2 always,
 * 10 = programmatic dependent launch of k_raster_faces / k_render_color on/off,
 * 11 = indexed (per-unique-vertex) path for MeshGrid / MeshModel sources on/off,
 * 12 = adaptive tile path: after 8 consecutive render_occup/render_color pairs that queued no large
 *      face the tile-path kernel is not launched and the setup kernel walks large faces itself,
 * 13 = fast shading (default 1): render_color keeps the barycentric weights in the reference's exact
 *      arithmetic but evaluates interpolation, normalisation, the view ray, lighting and the tone curve
 *      with FMA contraction and SFU rcp/rsqrt (colour within 1e-4 of the reference, ids/depth unaffected).
 *      Applies to constant-brdf and Lambert+Phong (constant shineness <= 64) materials; Cook-Torrance and
 *      interpreted programs always shade exactly;
 *      0 = every shading op in the reference's order with IEEE division/sqrt,
 * 14 = lean kernels (default 1): the default raster options / plain sources as compile-time constants in the
 *      rasteriser; with fast shading, Diffuse / Classic materials whose parameters are all constants on untextured
 *      rasters run shading kernels with compile-time raster flags and no operand tests,
 * 15 = indexed sources: mark every vertex record "not tame", so that every face takes the rasteriser's general
 *      path (float bounding box, x86 conversions, no tightening) instead of the per-vertex integer bounds,
 * 16 = plain square MeshGrid sources: independent persistent warps over row chunks staged through shared memory by
 *      cp.async (k_raster_grid; default 0: measured slower than the gather kernel every indexed source uses),
 * 17 = plain square MeshGrid sources: one quad (two faces, four vertex records) per thread (k_raster_quads, default 1);
 *      0 = one face per thread like every other indexed source,
 * 18 = indexed sources: when no set_faces* call happened since the previous render_occup, the vertex-stage blocks of this
 *      one do not wait for the previous render_color (they write the other of two per-vertex record sets; the blocks that
 *      clear the keys still wait), so consecutive frames overlap by that kernel's last wave (default 1),
 * 19 = render_color passes without frame glue / composite run as a smaller grid that walks the 256-pixel chunks with a
 *      grid stride: value / 4 chunks per CTA, at least five CTAs per SM (default 15 = 3.75 chunks; 0 = one CTA per chunk) */
int tina_raster_set_tuning(TinaRaster *r, int which, int value);
/* counters of the last render_occup (synchronises): faces culled, clipped, per-thread,
 * per-warp, queued for the tile path, tile-list entries */
int tina_raster_stats(TinaRaster *r, int64_t *out6_host);
/* ms of the last launch of: [0] k_raster_faces / k_raster_indexed, [1] k_frame_prologue (indexed sources only:
 * vertex records + the deferred key clear), [3] k_large_path,
 * [4] k_render_color ([2] unused; -1 = never recorded); needs tuning knob 4; synchronises on the events */
int tina_raster_kernel_times(TinaRaster *r, float *ms5_host);

/* ---- ParticleRaster (core/particle.py:4-161): sphere splats on the same Engine (shared depth / ids) ---- */
typedef struct TinaPars TinaPars;
/* flags: 1 = coloring, 2 = clipping (particle.py:6-12) */
int tina_pars_create(TinaPars **out, TinaEngine *e, int64_t maxpars, uint32_t flags);
int tina_pars_destroy(TinaPars *r);
/* set_object (particle.py:64-76) for SimpleParticles (+ParsTransform, pars/trans.py:22-31): verts [N,3], sizes [N],
 * colors [N,3] or NULL; trans_host (4x4) / scale optional; borrow=1 aliases the caller's buffers */
int tina_pars_set(TinaPars *r, const float *verts, const float *sizes, const float *colors, int64_t npars,
                  const float *trans_host, float scale, int borrow, void *stream);
int tina_pars_render_occup(TinaPars *r, void *stream);                    /* particle.py:78-127 */
int tina_pars_render_color(TinaPars *r, const TinaMaterial *mat_host, const TinaLighting *light_host, float *image,
                           uint32_t flags, const float *bg_host, void *stream); /* particle.py:129-161 */
int tina_pars_occup(TinaPars *r, int32_t *occup, void *stream);
/* G-buffer sinks for the current particles (particle.py:129-161 hands pos / normal / texcoord (0, 0) / colour of the visible
 * sphere point to any shader of a ShaderGroup): same sink kinds and layout as tina_raster_render_gbuffers */
int tina_pars_render_gbuffers(TinaPars *r, int nsinks, const int *kinds_host, void *const *outs_host, const int *ncomps_host,
                              const int *out_is_int_host, const float *params_host, void *stream);

/* ---- WireframeRaster (core/wireframe.py:4-95): depth-tested DDA lines on the same Engine ---- */
typedef struct TinaWire TinaWire;
/* flags: 2 = clipping (wireframe.py:7); linecolor_host: 3 floats or NULL for the default (.9, .6, 0) */
int tina_wire_create(TinaWire **out, TinaEngine *e, int64_t maxwires, uint32_t flags, const float *linecolor_host);
int tina_wire_destroy(TinaWire *w);
int tina_wire_set_color(TinaWire *w, const float *linecolor_host);
/* set_object (wireframe.py:32-38): npoly = 0: verts [N,2,3] borrowed; npoly > 0: polygon faces [N/npoly, npoly, 3]
 * turned into their N edges like MeshToWire (mesh/wire.py:19-27) */
int tina_wire_set(TinaWire *w, const float *verts, int64_t nwires, int npoly, void *stream);
/* wireframe.py:70-95 (render_occup is a no-op there): depth test + Shader.blend_color(1, linecolor) into each image */
int tina_wire_render_color(TinaWire *w, float *const *images_host, int nimages, void *stream);

/* ---- frame glue (scene/raster.py:176,202-203) -------------------------------- */
int tina_image_fill(float *image, int64_t npixels, const float *rgb_host, void *stream);
int tina_image_tonemap(float *image, int64_t nfloats, void *stream);
/* SSAO (postp/ssao.py, non-TAA mode).  render (:65-96): ambient-occlusion field ao [W,H] from the engine's depth, the
 * scene's world-normal G-buffer normals [W,H,3] (NormalShader, scene/raster.py:51-54), the sample table [nsamples,3] and
 * the rotation table [noise_size,noise_size,2] that the reference draws once at construction (:24-36).
 * apply (:38-49): image [W,H,3] *= 1 - box_{noise_size}(ao) / noise_size^2. */
int tina_engine_ssao_render(TinaEngine *e, const float *normals, const float *samples, int nsamples, const float *rotations,
                            int noise_size, float radius, float thresh, float factor, float *ao, void *stream);
int tina_image_ssao_apply(float *image, const float *ao, int W, int H, int noise_size, void *stream);
/* SSAO with fresh samples per pixel and frame (postp/ssao.py:52-56, 80-81 `taa=True`): make_sample() draws three uniform
 * numbers per sample.  The reference takes them from Taichi's global ti.random(), whose stream is not specified; here
 * they come from the Wang hash of tina/random.py:26-35 seeded with (pixel, frame, draw index), so a frame is
 * reproducible and the oracle restates it bit for bit.  apply_taa (:40-41): image *= 1 - ao. */
int tina_engine_ssao_render_taa(TinaEngine *e, const float *normals, int nsamples, float radius, float thresh, float factor,
                                uint32_t frame, float *ao, void *stream);
int tina_image_ssao_apply_taa(float *image, const float *ao, int W, int H, void *stream);

/* ---- SSR (postp/ssr.py) ------------------------------------------------------------------------------------------
 * material.sample(idir, nrm, sign, rng) of matr/material.py as a tree the device walks: node 0 is the root; a node's
 * parameters (MixMaterial / ScaleMaterial factor, Phong shineness, CookTorrance roughness + fresnel) are postfix value
 * programs in code[] made of TINA_OP_CONST / INPUT / TEXTURE / FRESNEL (p0/n0 first parameter, p1/n1 second). */
enum {
    TINA_SNODE_LAMBERT = 0,  /* material.py:398-405 */
    TINA_SNODE_PHONG = 1,    /* :459-472, p0 = shineness */
    TINA_SNODE_COOK = 2,     /* :364-384, p0 = roughness, p1 = fresnel */
    TINA_SNODE_EMISSION = 3, /* :679-681 */
    TINA_SNODE_MIX = 4,      /* :123-138, a = mat1, b = mat2, p0 = factor */
    TINA_SNODE_SCALE = 5,    /* :180-184, a = mat, p0 = factor */
    TINA_SNODE_ADD = 6,      /* :227-238, a = mat1, b = mat2 */
};
#define TINA_SAMPLE_MAX_NODES 16
#define TINA_SAMPLE_MAX_INSTR 48
typedef struct {
    int32_t kind, a, b, p0, n0, p1, n1, pad_;
} TinaSampleNode;
typedef struct {
    int32_t nnodes, ncode, ntex, pad_;
    const float *tex[TINA_MAX_TEX]; /* device, [w][h][c] f32 */
    int32_t tex_w[TINA_MAX_TEX], tex_h[TINA_MAX_TEX], tex_c[TINA_MAX_TEX];
    TinaSampleNode nodes[TINA_SAMPLE_MAX_NODES];
    TinaInstr code[TINA_SAMPLE_MAX_INSTR];
} TinaSampleMaterial;
/* SSR.render (ssr.py:44-103): per pixel with a normal, nsamples reflection rays drawn with material.sample() of the
 * pixel's material (table_host[mtlid[P]]) and marched through the engine's depth buffer for at most nsteps steps; a hit
 * adds bilerp(image) * weight.  normals [W,H,3], coors [W,H,2] or NULL (the reference's dummy (1,1) field: texcoord 0),
 * mtlid int32 [W,H], image [W,H,3] (read only), out4 [W,H,4].  Random numbers: taa = 0: WangHashRNG(P % blurring)
 * exactly as ssr.py:73-76 (tina/random.py:17-62); taa = 1 (ssr.py:72, Taichi's unspecified ti.random()): the same
 * hash seeded with (pixel, frame).  Reference defaults (:20-28): nsamples 32 / 12 (taa), nsteps 32 / 64, stepsize 2,
 * tolerance 15, blurring 4.
 * SSR.apply (:30-42): image = image * (1 - res.w) + res.xyz with res = out4 (taa) or its blurring x blurring box mean. */
int tina_engine_ssr_render(TinaEngine *e, const float *normals, const float *coors, const int32_t *mtlid,
                           const TinaSampleMaterial *table_host, int nmaterials, const float *image, int nsamples, int nsteps,
                           float stepsize, float tolerance, int blurring, int taa, uint32_t frame, float *out4, void *stream);
int tina_image_ssr_apply(float *image, const float *img4, int W, int H, int blurring, int taa, void *stream);
/* FXAA (postp/fxaa.py:28-68) in place on image [W,H,3]; scratch_lumi [W*H], scratch_copy [W*H*3] floats.
 * Reference defaults: abs_thresh 0.0625, rel_thresh 0.063, factor 1.  Out-of-image taps read 0. */
int tina_image_fxaa(float *image, int W, int H, float *scratch_lumi, float *scratch_copy, float abs_thresh,
                    float rel_thresh, float factor, void *stream);
/* Blooming (postp/blooming.py:45-68) in place; scratch_a/b: [(W/2)*(H/2)*3] floats; gwei: device [radius+1]
 * normalised Gaussian weights (blooming.py:26-37).  Reference defaults: thresh 1, scale 0.25, factor 1,
 * radius min(W,H)/16, sigma 1. */
int tina_image_bloom(float *image, int W, int H, float *scratch_a, float *scratch_b, const float *gwei, int radius,
                     float thresh, float scale, float factor, void *stream);
/* TAA accumulation, util/accumator.py:16-23: acc = acc * (1 - 1/count) + src * (1/count), count >= 1 */
int tina_image_accumulate(float *acc, const float *src, int64_t nfloats, int count, void *stream);

#ifdef __cplusplus
}
#endif
#endif
