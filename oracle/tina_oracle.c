/*
 * tina_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE f32, no contraction) of the reference's
 * triangle-raster path, used ONLY by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs as the checker / CPU baseline.
 * The product path (taichi_three_b200/) never imports or links this file.
 *
 * Parity pin: the reference (taichi-dev/taichi_three) ships no golden outputs and
 * its runtime dependency `taichi` (unpinned, setup.py:23; needs the 0.7.x API) is
 * not installable here.  This restatement is pinned against the reference's OWN
 * Python source executed under an f32 NumPy emulation of the Taichi runtime
 * (oracle/ref_shim + tests/golden/make_golden.py -> tests/golden/*.npz).  Against a
 * real Taichi JIT build parity remains unpinned.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile)
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tina_b200.h"

#define MAXDEPTH_F 1073741824.0f /* engine.py:12  maxdepth = 2**30 */

/* x86 cvttss2si semantics of Taichi's CPU backend for int(float): NaN / out of
 * range -> INT_MIN (SURVEY hard part 6). */
static int32_t f2i(float x) {
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)x;
}
/* common.py:130-137 */
static int32_t ifloor_(float x) { return f2i(floorf(x)); }
static int32_t iceil_(float x) { return f2i(ceilf(x)); }

/* common.py:169-177  mapply(mat, pos, wei) -> (res, rew) */
static void mapply(const float *M, const float *p, float w, float *r, float *rw) {
    for (int i = 0; i < 3; i++) r[i] = M[i * 4 + 3] * w;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i] += M[i * 4 + j] * p[j];
    float rew = M[15] * w;
    for (int i = 0; i < 3; i++) rew += M[12 + i] * p[i];
    *rw = rew;
}
/* common.py:181-183 */
static void mapply_pos(const float *M, const float *p, float *r) {
    float rw;
    mapply(M, p, 1.0f, r, &rw);
    for (int i = 0; i < 3; i++) r[i] = r[i] / rw;
}

static float dot3(const float *a, const float *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static void cross3(const float *a, const float *b, float *r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
/* taichi Matrix.normalized(): invlen = 1 / sqrt(norm_sqr); invlen * v */
static void normalize3(float *v) {
    float inv = 1.0f / sqrtf(dot3(v, v));
    for (int i = 0; i < 3; i++) v[i] = inv * v[i];
}

typedef struct {
    int ok;              /* passed cull + clip */
    int culled, clipped;
    float Avz, Bvz, Cvz; /* NDC z */
    float b[2], c[2], bcn[2], can[2], wsc[3];
    int32_t bot[2], top[2];
} FaceSetup;

/* triangle.py:93-113 */
static void face_setup(const float *V9, const float *W2V, int W, int H, uint32_t flags, FaceSetup *s) {
    float Av[3], Bv[3], Cv[3], rw[3], tmp[3];
    const float *Al = V9, *Bl = V9 + 3, *Cl = V9 + 6;
    mapply(W2V, Al, 1.0f, tmp, &rw[0]);
    for (int i = 0; i < 3; i++) Av[i] = tmp[i] / rw[0];
    mapply(W2V, Bl, 1.0f, tmp, &rw[1]);
    for (int i = 0; i < 3; i++) Bv[i] = tmp[i] / rw[1];
    mapply(W2V, Cl, 1.0f, tmp, &rw[2]);
    for (int i = 0; i < 3; i++) Cv[i] = tmp[i] / rw[2];
    s->ok = 0;
    s->culled = s->clipped = 0;
    /* triangle.py:95-98 */
    float facing = (Bv[0] - Av[0]) * (Cv[1] - Av[1]) - (Bv[1] - Av[1]) * (Cv[0] - Av[0]);
    if (facing <= 0 && (flags & TINA_CULLING)) {
        s->culled = 1;
        return;
    }
    /* triangle.py:100-104 */
    if (flags & TINA_CLIPPING) {
        int ina = 1, inb = 1, inc = 1;
        for (int i = 0; i < 3; i++) {
            ina &= (-1.0f <= Av[i]) && (Av[i] <= 1.0f);
            inb &= (-1.0f <= Bv[i]) && (Bv[i] <= 1.0f);
            inc &= (-1.0f <= Cv[i]) && (Cv[i] <= 1.0f);
        }
        if (!ina && !inb && !inc) {
            s->clipped = 1;
            return;
        }
    }
    /* engine.py:60-61 */
    float res[2] = {(float)W, (float)H};
    float a[2], b[2], c[2];
    for (int i = 0; i < 2; i++) {
        a[i] = (Av[i] * 0.5f + 0.5f) * res[i];
        b[i] = (Bv[i] * 0.5f + 0.5f) * res[i];
        c[i] = (Cv[i] * 0.5f + 0.5f) * res[i];
    }
    /* triangle.py:108-109 */
    int resi[2] = {W, H};
    for (int i = 0; i < 2; i++) {
        int32_t bot = ifloor_(fminf(fminf(a[i], b[i]), c[i]));
        int32_t top = iceil_(fmaxf(fmaxf(a[i], b[i]), c[i]));
        s->bot[i] = bot > 0 ? bot : 0;
        s->top[i] = top < resi[i] - 1 ? top : resi[i] - 1;
    }
    /* triangle.py:110-113 */
    float n = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
    for (int i = 0; i < 2; i++) {
        s->bcn[i] = (b[i] - c[i]) / n;
        s->can[i] = (c[i] - a[i]) / n;
        s->b[i] = b[i];
        s->c[i] = c[i];
    }
    for (int i = 0; i < 3; i++) s->wsc[i] = 1.0f / rw[i];
    s->Avz = Av[2];
    s->Bvz = Bv[2];
    s->Cvz = Cv[2];
    s->ok = 1;
}

/* triangle.py:115-119 (and :147-151) */
static void pixel_weights(const FaceSetup *s, int x, int y, const float *bias, float *wei) {
    float px = (float)x + bias[0], py = (float)y + bias[1];
    float w_bc = (px - s->b[0]) * s->bcn[1] - (py - s->b[1]) * s->bcn[0];
    float w_ca = (px - s->c[0]) * s->can[1] - (py - s->c[1]) * s->can[0];
    wei[0] = w_bc * s->wsc[0];
    wei[1] = w_ca * s->wsc[1];
    wei[2] = ((1.0f - w_bc) - w_ca) * s->wsc[2];
    float sum = (wei[0] + wei[1]) + wei[2];
    wei[0] = wei[0] / sum;
    wei[1] = wei[1] / sum;
    wei[2] = wei[2] / sum;
}

/*
 * triangle.py:89-131, serial execution order (face 0..N-1): the deterministic
 * restatement of the racy atomic_min/store pair (:123-125): lowest face id wins
 * exact depth ties, a face only wins with depth strictly below what is stored.
 * depth[] persists across calls (engine.py:68-70 clears it), occup[] is reset.
 * tie[] (optional): 1 where >= 2 covering faces of this call share the winning depth.
 * stats[] (optional): {culled, clipped, rasterised, candidate pixels, covered samples}
 */
void orc_render_occup(const float *verts, int64_t nfaces, const float *W2V, const float *bias, int W, int H,
                      uint32_t flags, int32_t *depth, int32_t *occup, uint8_t *tie, int64_t *stats) {
    int64_t npix = (int64_t)W * H;
    for (int64_t i = 0; i < npix; i++) occup[i] = -1; /* triangle.py:90-91 */
    if (tie) memset(tie, 0, npix);
    int64_t st[5] = {0, 0, 0, 0, 0};
    for (int64_t f = 0; f < nfaces; f++) {
        FaceSetup s;
        face_setup(verts + f * 9, W2V, W, H, flags, &s);
        if (!s.ok) {
            st[0] += s.culled;
            st[1] += s.clipped;
            continue;
        }
        st[2]++;
        for (int32_t x = s.bot[0]; x <= s.top[0]; x++)
            for (int32_t y = s.bot[1]; y <= s.top[1]; y++) {
                float wei[3];
                st[3]++;
                pixel_weights(&s, x, y, bias, wei);
                if (!(wei[0] >= 0 && wei[1] >= 0 && wei[2] >= 0)) continue; /* :120 */
                st[4]++;
                float depth_f = (wei[0] * s.Avz + wei[1] * s.Bvz) + wei[2] * s.Cvz; /* :121 */
                int32_t d = f2i(depth_f * MAXDEPTH_F);                             /* :122 */
                int64_t P = (int64_t)x * H + y;
                int32_t old = depth[P];
                if (old > d) { /* :123-125 */
                    depth[P] = d;
                    occup[P] = (int32_t)f;
                    if (tie) tie[P] = 0;
                } else if (tie && old == d && occup[P] != -1) {
                    tie[P] = 1;
                }
            }
    }
    if (stats) memcpy(stats, st, sizeof st);
}

/* Per-face setup record export (tests compare the CUDA setup kernel against it):
 * out[f*16 ..] = ok, bcn.xy, can.xy, b.xy, c.xy, wsc.xyz, z.xyz ; bbox[f*4..] = bot.xy, top.xy */
void orc_face_setup(const float *verts, int64_t nfaces, const float *W2V, int W, int H, uint32_t flags, float *out,
                    int32_t *bbox) {
    for (int64_t f = 0; f < nfaces; f++) {
        FaceSetup s;
        memset(&s, 0, sizeof s);
        face_setup(verts + f * 9, W2V, W, H, flags, &s);
        float *o = out + f * 16;
        o[0] = (float)s.ok;
        o[1] = s.bcn[0], o[2] = s.bcn[1], o[3] = s.can[0], o[4] = s.can[1];
        o[5] = s.b[0], o[6] = s.b[1], o[7] = s.c[0], o[8] = s.c[1];
        o[9] = s.wsc[0], o[10] = s.wsc[1], o[11] = s.wsc[2];
        o[12] = s.Avz, o[13] = s.Bvz, o[14] = s.Cvz, o[15] = 0;
        bbox[f * 4 + 0] = s.bot[0], bbox[f * 4 + 1] = s.bot[1];
        bbox[f * 4 + 2] = s.top[0], bbox[f * 4 + 3] = s.top[1];
    }
}

/*
 * The reference's parallel structure for TIMING ONLY (cpu_baseline): faces across
 * threads, atomic min on depth, racy occup store (triangle.py:92,123-125 on the
 * Taichi CPU backend = one task per face range on a thread pool).
 */
void orc_render_occup_parallel(const float *verts, int64_t nfaces, const float *W2V, const float *bias, int W, int H,
                               uint32_t flags, int32_t *depth, int32_t *occup) {
    int64_t npix = (int64_t)W * H;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < npix; i++) occup[i] = -1;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t f = 0; f < nfaces; f++) {
        FaceSetup s;
        face_setup(verts + f * 9, W2V, W, H, flags, &s);
        if (!s.ok) continue;
        for (int32_t x = s.bot[0]; x <= s.top[0]; x++)
            for (int32_t y = s.bot[1]; y <= s.top[1]; y++) {
                float wei[3];
                pixel_weights(&s, x, y, bias, wei);
                if (!(wei[0] >= 0 && wei[1] >= 0 && wei[2] >= 0)) continue;
                float depth_f = (wei[0] * s.Avz + wei[1] * s.Bvz) + wei[2] * s.Cvz;
                int32_t d = f2i(depth_f * MAXDEPTH_F);
                int64_t P = (int64_t)x * H + y;
                int32_t old = __atomic_load_n(&depth[P], __ATOMIC_RELAXED);
                while (old > d && !__atomic_compare_exchange_n(&depth[P], &old, d, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                }
                if (old > d)
                    if (__atomic_load_n(&depth[P], __ATOMIC_RELAXED) >= d) occup[P] = (int32_t)f;
            }
    }
}

/* ---- materials ---------------------------------------------------------------- */

typedef struct {
    float pos[3], color[3], normal[3], texcoord[3];
} ShadeInputs;

/* common.py:140-149 bilerp + nodes.py:107-111 Texture.__call__; the `+1` texel index
 * is clamped to the last texel (reference reads one past the end with weight 0, Appendix B) */
static void tex_sample(const float *tex, int w, int h, int c, const float *uv, float *out) {
    float p[2] = {uv[0] * (float)(w - 1), uv[1] * (float)(h - 1)};
    int32_t I[2] = {ifloor_(p[0]), ifloor_(p[1])};
    float x[2] = {p[0] - (float)I[0], p[1] - (float)I[1]};
    float y[2] = {1.0f - x[0], 1.0f - x[1]};
    int i0 = I[0] < 0 ? 0 : (I[0] > w - 1 ? w - 1 : I[0]);
    int j0 = I[1] < 0 ? 0 : (I[1] > h - 1 ? h - 1 : I[1]);
    int i1 = I[0] + 1 < 0 ? 0 : (I[0] + 1 > w - 1 ? w - 1 : I[0] + 1);
    int j1 = I[1] + 1 < 0 ? 0 : (I[1] + 1 > h - 1 ? h - 1 : I[1] + 1);
    for (int k = 0; k < 3; k++) {
        int kk = c == 1 ? 0 : k;
        float f11 = tex[((int64_t)i1 * h + j1) * c + kk];
        float f10 = tex[((int64_t)i1 * h + j0) * c + kk];
        float f00 = tex[((int64_t)i0 * h + j0) * c + kk];
        float f01 = tex[((int64_t)i0 * h + j1) * c + kk];
        out[k] = ((f11 * x[0] * x[1] + f10 * x[0] * y[1]) + f00 * y[0] * y[1]) + f01 * y[0] * x[1];
    }
}

static void op_bcast(float *v, int k) {
    const float c = v[k];
    v[0] = v[1] = v[2] = c;
}
/* nodes.py:114-126: (texcoord // size).sum() % 2 with Taichi's float semantics: a // b = floor(a / b) on the rounded f32
 * quotient, a % b = a - b * (a // b); the result replaces uv (broadcast) */
static void op_chess(float *uv, const float *size) {
    const float t = floorf(uv[0] / size[0]) + floorf(uv[1] / size[1]);
    const float m = t - 2.0f * floorf(t / 2.0f);
    uv[0] = uv[1] = uv[2] = m;
}

#define STK 16
/* evaluate one postfix program; see include/tina_b200.h for the op table */
static void run_program(const TinaMaterial *m, const float *const *texhost, int begin, int n, const ShadeInputs *in,
                        const float *nrm, const float *idir, const float *odir, float *out) {
    float st[STK][3];
    int sp = 0;
    for (int pc = begin; pc < begin + n; pc++) {
        const TinaInstr *I = &m->code[pc];
        switch (I->op) {
        case TINA_OP_CONST:
            for (int k = 0; k < 3; k++) st[sp][k] = I->c[k];
            sp++;
            break;
        case TINA_OP_INPUT: {
            const float *src = I->arg == 0 ? in->pos : I->arg == 1 ? in->color : I->arg == 2 ? in->normal : in->texcoord;
            for (int k = 0; k < 3; k++) st[sp][k] = src[k];
            sp++;
            break;
        }
        case TINA_OP_TEXTURE: {
            float uv[2] = {st[sp - 1][0], st[sp - 1][1]};
            tex_sample(texhost[I->arg], m->tex_w[I->arg], m->tex_h[I->arg], m->tex_c[I->arg], uv, st[sp - 1]);
            break;
        }
        case TINA_OP_FRESNEL: { /* material.py:69-83 */
            float *specular = st[sp - 1], *albedo = st[sp - 2], *metallic = st[sp - 3];
            for (int k = 0; k < 3; k++)
                metallic[k] = metallic[k] * albedo[k] + (1.0f - metallic[k]) * 0.16f * (specular[k] * specular[k]);
            sp -= 2;
            break;
        }
        case TINA_OP_LAMBERT: { /* material.py:392-393 */
            float v = (float)(1.0 / 3.14159265358979323846); /* Python folds 1 / ti.pi in f64 */
            for (int k = 0; k < 3; k++) st[sp][k] = v;
            sp++;
            break;
        }
        case TINA_OP_PHONG: { /* material.py:450-454, common.py:197-199 */
            float *mm = st[sp - 1];
            float I3[3] = {-idir[0], -idir[1], -idir[2]};
            float t = 2.0f * dot3(nrm, I3);
            float rdir[3];
            for (int k = 0; k < 3; k++) rdir[k] = I3[k] - t * nrm[k];
            float VoR = fmaxf(0.0f, dot3(odir, rdir));
            for (int k = 0; k < 3; k++) mm[k] = powf(VoR, mm[k]) * (mm[k] + 2.0f) / 2.0f;
            break;
        }
        case TINA_OP_COOK: { /* material.py:323-362 */
            float *f0 = st[sp - 1], *rough = st[sp - 2];
            const float EPS = 1e-10f, eps = 1e-6f;
            float half[3] = {idir[0] + odir[0], idir[1] + odir[1], idir[2] + odir[2]};
            normalize3(half);
            float NoH = fmaxf(EPS, dot3(half, nrm));
            float NoL = fmaxf(EPS, dot3(idir, nrm));
            float NoV = fmaxf(EPS, dot3(odir, nrm));
            float VoH = fminf((float)(1 - 1e-10), fmaxf(EPS, dot3(half, odir)));
            for (int k = 0; k < 3; k++) {
                float alpha2 = fmaxf(eps, rough[k] * rough[k]);
                float denom = 1.0f - (NoH * NoH) * (1.0f - alpha2);
                float ndf = alpha2 / (denom * denom);
                float kk = alpha2 / 2.0f;
                float vdf = 1.0f / ((NoV * kk + 1.0f) - kk);
                vdf *= 1.0f / ((NoL * kk + 1.0f) - kk);
                vdf /= 1.0f * (1.0f - alpha2) + 12.566370614359172f * alpha2; /* common.py:221-223 lerp */
                float fdf = f0[k] + (1.0f - f0[k]) * powf(1.0f - VoH, 5.0f);
                rough[k] = fdf * vdf * ndf;
            }
            sp -= 1;
            break;
        }
        case TINA_OP_MIX: { /* material.py:96-118 */
            float *b = st[sp - 1], *a = st[sp - 2], *fac = st[sp - 3];
            for (int k = 0; k < 3; k++) fac[k] = (1.0f - fac[k]) * a[k] + fac[k] * b[k];
            sp -= 2;
            break;
        }
        case TINA_OP_MUL: { /* material.py:157-176 */
            float *wei = st[sp - 1], *fac = st[sp - 2];
            for (int k = 0; k < 3; k++) fac[k] = fac[k] * wei[k];
            sp -= 1;
            break;
        }
        case TINA_OP_ADD: {
            float *b = st[sp - 1], *a = st[sp - 2];
            for (int k = 0; k < 3; k++) a[k] = a[k] + b[k];
            sp -= 1;
            break;
        }
        case TINA_OP_BCAST: /* uv.x / uv.y of LerpTexture, nodes.py:129-136 */
            op_bcast(st[sp - 1], I->arg);
            break;
        case TINA_OP_CHESS: /* ChessboardTexture's factor, nodes.py:114-126 */
            op_chess(st[sp - 2], st[sp - 1]);
            sp -= 1;
            break;
        }
    }
    for (int k = 0; k < 3; k++) out[k] = sp > 0 ? st[sp - 1][k] : 0.0f;
}

/* advans.py:32-35 */
static float aces(float c) { return c * (2.51f * c + 0.03f) / (c * (2.43f * c + 0.59f) + 0.14f); }

void orc_tonemap(float *img, int64_t n) {
    for (int64_t i = 0; i < n; i++) img[i] = aces(img[i]);
}

/*
 * triangle.py:134-153 + :32-49 (interpolate) + shader.py:82-93,119-131 +
 * lighting.py:84-98.  Writes image[P] for pixels with occup[P] != -1 only.
 * `textures` are HOST pointers replacing mat->tex (which holds device pointers
 * in the product).
 */
void orc_render_color(const float *verts, const float *norms, const float *coors, const int32_t *occup,
                      const float *W2V, const float *V2W, const float *bias, int W, int H, uint32_t flags,
                      const TinaMaterial *mat, const float *const *textures, const TinaLighting *L, float *image,
                      int parallel) {
#pragma omp parallel for schedule(dynamic, 64) if (parallel)
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int64_t P = (int64_t)x * H + y;
            int32_t f = occup[P];
            if (f == -1) continue;
            const float *V9 = verts + (int64_t)f * 9;
            FaceSetup s;
            /* the reference re-reads bcn/can/boo/coo/wsc cached by render_occup (:140-145);
             * recomputing them with the same ops gives the same bits */
            face_setup(V9, W2V, W, H, 0u, &s);
            float wei[3];
            pixel_weights(&s, x, y, bias, wei);
            ShadeInputs in;
            const float *A = V9, *B = V9 + 3, *C = V9 + 6;
            for (int k = 0; k < 3; k++) in.pos[k] = (wei[0] * A[k] + wei[1] * B[k]) + wei[2] * C[k]; /* :33 */
            if (flags & TINA_SMOOTHING) {
                const float *N9 = norms + (int64_t)f * 9;
                for (int k = 0; k < 3; k++) in.normal[k] = (wei[0] * N9[k] + wei[1] * N9[3 + k]) + wei[2] * N9[6 + k];
            } else {
                float e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
                float e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
                cross3(e1, e2, in.normal);
            }
            normalize3(in.normal); /* :41 */
            in.texcoord[0] = in.texcoord[1] = in.texcoord[2] = 0.0f;
            if (flags & TINA_TEXTURING) {
                const float *T6 = coors + (int64_t)f * 6;
                for (int k = 0; k < 2; k++) in.texcoord[k] = (wei[0] * T6[k] + wei[1] * T6[2 + k]) + wei[2] * T6[4 + k];
            }
            in.color[0] = in.color[1] = in.color[2] = 1.0f;
            /* shader.py:82-93 */
            float p[2] = {(float)x + bias[0], (float)y + bias[1]};
            float q[3] = {p[0] / (float)W * 2.0f - 1.0f, p[1] / (float)H * 2.0f - 1.0f, -1.0f};
            float ro[3], ro1[3], rd[3];
            mapply_pos(V2W, q, ro);
            q[2] = 1.0f;
            mapply_pos(V2W, q, ro1);
            for (int k = 0; k < 3; k++) rd[k] = ro1[k] - ro[k];
            normalize3(rd);
            float viewdir[3] = {-rd[0], -rd[1], -rd[2]};
            /* lighting.py:84-98 */
            float res[3] = {0, 0, 0}, tmp[3];
            float zero[3] = {0, 0, 0};
            run_program(mat, textures, mat->n_brdf + mat->n_ambient, mat->n_emission, &in, zero, zero, zero, tmp);
            for (int k = 0; k < 3; k++) res[k] += tmp[k];
            run_program(mat, textures, mat->n_brdf, mat->n_ambient, &in, zero, zero, zero, tmp);
            for (int k = 0; k < 3; k++) res[k] += L->ambient[k] * tmp[k];
            for (int l = 0; l < L->nlights; l++) {
                float ld[3];
                for (int k = 0; k < 3; k++) ld[k] = L->dirs[l][k] - in.pos[k] * L->dirs[l][3];
                float dist = sqrtf(dot3(ld, ld));
                for (int k = 0; k < 3; k++) ld[k] = ld[k] / dist;
                float cos_i = dot3(in.normal, ld);
                if (cos_i > 0) {
                    float d2 = dist * dist;
                    run_program(mat, textures, 0, mat->n_brdf, &in, in.normal, ld, viewdir, tmp);
                    for (int k = 0; k < 3; k++) res[k] += cos_i * (L->colors[l][k] / d2) * tmp[k];
                }
            }
            for (int k = 0; k < 3; k++) image[P * 3 + k] = res[k];
        }
}

/*
 * core/shader.py:21-109 G-buffer sinks for pixels with occup != -1: out[(x*H+y)*ncomp + k].
 * kind: 0 Const(param) 1 Position 2 Depth 3 Normal 4 ViewNormal 5 Texcoord 6 Color 7 Chessboard(param[0])
 *       8 Viewdir 9 Simple
 */
void orc_render_gbuffer(const float *verts, const float *norms, const float *coors, const int32_t *occup,
                        const int32_t *depth, const float *W2V, const float *V2W, const float *bias, int W, int H,
                        uint32_t flags, int kind, const float *param, float *out, int ncomp) {
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int64_t P = (int64_t)x * H + y;
            int32_t f = occup[P];
            if (f == -1) continue;
            const float *V9 = verts + (int64_t)f * 9;
            FaceSetup s;
            face_setup(V9, W2V, W, H, 0u, &s);
            float wei[3];
            pixel_weights(&s, x, y, bias, wei);
            const float *A = V9, *B = V9 + 3, *C = V9 + 6;
            float pos[3], nrm[3], uv[2] = {0, 0}, v[3] = {0, 0, 0};
            for (int k = 0; k < 3; k++) pos[k] = (wei[0] * A[k] + wei[1] * B[k]) + wei[2] * C[k];
            if (flags & TINA_SMOOTHING) {
                const float *N9 = norms + (int64_t)f * 9;
                for (int k = 0; k < 3; k++) nrm[k] = (wei[0] * N9[k] + wei[1] * N9[3 + k]) + wei[2] * N9[6 + k];
            } else {
                float e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
                cross3(e1, e2, nrm);
            }
            normalize3(nrm);
            if (flags & TINA_TEXTURING) {
                const float *T6 = coors + (int64_t)f * 6;
                for (int k = 0; k < 2; k++) uv[k] = (wei[0] * T6[k] + wei[1] * T6[2 + k]) + wei[2] * T6[4 + k];
            }
            float p[2] = {(float)x + bias[0], (float)y + bias[1]};
            float q[3] = {p[0] / (float)W * 2.0f - 1.0f, p[1] / (float)H * 2.0f - 1.0f, -1.0f}, ro[3], ro1[3], vd[3];
            mapply_pos(V2W, q, ro);
            q[2] = 1.0f;
            mapply_pos(V2W, q, ro1);
            for (int k = 0; k < 3; k++) vd[k] = ro1[k] - ro[k];
            normalize3(vd);
            for (int k = 0; k < 3; k++) vd[k] = -vd[k];
            switch (kind) {
            case 0: v[0] = param[0], v[1] = param[1], v[2] = param[2]; break;
            case 1: memcpy(v, pos, sizeof v); break;
            case 2: v[0] = v[1] = v[2] = (float)depth[P]; break;
            case 3: memcpy(v, nrm, sizeof v); break;
            case 4: { /* shader.py:51-58 mapply_dir(W2V, normal).normalized() */
                float r[3], rw;
                mapply(W2V, nrm, 0.0f, r, &rw);
                normalize3(r);
                memcpy(v, r, sizeof v);
                break;
            }
            case 5: v[0] = uv[0], v[1] = uv[1]; break;
            case 6: v[0] = v[1] = v[2] = 1.0f; break;
            case 7: { /* shader.py:73-79 */
                float fac = fmodf(floorf(p[0] / param[0]) + floorf(p[1] / param[0]), 2.0f);
                if (fac < 0) fac += 2.0f;
                v[0] = v[1] = v[2] = 0.4f * (1.0f - fac) + 0.9f * fac;
                break;
            }
            case 8: for (int k = 0; k < 3; k++) v[k] = vd[k] * 0.5f + 0.5f; break;
            default: v[0] = v[1] = v[2] = fabsf(dot3(nrm, vd)); break;
            }
            for (int k = 0; k < ncomp; k++) out[P * ncomp + k] = v[k];
        }
}

/* ---- ParticleRaster (core/particle.py) -------------------------------------------------- */
static void mapply_dir_n(const float *M, float d0, float d1, float d2, float *r) { /* common.py:186-189 + normalized */
    float d[3] = {d0, d1, d2}, rw;
    mapply(M, d, 0.0f, r, &rw);
    normalize3(r);
}

/* particle.py:78-127, serial order; flags: 2 = clipping.  depth persists, occup is reset. */
void orc_pars_occup(const float *verts, const float *sizes, int64_t npars, const float *W2V, const float *V2W,
                    const float *bias, int W, int H, uint32_t flags, int32_t *depth, int32_t *occup) {
    for (int64_t i = 0; i < (int64_t)W * H; i++) occup[i] = -1;
    float DXl[3], DYl[3];
    mapply_dir_n(V2W, 1, 0, 0, DXl);
    mapply_dir_n(V2W, 0, 1, 0, DYl);
    float res[2] = {(float)W, (float)H};
    int resi[2] = {W, H};
    for (int64_t f = 0; f < npars; f++) {
        const float *Al = verts + f * 3;
        float Rl = sizes[f], Av[3], t[3], q[3];
        mapply_pos(W2V, Al, Av);
        if ((flags & 2u) && !(-1.0f <= Av[2] && Av[2] <= 1.0f)) continue;
        float Rv[2];
        for (int k = 0; k < 3; k++) q[k] = Al[k] + DXl[k] * Rl;
        mapply_pos(W2V, q, t);
        Rv[0] = t[0] - Av[0];
        for (int k = 0; k < 3; k++) q[k] = Al[k] + DYl[k] * Rl;
        mapply_pos(W2V, q, t);
        Rv[1] = t[1] - Av[1];
        float Bv[4][2] = {{Av[0] - Rv[0], Av[1] - 0.0f}, {Av[0] + Rv[0], Av[1] + 0.0f}, {Av[0] - 0.0f, Av[1] - Rv[1]}, {Av[0] + 0.0f, Av[1] + Rv[1]}};
        float b[4][2];
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 2; k++) b[j][k] = (Bv[j][k] * 0.5f + 0.5f) * res[k];
        int32_t bot[2], top[2];
        for (int k = 0; k < 2; k++) {
            int32_t lo = ifloor_(fminf(b[0][k], b[2][k])), hi = iceil_(fmaxf(b[1][k], b[3][k]));
            bot[k] = lo > 0 ? lo : 0;
            top[k] = hi < resi[k] - 1 ? hi : resi[k] - 1;
        }
        int32_t d = f2i(Av[2] * MAXDEPTH_F);
        for (int32_t x = bot[0]; x <= top[0]; x++)
            for (int32_t y = bot[1]; y <= top[1]; y++) {
                float p[2] = {(float)x + bias[0], (float)y + bias[1]};
                float Pv[3] = {p[0] / res[0] * 2.0f - 1.0f, p[1] / res[1] * 2.0f - 1.0f, Av[2]}, Pl[3];
                mapply_pos(V2W, Pv, Pl);
                float e[3] = {Pl[0] - Al[0], Pl[1] - Al[1], Pl[2] - Al[2]};
                if ((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2] > Rl * Rl) continue;
                int64_t P = (int64_t)x * H + y;
                if (depth[P] > d) {
                    depth[P] = d;
                    occup[P] = (int32_t)f;
                }
            }
    }
}

/* particle.py:129-161 + shader.py:119-131 + lighting.py:84-98 */
void orc_pars_color(const float *verts, const float *sizes, const float *colors, const int32_t *occup, const float *W2V,
                    const float *V2W, const float *bias, int W, int H, const TinaMaterial *mat,
                    const float *const *textures, const TinaLighting *L, float *image) {
    float Zl[3];
    mapply_dir_n(V2W, 0, 0, 1, Zl);
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int64_t P = (int64_t)x * H + y;
            int32_t f = occup[P];
            if (f == -1) continue;
            const float *Al = verts + (int64_t)f * 3;
            float Rl = sizes[f], Av[3];
            mapply_pos(W2V, Al, Av);
            float p[2] = {(float)x + bias[0], (float)y + bias[1]};
            float Pv[3] = {p[0] / (float)W * 2.0f - 1.0f, p[1] / (float)H * 2.0f - 1.0f, Av[2]}, Pl[3], Dl[3];
            mapply_pos(V2W, Pv, Pl);
            for (int k = 0; k < 3; k++) Dl[k] = (Pl[k] - Al[k]) / Rl;
            float t = sqrtf(1.0f - dot3(Dl, Dl));
            for (int k = 0; k < 3; k++) Dl[k] = Dl[k] - Zl[k] * t;
            normalize3(Dl);
            ShadeInputs in;
            for (int k = 0; k < 3; k++) {
                in.normal[k] = Dl[k];
                in.pos[k] = Al[k] + Dl[k] * Rl;
                in.texcoord[k] = 0.0f;
                in.color[k] = colors ? colors[(int64_t)f * 3 + k] : 1.0f;
            }
            float q[3] = {Pv[0], Pv[1], -1.0f}, ro[3], ro1[3], rd[3];
            mapply_pos(V2W, q, ro);
            q[2] = 1.0f;
            mapply_pos(V2W, q, ro1);
            for (int k = 0; k < 3; k++) rd[k] = ro1[k] - ro[k];
            normalize3(rd);
            float viewdir[3] = {-rd[0], -rd[1], -rd[2]};
            float res[3] = {0, 0, 0}, tmp[3], zero[3] = {0, 0, 0};
            run_program(mat, textures, mat->n_brdf + mat->n_ambient, mat->n_emission, &in, zero, zero, zero, tmp);
            for (int k = 0; k < 3; k++) res[k] += tmp[k];
            run_program(mat, textures, mat->n_brdf, mat->n_ambient, &in, zero, zero, zero, tmp);
            for (int k = 0; k < 3; k++) res[k] += L->ambient[k] * tmp[k];
            for (int l = 0; l < L->nlights; l++) {
                float ld[3];
                for (int k = 0; k < 3; k++) ld[k] = L->dirs[l][k] - in.pos[k] * L->dirs[l][3];
                float dist = sqrtf(dot3(ld, ld));
                for (int k = 0; k < 3; k++) ld[k] = ld[k] / dist;
                float cos_i = dot3(in.normal, ld);
                if (cos_i > 0) {
                    float d2 = dist * dist;
                    run_program(mat, textures, 0, mat->n_brdf, &in, in.normal, ld, viewdir, tmp);
                    for (int k = 0; k < 3; k++) res[k] += cos_i * (L->colors[l][k] / d2) * tmp[k];
                }
            }
            for (int k = 0; k < 3; k++) image[P * 3 + k] = res[k];
        }
}

/* particle.py:129-158: the shader inputs of the visible sphere point (what a NormalShader / PositionShader in the
 * particle's ShaderGroup receives); pos, normal: [W][H][3], written where occup != -1 */
void orc_pars_attrs(const float *verts, const float *sizes, const int32_t *occup, const float *W2V, const float *V2W,
                    const float *bias, int W, int H, float *pos, float *normal) {
    float Zl[3];
    mapply_dir_n(V2W, 0, 0, 1, Zl);
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int64_t P = (int64_t)x * H + y;
            int32_t f = occup[P];
            if (f == -1) continue;
            const float *Al = verts + (int64_t)f * 3;
            float Rl = sizes[f], Av[3];
            mapply_pos(W2V, Al, Av);
            float p[2] = {(float)x + bias[0], (float)y + bias[1]};
            float Pv[3] = {p[0] / (float)W * 2.0f - 1.0f, p[1] / (float)H * 2.0f - 1.0f, Av[2]}, Pl[3], Dl[3];
            mapply_pos(V2W, Pv, Pl);
            for (int k = 0; k < 3; k++) Dl[k] = (Pl[k] - Al[k]) / Rl;
            float t = sqrtf(1.0f - dot3(Dl, Dl));
            for (int k = 0; k < 3; k++) Dl[k] = Dl[k] - Zl[k] * t;
            normalize3(Dl);
            for (int k = 0; k < 3; k++) normal[P * 3 + k] = Dl[k], pos[P * 3 + k] = Al[k] + Dl[k] * Rl;
        }
}

/* ---- WireframeRaster (core/wireframe.py:49-95), serial order; flags: 2 = clipping.  verts [N][2][3].
 * On a depth win the pixel of `image` becomes img * (1 - 1) + color * 1 (Shader.blend_color, shader.py:133-135). */
void orc_wire_render(const float *verts, int64_t nwires, const float *W2V, const float *bias, int W, int H, uint32_t flags,
                     const float *color, int32_t *depth, float *image) {
    float res[2] = {(float)W, (float)H};
    for (int64_t f = 0; f < nwires; f++) {
        const float *Al = verts + f * 6, *Bl = Al + 3;
        float Av[3], Bv[3], t[3], rwa, rwb;
        mapply(W2V, Al, 1.0f, t, &rwa);
        for (int k = 0; k < 3; k++) Av[k] = t[k] / rwa;
        mapply(W2V, Bl, 1.0f, t, &rwb);
        for (int k = 0; k < 3; k++) Bv[k] = t[k] / rwb;
        if (flags & 2u) {
            int ina = 1, inb = 1;
            for (int k = 0; k < 3; k++) {
                ina &= (-1.0f <= Av[k]) && (Av[k] <= 1.0f);
                inb &= (-1.0f <= Bv[k]) && (Bv[k] <= 1.0f);
            }
            if (!ina && !inb) continue;
        }
        float a[2], b[2];
        for (int k = 0; k < 2; k++) a[k] = (Av[k] * 0.5f + 0.5f) * res[k], b[k] = (Bv[k] * 0.5f + 0.5f) * res[k];
        float wsc[2] = {1.0f / rwa, 1.0f / rwb};
        float dlt[2] = {b[0] - a[0], b[1] - a[1]}, adlt[2] = {fabsf(dlt[0]), fabsf(dlt[1])}, k[2] = {1.0f, 1.0f};
        int32_t siz;
        if (adlt[0] >= adlt[1]) {
            k[0] = dlt[0] >= 0 ? 1.0f : -1.0f;
            k[1] = k[0] * dlt[1] / dlt[0];
            siz = f2i(adlt[0]);
        } else {
            k[1] = dlt[1] >= 0 ? 1.0f : -1.0f;
            k[0] = k[1] * dlt[0] / dlt[1];
            siz = f2i(adlt[1]);
        }
        for (int64_t i = 0; i < (int64_t)siz + 1; i++) {
            float fi = (float)i;
            float pos[2] = {(a[0] + k[0] * fi) + bias[0], (a[1] + k[1] * fi) + bias[1]};
            int32_t Px = ifloor_(pos[0]), Py = ifloor_(pos[1]);
            if (!(0 <= Px && Px < W && 0 <= Py && Py < H)) continue;
            float cor = fi / (float)siz;
            float wei[2] = {(1.0f - cor) * wsc[0], cor * wsc[1]};
            float sum = wei[0] + wei[1];
            wei[0] = wei[0] / sum, wei[1] = wei[1] / sum;
            int32_t d = f2i((wei[0] * Av[2] + wei[1] * Bv[2]) * MAXDEPTH_F);
            int64_t P = (int64_t)Px * H + Py;
            if (depth[P] > d) {
                depth[P] = d;
                if (image)
                    for (int c = 0; c < 3; c++) image[P * 3 + c] = image[P * 3 + c] * 0.0f + color[c] * 1.0f;
            }
        }
    }
}

/* ---- post effects (postp/fxaa.py:28-68, postp/blooming.py:45-68); x-major [W][H]; out-of-shape reads = 0 ---- */
static float ld2(const float *f, int W, int H, int x, int y) { return (x < 0 || y < 0 || x >= W || y >= H) ? 0.0f : f[(int64_t)x * H + y]; }
static void ld2v(const float *f, int W, int H, int x, int y, float *o) {
    if (x < 0 || y < 0 || x >= W || y >= H) {
        o[0] = o[1] = o[2] = 0.0f;
        return;
    }
    memcpy(o, f + ((int64_t)x * H + y) * 3, 3 * sizeof(float));
}
static void bilerp3(const float *f, int W, int H, float px, float py, float *o) { /* common.py:140-149 */
    int32_t I0 = ifloor_(px), I1 = ifloor_(py);
    float x0 = px - (float)I0, x1 = py - (float)I1, y0 = 1.0f - x0, y1 = 1.0f - x1, a[3], b[3], c[3], d[3];
    ld2v(f, W, H, I0 + 1, I1 + 1, a), ld2v(f, W, H, I0 + 1, I1, b), ld2v(f, W, H, I0, I1, c), ld2v(f, W, H, I0, I1 + 1, d);
    for (int k = 0; k < 3; k++) o[k] = ((a[k] * x0 * x1 + b[k] * x0 * y1) + c[k] * y0 * y1) + d[k] * y0 * x1;
}
static float clamp01(float x) { return fminf(1.0f, fmaxf(0.0f, x)); }

void orc_fxaa(float *image, int W, int H, float abs_thresh, float rel_thresh, float factor) {
    int64_t npix = (int64_t)W * H;
    float *lumi = malloc(sizeof(float) * npix), *copy = malloc(sizeof(float) * npix * 3);
    for (int64_t i = 0; i < npix; i++) {
        lumi[i] = clamp01((0.2989f * image[i * 3] + 0.587f * image[i * 3 + 1]) + 0.114f * image[i * 3 + 2]);
        memcpy(copy + i * 3, image + i * 3, 3 * sizeof(float));
    }
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int64_t i = (int64_t)x * H + y;
            float m = lumi[i], n = ld2(lumi, W, H, x, y + 1), e = ld2(lumi, W, H, x + 1, y), s = ld2(lumi, W, H, x, y - 1);
            float w = ld2(lumi, W, H, x - 1, y), ne = ld2(lumi, W, H, x + 1, y + 1), nw = ld2(lumi, W, H, x - 1, y + 1);
            float se = ld2(lumi, W, H, x + 1, y - 1), sw = ld2(lumi, W, H, x - 1, y - 1);
            float hi = fmaxf(fmaxf(fmaxf(fmaxf(m, n), e), s), w), lo = fminf(fminf(fminf(fminf(m, n), e), s), w);
            float c = hi - lo;
            if (c < abs_thresh || c < rel_thresh * hi) continue;
            float filt = 2.0f * (((n + e) + s) + w);
            filt += ((ne + nw) + se) + sw;
            filt = fabsf(filt / 12.0f - m);
            filt = clamp01(filt / c);
            float t = clamp01((filt - 0.0f) / (1.0f - 0.0f)), sm = t * t * (3.0f - 2.0f * t);
            float blend = (sm * sm) * factor;
            float hori = fabsf((n + s) - 2.0f * m) * 2.0f;
            hori += fabsf((ne + se) - 2.0f * e);
            hori += fabsf((nw + sw) - 2.0f * w);
            float vert = fabsf((e + w) - 2.0f * m) * 2.0f;
            vert += fabsf((ne + nw) - 2.0f * n);
            vert += fabsf((se + sw) - 2.0f * s);
            int is_hori = hori >= vert;
            float plumi = is_hori ? n : e, nlumi = is_hori ? s : w;
            if (fabsf(plumi - m) < fabsf(nlumi - m)) blend = -blend;
            bilerp3(copy, W, H, (float)x + blend * (is_hori ? 0.0f : 1.0f), (float)y + blend * (is_hori ? 1.0f : 0.0f), image + i * 3);
        }
    free(lumi), free(copy);
}

static float bloom_filter(float x, float thresh, float scale, float factor) {
    float t = fmaxf(0.0f, x - thresh);
    t = 1.0f - 1.0f / (1.0f + scale * t);
    return factor * t;
}
void orc_bloom(float *image, int W, int H, const float *gwei, int radius, float thresh, float scale, float factor) {
    int hw = W / 2, hh = H / 2;
    float *a = malloc(sizeof(float) * hw * hh * 3), *b = malloc(sizeof(float) * hw * hh * 3);
    for (int x = 0; x < hw; x++)
        for (int y = 0; y < hh; y++) {
            float r[3] = {0, 0, 0}, c[3];
            for (int jx = 0; jx < 2; jx++)
                for (int jy = 0; jy < 2; jy++) {
                    ld2v(image, W, H, x * 2 + jx, y * 2 + jy, c);
                    for (int k = 0; k < 3; k++) r[k] += bloom_filter(c[k], thresh, scale, factor);
                }
            for (int k = 0; k < 3; k++) a[((int64_t)x * hh + y) * 3 + k] = r[k] / 4.0f;
        }
    for (int axis = 0; axis < 2; axis++) {
        const float *src = axis ? b : a;
        float *dst = axis ? a : b;
        for (int x = 0; x < hw; x++)
            for (int y = 0; y < hh; y++)
                for (int k = 0; k < 3; k++) {
                    float r = src[((int64_t)x * hh + y) * 3 + k] * gwei[0];
                    for (int i = 1; i <= radius; i++) {
                        int xa = axis ? x : (x - i > 0 ? x - i : 0), ya = axis ? (y - i > 0 ? y - i : 0) : y;
                        int xb = axis ? x : (x + i < hw - 1 ? x + i : hw - 1), yb = axis ? (y + i < hh - 1 ? y + i : hh - 1) : y;
                        float v = src[((int64_t)xa * hh + ya) * 3 + k];
                        v += src[((int64_t)xb * hh + yb) * 3 + k];
                        r += v * gwei[i];
                    }
                    dst[((int64_t)x * hh + y) * 3 + k] = r;
                }
    }
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            float o[3];
            bilerp3(a, hw, hh, (float)x / 2.0f, (float)y / 2.0f, o);
            for (int k = 0; k < 3; k++) image[((int64_t)x * H + y) * 3 + k] += o[k];
        }
    free(a), free(b);
}

/* ---- mesh providers feeding set_object ------------------------------------------ */

/* mesh/grid.py:26-35 MeshGrid.pre_compute; pos, nrm: [nx][ny][3] */
/* postp/ssao.py:65-96 render_at for every pixel (non-TAA mode: fixed sample / rotation tables, :24-36).
 * depth: int32 [W][H] (engine.depth), normals: [W][H][3] (the scene's NormalShader buffer), out ao: [W][H] */
void orc_ssao_render(const int32_t *depth, const float *normals, const float *W2V, const float *V2W, const float *bias,
                     int W, int H, const float *samples, int nsamples, const float *rotations, int noise, float radius,
                     float thresh, float factor, float *ao) {
    const float up[3] = {233.0f, 666.0f, 512.0f}; /* advans.py:97 */
    for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) {
            const int64_t P = (int64_t)i * H + j;
            const float *normal = normals + P * 3;
            const float p[2] = {(float)i + bias[0], (float)j + bias[1]};
            float vpos[3] = {p[0] / (float)W * 2.0f - 1.0f, p[1] / (float)H * 2.0f - 1.0f, (float)depth[P] / 1073741824.0f};
            float pos[3];
            mapply_pos(V2W, vpos, pos);
            /* shader.py:82-93 calc_viewdir */
            float q[3] = {vpos[0], vpos[1], -1.0f}, ro[3], ro1[3], rd[3];
            mapply_pos(V2W, q, ro);
            q[2] = 1.0f;
            mapply_pos(V2W, q, ro1);
            for (int k = 0; k < 3; k++) rd[k] = ro1[k] - ro[k];
            normalize3(rd);
            const float viewdir[3] = {-rd[0], -rd[1], -rd[2]};
            float t3[3], tv[3];
            for (int k = 0; k < 3; k++) t3[k] = pos[k] - radius * viewdir[k];
            mapply_pos(W2V, t3, tv);
            const float vradius = tv[2] - vpos[2];
            /* advans.py:97-100 tangentspace(normal) */
            float bitan[3], tan[3];
            cross3(normal, up, bitan);
            normalize3(bitan);
            cross3(bitan, normal, tan);
            const float *rot = rotations + ((int64_t)(i % noise) * noise + (j % noise)) * 2;
            float occ = 0.0f;
            for (int s = 0; s < nsamples; s++) {
                const float *sm = samples + (int64_t)s * 3;
                const float sx = rot[0] * sm[0] + rot[1] * sm[1], sy = -rot[0] * sm[0] + rot[1] * sm[1], sz = sm[2]; /* :86-88 */
                float sp[3], sv[3];
                for (int k = 0; k < 3; k++) sp[k] = pos[k] + ((tan[k] * sx + bitan[k] * sy) + normal[k] * sz) * radius;
                mapply_pos(W2V, sp, sv);
                const float Dx = (sv[0] * 0.5f + 0.5f) * (float)W, Dy = (sv[1] * 0.5f + 0.5f) * (float)H;
                if (0.0f <= Dx && Dx < (float)W && 0.0f <= Dy && Dy < (float)H) {
                    const float d = (float)depth[(int64_t)f2i(Dx) * H + f2i(Dy)] / 1073741824.0f;
                    if (d < sv[2]) {
                        float rc = vradius / (vpos[2] - d);
                        const float t = clamp01((fabsf(rc) - 0.0f) / (1.0f - 0.0f)); /* common.py:216-218 smoothstep */
                        occ += t * t * (3.0f - 2.0f * t);
                    }
                }
            }
            float a = occ / (float)nsamples;
            a = factor * (a - thresh);
            ao[P] = clamp01(a);
        }
}

/* postp/ssao.py:38-49 apply (non-TAA): out *= 1 - box(noise x noise)(ao) / noise^2; reads outside the field are 0 */
void orc_ssao_apply(float *image, const float *ao, int W, int H, int noise) {
    const int offs = noise / 2;
    for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) {
            float r = 0.0f;
            for (int k = 0; k < noise; k++)
                for (int l = 0; l < noise; l++) r += ld2(ao, W, H, i + k - offs, j + l - offs);
            const float f = 1.0f - r / (float)(noise * noise);
            for (int c = 0; c < 3; c++) image[((int64_t)i * H + j) * 3 + c] *= f;
        }
}

/* ---- SSR (postp/ssr.py) and the random streams it draws from (tina/random.py) ------------------------------------ */

/* random.py:26-35 WangHashRNG._noise */
static uint32_t wang(uint32_t v) {
    v = (v ^ 61u) ^ (v >> 16);
    v *= 9u;
    v ^= v << 4;
    v *= 0x27d4eb2du;
    v ^= v >> 15;
    return v;
}
typedef struct {
    uint32_t seed;
} WangRNG;
/* random.py:54-58: noise(seed) = (noise_int(seed) >> 1) * (2 / 4294967296) -- u32 -> f32 (rounded), times 2^-31; seed += 1 */
static float rng_random(WangRNG *r) {
    const uint32_t u = wang(r->seed) >> 1;
    r->seed += 1u;
    return (float)u * 4.656612873077393e-10f;
}
static uint32_t rng_random_int(WangRNG *r) { /* random.py:60-64 */
    const uint32_t u = wang(r->seed);
    r->seed += 1u;
    return u;
}

/* a parameter program of a TinaSampleMaterial: TINA_OP_CONST / INPUT / TEXTURE / FRESNEL only (matr/nodes.py, material.py:69-83) */
static void run_value(const TinaSampleMaterial *m, const float *const *texhost, int begin, int n, const ShadeInputs *in, float *out) {
    float st[STK][3];
    int sp = 0;
    for (int pc = begin; pc < begin + n; pc++) {
        const TinaInstr *I = &m->code[pc];
        switch (I->op) {
        case TINA_OP_CONST:
            for (int k = 0; k < 3; k++) st[sp][k] = I->c[k];
            sp++;
            break;
        case TINA_OP_INPUT: {
            const float *src = I->arg == 0 ? in->pos : I->arg == 1 ? in->color : I->arg == 2 ? in->normal : in->texcoord;
            for (int k = 0; k < 3; k++) st[sp][k] = src[k];
            sp++;
            break;
        }
        case TINA_OP_TEXTURE: {
            float uv[2] = {st[sp - 1][0], st[sp - 1][1]};
            tex_sample(texhost[I->arg], m->tex_w[I->arg], m->tex_h[I->arg], m->tex_c[I->arg], uv, st[sp - 1]);
            break;
        }
        case TINA_OP_FRESNEL: {
            float *specular = st[sp - 1], *albedo = st[sp - 2], *metallic = st[sp - 3];
            for (int k = 0; k < 3; k++)
                metallic[k] = metallic[k] * albedo[k] + (1.0f - metallic[k]) * 0.16f * (specular[k] * specular[k]);
            sp -= 2;
            break;
        }
        case TINA_OP_MIX: {
            float *b = st[sp - 1], *a = st[sp - 2], *fac = st[sp - 3];
            for (int k = 0; k < 3; k++) fac[k] = (1.0f - fac[k]) * a[k] + fac[k] * b[k];
            sp -= 2;
            break;
        }
        case TINA_OP_ADD: {
            float *b = st[sp - 1], *a = st[sp - 2];
            for (int k = 0; k < 3; k++) a[k] = a[k] + b[k];
            sp -= 1;
            break;
        }
        case TINA_OP_BCAST:
            op_bcast(st[sp - 1], I->arg);
            break;
        case TINA_OP_CHESS:
            op_chess(st[sp - 2], st[sp - 1]);
            sp -= 1;
            break;
        }
    }
    for (int k = 0; k < 3; k++) out[k] = sp > 0 ? st[sp - 1][k] : 0.0f;
}

/* advans.py:97-100 tangentspace(n) @ advans.py:105-108 spherical(h, p) */
static void tangent_spherical(const float *n, float h, float p, float *o) {
    const float up[3] = {233.0f, 666.0f, 512.0f};
    float bitan[3], tan[3];
    cross3(n, up, bitan);
    normalize3(bitan);
    cross3(bitan, n, tan);
    const float ang = p * 6.283185307179586f;
    const float s = sqrtf(fmaxf(0.0f, 1.0f - h * h));
    const float ux = s * cosf(ang), uy = s * sinf(ang);
    for (int k = 0; k < 3; k++) o[k] = (tan[k] * ux + bitan[k] * uy) + n[k] * h;
}
static void reflect_neg(const float *idir, const float *nrm, float *r) { /* common.py:198-199 reflect(-idir, nrm) */
    const float I3[3] = {-idir[0], -idir[1], -idir[2]};
    const float t = 2.0f * dot3(nrm, I3);
    for (int k = 0; k < 3; k++) r[k] = I3[k] - t * nrm[k];
}

/* material.sample(idir, nrm, 1, rng) (matr/material.py) for the sub-tree rooted at `node`: -> odir, wei (rough is unused by SSR) */
static void sample_node(const TinaSampleMaterial *m, const float *const *texhost, int node, const ShadeInputs *in,
                        const float *idir, const float *nrm, WangRNG *rng, float *odir, float *wei) {
    const TinaSampleNode *N = &m->nodes[node];
    float p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0};
    if (N->n0) run_value(m, texhost, N->p0, N->n0, in, p0);
    if (N->n1) run_value(m, texhost, N->p1, N->n1, in, p1);
    switch (N->kind) {
    case TINA_SNODE_LAMBERT: { /* material.py:398-405 */
        float u = rng_random(rng);
        const float v = rng_random(rng);
        u = sqrtf(u);
        tangent_spherical(nrm, u, v, odir);
        normalize3(odir);
        wei[0] = wei[1] = wei[2] = 1.0f;
        break;
    }
    case TINA_SNODE_PHONG: { /* :459-472; the shineness is a scalar there */
        const float mm = p0[0];
        float u = rng_random(rng);
        const float v = rng_random(rng);
        u = powf(u, 1.0f / (mm + 1.0f));
        float rdir[3];
        reflect_neg(idir, nrm, rdir);
        tangent_spherical(rdir, u, v, odir);
        float w = 1.0f;
        if (dot3(odir, nrm) < 0.0f) {
            for (int k = 0; k < 3; k++) odir[k] = -odir[k];
            w = 0.0f;
        }
        wei[0] = wei[1] = wei[2] = w;
        break;
    }
    case TINA_SNODE_COOK: { /* :364-384 sample, :323-357 sub_brdf */
        const float EPS = 1e-10f, eps = 1e-6f;
        const float alpha2s = fmaxf(0.0f, p0[0] * p0[0]);
        float u = rng_random(rng);
        const float v = rng_random(rng);
        u = sqrtf((1.0f - u) / (1.0f - u * (1.0f - alpha2s)));
        float rdir[3];
        reflect_neg(idir, nrm, rdir);
        tangent_spherical(rdir, u, v, odir);
        float half[3] = {idir[0] + odir[0], idir[1] + odir[1], idir[2] + odir[2]};
        normalize3(half);
        const float NoL = fmaxf(EPS, dot3(idir, nrm));
        const float NoV = fmaxf(EPS, dot3(odir, nrm));
        const float VoH = fminf((float)(1 - 1e-10), fmaxf(EPS, dot3(half, odir)));
        const int flip = dot3(odir, nrm) < 0.0f;
        for (int k = 0; k < 3; k++) {
            const float alpha2 = fmaxf(eps, p0[k] * p0[k]);
            const float kk = alpha2 / 2.0f;
            float vdf = 1.0f / ((NoV * kk + 1.0f) - kk);
            vdf *= 1.0f / ((NoL * kk + 1.0f) - kk);
            vdf /= 1.0f * (1.0f - alpha2) + 12.566370614359172f * alpha2;
            float fdf = p1[k] + (1.0f - p1[k]) * powf(1.0f - VoH, 5.0f);
            if (flip) fdf = 0.0f;
            wei[k] = fdf * vdf;
        }
        if (flip)
            for (int k = 0; k < 3; k++) odir[k] = -odir[k];
        break;
    }
    case TINA_SNODE_EMISSION: /* :679-681 */
        for (int k = 0; k < 3; k++) odir[k] = idir[k], wei[k] = 0.0f;
        break;
    case TINA_SNODE_MIX: { /* :123-138 */
        /* Vavg (common.py:36-40): a vector parameter is averaged, a scalar one is taken as it is (pad_ bit 0) */
        float avg = (N->pad_ & 1) ? p0[0] : ((p0[0] + p0[1]) + p0[2]) / 3.0f;
        float factor = avg; /* common.py:227-231 smoothlerp(avg, 0.12, 0.88) */
        if (factor != 0.0f && factor != 1.0f) {
            const float t = clamp01((factor - 0.0f) / (1.0f - 0.0f));
            factor = t * t * (3.0f - 2.0f * t);
            factor = 0.12f * (1.0f - factor) + 0.88f * factor;
        }
        if (rng_random(rng) < factor) {
            sample_node(m, texhost, N->b, in, idir, nrm, rng, odir, wei);
            for (int k = 0; k < 3; k++) wei[k] *= p0[k] / factor;
        } else {
            sample_node(m, texhost, N->a, in, idir, nrm, rng, odir, wei);
            for (int k = 0; k < 3; k++) wei[k] *= (1.0f - p0[k]) / (1.0f - factor);
        }
        break;
    }
    case TINA_SNODE_SCALE: /* :180-184 */
        sample_node(m, texhost, N->a, in, idir, nrm, rng, odir, wei);
        for (int k = 0; k < 3; k++) wei[k] = wei[k] * p0[k];
        break;
    case TINA_SNODE_ADD: /* :227-238 */
        sample_node(m, texhost, (rng_random_int(rng) % 2u == 0u) ? N->a : N->b, in, idir, nrm, rng, odir, wei);
        for (int k = 0; k < 3; k++) wei[k] *= 2.0f;
        break;
    }
}

/* postp/ssr.py:44-103 render / render_at.  depth int32 [W][H], normals [W][H][3], coors [W][H][2] or NULL, mtlid int32
 * [W][H], image [W][H][3]; out4 [W][H][4].  texhost[i] = host texture pointers of table[i] (TINA_MAX_TEX each).
 * taa = 0: rng = WangHashRNG(P % blurring) (:73-76); taa = 1: the same hash seeded with (P.x, P.y, frame) (the
 * reference's ti.random() stream is unspecified, include/tina_b200.h). */
void orc_ssr_render(const int32_t *depth, const float *normals, const float *coors, const int32_t *mtlid,
                    const TinaSampleMaterial *table, const float *const *texhost, int nmaterials, const float *image,
                    const float *W2V, const float *V2W, const float *bias, int W, int H, int nsamples, int nsteps,
                    float stepsize, float tolerance, int blurring, int taa, uint32_t frame, float *out4) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) {
            const int64_t P = (int64_t)i * H + j;
            float *o = out4 + P * 4;
            const float *normal = normals + P * 3;
            o[0] = o[1] = o[2] = o[3] = 0.0f;
            if ((normal[0] * normal[0] + normal[1] * normal[1]) + normal[2] * normal[2] < 1e-6f) continue; /* :47-48 */
            const int mid = mtlid[P];
            if (mid < 0 || mid >= nmaterials) continue; /* (VirtualMaterial: no branch taken -> zeros) */
            const TinaSampleMaterial *m = table + mid;
            const float *const *th = texhost + (int64_t)mid * TINA_MAX_TEX;
            const float p[2] = {(float)i + bias[0], (float)j + bias[1]};
            const float vpos[3] = {p[0] / (float)W * 2.0f - 1.0f, p[1] / (float)H * 2.0f - 1.0f, (float)depth[P] / MAXDEPTH_F};
            float pos[3];
            mapply_pos(V2W, vpos, pos);
            float q[3] = {vpos[0], vpos[1], -1.0f}, r0[3], r1[3], rd[3]; /* shader.py:82-93 calc_viewdir */
            mapply_pos(V2W, q, r0);
            q[2] = 1.0f;
            mapply_pos(V2W, q, r1);
            for (int k = 0; k < 3; k++) rd[k] = r1[k] - r0[k];
            normalize3(rd);
            const float viewdir[3] = {-rd[0], -rd[1], -rd[2]};
            ShadeInputs in;
            for (int k = 0; k < 3; k++) in.pos[k] = pos[k], in.color[k] = 1.0f, in.normal[k] = normal[k];
            in.texcoord[0] = coors ? coors[P * 2] : 0.0f, in.texcoord[1] = coors ? coors[P * 2 + 1] : 0.0f, in.texcoord[2] = 0.0f;
            WangRNG rng;
            if (!taa) rng.seed = wang((uint32_t)(j % blurring) ^ wang((uint32_t)(i % blurring))); /* random.py:44-52 on P % blurring */
            else rng.seed = wang(frame ^ wang((uint32_t)j ^ wang((uint32_t)i)));
            float res[4] = {0, 0, 0, 0};
            for (int s = 0; s < nsamples; s++) {
                float odir[3], wei[3];
                sample_node(m, th, 0, &in, viewdir, normal, &rng, odir, wei);
                const float ov = dot3(odir, viewdir);
                const float step = stepsize / (sqrtf(1.0f - ov * ov) * (float)nsteps);
                float t3[3], v0[3], v1[3];
                for (int k = 0; k < 3; k++) t3[k] = pos[k] - viewdir[k] / (float)nsteps;
                mapply_pos(W2V, t3, v0);
                mapply_pos(W2V, pos, v1);
                const float vtol = tolerance * (v0[2] - v1[2]);
                const float rr = rng_random(&rng);
                float ro[3];
                for (int k = 0; k < 3; k++) ro[k] = pos[k] + odir[k] * rr * step;
                for (int t = 0; t < nsteps; t++) {
                    float vro[3];
                    for (int k = 0; k < 3; k++) ro[k] += odir[k] * step;
                    mapply_pos(W2V, ro, vro);
                    if (!(-1.0f <= vro[0] && vro[0] <= 1.0f && -1.0f <= vro[1] && vro[1] <= 1.0f && -1.0f <= vro[2] && vro[2] <= 1.0f)) break;
                    const float Dx = (vro[0] * 0.5f + 0.5f) * (float)W, Dy = (vro[1] * 0.5f + 0.5f) * (float)H;
                    const int ix = f2i(Dx), iy = f2i(Dy);
                    const float d = (ix < 0 || iy < 0 || ix >= W || iy >= H) ? 0.0f : (float)depth[(int64_t)ix * H + iy] / MAXDEPTH_F;
                    if (vro[2] - vtol < d && d < vro[2]) {
                        float clr[3];
                        bilerp3(image, W, H, Dx, Dy, clr);
                        for (int k = 0; k < 3; k++) res[k] += clr[k] * wei[k];
                        res[3] += 1.0f;
                        break;
                    }
                }
            }
            for (int k = 0; k < 4; k++) o[k] = res[k] / (float)nsamples;
        }
}

/* postp/ssr.py:30-42 apply: res = img4 (taa) or its blurring x blurring box mean (reads outside the field are 0);
 * image = image * (1 - res.w) + res.xyz */
void orc_ssr_apply(float *image, const float *img4, int W, int H, int blurring, int taa) {
    const int offs = blurring / 2;
    for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) {
            float res[4] = {0, 0, 0, 0};
            if (taa) {
                memcpy(res, img4 + ((int64_t)i * H + j) * 4, sizeof res);
            } else {
                for (int k = 0; k < blurring; k++)
                    for (int l = 0; l < blurring; l++) {
                        const int x = i + k - offs, y = j + l - offs;
                        if (x < 0 || y < 0 || x >= W || y >= H) continue;
                        for (int c = 0; c < 4; c++) res[c] += img4[((int64_t)x * H + y) * 4 + c];
                    }
                for (int c = 0; c < 4; c++) res[c] /= (float)(blurring * blurring);
            }
            float *px = image + ((int64_t)i * H + j) * 3;
            for (int c = 0; c < 3; c++) {
                px[c] *= 1.0f - res[3];
                px[c] += res[c];
            }
        }
}

/* postp/ssao.py:52-56,65-96 with taa=True: make_sample() per sample from the hash stream seeded with (P.x, P.y, frame)
 * (include/tina_b200.h: tina_engine_ssao_render_taa) instead of the rotated table */
void orc_ssao_render_taa(const int32_t *depth, const float *normals, const float *W2V, const float *V2W, const float *bias,
                         int W, int H, int nsamples, float radius, float thresh, float factor, uint32_t frame, float *ao) {
    for (int i = 0; i < W; i++)
        for (int j = 0; j < H; j++) {
            const int64_t P = (int64_t)i * H + j;
            const float *normal = normals + P * 3;
            const float p[2] = {(float)i + bias[0], (float)j + bias[1]};
            float vpos[3] = {p[0] / (float)W * 2.0f - 1.0f, p[1] / (float)H * 2.0f - 1.0f, (float)depth[P] / MAXDEPTH_F};
            float pos[3];
            mapply_pos(V2W, vpos, pos);
            float q[3] = {vpos[0], vpos[1], -1.0f}, ro[3], ro1[3], rd[3];
            mapply_pos(V2W, q, ro);
            q[2] = 1.0f;
            mapply_pos(V2W, q, ro1);
            for (int k = 0; k < 3; k++) rd[k] = ro1[k] - ro[k];
            normalize3(rd);
            const float viewdir[3] = {-rd[0], -rd[1], -rd[2]};
            float t3[3], tv[3];
            for (int k = 0; k < 3; k++) t3[k] = pos[k] - radius * viewdir[k];
            mapply_pos(W2V, t3, tv);
            const float vradius = tv[2] - vpos[2];
            WangRNG rng;
            rng.seed = wang(frame ^ wang((uint32_t)j ^ wang((uint32_t)i)));
            float occ = 0.0f;
            for (int s = 0; s < nsamples; s++) {
                /* ssao.py:52-56 make_sample: u, v = random(), random(); r = lerp(random() ** 1.5, .01, 1); u = lerp(u, .01, 1) */
                float u = rng_random(&rng);
                const float v = rng_random(&rng);
                const float w = powf(rng_random(&rng), 1.5f);
                const float r = 0.01f * (1.0f - w) + 1.0f * w;
                u = 0.01f * (1.0f - u) + 1.0f * u;
                float dir[3], sp[3], sv[3];
                { /* tangentspace(normal) @ (spherical(u, v) * r): the sample is scaled before the basis change */
                    const float up[3] = {233.0f, 666.0f, 512.0f};
                    float bitan[3], tan[3];
                    cross3(normal, up, bitan);
                    normalize3(bitan);
                    cross3(bitan, normal, tan);
                    const float ang = v * 6.283185307179586f, sq = sqrtf(fmaxf(0.0f, 1.0f - u * u));
                    const float sx = sq * cosf(ang) * r, sy = sq * sinf(ang) * r, sz = u * r;
                    for (int k = 0; k < 3; k++) dir[k] = (tan[k] * sx + bitan[k] * sy) + normal[k] * sz;
                }
                for (int k = 0; k < 3; k++) sp[k] = pos[k] + dir[k] * radius;
                mapply_pos(W2V, sp, sv);
                const float Dx = (sv[0] * 0.5f + 0.5f) * (float)W, Dy = (sv[1] * 0.5f + 0.5f) * (float)H;
                if (0.0f <= Dx && Dx < (float)W && 0.0f <= Dy && Dy < (float)H) {
                    const float d = (float)depth[(int64_t)f2i(Dx) * H + f2i(Dy)] / MAXDEPTH_F;
                    if (d < sv[2]) {
                        float rc = vradius / (vpos[2] - d);
                        const float t = clamp01((fabsf(rc) - 0.0f) / (1.0f - 0.0f));
                        occ += t * t * (3.0f - 2.0f * t);
                    }
                }
            }
            float a = occ / (float)nsamples;
            a = factor * (a - thresh);
            ao[P] = clamp01(a);
        }
}

void orc_ssao_apply_taa(float *image, const float *ao, int W, int H) { /* ssao.py:40-41: out *= 1 - img */
    for (int64_t P = 0; P < (int64_t)W * H; P++)
        for (int c = 0; c < 3; c++) image[P * 3 + c] *= 1.0f - ao[P];
}

void orc_grid_normals(const float *pos, int nx, int ny, float *nrm) {
    for (int i = 0; i < nx; i++)
        for (int j = 0; j < ny; j++) {
            int i2 = i - 1 > 0 ? i - 1 : 0, j2 = j - 1 > 0 ? j - 1 : 0;
            int i1 = i + 1 < nx - 1 ? i + 1 : nx - 1, j1 = j + 1 < ny - 1 ? j + 1 : ny - 1;
            float dx[3], dy[3];
            for (int k = 0; k < 3; k++) {
                dy[k] = pos[((int64_t)i * ny + j1) * 3 + k] - pos[((int64_t)i * ny + j2) * 3 + k];
                dx[k] = pos[((int64_t)i1 * ny + j) * 3 + k] - pos[((int64_t)i2 * ny + j) * 3 + k];
            }
            float *o = nrm + ((int64_t)i * ny + j) * 3;
            cross3(dx, dy, o);
            normalize3(o);
        }
}

/* mesh/grid.py:45-58 _get_face_props for a [nx][ny][dim] property -> out[nfaces][3][dim] */
void orc_grid_faces(const float *prop, int nx, int ny, int dim, float *out) {
    int64_t nfaces = 2 * (int64_t)(nx - 1) * (ny - 1);
    int stride = nx - 1; /* sic: res.x - 1 for both div and mod (grid.py:46) */
    for (int64_t n = 0; n < nfaces; n++) {
        int64_t m = n / 2;
        int i = (int)(m / stride), j = (int)(m % stride);
        const float *a = prop + ((int64_t)i * ny + j) * dim, *b = prop + ((int64_t)(i + 1) * ny + j) * dim;
        const float *c = prop + ((int64_t)(i + 1) * ny + j + 1) * dim, *d = prop + ((int64_t)i * ny + j + 1) * dim;
        const float *src[3] = {a, b, c};
        if (n % 2 != 0) src[1] = c, src[2] = d;
        for (int k = 0; k < 3; k++) memcpy(out + (n * 3 + k) * dim, src[k], sizeof(float) * dim);
    }
}

/* mesh/trans.py:28-40 applied to [n][3] arrays; trans 4x4, trans_normal 3x3 row-major */
void orc_transform_verts(float *v, int64_t n, const float *trans) {
    for (int64_t i = 0; i < n; i++) {
        float r[3];
        mapply_pos(trans, v + i * 3, r);
        memcpy(v + i * 3, r, sizeof r);
    }
}
void orc_transform_norms(float *v, int64_t n, const float *tn) {
    for (int64_t i = 0; i < n; i++) {
        float *p = v + i * 3, r[3];
        for (int a = 0; a < 3; a++) r[a] = (tn[a * 3 + 0] * p[0] + tn[a * 3 + 1] * p[1]) + tn[a * 3 + 2] * p[2];
        memcpy(p, r, sizeof r);
    }
}

/* bench.py's reference arm: use `n` host threads whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    extern void omp_set_dynamic(int);
    omp_set_dynamic(0);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
