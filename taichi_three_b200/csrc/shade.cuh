// shade.cuh -- part of the single translation unit tina_b200.cu (included once, in order): K4 k_render_color (+ material ops / interpreter, lean and fast arms, fused sort-last composite) and k_gbuffer.
#pragma once

// ------------------------------------------------------------------------------------
// K4: deferred shading (render_color)
// ------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float a, float b, float c) { return V3{a, b, c}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 normalized(V3 v) { // taichi: invlen = 1/sqrt(norm_sqr); invlen * v
    float inv = 1.0f / sqrtf(dot3(v, v));
    return v3(inv * v.x, inv * v.y, inv * v.z);
}
// ---- relaxed arithmetic for SHADING only (tina_raster_set_tuning(.., TINA_TUNE_FAST_SHADING, 1), the default):
// colour is specified to 1e-4 (north_star), ids and depth to the bit, so everything that decides coverage --
// and the barycentric weights, which are ill-conditioned on slivers -- keeps the reference's exact op order,
// while the well-conditioned rest (interpolation, normalisation, view ray, lighting, tone curve) may contract
// to FMA and use the SFU reciprocal / rsqrt (<= 2 ulp).  Exact shading stays available as the other template arm.
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsq_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fdot3(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
template <bool FAST>
__device__ __forceinline__ V3 normalized_t(V3 v) {
    if (!FAST) return normalized(v);
    const float inv = rsq_fast(fdot3(v, v));
    return v3(inv * v.x, inv * v.y, inv * v.z);
}
__device__ __forceinline__ V3 mapply_pos3(const float *M, float p0, float p1, float p2) {
    float r0, r1, r2, rw;
    mapply(M, p0, p1, p2, 1.0f, r0, r1, r2, rw);
    return v3(fd(r0, rw), fd(r1, rw), fd(r2, rw));
}

struct ShadeIn {
    V3 pos, color, normal, texcoord;
};

// nodes.py:107-111 + common.py:140-149; the +1 texel is clamped (reference reads one past
// the end with weight 0 there)
__device__ V3 tex_sample(const float *__restrict__ tex, int w, int h, int c, float u, float v) {
    float p0 = u * (float)(w - 1), p1 = v * (float)(h - 1);
    int I0 = f2i(floorf(p0)), I1 = f2i(floorf(p1));
    float x0 = p0 - (float)I0, x1 = p1 - (float)I1;
    float y0 = 1.0f - x0, y1 = 1.0f - x1;
    int i0 = min(max(I0, 0), w - 1), j0 = min(max(I1, 0), h - 1);
    int i1 = min(max(I0 + 1, 0), w - 1), j1 = min(max(I1 + 1, 0), h - 1);
    float o[3];
    if (c == 4 && (((uintptr_t)tex) & 15) == 0) { // RGBA texels (the host pads RGB images): one 128-bit load per texel
        const float4 *t4 = reinterpret_cast<const float4 *>(tex);
        const float4 f11 = __ldg(t4 + (long long)i1 * h + j1), f10 = __ldg(t4 + (long long)i1 * h + j0);
        const float4 f00 = __ldg(t4 + (long long)i0 * h + j0), f01 = __ldg(t4 + (long long)i0 * h + j1);
        o[0] = ((f11.x * x0 * x1 + f10.x * x0 * y1) + f00.x * y0 * y1) + f01.x * y0 * x1;
        o[1] = ((f11.y * x0 * x1 + f10.y * x0 * y1) + f00.y * y0 * y1) + f01.y * y0 * x1;
        o[2] = ((f11.z * x0 * x1 + f10.z * x0 * y1) + f00.z * y0 * y1) + f01.z * y0 * x1;
        return v3(o[0], o[1], o[2]);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int kk = c == 1 ? 0 : k;
        float f11 = __ldg(tex + ((long long)i1 * h + j1) * c + kk);
        float f10 = __ldg(tex + ((long long)i1 * h + j0) * c + kk);
        float f00 = __ldg(tex + ((long long)i0 * h + j0) * c + kk);
        float f01 = __ldg(tex + ((long long)i0 * h + j1) * c + kk);
        o[k] = ((f11 * x0 * x1 + f10 * x0 * y1) + f00 * y0 * y1) + f01 * y0 * x1;
    }
    return v3(o[0], o[1], o[2]);
}

// ---- material ops shared by the VM and the specialised paths (same op order => same bits) ----
template <bool FAST = false>
__device__ __forceinline__ V3 op_phong(V3 mm, V3 nrm, V3 idir, V3 odir) { // material.py:450-454, common.py:197-199
    V3 I3 = v3(-idir.x, -idir.y, -idir.z);
    if (FAST) {
        const float t = 2.0f * fdot3(nrm, I3);
        V3 rdir = v3(fmaf(-t, nrm.x, I3.x), fmaf(-t, nrm.y, I3.y), fmaf(-t, nrm.z, I3.z));
        const float VoR = fmaxf(0.0f, fdot3(odir, rdir));
        const float m = mm.x;
        if (mm.x == mm.y && mm.x == mm.z && m >= 1.0f && m <= 1024.0f && m == truncf(m)) {
            // integer shineness (uniform over the launch): square-and-multiply, <= 2*log2(m) roundings
            const int e = (int)m;
            float r = (e & 1) ? VoR : 1.0f, b = VoR;
#pragma unroll
            for (int k = 1; k <= 10; k++) {
                if ((e >> k) == 0) break;
                b *= b;
                if ((e >> k) & 1) r *= b;
            }
            r *= fmaf(m, 0.5f, 1.0f);
            return v3(r, r, r);
        }
    }
    float t = 2.0f * dot3(nrm, I3);
    V3 rdir = v3(I3.x - t * nrm.x, I3.y - t * nrm.y, I3.z - t * nrm.z);
    float VoR = fmaxf(0.0f, dot3(odir, rdir));
    if (mm.x == mm.y && mm.x == mm.z) { // scalar shineness (the usual case): one powf
        float r = powf(VoR, mm.x) * (mm.x + 2.0f) / 2.0f;
        return v3(r, r, r);
    }
    return v3(powf(VoR, mm.x) * (mm.x + 2.0f) / 2.0f, powf(VoR, mm.y) * (mm.y + 2.0f) / 2.0f,
              powf(VoR, mm.z) * (mm.z + 2.0f) / 2.0f);
}
__device__ __forceinline__ V3 op_cook(V3 ro, V3 f0, V3 nrm, V3 idir, V3 odir) { // material.py:323-362
    const float EPS = 1e-10f, eps = 1e-6f;
    V3 half = normalized(v3(idir.x + odir.x, idir.y + odir.y, idir.z + odir.z));
    float NoH = fmaxf(EPS, dot3(half, nrm));
    float NoL = fmaxf(EPS, dot3(idir, nrm));
    float NoV = fmaxf(EPS, dot3(odir, nrm));
    float VoH = fminf(1.0f, fmaxf(EPS, dot3(half, odir))); // 1 - 1e-10 == 1.0f
    float fr = powf(1.0f - VoH, 5.0f);
    float rr[3] = {ro.x, ro.y, ro.z}, ff[3] = {f0.x, f0.y, f0.z}, o[3];
    // a grey roughness (the usual case) gives the same D and G for the three channels: evaluate them once --
    // the same operations on the same values, hence the same bits as the per-channel loop of the reference
    const bool grey = ro.x == ro.y && ro.x == ro.z;
    float ndf = 0.0f, vdf = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (k == 0 || !grey) {
            float alpha2 = fmaxf(eps, rr[k] * rr[k]);
            float denom = 1.0f - (NoH * NoH) * (1.0f - alpha2);
            ndf = alpha2 / (denom * denom);
            float kk = alpha2 / 2.0f;
            vdf = 1.0f / ((NoV * kk + 1.0f) - kk);
            vdf *= 1.0f / ((NoL * kk + 1.0f) - kk);
            vdf /= 1.0f * (1.0f - alpha2) + 12.566370614359172f * alpha2; // common.py:221-223 lerp(alpha2, 1, 4 pi)
        }
        float fdf = ff[k] + (1.0f - ff[k]) * fr;
        o[k] = fdf * vdf * ndf;
    }
    return v3(o[0], o[1], o[2]);
}
__device__ __forceinline__ V3 op_mix(V3 f, V3 a, V3 b) { // material.py:96-118
    return v3((1.0f - f.x) * a.x + f.x * b.x, (1.0f - f.y) * a.y + f.y * b.y, (1.0f - f.z) * a.z + f.z * b.z);
}
__device__ __forceinline__ V3 op_mix_fast(V3 f, V3 a, V3 b) {
    return v3(fmaf(f.x, b.x, (1.0f - f.x) * a.x), fmaf(f.y, b.y, (1.0f - f.y) * a.y), fmaf(f.z, b.z, (1.0f - f.z) * a.z));
}

__device__ __forceinline__ V3 op_bcast(V3 v, int k) { // uv.x / uv.y of LerpTexture (nodes.py:129-136), broadcast
    const float c = k == 0 ? v.x : k == 1 ? v.y : v.z;
    return v3(c, c, c);
}
// ChessboardTexture's mix factor (nodes.py:114-126): (texcoord // size).sum() % 2 with Taichi's float semantics,
// a // b = floor(a / b) on the rounded quotient and a % b = a - b * (a // b)
__device__ __forceinline__ V3 op_chess(V3 uv, V3 size) {
    const float t = floorf(uv.x / size.x) + floorf(uv.y / size.y);
    const float m = t - 2.0f * floorf(t / 2.0f);
    return v3(m, m, m);
}

#define STK 12
__device__ V3 run_program(const TinaMaterial &m, int begin, int n, const ShadeIn &in, V3 nrm, V3 idir, V3 odir, V3 *regs) {
    V3 st[STK];
    int sp = 0;
    for (int pc = begin; pc < begin + n; pc++) {
        const TinaInstr &I = m.code[pc];
        switch (I.op) {
        case TINA_OP_REG:
            st[sp++] = regs[I.arg & (TINA_MAX_REGS - 1)];
            break;
        case TINA_OP_STORE:
            regs[I.arg & (TINA_MAX_REGS - 1)] = st[--sp];
            break;
        case TINA_OP_CONST:
            st[sp++] = v3(I.c[0], I.c[1], I.c[2]);
            break;
        case TINA_OP_INPUT:
            st[sp++] = I.arg == 0 ? in.pos : I.arg == 1 ? in.color : I.arg == 2 ? in.normal : in.texcoord;
            break;
        case TINA_OP_TEXTURE: {
            V3 uv = st[sp - 1];
            st[sp - 1] = tex_sample(m.tex[I.arg], m.tex_w[I.arg], m.tex_h[I.arg], m.tex_c[I.arg], uv.x, uv.y);
            break;
        }
        case TINA_OP_FRESNEL: { // material.py:69-83
            V3 sp_ = st[sp - 1], al = st[sp - 2], me = st[sp - 3];
            V3 r;
            r.x = me.x * al.x + (1.0f - me.x) * 0.16f * (sp_.x * sp_.x);
            r.y = me.y * al.y + (1.0f - me.y) * 0.16f * (sp_.y * sp_.y);
            r.z = me.z * al.z + (1.0f - me.z) * 0.16f * (sp_.z * sp_.z);
            sp -= 2;
            st[sp - 1] = r;
            break;
        }
        case TINA_OP_LAMBERT: { // material.py:392-393
            const float v = 0.3183098861837907f;
            st[sp++] = v3(v, v, v);
            break;
        }
        case TINA_OP_PHONG:
            st[sp - 1] = op_phong(st[sp - 1], nrm, idir, odir);
            break;
        case TINA_OP_COOK: {
            V3 f0 = st[sp - 1], ro = st[sp - 2];
            sp -= 1;
            st[sp - 1] = op_cook(ro, f0, nrm, idir, odir);
            break;
        }
        case TINA_OP_MIX: {
            V3 b = st[sp - 1], a = st[sp - 2], f = st[sp - 3];
            sp -= 2;
            st[sp - 1] = op_mix(f, a, b);
            break;
        }
        case TINA_OP_MUL: { // material.py:157-176
            V3 w = st[sp - 1], f = st[sp - 2];
            sp -= 1;
            st[sp - 1] = v3(f.x * w.x, f.y * w.y, f.z * w.z);
            break;
        }
        case TINA_OP_ADD: {
            V3 b = st[sp - 1], a = st[sp - 2];
            sp -= 1;
            st[sp - 1] = v3(a.x + b.x, a.y + b.y, a.z + b.z);
            break;
        }
        case TINA_OP_BCAST:
            st[sp - 1] = op_bcast(st[sp - 1], I.arg);
            break;
        case TINA_OP_CHESS:
            sp -= 1;
            st[sp - 1] = op_chess(st[sp - 1], st[sp]);
            break;
        default:
            break;
        }
    }
    return sp > 0 ? st[sp - 1] : v3(0.f, 0.f, 0.f);
}

// TinaMaterial.prologue_form == 2: the prologue in three-address form (include/tina_b200.h, TINA_OP3): one interpreter
// step per operation, sources are value registers, shading inputs or inline constants.  Same operations on the same
// values as run_program.
__device__ void run_prologue3(const TinaMaterial &m, int begin, int n, const ShadeIn &in, V3 *vals) {
    int pc = begin;
    const int end = begin + n;
    while (pc < end) {
        const TinaInstr &I = m.code[pc++];
        const unsigned a = (unsigned)I.arg;
        V3 s[3];
        const int op = I.op & 0xff;
        const int ns = op == TINA_OP_TEXTURE || op == TINA_OP_REG ? 1 : (op == TINA_OP_MUL || op == TINA_OP_ADD ? 2 : 3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (k < ns) {
                const unsigned code = (a >> (8 + 8 * k)) & 0xffu;
                if (code < TINA_VM_VALUES) s[k] = vals[code];
                else if (code < TINA_VM_VALUES + 4) s[k] = code == 16 ? in.pos : code == 17 ? in.color : code == 18 ? in.normal : in.texcoord;
                else {
                    const TinaInstr &C = m.code[pc++];
                    s[k] = v3(C.c[0], C.c[1], C.c[2]);
                }
            }
        }
        V3 r;
        switch (op) {
        case TINA_OP_TEXTURE: {
            const int t = (int)I.c[0];
            r = tex_sample(m.tex[t], m.tex_w[t], m.tex_h[t], m.tex_c[t], s[0].x, s[0].y);
            break;
        }
        case TINA_OP_FRESNEL: // material.py:69-83; sources in push order: metallic, albedo, specular
            r.x = s[0].x * s[1].x + (1.0f - s[0].x) * 0.16f * (s[2].x * s[2].x);
            r.y = s[0].y * s[1].y + (1.0f - s[0].y) * 0.16f * (s[2].y * s[2].y);
            r.z = s[0].z * s[1].z + (1.0f - s[0].z) * 0.16f * (s[2].z * s[2].z);
            break;
        case TINA_OP_MIX: // push order: factor, a, b
            r = op_mix(s[0], s[1], s[2]);
            break;
        case TINA_OP_MUL:
            r = v3(s[0].x * s[1].x, s[0].y * s[1].y, s[0].z * s[1].z);
            break;
        case TINA_OP_ADD:
            r = v3(s[0].x + s[1].x, s[0].y + s[1].y, s[0].z + s[1].z);
            break;
        default: // TINA_OP_REG: copy
            r = s[0];
            break;
        }
        vals[a & (TINA_VM_VALUES - 1)] = r;
    }
}

// TinaMaterial.prologue_form == 1: the prologue of tina.PBR with a textured base colour (matr/material.py:701-706 after
// the host's folding / hoisting), slot layout in material.py:_PBR_TEX_PROLOGUE.  The same operations in the same order
// as interpreting those 24 slots -- without 24 interpreter dispatches and their local-memory stack traffic.
__device__ __forceinline__ void prologue_pbr_textured(const TinaMaterial &m, int b, const ShadeIn &in, V3 *regs) {
    auto C = [&](int i) { return v3(m.code[b + i].c[0], m.code[b + i].c[1], m.code[b + i].c[2]); };
    const int t = m.code[b + 1].arg;
    const V3 base = tex_sample(m.tex[t], m.tex_w[t], m.tex_h[t], m.tex_c[t], in.texcoord.x, in.texcoord.y);
    const V3 me = C(3), sp = C(5);
    V3 fr; // material.py:69-83
    fr.x = me.x * base.x + (1.0f - me.x) * 0.16f * (sp.x * sp.x);
    fr.y = me.y * base.y + (1.0f - me.y) * 0.16f * (sp.y * sp.y);
    fr.z = me.z * base.z + (1.0f - me.z) * 0.16f * (sp.z * sp.z);
    const V3 k9 = C(9), k19 = C(19);
    regs[0] = base, regs[1] = fr;
    regs[2] = v3(base.x * k9.x, base.y * k9.y, base.z * k9.z);
    regs[3] = op_mix(fr, base, C(14));
    regs[4] = op_mix(fr, v3(base.x * k19.x, base.y * k19.y, base.z * k19.z), C(21));
}

// prologue_form 3 / 4: tina.Classic / tina.Diffuse with a textured colour (material.py:_CLASSIC_TEX_PROLOGUE,
// _DIFFUSE_TEX_PROLOGUE), same operations in the same order as interpreting the slots
__device__ __forceinline__ void prologue_classic_textured(const TinaMaterial &m, int b, const ShadeIn &in, V3 *regs, bool classic) {
    auto C = [&](int i) { return v3(m.code[b + i].c[0], m.code[b + i].c[1], m.code[b + i].c[2]); };
    const int t = m.code[b + 1].arg;
    const V3 base = tex_sample(m.tex[t], m.tex_w[t], m.tex_h[t], m.tex_c[t], in.texcoord.x, in.texcoord.y);
    const V3 k4 = C(4);
    regs[0] = base;
    regs[1] = v3(base.x * k4.x, base.y * k4.y, base.z * k4.z);
    if (classic) {
        const V3 k14 = C(14);
        regs[2] = op_mix(C(7), base, C(9));
        regs[3] = op_mix(C(12), v3(base.x * k14.x, base.y * k14.y, base.z * k14.z), C(16));
    } else {
        const V3 k8 = C(8);
        regs[2] = v3(base.x * k8.x, base.y * k8.y, base.z * k8.z);
    }
}

// operand i of a specialised brdf shape: a constant or a prologue register
__device__ __forceinline__ V3 operand(const TinaMaterial &m, int i, const V3 *regs) {
    if (m.code[i].op == TINA_OP_REG) return regs[m.code[i].arg & (TINA_MAX_REGS - 1)];
    return v3(m.code[i].c[0], m.code[i].c[1], m.code[i].c[2]);
}
// a program that the host folded down to one constant / one register needs no interpreter
__device__ __forceinline__ V3 run_or_const(const TinaMaterial &m, int begin, int n, const ShadeIn &in, V3 *regs) {
    if (n == 1 && (m.code[begin].op == TINA_OP_CONST || m.code[begin].op == TINA_OP_REG)) return operand(m, begin, regs);
    const V3 zero = v3(0.f, 0.f, 0.f);
    return run_program(m, begin, n, in, zero, zero, zero, regs);
}

__device__ __forceinline__ float aces(float c) { // advans.py:32-35
    return c * (2.51f * c + 0.03f) / (c * (2.43f * c + 0.59f) + 0.14f);
}
template <bool FAST>
__device__ __forceinline__ float aces_t(float c) {
    if (!FAST) return aces(c);
    return c * fmaf(2.51f, c, 0.03f) * rcp_fast(fmaf(c, fmaf(2.43f, c, 0.59f), 0.14f));
}

// the part of triangle.py:93-113 that render_color re-reads from the setup cache (:140-145):
// b, c, bcn, can, wscale.  Same ops as setup_face for these values => same bits.
__device__ __forceinline__ void setup_weights_clip(float4 ca, float4 cb, float4 cc, const Cam &cam, Setup &s) {
    const float ax = ca.x, ay = ca.y, aw = ca.w, bx = cb.x, by = cb.y, bw = cb.w, cx = cc.x, cy = cc.y, cw = cc.w;
    const float rx = cam.fW, ry = cam.fH;
    float pax = fm(fa(fm(ax, 0.5f), 0.5f), rx), pay = fm(fa(fm(ay, 0.5f), 0.5f), ry);
    float pbx = fm(fa(fm(bx, 0.5f), 0.5f), rx), pby = fm(fa(fm(by, 0.5f), 0.5f), ry);
    float pcx = fm(fa(fm(cx, 0.5f), 0.5f), rx), pcy = fm(fa(fm(cy, 0.5f), 0.5f), ry);
    float n = fs(fm(fs(pbx, pax), fs(pcy, pay)), fm(fs(pby, pay), fs(pcx, pax)));
    {
        const float a[4] = {fs(pbx, pcx), fs(pby, pcy), fs(pcx, pax), fs(pcy, pay)};
        float q[4];
        div_many(a, n, q);
        s.bcnx = q[0], s.bcny = q[1], s.canx = q[2], s.cany = q[3];
    }
    s.bx = pbx, s.by = pby, s.cx = pcx, s.cy = pcy;
    s.w0 = fd(1.0f, aw), s.w1 = fd(1.0f, bw), s.w2 = fd(1.0f, cw);
}
__device__ __forceinline__ void setup_weights(const float *v, const Cam &cam, Setup &s) {
    setup_weights_clip(vertex_clip(cam, v[0], v[1], v[2]), vertex_clip(cam, v[3], v[4], v[5]), vertex_clip(cam, v[6], v[7], v[8]),
                       cam, s);
}

// brdf program shapes the host's constant folding produces for the stock materials
#define MAT_GENERIC 0 /* interpret the program                                             */
#define MAT_CONST 1   /* [X]                      tina.Diffuse (X = CONST or a prologue REGister) */
#define MAT_CLASSIC 2 /* [X f, X a, X m, PHONG, MIX]                  tina.Classic        */
#define MAT_PBR 3     /* [X f, X a, X ro, X f0, COOK, MIX]            tina.PBR            */

// shade one covered pixel: triangle.py:139-153 + :32-49 + shader.py:119-131 + lighting.py:84-98
// triangle.py:139-153 + :32-49: gather face f, recompute the weights at pixel P, interpolate
// CF >= 0: the raster's SMOOTHING / TEXTURING bits as a compile-time constant (lean kernels), else runtime `flags_rt`
template <bool IDX, bool FAST = false, int CF = -1, int CK = 0>
__device__ __forceinline__ void pixel_inputs(int P, unsigned f, const float *__restrict__ verts, const float *__restrict__ norms,
                                             const float *__restrict__ coors, const Cam &cam, uint32_t flags_rt, const Src &S,
                                             ShadeIn &in, float &px, float &py) {
    const uint32_t flags = CF >= 0 ? (uint32_t)CF : flags_rt;
    const int x = P / cam.H, y = P - x * cam.H;
    float vv[9], n9[9], t6[6];
    Setup s;
    bool nsign = false;
    if (IDX) { // gather the face's corners through the mesh's own indexing (per-unique-vertex arrays)
        int iv[3], it[3], in_[3], gi[3], gj[3];
        bool neg;
        corner_ids<CK>(S, (long long)f, iv, it, in_, gi, gj, neg);
        // every gather is issued before the first use of any of them (one exposed round trip, not three);
        // the sign of a negated normal is applied after the interpolation (-(x) commutes with rounding)
        const float4 ca = __ldg(S.recA + iv[0]), cb = __ldg(S.recA + iv[1]), cc = __ldg(S.recA + iv[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float *p = S.vpos + (long long)iv[k] * 3;
            vv[k * 3] = __ldg(p), vv[k * 3 + 1] = __ldg(p + 1), vv[k * 3 + 2] = __ldg(p + 2);
        }
        if (flags & TINA_SMOOTHING) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float *p = S.vnrm + (long long)in_[k] * 3;
                n9[k * 3] = __ldg(p), n9[k * 3 + 1] = __ldg(p + 1), n9[k * 3 + 2] = __ldg(p + 2);
            }
            nsign = neg;
        }
        if (flags & TINA_TEXTURING) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (S.kind == 1) { // grid.py:17-21: I / (res - 1)
                    t6[k * 2] = (float)gi[k] / (float)(S.nx - 1), t6[k * 2 + 1] = (float)gj[k] / (float)(S.ny - 1);
                } else {
                    const float *p = S.vtex + (long long)it[k] * 2;
                    t6[k * 2] = __ldg(p), t6[k * 2 + 1] = __ldg(p + 1);
                }
            }
        }
        setup_from_records(ca, cb, cc, s); // the vertex stage's viewport coordinates and 1/w: same bits as setup_weights
    } else {
        const float *v = verts + (long long)f * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) vv[k] = __ldg(v + k);
        if (flags & TINA_SMOOTHING) {
            const float *nn = norms + (long long)f * 9;
#pragma unroll
            for (int k = 0; k < 9; k++) n9[k] = __ldg(nn + k);
        }
        if (flags & TINA_TEXTURING) {
            const float *tt = coors + (long long)f * 6;
#pragma unroll
            for (int k = 0; k < 6; k++) t6[k] = __ldg(tt + k);
        }
        setup_weights(vv, cam, s);
    }
    px = fa((float)x, cam.bias[0]), py = fa((float)y, cam.bias[1]);
    PW w = pix_products(s, px, py);
    float q0, q1, q2;
    pix_finish(s, w, q0, q1, q2);
    // triangle.py:32-49 interpolate
    if (FAST) {
        in.pos = v3(fmaf(q2, vv[6], fmaf(q1, vv[3], q0 * vv[0])), fmaf(q2, vv[7], fmaf(q1, vv[4], q0 * vv[1])),
                    fmaf(q2, vv[8], fmaf(q1, vv[5], q0 * vv[2])));
        if (flags & TINA_SMOOTHING)
            in.normal = v3(fmaf(q2, n9[6], fmaf(q1, n9[3], q0 * n9[0])), fmaf(q2, n9[7], fmaf(q1, n9[4], q0 * n9[1])),
                           fmaf(q2, n9[8], fmaf(q1, n9[5], q0 * n9[2])));
    } else {
        in.pos = v3((q0 * vv[0] + q1 * vv[3]) + q2 * vv[6], (q0 * vv[1] + q1 * vv[4]) + q2 * vv[7],
                    (q0 * vv[2] + q1 * vv[5]) + q2 * vv[8]);
        if (flags & TINA_SMOOTHING)
            in.normal = v3((q0 * n9[0] + q1 * n9[3]) + q2 * n9[6], (q0 * n9[1] + q1 * n9[4]) + q2 * n9[7],
                           (q0 * n9[2] + q1 * n9[5]) + q2 * n9[8]);
    }
    if (nsign) in.normal = v3(-in.normal.x, -in.normal.y, -in.normal.z);
    if (!(flags & TINA_SMOOTHING))
        in.normal = cross3(v3(vv[3] - vv[0], vv[4] - vv[1], vv[5] - vv[2]), v3(vv[6] - vv[0], vv[7] - vv[1], vv[8] - vv[2]));
    in.normal = normalized_t<FAST>(in.normal);
    in.texcoord = v3(0.f, 0.f, 0.f);
    if (flags & TINA_TEXTURING) {
        in.texcoord.x = (q0 * t6[0] + q1 * t6[2]) + q2 * t6[4];
        in.texcoord.y = (q0 * t6[1] + q1 * t6[3]) + q2 * t6[5];
    }
    in.color = v3(1.f, 1.f, 1.f);
}

// shader.py:82-93 calc_viewdir
template <bool FAST = false>
__device__ __forceinline__ V3 view_direction(const Cam &cam, float px, float py) {
    if (FAST) {
        // same ray without the six divisions: with h0 = V2W (qx,qy,-1,1), h1 = V2W (qx,qy,+1,1) the reference's
        // ro1 - ro = h1.xyz/h1.w - h0.xyz/h0.w is parallel to h1.xyz*h0.w - h0.xyz*h1.w (sign of h0.w*h1.w)
        const float *V = cam.V2W;
        const float qx = fmaf(px, cam.inv2W, -1.0f), qy = fmaf(py, cam.inv2H, -1.0f);
        const float b0 = fmaf(V[0], qx, fmaf(V[1], qy, V[3])), b1 = fmaf(V[4], qx, fmaf(V[5], qy, V[7]));
        const float b2 = fmaf(V[8], qx, fmaf(V[9], qy, V[11])), b3 = fmaf(V[12], qx, fmaf(V[13], qy, V[15]));
        const float w0 = b3 - V[14], w1 = b3 + V[14];
        V3 d = v3(fmaf(b0 + V[2], w0, -(b0 - V[2]) * w1), fmaf(b1 + V[6], w0, -(b1 - V[6]) * w1),
                  fmaf(b2 + V[10], w0, -(b2 - V[10]) * w1));
        float inv = rsq_fast(fdot3(d, d));
        if (w0 * w1 > 0.0f) inv = -inv; // returns -rd
        return v3(d.x * inv, d.y * inv, d.z * inv);
    }
    const float qx = px / cam.fW * 2.0f - 1.0f, qy = py / cam.fH * 2.0f - 1.0f;
    V3 ro = mapply_pos3(cam.V2W, qx, qy, -1.0f), ro1 = mapply_pos3(cam.V2W, qx, qy, 1.0f);
    V3 rd = normalized(v3(ro1.x - ro.x, ro1.y - ro.y, ro1.z - ro.z));
    return v3(-rd.x, -rd.y, -rd.z);
}

// lighting.py:84-98 (+ the per-pixel prologue registers of the material program)
// LEANOPS: the host verified that the material has no prologue and that every operand of the brdf shape, the
// ambient and the emission program is a constant (or absent): no register file, no interpreter, no operand tests.
__device__ __forceinline__ V3 const_operand(const TinaMaterial &m, int i) { return v3(m.code[i].c[0], m.code[i].c[1], m.code[i].c[2]); }
template <int KIND, bool FAST = false, bool LEANOPS = false>
__device__ __forceinline__ V3 light_pixel(const ShadeIn &in, V3 viewdir, const TinaMaterial &mat, const TinaLighting &L) {
    V3 res = v3(0.f, 0.f, 0.f);
    V3 regs[LEANOPS ? 1 : TINA_VM_VALUES];
    if (LEANOPS) {
        if (mat.n_emission) res = const_operand(mat, mat.n_brdf + mat.n_ambient);
        if (mat.n_ambient) {
            const V3 am = const_operand(mat, mat.n_brdf);
            res.x += L.ambient[0] * am.x, res.y += L.ambient[1] * am.y, res.z += L.ambient[2] * am.z;
        }
    } else {
        if (mat.n_prologue) { // light-independent sub-expressions (texture samples, Fresnel factors ...), once per pixel
            const V3 zero = v3(0.f, 0.f, 0.f);
            if (mat.prologue_form == 1) prologue_pbr_textured(mat, mat.n_brdf + mat.n_ambient + mat.n_emission, in, regs);
            else if (mat.prologue_form >= 3) prologue_classic_textured(mat, mat.n_brdf + mat.n_ambient + mat.n_emission, in, regs, mat.prologue_form == 3);
            else if (mat.prologue_form == 2) run_prologue3(mat, mat.n_brdf + mat.n_ambient + mat.n_emission, mat.n_prologue, in, regs);
            else run_program(mat, mat.n_brdf + mat.n_ambient + mat.n_emission, mat.n_prologue, in, zero, zero, zero, regs);
        }
        V3 em = run_or_const(mat, mat.n_brdf + mat.n_ambient, mat.n_emission, in, regs);
        res.x += em.x, res.y += em.y, res.z += em.z;
        V3 am = run_or_const(mat, mat.n_brdf, mat.n_ambient, in, regs);
        res.x += L.ambient[0] * am.x, res.y += L.ambient[1] * am.y, res.z += L.ambient[2] * am.z;
    }
    for (int l = 0; l < L.nlights; l++) {
        const float lw = L.dirs[l][3];
        V3 ld = v3(L.dirs[l][0] - in.pos.x * lw, L.dirs[l][1] - in.pos.y * lw, L.dirs[l][2] - in.pos.z * lw);
        float cos_i, d2;
        if (FAST) {
            d2 = fdot3(ld, ld);
            const float inv = rsq_fast(d2);
            ld = v3(ld.x * inv, ld.y * inv, ld.z * inv);
            cos_i = fdot3(in.normal, ld);
        } else {
            float dist = sqrtf(dot3(ld, ld));
            ld = v3(ld.x / dist, ld.y / dist, ld.z / dist);
            cos_i = dot3(in.normal, ld);
            d2 = dist * dist;
        }
        if (cos_i > 0.0f) {
            V3 mc;
            if (LEANOPS && KIND == MAT_CONST) {
                mc = const_operand(mat, 0);
            } else if (LEANOPS && KIND == MAT_CLASSIC) {
                V3 ph = op_phong<FAST>(const_operand(mat, 2), in.normal, ld, viewdir);
                mc = FAST ? op_mix_fast(const_operand(mat, 0), const_operand(mat, 1), ph)
                          : op_mix(const_operand(mat, 0), const_operand(mat, 1), ph);
            } else if (KIND == MAT_CONST) {
                mc = operand(mat, 0, regs);
            } else if (KIND == MAT_CLASSIC) {
                V3 ph = op_phong<FAST>(operand(mat, 2, regs), in.normal, ld, viewdir);
                mc = FAST ? op_mix_fast(operand(mat, 0, regs), operand(mat, 1, regs), ph)
                          : op_mix(operand(mat, 0, regs), operand(mat, 1, regs), ph);
            } else if (KIND == MAT_PBR) {
                V3 ck = op_cook(operand(mat, 2, regs), operand(mat, 3, regs), in.normal, ld, viewdir);
                mc = op_mix(operand(mat, 0, regs), operand(mat, 1, regs), ck);
            } else {
                mc = run_program(mat, 0, mat.n_brdf, in, in.normal, ld, viewdir, regs);
            }
            if (FAST) {
                const float k = cos_i * rcp_fast(d2);
                res.x = fmaf(k * L.colors[l][0], mc.x, res.x);
                res.y = fmaf(k * L.colors[l][1], mc.y, res.y);
                res.z = fmaf(k * L.colors[l][2], mc.z, res.z);
            } else {
                res.x += cos_i * (L.colors[l][0] / d2) * mc.x;
                res.y += cos_i * (L.colors[l][1] / d2) * mc.y;
                res.z += cos_i * (L.colors[l][2] / d2) * mc.z;
            }
        }
    }
    return res;
}

// shade one covered pixel: shader.py:119-131 + lighting.py:84-98
// LEAN: 0 generic; else a lean kernel for rasters without texturing (compile-time flags, constant operands):
// 1 / 2 = flat / smooth with the source kind read at run time; 3 / 4 = flat / smooth on a plain MeshGrid source,
// 5 / 6 = on a plain MeshModel source (indexed, mode 0: corner_ids with compile-time kind)
template <int KIND, bool IDX, bool FAST, int LEAN = 0>
__device__ __forceinline__ V3 shade_pixel(int P, unsigned f, const float *__restrict__ verts, const float *__restrict__ norms,
                                       const float *__restrict__ coors, const Cam &cam, uint32_t flags,
                                       const TinaMaterial &mat, const TinaLighting &L, const Src &S) {
    ShadeIn in;
    float px, py;
    pixel_inputs<IDX, FAST, LEAN == 0 ? -1 : ((LEAN & 1) ? 0 : (int)TINA_SMOOTHING), LEAN <= 2 ? 0 : (LEAN <= 4 ? 1 : 2)>(
        P, f, verts, norms, coors, cam, flags, S, in, px, py);
    return light_pixel<KIND, FAST, LEAN != 0>(in, view_direction<FAST>(cam, px, py), mat, L);
}

// K4: one CTA per 256-pixel chunk, one thread per pixel (x-major, so a warp covers 32 consecutive y).
// Measured alternatives on C2 (profiles/r1_k4_variants.md): 4 pixels per thread with serial shading 62 us,
// 4-pixel classification + shared-memory compaction + CTA-wide shading 37 us, persistent CTAs striding over
// chunks (flags read in one batch, next key prefetched) 25.0 us, persistent warps over 32-pixel units 25-27 us,
// this mapping 25 us (29-31 us before the relaxed shading arithmetic).
#ifndef K4_THREADS
#define K4_THREADS 256
#endif
#ifndef K4_MINBLOCKS
#define K4_MINBLOCKS 4
#endif
// key buffers of all ranks for the fused composite (n == 0: plain render_color on the local keys)
struct PeerTab {
    const long long *p[TINA_MAX_PEERS];
    int n, self;
};

// GLUE: the instantiations that can finish other objects' pixels (TINA_COLOR_FINISH) and accumulate (TAA); the
// plain ones compile those paths away (carrying them as run-time options cost the C2 frame 2 us).
// COMP: the sort-last composite over peer memory (keys = MIN over every rank's buffer, image strip stored into the
// root rank's image): its own instantiations, so that the plain kernels carry none of it.
// PERSIST: a smaller grid walks the chunks with a grid stride (~3.75 chunks per CTA) instead of one CTA per chunk: a
// quarter of the CTA launches (launching the 8100 CTAs of a 1080p pass costs 4 us by itself, tools/micro/cta_turnover.cu),
// and the last CTA starts -- and lets the next frame's vertex stage go -- earlier.  C2 sustained frame 49.9 -> 48.0 us.
template <int KIND, bool IDX, bool FAST, int LEAN = 0, bool GLUE = false, bool COMP = false, bool PERSIST = false>
__global__ void __launch_bounds__(K4_THREADS, (PERSIST && LEAN != 0) ? 5 : K4_MINBLOCKS) // (lean + PERSIST: 48 registers = five CTAs per SM like the plain form)
k_render_color(const long long *__restrict__ keys, const float *__restrict__ verts, const float *__restrict__ norms,
               const float *__restrict__ coors, const __grid_constant__ Cam cam, uint32_t flags, unsigned base,
               unsigned nfaces, const __grid_constant__ TinaMaterial mat, const __grid_constant__ TinaLighting L,
               float *__restrict__ image, uint32_t cflags, float bg0, float bg1, float bg2,
               const __grid_constant__ Src S, const unsigned char *__restrict__ blkflags, unsigned *__restrict__ publish,
               const unsigned *__restrict__ counters, int pix_lo, int pix_hi, unsigned *__restrict__ pubstate,
               unsigned char flagval, const __grid_constant__ PeerTab peers, long long *__restrict__ keys_out,
               float *__restrict__ acc_rt, int acc_count) {
    static_assert((1 << FLAG_SHIFT) % K4_THREADS == 0 && (K4_THREADS * 3) % 4 == 0, "a K4 block lies inside one coverage chunk");
    float *const acc = GLUE ? acc_rt : nullptr;
    pdl_wait();
#ifndef K4_NO_TRIGGER
    // the rasteriser is complete: the next frame's vertex stage may start on the SMs this kernel's tail leaves idle (its
    // vertex blocks write the other record set and do not wait; its clear blocks wait for this grid to finish)
    pdl_launch_dependents();
#endif
    if (blockIdx.x == 0 && threadIdx.x == 0 && publish) { // tell the host how many faces needed the tile path
        // running counts live in device memory; the mapped host words are only written (posted stores, no PCIe
        // round trip).  The host reads them as a heuristic, a stale value is harmless.
        const unsigned nq = counters[0];
        const unsigned npub = pubstate[0] + 1u;              // publishes so far
        const unsigned streak = nq ? 0u : pubstate[1] + 1u;  // consecutive render_occup/render_color pairs without large faces
        pubstate[0] = npub, pubstate[1] = streak;
        publish[0] = nq, publish[1] = npub, publish[2] = streak;
    }
    const int npix = pix_hi; // this launch shades pixels [pix_lo, pix_hi); pix_lo is a multiple of 256
    const bool fill = (cflags & TINA_COLOR_FILL_BG) != 0;
    // TINA_COLOR_FINISH: the frame's last shading pass also finishes the pixels it does not own (tonemap + TAA
    // accumulation, scene/raster.py:202-207) instead of leaving them to separate full-screen passes
    const bool finish = GLUE && (cflags & TINA_COLOR_FINISH) != 0 && !fill;
    float r = bg0, g = bg1, b = bg2;
    if (fill && (cflags & TINA_COLOR_TONEMAP)) r = aces(r), g = aces(g), b = aces(b);
    // util/accumator.py:16-23 for one pixel (the same f32 operations as k_accumulate)
    auto accumulate = [&](long long P_, float cr, float cg, float cb) {
        const float inv = __fdiv_rn(1.0f, (float)acc_count), keep = __fsub_rn(1.0f, inv);
        float *a = acc + P_ * 3;
        a[0] = __fadd_rn(__fmul_rn(a[0], keep), __fmul_rn(cr, inv));
        a[1] = __fadd_rn(__fmul_rn(a[1], keep), __fmul_rn(cg, inv));
        a[2] = __fadd_rn(__fmul_rn(a[2], keep), __fmul_rn(cb, inv));
    };
    // a pixel of another object (or the background) in the frame's last pass
    auto finish_pixel = [&](long long P_) {
        float *o = image + P_ * 3;
        float cr = o[0], cg = o[1], cb = o[2];
        if (cflags & TINA_COLOR_TONEMAP) {
            cr = aces(cr), cg = aces(cg), cb = aces(cb);
            o[0] = cr, o[1] = cg, o[2] = cb;
        }
        if (acc) accumulate(P_, cr, cg, cb);
    };
    // (CTAs take the chunks in plain order: spreading covered and background chunks over the launch -- CTA b -> chunk
    // (b % 8) * n / 8 + b / 8 -- measured 1.3 us slower on C2, profiles/r2_k4_variants.md)
    static_assert(!(PERSIST && (COMP || GLUE)), "the persistent form serves the plain passes only");
    auto shade_chunk = [&](const long long p0) {
    // flagval != 0: this object's render_occup was the engine's last, its flags carry its own stamp -> a chunk with
    // any other value holds none of its pixels.  flagval == 0: only "nothing rasterised here since the clear" is known.
    const unsigned char cf = blkflags ? blkflags[p0 >> FLAG_SHIFT] : (unsigned char)1;
    if (blkflags && (flagval ? cf != flagval : cf == 0)) {
        if (finish) {
            if (p0 + threadIdx.x < npix) finish_pixel(p0 + threadIdx.x);
        } else if (fill) {
            const int np = (int)min((long long)K4_THREADS, (long long)npix - p0);
            const int t = threadIdx.x;
            if (acc) { // background + accumulation: per-pixel read-modify-write of the accumulator
                if (t < np) {
                    float *out = image + (p0 + t) * 3;
                    out[0] = r, out[1] = g, out[2] = b;
                    accumulate(p0 + t, r, g, b);
                }
            } else if (np == K4_THREADS && (((uintptr_t)image) & 15) == 0) { // K4_THREADS * 12 contiguous, 16-byte aligned bytes
                if (t < K4_THREADS * 3 / 4) {
                    const int m = t % 3;
                    const float4 v = m == 0 ? make_float4(r, g, b, r) : m == 1 ? make_float4(g, b, r, g) : make_float4(b, r, g, b);
                    __stcs(reinterpret_cast<float4 *>(image + p0 * 3) + t, v);
                }
            } else if (t < np) {
                float *out = image + (p0 + t) * 3;
                out[0] = r, out[1] = g, out[2] = b;
            }
        }
        return;
    }
    const long long Pl = p0 + threadIdx.x;
    if (COMP) {
        // Sort-last composite fused into the shading pass: the winner of a pixel is the MIN of the packed keys of every
        // rank, read straight from the peers' key buffers over NVLink (L2-coherent loads: a peer's buffer changes between
        // frames, L1 must not keep it); every peer's key is requested before the first one is used (one round trip per
        // pixel, not one per peer); the composited key is kept in the local buffer.
        // The image usually lives on another GPU too (SharedImage on the root rank).  Three 4-byte stores per pixel
        // cross NVLink as partially filled sectors -- 0.97 ms for a 57 MB strip on 8 GPUs, twice what the root's port needs
        // for all seven strips (tools/c5_breakdown.py) -- so the CTA stages its 3 KB in shared memory and sends whole
        // 16-byte pieces, as long as every pixel of the CTA has a colour (a face, or the background fill).
        __shared__ __align__(16) float simg[K4_THREADS * 3];
        const bool valid = Pl < npix;
        const int P = valid ? (int)Pl : 0;
        bool have = false;
        float cr = r, cg = g, cb = b;
        if (valid) {
            long long kk[TINA_MAX_PEERS];
#pragma unroll
            for (int q = 0; q < TINA_MAX_PEERS; q++) kk[q] = q < peers.n ? __ldcg(peers.p[q] + P) : 0x7fffffffffffffffll;
            long long k = kk[0];
#pragma unroll
            for (int q = 1; q < TINA_MAX_PEERS; q++) k = kk[q] < k ? kk[q] : k;
            keys_out[P] = k;
            const unsigned idc = (unsigned)(unsigned long long)k, fidc = idc - 1u - base;
            if (idc == 0u || fidc >= nfaces) {
                have = fill;
            } else {
                V3 c = shade_pixel<KIND, IDX, FAST, LEAN>(P, fidc, verts, norms, coors, cam, flags, mat, L, S);
                if (cflags & TINA_COLOR_TONEMAP) c.x = aces_t<FAST>(c.x), c.y = aces_t<FAST>(c.y), c.z = aces_t<FAST>(c.z);
                cr = c.x, cg = c.y, cb = c.z, have = true;
            }
        }
        simg[threadIdx.x * 3] = cr, simg[threadIdx.x * 3 + 1] = cg, simg[threadIdx.x * 3 + 2] = cb;
        const bool whole = __syncthreads_and(have && valid) && (((uintptr_t)image) & 15) == 0; // (p0 * 12 is a multiple of 16)
        if (whole) {
            if (threadIdx.x < K4_THREADS * 3 / 4)
                __stcs(reinterpret_cast<float4 *>(image + p0 * 3) + threadIdx.x, reinterpret_cast<const float4 *>(simg)[threadIdx.x]);
        } else if (have && valid) {
            float *o = image + (long long)P * 3;
            __stcs(o, cr), __stcs(o + 1, cg), __stcs(o + 2, cb);
        }
        return;
    }
    if (Pl >= npix) return;
    const int P = (int)Pl;
    const unsigned id = (unsigned)(unsigned long long)__ldcs(keys + P);
    const unsigned fid = id - 1u - base;
    float *out = image + (long long)P * 3;
    if (id == 0u || fid >= nfaces) { // triangle.py:137-138 (occup == -1)
        if (finish) {
            finish_pixel(P);
        } else if (fill) {
            __stcs(out, r), __stcs(out + 1, g), __stcs(out + 2, b);
            if (acc) accumulate(P, r, g, b);
        }
        return;
    }
    V3 c = shade_pixel<KIND, IDX, FAST, LEAN>(P, fid, verts, norms, coors, cam, flags, mat, L, S);
    if (cflags & TINA_COLOR_TONEMAP) c.x = aces_t<FAST>(c.x), c.y = aces_t<FAST>(c.y), c.z = aces_t<FAST>(c.z);
    __stcs(out, c.x), __stcs(out + 1, c.y), __stcs(out + 2, c.z);
    if (acc) accumulate(P, c.x, c.y, c.z);
    };
    if (!PERSIST) {
        shade_chunk((long long)pix_lo + (long long)blockIdx.x * K4_THREADS);
    } else {
        const unsigned nch = ((unsigned)(pix_hi - pix_lo) + K4_THREADS - 1u) / K4_THREADS;
        for (unsigned c = blockIdx.x; c < nch; c += gridDim.x) shade_chunk((long long)pix_lo + (long long)c * K4_THREADS);
    }
}

// G-buffer sinks (core/shader.py:21-109, probe.py:21-23): attributes of the visible surface per pixel.  One launch
// serves every sink of a ShaderGroup (scene/raster.py:101-107): the face is gathered and the weights recomputed once.
#define TINA_MAX_SINKS 8
struct SinkTab {
    int n;
    int kind[TINA_MAX_SINKS], ncomp[TINA_MAX_SINKS], is_int[TINA_MAX_SINKS];
    void *out[TINA_MAX_SINKS];
    float p[TINA_MAX_SINKS][3];
};
// the sinks of one visible pixel: shader.py:21-109, probe.py:21-23 (shared by the triangle and the particle rasteriser)
__device__ __forceinline__ void write_sinks(const SinkTab &T, int P, long long key, unsigned f, const ShadeIn &in, float px, float py, V3 vd,
                                            const Cam &cam) {
    for (int k = 0; k < T.n; k++) {
        const int kind = T.kind[k], ncomp = T.ncomp[k];
        const float p0 = T.p[k][0], p1 = T.p[k][1], p2 = T.p[k][2];
        float v[3] = {0.f, 0.f, 0.f};
        if (kind == TINA_SINK_ELMID) { // probe.py:22: the face id itself (exact as an integer)
            if (T.is_int[k]) {
                int *o = reinterpret_cast<int *>(T.out[k]) + (long long)P * ncomp;
                for (int c = 0; c < ncomp; c++) o[c] = (int)f;
                continue;
            }
            v[0] = v[1] = v[2] = (float)f;
        } else if (kind == TINA_SINK_CONST) {
            v[0] = p0, v[1] = p1, v[2] = p2;
        } else if (kind == TINA_SINK_DEPTH) {
            v[0] = v[1] = v[2] = (float)(int)(key >> 32); // shader.py:39-42: engine.depth[P]
        } else if (kind == TINA_SINK_COLOR) {
            v[0] = in.color.x, v[1] = in.color.y, v[2] = in.color.z; // triangle.py:48: (1, 1, 1); particle.py:159: the particle's colour
        } else if (kind == TINA_SINK_POSITION) {
            v[0] = in.pos.x, v[1] = in.pos.y, v[2] = in.pos.z;
        } else if (kind == TINA_SINK_NORMAL) {
            v[0] = in.normal.x, v[1] = in.normal.y, v[2] = in.normal.z;
        } else if (kind == TINA_SINK_VIEWNORMAL) { // shader.py:51-58: mapply_dir(W2V, normal).normalized()
            float r0, r1, r2, rw;
            mapply(cam.W2V, in.normal.x, in.normal.y, in.normal.z, 0.0f, r0, r1, r2, rw);
            V3 n = normalized(v3(r0, r1, r2));
            v[0] = n.x, v[1] = n.y, v[2] = n.z;
        } else if (kind == TINA_SINK_TEXCOORD) {
            v[0] = in.texcoord.x, v[1] = in.texcoord.y;
        } else if (kind == TINA_SINK_CHESSBOARD) { // shader.py:73-79: lerp((p // size).sum() % 2, 0.4, 0.9)
            const float fac = fmodf(floorf(px / p0) + floorf(py / p0), 2.0f);
            const float m = fac < 0.0f ? fac + 2.0f : fac; // python-style modulo
            v[0] = v[1] = v[2] = 0.4f * (1.0f - m) + 0.9f * m;
        } else if (kind == TINA_SINK_VIEWDIR) { // shader.py:96-101
            v[0] = vd.x * 0.5f + 0.5f, v[1] = vd.y * 0.5f + 0.5f, v[2] = vd.z * 0.5f + 0.5f;
        } else { // TINA_SINK_SIMPLE, shader.py:104-109
            v[0] = v[1] = v[2] = fabsf(dot3(in.normal, vd));
        }
        if (T.is_int[k]) {
            int *o = reinterpret_cast<int *>(T.out[k]) + (long long)P * ncomp;
            for (int c = 0; c < ncomp; c++) o[c] = (int)v[c];
        } else {
            float *o = reinterpret_cast<float *>(T.out[k]) + (long long)P * ncomp;
            for (int c = 0; c < ncomp; c++) o[c] = v[c];
        }
    }
}

template <bool IDX>
__global__ void __launch_bounds__(256)
k_gbuffer(const long long *__restrict__ keys, const float *__restrict__ verts, const float *__restrict__ norms,
          const float *__restrict__ coors, const __grid_constant__ Cam cam, uint32_t flags, unsigned base, unsigned nfaces,
          const __grid_constant__ SinkTab T, const __grid_constant__ Src S) {
    pdl_wait();
    const int P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= cam.W * cam.H) return;
    const long long key = keys[P];
    const unsigned id = (unsigned)(unsigned long long)key;
    const unsigned f = id - 1u - base;
    if (id == 0u || f >= nfaces) return; // triangle.py:137-138: sinks are only written where this object is visible
    bool need_inputs = false, need_view = false;
    for (int k = 0; k < T.n; k++) {
        const int kd = T.kind[k];
        need_inputs |= kd == TINA_SINK_POSITION || kd == TINA_SINK_NORMAL || kd == TINA_SINK_VIEWNORMAL || kd == TINA_SINK_TEXCOORD ||
                       kd == TINA_SINK_CHESSBOARD || kd == TINA_SINK_VIEWDIR || kd == TINA_SINK_SIMPLE;
        need_view |= kd == TINA_SINK_VIEWDIR || kd == TINA_SINK_SIMPLE;
    }
    ShadeIn in;
    in.color = v3(1.0f, 1.0f, 1.0f); // triangle.py:48
    float px = 0.f, py = 0.f;
    V3 vd = v3(0.f, 0.f, 0.f);
    if (need_inputs) pixel_inputs<IDX>(P, f, verts, norms, coors, cam, flags, S, in, px, py);
    if (need_view) vd = view_direction(cam, px, py);
    write_sinks(T, P, key, f, in, px, py, vd, cam);
}
