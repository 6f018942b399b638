// ssr.cuh -- part of the single translation unit tina_b200.cu (included once, in order): screen-space reflections
// (postp/ssr.py), SSAO with per-frame samples (postp/ssao.py, taa=True) and the random streams they draw from
// (tina/random.py).  Plain IEEE f32 in the reference's operation order (the TU is compiled -fmad=false); what differs from
// the CPU restatement is the last ulp of sinf / cosf / powf / sqrtf, and a ray-march test that flips on it moves one sample
// of one pixel.
#pragma once

// random.py:26-35 WangHashRNG._noise
__device__ __forceinline__ unsigned wang_hash(unsigned v) {
    v = (v ^ 61u) ^ (v >> 16);
    v *= 9u;
    v ^= v << 4;
    v *= 0x27d4eb2du;
    v ^= v >> 15;
    return v;
}
struct WangRNG {
    unsigned seed;
    // random.py:54-58: (noise_int(seed) >> 1) * (2 / 4294967296): u32 -> f32 (rounded to nearest), times 2^-31; seed += 1
    __device__ __forceinline__ float random() {
        const unsigned u = wang_hash(seed) >> 1;
        seed += 1u;
        return __uint2float_rn(u) * 4.656612873077393e-10f;
    }
    __device__ __forceinline__ unsigned random_int() { // random.py:60-64
        const unsigned u = wang_hash(seed);
        seed += 1u;
        return u;
    }
};

// a parameter program of a TinaSampleMaterial (include/tina_b200.h): value ops only (CONST / INPUT / TEXTURE / FRESNEL and
// the MIX / ADD / BCAST / CHESS of the procedural textures)
#define SSR_STK 8
__device__ V3 run_value(const TinaSampleMaterial &m, int begin, int n, const ShadeIn &in) {
    V3 st[SSR_STK];
    int sp = 0;
    for (int pc = begin; pc < begin + n; pc++) {
        const TinaInstr &I = m.code[pc];
        switch (I.op) {
        case TINA_OP_CONST:
            st[sp++] = v3(I.c[0], I.c[1], I.c[2]);
            break;
        case TINA_OP_INPUT:
            st[sp++] = I.arg == 0 ? in.pos : I.arg == 1 ? in.color : I.arg == 2 ? in.normal : in.texcoord;
            break;
        case TINA_OP_TEXTURE: {
            const V3 uv = st[sp - 1];
            st[sp - 1] = tex_sample(m.tex[I.arg], m.tex_w[I.arg], m.tex_h[I.arg], m.tex_c[I.arg], uv.x, uv.y);
            break;
        }
        case TINA_OP_FRESNEL: { // material.py:69-83
            const V3 s_ = st[sp - 1], al = st[sp - 2], me = st[sp - 3];
            V3 r;
            r.x = me.x * al.x + (1.0f - me.x) * 0.16f * (s_.x * s_.x);
            r.y = me.y * al.y + (1.0f - me.y) * 0.16f * (s_.y * s_.y);
            r.z = me.z * al.z + (1.0f - me.z) * 0.16f * (s_.z * s_.z);
            sp -= 2;
            st[sp - 1] = r;
            break;
        }
        case TINA_OP_MIX: { // (lerp of the procedural textures, nodes.py:114-136)
            const V3 b = st[sp - 1], a = st[sp - 2], f = st[sp - 3];
            sp -= 2;
            st[sp - 1] = op_mix(f, a, b);
            break;
        }
        case TINA_OP_ADD: {
            const V3 b = st[sp - 1], a = st[sp - 2];
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

// advans.py:97-100 tangentspace(n) @ (advans.py:105-108 spherical(h, p) * scale)
__device__ __forceinline__ V3 tangent_spherical(V3 n, float h, float p, float scale, bool scaled) {
    const V3 bitan = normalized(cross3(n, v3(233.0f, 666.0f, 512.0f)));
    const V3 tan = cross3(bitan, n);
    const float ang = p * 6.283185307179586f;
    const float s = sqrtf(fmaxf(0.0f, 1.0f - h * h));
    float ux = s * cosf(ang), uy = s * sinf(ang), uz = h;
    if (scaled) ux = ux * scale, uy = uy * scale, uz = uz * scale;
    return v3((tan.x * ux + bitan.x * uy) + n.x * uz, (tan.y * ux + bitan.y * uy) + n.y * uz, (tan.z * ux + bitan.z * uy) + n.z * uz);
}
__device__ __forceinline__ V3 reflect_neg(V3 idir, V3 nrm) { // common.py:198-199 reflect(-idir, nrm)
    const V3 I3 = v3(-idir.x, -idir.y, -idir.z);
    const float t = 2.0f * dot3(nrm, I3);
    return v3(I3.x - t * nrm.x, I3.y - t * nrm.y, I3.z - t * nrm.z);
}

// material.sample(idir, nrm, 1, rng) (matr/material.py): descend from the root choosing a branch at Mix / Add nodes,
// sample the leaf, then apply the weights of the nodes passed on the way back up, innermost first -- the order in
// which the reference's nested calls multiply.  (rough, the third return value, is unused by SSR.)
#define SSR_MAX_DEPTH 8
__device__ void sample_material(const TinaSampleMaterial &m, const ShadeIn &in, V3 idir, V3 nrm, WangRNG &rng, V3 &odir, V3 &wei) {
    V3 pend[SSR_MAX_DEPTH];
    int np = 0, node = 0;
    for (int guard = 0; guard < TINA_SAMPLE_MAX_NODES; guard++) {
        const TinaSampleNode &N = m.nodes[node];
        if (N.kind < TINA_SNODE_MIX) break;
        if (N.kind == TINA_SNODE_MIX) { // material.py:123-138
            const V3 fac = run_value(m, N.p0, N.n0, in);
            // Vavg (common.py:36-40): a vector factor is averaged, a scalar one taken as it is (pad_ bit 0)
            float factor = (N.pad_ & 1) ? fac.x : ((fac.x + fac.y) + fac.z) / 3.0f;
            if (factor != 0.0f && factor != 1.0f) { // common.py:227-231 smoothlerp(., 0.12, 0.88)
                const float t = clamp01((factor - 0.0f) / (1.0f - 0.0f));
                factor = t * t * (3.0f - 2.0f * t);
                factor = 0.12f * (1.0f - factor) + 0.88f * factor;
            }
            if (rng.random() < factor) {
                node = N.b;
                if (np < SSR_MAX_DEPTH) pend[np++] = v3(fac.x / factor, fac.y / factor, fac.z / factor);
            } else {
                node = N.a;
                const float d = 1.0f - factor;
                if (np < SSR_MAX_DEPTH) pend[np++] = v3((1.0f - fac.x) / d, (1.0f - fac.y) / d, (1.0f - fac.z) / d);
            }
        } else if (N.kind == TINA_SNODE_SCALE) { // :180-184
            if (np < SSR_MAX_DEPTH) pend[np++] = run_value(m, N.p0, N.n0, in);
            node = N.a;
        } else { // TINA_SNODE_ADD, :227-238
            node = (rng.random_int() % 2u == 0u) ? N.a : N.b;
            if (np < SSR_MAX_DEPTH) pend[np++] = v3(2.0f, 2.0f, 2.0f);
        }
    }
    const TinaSampleNode &N = m.nodes[node];
    switch (N.kind) {
    case TINA_SNODE_LAMBERT: { // :398-405
        float u = rng.random();
        const float v = rng.random();
        u = sqrtf(u);
        odir = normalized(tangent_spherical(nrm, u, v, 1.0f, false));
        wei = v3(1.0f, 1.0f, 1.0f);
        break;
    }
    case TINA_SNODE_PHONG: { // :459-472 (the shineness is a scalar there)
        const float mm = run_value(m, N.p0, N.n0, in).x;
        float u = rng.random();
        const float v = rng.random();
        u = powf(u, 1.0f / (mm + 1.0f));
        odir = tangent_spherical(reflect_neg(idir, nrm), u, v, 1.0f, false);
        float w = 1.0f;
        if (dot3(odir, nrm) < 0.0f) odir = v3(-odir.x, -odir.y, -odir.z), w = 0.0f;
        wei = v3(w, w, w);
        break;
    }
    case TINA_SNODE_COOK: { // :364-384 sample, :323-357 sub_brdf
        const V3 ro = run_value(m, N.p0, N.n0, in), f0 = run_value(m, N.p1, N.n1, in);
        const float EPS = 1e-10f, eps = 1e-6f;
        const float alpha2s = fmaxf(0.0f, ro.x * ro.x);
        float u = rng.random();
        const float v = rng.random();
        u = sqrtf((1.0f - u) / (1.0f - u * (1.0f - alpha2s)));
        odir = tangent_spherical(reflect_neg(idir, nrm), u, v, 1.0f, false);
        const V3 half = normalized(v3(idir.x + odir.x, idir.y + odir.y, idir.z + odir.z));
        const float NoL = fmaxf(EPS, dot3(idir, nrm)), NoV = fmaxf(EPS, dot3(odir, nrm));
        const float VoH = fminf(1.0f, fmaxf(EPS, dot3(half, odir))); // 1 - 1e-10 == 1.0f
        const bool flip = dot3(odir, nrm) < 0.0f;
        const float fr = powf(1.0f - VoH, 5.0f);
        const float rr[3] = {ro.x, ro.y, ro.z}, ff[3] = {f0.x, f0.y, f0.z};
        float o[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float alpha2 = fmaxf(eps, rr[k] * rr[k]);
            const float kk = alpha2 / 2.0f;
            float vdf = 1.0f / ((NoV * kk + 1.0f) - kk);
            vdf *= 1.0f / ((NoL * kk + 1.0f) - kk);
            vdf /= 1.0f * (1.0f - alpha2) + 12.566370614359172f * alpha2;
            float fdf = ff[k] + (1.0f - ff[k]) * fr;
            if (flip) fdf = 0.0f;
            o[k] = fdf * vdf;
        }
        if (flip) odir = v3(-odir.x, -odir.y, -odir.z);
        wei = v3(o[0], o[1], o[2]);
        break;
    }
    default: // TINA_SNODE_EMISSION, :679-681
        odir = idir;
        wei = v3(0.0f, 0.0f, 0.0f);
        break;
    }
    for (int k = np - 1; k >= 0; k--) wei = v3(wei.x * pend[k].x, wei.y * pend[k].y, wei.z * pend[k].z);
}

struct SsrArgs {
    int nsamples, nsteps, blurring, taa, nmaterials;
    float stepsize, tolerance;
    unsigned frame;
};

// postp/ssr.py:44-103 render / render_at: one thread per pixel
__global__ void __launch_bounds__(128)
k_ssr_render(const long long *__restrict__ keys, const float *__restrict__ normals, const float *__restrict__ coors,
             const int *__restrict__ mtlid, const TinaSampleMaterial *__restrict__ table, const float *__restrict__ image,
             const __grid_constant__ Cam cam, const __grid_constant__ SsrArgs A, float4 *__restrict__ out4) {
    const int P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= cam.W * cam.H) return;
    const int i = P / cam.H, j = P - i * cam.H;
    const V3 normal = v3(normals[(long long)P * 3], normals[(long long)P * 3 + 1], normals[(long long)P * 3 + 2]);
    float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mid = mtlid[P];
    if (!((normal.x * normal.x + normal.y * normal.y) + normal.z * normal.z < 1e-6f) && mid >= 0 && mid < A.nmaterials) { // :47-48
        const TinaSampleMaterial &m = table[mid];
        const float px = (float)i + cam.bias[0], py = (float)j + cam.bias[1];
        const float vx = px / cam.fW * 2.0f - 1.0f, vy = py / cam.fH * 2.0f - 1.0f;
        const float vz = (float)(int)(keys[P] >> 32) / 1073741824.0f;
        const V3 pos = mapply_pos3(cam.V2W, vx, vy, vz);
        const V3 viewdir = view_direction<false>(cam, px, py);
        ShadeIn in;
        in.pos = pos, in.color = v3(1.0f, 1.0f, 1.0f), in.normal = normal;
        in.texcoord = coors ? v3(coors[(long long)P * 2], coors[(long long)P * 2 + 1], 0.0f) : v3(0.0f, 0.0f, 0.0f);
        WangRNG rng;
        if (!A.taa) rng.seed = wang_hash((unsigned)(j % A.blurring) ^ wang_hash((unsigned)(i % A.blurring))); // random.py:44-52 on P % blurring
        else rng.seed = wang_hash(A.frame ^ wang_hash((unsigned)j ^ wang_hash((unsigned)i)));
        const float fn = (float)A.nsteps;
        // (:86-88: the tolerance term does not depend on the sample)
        const float vtol = A.tolerance * (mapply_pos3(cam.W2V, pos.x - viewdir.x / fn, pos.y - viewdir.y / fn, pos.z - viewdir.z / fn).z -
                                          mapply_pos3(cam.W2V, pos.x, pos.y, pos.z).z);
        for (int s = 0; s < A.nsamples; s++) {
            V3 odir, wei;
            sample_material(m, in, viewdir, normal, rng, odir, wei);
            const float ov = dot3(odir, viewdir);
            const float step = A.stepsize / (sqrtf(1.0f - ov * ov) * fn);
            const float rr = rng.random();
            V3 ro = v3(pos.x + odir.x * rr * step, pos.y + odir.y * rr * step, pos.z + odir.z * rr * step);
            for (int t = 0; t < A.nsteps; t++) {
                ro = v3(ro.x + odir.x * step, ro.y + odir.y * step, ro.z + odir.z * step);
                const V3 vro = mapply_pos3(cam.W2V, ro.x, ro.y, ro.z);
                if (!(-1.0f <= vro.x && vro.x <= 1.0f && -1.0f <= vro.y && vro.y <= 1.0f && -1.0f <= vro.z && vro.z <= 1.0f)) break;
                const float Dx = (vro.x * 0.5f + 0.5f) * cam.fW, Dy = (vro.y * 0.5f + 0.5f) * cam.fH;
                const int ix = f2i(Dx), iy = f2i(Dy);
                const float d = (ix < 0 || iy < 0 || ix >= cam.W || iy >= cam.H)
                                    ? 0.0f
                                    : (float)(int)(keys[(long long)ix * cam.H + iy] >> 32) / 1073741824.0f;
                if (vro.z - vtol < d && d < vro.z) {
                    const V3 clr = bilerp3(image, cam.W, cam.H, Dx, Dy);
                    res.x += clr.x * wei.x, res.y += clr.y * wei.y, res.z += clr.z * wei.z, res.w += 1.0f;
                    break;
                }
            }
        }
        const float ns = (float)A.nsamples;
        res = make_float4(res.x / ns, res.y / ns, res.z / ns, res.w / ns);
    }
    out4[P] = res;
}

// postp/ssr.py:30-42 apply: res = img4 (taa) or its blurring x blurring box mean (reads outside the field are 0);
// image = image * (1 - res.w) + res.xyz
__global__ void k_ssr_apply(float *__restrict__ image, const float4 *__restrict__ img4, int W, int H, int blurring, int taa) {
    const int P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= W * H) return;
    const int i = P / H, j = P - i * H, offs = blurring / 2;
    float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
    if (taa) {
        res = img4[P];
    } else {
        for (int k = 0; k < blurring; k++)
            for (int l = 0; l < blurring; l++) {
                const int x = i + k - offs, y = j + l - offs;
                if (x < 0 || y < 0 || x >= W || y >= H) continue;
                const float4 v = __ldg(img4 + (long long)x * H + y);
                res.x += v.x, res.y += v.y, res.z += v.z, res.w += v.w;
            }
        const float n = (float)(blurring * blurring);
        res = make_float4(res.x / n, res.y / n, res.z / n, res.w / n);
    }
    float *px = image + (long long)P * 3;
    const float keep = 1.0f - res.w;
    px[0] = px[0] * keep + res.x, px[1] = px[1] * keep + res.y, px[2] = px[2] * keep + res.z;
}

// postp/ssao.py:52-56,65-96 with taa=True: make_sample() per sample from the hash stream seeded with (P.x, P.y, frame)
__global__ void __launch_bounds__(256)
k_ssao_render_taa(const long long *__restrict__ keys, const float *__restrict__ normals, const __grid_constant__ Cam cam, int nsamples,
                  float radius, float thresh, float factor, unsigned frame, float *__restrict__ ao) {
    const int P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= cam.W * cam.H) return;
    const int i = P / cam.H, j = P - i * cam.H;
    const V3 normal = v3(normals[(long long)P * 3], normals[(long long)P * 3 + 1], normals[(long long)P * 3 + 2]);
    const float px = (float)i + cam.bias[0], py = (float)j + cam.bias[1];
    const float vx = px / cam.fW * 2.0f - 1.0f, vy = py / cam.fH * 2.0f - 1.0f;
    const float vz = (float)(int)(keys[P] >> 32) / 1073741824.0f;
    const V3 pos = mapply_pos3(cam.V2W, vx, vy, vz);
    const V3 viewdir = view_direction<false>(cam, px, py);
    const float vradius = mapply_pos3(cam.W2V, pos.x - radius * viewdir.x, pos.y - radius * viewdir.y, pos.z - radius * viewdir.z).z - vz;
    WangRNG rng;
    rng.seed = wang_hash(frame ^ wang_hash((unsigned)j ^ wang_hash((unsigned)i)));
    float occ = 0.0f;
    for (int s = 0; s < nsamples; s++) {
        float u = rng.random();
        const float v = rng.random();
        const float w = powf(rng.random(), 1.5f);
        const float r = 0.01f * (1.0f - w) + 1.0f * w; // common.py:221-223 lerp(w, 0.01, 1)
        u = 0.01f * (1.0f - u) + 1.0f * u;
        const V3 dir = tangent_spherical(normal, u, v, r, true);
        const V3 sv = mapply_pos3(cam.W2V, pos.x + dir.x * radius, pos.y + dir.y * radius, pos.z + dir.z * radius);
        const float Dx = (sv.x * 0.5f + 0.5f) * cam.fW, Dy = (sv.y * 0.5f + 0.5f) * cam.fH;
        if (0.0f <= Dx && Dx < cam.fW && 0.0f <= Dy && Dy < cam.fH) {
            const float d = (float)(int)(keys[(long long)f2i(Dx) * cam.H + f2i(Dy)] >> 32) / 1073741824.0f;
            if (d < sv.z) {
                const float rc = vradius / (vz - d);
                const float t = clamp01((fabsf(rc) - 0.0f) / (1.0f - 0.0f));
                occ += t * t * (3.0f - 2.0f * t);
            }
        }
    }
    float a = occ / (float)nsamples;
    a = factor * (a - thresh);
    ao[P] = clamp01(a);
}
__global__ void k_ssao_apply_taa(float *__restrict__ image, const float *__restrict__ ao, long long npix) { // ssao.py:40-41
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= npix) return;
    const float f = 1.0f - ao[P];
    image[P * 3] *= f, image[P * 3 + 1] *= f, image[P * 3 + 2] *= f;
}
