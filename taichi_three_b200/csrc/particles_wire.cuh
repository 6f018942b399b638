// particles_wire.cuh -- part of the single translation unit tina_b200.cu (included once, in order): ParticleRaster and WireframeRaster kernels.
#pragma once

// ------------------------------------------------------------------------------------
// ParticleRaster (core/particle.py:78-148): sphere splats sharing the engine's key buffer
// ------------------------------------------------------------------------------------
struct ParSetup {
    float ax, ay, az;     // particle centre (world)
    float rl;             // radius
    float avz;            // NDC z of the centre (= depth of every covered pixel)
    int botx, boty, topx, topy;
};

// common.py:186-189 mapply_dir(M, d) for a unit axis, then .normalized()
__device__ __forceinline__ V3 axis_dir(const float *M, float d0, float d1, float d2) {
    float r0, r1, r2, rw;
    mapply(M, d0, d1, d2, 0.0f, r0, r1, r2, rw);
    return normalized(v3(r0, r1, r2));
}

// particle.py:96-120.  returns false when the particle is clipped
__device__ __forceinline__ bool par_setup(const float *__restrict__ verts, const float *__restrict__ sizes, long long f,
                                          const Cam &cam, uint32_t flags, ParSetup &s) {
    s.ax = __ldg(verts + f * 3), s.ay = __ldg(verts + f * 3 + 1), s.az = __ldg(verts + f * 3 + 2);
    s.rl = __ldg(sizes + f);
    const V3 av = mapply_pos3(cam.W2V, s.ax, s.ay, s.az);
    s.avz = av.z;
    if ((flags & 2u) && !((-1.0f <= av.z) & (av.z <= 1.0f))) return false;
    const V3 dx = axis_dir(cam.V2W, 1.f, 0.f, 0.f), dy = axis_dir(cam.V2W, 0.f, 1.f, 0.f);
    const float rvx = mapply_pos3(cam.W2V, s.ax + dx.x * s.rl, s.ay + dx.y * s.rl, s.az + dx.z * s.rl).x - av.x;
    const float rvy = mapply_pos3(cam.W2V, s.ax + dy.x * s.rl, s.ay + dy.y * s.rl, s.az + dy.z * s.rl).y - av.y;
    const float rx = cam.fW, ry = cam.fH;
    // Bv = [Av - (Rv.x,0,0), Av + (Rv.x,0,0), Av - (0,Rv.y,0), Av + (0,Rv.y,0)]; b = to_viewport(Bv)
    const float b0x = ((av.x - rvx) * 0.5f + 0.5f) * rx, b0y = ((av.y - 0.0f) * 0.5f + 0.5f) * ry;
    const float b1x = ((av.x + rvx) * 0.5f + 0.5f) * rx, b1y = ((av.y + 0.0f) * 0.5f + 0.5f) * ry;
    const float b2x = ((av.x - 0.0f) * 0.5f + 0.5f) * rx, b2y = ((av.y - rvy) * 0.5f + 0.5f) * ry;
    const float b3x = ((av.x + 0.0f) * 0.5f + 0.5f) * rx, b3y = ((av.y + rvy) * 0.5f + 0.5f) * ry;
    s.botx = max(ifloor_x86(fminf(b0x, b2x)), 0), s.boty = max(ifloor_x86(fminf(b0y, b2y)), 0);
    s.topx = min(iceil_x86(fmaxf(b1x, b3x)), cam.W - 1), s.topy = min(iceil_x86(fmaxf(b1y, b3y)), cam.H - 1);
    return true;
}

// particle.py:121-127: world position of the pixel on the particle's depth plane; inside the sphere?
__device__ __forceinline__ bool par_hit(const ParSetup &s, const Cam &cam, int x, int y, V3 &pl) {
    const float px = (float)x + cam.bias[0], py = (float)y + cam.bias[1];
    pl = mapply_pos3(cam.V2W, px / cam.fW * 2.0f - 1.0f, py / cam.fH * 2.0f - 1.0f, s.avz);
    const float dx = pl.x - s.ax, dy = pl.y - s.ay, dz = pl.z - s.az;
    return !((dx * dx + dy * dy) + dz * dz > s.rl * s.rl);
}

// one thread per particle; small discs are walked by their thread, bigger ones by the whole warp
__global__ void __launch_bounds__(256)
k_pars_occup(const float *__restrict__ verts, const float *__restrict__ sizes, long long npars, const __grid_constant__ Cam cam,
             uint32_t flags, unsigned base, long long *__restrict__ keys, unsigned char *__restrict__ blkflags) {
    pdl_wait();
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    ParSetup s;
    bool ok = false;
    int area = 0;
    if (f < npars) {
        ok = par_setup(verts, sizes, f, cam, flags, s);
        if (ok) {
            const int w = s.topx - s.botx + 1, h = s.topy - s.boty + 1;
            area = (w > 0 && h > 0) ? w * h : 0;
        }
    }
    const unsigned id = base + (unsigned)f + 1u;
    if (ok && area > 0 && area <= 32) {
        const long long key = pack_key(f2i(s.avz * 1073741824.0f), id);
        for (int x = s.botx; x <= s.topx; x++)
            for (int y = s.boty; y <= s.topy; y++) {
                V3 pl;
                if (!par_hit(s, cam, x, y, pl)) continue;
                const long long P = (long long)x * cam.H + y;
                atomicMin(keys + P, key);
                blkflags[P >> FLAG_SHIFT] = 1;
            }
    }
    unsigned big = __ballot_sync(0xffffffffu, ok && area > 32);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        ParSetup t;
        t.ax = __shfl_sync(0xffffffffu, s.ax, src), t.ay = __shfl_sync(0xffffffffu, s.ay, src);
        t.az = __shfl_sync(0xffffffffu, s.az, src), t.rl = __shfl_sync(0xffffffffu, s.rl, src);
        t.avz = __shfl_sync(0xffffffffu, s.avz, src);
        t.botx = __shfl_sync(0xffffffffu, s.botx, src), t.boty = __shfl_sync(0xffffffffu, s.boty, src);
        t.topx = __shfl_sync(0xffffffffu, s.topx, src), t.topy = __shfl_sync(0xffffffffu, s.topy, src);
        const unsigned tid_ = __shfl_sync(0xffffffffu, id, src);
        const long long key = pack_key(f2i(t.avz * 1073741824.0f), tid_);
        const int h = t.topy - t.boty + 1, n = (t.topx - t.botx + 1) * h;
        const float rh = __frcp_rn((float)h);
        for (int k = (int)lane; k < n; k += 32) {
            const int q = (int)(((float)k + 0.5f) * rh); // k / h, exact for k < 2^21
            const int x = t.botx + q, y = t.boty + (k - q * h);
            V3 pl;
            if (!par_hit(t, cam, x, y, pl)) continue;
            const long long P = (long long)x * cam.H + y;
            atomicMin(keys + P, key);
            blkflags[P >> FLAG_SHIFT] = 1;
        }
    }
}

// particle.py:129-161 + shader.py:119-131 + lighting.py:84-98
template <int KIND>
__global__ void __launch_bounds__(256)
k_pars_color(const long long *__restrict__ keys, const float *__restrict__ verts, const float *__restrict__ sizes,
             const float *__restrict__ colors, const __grid_constant__ Cam cam, unsigned base, unsigned npars,
             const __grid_constant__ TinaMaterial mat, const __grid_constant__ TinaLighting L, float *__restrict__ image,
             uint32_t cflags, float bg0, float bg1, float bg2, const unsigned char *__restrict__ blkflags) {
    pdl_wait();
    const int npix = cam.W * cam.H;
    const int P = blockIdx.x * 256 + threadIdx.x;
    if (P >= npix) return;
    float *out = image + (long long)P * 3;
    unsigned f = 0xffffffffu;
    if (blkflags[blockIdx.x]) {
        const unsigned id = (unsigned)(unsigned long long)keys[P];
        f = id - 1u - base;
        if (id == 0u) f = 0xffffffffu;
    }
    if (f >= npars) { // particle.py:131-133 (occup == -1)
        if (cflags & TINA_COLOR_FILL_BG) {
            float r = bg0, g = bg1, b = bg2;
            if (cflags & TINA_COLOR_TONEMAP) r = aces(r), g = aces(g), b = aces(b);
            out[0] = r, out[1] = g, out[2] = b;
        }
        return;
    }
    const int x = P / cam.H, y = P - x * cam.H;
    ParSetup s;
    s.ax = __ldg(verts + (long long)f * 3), s.ay = __ldg(verts + (long long)f * 3 + 1), s.az = __ldg(verts + (long long)f * 3 + 2);
    s.rl = __ldg(sizes + f);
    s.avz = mapply_pos3(cam.W2V, s.ax, s.ay, s.az).z;
    V3 pl;
    par_hit(s, cam, x, y, pl);
    // Dl = (Pl - Al) / Rl;  Dl -= Zl * sqrt(1 - |Dl|^2);  Dl = Dl.normalized()
    V3 d = v3((pl.x - s.ax) / s.rl, (pl.y - s.ay) / s.rl, (pl.z - s.az) / s.rl);
    const V3 zl = axis_dir(cam.V2W, 0.f, 0.f, 1.f);
    const float t = sqrtf(1.0f - dot3(d, d));
    d = normalized(v3(d.x - zl.x * t, d.y - zl.y * t, d.z - zl.z * t));
    ShadeIn in;
    in.normal = d;
    in.pos = v3(s.ax + d.x * s.rl, s.ay + d.y * s.rl, s.az + d.z * s.rl);
    in.texcoord = v3(0.f, 0.f, 0.f);
    in.color = colors ? v3(__ldg(colors + (long long)f * 3), __ldg(colors + (long long)f * 3 + 1), __ldg(colors + (long long)f * 3 + 2))
                      : v3(1.f, 1.f, 1.f);
    const float px = (float)x + cam.bias[0], py = (float)y + cam.bias[1];
    V3 c = light_pixel<KIND>(in, view_direction(cam, px, py), mat, L);
    if (cflags & TINA_COLOR_TONEMAP) c.x = aces(c.x), c.y = aces(c.y), c.z = aces(c.z);
    out[0] = c.x, out[1] = c.y, out[2] = c.z;
}

// G-buffer sinks for particles: ParticleRaster.render_color hands pos / normal / texcoord / color of the visible sphere
// point to ANY shader (particle.py:129-161), e.g. the NormalShader of Scene(ssao=True) -- same inputs as k_pars_color
__global__ void __launch_bounds__(256)
k_pars_gbuffer(const long long *__restrict__ keys, const float *__restrict__ verts, const float *__restrict__ sizes,
               const float *__restrict__ colors, const __grid_constant__ Cam cam, unsigned base, unsigned npars,
               const __grid_constant__ SinkTab T) {
    pdl_wait();
    const int P = blockIdx.x * 256 + threadIdx.x;
    if (P >= cam.W * cam.H) return;
    const long long key = keys[P];
    const unsigned id = (unsigned)(unsigned long long)key;
    const unsigned f = id - 1u - base;
    if (id == 0u || f >= npars) return; // particle.py:131-133
    const int x = P / cam.H, y = P - x * cam.H;
    ParSetup s;
    s.ax = __ldg(verts + (long long)f * 3), s.ay = __ldg(verts + (long long)f * 3 + 1), s.az = __ldg(verts + (long long)f * 3 + 2);
    s.rl = __ldg(sizes + f);
    s.avz = mapply_pos3(cam.W2V, s.ax, s.ay, s.az).z;
    V3 pl;
    par_hit(s, cam, x, y, pl);
    V3 d = v3((pl.x - s.ax) / s.rl, (pl.y - s.ay) / s.rl, (pl.z - s.az) / s.rl);
    const V3 zl = axis_dir(cam.V2W, 0.f, 0.f, 1.f);
    const float t = sqrtf(1.0f - dot3(d, d));
    d = normalized(v3(d.x - zl.x * t, d.y - zl.y * t, d.z - zl.z * t));
    ShadeIn in;
    in.normal = d;
    in.pos = v3(s.ax + d.x * s.rl, s.ay + d.y * s.rl, s.az + d.z * s.rl);
    in.texcoord = v3(0.f, 0.f, 0.f);
    in.color = colors ? v3(__ldg(colors + (long long)f * 3), __ldg(colors + (long long)f * 3 + 1), __ldg(colors + (long long)f * 3 + 2))
                      : v3(1.f, 1.f, 1.f);
    const float px = (float)x + cam.bias[0], py = (float)y + cam.bias[1];
    write_sinks(T, P, key, f, in, px, py, view_direction(cam, px, py), cam);
}

// ------------------------------------------------------------------------------------
// WireframeRaster (core/wireframe.py:70-95): depth-tested DDA lines on the engine's key buffer
// ------------------------------------------------------------------------------------
struct WireSetup {
    float ax, ay, kx, ky;  // viewport start, DDA step (wireframe.py:49-66)
    float w0, w1, z0, z1;  // 1/w and NDC z of the two ends
    int siz;               // steps: i = 0..siz
    int i0, i1;            // sub-range that can touch the screen
};

// wireframe.py:73-86 + draw_line set-up.  false = clipped / nothing to draw
__device__ __forceinline__ bool wire_setup(const float *__restrict__ v, const Cam &cam, uint32_t flags, WireSetup &s) {
    float ax, ay, az, aw, bx, by, bz, bw;
    mapply(cam.W2V, v[0], v[1], v[2], 1.0f, ax, ay, az, aw);
    mapply(cam.W2V, v[3], v[4], v[5], 1.0f, bx, by, bz, bw);
    ax = fd(ax, aw), ay = fd(ay, aw), az = fd(az, aw);
    bx = fd(bx, bw), by = fd(by, bw), bz = fd(bz, bw);
    if (flags & 2u) {
        const bool ina = in_unit2(ax, ay) & (fabsf(az) <= 1.0f), inb = in_unit2(bx, by) & (fabsf(bz) <= 1.0f);
        if (!ina && !inb) return false;
    }
    const float rx = cam.fW, ry = cam.fH;
    const float pax = fm(fa(fm(ax, 0.5f), 0.5f), rx), pay = fm(fa(fm(ay, 0.5f), 0.5f), ry);
    const float pbx = fm(fa(fm(bx, 0.5f), 0.5f), rx), pby = fm(fa(fm(by, 0.5f), 0.5f), ry);
    const float dx = fs(pbx, pax), dy = fs(pby, pay), adx = fabsf(dx), ady = fabsf(dy);
    s.kx = 1.0f, s.ky = 1.0f;
    bool xmajor = adx >= ady;
    if (xmajor) {
        s.kx = dx >= 0.0f ? 1.0f : -1.0f;
        s.ky = fd(fm(s.kx, dy), dx);
        s.siz = f2i(adx);
    } else {
        s.ky = dy >= 0.0f ? 1.0f : -1.0f;
        s.kx = fd(fm(s.ky, dx), dy);
        s.siz = f2i(ady);
    }
    s.ax = pax, s.ay = pay;
    s.w0 = fd(1.0f, aw), s.w1 = fd(1.0f, bw), s.z0 = az, s.z1 = bz;
    if (s.siz < 0) return false; // range(siz + 1) is empty
    // i-range whose major coordinate can fall on the screen (conservative, the per-pixel test stays exact)
    const double am = xmajor ? (double)pax : (double)pay, sg = xmajor ? (double)s.kx : (double)s.ky;
    const double lim = xmajor ? (double)cam.W : (double)cam.H;
    const double mg = 2.0 + 1e-6 * (fabs(am) + lim + (double)s.siz); // >= 8 ulp of the f32 positions involved
    double t0 = (-mg - am) / sg, t1 = (lim + mg - am) / sg;
    if (t0 > t1) {
        const double t = t0;
        t0 = t1, t1 = t;
    }
    if (!(t0 == t0) || !(t1 == t1)) t0 = 0.0, t1 = (double)s.siz; // NaN start: let the per-pixel test decide
    s.i0 = (int)fmax(0.0, floor(t0)), s.i1 = (int)fmin((double)s.siz, ceil(t1));
    return s.i0 <= s.i1;
}

// wireframe.py:87-95 for step i: pixel + depth; false = off screen
__device__ __forceinline__ bool wire_pixel(const WireSetup &s, const Cam &cam, int i, int &P, int &depth) {
    const float fi = (float)i;
    const float px = fa(fa(s.ax, fm(s.kx, fi)), cam.bias[0]), py = fa(fa(s.ay, fm(s.ky, fi)), cam.bias[1]);
    const int x = ifloor_x86(px), y = ifloor_x86(py);
    if (x < 0 || x >= cam.W || y < 0 || y >= cam.H) return false;
    const float cor = fd(fi, (float)s.siz);
    float w0 = fm(fs(1.0f, cor), s.w0), w1 = fm(cor, s.w1);
    const float sum = fa(w0, w1);
    w0 = fd(w0, sum), w1 = fd(w1, sum);
    depth = f2i(fm(fa(fm(w0, s.z0), fm(w1, s.z1)), 1073741824.0f));
    P = x * cam.H + y;
    return true;
}

__global__ void __launch_bounds__(256)
k_wire_occup(const float *__restrict__ verts, long long nwires, const __grid_constant__ Cam cam, uint32_t flags, unsigned base,
             long long *__restrict__ keys, unsigned char *__restrict__ blkflags) {
    pdl_wait();
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    WireSetup s;
    bool ok = false;
    if (f < nwires) {
        float v[6];
#pragma unroll
        for (int k = 0; k < 6; k++) v[k] = __ldg(verts + f * 6 + k);
        ok = wire_setup(v, cam, flags, s);
    }
    const unsigned id = base + (unsigned)f + 1u;
    const int len = ok ? s.i1 - s.i0 + 1 : 0;
    if (ok && len <= 48) {
        for (int i = s.i0; i <= s.i1; i++) {
            int P, d;
            if (!wire_pixel(s, cam, i, P, d)) continue;
            atomicMin(keys + P, pack_key(d, id));
            blkflags[P >> FLAG_SHIFT] = 1;
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, ok && len > 48);
    while (big) { // long lines: the whole warp walks them
        const int src = __ffs(big) - 1;
        big &= big - 1;
        WireSetup t;
        t.ax = __shfl_sync(0xffffffffu, s.ax, src), t.ay = __shfl_sync(0xffffffffu, s.ay, src);
        t.kx = __shfl_sync(0xffffffffu, s.kx, src), t.ky = __shfl_sync(0xffffffffu, s.ky, src);
        t.w0 = __shfl_sync(0xffffffffu, s.w0, src), t.w1 = __shfl_sync(0xffffffffu, s.w1, src);
        t.z0 = __shfl_sync(0xffffffffu, s.z0, src), t.z1 = __shfl_sync(0xffffffffu, s.z1, src);
        t.siz = __shfl_sync(0xffffffffu, s.siz, src);
        t.i0 = __shfl_sync(0xffffffffu, s.i0, src), t.i1 = __shfl_sync(0xffffffffu, s.i1, src);
        const unsigned tid_ = __shfl_sync(0xffffffffu, id, src);
        for (int i = t.i0 + (int)lane; i <= t.i1; i += 32) {
            int P, d;
            if (!wire_pixel(t, cam, i, P, d)) continue;
            atomicMin(keys + P, pack_key(d, tid_));
            blkflags[P >> FLAG_SHIFT] = 1;
        }
    }
}

// Shader.blend_color(factor = 1) (shader.py:133-135) where a wire of this object owns the pixel
__global__ void k_wire_color(const long long *__restrict__ keys, unsigned base, unsigned nwires, float *__restrict__ image, int npix,
                             float c0, float c1, float c2, const unsigned char *__restrict__ blkflags) {
    pdl_wait();
    const int P = blockIdx.x * 256 + threadIdx.x;
    if (P >= npix || !blkflags[blockIdx.x]) return;
    const unsigned id = (unsigned)(unsigned long long)keys[P];
    if (id == 0u || id - 1u - base >= nwires) return;
    float *o = image + (long long)P * 3; // lerp(1, img, color) = img * (1 - 1) + color * 1
    o[0] = o[0] * 0.0f + c0 * 1.0f, o[1] = o[1] * 0.0f + c1 * 1.0f, o[2] = o[2] * 0.0f + c2 * 1.0f;
}

// wires of a polygon mesh (mesh/wire.py:19-27): wire n = corners (n % p, (n + 1) % p) of face n / p
__global__ void k_wires_from_faces(const float *__restrict__ faces, long long nwires, int npoly, float *__restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nwires) return;
    const long long f = n / npoly;
    const int e1 = (int)(n % npoly), e2 = (int)((n + 1) % npoly);
    const float *a = faces + (f * npoly + e1) * 3, *b = faces + (f * npoly + e2) * 3;
    float *o = out + n * 6;
    o[0] = a[0], o[1] = a[1], o[2] = a[2], o[3] = b[0], o[4] = b[1], o[5] = b[2];
}
