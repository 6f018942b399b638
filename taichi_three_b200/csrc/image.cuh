// image.cuh -- part of the single translation unit tina_b200.cu (included once, in order): FXAA, bloom, the division self-test, clear / fill / tonemap / accumulate.
#pragma once

// ------------------------------------------------------------------------------------
// image-space post effects (postp/fxaa.py, postp/blooming.py); fields are x-major [W][H]
// ------------------------------------------------------------------------------------
// a dense field read outside its shape yields 0 here (the reference reads out of bounds at the borders)
__device__ __forceinline__ float ld2(const float *f, int W, int H, int x, int y) {
    return (x < 0 || y < 0 || x >= W || y >= H) ? 0.0f : f[(long long)x * H + y];
}
__device__ __forceinline__ V3 ld2v(const float *f, int W, int H, int x, int y) {
    if (x < 0 || y < 0 || x >= W || y >= H) return v3(0.f, 0.f, 0.f);
    const float *p = f + ((long long)x * H + y) * 3;
    return v3(p[0], p[1], p[2]);
}
// common.py:140-149 on a vec3 field
__device__ __forceinline__ V3 bilerp3(const float *f, int W, int H, float px, float py) {
    const int I0 = f2i(floorf(px)), I1 = f2i(floorf(py));
    const float x0 = px - (float)I0, x1 = py - (float)I1, y0 = 1.0f - x0, y1 = 1.0f - x1;
    const V3 a = ld2v(f, W, H, I0 + 1, I1 + 1), b = ld2v(f, W, H, I0 + 1, I1), c = ld2v(f, W, H, I0, I1), d = ld2v(f, W, H, I0, I1 + 1);
    return v3(((a.x * x0 * x1 + b.x * x0 * y1) + c.x * y0 * y1) + d.x * y0 * x1,
              ((a.y * x0 * x1 + b.y * x0 * y1) + c.y * y0 * y1) + d.y * y0 * x1,
              ((a.z * x0 * x1 + b.z * x0 * y1) + c.z * y0 * y1) + d.z * y0 * x1);
}
__device__ __forceinline__ float clamp01(float x) { return fminf(1.0f, fmaxf(0.0f, x)); }

// fxaa.py:29-32
__global__ void k_fxaa_lumi(const float *__restrict__ image, float *__restrict__ lumi, float *__restrict__ copy, long long npix) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float r = image[i * 3], g = image[i * 3 + 1], b = image[i * 3 + 2];
    lumi[i] = clamp01((0.2989f * r + 0.587f * g) + 0.114f * b);
    copy[i * 3] = r, copy[i * 3 + 1] = g, copy[i * 3 + 2] = b;
}
// fxaa.py:33-68
__global__ void k_fxaa_apply(float *__restrict__ image, const float *__restrict__ lumi, const float *__restrict__ copy, int W,
                             int H, float abs_thresh, float rel_thresh, float factor) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)W * H) return;
    const int x = (int)(i / H), y = (int)(i - (long long)x * H);
    const float m = lumi[i], n = ld2(lumi, W, H, x, y + 1), e = ld2(lumi, W, H, x + 1, y), s = ld2(lumi, W, H, x, y - 1);
    const float w = ld2(lumi, W, H, x - 1, y), ne = ld2(lumi, W, H, x + 1, y + 1), nw = ld2(lumi, W, H, x - 1, y + 1);
    const float se = ld2(lumi, W, H, x + 1, y - 1), sw = ld2(lumi, W, H, x - 1, y - 1);
    const float hi = fmaxf(fmaxf(fmaxf(fmaxf(m, n), e), s), w), lo = fminf(fminf(fminf(fminf(m, n), e), s), w);
    const float c = hi - lo;
    if (c < abs_thresh || c < rel_thresh * hi) return;
    float filt = 2.0f * (((n + e) + s) + w);
    filt += ((ne + nw) + se) + sw;
    filt = fabsf(filt / 12.0f - m);
    filt = clamp01(filt / c);
    const float t = clamp01((filt - 0.0f) / (1.0f - 0.0f)); // smoothstep (common.py:203-205)
    const float sm = t * t * (3.0f - 2.0f * t);
    float blend = (sm * sm) * factor;
    float hori = fabsf((n + s) - 2.0f * m) * 2.0f;
    hori += fabsf((ne + se) - 2.0f * e);
    hori += fabsf((nw + sw) - 2.0f * w);
    float vert = fabsf((e + w) - 2.0f * m) * 2.0f;
    vert += fabsf((ne + nw) - 2.0f * n);
    vert += fabsf((se + sw) - 2.0f * s);
    const bool is_hori = hori >= vert;
    const float plumi = is_hori ? n : e, nlumi = is_hori ? s : w;
    if (fabsf(plumi - m) < fabsf(nlumi - m)) blend = -blend;
    const V3 r = bilerp3(copy, W, H, (float)x + blend * (is_hori ? 0.0f : 1.0f), (float)y + blend * (is_hori ? 1.0f : 0.0f));
    image[i * 3] = r.x, image[i * 3 + 1] = r.y, image[i * 3 + 2] = r.z;
}

// blooming.py:39-43 filter + :47-51 2x2 average into the half-resolution buffer
__device__ __forceinline__ float bloom_filter(float x, float thresh, float scale, float factor) {
    float t = fmaxf(0.0f, x - thresh);
    t = 1.0f - 1.0f / (1.0f + scale * t);
    return factor * t;
}
__global__ void k_bloom_down(const float *__restrict__ image, float *__restrict__ half, int W, int H, int hw, int hh, float thresh,
                             float scale, float factor) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)hw * hh) return;
    const int x = (int)(i / hh), y = (int)(i - (long long)x * hh);
    float r[3] = {0.f, 0.f, 0.f};
    for (int jx = 0; jx < 2; jx++)
        for (int jy = 0; jy < 2; jy++) {
            const V3 c = ld2v(image, W, H, x * 2 + jx, y * 2 + jy);
            r[0] += bloom_filter(c.x, thresh, scale, factor), r[1] += bloom_filter(c.y, thresh, scale, factor);
            r[2] += bloom_filter(c.z, thresh, scale, factor);
        }
    half[i * 3] = r[0] / 4.0f, half[i * 3 + 1] = r[1] / 4.0f, half[i * 3 + 2] = r[2] / 4.0f;
}
// blooming.py:52-67 separable blur with clamped taps; axis 0 = x, 1 = y
__global__ void k_bloom_blur(const float *__restrict__ src, float *__restrict__ dst, int hw, int hh, const float *__restrict__ gwei,
                             int radius, int axis) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)hw * hh) return;
    const int x = (int)(i / hh), y = (int)(i - (long long)x * hh);
    const float g0 = gwei[0];
    float r0 = src[i * 3] * g0, r1 = src[i * 3 + 1] * g0, r2 = src[i * 3 + 2] * g0;
    for (int k = 1; k <= radius; k++) {
        const int xa = axis ? x : max(0, x - k), ya = axis ? max(0, y - k) : y;
        const int xb = axis ? x : min(hw - 1, x + k), yb = axis ? min(hh - 1, y + k) : y;
        const float *a = src + ((long long)xa * hh + ya) * 3, *b = src + ((long long)xb * hh + yb) * 3;
        const float g = gwei[k];
        r0 += (a[0] + b[0]) * g, r1 += (a[1] + b[1]) * g, r2 += (a[2] + b[2]) * g;
    }
    dst[i * 3] = r0, dst[i * 3 + 1] = r1, dst[i * 3 + 2] = r2;
}
// blooming.py:68: image[I] += bilerp(img, I / 2)
__global__ void k_bloom_up(float *__restrict__ image, const float *__restrict__ half, int W, int H, int hw, int hh) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)W * H) return;
    const int x = (int)(i / H), y = (int)(i - (long long)x * H);
    const V3 b = bilerp3(half, hw, hh, (float)x / 2.0f, (float)y / 2.0f);
    image[i * 3] += b.x, image[i * 3 + 1] += b.y, image[i * 3 + 2] += b.z;
}

static int material_kind(const TinaMaterial *m) {
    const TinaInstr *c = m->code;
    auto isc = [&](int i) { return c[i].op == TINA_OP_CONST || c[i].op == TINA_OP_REG; };
    if (m->n_brdf == 1 && isc(0)) return MAT_CONST;
    if (m->n_brdf == 5 && isc(0) && isc(1) && isc(2) && c[3].op == TINA_OP_PHONG && c[4].op == TINA_OP_MIX) return MAT_CLASSIC;
    if (m->n_brdf == 6 && isc(0) && isc(1) && isc(2) && isc(3) && c[4].op == TINA_OP_COOK && c[5].op == TINA_OP_MIX)
        return MAT_PBR;
    return MAT_GENERIC;
}

// ---- SSAO (postp/ssao.py, non-TAA mode: sample and rotation tables fixed at construction) ----
// postp/ssao.py:65-96 render_at; same op order as the reference (the AO term has hard tests: d < sample.z, D inside)
__global__ void __launch_bounds__(256)
k_ssao_render(const long long *__restrict__ keys, const float *__restrict__ normals, const __grid_constant__ Cam cam,
              const float *__restrict__ samples, int nsamples, const float *__restrict__ rotations, int noise, float radius,
              float thresh, float factor, float *__restrict__ ao) {
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
    const V3 bitan = normalized(cross3(normal, v3(233.0f, 666.0f, 512.0f))); // advans.py:97-100 tangentspace
    const V3 tan = cross3(bitan, normal);
    const float *rot = rotations + ((long long)(i % noise) * noise + (j % noise)) * 2;
    const float r0 = rot[0], r1 = rot[1];
    float occ = 0.0f;
    for (int s = 0; s < nsamples; s++) {
        const float s0 = __ldg(samples + s * 3), s1 = __ldg(samples + s * 3 + 1), sz = __ldg(samples + s * 3 + 2);
        const float sx = r0 * s0 + r1 * s1, sy = -r0 * s0 + r1 * s1; // ssao.py:86-88 (sic: not a rotation matrix)
        const V3 sv = mapply_pos3(cam.W2V, pos.x + ((tan.x * sx + bitan.x * sy) + normal.x * sz) * radius,
                                  pos.y + ((tan.y * sx + bitan.y * sy) + normal.y * sz) * radius,
                                  pos.z + ((tan.z * sx + bitan.z * sy) + normal.z * sz) * radius);
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

// postp/ssao.py:38-49 apply: out *= 1 - box(noise x noise)(ao) / noise^2 (reads outside the field are 0)
__global__ void k_ssao_apply(float *__restrict__ image, const float *__restrict__ ao, int W, int H, int noise) {
    const int P = blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= W * H) return;
    const int i = P / H, j = P - i * H, offs = noise / 2;
    float r = 0.0f;
    for (int k = 0; k < noise; k++)
        for (int l = 0; l < noise; l++) r += ld2(ao, W, H, i + k - offs, j + l - offs);
    const float f = 1.0f - r / (float)(noise * noise);
    image[(long long)P * 3] *= f, image[(long long)P * 3 + 1] *= f, image[(long long)P * 3 + 2] *= f;
}

// ------------------------------------------------------------------------------------
// self-test: div_many against __fdiv_rn
// ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long splitmix(unsigned long long &x) {
    unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void k_selftest_division(unsigned long long per_thread, unsigned long long seed, unsigned long long *mismatch) {
    unsigned long long st = seed + 0x632be59bd9b4e019ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long bad = 0;
    for (unsigned long long i = 0; i < per_thread; i++) {
        const unsigned long long r0 = splitmix(st), r1 = splitmix(st);
        float d = __uint_as_float((unsigned)r0), a[3] = {__uint_as_float((unsigned)(r0 >> 32)), __uint_as_float((unsigned)r1),
                                                          __uint_as_float((unsigned)(r1 >> 32))};
        const unsigned mode = (unsigned)(i & 7);
        if (mode >= 2) { // operands of moderate exponent (the window the shared path serves), random mantissas
            const unsigned ex = 127u - 40u + (unsigned)(splitmix(st) % 81u);
            d = __uint_as_float((__float_as_uint(d) & 0x807fffffu) | (ex << 23));
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const unsigned ea = 127u - 40u + (unsigned)(splitmix(st) % 81u);
                a[k] = __uint_as_float((__float_as_uint(a[k]) & 0x807fffffu) | (ea << 23));
            }
            if (mode == 3) a[0] = d;                                              // quotient exactly 1
            if (mode == 4) a[1] = __uint_as_float(__float_as_uint(d) + 1u);       // quotient just above 1
            if (mode == 5) a[2] = 0.0f, a[0] = -0.0f;                             // signed zeros
            if (mode == 6) d = __uint_as_float((__float_as_uint(d) & 0xff800000u) | 0x7fffffu); // all-ones mantissa
            if (mode == 7) a[0] = 1.0f;                                           // reciprocals
        }
        float q[3];
        div_many(a, d, q);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float ref = __fdiv_rn(a[k], d);
            const bool same = __float_as_uint(ref) == __float_as_uint(q[k]) || (ref != ref && q[k] != q[k]);
            bad += same ? 0 : 1;
        }
    }
    if (bad) atomicAdd(mismatch, bad);
}
extern "C" int tina_selftest_division(int device, uint64_t nquotients, uint64_t seed, uint64_t *mismatch_host) {
    if (!mismatch_host) return fail(-1, "tina_selftest_division: null argument");
    DevGuard guard_(device);
    unsigned long long *d_bad = nullptr;
    CK(cudaMalloc(&d_bad, sizeof *d_bad));
    CK(cudaMemset(d_bad, 0, sizeof *d_bad));
    const unsigned blocks = 148 * 8, threads = 256;
    const unsigned long long per = (nquotients / 3 + (unsigned long long)blocks * threads - 1) / ((unsigned long long)blocks * threads);
    k_selftest_division<<<blocks, threads>>>(per, seed, d_bad);
    cudaError_t err = cudaDeviceSynchronize();
    unsigned long long bad = 0;
    if (err == cudaSuccess) err = cudaMemcpy(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost);
    cudaFree(d_bad);
    if (err != cudaSuccess) return fail(-2, "selftest: %s", cudaGetErrorString(err));
    *mismatch_host = bad;
    return 0;
}

// ------------------------------------------------------------------------------------
// small full-screen kernels
// ------------------------------------------------------------------------------------
// one 256-thread block per 256-pixel coverage chunk; selective: chunks nobody wrote since the last clear (flag 0) are skipped
__global__ void k_clear_keys(long long *keys, int n, unsigned char *blkflags, int selective) {
    pdl_launch_dependents(); // let the next kernel's launch overlap this one (it waits before touching memory)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (selective && blkflags[blockIdx.x] == 0) return;
    if (i < n) keys[i] = (long long)MAXDEPTH_I << 32; // engine.py:68-70, winner = none
    __syncthreads(); // (every thread has read the flag)
    if (threadIdx.x == 0) blkflags[blockIdx.x] = 0;
}
__global__ void k_depth(const long long *keys, int32_t *depth, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) depth[i] = (int32_t)(keys[i] >> 32);
}
__global__ void k_occup(const long long *keys, int32_t *occup, int n, unsigned base, unsigned nfaces) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned id = (unsigned)(unsigned long long)keys[i];
    unsigned f = id - 1u - base;
    occup[i] = (id != 0u && f < nfaces) ? (int32_t)f : -1;
}
__global__ void k_fill(float *img, long long npix, float r, float g, float b) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) img[i * 3] = r, img[i * 3 + 1] = g, img[i * 3 + 2] = b;
}
__global__ void k_tonemap(float *img, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) img[i] = aces(img[i]);
}
// util/accumator.py:16-23: img = img * (1 - 1/count) + src * (1/count)
__global__ void k_accumulate(float *acc, const float *src, long long n, int count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inv = __fdiv_rn(1.0f, (float)count);
    acc[i] = __fadd_rn(__fmul_rn(acc[i], __fsub_rn(1.0f, inv)), __fmul_rn(src[i], inv));
}
__global__ void k_tonemap4(float4 *img, long long n4) { // 16-byte aligned images: 128-bit accesses
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 v = img[i];
        img[i] = make_float4(aces(v.x), aces(v.y), aces(v.z), aces(v.w));
    }
}
