// vertex_stage.cuh -- part of the single translation unit tina_b200.cu (included once, in order): K0 set_object adapters and the per-unique-vertex stage of indexed sources.
#pragma once

// ------------------------------------------------------------------------------------
// K0: set_object adapters
// ------------------------------------------------------------------------------------
#define TINA_MAX_XFORMS 4 /* nested MeshTransform wrappers, applied innermost first like the reference's call chain */
struct Xform {
    float t[TINA_MAX_XFORMS][16];
    float tn[TINA_MAX_XFORMS][9];
    int has_t; // number of transforms in the chain
};
// mesh/trans.py:28-40 for a chain of wrappers: one rounding sequence per wrapper, innermost first
__device__ __forceinline__ void xform_pos(const Xform &X, float &a, float &b, float &c) {
    for (int k = 0; k < X.has_t; k++) {
        V3 r = mapply_pos3(X.t[k], a, b, c);
        a = r.x, b = r.y, c = r.z;
    }
}
__device__ __forceinline__ void xform_nrm(const Xform &X, float &a, float &b, float &c) { // trans_normal @ norm, not re-normalised
    for (int k = 0; k < X.has_t; k++) {
        const float *tn = X.tn[k];
        const float ra = (tn[0] * a + tn[1] * b) + tn[2] * c;
        const float rb = (tn[3] * a + tn[4] * b) + tn[5] * c;
        const float rc = (tn[6] * a + tn[7] * b) + tn[8] * c;
        a = ra, b = rb, c = rc;
    }
}

// mesh/model.py:56-73 (+ trans.py:28-40, cull.py:6-57).  One thread per output corner.
__global__ void k_gather_indexed(const float *__restrict__ v, const float *__restrict__ vt, const float *__restrict__ vn,
                                 const int32_t *__restrict__ faces, long long nout, const __grid_constant__ Xform X,
                                 uint32_t mode, float *__restrict__ overts, float *__restrict__ onorms,
                                 float *__restrict__ ocoors) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nout * 3) return;
    long long n = t / 3;
    int k = (int)(t - n * 3);
    long long src = (mode & 1u) ? (n >> 1) : n;
    bool flip = ((mode & 2u) != 0) != (((mode & 1u) != 0) && (n & 1)); // a flip around a double-sided wrapper un-reverses the odd copies
    bool neg = ((mode & 1u) && (n & 1)) != ((mode & 4u) != 0);
    int ks = flip ? 2 - k : k;
    const int32_t *fc = faces + (src * 3 + ks) * 3;
    {
        const float *p = v + (long long)(uint32_t)fc[0] * 3;
        float a = p[0], b = p[1], c = p[2];
        xform_pos(X, a, b, c);
        float *o = overts + t * 3;
        o[0] = a, o[1] = b, o[2] = c;
    }
    if (ocoors) {
        const float *p = vt + (long long)(uint32_t)fc[1] * 2;
        ocoors[t * 2] = p[0], ocoors[t * 2 + 1] = p[1];
    }
    if (onorms) {
        const float *p = vn + (long long)(uint32_t)fc[2] * 3;
        float a = p[0], b = p[1], c = p[2];
        xform_nrm(X, a, b, c);
        if (neg) a = -a, b = -b, c = -c;
        float *o = onorms + t * 3;
        o[0] = a, o[1] = b, o[2] = c;
    }
}

// mesh/grid.py:26-35
__global__ void k_grid_normals(const float *__restrict__ pos, int nx, int ny, float *__restrict__ nrm) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)nx * ny) return;
    int i = (int)(t / ny), j = (int)(t - (long long)i * ny);
    int i2 = max(i - 1, 0), j2 = max(j - 1, 0), i1 = min(i + 1, nx - 1), j1 = min(j + 1, ny - 1);
    const float *pa = pos + ((long long)i * ny + j1) * 3, *pb = pos + ((long long)i * ny + j2) * 3;
    const float *pc = pos + ((long long)i1 * ny + j) * 3, *pd = pos + ((long long)i2 * ny + j) * 3;
    V3 dy = v3(pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]);
    V3 dx = v3(pc[0] - pd[0], pc[1] - pd[1], pc[2] - pd[2]);
    V3 r = normalized(cross3(dx, dy));
    nrm[t * 3] = r.x, nrm[t * 3 + 1] = r.y, nrm[t * 3 + 2] = r.z;
}

// mesh/grid.py:45-58 (+ trans / cull wrappers).  One thread per output corner.
__global__ void k_grid_faces(const float *__restrict__ pos, const float *__restrict__ nrm, int nx, int ny, long long nout,
                             const __grid_constant__ Xform X, uint32_t mode, float *__restrict__ overts,
                             float *__restrict__ onorms, float *__restrict__ ocoors) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nout * 3) return;
    long long n = t / 3;
    int k = (int)(t - n * 3);
    long long src = (mode & 1u) ? (n >> 1) : n;
    bool flip = ((mode & 2u) != 0) != (((mode & 1u) != 0) && (n & 1)); // a flip around a double-sided wrapper un-reverses the odd copies
    bool neg = ((mode & 1u) && (n & 1)) != ((mode & 4u) != 0);
    int ks = flip ? 2 - k : k;
    const int stride = nx - 1; // sic (grid.py:46)
    long long m = src >> 1;
    int i = (int)(m / stride), j = (int)(m % stride);
    // corners a=[i,j] b=[i+1,j] c=[i+1,j+1] d=[i,j+1]; even: (a,b,c), odd: (a,c,d)
    int ci, cj;
    if (ks == 0) ci = i, cj = j;
    else if ((src & 1) == 0) ci = i + 1, cj = (ks == 1) ? j : j + 1;
    else ci = (ks == 1) ? i + 1 : i, cj = j + 1;
    ci = min(ci, nx - 1), cj = min(cj, ny - 1); // (reference: out of bounds for nx != ny, grid.py:46)
    long long vi = (long long)ci * ny + cj;
    {
        float a = pos[vi * 3], b = pos[vi * 3 + 1], c = pos[vi * 3 + 2];
        xform_pos(X, a, b, c);
        overts[t * 3] = a, overts[t * 3 + 1] = b, overts[t * 3 + 2] = c;
    }
    if (ocoors) { // grid.py:17-21: I / (res - 1)
        ocoors[t * 2] = (float)ci / (float)(nx - 1);
        ocoors[t * 2 + 1] = (float)cj / (float)(ny - 1);
    }
    if (onorms) {
        float a = nrm[vi * 3], b = nrm[vi * 3 + 1], c = nrm[vi * 3 + 2];
        xform_nrm(X, a, b, c);
        if (neg) a = -a, b = -b, c = -c;
        onorms[t * 3] = a, onorms[t * 3 + 1] = b, onorms[t * 3 + 2] = c;
    }
}

// ------------------------------------------------------------------------------------
// vertex stage for indexed sources (MeshGrid / MeshModel): per UNIQUE vertex / normal
// ------------------------------------------------------------------------------------
// mesh/trans.py:28-40 per unique vertex / normal (instead of per face corner)
__global__ void k_vtx_world(const float *__restrict__ v, long long nv, const float *__restrict__ vn, long long nvn,
                            const __grid_constant__ Xform X, float *__restrict__ vpos, float *__restrict__ vnrm) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nv) {
        float a = v[t * 3], b = v[t * 3 + 1], c = v[t * 3 + 2];
        xform_pos(X, a, b, c);
        vpos[t * 3] = a, vpos[t * 3 + 1] = b, vpos[t * 3 + 2] = c;
    }
    if (vnrm && t < nvn) {
        float a = vn[t * 3], b = vn[t * 3 + 1], c = vn[t * 3 + 2];
        xform_nrm(X, a, b, c);
        vnrm[t * 3] = a, vnrm[t * 3 + 1] = b, vnrm[t * 3 + 2] = c;
    }
}

// (the camera-dependent part of the vertex stage, k_frame_prologue, lives in raster_indexed.cuh)

// pars/trans.py:22-31
__global__ void k_pars_transform(const float *__restrict__ v, const float *__restrict__ sz, long long n,
                                 const __grid_constant__ Xform X, float scale, float *__restrict__ ov, float *__restrict__ osz) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float a = v[t * 3], b = v[t * 3 + 1], c = v[t * 3 + 2];
    xform_pos(X, a, b, c);
    ov[t * 3] = a, ov[t * 3 + 1] = b, ov[t * 3 + 2] = c;
    osz[t] = scale * sz[t];
}

struct IndexedState {
    Src src;         // src.kind != 0: K1/K3/K4 fetch corners through the mesh's own indexing
    int enabled;     // tuning knob 11
    int expanded;    // overts / onorms / ocoors hold the current object
    // arguments of the last set_faces_indexed / _grid, for lazy materialisation of the expanded arrays
    const float *a_v, *a_vt, *a_vn, *a_pos;
    const int32_t *a_faces;
    int a_nx, a_ny;
    Xform a_X;
    uint32_t a_mode;
    int64_t a_nout;
    // owned per-vertex buffers
    float *vpos_w, *vnrm_w;
    // per-vertex records, written by k_frame_prologue in render_occup (raster_indexed.cuh).  Two sets used by alternate
    // render_occup calls: the vertex stage of the next call may then run while the shading kernel of this one still
    // reads its records (programmatic dependent launch, see k_frame_prologue); recA / recB = the set of the last call
    float4 *recA, *recA2[2];
    uint4 *recB, *recB2[2];
    unsigned rec_parity;
    int64_t vpos_cap, vnrm_cap, recA_cap[2], recB_cap[2], nv, nvn;
    int force_general; // tuning knob 15: mark every vertex non-tame (every face takes K1's general path)
};
