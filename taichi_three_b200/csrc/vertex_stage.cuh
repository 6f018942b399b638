// vertex_stage.cuh -- part of the single translation unit tina_b200.cu (included once, in order): K0 set_object adapters and the per-unique-vertex stage of indexed sources.
#pragma once

// ------------------------------------------------------------------------------------
// K0: set_object adapters
// ------------------------------------------------------------------------------------
struct Xform {
    float t[16];
    float tn[9];
    int has_t;
};

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
    bool flip = (mode & 2u) || ((mode & 1u) && (n & 1));
    bool neg = ((mode & 1u) && (n & 1)) != ((mode & 4u) != 0);
    int ks = flip ? 2 - k : k;
    const int32_t *fc = faces + (src * 3 + ks) * 3;
    {
        const float *p = v + (long long)(uint32_t)fc[0] * 3;
        float a = p[0], b = p[1], c = p[2];
        if (X.has_t) {
            V3 r = mapply_pos3(X.t, a, b, c);
            a = r.x, b = r.y, c = r.z;
        }
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
        if (X.has_t) { // trans.py:38-40: trans_normal @ norm, not re-normalised
            float ra = (X.tn[0] * a + X.tn[1] * b) + X.tn[2] * c;
            float rb = (X.tn[3] * a + X.tn[4] * b) + X.tn[5] * c;
            float rc = (X.tn[6] * a + X.tn[7] * b) + X.tn[8] * c;
            a = ra, b = rb, c = rc;
        }
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
    bool flip = (mode & 2u) || ((mode & 1u) && (n & 1));
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
        if (X.has_t) {
            V3 r = mapply_pos3(X.t, a, b, c);
            a = r.x, b = r.y, c = r.z;
        }
        overts[t * 3] = a, overts[t * 3 + 1] = b, overts[t * 3 + 2] = c;
    }
    if (ocoors) { // grid.py:17-21: I / (res - 1)
        ocoors[t * 2] = (float)ci / (float)(nx - 1);
        ocoors[t * 2 + 1] = (float)cj / (float)(ny - 1);
    }
    if (onorms) {
        float a = nrm[vi * 3], b = nrm[vi * 3 + 1], c = nrm[vi * 3 + 2];
        if (X.has_t) {
            float ra = (X.tn[0] * a + X.tn[1] * b) + X.tn[2] * c;
            float rb = (X.tn[3] * a + X.tn[4] * b) + X.tn[5] * c;
            float rc = (X.tn[6] * a + X.tn[7] * b) + X.tn[8] * c;
            a = ra, b = rb, c = rc;
        }
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
        V3 r = mapply_pos3(X.t, v[t * 3], v[t * 3 + 1], v[t * 3 + 2]);
        vpos[t * 3] = r.x, vpos[t * 3 + 1] = r.y, vpos[t * 3 + 2] = r.z;
    }
    if (vnrm && t < nvn) {
        const float a = vn[t * 3], b = vn[t * 3 + 1], c = vn[t * 3 + 2];
        vnrm[t * 3] = (X.tn[0] * a + X.tn[1] * b) + X.tn[2] * c;
        vnrm[t * 3 + 1] = (X.tn[3] * a + X.tn[4] * b) + X.tn[5] * c;
        vnrm[t * 3 + 2] = (X.tn[6] * a + X.tn[7] * b) + X.tn[8] * c;
    }
}

// (the camera-dependent part of the vertex stage, k_frame_prologue, lives in raster_indexed.cuh)

// pars/trans.py:22-31
__global__ void k_pars_transform(const float *__restrict__ v, const float *__restrict__ sz, long long n,
                                 const __grid_constant__ Xform X, float scale, float *__restrict__ ov, float *__restrict__ osz) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    V3 r = mapply_pos3(X.t, v[t * 3], v[t * 3 + 1], v[t * 3 + 2]);
    ov[t * 3] = r.x, ov[t * 3 + 1] = r.y, ov[t * 3 + 2] = r.z;
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
    float4 *recA; // per-vertex records, written by k_frame_prologue in render_occup (raster_indexed.cuh)
    uint4 *recB;
    int64_t vpos_cap, vnrm_cap, recA_cap, recB_cap, nv, nvn;
    int force_general; // tuning knob 15: mark every vertex non-tame (every face takes K1's general path)
};
