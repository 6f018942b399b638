// raster.cuh -- part of the single translation unit tina_b200.cu (included once, in order): K1 k_raster_faces (phase A, compaction, phase B candidate walk) and K2+K3 k_large_path (tile path).
#pragma once

// ------------------------------------------------------------------------------------
// K1: transform + cull/clip/bbox (phase A), compaction, setup + coverage + atomicMin (phase B)
// ------------------------------------------------------------------------------------
// Candidate tightening.  The reference tests every pixel P of the integer bbox
// [floor(min), ceil(max)] (triangle.py:108-114) at the sample s = P + bias.  A sample whose x
// (or y) lies outside the vertices' float range by more than a margin mu is rejected by the
// reference for every *well-conditioned* triangle, so those pixels need not be visited:
//   with exact barycentrics l_k of s (sum 1), sx < minx - mu gives sum_k l_k (v_kx - sx) = 0
//   with every (v_kx - sx) in (mu, D], hence some l_i < -mu/(2D) and some l_j >= 1/3.
//   The reference's computed weights differ from l_k by at most eta = (2 rho + 15 eps) max(1, Rb),
//   rho = 2^-20 + 2^-22 the relative error of bcn/can when the area n has not cancelled by
//   more than 4x (guard G1), Rb = 2 Lmax^2 / |n| >= |l_k| and >= the magnitude of every term,
//   Lmax = extent + 2 >= |s - v|.  Under guard G3 (Lmax * max(1, Rb) <= 512) eta << mu/(2D),
//   so computed weight i is a normal negative number and weight j a normal positive one; with
//   all 1/w in [2^-20, 2^20] (G0) the products keep those signs, the quotients by `sum` have
//   opposite signs (or are +-inf), and `all(wei >= 0)` (triangle.py:120) is false.
// Faces failing any guard walk the full reference bbox.  mu: TIGHTEN_M = 2^-5 minus the
// rounding of the bound computation (<= 3 ulp at |coord| <= 2^15, guard G2) > 0.019.
// tests/test_gpu_parity.py::test_tightening_is_exact checks tightened == untightened bits on
// adversarial micro-triangle / sliver sets.
#define TIGHTEN_M 0.03125f

struct FaceA {          // phase-A result for one face
    float ax, ay, bx, by, cx, cy; // viewport coords (engine.py:60-61)
    float zc0, zc1, zc2;          // clip-space z (divided by w in phase B)
    float w0, w1, w2;             // clip-space w
    int botx, boty, topx, topy;   // reference bbox (clamped)
    int xlo, ylo, xhi, yhi;       // candidate range actually walked
};

// -1 <= fd(zc, w) <= 1 without the division in the common case
__device__ __forceinline__ bool z_in_range(float zc, float w) {
    const float az = fabsf(zc);
    if (w > 0.0f && az <= w) return true;                 // |zc/w| <= 1 => |fd| <= 1 (rounding is monotonic)
    if (w > 0.0f && az > fm(w, 1.000001f) && w < 1e30f) return false; // ratio > 1 + 2^-24 => fd > 1
    const float z = fd(zc, w);
    return (-1.0f <= z) & (z <= 1.0f);
}

// ---- where a face's corners live ------------------------------------------------------------
// kind 0: expanded [N,3,3] arrays (SimpleMesh, or after tina_raster_materialize)
// kind 1/2: the mesh's own indexing (MeshGrid / MeshModel) over per-UNIQUE-vertex arrays written by
// the vertex stage (k_vtx_*): world position, world normal, and clip coordinates.  Every vertex is
// shared by ~6 faces, so transforming it once instead of once per face removes most of phase A's
// arithmetic and lets K1/K4 gather from a few tens of MB that stay L2-resident instead of the
// expanded copies.  Per-vertex values are computed with the same ops => same bits.
struct FastDiv { // unsigned division by a launch-invariant divisor (Granlund-Montgomery)
    unsigned mul, sh1, sh2, d;
};
__host__ __device__ inline unsigned fastdiv(unsigned n, const FastDiv &f) {
#ifdef __CUDA_ARCH__
    const unsigned t = __umulhi(f.mul, n);
#else
    const unsigned t = (unsigned)(((unsigned long long)f.mul * n) >> 32);
#endif
    return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
static FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d;
    unsigned l = 0;
    while ((1ull << l) < d) l++;
    f.mul = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.sh1 = l < 1 ? l : 1;
    f.sh2 = l > 0 ? l - 1 : 0;
    return f;
}

struct Src {
    int kind;
    uint32_t mode;          // 1 double sided (MeshNoCulling), 2 flip winding, 4 negate normals
    int nx, ny;             // grid
    FastDiv div_stride;     // grid: division by (nx - 1)
    const int32_t *faces;   // model: [N,3,3] = [corner][v, vt, vn]
    const float *vpos;      // world positions per unique vertex
    const float *vnrm;      // world normals per unique normal
    const float *vtex;      // model: texture coordinates per unique vt
    // per-unique-vertex records written by the vertex stage (k_frame_prologue, raster_indexed.cuh):
    const float4 *recA;     // (viewport x, viewport y, z/w, 1/w): engine.py:52-53,60-61 + triangle.py:113
    const uint4 *recB;      // (candidate lower bounds x|y<<16, upper bounds x|y<<16, bits of x/w, bits of y/w)
};

// corner k of output face n -> vertex / texcoord / normal ids (mesh/grid.py:45-58, mesh/model.py:56-73,
// mesh/cull.py:6-57).  For grids it[] is unused and (gi, gj) are the corner's grid coordinates.
// CK = 0: kind and mode read from S; CK = 1 / 2: compile-time kind (grid / model) with mode 0 (lean kernels)
template <int CK = 0>
__device__ __forceinline__ void corner_ids(const Src &S, long long n, int iv[3], int it[3], int in_[3], int gi[3], int gj[3],
                                           bool &neg) {
    const uint32_t mode = CK ? 0u : S.mode;
    const int kind = CK ? CK : S.kind;
    const long long src = (mode & 1u) ? (n >> 1) : n;
    const bool odd = (mode & 1u) && (n & 1);
    const bool flip = ((mode & 2u) != 0) != odd;
    neg = odd != ((mode & 4u) != 0);
    if (kind == 1) {
        const unsigned stride = (unsigned)(S.nx - 1); // sic (grid.py:46)
        const unsigned m = (unsigned)(src >> 1);
        const unsigned qi = fastdiv(m, S.div_stride);
        const int i = (int)qi, j = (int)(m - qi * stride);
        const bool second = (src & 1) != 0; // even: (a,b,c), odd: (a,c,d); a=[i,j] b=[i+1,j] c=[i+1,j+1] d=[i,j+1]
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int ks = flip ? 2 - k : k;
            int ci, cj;
            if (ks == 0) ci = i, cj = j;
            else if (!second) ci = i + 1, cj = (ks == 1) ? j : j + 1;
            else ci = (ks == 1) ? i + 1 : i, cj = j + 1;
            // (the reference indexes out of bounds for nx != ny, grid.py:46; stay inside the arrays)
            ci = min(ci, S.nx - 1), cj = min(cj, S.ny - 1);
            gi[k] = ci, gj[k] = cj;
            iv[k] = in_[k] = it[k] = ci * S.ny + cj;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int ks = flip ? 2 - k : k;
            const int32_t *fc = S.faces + (src * 3 + ks) * 3;
            iv[k] = __ldg(fc), it[k] = __ldg(fc + 1), in_[k] = __ldg(fc + 2);
            gi[k] = gj[k] = 0;
        }
    }
}

// world-space corner positions of face f (kind 0: expanded array)
__device__ __forceinline__ void face_world_verts(const Src &S, const float *__restrict__ verts, long long f, float vv[9]) {
    if (S.kind == 0) {
        const float *v = verts + f * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) vv[k] = __ldg(v + k);
    } else {
        int iv[3], it[3], in_[3], gi[3], gj[3];
        bool neg;
        corner_ids(S, f, iv, it, in_, gi, gj, neg);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float *p = S.vpos + (long long)iv[k] * 3;
            vv[k * 3] = __ldg(p), vv[k * 3 + 1] = __ldg(p + 1), vv[k * 3 + 2] = __ldg(p + 2);
        }
    }
}

// engine.py:52-53 for one vertex, kept un-divided in z and w: (x/w, y/w, z_clip, w_clip)
__device__ __forceinline__ float4 vertex_clip(const Cam &cam, float p0, float p1, float p2) {
    float x, y, z, w;
    mapply(cam.W2V, p0, p1, p2, 1.0f, x, y, z, w);
    const float a[2] = {x, y};
    float q[2];
    div_many(a, w, q);
    return make_float4(q[0], q[1], z, w);
}

__device__ __forceinline__ int face_phase_a_clip(float4 ca, float4 cb, float4 cc, const Cam &cam, uint32_t flags, int tighten,
                                                 FaceA &f);

// triangle.py:93-109.  returns 0 ok, 1 culled, 2 clipped
__device__ __forceinline__ int face_phase_a(const float *v, const Cam &cam, uint32_t flags, int tighten, FaceA &f) {
    return face_phase_a_clip(vertex_clip(cam, v[0], v[1], v[2]), vertex_clip(cam, v[3], v[4], v[5]),
                             vertex_clip(cam, v[6], v[7], v[8]), cam, flags, tighten, f);
}

__device__ __forceinline__ int face_phase_a_clip(float4 ca, float4 cb, float4 cc, const Cam &cam, uint32_t flags, int tighten,
                                                 FaceA &f) {
    const float ax = ca.x, ay = ca.y, bx = cb.x, by = cb.y, cx = cc.x, cy = cc.y;
    f.zc0 = ca.z, f.w0 = ca.w, f.zc1 = cb.z, f.w1 = cb.w, f.zc2 = cc.z, f.w2 = cc.w;
    float facing = fs(fm(fs(bx, ax), fs(cy, ay)), fm(fs(by, ay), fs(cx, ax)));
    if (facing <= 0.0f && (flags & TINA_CULLING)) return 1;
    if (flags & TINA_CLIPPING) {
        bool ina = in_unit2(ax, ay), inb = in_unit2(bx, by), inc = in_unit2(cx, cy);
        if (ina) ina = z_in_range(f.zc0, f.w0);
        if (!ina && inb) inb = z_in_range(f.zc1, f.w1);
        if (!ina && !inb && inc) inc = z_in_range(f.zc2, f.w2);
        if (!ina && !inb && !inc) return 2;
    }
    const float rx = cam.fW, ry = cam.fH;
    f.ax = fm(fa(fm(ax, 0.5f), 0.5f), rx), f.ay = fm(fa(fm(ay, 0.5f), 0.5f), ry);
    f.bx = fm(fa(fm(bx, 0.5f), 0.5f), rx), f.by = fm(fa(fm(by, 0.5f), 0.5f), ry);
    f.cx = fm(fa(fm(cx, 0.5f), 0.5f), rx), f.cy = fm(fa(fm(cy, 0.5f), 0.5f), ry);
    const float minx = fminf(fminf(f.ax, f.bx), f.cx), miny = fminf(fminf(f.ay, f.by), f.cy);
    const float maxx = fmaxf(fmaxf(f.ax, f.bx), f.cx), maxy = fmaxf(fmaxf(f.ay, f.by), f.cy);
    f.botx = max(ifloor_x86(minx), 0), f.boty = max(ifloor_x86(miny), 0);
    f.topx = min(iceil_x86(maxx), cam.W - 1), f.topy = min(iceil_x86(maxy), cam.H - 1);
    f.xlo = f.botx, f.ylo = f.boty, f.xhi = f.topx, f.yhi = f.topy;
    if (tighten) {
        const float P1 = fm(fs(f.bx, f.ax), fs(f.cy, f.ay)), P2 = fm(fs(f.by, f.ay), fs(f.cx, f.ax));
        const float n = fabsf(fs(P1, P2));
        const float ext = fmaxf(maxx - minx, maxy - miny), L = ext + 2.0f;
        const float wmin = fminf(fminf(f.w0, f.w1), f.w2), wmax = fmaxf(fmaxf(f.w0, f.w1), f.w2);
        bool ok = (wmin >= 9.5367431640625e-07f) & (wmax <= 1048576.0f);                        // G0
        ok &= n >= 0.25f * (fabsf(P1) + fabsf(P2));                                              // G1
        ok &= (minx >= -32768.0f) & (miny >= -32768.0f) & (maxx <= 32768.0f) & (maxy <= 32768.0f); // G2
        ok &= (L * fmaxf(n, 2.0f * L * L) <= 512.0f * n);                                        // G3: L*max(1,Rb) <= 512
        // (bias in [0, 1] is checked once on the host: tina_raster_render_occup drops `tighten` otherwise)
        if (ok) {
            f.xlo = max(f.botx, __float2int_ru(fs(fs(minx, TIGHTEN_M), cam.bias[0])));
            f.xhi = min(f.topx, __float2int_rd(fs(fa(maxx, TIGHTEN_M), cam.bias[0])));
            f.ylo = max(f.boty, __float2int_ru(fs(fs(miny, TIGHTEN_M), cam.bias[1])));
            f.yhi = min(f.topy, __float2int_rd(fs(fa(maxy, TIGHTEN_M), cam.bias[1])));
        }
    }
    return 0;
}

// triangle.py:110-113 from the phase-A record (same ops as setup_face => same bits)
__device__ __forceinline__ void face_phase_b(const FaceA &f, Setup &s) {
    float n = fs(fm(fs(f.bx, f.ax), fs(f.cy, f.ay)), fm(fs(f.by, f.ay), fs(f.cx, f.ax)));
    {
        const float a[4] = {fs(f.bx, f.cx), fs(f.by, f.cy), fs(f.cx, f.ax), fs(f.cy, f.ay)};
        float q[4];
        div_many(a, n, q);
        s.bcnx = q[0], s.bcny = q[1], s.canx = q[2], s.cany = q[3];
    }
    s.bx = f.bx, s.by = f.by, s.cx = f.cx, s.cy = f.cy;
    {
        float q[2];
        const float a0[2] = {1.0f, f.zc0}, a1[2] = {1.0f, f.zc1}, a2[2] = {1.0f, f.zc2};
        div_many(a0, f.w0, q), s.w0 = q[0], s.z0 = q[1];
        div_many(a1, f.w1, q), s.w1 = q[0], s.z1 = q[1];
        div_many(a2, f.w2, q), s.w2 = q[0], s.z2 = q[1];
    }
}

#define SURV_WORDS 15            /* planes of the survivor records between phase A and B */
#define WALK_MAX_T 8192          /* most candidate pixels one warp deals out in the shared walk (32 x tiny_max 256) */
#define WALK_WORDS (32 * 20 + WALK_MAX_T / 32 + 8) /* per warp: setups [32][5] float4, start bits, rank table */
#define HQ_CAP 64 /* per-warp deferred-hit queue entries */
// Append this warp's large faces to the tile-path queue (warp-aggregated), with their finished edge setups, and
// add the warp's stats.  Called by whole warps.
// `mk(Setup&)` computes the face's finished edge setup (only evaluated for queued faces).
template <typename MakeSetup>
__device__ __forceinline__ void queue_large_faces(int botx, int boty, int topx, int topy, MakeSetup mk, bool big, bool queued,
                                                  bool surv, int rc, unsigned gface,
                                                  unsigned lane, uint4 *__restrict__ queue, unsigned *__restrict__ counters,
                                                  unsigned queue_cap, float4 *__restrict__ qsetup, unsigned qsetup_cap,
                                                  int inline_large, int collect_stats) {
    // queue the large ones for the tile path (warp-aggregated append)
    {
        if (inline_large) {
            const unsigned bm = __ballot_sync(0xffffffffu, big);
            if (bm && lane == 0) atomicAdd(&counters[0], __popc(bm));
        }
        const unsigned qm = __ballot_sync(0xffffffffu, queued);
        if (qm) {
            unsigned slot = 0;
            if (lane == (unsigned)(__ffs(qm) - 1)) slot = atomicAdd(&counters[0], __popc(qm));
            slot = __shfl_sync(0xffffffffu, slot, __ffs(qm) - 1);
            if (queued) {
                unsigned my = slot + __popc(qm & ((1u << lane) - 1u));
                if (my < queue_cap)
                    queue[my] = make_uint4(gface, (unsigned)botx | ((unsigned)boty << 16),
                                           (unsigned)topx | ((unsigned)topy << 16), 0u);
                if (my < qsetup_cap) { // finished edge setup, so that no tile has to redo its 16 divisions
                    Setup q;
                    mk(q);
                    float4 *o = qsetup + (size_t)my * 4;
                    o[0] = make_float4(q.bcnx, q.bcny, q.canx, q.cany);
                    o[1] = make_float4(q.bx, q.by, q.cx, q.cy);
                    o[2] = make_float4(q.w0, q.w1, q.w2, q.z0);
                    o[3] = make_float4(q.z1, q.z2, 0.f, 0.f);
                }
            }
        }
        if (collect_stats) {
            unsigned m1 = __ballot_sync(0xffffffffu, rc == 1), m2 = __ballot_sync(0xffffffffu, rc == 2);
            unsigned m3 = __ballot_sync(0xffffffffu, surv);
            if (lane == 0) {
                if (m1) atomicAdd(&counters[4], __popc(m1));
                if (m2) atomicAdd(&counters[5], __popc(m2));
                if (m3) atomicAdd(&counters[6], __popc(m3));
                if (qm) atomicAdd(&counters[7], __popc(qm));
            }
        }
    }
}

// Phase B walk of one warp's survivors (lane = one face: setup s, candidate range f.xlo..f.yhi, cnt candidates,
// id = global face id + 1).  `wk` is the warp's WALK_WORDS-word scratch region in shared memory (free for its use),
// `hq` the warp's deferred-hit queue.  Called by whole warps.
// SHARED = false: per-lane walk only (no scratch needed: wk / hq may be null).
template <bool SHARED = true>
__device__ __forceinline__ void walk_candidates(const FaceA &f, const Setup &s, unsigned id, int cnt, int col, unsigned lane,
                                                const Cam &cam, long long *__restrict__ keys,
                                                unsigned char *__restrict__ blkflags, unsigned char flagval, int precheck,
                                                int balance, float *wk, unsigned (*hq)[2]) {
    const float bxs = cam.bias[0], bys = cam.bias[1];
    // How uneven is this warp?  M = longest lane, T = total candidate pixels.
    int M = 0, T = 0;
    bool shared_walk = false;
    if (SHARED) {
        M = __reduce_max_sync(0xffffffffu, cnt), T = __reduce_add_sync(0xffffffffu, cnt); // REDUX: one instruction each
        shared_walk = (balance == 2) || (balance == 1 && (long long)M * 40 > (long long)((T + 31) >> 5) * 75 + 150);
    }
    if (!SHARED || !shared_walk || T > WALK_MAX_T) {
        // per-lane walk of the candidate range, x-outer / y-inner like triangle.py:114; the inner loop
        // only does the cheap exact reject, candidates fall out to the division + atomic part
        int x = f.xlo, y = f.ylo;
        while (x <= f.xhi) {
            PW w;
            int hx = x, hy = y;
            bool cand = false;
            while (x <= f.xhi) {
                w = pix_products(s, fa((float)x, bxs), fa((float)y, bys));
                hx = x, hy = y;
                if (++y > f.yhi) y = f.ylo, ++x;
                if (!pix_fast_reject(w)) {
                    cand = true;
                    break;
                }
            }
            if (cand) {
                float q0, q1, q2;
                if (pix_finish(s, w, q0, q1, q2)) {
                    long long key = pack_key(pix_depth(s, q0, q1, q2), id);
                    const long long P = (long long)hx * cam.H + hy;
                    long long *dst = keys + P;
                    if (!precheck || __ldcg(dst) > key) atomicMin(dst, key);
                    blkflags[P >> FLAG_SHIFT] = flagval;
                }
            }
        }
        return;
    }
    // Warp-shared walk (soups: lanes with 1 and lanes with 60 candidate pixels in one warp): the warp's T
    // candidate pixels are dealt 32 at a time to the lanes, and pixels that survive the cheap reject are parked in
    // a queue so that the division + atomic part always runs with full lanes.
    //   Who owns candidate k?  Face j's candidates are [off_j, off_j + cnt_j).  Every face with cnt > 0 sets bit
    //   off_j in a T-bit array; in iteration `it` all lanes look at the same word of it (bits 32 it .. 32 it + 31):
    //   owner(k) = rank-th non-empty face, rank = starts before the window (a running count) + starts at or below
    //   lane inside it - 1.  One broadcast load and a popcount instead of a five-step shuffle search.
    //   Setups live as 5 x float4 per face (stride 20 words: conflict-free for neighbouring faces), so a
    //   candidate costs four 128-bit shared loads instead of fourteen 32-bit ones.
    float4 *A = reinterpret_cast<float4 *>(wk);                    // [32][5]
    unsigned *bits = reinterpret_cast<unsigned *>(wk + 32 * 20);   // [WALK_MAX_T / 32]
    unsigned char *tab = reinterpret_cast<unsigned char *>(bits + WALK_MAX_T / 32); // [32] rank -> lane
    int off = cnt; // exclusive prefix sum of cnt over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, off, d);
        if ((int)lane >= d) off += t;
    }
    off -= cnt;
    const int ch_own = f.yhi - f.ylo + 1;
    A[lane * 5 + 0] = make_float4(s.bcnx, s.bcny, s.canx, s.cany);
    A[lane * 5 + 1] = make_float4(s.bx, s.by, s.cx, s.cy);
    A[lane * 5 + 2] = make_float4(s.w0, s.w1, s.w2, __int_as_float(f.xlo | (f.ylo << 16)));
    A[lane * 5 + 3] = make_float4(s.z0, s.z1, s.z2, __int_as_float((int)id));
    A[lane * 5 + 4] = make_float4(__int_as_float(ch_own), __frcp_rn((float)max(ch_own, 1)), __int_as_float(off), 0.0f);
    for (int w = (int)lane; w < (T + 31) >> 5; w += 32) bits[w] = 0u;
    const unsigned nz = __ballot_sync(0xffffffffu, cnt > 0);
    __syncwarp();
    if (cnt > 0) {
        tab[__popc(nz & ((1u << lane) - 1u))] = (unsigned char)lane;
        atomicOr(&bits[off >> 5], 1u << (off & 31));
    }
    __syncwarp();
    int hqn = 0;
    auto load_setup = [&](int j, Setup &t, float4 &c2) {
        const float4 c0 = A[j * 5 + 0], c1 = A[j * 5 + 1];
        c2 = A[j * 5 + 2];
        t.bcnx = c0.x, t.bcny = c0.y, t.canx = c0.z, t.cany = c0.w;
        t.bx = c1.x, t.by = c1.y, t.cx = c1.z, t.cy = c1.w;
        t.w0 = c2.x, t.w1 = c2.y, t.w2 = c2.z;
    };
    auto drain = [&](int e) { // finish one parked pixel: divisions, depth, atomicMin
        const int j = (int)hq[e][0];
        const int hx = (int)(hq[e][1] & 0xffffu), hy = (int)(hq[e][1] >> 16);
        Setup t;
        float4 c2;
        load_setup(j, t, c2);
        PW w = pix_products(t, fa((float)hx, bxs), fa((float)hy, bys));
        float q0, q1, q2;
        if (pix_finish(t, w, q0, q1, q2)) {
            const float4 c3 = A[j * 5 + 3];
            t.z0 = c3.x, t.z1 = c3.y, t.z2 = c3.z;
            long long key = pack_key(pix_depth(t, q0, q1, q2), (unsigned)__float_as_int(c3.w));
            const long long P = (long long)hx * cam.H + hy;
            long long *dst = keys + P;
            if (!precheck || __ldcg(dst) > key) atomicMin(dst, key);
            blkflags[P >> FLAG_SHIFT] = flagval;
        }
    };
    int before = 0; // non-empty faces that start before the current 32-candidate window
    for (int k0 = 0; k0 < T; k0 += 32) {
        const int k = k0 + (int)lane;
        const unsigned word = bits[k0 >> 5];
        bool cand = false;
        int x = 0, y = 0, j = 0;
        if (k < T) {
            j = tab[before + __popc(word & (0xffffffffu >> (31 - lane))) - 1];
            const float4 c4 = A[j * 5 + 4];
            const int p = k - __float_as_int(c4.z);
            const int ch = __float_as_int(c4.x);
            const int q = (int)(((float)p + 0.5f) * c4.y); // p / ch, exact for p < 2^21
            Setup t;
            float4 c2;
            load_setup(j, t, c2);
            const int xy = __float_as_int(c2.w);
            x = (xy & 0xffff) + q, y = (int)((unsigned)xy >> 16) + (p - q * ch);
            cand = !pix_fast_reject(pix_products(t, fa((float)x, bxs), fa((float)y, bys)));
        }
        before += __popc(word);
        const unsigned cm = __ballot_sync(0xffffffffu, cand);
        if (cand) {
            const int slot = hqn + __popc(cm & ((1u << lane) - 1u));
            hq[slot][0] = (unsigned)j, hq[slot][1] = (unsigned)x | ((unsigned)y << 16);
        }
        hqn += __popc(cm);
        __syncwarp();
        if (hqn >= 32) {
            hqn -= 32;
            drain(hqn + (int)lane);
            __syncwarp();
        }
    }
    if ((int)lane < hqn) drain((int)lane);
}

// stats layout in counters[]: [4] culled [5] clipped [6] survivors (phase B) [7] queued
// LEAN = 0: every option read at run time.  LEAN = 3: the default configuration as compile-time constants --
// culling + clipping on, tightening on, no key pre-read, no stats -- which removes the option tests from the
// per-face path.
// This kernel serves the expanded [N,3,3] arrays (SimpleMesh, soups); indexed sources (MeshGrid / MeshModel) have
// their own kernel on per-vertex records (raster_indexed.cuh).
template <int LEAN = 0>
// (register budget 41-42: the shared walk then keeps its shared-memory addresses in registers: C3 1.67 -> 1.57 ms)
__global__ void __launch_bounds__(K1_THREADS, 5)
k_raster_faces(const float *__restrict__ verts, long long nfaces, const __grid_constant__ Cam cam, uint32_t flags_rt,
               unsigned base, long long *__restrict__ keys, uint4 *__restrict__ queue, unsigned *__restrict__ counters,
               unsigned queue_cap, int tiny_max, int tighten_rt, int precheck_rt, int balance, int collect_stats_rt,
               const __grid_constant__ Src S, unsigned char *__restrict__ blkflags, unsigned *__restrict__ next_counters,
               int inline_large, float4 *__restrict__ qsetup, unsigned qsetup_cap, unsigned char flagval) {
    static_assert(LEAN == 0 || LEAN == 3, "lean variant: 3 = expanded arrays with the default options");
    const uint32_t flags = LEAN ? (uint32_t)(TINA_CULLING | TINA_CLIPPING) : flags_rt;
    const int tighten = LEAN ? 1 : tighten_rt, precheck = LEAN ? 0 : precheck_rt, collect_stats = LEAN ? 0 : collect_stats_rt;
    // staging of the CTA's vertices, later reused for the compacted survivor records (SoA)
    // staged vertices (expanded sources), then the compacted survivor records, then the warps' walk scratch
    constexpr int SM_WORDS = K1_THREADS * SURV_WORDS > (K1_THREADS / 32) * WALK_WORDS ? K1_THREADS * SURV_WORDS : (K1_THREADS / 32) * WALK_WORDS;
    __shared__ __align__(128) float sm[SM_WORDS];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ unsigned s_nsurv;
    __shared__ unsigned s_hq[K1_THREADS / 32][HQ_CAP][2];
    // PDL: only the launch latency is overlapped with the predecessor; every global access (the
    // vertices may have been written by the kernel just before us) comes after the wait
    pdl_wait();
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31;
    if (blockIdx.x == 0 && tid < 8) next_counters[tid] = 0u; // counter set of the NEXT render_occup (3 sets rotate)
    const long long f0 = (long long)blockIdx.x * K1_THREADS;
    const int n = (int)min((long long)K1_THREADS, nfaces - f0);
    const float *src = verts + f0 * 9;
    const int nfl = n * 9;
    if (tid == 0) s_nsurv = 0;
    // the CTA's 256 x 36 B of vertices arrive with ONE bulk-copy instruction (TMA, UBLKCP)
    const bool bulk = ((((uintptr_t)src) & 15) == 0) && ((nfl & 3) == 0);
    if (bulk) {
        if (tid == 0) mbar_init(&s_mbar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s_mbar, (uint32_t)nfl * 4u);
            bulk_g2s(sm, src, (uint32_t)nfl * 4u, &s_mbar);
        }
        mbar_wait(&s_mbar, 0);
    } else {
        for (int i = tid; i < nfl; i += K1_THREADS) sm[i] = __ldg(src + i);
        __syncthreads();
    }

    // ---- phase A ----
    FaceA f;
    int rc = 3; // 3 = inactive lane
    int cnt = 0, refarea = 0;
    if (tid < n) {
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = sm[tid * 9 + k];
        rc = face_phase_a(v, cam, flags, tighten, f);
        if (rc == 0) {
            const int rw_ = f.topx - f.botx + 1, rh_ = f.topy - f.boty + 1;
            refarea = (rw_ > 0 && rh_ > 0) ? rw_ * rh_ : 0;
            const int cw = f.xhi - f.xlo + 1, ch = f.yhi - f.ylo + 1;
            cnt = (refarea > 0 && cw > 0 && ch > 0) ? cw * ch : 0;
        }
    }
    // faces with many candidate pixels go to the tile path -- unless the host launched us without it
    // (inline_large: recent frames queued nothing); then they are walked here and only counted
    const bool big = (rc == 0) && (cnt > tiny_max);
    const bool queued = big && !inline_large;
    const bool surv = (rc == 0) && (cnt > 0) && !queued;
    __syncthreads(); // everyone has read its vertices: sm can be overwritten

    // compaction of survivors (warp-aggregated slots)
    {
        const unsigned m = __ballot_sync(0xffffffffu, surv);
        unsigned slot = 0;
        if (m) {
            if (lane == (unsigned)(__ffs(m) - 1)) slot = atomicAdd(&s_nsurv, __popc(m));
            slot = __shfl_sync(0xffffffffu, slot, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
        }
        if (surv) {
            float *r = sm + slot;
            r[0 * K1_THREADS] = f.ax, r[1 * K1_THREADS] = f.ay, r[2 * K1_THREADS] = f.bx, r[3 * K1_THREADS] = f.by;
            r[4 * K1_THREADS] = f.cx, r[5 * K1_THREADS] = f.cy;
            r[6 * K1_THREADS] = f.zc0, r[7 * K1_THREADS] = f.zc1, r[8 * K1_THREADS] = f.zc2;
            r[9 * K1_THREADS] = f.w0, r[10 * K1_THREADS] = f.w1, r[11 * K1_THREADS] = f.w2;
            r[12 * K1_THREADS] = __int_as_float(f.xlo | (f.xhi << 16));
            r[13 * K1_THREADS] = __int_as_float(f.ylo | (f.yhi << 16));
            r[14 * K1_THREADS] = __int_as_float(tid);
        }
    }
    queue_large_faces(f.botx, f.boty, f.topx, f.topy, [&](Setup &q) { face_phase_b(f, q); }, big, queued, surv, rc,
                      (unsigned)(f0 + tid), lane, queue, counters, queue_cap, qsetup, qsetup_cap, inline_large, collect_stats);
    __syncthreads();

    // ---- phase B: dense over survivors ----
    const int nsurv = (int)s_nsurv;
    const bool idle_warp = (tid & ~31) >= nsurv;
    const bool act = tid < nsurv;
    Setup s;
    unsigned id = 0;
    cnt = 0;
    f.xlo = f.ylo = 0, f.xhi = f.yhi = -1;
    if (act) {
        const float *r = sm + tid;
        f.ax = r[0 * K1_THREADS], f.ay = r[1 * K1_THREADS], f.bx = r[2 * K1_THREADS], f.by = r[3 * K1_THREADS];
        f.cx = r[4 * K1_THREADS], f.cy = r[5 * K1_THREADS];
        f.zc0 = r[6 * K1_THREADS], f.zc1 = r[7 * K1_THREADS], f.zc2 = r[8 * K1_THREADS];
        f.w0 = r[9 * K1_THREADS], f.w1 = r[10 * K1_THREADS], f.w2 = r[11 * K1_THREADS];
        const int xb = __float_as_int(r[12 * K1_THREADS]), yb = __float_as_int(r[13 * K1_THREADS]);
        f.xlo = xb & 0xffff, f.xhi = (int)((unsigned)xb >> 16), f.ylo = yb & 0xffff, f.yhi = (int)((unsigned)yb >> 16);
        id = base + (unsigned)(f0 + __float_as_int(r[14 * K1_THREADS])) + 1u;
        face_phase_b(f, s);
        cnt = (f.xhi - f.xlo + 1) * (f.yhi - f.ylo + 1);
    }
    __syncthreads(); // every warp has taken its survivors out of `sm`: from here on it is per-warp walk scratch
    if (idle_warp) return;
    walk_candidates(f, s, id, cnt, tid, lane, cam, keys, blkflags, flagval, precheck, balance, sm + (tid >> 5) * WALK_WORDS,
                    s_hq[tid >> 5]);
}

// ------------------------------------------------------------------------------------
// K2+K3: the tile path for queued (large) triangles, one cooperative persistent kernel
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_range(const uint4 &q, int &tx0, int &ty0, int &tx1, int &ty1) {
    tx0 = (int)(q.y & 0xffffu) / TILE, ty0 = (int)(q.y >> 16) / TILE;
    tx1 = (int)(q.z & 0xffffu) / TILE, ty1 = (int)(q.z >> 16) / TILE;
}

__device__ __forceinline__ unsigned ld_volatile(const unsigned *p) { return *((const volatile unsigned *)p); }

// sense-reversing grid barrier; the kernel is launched cooperatively so every CTA is resident
__device__ void grid_barrier(unsigned *bar) { // bar[0] = arrivals, bar[1] = generation
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = ld_volatile(&bar[1]);
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(&bar[1], 1u);
        } else {
            while (ld_volatile(&bar[1]) == gen) __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}

#define K3_CHUNK 128
struct SetupSoA {
    float f[14][K3_CHUNK];
    int bot[K3_CHUNK], top[K3_CHUNK];
    unsigned id[K3_CHUNK];
};

// One 16x16 tile, one thread per pixel ("pixel owner"): the tile's keys are read once,
// min-merged in registers against every listed triangle (setups staged in shared memory,
// broadcast reads), written back once, coalesced.  No atomics.
__device__ void raster_tile(int tile, bool scan_mode, unsigned nq, const Src &SRC, const float *__restrict__ verts, const Cam &cam,
                            unsigned base, long long *__restrict__ keys, const uint4 *__restrict__ queue,
                            const unsigned *__restrict__ tile_offs, const unsigned *__restrict__ tile_list, int tiles_y,
                            SetupSoA &S, unsigned &s_cnt, unsigned char *__restrict__ blkflags,
                            const float4 *__restrict__ qsetup, unsigned qsetup_cap, unsigned char flagval) {
    unsigned beg = 0, end = nq;
    if (!scan_mode) {
        beg = tile_offs[tile], end = tile_offs[tile + 1];
        if (beg == end) return;
    }
    const int tid = threadIdx.x;
    const int tx = tile / tiles_y, ty = tile % tiles_y;
    const int x0 = tx * TILE, y0 = ty * TILE;
    const int x = x0 + (tid >> 4), y = y0 + (tid & 15);
    const bool inb = (x < cam.W) && (y < cam.H);
    long long *dst = keys + ((long long)x * cam.H + y);
    long long orig = LLONG_MIN, best = LLONG_MIN;
    bool loaded = false;
    const float px = fa((float)x, cam.bias[0]), py = fa((float)y, cam.bias[1]);

    for (unsigned c0 = beg; c0 < end; c0 += K3_CHUNK) {
        const unsigned cn = min((unsigned)K3_CHUNK, end - c0);
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        if ((unsigned)tid < cn) {
            const unsigned qi = scan_mode ? (c0 + tid) : tile_list[c0 + tid];
            const uint4 q = queue[qi];
            bool take = true;
            if (scan_mode) {
                int bx0 = (int)(q.y & 0xffffu), by0 = (int)(q.y >> 16), bx1 = (int)(q.z & 0xffffu), by1 = (int)(q.z >> 16);
                take = !(bx1 < x0 || bx0 >= x0 + TILE || by1 < y0 || by0 >= y0 + TILE);
            }
            if (take) {
                Setup s;
                if (qi < qsetup_cap) { // K1 stored the finished setup next to the queue entry
                    const float4 *o = qsetup + (size_t)qi * 4;
                    const float4 a = __ldg(o), b = __ldg(o + 1), c = __ldg(o + 2), d = __ldg(o + 3);
                    s.bcnx = a.x, s.bcny = a.y, s.canx = a.z, s.cany = a.w;
                    s.bx = b.x, s.by = b.y, s.cx = b.z, s.cy = b.w;
                    s.w0 = c.x, s.w1 = c.y, s.w2 = c.z, s.z0 = c.w, s.z1 = d.x, s.z2 = d.y;
                } else {
                    float vv[9];
                    face_world_verts(SRC, verts, (long long)q.x, vv);
                    setup_face(vv, cam, 0u, s); // same ops as K1 => same bits
                }
                const unsigned slot = atomicAdd(&s_cnt, 1u);
                S.f[0][slot] = s.bcnx, S.f[1][slot] = s.bcny, S.f[2][slot] = s.canx, S.f[3][slot] = s.cany;
                S.f[4][slot] = s.bx, S.f[5][slot] = s.by, S.f[6][slot] = s.cx, S.f[7][slot] = s.cy;
                S.f[8][slot] = s.w0, S.f[9][slot] = s.w1, S.f[10][slot] = s.w2;
                S.f[11][slot] = s.z0, S.f[12][slot] = s.z1, S.f[13][slot] = s.z2;
                S.bot[slot] = (int)q.y, S.top[slot] = (int)q.z;
                S.id[slot] = base + q.x + 1u;
            }
        }
        __syncthreads();
        const unsigned m = s_cnt;
        if (m && !loaded) { // first touch of this tile's keys
            orig = inb ? *dst : LLONG_MIN;
            best = orig;
            loaded = true;
        }
        if (inb) {
            for (unsigned j = 0; j < m; j++) {
                const int bot = S.bot[j], top = S.top[j];
                if (x < (bot & 0xffff) || x > (top & 0xffff) || y < (int)((unsigned)bot >> 16) || y > (int)((unsigned)top >> 16))
                    continue;
                Setup s;
                s.bcnx = S.f[0][j], s.bcny = S.f[1][j], s.canx = S.f[2][j], s.cany = S.f[3][j];
                s.bx = S.f[4][j], s.by = S.f[5][j], s.cx = S.f[6][j], s.cy = S.f[7][j];
                s.w0 = S.f[8][j], s.w1 = S.f[9][j], s.w2 = S.f[10][j];
                PW w = pix_products(s, px, py);
                if (pix_fast_reject(w)) continue;
                float q0, q1, q2;
                if (!pix_finish(s, w, q0, q1, q2)) continue;
                s.z0 = S.f[11][j], s.z1 = S.f[12][j], s.z2 = S.f[13][j];
                long long key = pack_key(pix_depth(s, q0, q1, q2), S.id[j]);
                best = key < best ? key : best;
            }
        }
    }
    if (inb && loaded && best < orig) {
        *dst = best;
        blkflags[((long long)x * cam.H + y) >> FLAG_SHIFT] = flagval;
    }
}

// counters: [0] queue count [1] list entries [2] overflow; bar = counters + 8 (arrivals, generation)
__global__ void __launch_bounds__(TILE_PIX)
k_large_path(const float *__restrict__ verts, const __grid_constant__ Cam cam, unsigned base,
             long long *__restrict__ keys, const uint4 *__restrict__ queue, unsigned *__restrict__ counters,
             unsigned *__restrict__ next_counters, unsigned *__restrict__ bar, unsigned queue_cap,
             unsigned *__restrict__ tile_count, unsigned *__restrict__ tile_offs, unsigned *__restrict__ tile_cursor,
             unsigned *__restrict__ tile_list, unsigned list_cap, int tiles_y, int ntiles, unsigned scan_max,
             const __grid_constant__ Src SRC, unsigned char *__restrict__ blkflags, const float4 *__restrict__ qsetup,
             unsigned qsetup_cap, unsigned char flagval) {
    (void)next_counters;
    const unsigned nq = min(counters[0], queue_cap);
    if (nq == 0) return; // nothing queued: the tile path is idle
    __shared__ SetupSoA S;
    __shared__ unsigned s_cnt, s_total;
    __shared__ unsigned s_warp[32];
    bool scan_mode = nq <= scan_max; // small queue: every tile tests the queued bboxes itself
    if (!scan_mode) {
        const int lane = threadIdx.x & 31;
        const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
        // K2a: count overlapped tiles, one warp per queued triangle
        for (unsigned i = warp; i < nq; i += nwarps) {
            int tx0, ty0, tx1, ty1;
            tile_range(queue[i], tx0, ty0, tx1, ty1);
            const int th = ty1 - ty0 + 1, nt = (tx1 - tx0 + 1) * th;
            for (int k = lane; k < nt; k += 32) atomicAdd(&tile_count[(tx0 + k / th) * tiles_y + (ty0 + k % th)], 1u);
        }
        grid_barrier(bar);
        // K2b: exclusive prefix sum over the per-tile counts (CTA 0, warp shuffles)
        if (blockIdx.x == 0) {
            unsigned carry = 0;
            const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
            for (int b0 = 0; b0 < ntiles; b0 += blockDim.x) {
                const int i = b0 + threadIdx.x;
                const unsigned c = (i < ntiles) ? __ldcg(&tile_count[i]) : 0u;
                if (i < ntiles) tile_count[i] = 0u; // leave the histogram clean for the next call
                unsigned incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) s_warp[wid] = incl;
                __syncthreads();
                if (wid == 0) {
                    unsigned v = (lane < nw) ? s_warp[lane] : 0u, iv = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        unsigned t = __shfl_up_sync(0xffffffffu, iv, d);
                        if (lane >= d) iv += t;
                    }
                    s_warp[lane] = iv - v; // exclusive warp offsets
                    if (lane == 31) s_total = iv;
                }
                __syncthreads();
                const unsigned excl = carry + s_warp[wid] + incl - c;
                if (i < ntiles) tile_offs[i] = excl, tile_cursor[i] = excl;
                carry += s_total;
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                tile_offs[ntiles] = carry;
                counters[1] = carry;
                counters[2] = (carry > list_cap) ? 1u : 0u; // lists would overflow: fall back to bbox scanning
            }
        }
        grid_barrier(bar);
        scan_mode = __ldcg(&counters[2]) != 0;
        if (!scan_mode) {
            // K2c: scatter queue indices into the per-tile lists
            for (unsigned i = warp; i < nq; i += nwarps) {
                int tx0, ty0, tx1, ty1;
                tile_range(queue[i], tx0, ty0, tx1, ty1);
                const int th = ty1 - ty0 + 1, nt = (tx1 - tx0 + 1) * th;
                for (int k = lane; k < nt; k += 32) {
                    unsigned pos = atomicAdd(&tile_cursor[(tx0 + k / th) * tiles_y + (ty0 + k % th)], 1u);
                    tile_list[pos] = i;
                }
            }
        }
        grid_barrier(bar);
    }
    // K3: tiles round-robin over the persistent CTAs
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        raster_tile(tile, scan_mode, nq, SRC, verts, cam, base, keys, queue, tile_offs, tile_list, tiles_y, S, s_cnt, blkflags,
                    qsetup, qsetup_cap, flagval);
}
