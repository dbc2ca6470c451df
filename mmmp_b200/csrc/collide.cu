// Kernel (c): configuration and discretised-edge collision checks for spatial
// (Panda / KUKA) worlds.  One configuration per thread: FK of the serial chain
// with link transforms staged in shared memory; the thread parks its 3x4 frame
// poses and the world-space centres of its link bounding spheres in a private
// shared-memory column (conflict-free, lane-contiguous).
// Every collision link carries a sphere decomposition fitted offline plus a
// bounding sphere: obstacle, self and robot-robot tests first cull on the
// bounding spheres and only then descend to the sphere level, where one link's
// spheres are mapped into the OTHER link's frame so the inner loop reads its
// constants as shared-memory broadcasts.  Link pairs whose relative pose depends
// on two joints only may be decided by an exact offline bit table instead.
//
// Edge checks (OR over the interpolation steps) do not evaluate every step: an
// evaluation also yields the clearance of the sphere model at that step, and the
// lever arms RobotDev::reach bound how far any sphere centre can travel per unit
// of the interpolation parameter, so one evaluation proves a whole interval of
// steps collision free.  The verdict is the same OR over all steps (the skipped
// ones are proven free with a 0.1 mm slack that absorbs fp32 rounding); the
// exhaustive kernel is kept as the cross-check (MMMP_CHECK_EXHAUSTIVE).
//
// Replaces (reference): Environment.robot_self_collision / robot_obstacle_collision /
// robot_robot_collision planners/prm/pybullet/env.py:51-84 as composed by
// DecoupledPRM.is_collision_free_sample / _edge decoupled_prm.py:840-863 and
// CoupledPRM.is_collision_free_sample / _edge coupled/coupled_prm.py:143-187.
#include <cstdlib>
#include "common.cuh"
#include "obstacle.cuh"

namespace {

#define MMMP_GAP_SLACK 1.0e-4f   // metres: clearance kept in hand when a step is skipped
#ifndef MMMP_COLLIDE_MINB
#define MMMP_COLLIDE_MINB 5       // resident CTAs per SM the register budget must allow (96 registers; the
                                  // 216 B-per-thread column lets 5 fit in shared memory as well)
#endif
#ifndef MMMP_EDGE_LANES
#define MMMP_EDGE_LANES 4          // lanes per edge in the step-skipping pass (4: 8 edges per warp)
#endif
#ifndef MMMP_EDGE_DEFER_ROUNDS
#define MMMP_EDGE_DEFER_ROUNDS 2    // rounds after which an edge may still be handed to the second pass
#endif

#ifdef MMMP_EDGE_DIAG
// tuning aid (variant builds only, tools/edge_diag.py): which clearance term limited an evaluation
// that proved less than `g_diag_steps` steps.  [0..7] obstacle term of joint j, [16+pi] self pair
// pi, [32+n] evaluations that proved n steps per side (n capped at 15), [63] evaluations.
__device__ unsigned long long g_diag[64];
__device__ float g_diag_spt = 50.f;
#define MMMP_DIAG_MIN(expr, id)            \
    do {                                   \
        const float _t = (expr);           \
        if (_t < tsafe) {                  \
            tsafe = _t;                    \
            why = (id);                    \
        }                                  \
    } while (0)
#else
#define MMMP_DIAG_MIN(expr, id) tsafe = fminf(tsafe, (expr))
#endif

struct SpatialSmem {
    int n_r;
    int n_obs;
    int n_planes;                      // obstacles [0, n_planes) are planes
    int slots;                         // frame slots per thread (stored frames of all robots)
    int links;                         // link slots per thread (stored bounding-sphere centres)
    int mode;                          // RobotDev::frame_store / link_store row in use
    int any_tables;
    int col[MMMP_MAX_ROBOTS];          // first column of local robot lr in a q row
    int frame_slot[MMMP_MAX_ROBOTS];   // first frame slot of local robot lr
    int link_slot[MMMP_MAX_ROBOTS];    // first link slot of local robot lr
};

struct SpatialArgs {
    const WorldDev* W;
    int n_r;
    int robots[MMMP_MAX_ROBOTS];
    int col[MMMP_MAX_ROBOTS];
    uint32_t flags;
    int slots;  // frame slots per thread
    int links;  // link slots per thread
    int mode;   // 0: only what the self pairs read back; 1: every link (robot-robot tests)
};

// per-thread shared-memory column, element e of this thread at fr[e * stride]:
//   [0, 9*links)            per stored link: world centre of its bounding sphere (3), world end points of
//                           its bounding capsule (3 + 3)
//   [9*links, +slots)       mode 1: travel bound of each stored frame per unit t
// (RobotDev::link_store / frame_store map links and frames to slots.)  Round 1 parked the 3x4 POSES
// of the frames instead (360 B per thread for a Panda: 53 KB per CTA, 4 CTAs per SM).  Poses are only
// needed when a link pair survives both its bounding spheres and its capsules -- 2 % of the free
// configurations -- and are then recomputed by a second sweep over the joints (frames_of); what is
// read back for every pair is 9 floats per link: 216 B per thread, 37 KB per CTA, so 5 CTAs fit
// beside 96 registers per thread.
#define MMMP_COLUMN_FLOATS(slots, links, mode) (9 * (links) + ((mode) ? (slots) : 0))

// --- q loaders -------------------------------------------------------------
struct QConfig {
    const float* row;
    __device__ __forceinline__ float get(int lr, int c) const { return __ldg(row + c); }
    __device__ __forceinline__ float delta(int lr, int c) const { return 0.f; }
};
struct QEdge {
    const float* a;
    const float* b;
    float t;
    __device__ __forceinline__ float get(int lr, int c) const {
        const float qa = __ldg(a + c), qb = __ldg(b + c);
        return fmaf(t, qb - qa, qa);
    }
    __device__ __forceinline__ float delta(int lr, int c) const { return __ldg(b + c) - __ldg(a + c); }
};
struct QPair {
    const float* ra;
    const float* rb;
    __device__ __forceinline__ float get(int lr, int c) const { return __ldg((lr == 0 ? ra : rb) + c); }
    __device__ __forceinline__ float delta(int lr, int c) const { return 0.f; }
};

__device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// carve dynamic shared memory and stage the world tables (all threads of the CTA)
__device__ __forceinline__ void stage_world(const SpatialArgs& A, unsigned char* smem, SpatialSmem*& hdr,
                                            RobotDev*& robs, ObstacleDev*& obs, float*& frames) {
    hdr = reinterpret_cast<SpatialSmem*>(smem);
    size_t off = align16(sizeof(SpatialSmem));
    robs = reinterpret_cast<RobotDev*>(smem + off);
    off += align16(sizeof(RobotDev) * A.n_r);
    obs = reinterpret_cast<ObstacleDev*>(smem + off);
    off += align16(sizeof(ObstacleDev) * A.W->n_obstacles);
    frames = reinterpret_cast<float*>(smem + off);
    const int n_obs = A.W->n_obstacles;
    if (threadIdx.x == 0) {
        hdr->n_r = A.n_r;
        hdr->n_obs = n_obs;
        hdr->n_planes = A.W->n_planes;
        hdr->slots = A.slots;
        hdr->links = A.links;
        hdr->mode = A.mode;
        int slot = 0, lslot = 0, tables = 0;
        for (int lr = 0; lr < A.n_r; ++lr) {
            const RobotDev& R = A.W->robots[A.robots[lr]];
            hdr->col[lr] = A.col[lr];
            hdr->frame_slot[lr] = slot;
            hdr->link_slot[lr] = lslot;
            slot += R.n_frame_store[A.mode];
            lslot += R.n_link_store[A.mode];
            tables += R.n_tables;
        }
        hdr->any_tables = tables;
    }
    for (int lr = 0; lr < A.n_r; ++lr) {
        const int* src = reinterpret_cast<const int*>(&A.W->robots[A.robots[lr]]);
        int* dst = reinterpret_cast<int*>(&robs[lr]);
        for (int i = threadIdx.x; i < (int)(sizeof(RobotDev) / 4); i += blockDim.x) dst[i] = src[i];
    }
    {
        const int* src = reinterpret_cast<const int*>(A.W->obstacles);
        int* dst = reinterpret_cast<int*>(obs);
        for (int i = threadIdx.x; i < (int)(sizeof(ObstacleDev) / 4) * n_obs; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
}

// Poses of frames fa and fb of robot lr (frame f = after joint f - 1; frame 0 is the identity because
// base-link spheres are stored in world coordinates), recomputed with the SAME operations as the
// sweep of spatial_eval, hence bit-identical to what that sweep held.
template <class QL>
__device__ __noinline__ void frames_of(const RobotDev& R, const QL& ql, int lr, int col, int fa, int fb, Affine& TA,
                                       Affine& TB) {
#pragma unroll
    for (int i = 0; i < 12; ++i) TA.m[i] = TB.m[i] = (i == 0 || i == 5 || i == 10) ? 1.f : 0.f;
    Affine T;
    affine_load(T, R.base);
    const int last = fa > fb ? fa : fb;
#pragma unroll 1
    for (int j = 0; j < last; ++j) {
        const float qj = ql.get(lr, col + j);
        float s, c;
#ifdef MMMP_COLLIDE_ACCURATE_SINCOS
        sincosf(qj, &s, &c);
#else
        __sincosf(qj, &s, &c);
#endif
        if (R.origin_rx[j]) affine_step_rx(T, R.origin[j], s, c);
        else affine_step(T, R.origin[j], s, c);
        if (j + 1 == fa) TA = T;
        if (j + 1 == fb) TB = T;
    }
}

// spheres of link lb (robot RB, pose TB) against spheres of link la (robot RA, pose TA); the
// caller has already found the two bounding spheres overlapping (GAP: closer than `want`).
// With GAP the smallest clearance met on the way is folded into `gap`: a lower bound of the
// true one, exact wherever it is below `want` (a sphere of B further than `want` from A's
// bounding sphere contributes that distance instead of its exact clearance).
template <bool GAP>
__device__ __forceinline__ bool link_pair_hit(const RobotDev& RA, int la, const Affine& TA, const RobotDev& RB, int lb,
                                              const Affine& TB, float want, float& gap) {
    const float4 ba = RA.link_bound[la];
    // M = TA^-1 * TB  (rigid): R = RA^T RB, t = RA^T (tB - tA)
    Affine M;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            M.m[4 * i + j] = fmaf(TA.m[i], TB.m[j], fmaf(TA.m[4 + i], TB.m[4 + j], TA.m[8 + i] * TB.m[8 + j]));
        const float tx = TB.m[3] - TA.m[3], ty = TB.m[7] - TA.m[7], tz = TB.m[11] - TA.m[11];
        M.m[4 * i + 3] = fmaf(TA.m[i], tx, fmaf(TA.m[4 + i], ty, TA.m[8 + i] * tz));
    }
    bool hit = false;
    const int a0 = RA.link_sphere_begin[la], a1 = RA.link_sphere_begin[la + 1];
    for (int sb = RB.link_sphere_begin[lb]; sb < RB.link_sphere_begin[lb + 1]; ++sb) {
        const float4 S = RB.sphere[sb];
        const float3 c = affine_point(M, S.x, S.y, S.z);
        {
            const float dx = c.x - ba.x, dy = c.y - ba.y, dz = c.z - ba.z;
            const float rr = S.w + ba.w;
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            if (GAP) {  // clearance already as large as the caller can use: no need for the exact one
                const float gs = fast_sqrt(d2) - rr;
                if (gs > want) {
                    gap = fminf(gap, gs);
                    continue;
                }
            } else if (d2 > rr * rr) {
                continue;
            }
        }
        float m2 = 3.0e38f;
        for (int sa = a0; sa < a1; ++sa) {
            const float4 P = RA.sphere[sa];
            const float dx = c.x - P.x, dy = c.y - P.y, dz = c.z - P.z;
            const float rr = S.w + P.w;
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            hit |= d2 <= rr * rr;
            if (GAP) m2 = fminf(m2, fast_sqrt(d2) - rr);
        }
        if (hit) return true;
        if (GAP) gap = fminf(gap, m2);
    }
    return false;
}

// Squared distance between two segments p1-q1 and p2-q2 (the world-space axes of two bounding
// capsules).  The closest pair of points is found in fp32 (Ericson's segment / segment clamp
// sequence); whatever rounding does to the parameters, the value is a distance between two points
// that DO lie on the segments, so it can exceed the true minimum only by the (second order) effect
// of the parameter error -- callers keep a 0.1 mm margin.
__device__ __forceinline__ float segment_d2(float3 p1, float3 q1, float3 p2, float3 q2) {
    const float d1x = q1.x - p1.x, d1y = q1.y - p1.y, d1z = q1.z - p1.z;
    const float d2x = q2.x - p2.x, d2y = q2.y - p2.y, d2z = q2.z - p2.z;
    const float rx = p1.x - p2.x, ry = p1.y - p2.y, rz = p1.z - p2.z;
    const float a = fmaf(d1x, d1x, fmaf(d1y, d1y, d1z * d1z)), e = fmaf(d2x, d2x, fmaf(d2y, d2y, d2z * d2z));
    const float f = fmaf(d2x, rx, fmaf(d2y, ry, d2z * rz)), c = fmaf(d1x, rx, fmaf(d1y, ry, d1z * rz));
    const float b = fmaf(d1x, d2x, fmaf(d1y, d2y, d1z * d2z));
    const float denom = fmaf(a, e, -b * b);
    const float inv_a = __fdividef(1.f, fmaxf(a, 1e-20f)), inv_e = __fdividef(1.f, fmaxf(e, 1e-20f));
    float sp = denom > 1e-7f * a * e ? __saturatef(__fdividef(fmaf(b, f, -c * e), denom)) : 0.f;
    const float tp = __saturatef(fmaf(b, sp, f) * inv_e);   // closest point of segment 2 to P(sp) ...
    sp = __saturatef(fmaf(b, tp, -c) * inv_a);              // ... and of segment 1 to Q(tp): covers Ericson's
                                                            // clamp cases and degenerate (point) segments
    const float ex = fmaf(sp, d1x, rx) - tp * d2x, ey = fmaf(sp, d1y, ry) - tp * d2y, ez = fmaf(sp, d1z, rz) - tp * d2z;
    return fmaf(ex, ex, fmaf(ey, ey, ez * ez));
}

#define MMMP_CAPSULE_MARGIN 1.0e-4f   // metres kept in hand when a capsule test skips a descent

// world end points of link slot ls's capsule from a thread's column (see MMMP_COLUMN_FLOATS)
__device__ __forceinline__ void load_capsule(const float* cl, int stride, int ls, float3& a, float3& b) {
    const float* p = cl + (size_t)(9 * ls + 3) * stride;
    a = make_float3(p[0], p[stride], p[2 * stride]);
    b = make_float3(p[3 * stride], p[4 * stride], p[5 * stride]);
}

// Link pair whose bounding spheres overlap (GAP: are closer than `want`): capsule against capsule
// first (world end points parked by the sweep); only if that cannot separate them either (GAP:
// cannot prove `want`) are the two frame poses recomputed and the spheres of the two links tested.
// Returns true on a collision; GAP: g = clearance proven for the pair.
template <class QL, bool GAP>
__device__ __forceinline__ bool link_pair_refine(const RobotDev& RA, int la, int lra, int cola, const float* cla, int lsa,
                                                 const RobotDev& RB, int lb, int lrb, int colb, const float* clb, int lsb,
                                                 const QL& ql, int stride, bool capsules, float want, float& g) {
    if (capsules && RA.cap_a[la].w >= 0.f && RB.cap_a[lb].w >= 0.f) {
        float3 p1, q1, p2, q2;
        load_capsule(cla, stride, lsa, p1, q1);
        load_capsule(clb, stride, lsb, p2, q2);
        const float d2 = segment_d2(p1, q1, p2, q2);
        const float rr = RA.cap_a[la].w + RB.cap_a[lb].w;
        if (GAP) {
            const float gc = fast_sqrt(d2) - rr - MMMP_CAPSULE_MARGIN;
            if (gc > want) {
                g = fmaxf(g, gc);
                return false;
            }
        } else {
            const float lim = rr + MMMP_CAPSULE_MARGIN;
            if (d2 > lim * lim) return false;
        }
    }
    Affine TA, TB, unused;
    const int fa = RA.link_frame[la], fb = RB.link_frame[lb];
    if (&RA == &RB) {
        frames_of(RA, ql, lra, cola, fa, fb, TA, TB);
    } else {
        frames_of(RA, ql, lra, cola, fa, 0, TA, unused);
        frames_of(RB, ql, lrb, colb, fb, 0, TB, unused);
    }
    g = __int_as_float(0x7f800000);
    return link_pair_hit<GAP>(RA, la, TA, RB, lb, TB, want, g);
}

// bit of the pair table P at joint values (qu, qw); no branch, so that look-ups of several
// steps can be in flight together
__device__ __forceinline__ unsigned table_bit(const WorldDev* __restrict__ Wg, const PairTableDev& P, float qu,
                                              float qw) {
    int iu = (int)floorf((qu - P.lo_u) * P.inv_du);
    int iv = (int)floorf((qw - P.lo_v) * P.inv_dv);
    iu = min(max(iu, 0), P.n_u - 1);
    iv = min(max(iv, 0), P.n_v - 1);
    const int cell = iu * P.n_v + iv;
    return (__ldg(Wg->table_bits + P.word_offset + (cell >> 5)) >> (cell & 31)) & 1u;
}

template <class QL>
__device__ __forceinline__ bool tables_hit(const WorldDev* __restrict__ Wg, const RobotDev& R, int lr, int col,
                                           const QL& ql) {
    unsigned hit = 0;
    for (int t = 0; t < R.n_tables; ++t) {
        const PairTableDev& P = R.table[t];
        hit |= table_bit(Wg, P, ql.get(lr, col + P.joint_u), ql.get(lr, col + P.joint_v));
    }
    return hit != 0;
}

// One configuration.  Returns < 0 when it is in collision.  Otherwise: GAP = false returns 1;
// GAP = true returns the half-width, in units of the interpolation parameter t, of the
// interval around this configuration's t on which the edge ql is proven collision free
// (0 when nothing can be proven).  The collision verdict is computed by the same
// expressions for both values of GAP.  skip_tables: the caller has looked up the pair
// tables itself.  t_want (GAP): the half-width the caller would like; a bounding-sphere
// clearance that already proves it is taken as is, a smaller one is refined at the sphere
// level (tighter, dearer).
template <class QL, bool GAP>
__device__ float spatial_eval(const WorldDev* __restrict__ Wg, const SpatialSmem* hdr, const RobotDev* robs,
                              const ObstacleDev* obs, uint32_t flags, float* fr, int stride, const QL& ql,
                              bool skip_tables, float t_want) {
    const float HIT = -1.f;
    const float INF = __int_as_float(0x7f800000);
    const int n_r = hdr->n_r;
    const int mode = hdr->mode;
    const int n_obs = (flags & MMMP_CHECK_OBSTACLES) ? hdr->n_obs : 0;
    const int n_planes = hdr->n_planes;
    const bool capsules = !(flags & MMMP_CHECK_NO_CAPSULES);
    float* vel = fr + (size_t)hdr->links * 9 * stride;
    float tsafe = INF;
#ifdef MMMP_EDGE_DIAG
    int why = 30;
#endif
    for (int lr = 0; lr < n_r; ++lr) {
        const RobotDev& R = robs[lr];
        const int col = hdr->col[lr];
        if (n_obs > 0 && R.base_hits_obstacles) return HIT;
        Affine T;
        affine_load(T, R.base);
        const int slot0 = hdr->frame_slot[lr];
        float* cl = fr + (size_t)hdr->link_slot[lr] * 9 * stride;
        for (int l = R.frame_link_begin[0]; l < R.frame_link_begin[1]; ++l) {  // base links: world coordinates
            const int ls = R.link_store[mode][l];
            if (ls < 0) continue;
            const float4 B = R.link_bound[l], ca = R.cap_a[l], cb = R.cap_b[l];
            float* p = cl + (size_t)9 * ls * stride;
            p[0] = B.x, p[stride] = B.y, p[2 * stride] = B.z;
            p[3 * stride] = ca.x, p[4 * stride] = ca.y, p[5 * stride] = ca.z;
            p[6 * stride] = cb.x, p[7 * stride] = cb.y, p[8 * stride] = cb.z;
        }
        bool hit = false;
        float adq[MMMP_MAX_JOINTS];
#pragma unroll
        for (int j = 0; j < MMMP_MAX_JOINTS; ++j) adq[j] = 0.f;
#pragma unroll 1
        for (int j = 0; j < R.n_joints; ++j) {
            const float qj = ql.get(lr, col + j);
            // MUFU sine/cosine: |error| < 1e-6 on the joint ranges (|q| < 3.8), i.e. a few micrometres
            // at the tool -- three orders below the sphere-fit margin and inside MMMP_GAP_SLACK;
            // the accurate sincosf was a quarter of the instructions of a verdict-only evaluation
            float s, c;
#ifdef MMMP_COLLIDE_ACCURATE_SINCOS
            sincosf(qj, &s, &c);
#else
            __sincosf(qj, &s, &c);
#endif
            if (R.origin_rx[j]) affine_step_rx(T, R.origin[j], s, c);
            else affine_step(T, R.origin[j], s, c);
            const int fs = R.frame_store[mode][j + 1];
            float inv_v = 0.f, want = 0.f;
            if (GAP) {
                const float dj = fabsf(ql.delta(lr, col + j));
                float v = 0.f;
#pragma unroll
                for (int m = 0; m < MMMP_MAX_JOINTS; ++m) {
                    if (m == j) adq[m] = dj;
                    if (m <= j) v = fmaf(adq[m], R.reach[m][j + 1], v);
                }
                if (mode && fs >= 0) vel[(slot0 + fs) * stride] = v;
                inv_v = __fdividef(1.f, fmaxf(v, 1e-20f));
                want = fmaf(v, t_want, MMMP_GAP_SLACK);
            }
            float gmin = INF;
            // links hanging off frame j+1: bounding sphere first, spheres on demand
            for (int l = R.frame_link_begin[j + 1]; l < R.frame_link_begin[j + 2]; ++l) {
                const float4 B = R.link_bound[l];
                if (B.w < 0.f) continue;
                const float3 wb = affine_point(T, B.x, B.y, B.z);
                const int ls = R.link_store[mode][l];
                if (ls >= 0) {  // what the link-pair tests read back: centre and capsule end points, world space
                    const float4 ca = R.cap_a[l], cb = R.cap_b[l];
                    const float3 wa = affine_point(T, ca.x, ca.y, ca.z), wc = affine_point(T, cb.x, cb.y, cb.z);
                    float* p = cl + (size_t)9 * ls * stride;
                    p[0] = wb.x, p[stride] = wb.y, p[2 * stride] = wb.z;
                    p[3 * stride] = wa.x, p[4 * stride] = wa.y, p[5 * stride] = wa.z;
                    p[6 * stride] = wc.x, p[7 * stride] = wc.y, p[8 * stride] = wc.z;
                }
                for (int o = 0; o < n_obs; ++o) {
                    float g = INF;
                    bool bound_hit;
                    if (o < n_planes) {  // the same expressions as sphere_vs_obstacle's plane case, no dispatch
                        const float4 P = *reinterpret_cast<const float4*>(obs[o].p);
                        const float sd = fmaf(P.x, wb.x, fmaf(P.y, wb.y, P.z * wb.z)) - P.w;
                        g = sd - B.w;
                        bound_hit = sd <= B.w;
                    } else {
                        bound_hit = sphere_vs_obstacle<GAP>(wb.x, wb.y, wb.z, B.w, obs[o], g);
                    }
                    if (GAP ? g > want : !bound_hit) {
                        if (GAP) gmin = fminf(gmin, g);
                        continue;
                    }
                    for (int sp = R.link_sphere_begin[l]; sp < R.link_sphere_begin[l + 1]; ++sp) {
                        const float4 S = R.sphere[sp];
                        const float3 w = affine_point(T, S.x, S.y, S.z);
                        hit |= sphere_vs_obstacle<GAP>(w.x, w.y, w.z, S.w, obs[o], gmin);
                    }
                    if (hit) return HIT;
                }
            }
            if (GAP && n_obs > 0) MMMP_DIAG_MIN((gmin - MMMP_GAP_SLACK) * inv_v, j);
        }
        if (flags & MMMP_CHECK_SELF) {
            if (!skip_tables && tables_hit(Wg, R, lr, col, ql)) return HIT;
            for (int pi = 0; pi < R.n_self_pairs; ++pi) {
                const uchar2 pr = R.self_pair[pi];
                const int fa = R.link_frame[pr.x], fb = R.link_frame[pr.y];
                const float ra = R.link_bound[pr.x].w, rb = R.link_bound[pr.y].w;
                const int sa = R.link_store[mode][pr.x], sb = R.link_store[mode][pr.y];
                const float dx = cl[(9 * sa + 0) * stride] - cl[(9 * sb + 0) * stride];
                const float dy = cl[(9 * sa + 1) * stride] - cl[(9 * sb + 1) * stride];
                const float dz = cl[(9 * sa + 2) * stride] - cl[(9 * sb + 2) * stride];
                const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                const float rr = ra + rb;
                float g = INF, v = 0.f, want = 0.f;
                if (GAP) {
                    // relative travel of link b in link a's frame: joints fa .. fb-1 only
#pragma unroll
                    for (int m = 0; m < MMMP_MAX_JOINTS; ++m)
                        if (m >= fa && m < fb) v = fmaf(adq[m], R.reach[m][fb], v);
                    want = fmaf(v, t_want, MMMP_GAP_SLACK);
                    g = fast_sqrt(d2) - rr;
                }
                if (GAP ? g <= want : d2 <= rr * rr) {
                    if (link_pair_refine<QL, GAP>(R, pr.x, lr, col, cl, sa, R, pr.y, lr, col, cl, sb, ql, stride, capsules,
                                                  want, g))
                        return HIT;
                }
                if (GAP) MMMP_DIAG_MIN(__fdividef(g - MMMP_GAP_SLACK, fmaxf(v, 1e-20f)), 16 + pi);
            }
        }
    }
    if ((flags & MMMP_CHECK_ROBOTS) && n_r > 1) {
        for (int l1 = 0; l1 < n_r; ++l1) {
            const RobotDev& R1 = robs[l1];
            const float* c1 = fr + (size_t)hdr->link_slot[l1] * 9 * stride;
            for (int l2 = l1 + 1; l2 < n_r; ++l2) {
                const RobotDev& R2 = robs[l2];
                const float* c2 = fr + (size_t)hdr->link_slot[l2] * 9 * stride;
                for (int la = 0; la < R1.n_links; ++la) {
                    const float ra = R1.link_bound[la].w;
                    if (ra < 0.f) continue;
                    const int fa = R1.link_frame[la];
                    const int sa = R1.link_store[mode][la];
                    const float ax = c1[(9 * sa + 0) * stride], ay = c1[(9 * sa + 1) * stride],
                                az = c1[(9 * sa + 2) * stride];
                    for (int lb = 0; lb < R2.n_links; ++lb) {
                        const float rb = R2.link_bound[lb].w;
                        if (rb < 0.f) continue;
                        const int fb = R2.link_frame[lb];
                        if (fa == 0 && fb == 0) continue;  // static pair
                        const int sb = R2.link_store[mode][lb];
                        const float dx = ax - c2[(9 * sb + 0) * stride], dy = ay - c2[(9 * sb + 1) * stride],
                                    dz = az - c2[(9 * sb + 2) * stride];
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        const float rr = ra + rb;
                        float g = INF, v = 0.f, want = 0.f;
                        if (GAP) {
                            const float va = fa > 0 ? vel[(hdr->frame_slot[l1] + R1.frame_store[mode][fa]) * stride] : 0.f;
                            const float vb = fb > 0 ? vel[(hdr->frame_slot[l2] + R2.frame_store[mode][fb]) * stride] : 0.f;
                            v = va + vb;
                            want = fmaf(v, t_want, MMMP_GAP_SLACK);
                            g = fast_sqrt(d2) - rr;
                        }
                        if (GAP ? g <= want : d2 <= rr * rr) {
                            if (link_pair_refine<QL, GAP>(R1, la, l1, hdr->col[l1], c1, sa, R2, lb, l2, hdr->col[l2], c2, sb, ql,
                                                          stride, capsules, want, g))
                                return HIT;
                        }
                        if (GAP) MMMP_DIAG_MIN(__fdividef(g - MMMP_GAP_SLACK, fmaxf(v, 1e-20f)), 31);
                    }
                }
            }
        }
    }
    if (!GAP) return 1.f;
#ifdef MMMP_EDGE_DIAG
    {
        const int n = (int)fminf(fmaxf(tsafe, 0.f) * g_diag_spt, 15.f);
        atomicAdd(g_diag + 32 + n, 1ull);
        atomicAdd(g_diag + 63, 1ull);
        if (n < 1) atomicAdd(g_diag + why, 1ull);
    }
#endif
    return fmaxf(tsafe, 0.f);
}

template <class QL>
__device__ __forceinline__ bool spatial_in_collision(const WorldDev* __restrict__ Wg, const SpatialSmem* hdr,
                                                     const RobotDev* robs, const ObstacleDev* obs, uint32_t flags,
                                                     float* fr, int stride, const QL& ql) {
    return spatial_eval<QL, false>(Wg, hdr, robs, obs, flags, fr, stride, ql, false, 0.f) < 0.f;
}

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(128, MMMP_COLLIDE_MINB)
collide_configs_kernel(SpatialArgs A, const float* __restrict__ q, int ld, int64_t N, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    SpatialSmem* hdr;
    RobotDev* robs;
    ObstacleDev* obs;
    float* frames;
    stage_world(A, smem, hdr, robs, obs, frames);
    float* cs = frames + threadIdx.x;
    const int stride = blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        QConfig ql{q + i * ld};
        out[i] = spatial_in_collision(A.W, hdr, robs, obs, A.flags, cs, stride, ql) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128, MMMP_COLLIDE_MINB)
collide_pairs_kernel(SpatialArgs A, const float* __restrict__ qa, const float* __restrict__ qb, int ld,
                     int64_t N, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    SpatialSmem* hdr;
    RobotDev* robs;
    ObstacleDev* obs;
    float* frames;
    stage_world(A, smem, hdr, robs, obs, frames);
    float* cs = frames + threadIdx.x;
    const int stride = blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        QPair ql{qa + i * ld, qb + i * ld};
        out[i] = spatial_in_collision(A.W, hdr, robs, obs, A.flags, cs, stride, ql) ? 1 : 0;
    }
}

// Exhaustive edge check: one warp per edge; lane l takes steps l*rounds + r in round r so that
// every round spreads over the whole edge, warp-ballot early exit.  Verdict = OR over all
// steps = reference semantics (decoupled_prm.py:855-863).
__global__ void __launch_bounds__(128, MMMP_COLLIDE_MINB)
collide_edges_kernel(SpatialArgs A, const float* __restrict__ qa_base, const float* __restrict__ qb_base, int ld,
                     const int32_t* __restrict__ edges, int64_t E, int n_steps, double step,
                     uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    SpatialSmem* hdr;
    RobotDev* robs;
    ObstacleDev* obs;
    float* frames;
    stage_world(A, smem, hdr, robs, obs, frames);
    float* cs = frames + threadIdx.x;
    const int stride = blockDim.x;
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int rounds = (n_steps + 31) >> 5;
    for (int64_t e = (int64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); e < E;
         e += (int64_t)gridDim.x * warps_per_cta) {
        const float *pa, *pb;
        if (edges) {
            const int ia = __ldg(edges + 2 * e), ib = __ldg(edges + 2 * e + 1);
            if (ia < 0 || ib < 0) {  // empty neighbour slot: reported as blocked
                if (lane == 0) out[e] = 1;
                continue;
            }
            pa = qa_base + (int64_t)ia * ld;
            pb = qa_base + (int64_t)ib * ld;
        } else {
            pa = qa_base + e * ld;
            pb = qb_base + e * ld;
        }
        bool any = false;
        for (int r = 0; r < rounds && !any; ++r) {
            const int i = lane * rounds + r;
            bool hit = false;
            if (i < n_steps) {
                QEdge ql{pa, pb, (float)((double)i * step)};
                hit = spatial_in_collision(A.W, hdr, robs, obs, A.flags, cs, stride, ql);
            }
            any = __any_sync(0xffffffffu, hit);
        }
        if (lane == 0) out[e] = any ? 1 : 0;
    }
}

// position of the r-th (0-based) set bit of m, r < popcount(m): six popcount halvings (the __fns
// intrinsic this replaces is a software loop: 4-5 % of the edge kernels' samples,
// profiles/r02_edge_full.csv)
__device__ __forceinline__ int nth_set_bit(uint64_t m, int r) {
    uint32_t w = (uint32_t)m;
    int pos = 0;
    int c = __popc(w);
    if (r >= c) {
        r -= c;
        w = (uint32_t)(m >> 32);
        pos = 32;
    }
    c = __popc(w & 0xffffu);
    if (r >= c) {
        r -= c;
        w >>= 16;
        pos += 16;
    }
    c = __popc(w & 0xffu);
    if (r >= c) {
        r -= c;
        w >>= 8;
        pos += 8;
    }
    c = __popc(w & 0xfu);
    if (r >= c) {
        r -= c;
        w >>= 4;
        pos += 4;
    }
    c = __popc(w & 0x3u);
    if (r >= c) {
        r -= c;
        w >>= 2;
        pos += 2;
    }
    return pos + (r >= (int)(w & 1u) ? 1 : 0);
}

// Edge check that evaluates a step only while nothing proves it free (n_steps <= 64).
// A warp is cut into 32/L groups of L lanes; a group owns one edge at a time and keeps the
// set of still undecided steps as a bit mask.  Each round every lane of the group evaluates
// one undecided step (spread evenly over the mask); an evaluation that collides decides the
// edge, one that does not clears the steps its clearance covers.  A group whose mask is
// empty writes the verdict and takes the next edge of the warp's contiguous chunk, so slow
// edges never stall the other groups.
template <int L>
__global__ void __launch_bounds__(128, MMMP_COLLIDE_MINB)
collide_edges_adaptive_kernel(SpatialArgs A, const float* __restrict__ qa_base, const float* __restrict__ qb_base,
                              int ld, const int32_t* __restrict__ edges, int64_t E,
                              const int32_t* __restrict__ E_dev, int n_steps, double step,
                              float want_scale, int hard_after, int64_t* __restrict__ hard_edge,
                              unsigned long long* __restrict__ hard_mask, unsigned long long* __restrict__ hard_count,
                              unsigned long long* __restrict__ stats, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    SpatialSmem* hdr;
    RobotDev* robs;
    ObstacleDev* obs;
    float* frames;
    stage_world(A, smem, hdr, robs, obs, frames);
    float* cs = frames + threadIdx.x;
    const int stride = blockDim.x;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int group = lane / L, gl = lane % L;
    const unsigned gmask = ((L == 32) ? FULL : ((1u << L) - 1u)) << (group * L);
    const unsigned leaders_below = (L == 32) ? 0u : (0x11111111u * (L == 4) + 0x01010101u * (L == 8) +
                                                     0x00010001u * (L == 16)) & ((1u << (group * L)) - 1u);
    const int warps_per_cta = blockDim.x >> 5;
    const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;
    const int64_t warp_id = (int64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    if (E_dev) {  // the list was compacted on the device: only its first *E_dev entries are edges
        const int64_t live = __ldg(E_dev);
        E = live < E ? live : E;
    }
    const int64_t chunk = (E + n_warps - 1) / n_warps;
    int64_t next = warp_id * chunk;
    const int64_t w_end = next + chunk < E ? next + chunk : E;
    const uint64_t all_steps = n_steps >= 64 ? ~0ull : ((1ull << n_steps) - 1ull);
    const float steps_per_t = (float)(0.999 / step);
    const bool use_tables = (A.flags & MMMP_CHECK_SELF) && hdr->any_tables;

    int64_t e = -1;
    uint64_t mask = 0;
    bool collided = false, deferred = false;
    int rounds = 0;
    const float *pa = qa_base, *pb = qa_base;
    unsigned n_evals = 0;
    while (true) {
        // ---- groups without undecided steps: write the verdict, take the next edge
        const bool need = mask == 0;
        if (need && e >= 0 && gl == 0 && !deferred) out[e] = collided ? 1 : 0;
        const unsigned needb = __ballot_sync(FULL, need && gl == 0);
        bool table_hit = false;
        if (need) {
            const int64_t cand = next + __popc(needb & leaders_below);
            e = -1;
            if (cand < w_end) {
                e = cand;
                collided = false;
                deferred = false;
                rounds = 0;
                mask = all_steps;
                if (edges) {
                    const int ia = __ldg(edges + 2 * e), ib = __ldg(edges + 2 * e + 1);
                    if (ia < 0 || ib < 0) {  // empty neighbour slot: reported as blocked
                        collided = true;
                        mask = 0;
                    }
                    pa = qa_base + (int64_t)(ia < 0 ? 0 : ia) * ld;
                    pb = qa_base + (int64_t)(ib < 0 ? 0 : ib) * ld;
                } else {
                    pa = qa_base + e * ld;
                    pb = qb_base + e * ld;
                }
                if (use_tables && mask) {  // pair tables depend on q alone: look every step up
                    unsigned bits = 0;
                    for (int lr = 0; lr < hdr->n_r; ++lr) {
                        const RobotDev& R = robs[lr];
                        for (int tb = 0; tb < R.n_tables; ++tb) {
                            const PairTableDev& P = R.table[tb];
                            const int cu = hdr->col[lr] + P.joint_u, cv = hdr->col[lr] + P.joint_v;
                            const float au = __ldg(pa + cu), av = __ldg(pa + cv);
                            const float du = __ldg(pb + cu) - au, dv = __ldg(pb + cv) - av;
#pragma unroll 4
                            for (int i = gl; i < n_steps; i += L) {  // independent look-ups: several in flight
                                const float t = (float)((double)i * step);
                                bits |= table_bit(A.W, P, fmaf(t, du, au), fmaf(t, dv, av));  // = QEdge::get
                            }
                        }
                    }
                    table_hit = bits != 0;
                }
            }
        }
        next += __popc(needb);
        {
            const unsigned tb = __ballot_sync(FULL, table_hit);
            if (need && (tb & gmask)) {
                collided = true;
                mask = 0;
            }
        }
        if (!__any_sync(FULL, e >= 0)) break;
        // ---- one evaluation per lane
        const int cnt = __popcll(mask);
        const int rank = ((2 * gl + 1) * cnt) / (2 * L);
        const bool act = cnt > 0 && (gl == 0 || rank != ((2 * gl - 1) * cnt) / (2 * L));
        uint64_t cover = 0;
        bool hit = false;
        if (act) {
            ++n_evals;
            const int i = nth_set_bit(mask, rank);
            QEdge ql{pa, pb, (float)((double)i * step)};
            // what this lane can use: its share of the undecided steps, on either side
            const float t_want = (float)(cnt / (2 * L) + 1) * want_scale;
            const float ts = spatial_eval<QEdge, true>(A.W, hdr, robs, obs, A.flags, cs, stride, ql, true, t_want);
            if (ts < 0.f) {
                hit = true;
            } else {
                const int n = (int)fminf(ts * steps_per_t, 64.f);
                const int lo = max(i - n, 0), hi = min(i + n, 63);
                cover = (~0ull >> (63 - hi)) & (~0ull << lo);
            }
        }
#pragma unroll
        for (int off = L >> 1; off > 0; off >>= 1) cover |= __shfl_xor_sync(FULL, cover, off);
        const unsigned hb = __ballot_sync(FULL, hit);
        if (hb & gmask) {
            collided = true;
            mask = 0;
        } else {
            mask &= ~cover;
        }
        ++rounds;
        // An edge on which the first rounds prove little (clearance below the travel per step) will
        // have nearly every step evaluated: hand it, with its undecided steps, to the warp-per-edge
        // pass that follows (coherent lanes, verdict-only evaluation).
        const bool defer = hard_edge && mask != 0 && rounds <= MMMP_EDGE_DEFER_ROUNDS && __popcll(mask) > hard_after;
        const unsigned db = __ballot_sync(FULL, defer && gl == 0);
        if (db) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(hard_count, (unsigned long long)__popc(db));
            base = __shfl_sync(FULL, base, 0);
            if (defer) {
                if (gl == 0) {
                    const unsigned long long at = base + __popc(db & leaders_below);
                    hard_edge[at] = e;
                    hard_mask[at] = mask;
                }
                deferred = true;
                mask = 0;
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) n_evals += __shfl_xor_sync(FULL, n_evals, off);
        if (lane == 0) {
            atomicAdd(stats + 1, (unsigned long long)n_evals);
            const int64_t first = warp_id * chunk;
            if (w_end > first) atomicAdd(stats, (unsigned long long)(w_end - first));
        }
    }
}

// Second pass over the edges the kernel above deferred: a warp takes MMMP_HARD_BATCH edges at
// a time and deals their undecided steps (bit masks) out to its lanes as one list, so a round
// spans one or two edges (coherent lanes) and only the last round of a batch is partly idle.
// Verdict-only evaluation; an edge found colliding drops its remaining steps.  The pair tables
// were already looked up for every step of a deferred edge.
#define MMMP_HARD_BATCH 8
__global__ void __launch_bounds__(128, MMMP_COLLIDE_MINB)
collide_edges_hard_kernel(SpatialArgs A, const float* __restrict__ qa_base, const float* __restrict__ qb_base, int ld,
                          const int32_t* __restrict__ edges, const int64_t* __restrict__ hard_edge,
                          const unsigned long long* __restrict__ hard_mask,
                          const unsigned long long* __restrict__ hard_count, double step,
                          unsigned long long* __restrict__ stats, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    SpatialSmem* hdr;
    RobotDev* robs;
    ObstacleDev* obs;
    float* frames;
    stage_world(A, smem, hdr, robs, obs, frames);
    float* cs = frames + threadIdx.x;
    const int stride = blockDim.x;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int64_t n_hard = (int64_t)*hard_count;
    unsigned n_evals = 0;
    for (int64_t h0 = ((int64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5)) * MMMP_HARD_BATCH; h0 < n_hard;
         h0 += (int64_t)gridDim.x * warps_per_cta * MMMP_HARD_BATCH) {
        // lane b < BATCH holds edge h0 + b: id, mask, end points, running verdict
        int64_t my_e = -1;
        uint64_t my_mask = 0;
        if (lane < MMMP_HARD_BATCH && h0 + lane < n_hard) {
            my_e = hard_edge[h0 + lane];
            my_mask = hard_mask[h0 + lane];
        }
        int incl = __popcll(my_mask);
#pragma unroll
        for (int o = 1; o < MMMP_HARD_BATCH; o <<= 1) {
            const int up = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += up;
        }
        const int total = __shfl_sync(FULL, incl, MMMP_HARD_BATCH - 1);
        unsigned dead = 0;  // bit b: edge b of the batch already found colliding
        for (int g0 = 0; g0 < total; g0 += 32) {
            const int g = g0 + lane;
            int b = 0;  // edge of the batch that owns list item g
#pragma unroll
            for (int bb = 0; bb < MMMP_HARD_BATCH - 1; ++bb) b += (g >= __shfl_sync(FULL, incl, bb)) ? 1 : 0;
            const int before = __shfl_sync(FULL, incl, b) - __popcll(__shfl_sync(FULL, my_mask, b));
            const int64_t e = __shfl_sync(FULL, my_e, b);
            const uint64_t mask = __shfl_sync(FULL, my_mask, b);
            bool hit = false;
            if (g < total && !((dead >> b) & 1u)) {
                ++n_evals;
                const int i = nth_set_bit(mask, g - before);
                const float *pa, *pb;
                if (edges) {
                    pa = qa_base + (int64_t)__ldg(edges + 2 * e) * ld;
                    pb = qa_base + (int64_t)__ldg(edges + 2 * e + 1) * ld;
                } else {
                    pa = qa_base + e * ld;
                    pb = qb_base + e * ld;
                }
                QEdge ql{pa, pb, (float)((double)i * step)};
                hit = spatial_eval<QEdge, false>(A.W, hdr, robs, obs, A.flags, cs, stride, ql, true, 0.f) < 0.f;
            }
#pragma unroll
            for (int bb = 0; bb < MMMP_HARD_BATCH; ++bb)
                if (__any_sync(FULL, hit && b == bb)) dead |= 1u << bb;
        }
        if (my_e >= 0) out[my_e] = (dead >> lane) & 1u;
    }
    if (stats) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) n_evals += __shfl_xor_sync(FULL, n_evals, off);
        if (lane == 0) atomicAdd(stats + 3, (unsigned long long)n_evals);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(stats + 2, (unsigned long long)n_hard);
    }
}

// ------------------------------------------------------------------ host side
int fill_args(const mmmp_world* w, int robot, uint32_t flags, SpatialArgs& A) {
    if (!w || w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    const WorldDev* H = w->spatial_h;
    A.W = w->spatial_d;
    A.flags = flags;
    if (robot >= 0) {
        if (robot >= H->n_robots) return MMMP_ERR_ARG;
        A.n_r = 1;
        A.robots[0] = robot;
        A.col[0] = 0;
        A.mode = 0;
        A.slots = H->robots[robot].n_frame_store[0];
        A.links = H->robots[robot].n_link_store[0];
    } else {
        A.n_r = H->n_robots;
        A.mode = ((flags & MMMP_CHECK_ROBOTS) && H->n_robots > 1) ? 1 : 0;
        A.slots = 0;
        A.links = 0;
        for (int r = 0; r < H->n_robots; ++r) {
            A.robots[r] = r;
            A.col[r] = H->robots[r].dof_offset;
            A.slots += H->robots[r].n_frame_store[A.mode];
            A.links += H->robots[r].n_link_store[A.mode];
        }
    }
    return MMMP_OK;
}

size_t smem_bytes(const mmmp_world* w, const SpatialArgs& A, int block) {
    size_t off = (sizeof(SpatialSmem) + 15) & ~(size_t)15;
    off += (sizeof(RobotDev) * A.n_r + 15) & ~(size_t)15;
    off += (sizeof(ObstacleDev) * w->spatial_h->n_obstacles + 15) & ~(size_t)15;
    off += (size_t)MMMP_COLUMN_FLOATS(A.slots, A.links, A.mode) * sizeof(float) * block;
    return off;
}

template <class K>
int pick_launch(K kernel, const mmmp_world* w, const SpatialArgs& A, int64_t work_items_per_cta_hint,
                int64_t total_items, int& block, size_t& smem, int& grid) {
    int dev = 0, sms = 0, max_smem = 0;
    MMMP_CUDA_TRY(cudaGetDevice(&dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // CTA size: the staged world tables are a fixed cost per CTA and the columns scale with the
    // threads, so for wide columns (two-robot composite rows: 848 B per thread + 19 KB of tables)
    // smaller CTAs keep more THREADS resident -- take the size with the most threads per SM
    block = 0;
    int per_sm = 1;
    // the attribute belongs to the FUNCTION, not to the launch: always the device maximum, so that host
    // threads launching for different worlds (or trying different CTA sizes) never lower it under each other
    MMMP_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    for (int b = 128; b >= 32; b >>= 1) {
        const size_t need = smem_bytes(w, A, b);
        if (need > (size_t)max_smem) continue;
        int n = 0;
        MMMP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, b, need));
        if (n * b > per_sm * block) {
            block = b;
            per_sm = n;
            smem = need;
        }
    }
    if (block == 0) return MMMP_ERR_LIMIT;
    int64_t want = (total_items + work_items_per_cta_hint - 1) / work_items_per_cta_hint;
    const int64_t resident = (int64_t)sms * per_sm;
    // persistent grid: a whole number of waves, never more CTAs than work
    grid = (int)(want < resident ? (want < 1 ? 1 : want) : resident);
    return MMMP_OK;
}

unsigned long long* g_edge_stats = nullptr;  // device counters, enabled by mmmp_edge_stats

}  // namespace

extern "C" int mmmp_edge_stats(int enable, unsigned long long* out4_h) {
    constexpr size_t bytes = 4 * sizeof(unsigned long long);
    if (enable && !g_edge_stats) {
        MMMP_CUDA_TRY(cudaMalloc(&g_edge_stats, bytes));
        MMMP_CUDA_TRY(cudaMemset(g_edge_stats, 0, bytes));
    }
    if (out4_h && g_edge_stats) {
        MMMP_CUDA_TRY(cudaDeviceSynchronize());
        MMMP_CUDA_TRY(cudaMemcpy(out4_h, g_edge_stats, bytes, cudaMemcpyDeviceToHost));
        MMMP_CUDA_TRY(cudaMemset(g_edge_stats, 0, bytes));
    }
    if (!enable && g_edge_stats) {
        cudaFree(g_edge_stats);
        g_edge_stats = nullptr;
    }
    return MMMP_OK;
}

extern "C" int mmmp_planar_collide_configs(const mmmp_world* w, int arm, const double* q, int ld, int64_t N,
                                           uint32_t flags, uint8_t* out, void* stream);
extern "C" int mmmp_planar_collide_edges(const mmmp_world* w, int arm, const double* qa, const double* qb,
                                         int ld, const int32_t* edges, int64_t E, int n_steps, double step,
                                         uint32_t flags, uint8_t* out, void* stream);

extern "C" int mmmp_collide_configs(const mmmp_world* w, int robot, const void* q, int ld, int64_t N,
                                    uint32_t flags, uint8_t* out, void* stream) {
    if (!w || !q || !out || N < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (N == 0) return MMMP_OK;
    if (w->kind == MMMP_WORLD_PLANAR)
        return mmmp_planar_collide_configs(w, robot, (const double*)q, ld, N, flags, out, stream);
    SpatialArgs A;
    int rc = fill_args(w, robot, flags, A);
    if (rc) return rc;
    int block, grid;
    size_t smem;
    rc = pick_launch(collide_configs_kernel, w, A, 128, N, block, smem, grid);
    if (rc) return rc;
    grid = mmmp_div_up(N, block) < grid ? mmmp_div_up(N, block) : grid;
    MMMP_LAUNCH(collide_configs_kernel, grid, block, smem, stream, A, (const float*)q, ld, N, out);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_collide_robot_pairs(const mmmp_world* w, int ra, int rb, const void* qa, const void* qb,
                                        int ld, int64_t N, uint8_t* out, void* stream) {
    if (!w || !qa || !qb || !out || N < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (N == 0) return MMMP_OK;
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    const WorldDev* H = w->spatial_h;
    if (ra < 0 || rb < 0 || ra >= H->n_robots || rb >= H->n_robots || ra == rb) return MMMP_ERR_ARG;
    SpatialArgs A;
    A.W = w->spatial_d;
    A.flags = MMMP_CHECK_ROBOTS;
    A.n_r = 2;
    A.robots[0] = ra;
    A.robots[1] = rb;
    A.col[0] = A.col[1] = 0;
    A.mode = 1;
    A.slots = H->robots[ra].n_frame_store[1] + H->robots[rb].n_frame_store[1];
    A.links = H->robots[ra].n_link_store[1] + H->robots[rb].n_link_store[1];
    int block, grid;
    size_t smem;
    int rc = pick_launch(collide_pairs_kernel, w, A, 128, N, block, smem, grid);
    if (rc) return rc;
    grid = mmmp_div_up(N, block) < grid ? mmmp_div_up(N, block) : grid;
    MMMP_LAUNCH(collide_pairs_kernel, grid, block, smem, stream, A, (const float*)qa, (const float*)qb, ld, N, out);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

static int launch_edges(const mmmp_world* w, int robot, const float* qa, const float* qb, int ld,
                        const int32_t* edges, int64_t E, int n_steps, double step, uint32_t flags,
                        uint8_t* out, void* stream, const int32_t* E_dev = nullptr) {
    if (n_steps < 1 || n_steps > 4096) return MMMP_ERR_ARG;
    SpatialArgs A;
    int rc = fill_args(w, robot, flags & ~(uint32_t)MMMP_CHECK_EXHAUSTIVE, A);
    if (rc) return rc;
    int block, grid;
    size_t smem;
    if ((flags & MMMP_CHECK_EXHAUSTIVE) || n_steps > 64 || n_steps < 4 || !(step > 0.0)) {
        rc = pick_launch(collide_edges_kernel, w, A, 4, E, block, smem, grid);
        if (rc) return rc;
        const int wpc = block / 32;
        grid = mmmp_div_up(E, wpc) < grid ? mmmp_div_up(E, wpc) : grid;
        MMMP_LAUNCH(collide_edges_kernel, grid, block, smem, stream, A, qa, qb, ld, edges, E, n_steps, step, out);
    } else {
        constexpr int L = MMMP_EDGE_LANES;
        rc = pick_launch(collide_edges_adaptive_kernel<L>, w, A, 4 * (32 / L) * 8, E, block, smem, grid);
        if (rc) return rc;
        // clearance a lane asks for = (its share of the undecided steps) * step * refine.  Measured on
        // B200 (10M roadmap edges, 50 steps): refining bounding-sphere clearances at the sphere level
        // saves a third of the evaluations but costs 2-4x the time, so the default takes them as is.
        double refine = 0.0;
        if (const char* s = getenv("MMMP_EDGE_REFINE")) refine = atof(s);
        const float want_scale = (float)(step / 0.999 * refine);
        // edges still holding more than hard_after undecided steps after their first two rounds go
        // to the warp-per-edge pass (negative: never)
        int hard_after = (n_steps * 3) / 5;
        if (const char* s = getenv("MMMP_EDGE_HARD")) hard_after = atoi(s);
        cudaStream_t st = (cudaStream_t)stream;
        int64_t* hard_edge = nullptr;
        unsigned long long *hard_mask = nullptr, *hard_count = nullptr;
        int block2 = 0, grid2 = 0;
        size_t smem2 = 0;
        if (hard_after >= 0 && pick_launch(collide_edges_hard_kernel, w, A, 4, E, block2, smem2, grid2) != MMMP_OK)
            hard_after = -1;
        if (hard_after >= 0) {
            mmmp_pool_keep();
            MMMP_CUDA_TRY(cudaMallocAsync(&hard_edge, sizeof(int64_t) * E, st));
            cudaError_t e1 = cudaMallocAsync(&hard_mask, sizeof(unsigned long long) * (E + 1), st);
            if (e1 != cudaSuccess) {
                cudaFreeAsync(hard_edge, st);
                return MMMP_ERR_CUDA_BASE + (int)e1;
            }
            hard_count = hard_mask + E;
            cudaMemsetAsync(hard_count, 0, sizeof(unsigned long long), st);
        }
        MMMP_LAUNCH(collide_edges_adaptive_kernel<L>, grid, block, smem, stream, A, qa, qb, ld, edges, E, E_dev, n_steps,
                    step, want_scale, hard_after, hard_edge, hard_mask, hard_count, g_edge_stats, out);
        if (hard_edge) {
            MMMP_LAUNCH(collide_edges_hard_kernel, grid2, block2, smem2, stream, A, qa, qb, ld, edges, hard_edge,
                        hard_mask, hard_count, step, g_edge_stats, out);
            cudaFreeAsync(hard_edge, st);
            cudaFreeAsync(hard_mask, st);
        }
    }
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_collide_edges(const mmmp_world* w, int robot, const void* q, int ld, const int32_t* edges,
                                  int64_t E, int n_steps, double step, uint32_t flags, uint8_t* out,
                                  void* stream) {
    if (!w || !q || !edges || !out || E < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (E == 0) return MMMP_OK;
    if (w->kind == MMMP_WORLD_PLANAR)
        return mmmp_planar_collide_edges(w, robot, (const double*)q, nullptr, ld, edges, E, n_steps, step, flags,
                                         out, stream);
    return launch_edges(w, robot, (const float*)q, nullptr, ld, edges, E, n_steps, step, flags, out, stream);
}

extern "C" int mmmp_collide_segments(const mmmp_world* w, int robot, const void* qa, const void* qb, int ld,
                                     int64_t E, int n_steps, double step, uint32_t flags, uint8_t* out,
                                     void* stream) {
    if (!w || !qa || !qb || !out || E < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (E == 0) return MMMP_OK;
    if (w->kind == MMMP_WORLD_PLANAR)
        return mmmp_planar_collide_edges(w, robot, (const double*)qa, (const double*)qb, ld, nullptr, E, n_steps,
                                         step, flags, out, stream);
    return launch_edges(w, robot, (const float*)qa, (const float*)qb, ld, nullptr, E, n_steps, step, flags, out,
                        stream);
}

// ---------------------------------------------------------------- verdicts of a candidate table
// The reference's accept loop (decoupled_prm.py:496-503) checks the edge node -> neighbour for every
// candidate slot it reaches.  When i is a candidate of j AND j a candidate of i, the two checks visit
// the same configurations: t = s*step, s = 0 .. n_steps-1 from either end is the same interior grid
// when n_steps*step == 1 (np.arange(0, 1, local_step), :859), plus the end the walk starts from -- a
// roadmap node, collision free by construction.  Such a pair is evaluated ONCE, from the lower node id
// to the higher (the direction of the first of the reference's two checks: nodes are processed in id
// order), and both slots receive that verdict.
namespace {

enum : uint8_t { SLOT_EVAL = 0, SLOT_EMPTY = 1, SLOT_SHARED = 2 };

__device__ __forceinline__ int find_in_row(const int32_t* __restrict__ row, int k, int32_t what) {
    for (int s = 0; s < k; ++s)
        if (__ldg(row + s) == what) return s;
    return -1;
}

// one thread per slot of rows [row_lo, row_hi): who evaluates the pair, in which direction
__global__ void __launch_bounds__(256)
knn_edges_classify_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int64_t row_lo, int64_t row_hi, int64_t win_lo,
                          int64_t win_hi, int share, uint8_t* __restrict__ skip, int32_t* __restrict__ pair,
                          uint8_t* __restrict__ free_out) {
    const int64_t total = (row_hi - row_lo) * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = row_lo + e / k;
        const int32_t j = __ldg(nbr + row_lo * k + e);
        uint8_t code = SLOT_EVAL;
        int32_t a = (int32_t)i, b = j;
        if (j < 0 || j >= n) {   // an id outside the table is treated like an empty slot: blocked, never read
            code = SLOT_EMPTY;
            b = -1;
        } else if (share && j != i && find_in_row(nbr + (int64_t)j * k, k, (int32_t)i) >= 0) {
            a = i < j ? (int32_t)i : j;    // mutual candidates: the reference's first check runs low -> high
            b = i < j ? j : (int32_t)i;
            // one of the two slots evaluates, chosen by the parity of i + j so that no block of rows
            // ends up with more than its share; the partner must be inside the window of rows whose
            // verdicts the caller will see
            const int32_t owner = ((i + j) & 1) ? a : b;
            if (owner != (int32_t)i && j >= win_lo && j < win_hi) code = SLOT_SHARED;
        }
        skip[e] = code;
        pair[2 * e] = a;
        pair[2 * e + 1] = b;
        free_out[e] = code == SLOT_SHARED ? 2 : 0;
    }
}

// edges[p] = pair of the p-th evaluated slot, (-1, -1) behind the last one
__global__ void __launch_bounds__(256)
knn_edges_emit_kernel(const int32_t* __restrict__ slots, const int32_t* __restrict__ count,
                      const int32_t* __restrict__ pair, int64_t total, int32_t* __restrict__ edges) {
    const int64_t live = __ldg(count);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        int2 v = make_int2(-1, -1);
        if (p < live) v = __ldg(reinterpret_cast<const int2*>(pair) + __ldg(slots + p));
        reinterpret_cast<int2*>(edges)[p] = v;
    }
}

__global__ void __launch_bounds__(256)
knn_edges_expand_kernel(const int32_t* __restrict__ slots, const int32_t* __restrict__ count,
                        const uint8_t* __restrict__ hit, int64_t total, uint8_t* __restrict__ free_out) {
    const int64_t live = __ldg(count);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total && p < live;
         p += (int64_t)gridDim.x * blockDim.x)
        free_out[__ldg(slots + p)] = hit[p] ? 0 : 1;
}

// slots marked 2 take the verdict of the partner's slot (never itself marked 2); the table holds rows
// [row_lo, row_hi), both ends of a marked slot lie inside
__global__ void __launch_bounds__(256)
knn_edges_resolve_kernel(const int32_t* __restrict__ nbr, int k, int64_t row_lo, int64_t row_hi,
                         uint8_t* __restrict__ free_io) {
    const int64_t total = (row_hi - row_lo) * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        if (free_io[e] != 2) continue;
        const int64_t i = row_lo + e / k;
        const int32_t j = __ldg(nbr + row_lo * k + e);
        uint8_t v = 0;
        if (j >= row_lo && j < row_hi) {
            const int s = find_in_row(nbr + (int64_t)j * k, k, (int32_t)i);
            if (s >= 0) v = free_io[(j - row_lo) * k + s];
        }
        free_io[e] = v == 1 ? 1 : 0;
    }
}

int slot_grid(int64_t items) {
    int64_t g = (items + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

extern "C" int mmmp_compact_free(const uint8_t* flags, int64_t N, int32_t* idx_out, int32_t* count, void* stream);

extern "C" int mmmp_resolve_knn_edges(const int32_t* nbr_d, int64_t n, int k, int64_t row_lo, int64_t row_hi,
                                      uint8_t* free_d, void* stream) {
    if (!nbr_d || !free_d || n < 0 || k < 1 || row_lo < 0 || row_hi > n || row_lo > row_hi) return MMMP_ERR_ARG;
    if (row_lo == row_hi) return MMMP_OK;
    MMMP_LAUNCH(knn_edges_resolve_kernel, slot_grid((row_hi - row_lo) * k), 256, 0, stream, nbr_d, k, row_lo, row_hi,
                free_d);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_collide_knn_edges(const mmmp_world* w, int robot, const void* q, int ld, const int32_t* nbr_d,
                                      int64_t n, int k, int64_t row_lo, int64_t row_hi, int share_across_rows,
                                      int n_steps, double step, uint32_t flags, uint8_t* free_d, void* stream) {
    if (!w || !q || !nbr_d || !free_d || n < 0 || k < 1 || row_lo < 0 || row_hi > n || row_lo > row_hi)
        return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    const int64_t total = (row_hi - row_lo) * k;
    if (total == 0) return MMMP_OK;
    if (total > 0x7fffffff) return MMMP_ERR_ARG;
    // the two walks visit the same configurations only on a grid that is symmetric about t = 1/2
    double sym = (double)n_steps * step - 1.0;
    const int share = (sym < 0 ? -sym : sym) < 1e-9 && !getenv("MMMP_EDGES_NO_SHARE");
    cudaStream_t st = (cudaStream_t)stream;
    mmmp_pool_keep();
    // skip [total] u8 | hit [total] u8 | count i32 (+pad) | slots [total] i32 | pair [total] int2 | edges [total] int2
    const size_t off_hit = (size_t)((total + 15) & ~(int64_t)15);
    const size_t off_count = 2 * off_hit, off_slots = off_count + 16;
    const size_t off_pair = off_slots + sizeof(int32_t) * (size_t)((total + 3) & ~(int64_t)3);
    const size_t off_edges = off_pair + sizeof(int32_t) * 2 * (size_t)total;
    unsigned char* buf = nullptr;
    MMMP_CUDA_TRY(cudaMallocAsync(&buf, off_edges + sizeof(int32_t) * 2 * (size_t)total, st));
    uint8_t *skip = buf, *hit = buf + off_hit;
    int32_t* count = reinterpret_cast<int32_t*>(buf + off_count);
    int32_t* slots = reinterpret_cast<int32_t*>(buf + off_slots);
    int32_t* pair = reinterpret_cast<int32_t*>(buf + off_pair);
    int32_t* edges = reinterpret_cast<int32_t*>(buf + off_edges);
    const int64_t win_lo = share_across_rows ? 0 : row_lo, win_hi = share_across_rows ? n : row_hi;
    int rc = MMMP_OK;
    MMMP_LAUNCH(knn_edges_classify_kernel, slot_grid(total), 256, 0, stream, nbr_d, n, k, row_lo, row_hi, win_lo, win_hi,
                share, skip, pair, free_d);
    rc = mmmp_compact_free(skip, total, slots, count, stream);
    if (rc == MMMP_OK) {
        MMMP_LAUNCH(knn_edges_emit_kernel, slot_grid(total), 256, 0, stream, slots, count, pair, total, edges);
        rc = launch_edges(w, robot, (const float*)q, nullptr, ld, edges, total, n_steps, step, flags, hit, stream, count);
    }
    if (rc == MMMP_OK) {
        MMMP_LAUNCH(knn_edges_expand_kernel, slot_grid(total), 256, 0, stream, slots, count, hit, total, free_d);
        if (share && !share_across_rows)
            MMMP_LAUNCH(knn_edges_resolve_kernel, slot_grid(total), 256, 0, stream, nbr_d, k, row_lo, row_hi, free_d);
    }
    cudaFreeAsync(buf, st);
    if (rc != MMMP_OK) return rc;
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

#ifdef MMMP_EDGE_DIAG
extern "C" int mmmp_edge_diag(unsigned long long* out64_h) {
    if (out64_h) MMMP_CUDA_TRY(cudaMemcpyFromSymbol(out64_h, g_diag, sizeof(unsigned long long) * 64));
    unsigned long long zero[64] = {0};
    MMMP_CUDA_TRY(cudaMemcpyToSymbol(g_diag, zero, sizeof(zero)));
    return MMMP_OK;
}
#endif
