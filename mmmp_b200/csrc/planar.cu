// Planar worlds: PlanarArm FK and the analytic 2-D collision predicates, in
// fp64 with explicitly rounded (non-fused) arithmetic in the reference's own
// operation order, so that verdicts are identical to the Python reference
// except when a predicate value lies within an ulp of its threshold.
//
// Reference: robots/planar_arm.py:23-40 (FK);
// planners/prm/planar/env.py:78-229 (predicates);
// wrappers planners/prm/planar/single_arm/prm.py:104-169,
// multi_arm/coupled/coupled_prm.py:83-109,
// multi_arm/decoupled/decoupled_prm.py:126-152.
#include "common.cuh"

namespace {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

struct P2 {
    double x, y;
};

// ccw(a,b,c): (b0-a0)*(c1-a1) - (b1-a1)*(c0-a0) > 0          env.py:134-137
__device__ __forceinline__ bool ccw(P2 a, P2 b, P2 c) {
    const double cp = dsub(dmul(dsub(b.x, a.x), dsub(c.y, a.y)), dmul(dsub(b.y, a.y), dsub(c.x, a.x)));
    return cp > 0.0;
}
// link_link_collision(p1,q1,p2,q2)                             env.py:132-141
__device__ __forceinline__ bool seg_seg(P2 p1, P2 q1, P2 p2, P2 q2) {
    return (ccw(p1, q1, p2) != ccw(p1, q1, q2)) && (ccw(p2, q2, p1) != ccw(p2, q2, q1));
}
__device__ __forceinline__ double dot2(double ax, double ay, double bx, double by) {
    return dadd(dmul(ax, bx), dmul(ay, by));
}
// joint_link_collision(joint, lower, upper)                    env.py:119-130
__device__ __forceinline__ bool joint_link(P2 j, P2 lo, P2 hi, double radius) {
    const double lx = dsub(hi.x, lo.x), ly = dsub(hi.y, lo.y);
    const double jx = dsub(j.x, lo.x), jy = dsub(j.y, lo.y);
    const double proj = dot2(jx, jy, lx, ly) / dot2(lx, ly, lx, ly);
    if (proj < 0.0 || proj > 1.0) return false;  // NaN (zero-length link) falls through like numpy
    const double cr = fabs(dsub(dmul(lx, jy), dmul(ly, jx)));
    const double dist = cr / sqrt(dot2(lx, ly, lx, ly));
    return dist <= radius;
}
// line_intersects_rect(p, q, rect)                             env.py:168-188
__device__ __forceinline__ bool seg_rect(P2 p, P2 q, const double* o) {
    const P2 v0{o[0], o[1]}, v1{o[2], o[1]}, v2{o[2], o[3]}, v3{o[0], o[3]};
    return seg_seg(p, q, v0, v1) || seg_seg(p, q, v1, v2) || seg_seg(p, q, v2, v3) || seg_seg(p, q, v3, v0);
}
// line_intersects_circle(p, q, circle)                         env.py:190-205
__device__ __forceinline__ bool seg_circle(P2 p, P2 q, const double* o) {
    const double vx = dsub(o[0], p.x), vy = dsub(o[1], p.y);
    const double dx = dsub(q.x, p.x), dy = dsub(q.y, p.y);
    const double t = dot2(vx, vy, dx, dy) / dot2(dx, dy, dx, dy);
    const double cx = dadd(p.x, dmul(t, dx)), cy = dadd(p.y, dmul(t, dy));
    const double ex = dsub(cx, o[0]), ey = dsub(cy, o[1]);
    const double n = sqrt(dadd(dmul(ex, ex), dmul(ey, ey)));
    return (n <= o[2]) && (0.0 <= t) && (t <= 1.0);
}

struct PlanarQ {
    const double* a;
    const double* b;
    double t;
    bool lerp;
    __device__ __forceinline__ double get(int c) const {
        if (!lerp) return a[c];
        // sample = q1 + t * (q2 - q1)                          single_arm/prm.py:166
        return dadd(a[c], dmul(t, dsub(b[c], a[c])));
    }
};

// FK of arm `arm` into px/py (link end points), planar_arm.py:23-40
__device__ __forceinline__ void planar_fk(const PlanarDev& W, int arm, int col, const PlanarQ& q, double* px,
                                          double* py) {
    const int n = W.n_links[arm];
    double cum = 0.0, x = W.base[arm][0], y = W.base[arm][1];
    for (int i = 0; i < n; ++i) {
        const double qi = q.get(col + i);
        cum = (i == 0) ? qi : dadd(cum, qi);
        double s, c;
        sincos(cum, &s, &c);
        x = dadd(x, dmul(W.link_len[arm][i], c));
        y = dadd(y, dmul(W.link_len[arm][i], s));
        px[i] = x;
        py[i] = y;
    }
}

__device__ __forceinline__ bool in_bounds(const PlanarDev& W, double x, double y) {
    return W.bounds[0] <= x && x <= W.bounds[1] && W.bounds[2] <= y && y <= W.bounds[3];
}

__device__ bool planar_in_collision(const PlanarDev& W, int arm_sel, uint32_t flags, const PlanarQ& q) {
    double px[MMMP_PLANAR_MAX_LINKS], py[MMMP_PLANAR_MAX_LINKS];
    const int a0 = arm_sel >= 0 ? arm_sel : 0;
    const int a1 = arm_sel >= 0 ? arm_sel + 1 : W.n_arms;
    for (int a = a0; a < a1; ++a) {
        const int col = arm_sel >= 0 ? 0 : W.dof_offset[a];
        const int lb = arm_sel >= 0 ? 0 : W.link_begin[a];
        planar_fk(W, a, col, q, px + lb, py + lb);
    }
    for (int a = a0; a < a1; ++a) {
        const int lb = arm_sel >= 0 ? 0 : W.link_begin[a];
        const int n = W.n_links[a];
        const P2 base{W.base[a][0], W.base[a][1]};
        if (flags & MMMP_CHECK_BOUNDS) {  // env.py:220-229
            if (!in_bounds(W, base.x, base.y)) return true;
            for (int i = 0; i < n; ++i)
                if (!in_bounds(W, px[lb + i], py[lb + i])) return true;
        }
        if (flags & MMMP_CHECK_SELF) {  // env.py:208-217
            for (int i = 0; i < n; ++i) {
                const P2 pi{px[lb + i], py[lb + i]};
                const P2 prev = (i == 0) ? base : P2{px[lb + i - 1], py[lb + i - 1]};
                for (int j = i + 2; j < n; ++j) {
                    const P2 pj{px[lb + j], py[lb + j]}, pj1{px[lb + j - 1], py[lb + j - 1]};
                    if (seg_seg(pi, prev, pj, pj1)) return true;
                }
            }
        }
        if (flags & MMMP_CHECK_OBSTACLES) {  // env.py:148-166
            for (int o = 0; o < W.n_obstacles; ++o)
                for (int i = 0; i < n; ++i) {
                    const P2 pi{px[lb + i], py[lb + i]};
                    const P2 prev = (i == 0) ? base : P2{px[lb + i - 1], py[lb + i - 1]};
                    const bool h = (W.obs_type[o] == MMMP_POBS_RECT) ? seg_rect(pi, prev, W.obs[o])
                                                                      : seg_circle(pi, prev, W.obs[o]);
                    if (h) return true;
                }
        }
    }
    if ((flags & MMMP_CHECK_ROBOTS) && arm_sel < 0) {  // env.py:78-117
        for (int a = 0; a < W.n_arms; ++a)
            for (int b = a + 1; b < W.n_arms; ++b) {
                const int la = W.link_begin[a], lb2 = W.link_begin[b];
                const int na = W.n_links[a], nb = W.n_links[b];
                const P2 ba{W.base[a][0], W.base[a][1]}, bb{W.base[b][0], W.base[b][1]};
                for (int i = 0; i < na; ++i) {
                    const P2 pi{px[la + i], py[la + i]};
                    const P2 prev = (i == 0) ? ba : P2{px[la + i - 1], py[la + i - 1]};
                    for (int j = 0; j < nb; ++j) {
                        const P2 pj{px[lb2 + j], py[lb2 + j]};
                        const P2 prev2 = (j == 0) ? bb : P2{px[lb2 + j - 1], py[lb2 + j - 1]};
                        if (seg_seg(pi, prev, pj, prev2)) return true;
                    }
                }
                for (int i = 0; i < na; ++i) {
                    const P2 ji{px[la + i], py[la + i]};
                    for (int j = 0; j < nb; ++j) {
                        const P2 lo = (j == 0) ? bb : P2{px[lb2 + j - 1], py[lb2 + j - 1]};
                        const P2 hi{px[lb2 + j], py[lb2 + j]};
                        if (joint_link(ji, lo, hi, W.joint_link_radius)) return true;
                    }
                }
                for (int i = 0; i < nb; ++i) {
                    const P2 ji{px[lb2 + i], py[lb2 + i]};
                    for (int j = 0; j < na; ++j) {
                        const P2 lo = (j == 0) ? ba : P2{px[la + j - 1], py[la + j - 1]};
                        const P2 hi{px[la + j], py[la + j]};
                        if (joint_link(ji, lo, hi, W.joint_link_radius)) return true;
                    }
                }
            }
    }
    return false;
}

__global__ void __launch_bounds__(128)
planar_fk_kernel(const PlanarDev* __restrict__ Wg, int arm, const double* __restrict__ q, int ld, int64_t N,
                 double* __restrict__ out) {
    __shared__ PlanarDev W;
    for (int i = threadIdx.x; i < (int)(sizeof(PlanarDev) / 4); i += blockDim.x)
        reinterpret_cast<int*>(&W)[i] = reinterpret_cast<const int*>(Wg)[i];
    __syncthreads();
    const int n = W.n_links[arm];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double px[MMMP_MAX_JOINTS], py[MMMP_MAX_JOINTS];
        PlanarQ ql{q + i * ld, nullptr, 0.0, false};
        planar_fk(W, arm, 0, ql, px, py);
        for (int l = 0; l < n; ++l) {
            out[(i * n + l) * 2 + 0] = px[l];
            out[(i * n + l) * 2 + 1] = py[l];
        }
    }
}

__global__ void __launch_bounds__(128)
planar_configs_kernel(const PlanarDev* __restrict__ Wg, int arm, const double* __restrict__ q, int ld, int64_t N,
                      uint32_t flags, uint8_t* __restrict__ out) {
    __shared__ PlanarDev W;
    for (int i = threadIdx.x; i < (int)(sizeof(PlanarDev) / 4); i += blockDim.x)
        reinterpret_cast<int*>(&W)[i] = reinterpret_cast<const int*>(Wg)[i];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        PlanarQ ql{q + i * ld, nullptr, 0.0, false};
        out[i] = planar_in_collision(W, arm, flags, ql) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(128)
planar_edges_kernel(const PlanarDev* __restrict__ Wg, int arm, const double* __restrict__ qa_base,
                    const double* __restrict__ qb_base, int ld, const int32_t* __restrict__ edges, int64_t E,
                    int n_steps, double step, uint32_t flags, uint8_t* __restrict__ out) {
    __shared__ PlanarDev W;
    for (int i = threadIdx.x; i < (int)(sizeof(PlanarDev) / 4); i += blockDim.x)
        reinterpret_cast<int*>(&W)[i] = reinterpret_cast<const int*>(Wg)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int rounds = (n_steps + 31) >> 5;
    for (int64_t e = (int64_t)blockIdx.x * wpc + (threadIdx.x >> 5); e < E; e += (int64_t)gridDim.x * wpc) {
        const double *pa, *pb;
        if (edges) {
            const int ia = edges[2 * e], ib = edges[2 * e + 1];
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
                PlanarQ ql{pa, pb, dmul((double)i, step), true};  // np.arange(0,1,step)[i] = i*step
                hit = planar_in_collision(W, arm, flags, ql);
            }
            any = __any_sync(0xffffffffu, hit);
        }
        if (lane == 0) out[e] = any ? 1 : 0;
    }
}

int planar_grid(int64_t items, int per_cta) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (items + per_cta - 1) / per_cta;
    const int64_t cap = (int64_t)sms * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

extern "C" int mmmp_planar_fk(const mmmp_world* w, int arm, const double* q, int ld, int64_t N, double* out,
                              void* stream) {
    if (!w || !q || !out || N < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (w->kind != MMMP_WORLD_PLANAR) return MMMP_ERR_KIND;
    if (arm < 0 || arm >= w->planar_h->n_arms) return MMMP_ERR_ARG;
    if (N == 0) return MMMP_OK;
    MMMP_LAUNCH(planar_fk_kernel, planar_grid(N, 128), 128, 0, stream, w->planar_d, arm, q, ld, N, out);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_planar_collide_configs(const mmmp_world* w, int arm, const double* q, int ld, int64_t N,
                                           uint32_t flags, uint8_t* out, void* stream) {
    if (arm >= w->planar_h->n_arms) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    MMMP_LAUNCH(planar_configs_kernel, planar_grid(N, 128), 128, 0, stream, w->planar_d, arm, q, ld, N, flags, out);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_planar_collide_edges(const mmmp_world* w, int arm, const double* qa, const double* qb, int ld,
                                         const int32_t* edges, int64_t E, int n_steps, double step, uint32_t flags,
                                         uint8_t* out, void* stream) {
    if (arm >= w->planar_h->n_arms) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (n_steps < 1 || n_steps > 4096) return MMMP_ERR_ARG;
    MMMP_LAUNCH(planar_edges_kernel, planar_grid(E, 4), 128, 0, stream, w->planar_d, arm, qa, qb, ld, edges, E,
                n_steps, step, flags, out);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}
