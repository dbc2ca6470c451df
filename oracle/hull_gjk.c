/*
 * oracle/hull_gjk.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (float64) of what the reference obtains from pybullet's
 * getClosestPoints(bodyA, bodyB, distance=0.0) at
 *   planners/prm/pybullet/env.py:65  (robot vs obstacle),
 *   planners/prm/pybullet/env.py:72  (robot vs robot),
 *   planners/prm/pybullet/env.py:81  (link i vs link j of one robot):
 * "is the distance between two convex bodies <= threshold".
 * The arithmetic lives in pybullet==3.2.6 (Bullet 3.x, requirements.txt), which
 * is absent from the reference tree and not installable here; Bullet treats a
 * URDF collision mesh as its CONVEX HULL and answers with a GJK/EPA closest
 * point query.  This file restates that published algorithm: GJK distance
 * between convex hulls of point sets (optionally inflated by a radius =
 * Bullet's collision margin / a sphere or capsule primitive).
 *
 * PARITY UNPINNED: no pybullet in the build container nor on the GPU box
 * (probed: profiles/r02_pybullet_probe.txt) and none of the reference's tests
 * records a collision verdict, so this oracle is pinned only by known answers and
 * by an independent algorithm (LP feasibility / QP distance on the real Panda
 * hulls): tests/test_oracle_cpu.py::test_hull_oracle_known_answers and
 * ::test_hull_oracle_vs_linear_programming -- not by reference output.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call this code.
 */
#include <math.h>
#include <stddef.h>

typedef struct {
    const double* v; /* n x 3 world coordinates */
    int n;
    double radius;   /* Minkowski inflation */
} hull_shape;

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void sub3(const double* a, const double* b, double* o) {
    o[0] = a[0] - b[0];
    o[1] = a[1] - b[1];
    o[2] = a[2] - b[2];
}
static void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

static int support_idx(const hull_shape* s, const double* d) {
    int best = 0;
    double bv = dot3(s->v, d);
    for (int i = 1; i < s->n; ++i) {
        const double t = dot3(s->v + 3 * i, d);
        if (t > bv) {
            bv = t;
            best = i;
        }
    }
    return best;
}

/* closest point to the origin on segment / triangle; returns the number of
 * vertices that support it and compacts them to the front of p[] */
static int closest_segment(double p[][3], double* out) {
    double ab[3];
    sub3(p[1], p[0], ab);
    const double den = dot3(ab, ab);
    double t = den > 0 ? -dot3(p[0], ab) / den : 0.0;
    if (t <= 0) {
        for (int i = 0; i < 3; ++i) out[i] = p[0][i];
        return 1;
    }
    if (t >= 1) {
        for (int i = 0; i < 3; ++i) {
            out[i] = p[1][i];
            p[0][i] = p[1][i];
        }
        return 1;
    }
    for (int i = 0; i < 3; ++i) out[i] = p[0][i] + t * ab[i];
    return 2;
}

static int closest_triangle(double p[][3], double* out) {
    /* Voronoi-region walk (Ericson, Real-Time Collision Detection 5.1.5), query point = origin */
    double ab[3], ac[3], ap[3], bp[3], cp[3];
    const double* a = p[0];
    const double* b = p[1];
    const double* c = p[2];
    sub3(b, a, ab);
    sub3(c, a, ac);
    for (int i = 0; i < 3; ++i) {
        ap[i] = -a[i];
        bp[i] = -b[i];
        cp[i] = -c[i];
    }
    const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0 && d2 <= 0) {
        for (int i = 0; i < 3; ++i) out[i] = a[i];
        return 1;
    }
    const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) {
        for (int i = 0; i < 3; ++i) {
            out[i] = b[i];
            p[0][i] = b[i];
        }
        return 1;
    }
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        const double v = d1 / (d1 - d3);
        for (int i = 0; i < 3; ++i) out[i] = a[i] + v * ab[i];
        return 2; /* a, b already in slots 0,1 */
    }
    const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) {
        for (int i = 0; i < 3; ++i) {
            out[i] = c[i];
            p[0][i] = c[i];
        }
        return 1;
    }
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        const double w = d2 / (d2 - d6);
        for (int i = 0; i < 3; ++i) {
            out[i] = a[i] + w * ac[i];
            p[1][i] = c[i];
        }
        return 2; /* a, c */
    }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int i = 0; i < 3; ++i) {
            out[i] = b[i] + w * (c[i] - b[i]);
        }
        for (int i = 0; i < 3; ++i) {
            p[0][i] = b[i];
            p[1][i] = c[i];
        }
        return 2; /* b, c */
    }
    const double den = 1.0 / (va + vb + vc);
    const double v = vb * den, w = vc * den;
    for (int i = 0; i < 3; ++i) out[i] = a[i] + ab[i] * v + ac[i] * w;
    return 3;
}

/* tetrahedron: returns 4 if the origin is inside (intersection) */
static int closest_tetra(double p[][3], double* out) {
    static const int faces[4][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 3, 1}, {1, 2, 3, 0}};
    double best = INFINITY;
    int bestn = 4;
    double bestp[3][3], bestout[3] = {0, 0, 0};
    int outside_any = 0;
    for (int f = 0; f < 4; ++f) {
        const double* a = p[faces[f][0]];
        const double* b = p[faces[f][1]];
        const double* c = p[faces[f][2]];
        const double* d = p[faces[f][3]];
        double ab[3], ac[3], n[3], ad[3];
        sub3(b, a, ab);
        sub3(c, a, ac);
        cross3(ab, ac, n);
        sub3(d, a, ad);
        const double sd = dot3(n, ad);        /* side of the opposite vertex */
        const double so = -dot3(n, a);        /* side of the origin */
        /* flat tetra (the fourth vertex within rounding of the face plane): the side tests
         * below would be decided by noise and could report a bogus containment, so the
         * face is only a candidate for the closest feature */
        if (fabs(sd) <= 1e-10 * sqrt(dot3(n, n) * dot3(ad, ad))) {
            outside_any = 1;
        } else if ((so > 0) == (sd > 0) || so == 0.0) {
            continue; /* origin on the inner side of this face */
        } else {
            outside_any = 1;
        }
        double q[3][3], o[3];
        for (int i = 0; i < 3; ++i) {
            q[0][i] = a[i];
            q[1][i] = b[i];
            q[2][i] = c[i];
        }
        const int m = closest_triangle(q, o);
        const double dd = dot3(o, o);
        if (dd < best) {
            best = dd;
            bestn = m;
            for (int k = 0; k < 3; ++k)
                for (int i = 0; i < 3; ++i) bestp[k][i] = q[k][i];
            for (int i = 0; i < 3; ++i) bestout[i] = o[i];
        }
    }
    if (!outside_any) return 4;
    for (int k = 0; k < bestn; ++k)
        for (int i = 0; i < 3; ++i) p[k][i] = bestp[k][i];
    for (int i = 0; i < 3; ++i) out[i] = bestout[i];
    return bestn;
}

/* distance between conv(A)+ball(rA) and conv(B)+ball(rB); 0 when they overlap or touch */
double hull_gjk_distance(const double* va, int na, double ra, const double* vb, int nb, double rb) {
    hull_shape A = {va, na, ra}, B = {vb, nb, rb};
    double simplex[4][3];
    double v[3];
    int ns;
    sub3(A.v, B.v, v);
    for (int i = 0; i < 3; ++i) simplex[0][i] = v[i];
    ns = 1;
    double dist2 = dot3(v, v);
    for (int it = 0; it < 256; ++it) {
        if (dist2 <= 1e-30) return 0.0;
        double nd[3] = {-v[0], -v[1], -v[2]};
        const int ia = support_idx(&A, nd);
        const int ib = support_idx(&B, v);
        double w[3];
        sub3(A.v + 3 * ia, B.v + 3 * ib, w);
        const double vw = dot3(v, w);
        /* v is optimal when no support point lies further along -v */
        if (dist2 - vw <= 1e-14 * dist2 + 1e-300) break;
        /* duplicate vertex => no progress */
        int dup = 0;
        for (int k = 0; k < ns; ++k)
            if (simplex[k][0] == w[0] && simplex[k][1] == w[1] && simplex[k][2] == w[2]) dup = 1;
        if (dup) break;
        for (int i = 0; i < 3; ++i) simplex[ns][i] = w[i];
        ++ns;
        double nv[3];
        if (ns == 2) ns = closest_segment(simplex, nv);
        else if (ns == 3) ns = closest_triangle(simplex, nv);
        else {
            ns = closest_tetra(simplex, nv);
            if (ns == 4) return 0.0;
        }
        const double nd2 = dot3(nv, nv);
        if (nd2 >= dist2) break; /* numerical stall */
        for (int i = 0; i < 3; ++i) v[i] = nv[i];
        dist2 = nd2;
    }
    const double d = sqrt(dist2) - ra - rb;
    return d > 0 ? d : 0.0;
}

/* batched: pairs of world-space vertex arrays given by offsets */
void hull_gjk_distance_batch(const double* verts, const int* off_a, const int* n_a, const double* r_a,
                             const int* off_b, const int* n_b, const double* r_b, int n_pairs, double* out) {
    for (int i = 0; i < n_pairs; ++i)
        out[i] = hull_gjk_distance(verts + 3 * (size_t)off_a[i], n_a[i], r_a[i], verts + 3 * (size_t)off_b[i], n_b[i],
                                   r_b[i]);
}

/* out[i] = distance between transformed hulls: A under rigid 3x4 TA[i], B under TB[i].
 * scratch must hold (na + nb) * 3 doubles. */
void hull_gjk_posed_batch(const double* va, int na, double ra, const double* TA, const double* vb, int nb, double rb,
                          const double* TB, int n, double* scratch, double* out) {
    double* wa = scratch;
    double* wb = scratch + 3 * (size_t)na;
    for (int i = 0; i < n; ++i) {
        const double* a = TA + 12 * (size_t)i;
        const double* b = TB + 12 * (size_t)i;
        for (int k = 0; k < na; ++k)
            for (int r = 0; r < 3; ++r)
                wa[3 * k + r] = a[4 * r] * va[3 * k] + a[4 * r + 1] * va[3 * k + 1] + a[4 * r + 2] * va[3 * k + 2] + a[4 * r + 3];
        for (int k = 0; k < nb; ++k)
            for (int r = 0; r < 3; ++r)
                wb[3 * k + r] = b[4 * r] * vb[3 * k] + b[4 * r + 1] * vb[3 * k + 1] + b[4 * r + 2] * vb[3 * k + 2] + b[4 * r + 3];
        out[i] = hull_gjk_distance(wa, na, ra, wb, nb, rb);
    }
}

/* min over vertices of n.x - d (signed distance of a hull to the half-space n.x <= d) */
void hull_plane_posed_batch(const double* va, int na, double ra, const double* TA, const double* plane, int n,
                            double* out) {
    for (int i = 0; i < n; ++i) {
        const double* a = TA + 12 * (size_t)i;
        double best = INFINITY;
        for (int k = 0; k < na; ++k) {
            double w[3];
            for (int r = 0; r < 3; ++r)
                w[r] = a[4 * r] * va[3 * k] + a[4 * r + 1] * va[3 * k + 1] + a[4 * r + 2] * va[3 * k + 2] + a[4 * r + 3];
            const double sd = plane[0] * w[0] + plane[1] * w[1] + plane[2] * w[2] - plane[3];
            if (sd < best) best = sd;
        }
        best -= ra;
        out[i] = best > 0 ? best : 0.0;
    }
}
