// Sphere-vs-primitive overlap tests (touching counts as collision, matching
// pybullet getClosestPoints(distance=0) returning a point at distance <= 0:
// reference planners/prm/pybullet/env.py:65-67).
#pragma once
#include "common.cuh"

// sqrt for CLEARANCES only (never for a verdict): x * MUFU.RSQ(x), without the denormal fix-up
// rsqrtf() carries (6 % of the edge kernel's instructions, profiles/r01_edge_full.csv).  The
// ~2 ulp error is far inside MMMP_GAP_SLACK.
__device__ __forceinline__ float fast_sqrt(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaxf(x, 1e-30f)));
    return x * y;
}

// Returns the overlap verdict.  With GAP the sphere's clearance from the primitive (distance
// between the surfaces, meaningful when there is no overlap) is folded into `gap` by min; the
// verdict is computed by the same expressions either way.
template <bool GAP>
__device__ __forceinline__ bool sphere_vs_obstacle(float cx, float cy, float cz, float r, const ObstacleDev& O,
                                                   float& gap) {
    const float* p = O.p;
    switch (O.type) {
        case MMMP_OBS_PLANE: {
            const float sd = fmaf(p[0], cx, fmaf(p[1], cy, p[2] * cz)) - p[3];
            if (GAP) gap = fminf(gap, sd - r);
            return sd <= r;
        }
        case MMMP_OBS_SPHERE: {
            const float dx = cx - p[0], dy = cy - p[1], dz = cz - p[2];
            const float rr = r + p[3];
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            if (GAP) gap = fminf(gap, fast_sqrt(d2) - rr);
            return d2 <= rr * rr;
        }
        case MMMP_OBS_BOX: {
            const float dx = cx - p[0], dy = cy - p[1], dz = cz - p[2];
            // box-frame coordinates: R^T d  (R row-major at p[6..14])
            const float bx = fmaf(p[6], dx, fmaf(p[9], dy, p[12] * dz));
            const float by = fmaf(p[7], dx, fmaf(p[10], dy, p[13] * dz));
            const float bz = fmaf(p[8], dx, fmaf(p[11], dy, p[14] * dz));
            const float ex = fmaxf(fabsf(bx) - p[3], 0.f);
            const float ey = fmaxf(fabsf(by) - p[4], 0.f);
            const float ez = fmaxf(fabsf(bz) - p[5], 0.f);
            const float e2 = fmaf(ex, ex, fmaf(ey, ey, ez * ez));
            if (GAP) gap = fminf(gap, fast_sqrt(e2) - r);
            return e2 <= r * r;
        }
        case MMMP_OBS_CYLINDER: {
            const float dx = cx - p[0], dy = cy - p[1], dz = cz - p[2];
            const float h = fmaf(p[3], dx, fmaf(p[4], dy, p[5] * dz));
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const float rad = sqrtf(fmaxf(d2 - h * h, 0.f));
            const float eh = fmaxf(fabsf(h) - p[6], 0.f);
            const float er = fmaxf(rad - p[7], 0.f);
            const float e2 = fmaf(eh, eh, er * er);
            if (GAP) gap = fminf(gap, fast_sqrt(e2) - r);
            return e2 <= r * r;
        }
    }
    return false;
}

__device__ __forceinline__ bool sphere_hits_obstacle(float cx, float cy, float cz, float r,
                                                     const ObstacleDev& O) {
    float unused = 0.f;
    return sphere_vs_obstacle<false>(cx, cy, cz, r, O, unused);
}
