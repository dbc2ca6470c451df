// Sphere-vs-primitive overlap tests (touching counts as collision, matching
// pybullet getClosestPoints(distance=0) returning a point at distance <= 0:
// reference planners/prm/pybullet/env.py:65-67).
#pragma once
#include "common.cuh"

__device__ __forceinline__ bool sphere_hits_obstacle(float cx, float cy, float cz, float r,
                                                     const ObstacleDev& O) {
    const float* p = O.p;
    switch (O.type) {
        case MMMP_OBS_PLANE: {
            const float sd = fmaf(p[0], cx, fmaf(p[1], cy, p[2] * cz)) - p[3];
            return sd <= r;
        }
        case MMMP_OBS_SPHERE: {
            const float dx = cx - p[0], dy = cy - p[1], dz = cz - p[2];
            const float rr = r + p[3];
            return fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= rr * rr;
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
            return fmaf(ex, ex, fmaf(ey, ey, ez * ez)) <= r * r;
        }
        case MMMP_OBS_CYLINDER: {
            const float dx = cx - p[0], dy = cy - p[1], dz = cz - p[2];
            const float h = fmaf(p[3], dx, fmaf(p[4], dy, p[5] * dz));
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const float rad = sqrtf(fmaxf(d2 - h * h, 0.f));
            const float eh = fmaxf(fabsf(h) - p[6], 0.f);
            const float er = fmaxf(rad - p[7], 0.f);
            return fmaf(eh, eh, er * er) <= r * r;
        }
    }
    return false;
}
