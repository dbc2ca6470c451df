// Batched damped-least-squares inverse kinematics of a serial chain.
//
// Replaces (reference): Panda.solve_closest_arm_ik robots/panda.py:207-208 ->
// utils/ik_utils.py:25-27 -> pybullet.calculateInverseKinematics, as used one pose at a
// time by DecoupledPRM.sample_valid_in_task_space decoupled_prm.py:323-344.  pybullet's
// solver is a damped-least-squares iteration on the geometric Jacobian started from the
// body's current joint state; its arithmetic lives in Bullet (absent here), so this kernel
// restates the published method and PARITY WITH PYBULLET IS UNPINNED; it is pinned against
// the float64 restatement oracle/ik.py.
//
// One problem per thread, fp32 (the sampler rounds the answer to 0.01 rad).  Iteration:
//   FK with the joint axes a_j and joint positions p_j kept in registers; tool pose
//   T = frame_n * origin[n];  e = [p* - p ; 0.5 (n x n* + o x o* + a x a*)];
//   J_j = [a_j x (p - p_j) ; a_j];  dq = J^T (J J^T + lambda^2 I)^-1 e  (6x6 Cholesky);
//   |dq|_inf limited to max_step;  q <- clamp(q + dq, lo, hi);
// until |e_p| < tol_pos and |e_r| < tol_rot or max_iters.  target_quat == NULL solves for
// the position only (the orientation rows vanish).
#include "common.cuh"

namespace {

struct IkLimits {
    float lo[MMMP_MAX_JOINTS], hi[MMMP_MAX_JOINTS];
};

__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

__global__ void __launch_bounds__(128)
ik_dls_kernel(const WorldDev* __restrict__ W, int robot, const double* __restrict__ tpos, const double* __restrict__ tquat,
              const double* __restrict__ q0, int ld0, int64_t N, int max_iters, float lambda2, float tol_pos,
              float tol_rot, float max_step, IkLimits lim, double* __restrict__ q_out, int ld,
              uint8_t* __restrict__ converged, float* __restrict__ residual) {
    __shared__ float s_origin[MMMP_MAX_JOINTS + 1][12];
    __shared__ float s_base[12];
    const RobotDev& Rg = W->robots[robot];
    const int n = Rg.n_joints;
    for (int i = threadIdx.x; i < (n + 1) * 12; i += blockDim.x) s_origin[i / 12][i % 12] = Rg.origin[i / 12][i % 12];
    if (threadIdx.x < 12) s_base[threadIdx.x] = Rg.base[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        float q[MMMP_MAX_JOINTS];
#pragma unroll
        for (int j = 0; j < MMMP_MAX_JOINTS; ++j) q[j] = j < n ? (float)q0[(ld0 ? i * ld0 : 0) + j] : 0.f;
        const float tp[3] = {(float)tpos[3 * i], (float)tpos[3 * i + 1], (float)tpos[3 * i + 2]};
        float Rt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // target rotation, row major
        const bool with_rot = tquat != nullptr;
        if (with_rot) {
            const float x = (float)tquat[4 * i], y = (float)tquat[4 * i + 1], z = (float)tquat[4 * i + 2],
                        w = (float)tquat[4 * i + 3];
            const float s = 2.f / fmaxf(x * x + y * y + z * z + w * w, 1e-30f);
            Rt[0] = 1.f - s * (y * y + z * z); Rt[1] = s * (x * y - z * w);       Rt[2] = s * (x * z + y * w);
            Rt[3] = s * (x * y + z * w);       Rt[4] = 1.f - s * (x * x + z * z); Rt[5] = s * (y * z - x * w);
            Rt[6] = s * (x * z - y * w);       Rt[7] = s * (y * z + x * w);       Rt[8] = 1.f - s * (x * x + y * y);
        }
        bool done = false;
        float ep_n = 0.f, er_n = 0.f;
        for (int it = 0; it <= max_iters; ++it) {
            // ---- FK, joint axes and positions
            float ax[MMMP_MAX_JOINTS][3], pj[MMMP_MAX_JOINTS][3];
            Affine T;
            affine_load(T, s_base);
#pragma unroll
            for (int j = 0; j < MMMP_MAX_JOINTS; ++j) {
                if (j < n) {
                    float s, c;
                    sincosf(q[j], &s, &c);
                    affine_step(T, s_origin[j], s, c);
                    ax[j][0] = T.m[2]; ax[j][1] = T.m[6]; ax[j][2] = T.m[10];
                    pj[j][0] = T.m[3]; pj[j][1] = T.m[7]; pj[j][2] = T.m[11];
                }
            }
            affine_mul(T, s_origin[n]);
            // ---- task error
            float e[6];
            e[0] = tp[0] - T.m[3]; e[1] = tp[1] - T.m[7]; e[2] = tp[2] - T.m[11];
            e[3] = e[4] = e[5] = 0.f;
            if (with_rot) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {  // 0.5 * sum over columns of (current x target)
                    const float cur[3] = {T.m[c], T.m[4 + c], T.m[8 + c]}, tar[3] = {Rt[c], Rt[3 + c], Rt[6 + c]};
                    float x[3];
                    cross3(cur, tar, x);
                    e[3] += 0.5f * x[0]; e[4] += 0.5f * x[1]; e[5] += 0.5f * x[2];
                }
            }
            ep_n = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
            er_n = sqrtf(e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
            if (ep_n < tol_pos && er_n < tol_rot) {
                done = true;
                break;
            }
            if (it == max_iters) break;
            // ---- Jacobian and A = J J^T + lambda^2 I
            float J[MMMP_MAX_JOINTS][6];
            float A[6][6];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) A[r][c] = r == c ? lambda2 : 0.f;
#pragma unroll
            for (int j = 0; j < MMMP_MAX_JOINTS; ++j) {
                if (j < n) {
                    const float d[3] = {T.m[3] - pj[j][0], T.m[7] - pj[j][1], T.m[11] - pj[j][2]};
                    cross3(ax[j], d, J[j]);
                    J[j][3] = with_rot ? ax[j][0] : 0.f;
                    J[j][4] = with_rot ? ax[j][1] : 0.f;
                    J[j][5] = with_rot ? ax[j][2] : 0.f;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c <= r; ++c) A[r][c] = fmaf(J[j][r], J[j][c], A[r][c]);
                }
            }
            // ---- Cholesky A = L L^T (lower triangle in place), solve L L^T y = e
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float dgl = A[c][c];
#pragma unroll
                for (int k = 0; k < c; ++k) dgl -= A[c][k] * A[c][k];
                const float inv = rsqrtf(fmaxf(dgl, 1e-20f));
                A[c][c] = dgl * inv;
#pragma unroll
                for (int r = c + 1; r < 6; ++r) {
                    float v = A[r][c];
#pragma unroll
                    for (int k = 0; k < c; ++k) v -= A[r][k] * A[c][k];
                    A[r][c] = v * inv;
                }
            }
            float y[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                float v = e[r];
#pragma unroll
                for (int k = 0; k < r; ++k) v -= A[r][k] * y[k];
                y[r] = v / fmaxf(A[r][r], 1e-10f);
            }
#pragma unroll
            for (int r = 5; r >= 0; --r) {
                float v = y[r];
#pragma unroll
                for (int k = r + 1; k < 6; ++k) v -= A[k][r] * y[k];
                y[r] = v / fmaxf(A[r][r], 1e-10f);
            }
            // ---- step
            float dq[MMMP_MAX_JOINTS], big = 0.f;
#pragma unroll
            for (int j = 0; j < MMMP_MAX_JOINTS; ++j) {
                dq[j] = 0.f;
                if (j < n) {
#pragma unroll
                    for (int r = 0; r < 6; ++r) dq[j] = fmaf(J[j][r], y[r], dq[j]);
                    big = fmaxf(big, fabsf(dq[j]));
                }
            }
            const float scale = big > max_step ? max_step / big : 1.f;
#pragma unroll
            for (int j = 0; j < MMMP_MAX_JOINTS; ++j)
                if (j < n) q[j] = fminf(fmaxf(fmaf(scale, dq[j], q[j]), lim.lo[j]), lim.hi[j]);
        }
#pragma unroll
        for (int j = 0; j < MMMP_MAX_JOINTS; ++j)
            if (j < n) q_out[i * ld + j] = (double)q[j];
        if (converged) converged[i] = done ? 1 : 0;
        if (residual) {
            residual[2 * i] = ep_n;
            residual[2 * i + 1] = er_n;
        }
    }
}

}  // namespace

extern "C" int mmmp_ik_dls(const mmmp_world* w, int robot, const double* target_pos, const double* target_quat,
                           const double* q0, int ld0, int64_t N, const double* lo_h, const double* hi_h, int max_iters,
                           double lambda, double tol_pos, double tol_rot, double max_step, double* q_out, int ld,
                           uint8_t* converged, float* residual, void* stream) {
    if (!w || !target_pos || !q0 || !q_out || !lo_h || !hi_h || N < 0 || max_iters < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    if (robot < 0 || robot >= w->spatial_h->n_robots) return MMMP_ERR_ARG;
    const int n = w->spatial_h->robots[robot].n_joints;
    if (ld < n || (ld0 != 0 && ld0 < n) || !(lambda >= 0.0) || !(max_step > 0.0)) return MMMP_ERR_ARG;
    if (N == 0) return MMMP_OK;
    IkLimits lim;
    for (int j = 0; j < MMMP_MAX_JOINTS; ++j) {
        lim.lo[j] = j < n ? (float)lo_h[j] : 0.f;
        lim.hi[j] = j < n ? (float)hi_h[j] : 0.f;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (N + 127) / 128;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    MMMP_LAUNCH(ik_dls_kernel, (int)grid, 128, 0, stream, w->spatial_d, robot, target_pos, target_quat, q0, ld0, N,
                max_iters, (float)(lambda * lambda), (float)tol_pos, (float)tol_rot, (float)max_step, lim, q_out, ld,
                converged, residual);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}
