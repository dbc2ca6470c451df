// Kernel (b): batched forward kinematics, one configuration per thread, chain
// constants (base pose + per-joint fixed transforms) staged in shared memory.
//
// Reference: Panda.modified_homogeneous_transform / forward_kinematics /
// positions_from_fk robots/panda.py:111-173 (T = T_base * prod_i T_i(a,d,alpha,q_i),
// row 7 = fixed flange); the generic URDF chain pybullet evaluates behind
// resetJointState (utils/pb_joint_utils.py:76-88) for the KUKA iiwa.
#include "common.cuh"

namespace {

struct ChainSmem {
    int n_joints;
    float base[12];
    float origin[MMMP_MAX_JOINTS + 1][12];
};

// Positions leave through shared memory: a thread owns one configuration, i.e. 3 (n + 1) output
// floats at a stride of 3 (n + 1) floats -- written directly that is 32 scattered 4-byte stores per
// instruction (round 1: 1.07 TB/s, profiles/r02_fk_full.csv: lg_throttle / mio_throttle stalls).
// The 32 rows of a warp are contiguous in the output, so the warp parks them in a padded
// shared-memory tile (row stride 3 (n + 1) + 1: conflict free) and copies the tile out with
// consecutive lanes on consecutive addresses.  The configuration row comes in as two 16-byte loads.
constexpr int FK_WARPS = 8;
constexpr int FK_ROW_MAX = 3 * (MMMP_MAX_JOINTS + 1) + 1;

__global__ void __launch_bounds__(32 * FK_WARPS)
fk_chain_kernel(const WorldDev* __restrict__ W, int robot, const float* __restrict__ q, int ld, int64_t N,
                float* __restrict__ out_pos, float* __restrict__ out_T) {
    __shared__ ChainSmem C;
    __shared__ float stage[FK_WARPS][32 * FK_ROW_MAX];
    {
        const RobotDev& R = W->robots[robot];
        if (threadIdx.x == 0) C.n_joints = R.n_joints;
        for (int i = threadIdx.x; i < 12; i += blockDim.x) C.base[i] = R.base[i];
        for (int i = threadIdx.x; i < 12 * (MMMP_MAX_JOINTS + 1); i += blockDim.x)
            (&C.origin[0][0])[i] = (&R.origin[0][0])[i];
    }
    __syncthreads();
    const int n = C.n_joints;
    const int L = n + 1;
    const int row = 3 * L, srow = row + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = stage[warp];
    const bool vec = (ld & 3) == 0 && ((uintptr_t)q & 15) == 0;
    for (int64_t base = ((int64_t)blockIdx.x * FK_WARPS + warp) * 32; base < N; base += (int64_t)gridDim.x * FK_WARPS * 32) {
        const int64_t i = base + lane;
        if (i < N) {
            Affine T;
            affine_load(T, C.base);
            const float* qi = q + i * ld;
            float qv[MMMP_MAX_JOINTS];
            if (vec) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(qi));
                const float4 b = ld >= 8 ? __ldg(reinterpret_cast<const float4*>(qi) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                qv[0] = a.x, qv[1] = a.y, qv[2] = a.z, qv[3] = a.w, qv[4] = b.x, qv[5] = b.y, qv[6] = b.z, qv[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < MMMP_MAX_JOINTS; ++j) qv[j] = j < n ? __ldg(qi + j) : 0.f;
            }
            float* mine = tile + lane * srow;
#pragma unroll
            for (int j = 0; j < MMMP_MAX_JOINTS; ++j) {
                if (j < n) {
                    float s, c;
                    sincosf(qv[j], &s, &c);
                    affine_step(T, C.origin[j], s, c);
                    mine[3 * j + 0] = T.m[3];
                    mine[3 * j + 1] = T.m[7];
                    mine[3 * j + 2] = T.m[11];
                    if (out_T) {
                        float* t = out_T + (i * L + j) * 12;
#pragma unroll
                        for (int k = 0; k < 12; ++k) t[k] = T.m[k];
                    }
                }
            }
            affine_mul(T, C.origin[n]);
            mine[3 * n + 0] = T.m[3];
            mine[3 * n + 1] = T.m[7];
            mine[3 * n + 2] = T.m[11];
            if (out_T) {
                float* t = out_T + (i * L + n) * 12;
#pragma unroll
                for (int k = 0; k < 12; ++k) t[k] = T.m[k];
            }
        }
        __syncwarp();
        const int rows = (int)((N - base) < 32 ? (N - base) : 32);
        float* dst = out_pos + base * row;
        for (int e = lane; e < rows * row; e += 32) {
            const int r = e / row;
            dst[e] = tile[r * srow + (e - r * row)];
        }
        __syncwarp();
    }
}

struct FrameSel {
    int n;
    int frame[MMMP_MAX_GROUPS];  // 1-based frame numbers (tip = n_joints + 1)
    float weight[MMMP_MAX_GROUPS];
};

// FK-metric feature rows: xyz of the selected frames, zero padded to ldf; and the filter
// row of the neighbour scan: the weighted centroid sum_g w_g p_g (xyz, 0); dual rows: see
// mmmp_fk_features in include/mmmp_b200.h
__global__ void __launch_bounds__(256)
fk_features_kernel(const WorldDev* __restrict__ W, int robot, const float* __restrict__ q, int ld, int64_t N,
                   FrameSel sel, float* __restrict__ out, int ldf, float4* __restrict__ out_filter, int dual) {
    __shared__ ChainSmem C;
    {
        const RobotDev& R = W->robots[robot];
        if (threadIdx.x == 0) C.n_joints = R.n_joints;
        for (int i = threadIdx.x; i < 12; i += blockDim.x) C.base[i] = R.base[i];
        for (int i = threadIdx.x; i < 12 * (MMMP_MAX_JOINTS + 1); i += blockDim.x)
            (&C.origin[0][0])[i] = (&R.origin[0][0])[i];
    }
    __syncthreads();
    const int n = C.n_joints;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        Affine T;
        affine_load(T, C.base);
        const float* qi = q + i * ld;
        float* o = out + i * ldf;
        float cx = 0.f, cy = 0.f, cz = 0.f;   // weighted sum over ALL groups
        float ax = 0.f, ay = 0.f, az = 0.f;   // ... over the first floor(n / 2) groups
        for (int j = 0; j <= n; ++j) {
            if (j < n) {
                float s, c;
                sincosf(qi[j], &s, &c);
                affine_step(T, C.origin[j], s, c);
            } else {
                affine_mul(T, C.origin[n]);
            }
            for (int g = 0; g < sel.n; ++g)
                if (sel.frame[g] == j + 1) {
                    o[3 * g + 0] = T.m[3];
                    o[3 * g + 1] = T.m[7];
                    o[3 * g + 2] = T.m[11];
                    cx = fmaf(sel.weight[g], T.m[3], cx);
                    cy = fmaf(sel.weight[g], T.m[7], cy);
                    cz = fmaf(sel.weight[g], T.m[11], cz);
                    if (g < sel.n / 2) {
                        ax = fmaf(sel.weight[g], T.m[3], ax);
                        ay = fmaf(sel.weight[g], T.m[7], ay);
                        az = fmaf(sel.weight[g], T.m[11], az);
                    }
                }
        }
        for (int c = 3 * sel.n; c < ldf; ++c) o[c] = 0.f;
        if (out_filter && !dual) out_filter[i] = make_float4(cx, cy, cz, 0.f);
        if (out_filter && dual) {  // C = A + B (the centroid), D = A - B = 2 A - C
            const float dx = 2.f * ax - cx, dy = 2.f * ay - cy, dz = 2.f * az - cz;
            out_filter[2 * i] = make_float4(cx, cy, cz, dx);
            out_filter[2 * i + 1] = make_float4(dy, dz, 0.f, 0.f);
        }
    }
}

// fp64 twin, non-fused arithmetic in the reference's matmul order
// (np.matmul(T, T_i) row by column, k ascending): used for the fp64 re-rank of
// neighbour candidates, where the metric must follow robots/panda.py:248-254.
__global__ void __launch_bounds__(128)
fk_chain_f64_kernel(const WorldDev* __restrict__ W, int robot, const double* __restrict__ q, int ld, int64_t N,
                    double* __restrict__ out_pos) {
    __shared__ RobotDev64 C;
    for (int i = threadIdx.x; i < (int)(sizeof(RobotDev64) / 4); i += blockDim.x)
        reinterpret_cast<int*>(&C)[i] = reinterpret_cast<const int*>(&W->robots64[robot])[i];
    __syncthreads();
    const int n = C.n_joints;
    const int L = n + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        double T[12];
        for (int k = 0; k < 12; ++k) T[k] = C.base[k];
        const double* qi = q + i * ld;
        for (int j = 0; j <= n; ++j) {
            // M = O * Rz(q): the reference builds T_i in one go (panda.py:111-122);
            // its entries are products of two factors, so O*Rz with exact 0/1
            // entries of Rz reproduces them up to the rounding of ct*ca etc.
            const double* O = C.origin[j];
            double s = 0.0, c = 1.0;
            if (j < n) sincos(qi[j], &s, &c);
            double M[12];
            for (int r = 0; r < 3; ++r) {
                M[4 * r + 0] = __dadd_rn(__dmul_rn(O[4 * r + 0], c), __dmul_rn(O[4 * r + 1], s));
                M[4 * r + 1] = __dsub_rn(__dmul_rn(O[4 * r + 1], c), __dmul_rn(O[4 * r + 0], s));
                M[4 * r + 2] = O[4 * r + 2];
                M[4 * r + 3] = O[4 * r + 3];
            }
            double Rr[12];
            for (int r = 0; r < 3; ++r) {
                for (int cc = 0; cc < 4; ++cc) {
                    double acc = __dmul_rn(T[4 * r + 0], M[cc]);
                    acc = __dadd_rn(acc, __dmul_rn(T[4 * r + 1], M[4 + cc]));
                    acc = __dadd_rn(acc, __dmul_rn(T[4 * r + 2], M[8 + cc]));
                    if (cc == 3) acc = __dadd_rn(acc, T[4 * r + 3]);
                    Rr[4 * r + cc] = acc;
                }
            }
            for (int k = 0; k < 12; ++k) T[k] = Rr[k];
            double* p = out_pos + (i * L + j) * 3;
            p[0] = T[3];
            p[1] = T[7];
            p[2] = T[11];
        }
    }
}

int fk_grid(int64_t N, int block) {
    int64_t g = (N + block - 1) / block;
    const int64_t cap = 148 * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int mmmp_fk_chain(const mmmp_world* w, int robot, const float* q, int ld, int64_t N, float* out_pos,
                             float* out_T, void* stream) {
    if (!w || !q || !out_pos || N < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    if (robot < 0 || robot >= w->spatial_h->n_robots || ld < w->spatial_h->robots[robot].n_joints) return MMMP_ERR_ARG;
    if (N == 0) return MMMP_OK;
    MMMP_LAUNCH(fk_chain_kernel, fk_grid(N, 256), 256, 0, stream, w->spatial_d, robot, q, ld, N, out_pos, out_T);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_fk_chain_f64(const mmmp_world* w, int robot, const double* q, int ld, int64_t N,
                                 double* out_pos, void* stream) {
    if (!w || !q || !out_pos || N < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    if (robot < 0 || robot >= w->spatial_h->n_robots || ld < w->spatial_h->robots[robot].n_joints) return MMMP_ERR_ARG;
    if (N == 0) return MMMP_OK;
    MMMP_LAUNCH(fk_chain_f64_kernel, fk_grid(N, 128), 128, 0, stream, w->spatial_d, robot, q, ld, N, out_pos);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_fk_features(const mmmp_world* w, int robot, const float* q, int ld, int64_t N,
                                const int32_t* frames, const float* weights, int n_frames, float* out, int ldf,
                                float* out_filter, int ldfilt, void* stream) {
    if (!w || !q || !out || !frames || N < 0) return MMMP_ERR_ARG;
    MMMP_CHECK_DEVICE(w);
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    if (robot < 0 || robot >= w->spatial_h->n_robots || ld < w->spatial_h->robots[robot].n_joints) return MMMP_ERR_ARG;
    if (n_frames < 1 || n_frames > MMMP_MAX_GROUPS || ldf < 3 * n_frames) return MMMP_ERR_ARG;
    FrameSel sel;
    sel.n = n_frames;
    for (int g = 0; g < MMMP_MAX_GROUPS; ++g) {
        sel.frame[g] = -1;
        sel.weight[g] = 0.f;
    }
    for (int g = 0; g < n_frames; ++g) {
        if (frames[g] < 1 || frames[g] > w->spatial_h->robots[robot].n_joints + 1) return MMMP_ERR_ARG;
        sel.frame[g] = frames[g];
        sel.weight[g] = weights ? weights[g] : 1.f;
    }
    if (out_filter && ((((uintptr_t)out_filter) & 15) || (ldfilt != 4 && ldfilt != 8))) return MMMP_ERR_ARG;
    if (N == 0) return MMMP_OK;
    MMMP_LAUNCH(fk_features_kernel, fk_grid(N, 256), 256, 0, stream, w->spatial_d, robot, q, ld, N, sel, out, ldf,
                reinterpret_cast<float4*>(out_filter), ldfilt == 8 ? 1 : 0);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}
