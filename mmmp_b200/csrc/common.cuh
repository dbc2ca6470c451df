// Shared device/host definitions for the mmmp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include "../../include/mmmp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "mmmp_b200 kernels are written for sm_100a (B200) only"
#endif

// MMMP_DEBUG=1 in the environment: the first failing CUDA call of an entry point is named on stderr
static inline void mmmp_note_error(const char* file, int line, int code) {
    if (getenv("MMMP_DEBUG")) fprintf(stderr, "mmmp_b200: CUDA error %d (%s) at %s:%d\n", code,
                                      cudaGetErrorString((cudaError_t)code), file, line);
}
#define MMMP_CUDA_TRY(expr)                                            \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) {                                       \
            mmmp_note_error(__FILE__, __LINE__, (int)_e);              \
            return MMMP_ERR_CUDA_BASE + (int)_e;                       \
        }                                                              \
    } while (0)

// every kernel launch of the library goes through this so that bench.py can
// report gpu_launches from the library's own counter
extern std::atomic<int64_t> g_mmmp_launches;
#define MMMP_LAUNCH(kernel, grid, block, smem, stream, ...)                        \
    do {                                                                           \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);  \
        g_mmmp_launches.fetch_add(1, std::memory_order_relaxed);                   \
    } while (0)
#define MMMP_LAUNCH_CHECK()                                                        \
    do {                                                                           \
        cudaError_t _e = cudaPeekAtLastError();                                    \
        if (_e != cudaSuccess) {                                                   \
            mmmp_note_error(__FILE__, __LINE__, (int)_e);                          \
            return MMMP_ERR_CUDA_BASE + (int)_e;                                   \
        }                                                                          \
    } while (0)

// ---------------------------------------------------------------- device world
#define MMMP_MAX_SELF_PAIRS 120

struct PairTableDev {
    int joint_u, joint_v, n_u, n_v;
    float lo_u, inv_du, lo_v, inv_dv;
    int word_offset;  // into WorldDev::table_bits
    int pad[3];
};

struct RobotDev {
    int n_joints, n_spheres, n_links, dof_offset;
    int n_self_pairs, n_tables, base_hits_obstacles, n_moving_links;
    float base[12];
    float origin[MMMP_MAX_JOINTS + 1][12];
    float4 link_bound[MMMP_MAX_LINKS];  // bounding sphere, frame local (frame-0 links: WORLD)
    // bounding CAPSULE of every link's spheres (fitted at world creation, contains them all): segment
    // cap_a.xyz -- cap_b.xyz in the link's frame (frame-0 links: WORLD), radius cap_a.w (< 0: none).
    // Link pairs whose bounding spheres overlap are tested capsule against capsule before anybody
    // descends to the sphere level (the links are elongated: the capsules are 2-3x tighter).
    float4 cap_a[MMMP_MAX_LINKS], cap_b[MMMP_MAX_LINKS];
    int link_frame[MMMP_MAX_LINKS];              // non-decreasing in the link id
    int frame_link_begin[MMMP_MAX_JOINTS + 2];   // links of frame f: [begin[f], begin[f+1])
    int link_sphere_begin[MMMP_MAX_LINKS + 1];
    float reach[MMMP_MAX_JOINTS][MMMP_MAX_JOINTS + 1];  // [m][f]: bound of |d centre / d q_m|, frame f
    int origin_rx[MMMP_MAX_JOINTS + 1];  // 1: origin[j] is a rotation about x plus a translation (mod-DH rows)
    int pad_rx[3];
    // What a collision thread parks in its shared-memory column, per mode (0: self pairs only,
    // 1: every link, for robot-robot tests): slot of frame f / of link l's bounding-sphere
    // centre, -1 = not needed later, not stored.
    int n_frame_store[2], n_link_store[2];
    signed char frame_store[2][MMMP_MAX_JOINTS + 2];
    signed char link_store[2][MMMP_MAX_LINKS];
    signed char pad_store[4];
    uchar2 self_pair[MMMP_MAX_SELF_PAIRS];
    PairTableDev table[MMMP_MAX_PAIR_TABLES];
    float4 sphere[MMMP_MAX_SPHERES];    // frame local xyz, r (frame-0 links: WORLD xyz)
};

struct RobotDev64 {  // fp64 chain for the re-ranking FK
    int n_joints;
    double base[12];
    double origin[MMMP_MAX_JOINTS + 1][12];
};

struct ObstacleDev {   // p first and the struct 16-byte sized: a plane's (n, d) is one 16-byte shared-memory load
    float p[16];
    int type;
    int body;
    int pad[2];
};

struct WorldDev {
    int n_robots, n_obstacles, dof, n_table_words;
    int n_planes, pad_w[3];   // obstacles are stored planes first: [0, n_planes) take the plane fast path
    RobotDev robots[MMMP_MAX_ROBOTS];
    ObstacleDev obstacles[MMMP_MAX_OBSTACLES];
    RobotDev64 robots64[MMMP_MAX_ROBOTS];
    unsigned int table_bits[MMMP_MAX_ROBOTS * MMMP_MAX_PAIR_TABLES * MMMP_PAIR_TABLE_WORDS / 4];
};

struct PlanarDev {
    int n_arms, n_obstacles, dof, total_links;
    int n_links[MMMP_MAX_ROBOTS];
    int dof_offset[MMMP_MAX_ROBOTS];
    int link_begin[MMMP_MAX_ROBOTS + 1];
    double link_len[MMMP_MAX_ROBOTS][MMMP_MAX_JOINTS];
    double base[MMMP_MAX_ROBOTS][2];
    double bounds[4];
    int obs_type[MMMP_MAX_OBSTACLES];
    double obs[MMMP_MAX_OBSTACLES][4];
    double joint_link_radius;
};

enum { MMMP_WORLD_SPATIAL = 1, MMMP_WORLD_PLANAR = 2 };

struct mmmp_world {
    int kind;
    int device;
    WorldDev* spatial_d;   // device copy
    WorldDev* spatial_h;   // host mirror (pinned not required)
    PlanarDev* planar_d;
    PlanarDev* planar_h;
};

// a world's tables live on the device that was current when it was created; kernels launched on
// another device would fault on them
#define MMMP_CHECK_DEVICE(w)                                                        \
    do {                                                                           \
        int _cur = -1;                                                             \
        MMMP_CUDA_TRY(cudaGetDevice(&_cur));                                       \
        if ((w) && _cur != (w)->device) return MMMP_ERR_ARG;                       \
    } while (0)

#define MMMP_PLANAR_MAX_LINKS 32

// ---------------------------------------------------------------- affine helpers
struct Affine {  // row-major 3x4
    float m[12];
};

__device__ __forceinline__ void affine_load(Affine& T, const float* __restrict__ p) {
#pragma unroll
    for (int i = 0; i < 12; ++i) T.m[i] = p[i];
}

// T <- T * O * Rz(q)   (O a fixed 3x4 transform, then a rotation about the new z)
__device__ __forceinline__ void affine_step(Affine& T, const float* __restrict__ O, float s, float c) {
    float r[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = T.m[4 * i + 0], b = T.m[4 * i + 1], d = T.m[4 * i + 2];
        r[4 * i + 0] = fmaf(a, O[0], fmaf(b, O[4], d * O[8]));
        r[4 * i + 1] = fmaf(a, O[1], fmaf(b, O[5], d * O[9]));
        r[4 * i + 2] = fmaf(a, O[2], fmaf(b, O[6], d * O[10]));
        r[4 * i + 3] = fmaf(a, O[3], fmaf(b, O[7], fmaf(d, O[11], T.m[4 * i + 3])));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float x = r[4 * i + 0], y = r[4 * i + 1];
        T.m[4 * i + 0] = fmaf(x, c, y * s);
        T.m[4 * i + 1] = fmaf(y, c, -x * s);
        T.m[4 * i + 2] = r[4 * i + 2];
        T.m[4 * i + 3] = r[4 * i + 3];
    }
}

// Same for an origin of the form [1 0 0 a; 0 ca -sa y; 0 sa ca z] (every modified-DH row): the
// products with its exact zeros and its exact one are dropped, the result is bit-identical.
__device__ __forceinline__ void affine_step_rx(Affine& T, const float* __restrict__ O, float s, float c) {
    float r[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = T.m[4 * i + 0], b = T.m[4 * i + 1], d = T.m[4 * i + 2];
        r[4 * i + 0] = a;
        r[4 * i + 1] = fmaf(b, O[5], d * O[9]);
        r[4 * i + 2] = fmaf(b, O[6], d * O[10]);
        r[4 * i + 3] = fmaf(a, O[3], fmaf(b, O[7], fmaf(d, O[11], T.m[4 * i + 3])));
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float x = r[4 * i + 0], y = r[4 * i + 1];
        T.m[4 * i + 0] = fmaf(x, c, y * s);
        T.m[4 * i + 1] = fmaf(y, c, -x * s);
        T.m[4 * i + 2] = r[4 * i + 2];
        T.m[4 * i + 3] = r[4 * i + 3];
    }
}

// T <- T * O (fixed transform only, used for the tip frame)
__device__ __forceinline__ void affine_mul(Affine& T, const float* __restrict__ O) {
    float r[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = T.m[4 * i + 0], b = T.m[4 * i + 1], d = T.m[4 * i + 2];
        r[4 * i + 0] = fmaf(a, O[0], fmaf(b, O[4], d * O[8]));
        r[4 * i + 1] = fmaf(a, O[1], fmaf(b, O[5], d * O[9]));
        r[4 * i + 2] = fmaf(a, O[2], fmaf(b, O[6], d * O[10]));
        r[4 * i + 3] = fmaf(a, O[3], fmaf(b, O[7], fmaf(d, O[11], T.m[4 * i + 3])));
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) T.m[i] = r[i];
}

__device__ __forceinline__ float3 affine_point(const Affine& T, float x, float y, float z) {
    float3 o;
    o.x = fmaf(T.m[0], x, fmaf(T.m[1], y, fmaf(T.m[2], z, T.m[3])));
    o.y = fmaf(T.m[4], x, fmaf(T.m[5], y, fmaf(T.m[6], z, T.m[7])));
    o.z = fmaf(T.m[8], x, fmaf(T.m[9], y, fmaf(T.m[10], z, T.m[11])));
    return o;
}

static inline int mmmp_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// The library's temporaries come from the device's default stream-ordered pool.  By default
// that pool hands freed memory back to the driver at every synchronisation, which turns each
// host sync into a multi-hundred-megabyte unmap/remap; keep the memory cached instead.
static inline void mmmp_pool_keep(void) {
    static thread_local int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_for = dev;
}
