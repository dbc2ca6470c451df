// Kernel (c): configuration and discretised-edge collision checks for spatial
// (Panda / KUKA) worlds.  One configuration per thread: FK of the serial chain
// with link transforms staged in shared memory; the thread parks its 3x4 frame
// poses in a private shared-memory column (conflict-free, lane-contiguous).
// Every collision link carries a sphere decomposition fitted offline plus a
// bounding sphere: obstacle, self and robot-robot tests first cull on the
// bounding spheres and only then descend to the sphere level, where one link's
// spheres are mapped into the OTHER link's frame so the inner loop reads its
// constants as shared-memory broadcasts.  Link pairs whose relative pose depends
// on two joints only may be decided by an exact offline bit table instead.
// Edge checks map one warp to one edge (lanes = interpolation steps) with
// warp-ballot early exit.
//
// Replaces (reference): Environment.robot_self_collision / robot_obstacle_collision /
// robot_robot_collision planners/prm/pybullet/env.py:51-84 as composed by
// DecoupledPRM.is_collision_free_sample / _edge decoupled_prm.py:840-863 and
// CoupledPRM.is_collision_free_sample / _edge coupled/coupled_prm.py:143-187.
#include "common.cuh"
#include "obstacle.cuh"

namespace {

struct SpatialSmem {
    int n_r;
    int n_obs;
    int col[MMMP_MAX_ROBOTS];          // first column of local robot lr in a q row
    int frame_slot[MMMP_MAX_ROBOTS];   // first frame slot of local robot lr
};

struct SpatialArgs {
    const WorldDev* W;
    int n_r;
    int robots[MMMP_MAX_ROBOTS];
    int col[MMMP_MAX_ROBOTS];
    uint32_t flags;
    int slots;  // frame slots (12 floats each) per thread
};

// --- q loaders -------------------------------------------------------------
struct QConfig {
    const float* row;
    __device__ __forceinline__ float get(int lr, int c) const { return __ldg(row + c); }
};
struct QEdge {
    const float* a;
    const float* b;
    float t;
    __device__ __forceinline__ float get(int lr, int c) const {
        const float qa = __ldg(a + c), qb = __ldg(b + c);
        return fmaf(t, qb - qa, qa);
    }
};
struct QPair {
    const float* ra;
    const float* rb;
    __device__ __forceinline__ float get(int lr, int c) const { return __ldg((lr == 0 ? ra : rb) + c); }
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
        int slot = 0;
        for (int lr = 0; lr < A.n_r; ++lr) {
            hdr->col[lr] = A.col[lr];
            hdr->frame_slot[lr] = slot;
            slot += A.W->robots[A.robots[lr]].n_joints;
        }
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

// frame f (1..n) of a robot from the thread's shared-memory column; f == 0 is the
// identity because frame-0 (base link) spheres are stored in world coordinates
__device__ __forceinline__ void load_frame(Affine& T, const float* fr, int stride, int slot0, int f) {
    if (f == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) T.m[i] = (i == 0 || i == 5 || i == 10) ? 1.f : 0.f;
        return;
    }
    const float* p = fr + (size_t)(slot0 + f - 1) * 12 * stride;
#pragma unroll
    for (int i = 0; i < 12; ++i) T.m[i] = p[i * stride];
}

// spheres of link lb (robot RB, pose TB) against spheres of link la (robot RA, pose TA)
__device__ __forceinline__ bool link_pair_hit(const RobotDev& RA, int la, const Affine& TA, const RobotDev& RB, int lb,
                                              const Affine& TB) {
    const float4 ba = RA.link_bound[la], bb = RB.link_bound[lb];
    const float3 wa = affine_point(TA, ba.x, ba.y, ba.z);
    const float3 wb = affine_point(TB, bb.x, bb.y, bb.z);
    {
        const float dx = wa.x - wb.x, dy = wa.y - wb.y, dz = wa.z - wb.z;
        const float rr = ba.w + bb.w;
        if (fmaf(dx, dx, fmaf(dy, dy, dz * dz)) > rr * rr) return false;
    }
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
            if (fmaf(dx, dx, fmaf(dy, dy, dz * dz)) > rr * rr) continue;
        }
        for (int sa = a0; sa < a1; ++sa) {
            const float4 P = RA.sphere[sa];
            const float dx = c.x - P.x, dy = c.y - P.y, dz = c.z - P.z;
            const float rr = S.w + P.w;
            hit |= fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= rr * rr;
        }
        if (hit) return true;
    }
    return false;
}

template <class QL>
__device__ bool spatial_in_collision(const WorldDev* __restrict__ Wg, const SpatialSmem* hdr, const RobotDev* robs,
                                     const ObstacleDev* obs, uint32_t flags, float* fr, int stride, const QL& ql) {
    const int n_r = hdr->n_r;
    const int n_obs = (flags & MMMP_CHECK_OBSTACLES) ? hdr->n_obs : 0;
    for (int lr = 0; lr < n_r; ++lr) {
        const RobotDev& R = robs[lr];
        const int col = hdr->col[lr];
        if (n_obs > 0 && R.base_hits_obstacles) return true;
        Affine T;
        affine_load(T, R.base);
        const int slot0 = hdr->frame_slot[lr];
        bool hit = false;
        float qv[MMMP_MAX_JOINTS];
#pragma unroll
        for (int j = 0; j < MMMP_MAX_JOINTS; ++j) qv[j] = 0.f;
#pragma unroll 1
        for (int j = 0; j < R.n_joints; ++j) {
            const float qj = ql.get(lr, col + j);
#pragma unroll
            for (int jj = 0; jj < MMMP_MAX_JOINTS; ++jj)
                if (jj == j) qv[jj] = qj;
            float s, c;
            sincosf(qj, &s, &c);
            affine_step(T, R.origin[j], s, c);
            float* p = fr + (size_t)(slot0 + j) * 12 * stride;
#pragma unroll
            for (int i = 0; i < 12; ++i) p[i * stride] = T.m[i];
            if (n_obs > 0) {
                // links hanging off frame j+1: bounding sphere first, spheres on demand
                for (int l = R.frame_link_begin[j + 1]; l < R.frame_link_begin[j + 2]; ++l) {
                    const float4 B = R.link_bound[l];
                    if (B.w < 0.f) continue;
                    const float3 wb = affine_point(T, B.x, B.y, B.z);
                    for (int o = 0; o < n_obs; ++o) {
                        if (!sphere_hits_obstacle(wb.x, wb.y, wb.z, B.w, obs[o])) continue;
                        for (int sp = R.link_sphere_begin[l]; sp < R.link_sphere_begin[l + 1]; ++sp) {
                            const float4 S = R.sphere[sp];
                            const float3 w = affine_point(T, S.x, S.y, S.z);
                            hit |= sphere_hits_obstacle(w.x, w.y, w.z, S.w, obs[o]);
                        }
                        if (hit) return true;
                    }
                }
            }
        }
        if (flags & MMMP_CHECK_SELF) {
            for (int t = 0; t < R.n_tables; ++t) {
                const PairTableDev& P = R.table[t];
                float qu = 0.f, qw = 0.f;
#pragma unroll
                for (int jj = 0; jj < MMMP_MAX_JOINTS; ++jj) {
                    if (jj == P.joint_u) qu = qv[jj];
                    if (jj == P.joint_v) qw = qv[jj];
                }
                int iu = (int)floorf((qu - P.lo_u) * P.inv_du);
                int iv = (int)floorf((qw - P.lo_v) * P.inv_dv);
                iu = min(max(iu, 0), P.n_u - 1);
                iv = min(max(iv, 0), P.n_v - 1);
                const int cell = iu * P.n_v + iv;
                if ((__ldg(Wg->table_bits + P.word_offset + (cell >> 5)) >> (cell & 31)) & 1u) return true;
            }
            for (int pi = 0; pi < R.n_self_pairs; ++pi) {
                const uchar2 pr = R.self_pair[pi];
                Affine TA, TB;
                load_frame(TA, fr, stride, slot0, R.link_frame[pr.x]);
                load_frame(TB, fr, stride, slot0, R.link_frame[pr.y]);
                if (link_pair_hit(R, pr.x, TA, R, pr.y, TB)) return true;
            }
        }
    }
    if ((flags & MMMP_CHECK_ROBOTS) && n_r > 1) {
        for (int l1 = 0; l1 < n_r; ++l1) {
            const RobotDev& R1 = robs[l1];
            for (int l2 = l1 + 1; l2 < n_r; ++l2) {
                const RobotDev& R2 = robs[l2];
                for (int la = 0; la < R1.n_links; ++la) {
                    if (R1.link_bound[la].w < 0.f) continue;
                    Affine TA;
                    load_frame(TA, fr, stride, hdr->frame_slot[l1], R1.link_frame[la]);
                    for (int lb = 0; lb < R2.n_links; ++lb) {
                        if (R2.link_bound[lb].w < 0.f) continue;
                        if (R1.link_frame[la] == 0 && R2.link_frame[lb] == 0) continue;  // static pair
                        Affine TB;
                        load_frame(TB, fr, stride, hdr->frame_slot[l2], R2.link_frame[lb]);
                        if (link_pair_hit(R1, la, TA, R2, lb, TB)) return true;
                    }
                }
            }
        }
    }
    return false;
}

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(128)
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

__global__ void __launch_bounds__(128)
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

// Edge checks.  A warp is split into groups of L lanes; one group = one edge, lane g of the
// group takes steps g*rounds + r in round r (every round spreads over the whole edge).  After
// each round a warp ballot tells every group whether one of its steps collided: that edge is
// decided (early exit) and its lanes idle until all groups of the warp are done.  L is chosen
// on the host so that rounds * L wastes the fewest lanes (50 steps: L = 10 -> 3 edges per
// warp, 5 rounds, 94 % lane use; consecutive edges of a k-NN table share their first node,
// so the groups of a warp see similar configurations).  Verdict = OR over all steps =
// reference semantics (decoupled_prm.py:855-863).
__global__ void __launch_bounds__(128)
collide_edges_kernel(SpatialArgs A, const float* __restrict__ qa_base, const float* __restrict__ qb_base, int ld,
                     const int32_t* __restrict__ edges, int64_t E, int n_steps, double step, int L,
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
    const int epw = 32 / L;                 // edges per warp
    const int group = lane / L, gl = lane - group * L;
    const bool in_group = group < epw;
    const int rounds = (n_steps + L - 1) / L;
    const unsigned gmask = in_group ? (((L == 32) ? 0xffffffffu : ((1u << L) - 1u)) << (group * L)) : 0u;
    const int64_t warp_id = (int64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;
    for (int64_t e0 = warp_id * epw; e0 < E; e0 += n_warps * epw) {
        const int64_t e = e0 + group;
        bool valid = in_group && e < E;
        const float *pa = qa_base, *pb = qa_base;
        bool blocked = false;   // empty neighbour slot: reported as blocked
        if (valid) {
            if (edges) {
                const int ia = __ldg(edges + 2 * e), ib = __ldg(edges + 2 * e + 1);
                blocked = ia < 0 || ib < 0;
                pa = qa_base + (int64_t)(blocked ? 0 : ia) * ld;
                pb = qa_base + (int64_t)(blocked ? 0 : ib) * ld;
            } else {
                pa = qa_base + e * ld;
                pb = qb_base + e * ld;
            }
        }
        bool done = blocked || !valid;
        bool collided = blocked;
        for (int r = 0; r < rounds; ++r) {
            const int i = gl * rounds + r;
            bool hit = false;
            if (!done && i < n_steps) {
                QEdge ql{pa, pb, (float)((double)i * step)};
                hit = spatial_in_collision(A.W, hdr, robs, obs, A.flags, cs, stride, ql);
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            if (ballot & gmask) {
                collided = true;
                done = true;
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
        if (valid && gl == 0) out[e] = collided ? 1 : 0;
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
        A.slots = H->robots[robot].n_joints;
    } else {
        A.n_r = H->n_robots;
        A.slots = 0;
        for (int r = 0; r < H->n_robots; ++r) {
            A.robots[r] = r;
            A.col[r] = H->robots[r].dof_offset;
            A.slots += H->robots[r].n_joints;
        }
    }
    return MMMP_OK;
}

size_t smem_bytes(const mmmp_world* w, const SpatialArgs& A, int block) {
    size_t off = (sizeof(SpatialSmem) + 15) & ~(size_t)15;
    off += (sizeof(RobotDev) * A.n_r + 15) & ~(size_t)15;
    off += (sizeof(ObstacleDev) * w->spatial_h->n_obstacles + 15) & ~(size_t)15;
    off += (size_t)A.slots * 12 * sizeof(float) * block;
    return off;
}

template <class K>
int pick_launch(K kernel, const mmmp_world* w, const SpatialArgs& A, int64_t work_items_per_cta_hint,
                int64_t total_items, int& block, size_t& smem, int& grid) {
    int dev = 0, sms = 0, max_smem = 0;
    MMMP_CUDA_TRY(cudaGetDevice(&dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    block = 128;
    smem = smem_bytes(w, A, block);
    while (smem > (size_t)max_smem && block > 32) {
        block >>= 1;
        smem = smem_bytes(w, A, block);
    }
    if (smem > (size_t)max_smem) return MMMP_ERR_LIMIT;
    MMMP_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MMMP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
    if (per_sm < 1) per_sm = 1;
    const int64_t per_cta = work_items_per_cta_hint * (block / 128 > 0 ? 1 : 1);
    int64_t want = (total_items + per_cta - 1) / per_cta;
    const int64_t resident = (int64_t)sms * per_sm;
    // persistent grid: a whole number of waves, never more CTAs than work
    grid = (int)(want < resident ? (want < 1 ? 1 : want) : resident);
    return MMMP_OK;
}

}  // namespace

extern "C" int mmmp_planar_collide_configs(const mmmp_world* w, int arm, const double* q, int ld, int64_t N,
                                           uint32_t flags, uint8_t* out, void* stream);
extern "C" int mmmp_planar_collide_edges(const mmmp_world* w, int arm, const double* qa, const double* qb,
                                         int ld, const int32_t* edges, int64_t E, int n_steps, double step,
                                         uint32_t flags, uint8_t* out, void* stream);

extern "C" int mmmp_collide_configs(const mmmp_world* w, int robot, const void* q, int ld, int64_t N,
                                    uint32_t flags, uint8_t* out, void* stream) {
    if (!w || !q || !out || N < 0) return MMMP_ERR_ARG;
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
    A.slots = H->robots[ra].n_joints + H->robots[rb].n_joints;
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
                        uint8_t* out, void* stream) {
    if (n_steps < 1 || n_steps > 4096) return MMMP_ERR_ARG;
    SpatialArgs A;
    int rc = fill_args(w, robot, flags, A);
    if (rc) return rc;
    int block, grid;
    size_t smem;
    rc = pick_launch(collide_edges_kernel, w, A, 4, E, block, smem, grid);
    if (rc) return rc;
    // lanes per edge: minimise rounds / (edges per warp)
    int L = 32;
    {
        const int cand[6] = {32, 16, 10, 8, 5, 4};
        double best = 1e30;
        for (int c = 0; c < 6; ++c) {
            const int l = cand[c];
            const double cost = (double)((n_steps + l - 1) / l) / (double)(32 / l);
            if (cost < best - 1e-12) {
                best = cost;
                L = l;
            }
        }
    }
    const int wpc = block / 32 * (32 / L);  // edges per CTA pass
    grid = mmmp_div_up(E, wpc) < grid ? mmmp_div_up(E, wpc) : grid;
    MMMP_LAUNCH(collide_edges_kernel, grid, block, smem, stream, A, qa, qb, ld, edges, E, n_steps, step, L, out);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_collide_edges(const mmmp_world* w, int robot, const void* q, int ld, const int32_t* edges,
                                  int64_t E, int n_steps, double step, uint32_t flags, uint8_t* out,
                                  void* stream) {
    if (!w || !q || !edges || !out || E < 0) return MMMP_ERR_ARG;
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
    if (E == 0) return MMMP_OK;
    if (w->kind == MMMP_WORLD_PLANAR)
        return mmmp_planar_collide_edges(w, robot, (const double*)qa, (const double*)qb, ld, nullptr, E, n_steps,
                                         step, flags, out, stream);
    return launch_edges(w, robot, (const float*)qa, (const float*)qb, ld, nullptr, E, n_steps, step, flags, out,
                        stream);
}
