/*
 * mmmp_b200.h — C ABI of the B200-native PRM roadmap-construction hot path.
 *
 * This is the drop-in boundary for ViktorLaurens/MMMP's learning phase
 * (sample -> FK -> config/edge collision -> neighbour search -> adjacency).
 * The reference is pure Python calling pybullet / scipy; every entry point
 * below names the reference call site (path:line under the reference tree)
 * whose arithmetic it replaces.  A maintainer binds it with ctypes
 * (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary;
 *   - `*_d` pointers are DEVICE memory, `*_h` pointers are HOST memory;
 *   - every device entry point is asynchronous on `stream` (a cudaStream_t
 *     passed as void*, NULL = legacy default stream); the caller owns and
 *     allocates every output buffer; nothing is retained after return;
 *   - return value: 0 = MMMP_OK, 1..99 = argument error, 1000+e = cudaError e;
 *   - a call with a row count of 0 returns MMMP_OK without touching its buffers, but
 *     required pointers must still be non-NULL (NULL always means "argument missing");
 *   - configurations are row-major [N, ld] with ld >= D; the fp32 copies used
 *     by the spatial kernels are padded to ld % 4 == 0 for 16-byte loads.
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns 1000 + cudaErrorNoDevice (or whatever CUDA reports).
 */
#ifndef MMMP_B200_H
#define MMMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMMP_OK 0
#define MMMP_ERR_ARG 1
#define MMMP_ERR_LIMIT 2
#define MMMP_ERR_KIND 3
#define MMMP_ERR_CUDA_BASE 1000

#define MMMP_MAX_JOINTS 8     /* joints per serial chain                       */
#define MMMP_MAX_SPHERES 320  /* collision spheres per robot                   */
#define MMMP_MAX_PAIR_TABLES 2 /* exact 2-joint link-pair tables per robot       */
#define MMMP_PAIR_TABLE_WORDS 4096 /* 32-bit words per table (<= 131072 cells)  */
#define MMMP_MAX_LINKS 16     /* collision link ids per robot                  */
#define MMMP_MAX_ROBOTS 16    /* robots per world                              */
#define MMMP_MAX_OBSTACLES 64 /* obstacle primitives per world                 */
#define MMMP_MAX_K 64         /* neighbours per query (k1 <= 55 in reference)  */
#define MMMP_MAX_GROUPS 8     /* frame groups of the FK metric                 */

/* ---- world description (host side, copied to the device by *_create) ---- */

/* Serial chain with revolute-z joints: frame 0 = base link, frame f (1..n) =
 * base * prod_{j<f} origin[j] * Rz(q_j); the tip frame is frame n * origin[n].
 * Replaces: Panda.dh_params / modified_homogeneous_transform
 * (robots/panda.py:31-40,111-122), the URDF joint origins pybullet loads
 * (robots/robot.py:22; models/kuka_iiwa/urdf/iiwa14_spheres_collision.urdf)
 * and Robot.base_position/base_orientation (robots/robot.py:11-12).        */
/* Optional exact decision table for one self-collision link pair whose relative
 * pose depends on exactly two joints (u, v): cell (iu, iv) of a regular grid over
 * [lo_u,hi_u] x [lo_v,hi_v] holds 1 when the two link HULLS may touch anywhere in
 * the cell (fitted offline, conservative by the cell's Lipschitz bound).  The pair
 * is then decided by one bit lookup instead of sphere tests (Panda link5/link7,
 * whose hulls are never more than 4 cm apart).                                  */
typedef struct mmmp_pair_table {
    int32_t link_a, link_b;
    int32_t joint_u, joint_v;
    int32_t n_u, n_v;
    double lo_u, hi_u, lo_v, hi_v;
    uint32_t bits[MMMP_PAIR_TABLE_WORDS];     /* bit iu * n_v + iv                      */
} mmmp_pair_table;

typedef struct mmmp_robot_desc {
    int32_t n_joints;
    int32_t n_spheres;
    int32_t n_links;
    int32_t dof_offset;                       /* column of joint 0 in a composite row */
    int32_t n_pair_tables;
    int32_t reserved;
    double base[12];                          /* row-major 3x4 world_T_base            */
    double origin[MMMP_MAX_JOINTS + 1][12];   /* row-major 3x4, see above              */
    double sphere[MMMP_MAX_SPHERES][4];       /* x, y, z (frame local), radius         */
    int32_t sphere_link[MMMP_MAX_SPHERES];    /* collision link id, < n_links; sorted  */
    int32_t link_frame[MMMP_MAX_LINKS];       /* frame of link l: 0 = base, f = after joint f-1 */
    uint32_t self_mask[MMMP_MAX_LINKS];       /* bit b of [a]: test link a vs link b   */
    mmmp_pair_table pair_table[MMMP_MAX_PAIR_TABLES];
} mmmp_robot_desc;

enum { MMMP_OBS_PLANE = 0, MMMP_OBS_SPHERE = 1, MMMP_OBS_BOX = 2, MMMP_OBS_CYLINDER = 3 };

/* PLANE   : p = {nx,ny,nz,d}: solid half-space n.x <= d (unit n)
 * SPHERE  : p = {cx,cy,cz,r}
 * BOX     : p = {cx,cy,cz, hx,hy,hz, R row-major 3x3 (world_R_box)}
 * CYLINDER: p = {cx,cy,cz, ax,ay,az (unit axis), half_len, radius}
 * Replaces the pybullet bodies passed as `obstacles` to
 * Environment(agents, obstacles) (planners/prm/pybullet/env.py:16-19).    */
typedef struct mmmp_obstacle_desc {
    int32_t type;
    int32_t body;      /* obstacle body id (several primitives may share one) */
    double p[16];
} mmmp_obstacle_desc;

/* Planar world: replaces planners/prm/planar/env.py:21-47 (map bounds, arms,
 * Rectangle/Circle obstacles) and robots/planar_arm.py:12-21.              */
enum { MMMP_POBS_RECT = 0, MMMP_POBS_CIRCLE = 1 };
typedef struct mmmp_planar_desc {
    int32_t n_arms;
    int32_t n_obstacles;
    int32_t n_links[MMMP_MAX_ROBOTS];
    int32_t dof_offset[MMMP_MAX_ROBOTS];
    double link_len[MMMP_MAX_ROBOTS][MMMP_MAX_JOINTS];
    double base[MMMP_MAX_ROBOTS][2];
    double bounds[4];                          /* xmin, xmax, ymin, ymax            */
    int32_t obs_type[MMMP_MAX_OBSTACLES];
    double obs[MMMP_MAX_OBSTACLES][4];         /* RECT x,y,w,h ; CIRCLE cx,cy,r,-   */
    double joint_link_radius;                  /* 0.5 in planar/env.py:130          */
} mmmp_planar_desc;

typedef struct mmmp_world mmmp_world;          /* opaque; owns its device tables    */

int mmmp_world_create(const mmmp_robot_desc* robots_h, int n_robots,
                      const mmmp_obstacle_desc* obstacles_h, int n_obstacles,
                      mmmp_world** out);
int mmmp_planar_world_create(const mmmp_planar_desc* desc_h, mmmp_world** out);
int mmmp_world_destroy(mmmp_world* w);
/* total joint count of the composite configuration (sum over robots) */
int mmmp_world_dof(const mmmp_world* w);

/* ---- collision flags: which predicate families a check includes --------- */
enum {
    MMMP_CHECK_SELF = 1,        /* env.py:77-84 / planar env.py:208-217              */
    MMMP_CHECK_OBSTACLES = 2,   /* env.py:51-68 / planar env.py:148-205              */
    MMMP_CHECK_ROBOTS = 4,      /* env.py:70-75 / planar env.py:78-117               */
    MMMP_CHECK_BOUNDS = 8,      /* planar env.py:220-229 (ignored by spatial worlds) */
    MMMP_CHECK_EXHAUSTIVE = 16, /* edge checks only: evaluate every interpolation step even
                                   when a clearance bound already proves it free (same
                                   verdicts; the cross-check of the default path)       */
    MMMP_CHECK_NO_CAPSULES = 32 /* spatial worlds: skip the capsule stage between the links'
                                   bounding spheres and their sphere sets (same verdicts;
                                   the cross-check of the capsule cull)                  */
};

/* ---- (a) joint-space sampling -------------------------------------------
 * np.random.uniform(lo, hi) then np.round(.,2):
 * planners/prm/pybullet/decoupled/decoupled_prm.py:310-318,
 * planners/prm/pybullet/coupled/coupled_prm.py:88-94,
 * planners/prm/planar/single_arm/prm.py:52-58.
 * Counter-based Philox4x32-10: sample i, joint j depends only on
 * (seed, first_index + i, j).  round_dp < 0 disables rounding.
 * out64_d [N, ld64] (may be NULL), out32_d [N, ld32] (may be NULL; padding
 * columns are zero-filled).                                                */
int mmmp_sample_uniform(const double* lo_h, const double* hi_h, int D,
                        int64_t N, uint64_t seed, uint64_t first_index, int round_dp,
                        double* out64_d, int ld64, float* out32_d, int ld32,
                        void* stream);

/* fp64 [N, ld64] -> zero-padded fp32 [N, ld32] */
int mmmp_pack_f32(const double* q64_d, int ld64, int D, int64_t N,
                  float* out32_d, int ld32, void* stream);

/* ---- (b) forward kinematics ---------------------------------------------
 * Panda.positions_from_fk (robots/panda.py:151-173): out_pos_d
 * [N, n_joints+1, 3] = origins of frames 1..n and of the tip, fp32.
 * out_T_d (may be NULL) [N, n_joints+1, 12] row-major 3x4 frame poses
 * (Panda.forward_kinematics robots/panda.py:124-144 = last frame).         */
int mmmp_fk_chain(const mmmp_world* w, int robot, const float* q_d, int ld, int64_t N,
                  float* out_pos_d, float* out_T_d, void* stream);
/* feature rows of the FK metric: out_d [N, ldf] fp32 = xyz of the selected
 * frames (frames_h[g] in 1..n_joints+1, the tip being n_joints+1), zero padded.
 * Panda.task_space_pose_distance (robots/panda.py:248-254) only depends on
 * frames 3,4,5(=6),7,8 -> 5 groups with weights 1,1,2,1,1.
 * out_filter_d (may be NULL), 16-byte aligned: the filter rows of the neighbour scan.
 *   ldfilt = 4: [N, 4] = (C, 0), C = sum_g w_g p_g the weighted centroid -- triangle
 *               inequality: sum_g w_g |dp_g| >= |dC|;
 *   ldfilt = 8: [N, 8] = (C, D, 0, 0) "dual rows": with A = the weighted sum over the first
 *               floor(n/2) groups and B over the rest, C = A + B and D = A - B.  Both
 *               |dC| <= m and |dA|^2 + |dB|^2 = (|dC|^2 + |dD|^2) / 2 <= (|dA| + |dB|)^2 <= m^2
 *               hold for the metric value m, and the pair of tests passes ~4x fewer
 *               candidates than the centroid alone (the halves of the chain move apart
 *               long before their sum does).  mmmp_knn recognises these rows by
 *               kind = GROUP_L2_SUM, ldf = 8, fdim = 6.                              */
int mmmp_fk_features(const mmmp_world* w, int robot, const float* q_d, int ld, int64_t N,
                     const int32_t* frames_h, const float* weights_h, int n_frames,
                     float* out_d, int ldf, float* out_filter_d, int ldfilt, void* stream);
/* frame groups of the FK metric for mmmp_fk_features / mmmp_metric: frames whose
 * origin depends on q, coincident consecutive frames merged with multiplicity as
 * weight.  frames_h / weights_h hold MMMP_MAX_GROUPS entries.                 */
int mmmp_fk_metric_groups(const mmmp_world* w, int robot, int32_t* frames_h,
                          float* weights_h, int* n_groups_h);
/* same chain in fp64 (used to re-rank k-NN candidates; 1e-15 parity)       */
int mmmp_fk_chain_f64(const mmmp_world* w, int robot, const double* q_d, int ld, int64_t N,
                      double* out_pos_d, void* stream);
/* PlanarArm.forward_kinematics (robots/planar_arm.py:23-40), fp64:
 * out_pos_d [N, n_links, 2]                                                 */
int mmmp_planar_fk(const mmmp_world* w, int arm, const double* q_d, int ld, int64_t N,
                   double* out_pos_d, void* stream);

/* ---- (c) collision checks -----------------------------------------------
 * Spatial worlds: sphere decomposition of every link against obstacle
 * primitives, the robot's own non-adjacent links and other robots' links,
 * culled link by link with a bounding sphere and then a bounding CAPSULE
 * (fitted at world creation around the link's spheres: conservative);
 * replaces pybullet getClosestPoints(distance=0) at
 * planners/prm/pybullet/env.py:65,72,81 as composed by
 * DecoupledPRM.is_collision_free_sample (decoupled_prm.py:840-850) and
 * CoupledPRM.is_collision_free_sample (coupled/coupled_prm.py:143-160).
 * Planar worlds: the analytic fp64 predicates of planners/prm/planar/env.py.
 * robot >= 0: q rows hold that robot's joints only (decoupled planners);
 * robot  < 0: q rows are composite rows (all robots; coupled planners).
 * out_d[i] = 1 if configuration i is IN COLLISION, else 0.                 */
int mmmp_collide_configs(const mmmp_world* w, int robot, const void* q_d, int ld,
                         int64_t N, uint32_t flags, uint8_t* out_d, void* stream);

/* is_collision_free_edge (decoupled_prm.py:855-863, coupled_prm.py:179-187,
 * planar single_arm/prm.py:161-169): for t = i*step, i in [0, n_steps):
 * q = qa + t (qb - qa); out_d[e] = 1 if ANY interpolant collides.
 * edges_d [E,2] int32 rows (a, b) index q_d.  Spatial worlds with
 * n_steps <= 64 evaluate a step only when no earlier evaluation of the same
 * edge proves it free: an evaluation returns the clearance of the sphere
 * model, and |d centre / d q_m| <= lever arm bounds how far every sphere can
 * move per step, so the verdict equals the OR over all steps.  Edges on
 * which that proves little are finished by a warp-per-edge pass.            */
int mmmp_collide_edges(const mmmp_world* w, int robot, const void* q_d, int ld,
                       const int32_t* edges_d, int64_t E, int n_steps, double step,
                       uint32_t flags, uint8_t* out_d, void* stream);
/* same, with explicit endpoint arrays qa_d[e], qb_d[e] (start/goal hookup,
 * decoupled_prm.py:943-995; query-phase pair lists)                         */
int mmmp_collide_segments(const mmmp_world* w, int robot, const void* qa_d, const void* qb_d,
                          int ld, int64_t E, int n_steps, double step,
                          uint32_t flags, uint8_t* out_d, void* stream);
/* The verdicts of a whole candidate table: what connect_nearest_neighbors_degree
 * asks of is_collision_free_edge (decoupled_prm.py:496-503) for the slots of
 * rows [row_lo, row_hi) of nbr_d [n, k] (-1 = empty slot), spatial worlds.
 * free_d [row_hi - row_lo, k]: 1 = edge row -> neighbour collision free, 0 = not
 * (or empty slot).  When i and j name each other, the reference's two checks
 * i -> j and j -> i walk the same configurations (the t grid of :859 is symmetric
 * when n_steps * step == 1; the extra end point is a free node), so the pair is
 * evaluated ONCE, from the lower id to the higher -- the first of the reference's
 * two checks -- and both slots receive that verdict; every other slot is
 * evaluated row -> neighbour.  share_across_rows == 0: only pairs with both rows
 * inside [row_lo, row_hi) share an evaluation, free_d comes back complete.
 * share_across_rows != 0 (several ranks, each with a block of rows): the pair's
 * evaluation belongs to one of its two rows wherever that row lives; slots whose
 * verdict another row holds come back as 2 and are filled in by
 * mmmp_resolve_knn_edges on the gathered table.  Verdicts do not depend on how
 * the rows are cut.  n_steps * step != 1: nothing is shared.                 */
int mmmp_collide_knn_edges(const mmmp_world* w, int robot, const void* q_d, int ld,
                           const int32_t* nbr_d, int64_t n, int k, int64_t row_lo,
                           int64_t row_hi, int share_across_rows, int n_steps, double step,
                           uint32_t flags, uint8_t* free_d, void* stream);
/* free_d [row_hi - row_lo, k] (rows [row_lo, row_hi) of the table): every slot
 * marked 2 takes the verdict of its partner's slot (both rows inside the block) */
int mmmp_resolve_knn_edges(const int32_t* nbr_d, int64_t n, int k, int64_t row_lo,
                           int64_t row_hi, uint8_t* free_d, void* stream);
/* profiling aid: enable != 0 turns on four device counters of the spatial edge
 * kernels (edges seen, configurations evaluated by the step-skipping pass,
 * edges deferred to the warp-per-edge pass, configurations evaluated there);
 * out4_h (may be NULL) receives and resets them; enable == 0 frees them.    */
int mmmp_edge_stats(int enable, unsigned long long* out4_h);
/* robot_robot_collision(q1, q2, r1, r2) on a pair list
 * (decoupled_prm.py:871-876; cbsprm.py:165-183): qa_d rows are robot ra's
 * joints, qb_d rows robot rb's.                                            */
int mmmp_collide_robot_pairs(const mmmp_world* w, int ra, int rb, const void* qa_d,
                             const void* qb_d, int ld, int64_t N, uint8_t* out_d,
                             void* stream);

/* stable compaction of the rows with flag == 0 (rejection sampling):
 * idx_out_d [<=N] ascending row indices; count_d receives the count.        */
int mmmp_compact_free(const uint8_t* flags_d, int64_t N, int32_t* idx_out_d,
                      int32_t* count_d, void* stream);
/* out[i, :] = in[idx[i], :] for i < n ; row_bytes a multiple of 4             */
int mmmp_gather_rows(const void* in_d, int row_bytes, const int32_t* idx_d, int64_t n,
                     void* out_d, void* stream);

/* ---- (d) neighbour search ----------------------------------------------- */
enum {
    MMMP_METRIC_L2 = 0,            /* np.linalg.norm(q1-q2): decoupled_prm.py:824, coupled_prm.py:131 */
    MMMP_METRIC_GROUP_L2_SUM = 1,  /* sum_g w_g ||p_g(q1)-p_g(q2)||: robots/panda.py:248-254          */
    MMMP_METRIC_SQ_SUM = 2         /* sum ||.||^2: robots/planar_arm.py:54-61                          */
};

/* metric description: features are fp32 rows [N, ldf]; for GROUP_L2_SUM the
 * first 3*n_groups columns are xyz triples with weights w[g].               */
typedef struct mmmp_metric {
    int32_t kind;
    int32_t dim;                      /* feature columns used                       */
    int32_t n_groups;
    float weight[MMMP_MAX_GROUPS];
} mmmp_metric;

/* Brute-force k nearest neighbours: replaces the heap loops
 * decoupled_prm.py:480-504,527-550,943-995 and KDTree.query at
 * coupled/degree_coupled_prm.py:67, planar single_arm/degree_prm.py:178.
 * For query row i (global id query_id0 + i) the k nearest reference rows with
 * id in [id_lo, id_hi_i) excluding the query's own id when exclude_self;
 * id_hi_i = min(n_ref, causal ? query id : n_ref)  (causal = only earlier
 * nodes, the incremental planners).  Output ascending (dist, id); missing
 * slots get id -1 and dist +inf.  out_dist_d may be NULL.
 * Two row sets per side: FILTER rows fref_d/fquery_d [., ldf] whose first fdim
 * (<= 31) columns have a plain L2 distance that is a LOWER BOUND of the metric
 * -- every pair is tested against it in the hot loop -- and EXACT rows
 * xref_d/xquery_d [., ldx] on which metric_h is evaluated for the pairs that
 * pass.  L2 / SQ_SUM: pass the same rows for both (fdim = metric.dim).
 * GROUP_L2_SUM: exact rows = mmmp_fk_features out_d, filter rows = its
 * out_filter_d: centroid rows (ldf 4, fdim 3) or dual rows (ldf 8, fdim 6: the first
 * three columns bound the metric m by |d|^2 <= m^2, all six by |d|^2 <= 2 m^2; both
 * tests run in the hot loop, tiles are culled on the first three).  n_ref < 2^28.  */
int mmmp_knn(const float* fref_d, const float* xref_d, int64_t n_ref,
             const float* fquery_d, const float* xquery_d, int64_t n_query,
             int ldf, int fdim, int ldx, const mmmp_metric* metric_h, int k,
             int exclude_self, int causal, int64_t query_id0, int32_t* out_idx_d,
             float* out_dist_d, void* stream);

/* Multi-GPU form of mmmp_knn for a search of a set against itself:
 * the library orders the rows spatially before it scans them, and rank `shard`
 * of `n_shards` answers one contiguous slice of THAT order -- compact query
 * blocks, the same tile culling as on one GPU, 1/n_shards of the work -- instead
 * of a slice of the caller's row order.  out_rows_d [cap] receives the original
 * row of every answered query, out_idx_d / out_dist_d [cap, k] its neighbours,
 * *n_rows_h (host) their number; cap >= (ceil(ceil(n/128) / n_shards) + 1) * 128.
 * The union over the shards is every row exactly once.                      */
int mmmp_knn_shard(const float* fref_d, const float* xref_d, int64_t n_ref, int ldf, int fdim,
                   int ldx, const mmmp_metric* metric, int k, int exclude_self, int shard,
                   int n_shards, int32_t* out_rows_d, int64_t* n_rows_h, int32_t* out_idx_d,
                   float* out_dist_d, void* stream);

/* The same with explicit cuts: the query blocks [floor(B * frac_lo), floor(B * frac_hi))
 * of the B = ceil(n / 128) blocks of the scan order (0 <= frac_lo <= frac_hi <= 1).  Ranks
 * that use the cuts c_0 = 0 < c_1 < ... < c_G = 1 answer every row exactly once; moving the
 * cuts balances ranks whose slices cost differently (RoadmapBuilder adapts them from the
 * measured scan times).  cap >= (ceil(B * (frac_hi - frac_lo)) + 1) * 128.                 */
int mmmp_knn_range(const float* fref_d, const float* xref_d, int64_t n_ref, int ldf, int fdim,
                   int ldx, const mmmp_metric* metric, int k, int exclude_self, double frac_lo,
                   double frac_hi, int32_t* out_rows_d, int64_t* n_rows_h, int32_t* out_idx_d,
                   float* out_dist_d, void* stream);

/* Post-search exchange of a sharded search ("the local adjacency lists are gathered
 * afterwards"): the m rows a rank answered (out_rows_d of mmmp_knn_range, re-ranked by
 * mmmp_knn_refine) packed as records of 2 + 3k 32-bit words [row id, ambiguous, idx[k],
 * dist[k] fp64] in rec_d [cap, 2 + 3k] (records m..cap-1: row id -1), so that one all-gather
 * of equal blocks ships them; mmmp_scatter_neighbour_records writes any number of gathered
 * records into tables in the caller's row order (idx_d [n, k], dist_d [n, k] or NULL,
 * amb_d [n] or NULL), skipping row ids outside [0, n).                              */
int mmmp_pack_neighbour_records(const int32_t* rows_d, const int32_t* idx_d, const double* dist_d,
                                const uint8_t* amb_d, int64_t m, int k, int64_t cap, void* rec_d,
                                void* stream);
int mmmp_scatter_neighbour_records(const void* rec_d, int64_t n_records, int k, int64_t n,
                                   int32_t* idx_d, double* dist_d, uint8_t* amb_d, void* stream);

/* The scan order of the neighbour search (host function, no device work): position
 * of grid cell (x0, x1, x2), `bits` bits per axis, on the 3-D Hilbert curve the
 * library sorts its rows along.  Consecutive positions are face-adjacent cells,
 * which is what keeps a run of rows -- a tile, a query block -- compact
 * (tests/test_abi_cpu.py checks the bijection and the adjacency).            */
int mmmp_scan_order_cell(int x0, int x1, int x2, int bits, uint32_t* index_h);

/* profiling aid: enable != 0 turns on five device counters of the scan kernel
 * (exact evaluations, list insertions, warp drains, tiles loaded, most tiles
 * loaded by one CTA); out5_h (may be NULL) receives and resets them;
 * enable == 0 frees them.                                                   */
int mmmp_knn_stats(int enable, unsigned long long* out5_h);

/* Re-rank candidates in fp64 with the reference's own arithmetic order:
 * cand_idx_d [n_query, kc] (from mmmp_knn with kc >= k) -> out [n_query, k]
 * ascending (dist64, id).  L2/SQ_SUM: rows of ref64_d/query64_d [., ld64]
 * are the configurations (or fp64 features); GROUP_L2_SUM: they are fp64
 * frame positions [., ld64] from mmmp_fk_chain_f64 and metric->dim selects
 * the triples.  ambiguous_d (may be NULL) [n_query] = 1 when the fp32 scan
 * could have dropped a row of the float64 top k: the gap between its kc-th fp32
 * value and the float64 k-th distance is within max(tie_eps, 4 x the largest
 * |fp32 - float64| difference measured on the row's own candidates).  Such rows
 * are to be searched again with more candidates (core.knn_exact, RoadmapBuilder
 * and mmmp_roadmap_*_host do).                                              */
int mmmp_knn_refine(const double* ref64_d, const double* query64_d, int ld64,
                    const mmmp_metric* metric_h, const int32_t* cand_idx_d,
                    const float* cand_dist_d, int64_t n_query, int kc, int k,
                    double tie_eps, int32_t* out_idx_d, double* out_dist_d,
                    uint8_t* ambiguous_d, void* stream);

/* Radius search (dist <= r, or dist < r when strict): replaces
 * connect_nearest_neighbors_distance (decoupled_prm.py:553-571),
 * DistancePRM.add_n_samples (planar single_arm/distance_prm.py:40-52,
 * coupled/distance_coupled_prm.py:40-52) and KDTree.query_ball_point
 * (coupled/distance_coupled_prm.py:133).  The fp32 rows drive a conservative
 * filter; rows that pass are decided in fp64 on ref64_d/query64_d with the
 * reference's arithmetic (metric64_h describes those rows: L2 -> configs,
 * GROUP_L2_SUM / SQ_SUM -> fp64 frame / link positions, n_groups = frames),
 * so membership is exact.  Two-pass CSR: call with out_idx_d == NULL to fill
 * counts_d [n_query]; exclusive-scan them into offsets_d [n_query+1]
 * (mmmp_exclusive_scan_i32), then call again to fill out_idx_d / out_dist_d
 * (ids ascending within a row).                                             */
int mmmp_radius(const float* ref_d, const double* ref64_d, int64_t n_ref,
                const float* query_d, const double* query64_d, int64_t n_query,
                int ldf, int ld64, const mmmp_metric* metric_h,
                const mmmp_metric* metric64_h, double r, int strict, int exclude_self,
                int exclude_zero, int causal, int64_t query_id0, int32_t* counts_d,
                const int64_t* offsets_d, int32_t* out_idx_d, double* out_dist_d,
                void* stream);
int mmmp_exclusive_scan_i32(const int32_t* in_d, int64_t n, int64_t* out_d /* n+1 */,
                            void* stream);

/* ---- host-buffer entry points (the reference-facing call) ----------------
 * One call = the learning phase of a decoupled degree PRM for one robot on
 * caller-provided samples, host memory in, host memory out
 * (DecoupledPRM.generate_roadmap decoupled_prm.py:231-260 with the joint-space
 * sampler's accept test decoupled_prm.py:310-321):
 *   H2D samples -> config collision -> compaction of the free samples (= the
 *   nodes, ids in input order) -> metric features -> k-NN scan + fp64 re-rank
 *   -> edge collision of every (node, candidate) pair -> D2H.
 * q64_h [N, D] fp64.  Outputs (host, sized for N rows; the first *n_nodes_h
 * rows are valid): node_src_h [N] input row of node i, nbr_idx_h [N, k] int32
 * node ids ascending (dist, id) (-1 = none), nbr_dist_h [N, k] fp64 (may be
 * NULL), edge_free_h [N, k] u8 (1 = the straight edge is collision free).
 * metric_kind: MMMP_METRIC_L2 or MMMP_METRIC_GROUP_L2_SUM (Panda FK metric,
 * all chain frames, panda.py:248-254).  Synchronous.  timings_ms_h (may be
 * NULL) [8]: h2d, config collision+compaction, features, knn scan, re-rank,
 * edge collision, d2h, total (CUDA events).                                 */
int mmmp_roadmap_candidates_host(const mmmp_world* w, int robot, const double* q64_h,
                                 int64_t N, int D, int metric_kind, int k, int n_steps,
                                 double step, uint32_t flags, int64_t* n_nodes_h,
                                 int32_t* node_src_h, int32_t* nbr_idx_h,
                                 double* nbr_dist_h, uint8_t* edge_free_h,
                                 float* timings_ms_h);

/* The same call up to the ROADMAP: candidates as above with k = k1, then the greedy
 * degree-capped accept pass (connect_nearest_neighbors_degree decoupled_prm.py:496-503,
 * mmmp_greedy_degree_connect_device) and the connected components of the result
 * (decoupled_prm.py:613-636).  Additional host outputs, sized for N rows: adj_h [N, k2]
 * int32 node ids (-1 padded, the reference's list order), deg_h [N], labels_h [N] (may be
 * NULL; smallest node id of the component), n_components_h (may be NULL).  The neighbour
 * tables travel device->host on a second stream while the edges are checked, the verdicts
 * while the accept pass runs (page-locked host buffers make the overlap real).
 * timings_ms_h (may be NULL) [10]: h2d, accept test + compaction, features, scan, re-rank,
 * edge collision, d2h tail after the last kernel, total, accept pass, components.           */
int mmmp_roadmap_build_host(const mmmp_world* w, int robot, const double* q64_h, int64_t N, int D,
                            int metric_kind, int k1, int k2, int n_steps, double step,
                            uint32_t flags, int64_t* n_nodes_h, int32_t* node_src_h,
                            int32_t* nbr_idx_h, double* nbr_dist_h, uint8_t* edge_free_h,
                            int32_t* adj_h, int32_t* deg_h, int32_t* labels_h,
                            int64_t* n_components_h, float* timings_ms_h);

/* edges_d[e] = (e / k, nbr_d[e]) for a flat [n, k] neighbour table            */
int mmmp_edges_from_knn(const int32_t* nbr_d, int64_t n, int k, int32_t* edges_d,
                        void* stream);

/* Greedy sequential accept pass of connect_nearest_neighbors_degree
 * (decoupled_prm.py:496-503) on precomputed candidates, host side, native:
 * nodes in id order, candidates in (dist,id) order, accept while deg < k2 on
 * both ends, edge free and not already present.  adj_h [N, cap] int32 (-1
 * padded), deg_h [N].  cap_other_end = 0 reproduces connect_extra_nodes_*.  */
int mmmp_greedy_degree_connect(const int32_t* nbr_idx_h, const uint8_t* edge_free_h,
                               int64_t N, int k1, int k2, int cap, int32_t* adj_h,
                               int32_t* deg_h);

/* The same pass on the GPU, device tables in, device tables out (asynchronous on `stream`):
 * identical adjacency -- the same edges in the same order inside every list -- computed in a
 * few data-parallel rounds instead of one sequential sweep (csrc/greedy.cu: every node
 * bounds its degree at the time of each candidate from the decisions known so far; a
 * candidate is accepted once the cap provably cannot bind at either end, rejected once it
 * provably binds at one).  nbr_idx_d [N, k1], edge_free_d [N, k1] as produced by the
 * neighbour search and mmmp_collide_edges; adj_d [N, cap] (-1 padded), deg_d [N];
 * rounds_d (may be NULL): device scalar receiving the number of rounds.
 * N * k1 < 2^31.  This is what puts "roadmap build" on the device up to the adjacency.  */
int mmmp_greedy_degree_connect_device(const int32_t* nbr_idx_d, const uint8_t* edge_free_d,
                                      int64_t N, int k1, int k2, int cap, int32_t* adj_d,
                                      int32_t* deg_d, int32_t* rounds_d, void* stream);

/* Connected components of a roadmap: the partition computed by
 * compute_connected_components (decoupled_prm.py:613-636).  adj_d [N, cap]
 * int32 neighbour table, -1 padded (the accepted adjacency of
 * mmmp_greedy_degree_connect, or a candidate table with mask_d [N, cap] = 1
 * where the candidate edge is collision free; mask_d may be NULL); an edge may
 * be listed from either end.  labels_d[i] = smallest node id of i's component
 * (the node the reference's outer loop starts that component from);
 * n_components_d (may be NULL) receives the number of components.          */
int mmmp_connected_components(const int32_t* adj_d, const uint8_t* mask_d, int64_t N, int cap,
                              int32_t* labels_d, int64_t* n_components_d, void* stream);

/* Batched inverse kinematics for the task-space sampler:
 * Panda.solve_closest_arm_ik (robots/panda.py:207-208 -> utils/ik_utils.py:25-27
 * -> pybullet.calculateInverseKinematics) as called by sample_valid_in_task_space
 * (decoupled_prm.py:323-344).  Damped least squares on the geometric Jacobian of
 * the tip frame (frame n_joints * origin[n_joints]), one problem per thread:
 * dq = J^T (J J^T + lambda^2 I)^-1 e, |dq|_inf <= max_step, joints clamped to
 * [lo_h, hi_h], until |e_pos| < tol_pos and |e_rot| < tol_rot or max_iters.
 * target_pos_d [N,3], target_quat_d [N,4] (x,y,z,w; NULL = position only),
 * q0_d [N, ld0] start configurations (ld0 = 0: one row for all), q_out_d [N, ld]
 * fp64 (not rounded), converged_d [N] (may be NULL), residual_d [N,2] = final
 * |e_pos|, |e_rot| (may be NULL).  Parity with pybullet's own solver is
 * unpinned (pybullet absent); pinned against oracle/ik.py.                  */
int mmmp_ik_dls(const mmmp_world* w, int robot, const double* target_pos_d, const double* target_quat_d,
                const double* q0_d, int ld0, int64_t N, const double* lo_h, const double* hi_h,
                int max_iters, double lambda, double tol_pos, double tol_rot, double max_step,
                double* q_out_d, int ld, uint8_t* converged_d, float* residual_d, void* stream);

/* library / device introspection */
int mmmp_version(void);
int mmmp_device_info(int* sm_count, int* cc_major, int* cc_minor, int* clock_khz);
/* number of kernel launches issued by this library since load (bench.py's
 * gpu_launches)                                                             */
int64_t mmmp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MMMP_B200_H */
