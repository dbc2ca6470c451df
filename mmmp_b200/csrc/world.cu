// World (scene) tables: host descriptors -> device-resident constant tables.
// Replaces the state pybullet keeps behind loadURDF / the planar Environment
// object (reference: robots/robot.py:19-28, planners/prm/pybullet/env.py:15-29,
// planners/prm/planar/env.py:21-47).
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>
#include "common.cuh"
#include "obstacle.cuh"

std::atomic<int64_t> g_mmmp_launches{0};

// q-independent part: spheres of frame-0 links (already in world coordinates)
// against every obstacle primitive.  One thread per robot.
__global__ void world_static_kernel(WorldDev* W) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= W->n_robots) return;
    RobotDev& R = W->robots[r];
    int hit = 0;
    for (int l = 0; l < R.n_links; ++l) {
        if (R.link_frame[l] != 0) continue;
        for (int s = R.link_sphere_begin[l]; s < R.link_sphere_begin[l + 1]; ++s) {
            const float4 sp = R.sphere[s];
            for (int o = 0; o < W->n_obstacles; ++o)
                hit |= sphere_hits_obstacle(sp.x, sp.y, sp.z, sp.w, W->obstacles[o]) ? 1 : 0;
        }
    }
    R.base_hits_obstacles = hit;
}

static void affine_point_d(const double* T, const double* p, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = T[4 * i] * p[0] + T[4 * i + 1] * p[1] + T[4 * i + 2] * p[2] + T[4 * i + 3];
}

// bounding sphere of a set of spheres (Ritter pass over centres, then inflate)
static void bounding_sphere(const float4* sp, int n, float4* out) {
    if (n <= 0) {
        *out = make_float4(0.f, 0.f, 0.f, -1.f);
        return;
    }
    double c[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        c[0] += sp[i].x;
        c[1] += sp[i].y;
        c[2] += sp[i].z;
    }
    for (int k = 0; k < 3; ++k) c[k] /= n;
    // a few Badoiu-Clarkson steps on the inflated extent
    for (int it = 0; it < 64; ++it) {
        int far = 0;
        double fd = -1;
        for (int i = 0; i < n; ++i) {
            const double d = std::sqrt((sp[i].x - c[0]) * (sp[i].x - c[0]) + (sp[i].y - c[1]) * (sp[i].y - c[1]) +
                                       (sp[i].z - c[2]) * (sp[i].z - c[2])) + sp[i].w;
            if (d > fd) {
                fd = d;
                far = i;
            }
        }
        const double t = 1.0 / (it + 2);
        c[0] += (sp[far].x - c[0]) * t;
        c[1] += (sp[far].y - c[1]) * t;
        c[2] += (sp[far].z - c[2]) * t;
    }
    double r = 0;
    for (int i = 0; i < n; ++i) {
        const double d = std::sqrt((sp[i].x - c[0]) * (sp[i].x - c[0]) + (sp[i].y - c[1]) * (sp[i].y - c[1]) +
                                   (sp[i].z - c[2]) * (sp[i].z - c[2])) + sp[i].w;
        if (d > r) r = d;
    }
    *out = make_float4((float)c[0], (float)c[1], (float)c[2], (float)(r * (1.0 + 1e-6) + 1e-7));
}

// Bounding capsule of a set of spheres: axis = principal direction of the centres through their
// mean, radius = largest (distance from the axis line + sphere radius), segment ends pulled in as
// far as the hemispherical caps allow; the radius is then recomputed from the final segment so
// that containment holds by construction.
static void bounding_capsule(const float4* sp, int n, float4* out_a, float4* out_b) {
    if (n <= 0) {
        *out_a = make_float4(0.f, 0.f, 0.f, -1.f);
        *out_b = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    double mean[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        mean[0] += sp[i].x;
        mean[1] += sp[i].y;
        mean[2] += sp[i].z;
    }
    for (int k = 0; k < 3; ++k) mean[k] /= n;
    double C[3][3] = {{0}};
    for (int i = 0; i < n; ++i) {
        const double d[3] = {sp[i].x - mean[0], sp[i].y - mean[1], sp[i].z - mean[2]};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) C[a][b] += d[a] * d[b];
    }
    double u[3] = {1.0, 0.7, 0.4};
    for (int it = 0; it < 200; ++it) {  // power iteration
        double v[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) v[a] += C[a][b] * u[b];
        const double nv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (nv < 1e-30) break;  // all centres coincide: any axis
        for (int a = 0; a < 3; ++a) u[a] = v[a] / nv;
    }
    {
        const double nu = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        for (int a = 0; a < 3; ++a) u[a] /= nu;
    }
    double R = 0;
    for (int i = 0; i < n; ++i) {
        const double d[3] = {sp[i].x - mean[0], sp[i].y - mean[1], sp[i].z - mean[2]};
        const double t = d[0] * u[0] + d[1] * u[1] + d[2] * u[2];
        const double perp = std::sqrt(std::max(0.0, d[0] * d[0] + d[1] * d[1] + d[2] * d[2] - t * t));
        R = std::max(R, perp + sp[i].w);
    }
    double t_lo = 1e300, t_hi = -1e300;  // the segment may end where the caps still cover everything
    for (int i = 0; i < n; ++i) {
        const double d[3] = {sp[i].x - mean[0], sp[i].y - mean[1], sp[i].z - mean[2]};
        const double t = d[0] * u[0] + d[1] * u[1] + d[2] * u[2];
        const double perp2 = std::max(0.0, d[0] * d[0] + d[1] * d[1] + d[2] * d[2] - t * t);
        const double reach = std::sqrt(std::max(0.0, (R - sp[i].w) * (R - sp[i].w) - perp2));
        t_lo = std::min(t_lo, t + reach);
        t_hi = std::max(t_hi, t - reach);
    }
    if (t_lo > t_hi) t_lo = t_hi = 0.5 * (t_lo + t_hi);
    const double a[3] = {mean[0] + t_lo * u[0], mean[1] + t_lo * u[1], mean[2] + t_lo * u[2]};
    const double b[3] = {mean[0] + t_hi * u[0], mean[1] + t_hi * u[1], mean[2] + t_hi * u[2]};
    const float af[3] = {(float)a[0], (float)a[1], (float)a[2]}, bf[3] = {(float)b[0], (float)b[1], (float)b[2]};
    // radius from the segment as the kernels will see it (fp32 end points)
    double rad = 0;
    const double ab[3] = {(double)bf[0] - af[0], (double)bf[1] - af[1], (double)bf[2] - af[2]};
    const double ab2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
    for (int i = 0; i < n; ++i) {
        const double d[3] = {sp[i].x - (double)af[0], sp[i].y - (double)af[1], sp[i].z - (double)af[2]};
        double t = ab2 > 0 ? (d[0] * ab[0] + d[1] * ab[1] + d[2] * ab[2]) / ab2 : 0.0;
        t = std::min(1.0, std::max(0.0, t));
        const double e[3] = {d[0] - t * ab[0], d[1] - t * ab[1], d[2] - t * ab[2]};
        rad = std::max(rad, std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + sp[i].w);
    }
    *out_a = make_float4(af[0], af[1], af[2], (float)(rad * (1.0 + 1e-6) + 1e-6));
    *out_b = make_float4(bf[0], bf[1], bf[2], 0.f);
}

extern "C" int mmmp_world_create(const mmmp_robot_desc* robots, int n_robots,
                                 const mmmp_obstacle_desc* obstacles, int n_obstacles,
                                 mmmp_world** out) {
    if (!out || !robots || n_robots < 1 || (n_obstacles > 0 && !obstacles)) return MMMP_ERR_ARG;
    if (n_robots > MMMP_MAX_ROBOTS || n_obstacles > MMMP_MAX_OBSTACLES || n_obstacles < 0) return MMMP_ERR_LIMIT;
    WorldDev* H = (WorldDev*)calloc(1, sizeof(WorldDev));
    if (!H) return MMMP_ERR_ARG;
    H->n_robots = n_robots;
    H->n_obstacles = n_obstacles;
    int dof = 0;
    int table_words = 0;
    const int pool_words = (int)(sizeof(H->table_bits) / sizeof(unsigned int));
    for (int r = 0; r < n_robots; ++r) {
        const mmmp_robot_desc& d = robots[r];
        RobotDev& R = H->robots[r];
        RobotDev64& R64 = H->robots64[r];
        if (d.n_joints < 1 || d.n_joints > MMMP_MAX_JOINTS || d.n_spheres < 0 ||
            d.n_spheres > MMMP_MAX_SPHERES || d.n_links < 0 || d.n_links > MMMP_MAX_LINKS ||
            d.n_pair_tables < 0 || d.n_pair_tables > MMMP_MAX_PAIR_TABLES) {
            free(H);
            return MMMP_ERR_LIMIT;
        }
        R.n_joints = d.n_joints;
        R.n_spheres = d.n_spheres;
        R.n_links = d.n_links;
        R.dof_offset = d.dof_offset;
        R64.n_joints = d.n_joints;
        for (int i = 0; i < 12; ++i) {
            R.base[i] = (float)d.base[i];
            R64.base[i] = d.base[i];
        }
        for (int j = 0; j <= MMMP_MAX_JOINTS; ++j) R.origin_rx[j] = 0;
        for (int j = 0; j <= d.n_joints; ++j) {
            for (int i = 0; i < 12; ++i) {
                R.origin[j][i] = (float)d.origin[j][i];
                R64.origin[j][i] = d.origin[j][i];
            }
            const float* O = R.origin[j];
            R.origin_rx[j] = O[0] == 1.f && O[1] == 0.f && O[2] == 0.f && O[4] == 0.f && O[8] == 0.f;
        }
        // spheres sorted by link; link_sphere_begin[l] = #spheres with link < l
        for (int l = 0; l <= MMMP_MAX_LINKS; ++l) R.link_sphere_begin[l] = 0;
        int prev = 0;
        for (int l = 0; l < d.n_links; ++l) {
            if (d.link_frame[l] < 0 || d.link_frame[l] > d.n_joints) {
                free(H);
                return MMMP_ERR_ARG;
            }
            R.link_frame[l] = d.link_frame[l];
        }
        for (int s = 0; s < d.n_spheres; ++s) {
            const int l = d.sphere_link[s];
            if (l < prev || l >= d.n_links) {
                free(H);
                return MMMP_ERR_ARG;
            }
            prev = l;
            for (int g = l + 1; g <= MMMP_MAX_LINKS; ++g) R.link_sphere_begin[g]++;
            if (d.link_frame[l] == 0) {  // base link: store world coordinates
                double o[3];
                affine_point_d(d.base, d.sphere[s], o);
                R.sphere[s] = make_float4((float)o[0], (float)o[1], (float)o[2], (float)d.sphere[s][3]);
            } else {
                R.sphere[s] = make_float4((float)d.sphere[s][0], (float)d.sphere[s][1],
                                          (float)d.sphere[s][2], (float)d.sphere[s][3]);
            }
        }
        for (int f = 0; f < MMMP_MAX_JOINTS + 2; ++f) R.frame_link_begin[f] = 0;
        for (int l = 0; l < d.n_links; ++l) {
            if (l > 0 && d.link_frame[l] < d.link_frame[l - 1]) {  // links must be ordered by frame
                free(H);
                return MMMP_ERR_ARG;
            }
            for (int f = d.link_frame[l] + 1; f < MMMP_MAX_JOINTS + 2; ++f) R.frame_link_begin[f]++;
        }
        R.n_moving_links = 0;
        for (int l = 0; l < d.n_links; ++l) {
            bounding_sphere(R.sphere + R.link_sphere_begin[l], R.link_sphere_begin[l + 1] - R.link_sphere_begin[l],
                            &R.link_bound[l]);
            bounding_capsule(R.sphere + R.link_sphere_begin[l], R.link_sphere_begin[l + 1] - R.link_sphere_begin[l],
                             &R.cap_a[l], &R.cap_b[l]);
            if (R.link_frame[l] != 0) R.n_moving_links++;
        }
        // reach[m][f]: lever arm of joint m over the sphere centres of frame f (f > m), an upper
        // bound of |d centre / d q_m| over ALL configurations.  Enclose the centres of frame f in
        // a ball, then walk towards the base: seen from frame m+1 the ball may be anywhere on
        // its circle about that frame's z axis (= joint m's axis), so lever = |xy| + rho; the
        // swept set is enclosed by the ball centred on the axis, which the joint origin maps
        // into the next frame down.
        for (int m = 0; m < MMMP_MAX_JOINTS; ++m)
            for (int f = 0; f <= MMMP_MAX_JOINTS; ++f) R.reach[m][f] = 0.f;
        for (int f = 1; f <= d.n_joints; ++f) {
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            int cnt = 0;
            for (int s = 0; s < d.n_spheres; ++s) {
                if (d.link_frame[d.sphere_link[s]] != f) continue;
                for (int k = 0; k < 3; ++k) {
                    lo[k] = std::min(lo[k], d.sphere[s][k]);
                    hi[k] = std::max(hi[k], d.sphere[s][k]);
                }
                ++cnt;
            }
            if (!cnt) continue;
            double b[3] = {(lo[0] + hi[0]) / 2, (lo[1] + hi[1]) / 2, (lo[2] + hi[2]) / 2}, rho = 0;
            for (int s = 0; s < d.n_spheres; ++s) {
                if (d.link_frame[d.sphere_link[s]] != f) continue;
                const double dx = d.sphere[s][0] - b[0], dy = d.sphere[s][1] - b[1], dz = d.sphere[s][2] - b[2];
                rho = std::max(rho, std::sqrt(dx * dx + dy * dy + dz * dz));
            }
            for (int m = f - 1; m >= 0; --m) {
                const double arm = std::sqrt(b[0] * b[0] + b[1] * b[1]);
                R.reach[m][f] = (float)((arm + rho) * (1.0 + 1e-5) + 1e-7);
                const double axis_pt[3] = {0.0, 0.0, b[2]};
                rho += arm;
                affine_point_d(d.origin[m], axis_pt, b);
            }
        }
        // exact pair tables (deduplicated across robots by content)
        unsigned int table_mask[MMMP_MAX_LINKS];
        for (int l = 0; l < MMMP_MAX_LINKS; ++l) table_mask[l] = 0;
        R.n_tables = d.n_pair_tables;
        for (int t = 0; t < d.n_pair_tables; ++t) {
            const mmmp_pair_table& pt = d.pair_table[t];
            const int64_t cells = (int64_t)pt.n_u * pt.n_v;
            if (pt.n_u < 1 || pt.n_v < 1 || cells > (int64_t)MMMP_PAIR_TABLE_WORDS * 32 || pt.joint_u < 0 ||
                pt.joint_v < 0 || pt.joint_u >= d.n_joints || pt.joint_v >= d.n_joints || pt.link_a < 0 ||
                pt.link_b < 0 || pt.link_a >= d.n_links || pt.link_b >= d.n_links || !(pt.hi_u > pt.lo_u) ||
                !(pt.hi_v > pt.lo_v)) {
                free(H);
                return MMMP_ERR_ARG;
            }
            const int words = (int)((cells + 31) / 32);
            int off = -1;
            for (int o = 0; o + words <= table_words && off < 0; ++o)
                if (memcmp(H->table_bits + o, pt.bits, sizeof(unsigned int) * words) == 0) off = o;
            if (off < 0) {
                if (table_words + words > pool_words) {
                    free(H);
                    return MMMP_ERR_LIMIT;
                }
                off = table_words;
                memcpy(H->table_bits + off, pt.bits, sizeof(unsigned int) * words);
                table_words += words;
            }
            PairTableDev& T = R.table[t];
            T.joint_u = pt.joint_u;
            T.joint_v = pt.joint_v;
            T.n_u = pt.n_u;
            T.n_v = pt.n_v;
            T.lo_u = (float)pt.lo_u;
            T.lo_v = (float)pt.lo_v;
            T.inv_du = (float)(pt.n_u / (pt.hi_u - pt.lo_u));
            T.inv_dv = (float)(pt.n_v / (pt.hi_v - pt.lo_v));
            T.word_offset = off;
            table_mask[pt.link_a] |= 1u << pt.link_b;
            table_mask[pt.link_b] |= 1u << pt.link_a;
        }
        // self-collision link pairs from the mask, minus the pairs a table decides
        R.n_self_pairs = 0;
        for (int a = 0; a < d.n_links; ++a)
            for (int b = a + 1; b < d.n_links; ++b) {
                const bool on = ((d.self_mask[a] >> b) & 1u) || ((d.self_mask[b] >> a) & 1u);
                if (!on || ((table_mask[a] >> b) & 1u)) continue;
                if (R.link_frame[a] == 0 && R.link_frame[b] == 0) continue;  // static pair
                if (R.link_sphere_begin[a + 1] == R.link_sphere_begin[a] ||
                    R.link_sphere_begin[b + 1] == R.link_sphere_begin[b])
                    continue;
                R.self_pair[R.n_self_pairs++] = make_uchar2((unsigned char)a, (unsigned char)b);
            }
        // per-thread column layout of the collision kernels: which frames / link centres are read
        // back after the FK sweep (mode 0: by the self pairs; mode 1: by robot-robot tests too)
        for (int mode = 0; mode < 2; ++mode) {
            bool need_link[MMMP_MAX_LINKS] = {false};
            for (int l = 0; l < d.n_links; ++l)
                need_link[l] = mode == 1 && R.link_sphere_begin[l + 1] > R.link_sphere_begin[l];
            for (int p = 0; p < R.n_self_pairs; ++p) need_link[R.self_pair[p].x] = need_link[R.self_pair[p].y] = true;
            R.n_link_store[mode] = 0;
            for (int l = 0; l < MMMP_MAX_LINKS; ++l)
                R.link_store[mode][l] = (l < d.n_links && need_link[l]) ? (signed char)R.n_link_store[mode]++ : (signed char)-1;
            R.n_frame_store[mode] = 0;
            for (int f = 0; f < MMMP_MAX_JOINTS + 2; ++f) {
                bool need = false;
                for (int l = 0; l < d.n_links; ++l) need = need || (need_link[l] && R.link_frame[l] == f);
                R.frame_store[mode][f] = (f >= 1 && f <= d.n_joints && need) ? (signed char)R.n_frame_store[mode]++ : (signed char)-1;
            }
        }
        if (d.dof_offset + d.n_joints > dof) dof = d.dof_offset + d.n_joints;
    }
    H->dof = dof;
    H->n_table_words = table_words;
    // planes first (a verdict is an OR over the obstacles: their order is free), so that the kernels
    // can test them without the type dispatch
    int slot = 0;
    H->n_planes = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int o = 0; o < n_obstacles; ++o) {
            const int type = obstacles[o].type;
            if (type < MMMP_OBS_PLANE || type > MMMP_OBS_CYLINDER) {
                free(H);
                return MMMP_ERR_KIND;
            }
            if ((type == MMMP_OBS_PLANE) != (pass == 0)) continue;
            ObstacleDev& O = H->obstacles[slot++];
            O.type = type;
            O.body = obstacles[o].body;
            O.pad[0] = O.pad[1] = 0;
            for (int i = 0; i < 16; ++i) O.p[i] = (float)obstacles[o].p[i];
            if (pass == 0) H->n_planes++;
        }
    mmmp_world* w = new (std::nothrow) mmmp_world();
    if (!w) {
        free(H);
        return MMMP_ERR_ARG;
    }
    w->kind = MMMP_WORLD_SPATIAL;
    w->spatial_h = H;
    w->spatial_d = nullptr;
    w->planar_d = nullptr;
    w->planar_h = nullptr;
    cudaError_t e = cudaGetDevice(&w->device);
    if (e == cudaSuccess) e = cudaMalloc(&w->spatial_d, sizeof(WorldDev));
    if (e == cudaSuccess) e = cudaMemcpy(w->spatial_d, H, sizeof(WorldDev), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        MMMP_LAUNCH(world_static_kernel, 1, 32, 0, 0, w->spatial_d);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess)
        e = cudaMemcpy(H, w->spatial_d, sizeof(WorldDev), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        if (w->spatial_d) cudaFree(w->spatial_d);
        free(H);
        delete w;
        return MMMP_ERR_CUDA_BASE + (int)e;
    }
    *out = w;
    return MMMP_OK;
}

extern "C" int mmmp_planar_world_create(const mmmp_planar_desc* d, mmmp_world** out) {
    if (!out || !d) return MMMP_ERR_ARG;
    if (d->n_arms < 1 || d->n_arms > MMMP_MAX_ROBOTS || d->n_obstacles < 0 ||
        d->n_obstacles > MMMP_MAX_OBSTACLES)
        return MMMP_ERR_LIMIT;
    PlanarDev* H = (PlanarDev*)calloc(1, sizeof(PlanarDev));
    if (!H) return MMMP_ERR_ARG;
    H->n_arms = d->n_arms;
    H->n_obstacles = d->n_obstacles;
    int dof = 0, links = 0;
    for (int a = 0; a < d->n_arms; ++a) {
        if (d->n_links[a] < 1 || d->n_links[a] > MMMP_MAX_JOINTS) {
            free(H);
            return MMMP_ERR_LIMIT;
        }
        H->n_links[a] = d->n_links[a];
        H->dof_offset[a] = d->dof_offset[a];
        H->link_begin[a] = links;
        links += d->n_links[a];
        if (d->dof_offset[a] + d->n_links[a] > dof) dof = d->dof_offset[a] + d->n_links[a];
        for (int l = 0; l < d->n_links[a]; ++l) H->link_len[a][l] = d->link_len[a][l];
        H->base[a][0] = d->base[a][0];
        H->base[a][1] = d->base[a][1];
    }
    if (links > MMMP_PLANAR_MAX_LINKS) {
        free(H);
        return MMMP_ERR_LIMIT;
    }
    H->link_begin[d->n_arms] = links;
    H->total_links = links;
    H->dof = dof;
    for (int i = 0; i < 4; ++i) H->bounds[i] = d->bounds[i];
    for (int o = 0; o < d->n_obstacles; ++o) {
        if (d->obs_type[o] != MMMP_POBS_RECT && d->obs_type[o] != MMMP_POBS_CIRCLE) {
            free(H);
            return MMMP_ERR_KIND;
        }
        H->obs_type[o] = d->obs_type[o];
        for (int i = 0; i < 4; ++i) H->obs[o][i] = d->obs[o][i];
    }
    H->joint_link_radius = d->joint_link_radius;
    mmmp_world* w = new (std::nothrow) mmmp_world();
    if (!w) {
        free(H);
        return MMMP_ERR_ARG;
    }
    w->kind = MMMP_WORLD_PLANAR;
    w->planar_h = H;
    w->spatial_d = nullptr;
    w->spatial_h = nullptr;
    cudaError_t e = cudaGetDevice(&w->device);
    if (e == cudaSuccess) e = cudaMalloc(&w->planar_d, sizeof(PlanarDev));
    if (e == cudaSuccess) e = cudaMemcpy(w->planar_d, H, sizeof(PlanarDev), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (w->planar_d) cudaFree(w->planar_d);
        free(H);
        delete w;
        return MMMP_ERR_CUDA_BASE + (int)e;
    }
    *out = w;
    return MMMP_OK;
}

extern "C" int mmmp_world_destroy(mmmp_world* w) {
    if (!w) return MMMP_OK;
    if (w->spatial_d) cudaFree(w->spatial_d);
    if (w->planar_d) cudaFree(w->planar_d);
    free(w->spatial_h);
    free(w->planar_h);
    delete w;
    return MMMP_OK;
}

extern "C" int mmmp_world_dof(const mmmp_world* w) {
    if (!w) return -1;
    return w->kind == MMMP_WORLD_SPATIAL ? w->spatial_h->dof : w->planar_h->dof;
}

extern "C" int mmmp_version(void) { return 100; }

extern "C" int mmmp_device_info(int* sm_count, int* cc_major, int* cc_minor, int* clock_khz) {
    int dev = 0;
    MMMP_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    MMMP_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    if (clock_khz) *clock_khz = khz;
    return MMMP_OK;
}

extern "C" int64_t mmmp_launch_count(void) { return g_mmmp_launches.load(); }
