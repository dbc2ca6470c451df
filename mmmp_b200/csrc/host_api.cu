// Host-buffer entry points: the call a reference-side planner makes with numpy
// arrays (host memory in, host memory out), and the native host replay of the
// order-dependent greedy accept pass.
//
// Reference: DecoupledPRM.generate_roadmap decoupled_prm.py:231-260,
// sample_valid_in_joint_space_random :310-321, connect_nearest_neighbors_degree
// :480-504, is_collision_free_sample / _edge :840-863.
#include <cstring>
#include <vector>
#include "common.cuh"

namespace {

// edges[e] = (e / k, nbr[e]) for the flat [n, k] neighbour table
__global__ void __launch_bounds__(256)
edges_from_knn_kernel(const int32_t* __restrict__ nbr, int64_t n, int k, int32_t* __restrict__ edges) {
    const int64_t total = n * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        edges[2 * e] = (int32_t)(e / k);
        edges[2 * e + 1] = nbr[e];
    }
}

// *out |= 1 if any byte of f is non-zero
__global__ void __launch_bounds__(256)
any_flag_kernel(const uint8_t* __restrict__ f, int64_t n, int* __restrict__ out) {
    bool any = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        any = any || f[i] != 0;
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(out, 1);
}

thread_local cudaStream_t tl_alloc_stream = nullptr;  // stream the host call allocates on

struct DevBuf {  // stream-ordered temporaries from the (cached) default pool
    void* p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t bytes) {
        s = tl_alloc_stream;
        return cudaMallocAsync(&p, bytes ? bytes : 16, s);
    }
    ~DevBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    template <class T>
    T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace

extern "C" int mmmp_edges_from_knn(const int32_t* nbr_d, int64_t n, int k, int32_t* edges_d, void* stream) {
    if (!nbr_d || !edges_d || n < 0 || k < 1) return MMMP_ERR_ARG;
    if (n == 0) return MMMP_OK;
    int g = mmmp_div_up(n * k, 256);
    if (g > 148 * 16) g = 148 * 16;
    MMMP_LAUNCH(edges_from_knn_kernel, g, 256, 0, stream, nbr_d, n, k, edges_d);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}


// Frames whose origin depends on q, with coincident consecutive frames merged
// (multiplicity = group weight).  Frame f (1-based, tip = n_joints+1) sits at
// p_f = p_{f-1} + R_{f-1} t_{f-1}; it is q-independent while t_1..t_{f-1} are all
// zero and coincides with frame f-1 when t_{f-1} is zero.  Panda (robots/panda.py:31-40):
// frames 3,4,6(x2),7,8 -> weights 1,1,2,1,1.
extern "C" int mmmp_fk_metric_groups(const mmmp_world* w, int robot, int32_t* frames, float* weights, int* n_groups) {
    if (!w || !frames || !weights || !n_groups) return MMMP_ERR_ARG;
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    if (robot < 0 || robot >= w->spatial_h->n_robots) return MMMP_ERR_ARG;
    const RobotDev64& C = w->spatial_h->robots64[robot];
    const int L = C.n_joints + 1;
    auto zero_t = [&](int j) { return C.origin[j][3] == 0.0 && C.origin[j][7] == 0.0 && C.origin[j][11] == 0.0; };
    int ng = 0;
    for (int f = 1; f <= L; ++f) {
        bool moves = false;
        for (int g = 1; g < f; ++g)
            if (!zero_t(g)) moves = true;
        if (!moves) continue;
        if (zero_t(f - 1) && ng > 0 && frames[ng - 1] == f - 1) {
            frames[ng - 1] = f;
            weights[ng - 1] += 1.f;
            continue;
        }
        if (ng == MMMP_MAX_GROUPS) return MMMP_ERR_LIMIT;
        frames[ng] = f;
        weights[ng] = 1.f;
        ++ng;
    }
    if (ng == 0) return MMMP_ERR_ARG;
    *n_groups = ng;
    return MMMP_OK;
}

extern "C" int mmmp_greedy_degree_connect_device(const int32_t* nbr_idx_d, const uint8_t* edge_free_d, int64_t N,
                                                 int k1, int k2, int cap, int32_t* adj_d, int32_t* deg_d,
                                                 int32_t* rounds_d, void* stream);
extern "C" int mmmp_connected_components(const int32_t* adj_d, const uint8_t* mask_d, int64_t N, int cap,
                                         int32_t* labels_d, int64_t* n_components_d, void* stream);

// k2 > 0: also the accepted adjacency (greedy degree pass) and its components
static int roadmap_host_impl(const mmmp_world* w, int robot, const double* q64_h, int64_t N, int D, int metric_kind,
                             int k, int n_steps, double step, uint32_t flags, int k2, int64_t* n_nodes_h,
                             int32_t* node_src_h, int32_t* nbr_idx_h, double* nbr_dist_h, uint8_t* edge_free_h,
                             int32_t* adj_h, int32_t* deg_h, int32_t* labels_h, int64_t* n_components_h,
                             float* timings_ms_h) {
    if (!w || !q64_h || !n_nodes_h || !node_src_h || !nbr_idx_h || !edge_free_h) return MMMP_ERR_ARG;
    if (w->kind != MMMP_WORLD_SPATIAL) return MMMP_ERR_KIND;
    if (N < 1 || D < 1 || k < 1 || k > MMMP_MAX_K) return MMMP_ERR_ARG;
    if (k2 < 0 || (k2 > 0 && (!adj_h || !deg_h))) return MMMP_ERR_ARG;
    {
        int cur = -1;
        MMMP_CUDA_TRY(cudaGetDevice(&cur));
        if (cur != w->device) return MMMP_ERR_ARG;   // the world's tables live on another device
    }
    if (robot < 0 || robot >= w->spatial_h->n_robots) return MMMP_ERR_ARG;
    const int nj = w->spatial_h->robots[robot].n_joints;
    if (D != nj) return MMMP_ERR_ARG;
    if (metric_kind != MMMP_METRIC_L2 && metric_kind != MMMP_METRIC_GROUP_L2_SUM) return MMMP_ERR_KIND;
    auto row_width = [](int d) { return d <= 4 ? 4 : (d <= 8 ? 8 : (d <= 16 ? 16 : 32)); };
    const int ld32 = row_width(D);
    // the fp32 scan keeps k + 2 candidates for the fp64 re-rank; if that cannot separate rank k
    // from the last candidate anywhere (flagged `ambiguous`, ~1e-8 per query) the scan is repeated
    // once with k + 16
    int kc = k + 2;
    if (kc > MMMP_MAX_K) kc = MMMP_MAX_K;

    mmmp_pool_keep();
    cudaStream_t s, s2;   // s: compute; s2: device-to-host copies that overlap the next stage
    MMMP_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaError_t e2 = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    if (e2 != cudaSuccess) {
        cudaStreamDestroy(s);
        return MMMP_ERR_CUDA_BASE + (int)e2;
    }
    tl_alloc_stream = s;
    cudaEvent_t ev[11], ready;
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
    int rc = MMMP_OK;
    int64_t n_nodes = 0;
    {
        DevBuf q64, q32, coll, idx, cnt, nq64, nq32, feat, filt, pos64, cidx, cdist, ridx, rdist, ecoll, amb, adj,
            deg, labels, ncomp;
        const int L = nj + 1;
        auto t_at = [&](cudaError_t e, int line) {
            if (e != cudaSuccess && rc == MMMP_OK) {
                mmmp_note_error(__FILE__, line, (int)e);
                rc = MMMP_ERR_CUDA_BASE + (int)e;
            }
        };
        auto r_at = [&](int r, int line) {
            if (r != MMMP_OK && rc == MMMP_OK) {
                mmmp_note_error(__FILE__, line, r >= MMMP_ERR_CUDA_BASE ? r - MMMP_ERR_CUDA_BASE : -r);
                rc = r;
            }
        };
#define T(e) t_at((e), __LINE__)
#define R(r) r_at((r), __LINE__)
        T(q64.alloc(sizeof(double) * N * D));
        T(q32.alloc(sizeof(float) * N * ld32));
        T(coll.alloc(N));
        T(idx.alloc(sizeof(int32_t) * N));
        T(cnt.alloc(sizeof(int32_t)));
        if (rc == MMMP_OK) {
            cudaEventRecord(ev[0], s);
            T(cudaMemcpyAsync(q64.p, q64_h, sizeof(double) * N * D, cudaMemcpyHostToDevice, s));
            cudaEventRecord(ev[1], s);
            R(mmmp_pack_f32(q64.as<double>(), D, D, N, q32.as<float>(), ld32, s));
            R(mmmp_collide_configs(w, robot, q32.p, ld32, N, flags, coll.as<uint8_t>(), s));
            R(mmmp_compact_free(coll.as<uint8_t>(), N, idx.as<int32_t>(), cnt.as<int32_t>(), s));
            int32_t c = 0;
            T(cudaMemcpyAsync(&c, cnt.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            T(cudaStreamSynchronize(s));
            n_nodes = c;
        }
        if (rc == MMMP_OK && n_nodes > 0) {
            const int64_t n = n_nodes;
            T(nq64.alloc(sizeof(double) * n * D));
            T(nq32.alloc(sizeof(float) * n * ld32));
            const int kc_wide = k + 16 > MMMP_MAX_K ? MMMP_MAX_K : k + 16;
            T(cidx.alloc(sizeof(int32_t) * n * kc_wide));
            T(cdist.alloc(sizeof(float) * n * kc_wide));
            T(amb.alloc(((n + 3) & ~(int64_t)3) + sizeof(int)));
            T(ridx.alloc(sizeof(int32_t) * n * k));
            T(rdist.alloc(sizeof(double) * n * k));
            T(ecoll.alloc(n * k));
            mmmp_metric m32, m64;
            memset(&m32, 0, sizeof(m32));
            memset(&m64, 0, sizeof(m64));
            const float* feat_ptr = nullptr;
            const float* filt_ptr = nullptr;
            int ldfilt = 0;
            const double* ref64_ptr = nullptr;
            int ldf = ld32, ld64 = D;
            if (rc == MMMP_OK) {
                R(mmmp_gather_rows(q64.p, (int)(sizeof(double) * D), idx.as<int32_t>(), n, nq64.p, s));
                R(mmmp_gather_rows(q32.p, (int)(sizeof(float) * ld32), idx.as<int32_t>(), n, nq32.p, s));
                cudaEventRecord(ev[2], s);
                if (metric_kind == MMMP_METRIC_L2) {
                    m32.kind = m64.kind = MMMP_METRIC_L2;
                    m32.dim = m64.dim = D;
                    feat_ptr = nq32.as<float>();
                    ref64_ptr = nq64.as<double>();
                } else {
                    int frames[MMMP_MAX_GROUPS];
                    float wts[MMMP_MAX_GROUPS];
                    int ng = 0;
                    R(mmmp_fk_metric_groups(w, robot, frames, wts, &ng));
                    if (rc == MMMP_OK && ng == 0) rc = MMMP_ERR_ARG;
                    ldf = row_width(3 * ng);
                    m32.kind = MMMP_METRIC_GROUP_L2_SUM;
                    m32.dim = 3 * ng;
                    m32.n_groups = ng;
                    for (int g = 0; g < ng; ++g) m32.weight[g] = wts[g];
                    m64.kind = MMMP_METRIC_GROUP_L2_SUM;
                    m64.dim = 3 * L;
                    m64.n_groups = L;
                    ld64 = 3 * L;
                    T(feat.alloc(sizeof(float) * n * ldf));
                    T(pos64.alloc(sizeof(double) * n * 3 * L));
                    if (rc == MMMP_OK) {
                        T(filt.alloc(sizeof(float) * n * 8));
                        R(mmmp_fk_features(w, robot, nq32.as<float>(), ld32, n, frames, wts, ng, feat.as<float>(), ldf,
                                           filt.as<float>(), 8, s));      // dual filter rows
                        filt_ptr = filt.as<float>();
                        ldfilt = 8;
                        R(mmmp_fk_chain_f64(w, robot, nq64.as<double>(), D, n, pos64.as<double>(), s));
                    }
                    feat_ptr = feat.as<float>();
                    ref64_ptr = pos64.as<double>();
                }
                cudaEventRecord(ev[3], s);
            }
            if (rc == MMMP_OK) {
                if (!filt_ptr) {  // L2: the rows are their own filter
                    filt_ptr = feat_ptr;
                    ldfilt = ldf;
                }
                int* any_amb = reinterpret_cast<int*>(amb.as<uint8_t>() + ((n + 3) & ~(int64_t)3));
                for (int attempt = 0; attempt < 2 && rc == MMMP_OK; ++attempt) {
                    R(mmmp_knn(filt_ptr, feat_ptr, n, filt_ptr, feat_ptr, n, ldfilt,
                               filt_ptr != feat_ptr ? 6 : m32.dim, ldf, &m32, kc, 1, 0, 0,
                               cidx.as<int32_t>(), cdist.as<float>(), s));
                    if (attempt == 0) cudaEventRecord(ev[4], s);
                    R(mmmp_knn_refine(ref64_ptr, ref64_ptr, ld64, &m64, cidx.as<int32_t>(), cdist.as<float>(), n, kc, k,
                                      1e-6, ridx.as<int32_t>(), rdist.as<double>(), amb.as<uint8_t>(), s));
                    if (kc >= kc_wide) break;
                    int flagged = 0;
                    T(cudaMemsetAsync(any_amb, 0, sizeof(int), s));
                    int g = mmmp_div_up(n, 256);
                    MMMP_LAUNCH(any_flag_kernel, g > 148 * 8 ? 148 * 8 : g, 256, 0, s, amb.as<uint8_t>(), n, any_amb);
                    T(cudaPeekAtLastError());
                    T(cudaMemcpyAsync(&flagged, any_amb, sizeof(int), cudaMemcpyDeviceToHost, s));
                    T(cudaStreamSynchronize(s));
                    if (!flagged) break;
                    kc = kc_wide;
                }
                cudaEventRecord(ev[5], s);
                // the neighbour tables go home on the copy stream while the edges are checked
                cudaEventRecord(ready, s);
                cudaStreamWaitEvent(s2, ready, 0);
                T(cudaMemcpyAsync(node_src_h, idx.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s2));
                T(cudaMemcpyAsync(nbr_idx_h, ridx.p, sizeof(int32_t) * n * k, cudaMemcpyDeviceToHost, s2));
                if (nbr_dist_h) T(cudaMemcpyAsync(nbr_dist_h, rdist.p, sizeof(double) * n * k, cudaMemcpyDeviceToHost, s2));
                R(mmmp_collide_knn_edges(w, robot, nq32.p, ld32, ridx.as<int32_t>(), n, k, 0, n, 0, n_steps, step,
                                         flags, ecoll.as<uint8_t>(), s));
                cudaEventRecord(ev[6], s);
                // ... and the verdicts while the accept pass runs
                cudaEventRecord(ready, s);
                cudaStreamWaitEvent(s2, ready, 0);
                T(cudaMemcpyAsync(edge_free_h, ecoll.p, (size_t)n * k, cudaMemcpyDeviceToHost, s2));
                if (k2 > 0 && rc == MMMP_OK) {
                    T(adj.alloc(sizeof(int32_t) * n * k2));
                    T(deg.alloc(sizeof(int32_t) * n));
                    T(labels.alloc(sizeof(int32_t) * n));
                    T(ncomp.alloc(sizeof(int64_t)));
                    if (rc == MMMP_OK) {
                        R(mmmp_greedy_degree_connect_device(ridx.as<int32_t>(), ecoll.as<uint8_t>(), n, k, k2, k2,
                                                            adj.as<int32_t>(), deg.as<int32_t>(), nullptr, s));
                        cudaEventRecord(ev[7], s);
                        R(mmmp_connected_components(adj.as<int32_t>(), nullptr, n, k2, labels.as<int32_t>(),
                                                    ncomp.as<int64_t>(), s));
                        cudaEventRecord(ev[8], s);
                        T(cudaMemcpyAsync(adj_h, adj.p, sizeof(int32_t) * n * k2, cudaMemcpyDeviceToHost, s));
                        T(cudaMemcpyAsync(deg_h, deg.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
                        if (labels_h) T(cudaMemcpyAsync(labels_h, labels.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
                        if (n_components_h)
                            T(cudaMemcpyAsync(n_components_h, ncomp.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
                    }
                } else {
                    cudaEventRecord(ev[7], s);
                    cudaEventRecord(ev[8], s);
                }
                cudaEventRecord(ready, s2);
                cudaStreamWaitEvent(s, ready, 0);        // "d2h" ends when both streams have delivered
                cudaEventRecord(ev[9], s);
                T(cudaStreamSynchronize(s));
                T(cudaStreamSynchronize(s2));
                if (rc == MMMP_OK && timings_ms_h) {
                    // h2d, accept test + compaction, features, scan, re-rank, edge collision, [6] d2h tail,
                    // [7] total, [8] accept pass, [9] components
                    for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&timings_ms_h[i], ev[i], ev[i + 1]);
                    cudaEventElapsedTime(&timings_ms_h[6], ev[8], ev[9]);
                    cudaEventElapsedTime(&timings_ms_h[7], ev[0], ev[9]);
                    cudaEventElapsedTime(&timings_ms_h[8], ev[6], ev[7]);
                    cudaEventElapsedTime(&timings_ms_h[9], ev[7], ev[8]);
                }
            }
        }
    }
    *n_nodes_h = n_nodes;
    for (auto& e : ev) cudaEventDestroy(e);
    cudaEventDestroy(ready);
    cudaStreamDestroy(s2);
    cudaStreamDestroy(s);
    return rc;
}

#undef T
#undef R

extern "C" int mmmp_roadmap_candidates_host(const mmmp_world* w, int robot, const double* q64_h, int64_t N, int D,
                                            int metric_kind, int k, int n_steps, double step, uint32_t flags,
                                            int64_t* n_nodes_h, int32_t* node_src_h, int32_t* nbr_idx_h,
                                            double* nbr_dist_h, uint8_t* edge_free_h, float* timings_ms_h) {
    float t[10] = {0};
    const int rc = roadmap_host_impl(w, robot, q64_h, N, D, metric_kind, k, n_steps, step, flags, 0, n_nodes_h, node_src_h,
                                     nbr_idx_h, nbr_dist_h, edge_free_h, nullptr, nullptr, nullptr, nullptr,
                                     timings_ms_h ? t : nullptr);
    if (timings_ms_h)
        for (int i = 0; i < 8; ++i) timings_ms_h[i] = t[i];
    return rc;
}

extern "C" int mmmp_roadmap_build_host(const mmmp_world* w, int robot, const double* q64_h, int64_t N, int D,
                                       int metric_kind, int k1, int k2, int n_steps, double step, uint32_t flags,
                                       int64_t* n_nodes_h, int32_t* node_src_h, int32_t* nbr_idx_h, double* nbr_dist_h,
                                       uint8_t* edge_free_h, int32_t* adj_h, int32_t* deg_h, int32_t* labels_h,
                                       int64_t* n_components_h, float* timings_ms_h) {
    if (k2 < 1) return MMMP_ERR_ARG;
    return roadmap_host_impl(w, robot, q64_h, N, D, metric_kind, k1, n_steps, step, flags, k2, n_nodes_h, node_src_h,
                             nbr_idx_h, nbr_dist_h, edge_free_h, adj_h, deg_h, labels_h, n_components_h, timings_ms_h);
}

// Greedy accept pass, reference decoupled_prm.py:480-504: nodes in id order,
// candidates ascending (dist,id); stop when deg[node]==k2; accept iff the edge
// is free, not already present, and (cap_other_end) deg[neighbour] != k2.
extern "C" int mmmp_greedy_degree_connect(const int32_t* nbr_idx, const uint8_t* edge_free, int64_t N, int k1,
                                          int k2, int cap, int32_t* adj, int32_t* deg) {
    if (!nbr_idx || !edge_free || !adj || !deg || N < 0 || k1 < 1 || k2 < 1 || cap < k2) return MMMP_ERR_ARG;
    for (int64_t i = 0; i < N; ++i) deg[i] = 0;
    for (int64_t i = 0; i < N * (int64_t)cap; ++i) adj[i] = -1;
    for (int64_t i = 0; i < N; ++i) {
        for (int c = 0; c < k1; ++c) {
            if (deg[i] == k2) break;
            const int32_t j = nbr_idx[i * k1 + c];
            if (j < 0 || j >= N || j == i) continue;
            if (!edge_free[i * k1 + c]) continue;
            bool present = false;
            for (int t = 0; t < deg[i]; ++t)
                if (adj[i * cap + t] == j) {
                    present = true;
                    break;
                }
            if (present) continue;
            if (deg[j] == k2) continue;
            if (deg[i] >= cap || deg[j] >= cap) return MMMP_ERR_LIMIT;
            adj[i * cap + deg[i]++] = j;
            adj[(int64_t)j * cap + deg[j]++] = (int32_t)i;
        }
    }
    return MMMP_OK;
}
