// Kernel (d), scan stage: brute-force k nearest candidates in fp32 on the CUDA cores.
//
// Every (query, reference) pair is visited.  The hot loop evaluates a cheap LOWER BOUND
// of the metric as a squared L2 distance between "filter rows" in dot-product form
//        s = B_r - 2 <a, r>   <=   tau_f - A_q          (B_r, A_q = squared row norms)
// (dim FFMAs + a share of one FMNMX tree / compare per pair):
//   L2 / SQ_SUM metrics : filter row = the row itself (the bound is the metric);
//   FK metric sum_g w_g |dp_g| (robots/panda.py:248-254): filter row = the weighted
//       centroid sum_g w_g p_g (3 floats) -- triangle inequality
//       sum_g w_g |dp_g| >= | sum_g w_g dp_g |.
// A prepare kernel packs the filter rows as [x_0 .. x_{d-1}, 0 .., B_r] (B_r in the last
// of W = 4/8/16/32 columns), so the loop needs nothing but the row itself.
//
// CTA = 4 consumer warps + 1 producer warp (warp specialisation).  The producer streams
// the packed reference rows L2/HBM -> shared memory in contiguous tiles with the TMA
// engine (cp.async.bulk, 4-stage ring, full/empty mbarriers); consumer warps never meet
// at a CTA barrier.  Every consumer thread owns Q queries in registers and reads the
// tile rows as shared-memory broadcasts (LDS.128), 4 rows per step.  Pairs that pass the
// filter are appended to a small per-thread queue in shared memory; when any lane's
// queue is nearly full the warp drains it: exact fp32 metric on the "exact rows" (read
// from L2/HBM) and sorted insertion into the thread's candidate column, which tightens
// the thread's threshold.
//
// Reference call sites replaced: the heap loops of
// planners/prm/pybullet/decoupled/decoupled_prm.py:484-494,531-540,947-954 and
// KDTree.query at coupled/degree_coupled_prm.py:67, planar single_arm/degree_prm.py:178.
#include <math_constants.h>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int KNN_CWARPS = 4;                  // consumer warps per CTA
constexpr int KNN_CTHREADS = KNN_CWARPS * 32;  // consumer threads
constexpr int KNN_THREADS = KNN_CTHREADS + 32; // + producer warp
constexpr int KNN_STAGES = 4;
constexpr int KNN_TILE = 256;                  // reference rows per shared-memory tile
constexpr int KNN_RU = 4;                      // rows per inner-loop step
constexpr int KNN_QSLACK = 16;                 // queue entries a thread may hold before its warp drains
__host__ __device__ constexpr int knn_qcap(int Q) { return KNN_RU * Q + KNN_QSLACK; }
constexpr float KNN_SLACK = 2.0e-5f;           // relative slack of the fp32 filter

struct MetricDev {
    int kind, dim, n_groups, pad;
    float w[MMMP_MAX_GROUPS];
};

// ---- mbarrier / TMA bulk copy (sm_90+; sm_100a here), 32-bit shared addresses --------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"  // %2: suspend-time hint (ns)
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// ---- filter-row packing -----------------------------------------------------------------
// out[i] = [in[i][0..d), 0 .., (1 - eps) * |in[i][0..d)|^2]   (W columns)
__global__ void __launch_bounds__(256)
knn_pack_kernel(const float* __restrict__ in, int ld_in, int d, int64_t n, float* __restrict__ out, int W) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float* r = in + i * ld_in;
        float* o = out + i * W;
        float acc = 0.f;
        for (int c = 0; c < W - 1; ++c) {
            const float v = c < d ? r[c] : 0.f;
            acc = fmaf(v, v, acc);
            o[c] = v;
        }
        o[W - 1] = acc * (1.f - KNN_SLACK);
    }
}

struct KnnArgs {
    const float* pref;    // packed filter rows of the references [n_ref, W]
    const float* pquery;  // packed filter rows of the queries [n_query, W]
    const float* xref;    // exact rows [n_ref, ldx]
    const float* xquery;  // exact rows [n_query, ldx]
    int64_t n_ref, n_query, query_id0;
    int ldx, kc, flags;   // flags: 1 = exclude self, 2 = causal
    int tiles_per_split, n_splits;
    int32_t* out_idx;     // [n_query, n_splits, kc]
    float* out_val;       // candidate-list domain (L2/SQ_SUM: squared ; GROUP: distance)
    unsigned long long* stats;  // optional: [0] += exact evaluations, [1] += insertions, [2] += drains
    MetricDev M;
};

// exact fp32 metric in the candidate list's domain
__device__ __forceinline__ float exact_metric(const MetricDev& M, const float* __restrict__ a,
                                              const float* __restrict__ b) {
    float v = 0.f;
    if (M.kind == MMMP_METRIC_GROUP_L2_SUM) {
        for (int g = 0; g < M.n_groups; ++g) {
            const float dx = a[3 * g] - b[3 * g], dy = a[3 * g + 1] - b[3 * g + 1], dz = a[3 * g + 2] - b[3 * g + 2];
            v = fmaf(M.w[g], sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz))), v);
        }
    } else {
        for (int i = 0; i < M.dim; ++i) {
            const float d = a[i] - b[i];
            v = fmaf(d, d, v);
        }
    }
    return v;
}

template <int F4, int Q>
__global__ void __launch_bounds__(KNN_THREADS)
knn_scan_kernel(const __grid_constant__ KnnArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int W = F4 * 4;
    constexpr int LS = KNN_CTHREADS * Q;  // column stride of the per-query shared arrays
    constexpr int QCAP = knn_qcap(Q);
    // layout: tiles[STAGES][TILE][W] | lv[kc][LS] | li[kc][LS] | queue[QCAP][CTHREADS] | full[STAGES] empty[STAGES]
    float* tiles = reinterpret_cast<float*>(smem_raw);
    float* lv = tiles + KNN_STAGES * KNN_TILE * W;
    int* li = reinterpret_cast<int*>(lv + (size_t)A.kc * LS);
    uint32_t* queue = reinterpret_cast<uint32_t*>(li + (size_t)A.kc * LS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(queue + QCAP * KNN_CTHREADS);
    const uint32_t tiles_s = smem_u32(tiles);
    const uint32_t full_s = smem_u32(bars), empty_s = smem_u32(bars + KNN_STAGES);

    const int tid = threadIdx.x;
    const int kc = A.kc;
    if (tid == 0) {
        for (int s = 0; s < KNN_STAGES; ++s) {
            mbar_init(full_s + 8 * s, 1);
            mbar_init(empty_s + 8 * s, KNN_CWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- reference tiles of this CTA ---------------------------------------------
    const int64_t q_first = (int64_t)blockIdx.x * LS;
    int64_t ref_end = A.n_ref;
    if (A.flags & 2) {
        const int64_t last_q = A.query_id0 + q_first + LS;  // exclusive bound on this CTA's query ids
        if (last_q < ref_end) ref_end = last_q;
    }
    const int64_t total_tiles = (ref_end + KNN_TILE - 1) / KNN_TILE;
    const int64_t tile_lo = (int64_t)blockIdx.y * A.tiles_per_split;
    int64_t tile_hi = tile_lo + A.tiles_per_split;
    if (tile_hi > total_tiles) tile_hi = total_tiles;
    const int n_tiles = (int)(tile_hi > tile_lo ? tile_hi - tile_lo : 0);
    __syncthreads();

    if (tid >= KNN_CTHREADS) {
        // ================= producer warp: one elected lane drives the TMA ring =================
        if (tid == KNN_CTHREADS) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % KNN_STAGES;
                mbar_wait(empty_s + 8 * s, ((t / KNN_STAGES) & 1) ^ 1);
                const int64_t r0 = (tile_lo + t) * KNN_TILE;
                int64_t cnt = ref_end - r0;
                if (cnt > KNN_TILE) cnt = KNN_TILE;
                const uint32_t bytes = (uint32_t)(cnt * W * sizeof(float));
                mbar_expect_tx(full_s + 8 * s, bytes);
                tma_load_1d(tiles_s + (uint32_t)s * KNN_TILE * W * 4, A.pref + r0 * W, bytes, full_s + 8 * s);
            }
        }
        return;
    }

    // ================= consumer warps =================
    // queries: filter form in registers (a2 = -2 x, last column of the packed row = A_q)
    float a2[Q][W - 1], thr[Q], Aq[Q];
#pragma unroll
    for (int qq = 0; qq < Q; ++qq) {
        const int64_t qi = q_first + qq * KNN_CTHREADS + tid;
        const bool valid = qi < A.n_query;
#pragma unroll
        for (int i = 0; i < W - 1; ++i) a2[qq][i] = valid ? -2.f * __ldg(A.pquery + qi * W + i) : 0.f;
        Aq[qq] = valid ? __ldg(A.pquery + qi * W + (W - 1)) : 0.f;
        thr[qq] = valid ? CUDART_INF_F : -CUDART_INF_F;
        for (int c = 0; c < kc; ++c) {
            lv[c * LS + qq * KNN_CTHREADS + tid] = CUDART_INF_F;
            li[c * LS + qq * KNN_CTHREADS + tid] = -1;
        }
    }
    int qcnt = 0;
    uint32_t* myq = queue + tid;

    // Drain: every lane evaluates its queued (query, row) pairs exactly and inserts.
    auto drain = [&]() {
        int maxc = qcnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, o));
        for (int e = 0; e < maxc; ++e) {
            if (e < qcnt) {
                const uint32_t ent = myq[e * KNN_CTHREADS];
                const int qq = (int)(ent >> 28);
                const long long id = (long long)(ent & 0x0fffffffu);
                const int col = qq * KNN_CTHREADS + tid;
                const long long qi = q_first + col;
                const long long qid = A.query_id0 + qi;
                bool ok = id < ref_end;
                if ((A.flags & 1) && id == qid) ok = false;
                if ((A.flags & 2) && id >= qid) ok = false;
                if (ok) {
                    const float v = exact_metric(A.M, A.xquery + qi * A.ldx, A.xref + id * A.ldx);
                    float* lvc = lv + col;
                    int* lic = li + col;
                    if (v < lvc[(kc - 1) * LS]) {
                        int p = kc - 1;
                        while (p > 0) {
                            const float u = lvc[(p - 1) * LS];
                            if (!(u > v)) break;
                            lvc[p * LS] = u;
                            lic[p * LS] = lic[(p - 1) * LS];
                            --p;
                        }
                        lvc[p * LS] = v;
                        lic[p * LS] = (int)id;
                        const float tau = lvc[(kc - 1) * LS];
                        if (tau < CUDART_INF_F) {
                            const float tf = (A.M.kind == MMMP_METRIC_GROUP_L2_SUM) ? tau * tau : tau;
#pragma unroll
                            for (int k2 = 0; k2 < Q; ++k2)
                                if (k2 == qq) thr[k2] = tf * (1.f + KNN_SLACK) - Aq[k2];
                        }
                        if (A.stats) atomicAdd(A.stats + 1, 1ull);
                    }
                    if (A.stats) atomicAdd(A.stats, 1ull);
                }
            }
        }
        if (A.stats && (tid & 31) == 0) atomicAdd(A.stats + 2, 1ull);
        qcnt = 0;
    };

    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % KNN_STAGES;
        const int64_t r0 = (tile_lo + t) * KNN_TILE;
        const int cnt = (int)((ref_end - r0) < KNN_TILE ? (ref_end - r0) : KNN_TILE);
        mbar_wait(full_s + 8 * s, (t / KNN_STAGES) & 1);
        const uint32_t tile_s = tiles_s + (uint32_t)s * KNN_TILE * W * 4;
        const int cnt4 = (cnt + KNN_RU - 1) & ~(KNN_RU - 1);  // stale rows of a partial tile are rejected by id
        const uint32_t id0 = (uint32_t)r0;
        // software pipeline: the rows of step j+1 are loaded while step j is evaluated
        float4 rn[KNN_RU][F4];
#pragma unroll
        for (int u = 0; u < KNN_RU; ++u)
#pragma unroll
            for (int i = 0; i < F4; ++i) rn[u][i] = lds128(tile_s + (uint32_t)(u * W + 4 * i) * 4);
#pragma unroll 1
        for (int j = 0; j < cnt4; j += KNN_RU) {
            float4 r[KNN_RU][F4];
#pragma unroll
            for (int u = 0; u < KNN_RU; ++u)
#pragma unroll
                for (int i = 0; i < F4; ++i) r[u][i] = rn[u][i];
            // (the last prefetch of a tile wraps to the tile's first rows: harmless)
            const int jn = (j + KNN_RU < KNN_TILE) ? j + KNN_RU : 0;
#pragma unroll
            for (int u = 0; u < KNN_RU; ++u)
#pragma unroll
                for (int i = 0; i < F4; ++i) rn[u][i] = lds128(tile_s + (uint32_t)((jn + u) * W + 4 * i) * 4);
            float sv[Q][KNN_RU];
            float worst = CUDART_INF_F;  // min over queries of (best filter value - threshold)
#pragma unroll
            for (int qq = 0; qq < Q; ++qq) {
#pragma unroll
                for (int u = 0; u < KNN_RU; ++u) {
                    float acc = r[u][F4 - 1].w;  // B_r
#pragma unroll
                    for (int i = 0; i < F4; ++i) {
                        acc = fmaf(a2[qq][4 * i + 0], r[u][i].x, acc);
                        acc = fmaf(a2[qq][4 * i + 1], r[u][i].y, acc);
                        acc = fmaf(a2[qq][4 * i + 2], r[u][i].z, acc);
                        if (i < F4 - 1) acc = fmaf(a2[qq][4 * i + 3], r[u][i].w, acc);
                    }
                    sv[qq][u] = acc;
                }
                const float m = fminf(fminf(sv[qq][0], sv[qq][1]), fminf(sv[qq][2], sv[qq][3]));
                worst = fminf(worst, m - thr[qq]);
            }
            // one warp-uniform branch per step; everything below it is rare
            if (__any_sync(0xffffffffu, worst <= 0.f)) {
#pragma unroll
                for (int qq = 0; qq < Q; ++qq)
#pragma unroll
                    for (int u = 0; u < KNN_RU; ++u)
                        if (sv[qq][u] <= thr[qq]) {
                            myq[qcnt * KNN_CTHREADS] = ((uint32_t)qq << 28) | (id0 + (uint32_t)(j + u));
                            ++qcnt;
                        }
                if (__any_sync(0xffffffffu, qcnt > KNN_QSLACK)) drain();
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty_s + 8 * s);
    }
    drain();

    // ---- write the candidate columns ----------------------------------------
#pragma unroll
    for (int qq = 0; qq < Q; ++qq) {
        const int64_t qi = q_first + qq * KNN_CTHREADS + tid;
        if (qi >= A.n_query) continue;
        const int col = qq * KNN_CTHREADS + tid;
        const size_t o = ((size_t)qi * A.n_splits + blockIdx.y) * kc;
        for (int c = 0; c < kc; ++c) {
            A.out_idx[o + c] = li[c * LS + col];
            A.out_val[o + c] = lv[c * LS + col];
        }
    }
}

// merge n_splits sorted candidate lists per query into one list of kc and convert the
// list-domain value to a distance
__global__ void __launch_bounds__(128)
knn_merge_kernel(const int32_t* __restrict__ in_idx, const float* __restrict__ in_val, int64_t n_query, int n_splits,
                 int kc, int kind, int32_t* __restrict__ out_idx, float* __restrict__ out_dist) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_query; q += (int64_t)gridDim.x * blockDim.x) {
        int head[32];
        for (int s = 0; s < n_splits; ++s) head[s] = 0;
        const int32_t* bi = in_idx + (size_t)q * n_splits * kc;
        const float* bv = in_val + (size_t)q * n_splits * kc;
        for (int c = 0; c < kc; ++c) {
            int best = -1;
            float bvv = CUDART_INF_F;
            int bid = -1;
            for (int s = 0; s < n_splits; ++s) {
                if (head[s] >= kc) continue;
                const int id = bi[s * kc + head[s]];
                if (id < 0) continue;
                const float v = bv[s * kc + head[s]];
                if (best < 0 || v < bvv || (v == bvv && id < bid)) {
                    best = s;
                    bvv = v;
                    bid = id;
                }
            }
            if (best >= 0) head[best]++;
            out_idx[(size_t)q * kc + c] = bid;
            if (out_dist) {
                float d = bvv;
                if (best >= 0 && kind == MMMP_METRIC_L2) d = sqrtf(bvv);
                out_dist[(size_t)q * kc + c] = d;
            }
        }
    }
}

int width_for(int d) { return d + 1 <= 4 ? 4 : (d + 1 <= 8 ? 8 : (d + 1 <= 16 ? 16 : (d + 1 <= 32 ? 32 : 0))); }

int fill_metric(const mmmp_metric* m, int ldx, MetricDev& M) {
    if (!m) return MMMP_ERR_ARG;
    M.kind = m->kind;
    M.dim = m->dim;
    M.n_groups = m->n_groups;
    M.pad = 0;
    if (m->kind < MMMP_METRIC_L2 || m->kind > MMMP_METRIC_SQ_SUM) return MMMP_ERR_KIND;
    if (m->dim < 1 || m->dim > ldx) return MMMP_ERR_ARG;
    for (int g = 0; g < MMMP_MAX_GROUPS; ++g) M.w[g] = 0.f;
    if (m->kind == MMMP_METRIC_GROUP_L2_SUM) {
        if (m->n_groups < 1 || m->n_groups > MMMP_MAX_GROUPS || m->n_groups * 3 != m->dim) return MMMP_ERR_ARG;
        for (int g = 0; g < m->n_groups; ++g) M.w[g] = m->weight[g];
    }
    return MMMP_OK;
}

size_t scan_smem(int F4, int Q, int kc) {
    const size_t w = F4 * 4, ls = (size_t)KNN_CTHREADS * Q;
    return (size_t)KNN_STAGES * KNN_TILE * w * 4 + (size_t)kc * ls * 8 + (size_t)knn_qcap(Q) * KNN_CTHREADS * 4 +
           2 * KNN_STAGES * 8 + 16;
}

template <int F4, int Q>
int launch_scan(const KnnArgs& A, int n_qblocks, void* stream) {
    const size_t smem = scan_smem(F4, Q, A.kc);
    MMMP_CUDA_TRY(cudaFuncSetAttribute(knn_scan_kernel<F4, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_qblocks, A.n_splits, 1);
    MMMP_LAUNCH((knn_scan_kernel<F4, Q>), grid, KNN_THREADS, smem, stream, A);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

unsigned long long* g_knn_stats = nullptr;  // device counters, enabled by mmmp_knn_stats

}  // namespace

extern "C" int mmmp_knn_stats(int enable, unsigned long long* out3_h) {
    if (enable && !g_knn_stats) {
        MMMP_CUDA_TRY(cudaMalloc(&g_knn_stats, 3 * sizeof(unsigned long long)));
        MMMP_CUDA_TRY(cudaMemset(g_knn_stats, 0, 3 * sizeof(unsigned long long)));
    }
    if (out3_h && g_knn_stats) {
        MMMP_CUDA_TRY(cudaDeviceSynchronize());
        MMMP_CUDA_TRY(cudaMemcpy(out3_h, g_knn_stats, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        MMMP_CUDA_TRY(cudaMemset(g_knn_stats, 0, 3 * sizeof(unsigned long long)));
    }
    if (!enable && g_knn_stats) {
        cudaFree(g_knn_stats);
        g_knn_stats = nullptr;
    }
    return MMMP_OK;
}

extern "C" int mmmp_knn(const float* fref, const float* xref, int64_t n_ref, const float* fquery,
                        const float* xquery, int64_t n_query, int ldf, int fdim, int ldx, const mmmp_metric* metric,
                        int k, int exclude_self, int causal, int64_t query_id0, int32_t* out_idx, float* out_dist,
                        void* stream) {
    if (!fref || !xref || !fquery || !xquery || !out_idx || n_ref < 0 || n_query < 0 || k < 1 || k > MMMP_MAX_K)
        return MMMP_ERR_ARG;
    if (n_ref >= (1ll << 28)) return MMMP_ERR_LIMIT;
    if (fdim < 1 || fdim > ldf || fdim > 31) return MMMP_ERR_ARG;
    KnnArgs A;
    int rc = fill_metric(metric, ldx, A.M);
    if (rc) return rc;
    const int W = width_for(fdim);
    if (!W) return MMMP_ERR_LIMIT;
    const int F4 = W / 4;
    if (n_query == 0) return MMMP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 148, max_smem = 0;
    MMMP_CUDA_TRY(cudaGetDevice(&dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // queries per thread: as many as keep every SM busy with >= 2 CTAs of work
    // (measured on B200: Q = 2 beats 4 and 8 for every row width -- more resident warps win)
    const int qmax = F4 <= 1 ? 8 : (F4 <= 2 ? 4 : (F4 <= 4 ? 2 : 1));
    int Q = 1;
    if (qmax >= 2 && n_query >= (int64_t)sms * KNN_CTHREADS * 2 * 2) Q = 2;
    if (const char* e = getenv("MMMP_KNN_Q")) {
        const int q = atoi(e);
        if ((q == 1 || q == 2 || q == 4 || q == 8) && q <= qmax) Q = q;
    }
    while (Q > 1 && scan_smem(F4, Q, k) > (size_t)max_smem) Q >>= 1;
    if (scan_smem(F4, Q, k) > (size_t)max_smem) return MMMP_ERR_LIMIT;
    const int n_qblocks = mmmp_div_up(n_query, (int64_t)KNN_CTHREADS * Q);
    const int64_t total_tiles = (n_ref + KNN_TILE - 1) / KNN_TILE;
    int n_splits = 1;
    if (n_qblocks < 2 * sms) {
        n_splits = (2 * sms + n_qblocks - 1) / n_qblocks;
        if (n_splits > 32) n_splits = 32;
        if (n_splits > total_tiles) n_splits = (int)(total_tiles > 0 ? total_tiles : 1);
    }
    // packed filter rows (library-owned, stream ordered)
    float *pref = nullptr, *pquery = nullptr;
    MMMP_CUDA_TRY(cudaMallocAsync(&pref, (size_t)(n_ref > 0 ? n_ref : 1) * W * sizeof(float), s));
    if (n_ref > 0) {
        int g = mmmp_div_up(n_ref, 256);
        if (g > sms * 16) g = sms * 16;
        MMMP_LAUNCH(knn_pack_kernel, g, 256, 0, stream, fref, ldf, fdim, n_ref, pref, W);
    }
    const bool same = (fquery == fref && n_query == n_ref);
    if (same) {
        pquery = pref;
    } else {
        MMMP_CUDA_TRY(cudaMallocAsync(&pquery, (size_t)n_query * W * sizeof(float), s));
        int g = mmmp_div_up(n_query, 256);
        if (g > sms * 16) g = sms * 16;
        MMMP_LAUNCH(knn_pack_kernel, g, 256, 0, stream, fquery, ldf, fdim, n_query, pquery, W);
    }
    A.pref = pref;
    A.pquery = pquery;
    A.xref = xref;
    A.xquery = xquery;
    A.n_ref = n_ref;
    A.n_query = n_query;
    A.query_id0 = query_id0;
    A.ldx = ldx;
    A.kc = k;
    A.flags = (exclude_self ? 1 : 0) | (causal ? 2 : 0);
    A.n_splits = n_splits;
    A.tiles_per_split = (int)((total_tiles + n_splits - 1) / n_splits);
    if (A.tiles_per_split < 1) A.tiles_per_split = 1;
    A.stats = g_knn_stats;
    int32_t* tmp_idx = nullptr;
    float* tmp_val = nullptr;
    const size_t n_out = (size_t)n_query * n_splits * k;
    MMMP_CUDA_TRY(cudaMallocAsync(&tmp_idx, n_out * sizeof(int32_t), s));
    MMMP_CUDA_TRY(cudaMallocAsync(&tmp_val, n_out * sizeof(float), s));
    A.out_idx = tmp_idx;
    A.out_val = tmp_val;
    switch (F4 * 10 + Q) {
        case 11: rc = launch_scan<1, 1>(A, n_qblocks, stream); break;
        case 12: rc = launch_scan<1, 2>(A, n_qblocks, stream); break;
        case 14: rc = launch_scan<1, 4>(A, n_qblocks, stream); break;
        case 18: rc = launch_scan<1, 8>(A, n_qblocks, stream); break;
        case 21: rc = launch_scan<2, 1>(A, n_qblocks, stream); break;
        case 22: rc = launch_scan<2, 2>(A, n_qblocks, stream); break;
        case 24: rc = launch_scan<2, 4>(A, n_qblocks, stream); break;
        case 41: rc = launch_scan<4, 1>(A, n_qblocks, stream); break;
        case 42: rc = launch_scan<4, 2>(A, n_qblocks, stream); break;
        case 81: rc = launch_scan<8, 1>(A, n_qblocks, stream); break;
        default: rc = MMMP_ERR_LIMIT;
    }
    if (rc == MMMP_OK) {
        int g = mmmp_div_up(n_query, 128);
        if (g > sms * 8) g = sms * 8;
        MMMP_LAUNCH(knn_merge_kernel, g, 128, 0, stream, tmp_idx, tmp_val, n_query, n_splits, k, A.M.kind, out_idx,
                    out_dist);
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) rc = MMMP_ERR_CUDA_BASE + (int)e;
    }
    cudaFreeAsync(tmp_idx, s);
    cudaFreeAsync(tmp_val, s);
    if (!same) cudaFreeAsync(pquery, s);
    cudaFreeAsync(pref, s);
    return rc;
}
