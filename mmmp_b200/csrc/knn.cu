// Kernel (d), second half: fp64 re-rank of the scan's candidates and the radius search
// (the fp32 scan itself lives in knn_scan.cu).
//
// Re-rank kernel (fp64): the kc >= k candidates are re-evaluated in fp64 in the
// reference's arithmetic order and sorted by (distance, id) so that the result
// equals the reference's float64 ranking wherever fp32 could separate ranks.
// Radius kernel: fp32 lower-bound filter per (query, reference) pair, exact fp64 metric on
// the survivors, two passes (count, then fill a CSR table).
//
// Reference call sites replaced: heap loops decoupled_prm.py:480-504,527-550,
// 553-571,943-995; KDTree.query / query_ball_point coupled/degree_coupled_prm.py:67,82,
// coupled/distance_coupled_prm.py:133; planar single_arm/degree_prm.py:45-70,178,
// single_arm/distance_prm.py:40-52,126; metrics robots/panda.py:248-254,
// robots/planar_arm.py:54-61, decoupled_prm.py:824.
#include <math_constants.h>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 128;          // reference rows per shared-memory tile
constexpr float KNN_SLACK = 2.0e-5f;   // relative slack of the fp32 filter

struct MetricDev {
    int kind, dim, n_groups;
    float w[MMMP_MAX_GROUPS];
};

// per-column weight of the radius filter's quadratic form
__device__ __forceinline__ float filter_w2(const MetricDev& M, int col) {
    if (M.kind != MMMP_METRIC_GROUP_L2_SUM) return col < M.dim ? 1.f : 0.f;
    const int g = col / 3;
    return g < M.n_groups ? M.w[g] * M.w[g] : 0.f;
}

// ---- fp64 re-rank ------------------------------------------------------------
// distance in the reference's arithmetic:
//  L2            : np.linalg.norm(q1 - q2)          -> sqrt(sum_i d_i^2), i ascending
//  SQ_SUM        : sum_l np.linalg.norm(p1_l-p2_l)**2, l ascending (pairs)
//  GROUP_L2_SUM  : sum_f np.linalg.norm(p1_f-p2_f), f ascending over ALL frames
//                  (rows are fp64 frame positions, metric.dim = 3 * n_frames)
__device__ __forceinline__ double exact_metric_f64(int kind, int dim, int group, const double* a, const double* b) {
    if (kind == MMMP_METRIC_L2) {
        double s = 0.0;
        for (int i = 0; i < dim; ++i) {
            const double d = __dsub_rn(a[i], b[i]);
            s = __dadd_rn(s, __dmul_rn(d, d));
        }
        return sqrt(s);
    }
    double s = 0.0;
    for (int g = 0; g * group < dim; ++g) {
        double n2 = 0.0;
        for (int i = 0; i < group; ++i) {
            const double d = __dsub_rn(a[g * group + i], b[g * group + i]);
            n2 = __dadd_rn(n2, __dmul_rn(d, d));
        }
        const double n = sqrt(n2);
        s = __dadd_rn(s, kind == MMMP_METRIC_SQ_SUM ? __dmul_rn(n, n) : n);
    }
    return s;
}

// one query row, one thread: candidates evaluated and insertion-sorted in turn
__device__ void refine_serial(const double* __restrict__ ref, const double* __restrict__ query, int ld, int kind, int dim,
                              int group, const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_dist,
                              int64_t q, int kc, int k, double tie_eps, int32_t* __restrict__ out_idx,
                              double* __restrict__ out_dist, uint8_t* __restrict__ ambiguous) {
    double dv[MMMP_MAX_K];
    int di[MMMP_MAX_K];
    int n = 0;
    const double* a = query + q * ld;
    double err32 = 0.0;   // largest |fp32 scan value - float64 value| seen on this row's candidates
    for (int c = 0; c < kc; ++c) {
        const int id = cand_idx[(size_t)q * kc + c];
        if (id < 0) continue;
        const double d = exact_metric_f64(kind, dim, group, a, ref + (int64_t)id * ld);
        if (cand_dist) err32 = fmax(err32, fabs(d - (double)cand_dist[(size_t)q * kc + c]));
        int p = n++;
        while (p > 0 && (dv[p - 1] > d || (dv[p - 1] == d && di[p - 1] > id))) {
            dv[p] = dv[p - 1];
            di[p] = di[p - 1];
            --p;
        }
        dv[p] = d;
        di[p] = id;
    }
    // exact float64 ties straddling rank k: the reference's bounded heap pops the
    // smallest id first, i.e. keeps the LARGER ids (decoupled_prm.py:490-492).
    if (n > k && dv[k] == dv[k - 1]) {
        int lo = k - 1, hi = k;
        while (lo > 0 && dv[lo - 1] == dv[k - 1]) --lo;
        while (hi + 1 < n && dv[hi + 1] == dv[k - 1]) ++hi;
        const int need = k - lo;                 // slots left for the tie group [lo, hi]
        const int first = (hi - lo + 1) - need;  // skip this many of its smallest ids
        for (int c = lo; c < k; ++c) {
            dv[c] = dv[c + first];
            di[c] = di[c + first];
        }
    }
    for (int c = 0; c < k; ++c) {
        out_idx[(size_t)q * k + c] = c < n ? di[c] : -1;
        if (out_dist) out_dist[(size_t)q * k + c] = c < n ? dv[c] : CUDART_INF;
    }
    if (ambiguous) {
        // The fp32 scan kept the kc smallest fp32 values.  A row it dropped has an fp32 value of
        // at least vl (its kc-th); that row belongs in the float64 top k only if fp32 error can
        // bridge the gap between vl and the float64 k-th distance.  The error is not assumed: it
        // is measured on this row's own candidates (err32) and given a factor 4 in hand, with
        // tie_eps as the floor (an absolute 1e-6 ignored the metric's magnitude).
        uint8_t amb = 0;
        if (kc > k && n >= kc) {
            const double vl = (double)cand_dist[(size_t)q * kc + (kc - 1)];
            if (vl - dv[k - 1] <= fmax(tie_eps, 4.0 * err32)) amb = 1;
        }
        ambiguous[q] = amb;
    }
}

__global__ void __launch_bounds__(128)
knn_refine_kernel(const double* __restrict__ ref, const double* __restrict__ query, int ld, int kind, int dim, int group,
                  const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_dist, int64_t n_query, int kc,
                  int k, double tie_eps, int32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                  uint8_t* __restrict__ ambiguous) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_query; q += (int64_t)gridDim.x * blockDim.x)
        refine_serial(ref, query, ld, kind, dim, group, cand_idx, cand_dist, q, kc, k, tie_eps, out_idx, out_dist,
                      ambiguous);
}

// The same answer with G lanes per query row (kc <= G): lane c evaluates candidate c -- the kc gathers
// of a row are in flight together instead of one after the other --, ranks it among the row's
// candidates by (distance, id) with shuffles, and the lanes of rank < k write the row.  An exact
// float64 tie across rank k (the reference's heap rule above) sends the row to refine_serial.
template <int G>
__global__ void __launch_bounds__(128)
knn_refine_group_kernel(const double* __restrict__ ref, const double* __restrict__ query, int ld, int kind, int dim,
                        int group, const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_dist,
                        int64_t n_query, int kc, int k, double tie_eps, int32_t* __restrict__ out_idx,
                        double* __restrict__ out_dist, uint8_t* __restrict__ ambiguous) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane / G * G));
    const int gpc = blockDim.x / G;
    for (int64_t q = (int64_t)blockIdx.x * gpc + threadIdx.x / G; q < n_query; q += (int64_t)gridDim.x * gpc) {
        const int id = sub < kc ? cand_idx[(size_t)q * kc + sub] : -1;
        const bool valid = id >= 0;
        double d = CUDART_INF;
        if (valid) d = exact_metric_f64(kind, dim, group, query + q * ld, ref + (int64_t)id * ld);
        double err32 = (valid && cand_dist) ? fabs(d - (double)cand_dist[(size_t)q * kc + sub]) : 0.0;
        int rank = 0;
        for (int o = 0; o < G; ++o) {
            const double od = __shfl_sync(gmask, d, o, G);
            const int oi = __shfl_sync(gmask, id, o, G);
            rank += (oi >= 0 && (od < d || (od == d && oi < id))) ? 1 : 0;
        }
        const unsigned vmask = __ballot_sync(gmask, valid) & gmask;
        const int n = __popc(vmask);
        // the values at ranks k-1 and k (ranks are distinct among valid lanes: ids are)
        const unsigned at_k1 = __ballot_sync(gmask, valid && rank == k - 1) & gmask;
        const unsigned at_k = __ballot_sync(gmask, valid && rank == k) & gmask;
        const double dk1 = __shfl_sync(gmask, d, at_k1 ? (__ffs(at_k1) - 1) % G : 0, G);
        const double dk = __shfl_sync(gmask, d, at_k ? (__ffs(at_k) - 1) % G : 0, G);
        for (int o = G / 2; o > 0; o >>= 1) err32 = fmax(err32, __shfl_xor_sync(gmask, err32, o, G));
        if (at_k && at_k1 && dk == dk1) {   // tie across the cut: the sequential rule decides
            if (sub == 0)
                refine_serial(ref, query, ld, kind, dim, group, cand_idx, cand_dist, q, kc, k, tie_eps, out_idx,
                              out_dist, ambiguous);
            continue;
        }
        if (valid && rank < k) {
            out_idx[(size_t)q * k + rank] = id;
            if (out_dist) out_dist[(size_t)q * k + rank] = d;
        }
        if (sub >= n && sub < k) {
            out_idx[(size_t)q * k + sub] = -1;
            if (out_dist) out_dist[(size_t)q * k + sub] = CUDART_INF;
        }
        if (ambiguous && sub == 0) {
            uint8_t amb = 0;
            if (kc > k && n >= kc) {
                const double vl = (double)cand_dist[(size_t)q * kc + (kc - 1)];
                if (vl - dk1 <= fmax(tie_eps, 4.0 * err32)) amb = 1;
            }
            ambiguous[q] = amb;
        }
    }
}

// ---- radius search -------------------------------------------------------------
// thread per query, references streamed through shared memory; exact fp64 test
// for rows that survive the fp32 filter.
struct RadiusArgs {
    const float* ref;
    const float* query;
    const double* ref64;
    const double* query64;
    int64_t n_ref, n_query, query_id0;
    int ldf, ld64, dim64, group, strict, exclude_self, exclude_zero, causal;
    double r;
    int32_t* counts;
    const int64_t* offsets;
    int32_t* out_idx;
    double* out_dist;
    MetricDev M;
};

template <int F4>
__global__ void __launch_bounds__(KNN_THREADS)
radius_kernel(RadiusArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ROWF = F4 * 4;
    const int ldf = A.ldf;
    float* tile = reinterpret_cast<float*>(smem_raw);
    float* sB = tile + (size_t)KNN_TILE * ldf;
    const int tid = threadIdx.x;
    const MetricDev& M = A.M;
    const int64_t qi = (int64_t)blockIdx.x * KNN_THREADS + tid;
    const bool valid = qi < A.n_query;
    const int64_t qid = A.query_id0 + qi;
    float a[ROWF], a2[ROWF];
    float Aq = 0.f;
#pragma unroll
    for (int i = 0; i < ROWF; ++i) {
        const float v = (valid && i < ldf) ? A.query[qi * ldf + i] : 0.f;
        const float w2 = filter_w2(M, i);
        a[i] = v;
        a2[i] = -2.f * w2 * v;
        Aq = fmaf(w2 * v, v, Aq);
    }
    // filter threshold in the quadratic-form domain; SQ_SUM radii are already squared sums
    const float rf = (float)A.r;
    float tf = (M.kind == MMMP_METRIC_SQ_SUM) ? rf : rf * rf;
    tf = tf * (1.f + 8.f * KNN_SLACK) + 1e-12f;
    const float thr = valid ? tf - Aq * (1.f - KNN_SLACK) : -CUDART_INF_F;
    int64_t ref_end = A.n_ref;
    if (A.causal) {
        const int64_t last_q = A.query_id0 + (int64_t)(blockIdx.x + 1) * KNN_THREADS;
        if (last_q < ref_end) ref_end = last_q;
    }
    int count = 0;
    const int64_t obase = (A.out_idx && valid) ? A.offsets[qi] : 0;
    for (int64_t r0 = 0; r0 < ref_end; r0 += KNN_TILE) {
        const int cnt = (int)((ref_end - r0) < KNN_TILE ? (ref_end - r0) : KNN_TILE);
        __syncthreads();
        for (int i = tid; i < cnt * ldf / 4; i += KNN_THREADS)
            reinterpret_cast<float4*>(tile)[i] = reinterpret_cast<const float4*>(A.ref + r0 * ldf)[i];
        __syncthreads();
        for (int j = tid; j < cnt; j += KNN_THREADS) {
            const float* row = tile + (size_t)j * ldf;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < ROWF; ++i) acc = fmaf(filter_w2(M, i) * row[i], row[i], acc);
            sB[j] = acc * (1.f - KNN_SLACK);
        }
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            const float4* row = reinterpret_cast<const float4*>(tile + (size_t)j * ldf);
            float sacc = sB[j];
#pragma unroll
            for (int i = 0; i < F4; ++i) {
                const float4 r = row[i];
                sacc = fmaf(a2[4 * i + 0], r.x, sacc);
                sacc = fmaf(a2[4 * i + 1], r.y, sacc);
                sacc = fmaf(a2[4 * i + 2], r.z, sacc);
                sacc = fmaf(a2[4 * i + 3], r.w, sacc);
            }
            if (sacc <= thr) {
                const int64_t id = r0 + j;
                if (A.exclude_self && id == qid) continue;
                if (A.causal && id >= qid) continue;
                // the fp32 metric on the rows at hand settles every pair that is not within 1e-4 of
                // the radius; only that band (and the pairs to be reported, which need their
                // float64 distance) pays for the divergent float64 evaluation -- inline fp64 on
                // every filter survivor was 30 ms for 10 k x 10 k FK-metric rows
                {
                    float v32 = 0.f;
                    const float* rowf = tile + (size_t)j * ldf;
                    if (M.kind == MMMP_METRIC_GROUP_L2_SUM) {
                        for (int g = 0; g < M.n_groups; ++g) {
                            const float dx = a[3 * g] - rowf[3 * g], dy = a[3 * g + 1] - rowf[3 * g + 1],
                                        dz = a[3 * g + 2] - rowf[3 * g + 2];
                            v32 = fmaf(M.w[g], sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz))), v32);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < ROWF; ++i) {
                            const float dd = i < M.dim ? a[i] - rowf[i] : 0.f;
                            v32 = fmaf(dd, dd, v32);
                        }
                        if (M.kind == MMMP_METRIC_L2) v32 = sqrtf(v32);
                    }
                    if (v32 > rf * (1.f + 1e-4f) + 1e-4f) continue;   // certainly outside
                }
                const double d = exact_metric_f64(M.kind, A.dim64, A.group, A.query64 + qi * A.ld64,
                                                  A.ref64 + id * A.ld64);
                const bool in = A.strict ? (d < A.r) : (d <= A.r);
                if (!in) continue;
                if (A.exclude_zero && !(d > 0.0)) continue;
                if (A.out_idx) {
                    A.out_idx[obase + count] = (int32_t)id;
                    if (A.out_dist) A.out_dist[obase + count] = d;
                }
                ++count;
            }
        }
    }
    if (valid && A.counts && !A.out_idx) A.counts[qi] = count;
}

// feature rows are exactly 4, 8, 16 or 32 floats wide (zero padded)
int f4_for(int ldf) { return ldf == 4 ? 1 : (ldf == 8 ? 2 : (ldf == 16 ? 4 : (ldf == 32 ? 8 : 0))); }

int fill_metric(const mmmp_metric* m, int ldf, MetricDev& M) {
    if (!m) return MMMP_ERR_ARG;
    M.kind = m->kind;
    M.dim = m->dim;
    M.n_groups = m->n_groups;
    if (m->kind < MMMP_METRIC_L2 || m->kind > MMMP_METRIC_SQ_SUM) return MMMP_ERR_KIND;
    if (m->dim < 1 || m->dim > 32 || m->dim > ldf) return MMMP_ERR_ARG;
    for (int g = 0; g < MMMP_MAX_GROUPS; ++g) M.w[g] = 0.f;
    if (m->kind == MMMP_METRIC_GROUP_L2_SUM) {
        if (m->n_groups < 1 || m->n_groups > MMMP_MAX_GROUPS || m->n_groups * 3 != m->dim) return MMMP_ERR_ARG;
        for (int g = 0; g < m->n_groups; ++g) M.w[g] = m->weight[g];
    }
    return MMMP_OK;
}


}  // namespace


extern "C" int mmmp_knn_refine(const double* ref64, const double* query64, int ld64, const mmmp_metric* metric,
                               const int32_t* cand_idx, const float* cand_dist, int64_t n_query, int kc, int k,
                               double tie_eps, int32_t* out_idx, double* out_dist, uint8_t* ambiguous,
                               void* stream) {
    if (!ref64 || !query64 || !metric || !cand_idx || !out_idx || n_query < 0) return MMMP_ERR_ARG;
    if (k < 1 || kc < k || kc > MMMP_MAX_K || (ambiguous && !cand_dist)) return MMMP_ERR_ARG;
    if (metric->kind < MMMP_METRIC_L2 || metric->kind > MMMP_METRIC_SQ_SUM) return MMMP_ERR_KIND;
    if (metric->dim < 1 || metric->dim > ld64) return MMMP_ERR_ARG;
    if (n_query == 0) return MMMP_OK;
    const int group = metric->kind == MMMP_METRIC_L2 ? 1 : (metric->n_groups > 0 ? metric->dim / metric->n_groups : 1);
    if (kc <= 32 && !getenv("MMMP_REFINE_SERIAL")) {
        const int G = kc <= 16 ? 16 : 32;
        int g = mmmp_div_up(n_query, 128 / G);
        if (g > 148 * 32) g = 148 * 32;
        if (G == 16)
            MMMP_LAUNCH(knn_refine_group_kernel<16>, g, 128, 0, stream, ref64, query64, ld64, metric->kind, metric->dim,
                        group, cand_idx, cand_dist, n_query, kc, k, tie_eps, out_idx, out_dist, ambiguous);
        else
            MMMP_LAUNCH(knn_refine_group_kernel<32>, g, 128, 0, stream, ref64, query64, ld64, metric->kind, metric->dim,
                        group, cand_idx, cand_dist, n_query, kc, k, tie_eps, out_idx, out_dist, ambiguous);
    } else {
        int g = mmmp_div_up(n_query, 128);
        if (g > 148 * 8) g = 148 * 8;
        MMMP_LAUNCH(knn_refine_kernel, g, 128, 0, stream, ref64, query64, ld64, metric->kind, metric->dim, group,
                    cand_idx, cand_dist, n_query, kc, k, tie_eps, out_idx, out_dist, ambiguous);
    }
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_radius(const float* ref, const double* ref64, int64_t n_ref, const float* query,
                           const double* query64, int64_t n_query, int ldf, int ld64,
                           const mmmp_metric* metric, const mmmp_metric* metric64, double r, int strict,
                           int exclude_self, int exclude_zero, int causal, int64_t query_id0, int32_t* counts,
                           const int64_t* offsets, int32_t* out_idx, double* out_dist, void* stream) {
    if (!ref || !ref64 || !query || !query64 || !metric64 || n_ref < 0 || n_query < 0) return MMMP_ERR_ARG;
    if (!counts && !out_idx) return MMMP_ERR_ARG;
    if (out_idx && !offsets) return MMMP_ERR_ARG;
    if (ldf < 4 || (ldf & 3) || ldf > 32) return MMMP_ERR_ARG;
    RadiusArgs A;
    int rc = fill_metric(metric, ldf, A.M);
    if (rc) return rc;
    if (metric64->kind != metric->kind || metric64->dim < 1 || metric64->dim > ld64) return MMMP_ERR_ARG;
    if (n_query == 0) return MMMP_OK;
    const int F4 = f4_for(ldf);
    if (!F4) return MMMP_ERR_LIMIT;
    A.ref = ref;
    A.query = query;
    A.ref64 = ref64;
    A.query64 = query64;
    A.n_ref = n_ref;
    A.n_query = n_query;
    A.query_id0 = query_id0;
    A.ldf = ldf;
    A.ld64 = ld64;
    A.dim64 = metric64->dim;
    A.group = metric64->kind == MMMP_METRIC_L2 ? 1 : (metric64->n_groups > 0 ? metric64->dim / metric64->n_groups : 1);
    A.strict = strict;
    A.exclude_self = exclude_self;
    A.exclude_zero = exclude_zero;
    A.causal = causal;
    A.r = r;
    A.counts = counts;
    A.offsets = offsets;
    A.out_idx = out_idx;
    A.out_dist = out_dist;
    const size_t smem = (size_t)KNN_TILE * ldf * 4 + KNN_TILE * 4;
    const int grid = mmmp_div_up(n_query, KNN_THREADS);
    if (F4 == 1) MMMP_LAUNCH((radius_kernel<1>), grid, KNN_THREADS, smem, stream, A);
    else if (F4 == 2) MMMP_LAUNCH((radius_kernel<2>), grid, KNN_THREADS, smem, stream, A);
    else if (F4 == 4) MMMP_LAUNCH((radius_kernel<4>), grid, KNN_THREADS, smem, stream, A);
    else MMMP_LAUNCH((radius_kernel<8>), grid, KNN_THREADS, smem, stream, A);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

// ---- packed neighbour records: the post-search exchange of a sharded search ------------------
// A rank that answered m query rows of a sharded search (mmmp_knn_range + mmmp_knn_refine) packs
// them as records of 2 + 3k 32-bit words -- [row id, ambiguous, idx[k], dist[k] (fp64 as two
// words)] -- so that ONE all-gather ships everything, and every rank scatters the gathered
// records into tables in the caller's row order with one kernel.  Records beyond m (padding up
// to the fixed block size of the collective) carry row id -1 and are skipped by the scatter.
namespace {

__global__ void __launch_bounds__(256)
pack_records_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ idx, const double* __restrict__ dist,
                    const uint8_t* __restrict__ amb, int64_t m, int k, int64_t cap, uint32_t* __restrict__ rec) {
    const int rw = 2 + 3 * k;
    const int64_t total = cap * rw;
    const uint32_t* dist_w = reinterpret_cast<const uint32_t*>(dist);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / rw;
        const int c = (int)(t - r * rw);
        uint32_t v = 0;
        if (r >= m) v = c == 0 ? 0xffffffffu : 0u;
        else if (c == 0) v = (uint32_t)rows[r];
        else if (c == 1) v = amb ? amb[r] : 0u;
        else if (c < 2 + k) v = (uint32_t)idx[r * k + (c - 2)];
        else v = dist_w[r * 2 * k + (c - 2 - k)];
        rec[t] = v;
    }
}

__global__ void __launch_bounds__(256)
scatter_records_kernel(const uint32_t* __restrict__ rec, int64_t total_rec, int k, int64_t n, int32_t* __restrict__ idx,
                       double* __restrict__ dist, uint8_t* __restrict__ amb) {
    const int rw = 2 + 3 * k;
    const int64_t total = total_rec * rw;
    uint32_t* dist_w = reinterpret_cast<uint32_t*>(dist);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / rw;
        const int c = (int)(t - r * rw);
        const int64_t row = (int32_t)rec[r * rw];
        if (row < 0 || row >= n || c == 0) continue;
        const uint32_t v = rec[t];
        if (c == 1) {
            if (amb) amb[row] = (uint8_t)v;
        } else if (c < 2 + k) {
            idx[row * k + (c - 2)] = (int32_t)v;
        } else if (dist) {
            dist_w[row * 2 * k + (c - 2 - k)] = v;
        }
    }
}

}  // namespace

extern "C" int mmmp_pack_neighbour_records(const int32_t* rows_d, const int32_t* idx_d, const double* dist_d,
                                           const uint8_t* amb_d, int64_t m, int k, int64_t cap, void* rec_d,
                                           void* stream) {
    if (!rows_d || !idx_d || !dist_d || !rec_d || m < 0 || cap < m || k < 1 || k > MMMP_MAX_K) return MMMP_ERR_ARG;
    if (cap == 0) return MMMP_OK;
    int g = mmmp_div_up(cap * (2 + 3 * k), 256);
    if (g > 148 * 16) g = 148 * 16;
    MMMP_LAUNCH(pack_records_kernel, g, 256, 0, stream, rows_d, idx_d, dist_d, amb_d, m, k, cap, (uint32_t*)rec_d);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_scatter_neighbour_records(const void* rec_d, int64_t n_records, int k, int64_t n, int32_t* idx_d,
                                              double* dist_d, uint8_t* amb_d, void* stream) {
    if (!rec_d || !idx_d || n_records < 0 || n < 0 || k < 1 || k > MMMP_MAX_K) return MMMP_ERR_ARG;
    if (n_records == 0) return MMMP_OK;
    int g = mmmp_div_up(n_records * (2 + 3 * k), 256);
    if (g > 148 * 16) g = 148 * 16;
    MMMP_LAUNCH(scatter_records_kernel, g, 256, 0, stream, (const uint32_t*)rec_d, n_records, k, n, idx_d, dist_d, amb_d);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}
