// Kernel (d), scan stage: exact k nearest candidates in fp32 on the CUDA cores.
//
// The hot loop evaluates a cheap LOWER BOUND of the metric as a squared L2 distance
// between "filter rows" in dot-product form
//        s = B_r - 2 <a, r>   <=   tau_f - A_q          (B_r, A_q = squared row norms)
// (dim FFMAs + a share of one FMNMX tree / compare per pair):
//   L2 / SQ_SUM metrics : filter row = the row itself (the bound is the metric);
//   FK metric sum_g w_g |dp_g| (robots/panda.py:248-254): filter row = the weighted
//       centroid sum_g w_g p_g (3 floats) -- triangle inequality
//       sum_g w_g |dp_g| >= | sum_g w_g dp_g |.
// A pack kernel lays the filter rows out as [x_0 .. x_{d-1}, 0 .., B_r] (B_r in the last of
// W = 4/8/16/32 columns) and the exact rows as [x_0 .. x_{dim-1}] with an ODD row stride S,
// both in the scan order, so that a tile of either is one contiguous block.
//
// Spatial ordering + exact tile culling (non-causal searches with >= 4096 references):
// rows are counting-sorted by the position of their grid cell on the 3-D Hilbert curve (first
// three filter dimensions), so that a shared-memory tile of 128 consecutive rows and a CTA's 128*Q
// consecutive queries are both spatially compact; every tile carries its bounding box.
// The producer warp walks the tiles outward from the CTA's own position, 32 boxes per
// step, and only loads a tile whose box is closer to the CTA's query box than the CTA's
// current worst k-th best bound -- a tile that is skipped cannot contain a neighbour, so
// the result is identical to the exhaustive scan (which remains the path for causal /
// small searches).
//
// CTA = 4 consumer warps + 1 producer warp (warp specialisation).  The producer streams
// the filter rows AND the exact rows of a tile L2/HBM -> shared memory with the TMA engine
// (cp.async.bulk, 2-stage ring, full/empty mbarriers); consumer warps never meet at a CTA
// barrier.  Every consumer thread owns Q queries (filter form in registers, exact row in a
// shared-memory column) and reads the tile's filter rows as shared-memory broadcasts
// (LDS.128), 8 rows per step, software pipelined.  Pairs that pass the filter are appended
// to one compact queue per warp (ballot + popcount; 16-bit entries: owner, tile row); the
// warp drains it when it is nearly full and at the end of every tile, 32 pairs per round:
// exact fp32 metric against the tile's exact rows -- still resident in shared memory (odd
// stride: conflict free), so the drain never waits on a gather -- and the lane that
// evaluated a value that beats the owner's k-th best inserts it into the owner's sorted
// candidate column (shared memory); the owners read their tightened thresholds back after
// the drain.  Measured on B200 (1M x 1M, FK metric): 2 stages / 8 rows / Q = 1 is the
// fastest point of {2,3} x {4,8,16} x {1,2}; 72 registers and 45 KB of shared memory keep
// 5 CTAs on an SM (more resident CTAs beat deeper rings).  The per-thread queues this
// replaced (prefix sum, owner search, per-thread walk over parked values) cost 39.2 ms
// against 30.4 ms.
//
// Reference call sites replaced: the heap loops of
// planners/prm/pybullet/decoupled/decoupled_prm.py:484-494,531-540,947-954 and
// KDTree.query at coupled/degree_coupled_prm.py:67, planar single_arm/degree_prm.py:178.
#include <math_constants.h>
#include <cstdlib>
#include "common.cuh"

namespace {

#ifndef MMMP_KNN_CWARPS
#define MMMP_KNN_CWARPS 4
#endif
constexpr int KNN_CWARPS = MMMP_KNN_CWARPS;    // consumer warps per CTA
constexpr int KNN_CTHREADS = KNN_CWARPS * 32;  // consumer threads
constexpr int KNN_THREADS = KNN_CTHREADS + 32; // + producer warp
#ifndef MMMP_KNN_STAGES
#define MMMP_KNN_STAGES 2
#endif
constexpr int KNN_STAGES = MMMP_KNN_STAGES;
#ifndef MMMP_KNN_TILE
#define MMMP_KNN_TILE 128
#endif
constexpr int KNN_TILE = MMMP_KNN_TILE;        // reference rows per shared-memory tile
#ifndef MMMP_KNN_RU
#define MMMP_KNN_RU 8
#endif
// rows per inner-loop step.  Round 1 double-buffered the rows of a step in registers (software
// pipeline) with 8 rows for every width: 185 registers for 16-column rows, 636 B of spills for 32
// columns.  Measured in round 2 (profiles/r02_scan_sweep.txt): without the register double buffer --
// the resident warps hide the shared-memory latency -- and with MORE rows per step (the per-step
// vote / branch / queue check is paid less often) every width is faster: FK scan 30.7 -> 24.2 ms,
// 7-D L2 58.1 -> 50.3 ms, 14-D 43.2 -> 39.5 ms (1 M / 200 k rows, 12 candidates).
#ifndef MMMP_KNN_RU2
#define MMMP_KNN_RU2 32
#endif
#ifndef MMMP_KNN_RU4
#define MMMP_KNN_RU4 16
#endif
#ifndef MMMP_KNN_PREFETCH
#define MMMP_KNN_PREFETCH 0                    // 1: register double buffer of the rows (the round-1 loop)
#endif
__host__ __device__ constexpr int knn_ru(int F4) {
    return F4 == 1 ? MMMP_KNN_RU : (F4 == 2 ? MMMP_KNN_RU2 : (F4 == 4 ? MMMP_KNN_RU4 : 1));
}
#ifndef MMMP_KNN_QSLACK
#define MMMP_KNN_QSLACK 12
#endif
constexpr int KNN_QSLACK = MMMP_KNN_QSLACK;    // queue entries a thread may hold before its warp drains
__host__ __device__ constexpr int knn_qc(int F4) { return knn_ru(F4) + KNN_QSLACK; }  // queue capacity per (thread, query)
#ifndef MMMP_KNN_MINB
#define MMMP_KNN_MINB 5                        // resident CTAs per SM the register budget must allow, 4-column filter
                                               // rows (71 registers); 8 columns: 3 CTAs, 113 registers (5 would spill 348 B)
#endif
#ifndef MMMP_KNN_MINB2
#define MMMP_KNN_MINB2 4                       // the same for 8-column filter rows (joint-space L2 of a 7-DOF arm, dual FK rows)
#endif
#ifndef MMMP_KNN_MINB4
#define MMMP_KNN_MINB4 3                       // 16-column rows (14-D composite configurations)
#endif
#ifndef MMMP_KNN_MINB8
#define MMMP_KNN_MINB8 2                       // 32-column rows (28-D composite configurations)
#endif
#ifndef MMMP_KNN_WALK
#define MMMP_KNN_WALK 2                         // producer: blocks of 16 tile boxes per side and walk step (1/2/4/8: 27.5/27.3/27.3/28.4 ms)
#endif
constexpr int KNN_WALK = MMMP_KNN_WALK;
constexpr float KNN_SLACK = 2.0e-5f;           // relative slack of the fp32 filter
constexpr int KNN_CULL_MIN = 4096;             // references below which sorting does not pay

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
#ifndef MMMP_KNN_SPIN_NS
#define MMMP_KNN_SPIN_NS 0   // measured: 256 ns of sleep per poll costs 4 % (late pushes), the plain loop stays
#endif
// the producer's wait for a free ring slot.  The try_wait loop above re-issues every ~26 cycles
// and accounts for 38 % of the kernel's executed instructions (profiles/r01_knn_scan_full.csv),
// but the issue slots it takes are idle ones: polling with a sleep (MMMP_KNN_SPIN_NS > 0) frees
// them and still loses, because every push then starts late.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
#if MMMP_KNN_SPIN_NS > 0
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(MMMP_KNN_SPIN_NS);
    }
#else
    mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ float rsqrt_ftz(float x) {  // MUFU.RSQ without the denormal fix-up
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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

// ---- spatial ordering ------------------------------------------------------------------
__device__ __forceinline__ int float_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// bbox[0..2] = min, bbox[3..5] = max of the first `nd` (<= 3) columns, as ordered ints
__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ rows, int ld, int nd, int64_t n, int* __restrict__ bbox) {
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int c = 0; c < nd; ++c) {
            const float v = rows[i * ld + c];
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0 && c < nd) {
            atomicMin(bbox + c, float_ordered(lo[c]));
            atomicMax(bbox + 3 + c, float_ordered(hi[c]));
        }
    }
}

__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

#ifndef MMMP_KNN_HILBERT
#define MMMP_KNN_HILBERT 1
#endif
// Position of the grid cell (x0, x1, x2), `bits` bits per axis, on the 3-D Hilbert curve (Skilling's
// transpose algorithm, AIP Conf. Proc. 707, 2004).  Consecutive cells of the curve share a face, so a
// run of rows -- a tile, a CTA's query block -- never straddles one of the Z curve's jumps: the run's
// bounding box stays close to the cells it covers.
__host__ __device__ __forceinline__ uint32_t hilbert3(uint32_t x0, uint32_t x1, uint32_t x2, int bits) {
    uint32_t X[3] = {x0, x1, x2};
    const uint32_t M = 1u << (bits - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t;
    X[1] ^= t;
    X[2] ^= t;
    return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

// Cell of a row (bits per axis = `bits`) in the box `bbox`, numbered along the Hilbert curve (three
// sort dimensions) or the Z curve (fewer)
__device__ __forceinline__ uint32_t cell_of(const float* row, int nd, const int* bbox, int bits) {
    uint32_t key = 0, v3[3] = {0, 0, 0};
    const float cells = (float)(1 << bits);
    for (int c = 0; c < nd; ++c) {
        const float lo = ordered_float(bbox[c]), hi = ordered_float(bbox[3 + c]);
        const float ext = hi - lo;
        int v = ext > 0.f ? (int)((row[c] - lo) / ext * cells) : 0;
        v = min(max(v, 0), (1 << bits) - 1);
        v3[c] = (uint32_t)v;
        key |= spread3((uint32_t)v) << c;
    }
    if (MMMP_KNN_HILBERT && nd == 3) key = hilbert3(v3[0], v3[1], v3[2], bits);
    return key;
}

__global__ void __launch_bounds__(256)
cell_hist_kernel(const float* __restrict__ rows, int ld, int nd, int64_t n, const int* __restrict__ bbox, int bits,
                 uint32_t* __restrict__ keys, int32_t* __restrict__ hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = cell_of(rows + i * ld, nd, bbox, bits);
        keys[i] = k;
        atomicAdd(hist + k, 1);
    }
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(const uint32_t* __restrict__ keys, int64_t n, const int64_t* __restrict__ offsets,
                    int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        const int64_t pos = offsets[k] + atomicAdd(cursor + k, 1);
        perm[pos] = (int32_t)i;
    }
}

// The scatter above leaves the rows of a cell in the order their atomics landed.  One warp per
// cell puts them in ascending row order (rank by counting), so that the scan order is a pure
// function of the data: the shards of a multi-GPU search (mmmp_knn_shard) then cut the SAME
// order on every rank, and a search repeats bit for bit.
__global__ void __launch_bounds__(256)
cell_order_kernel(const int32_t* __restrict__ in, const int64_t* __restrict__ offsets, int n_cells,
                  int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n_cells;
         c += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t o0 = offsets[c];
        const int m = (int)(offsets[c + 1] - o0);
        for (int i = lane; i < m; i += 32) {
            const int32_t v = in[o0 + i];
            int rank = 0;
            for (int j = 0; j < m; ++j) rank += in[o0 + j] < v;
            out[o0 + rank] = v;
        }
    }
}

// ---- row packing ----------------------------------------------------------------------
// filter rows: out[p] = [in[perm[p]][0..d), 0 .., (1 - eps) * |.|^2]   (W columns); perm may be NULL
// dual rows (W = 8, d = 6): out[p] = [c0 d0 c1 d1 | c2 d2 (1 - eps) |c|^2 (1 - eps) |d|^2] -- the C and D
// halves interleaved, so that one packed FFMA2 (two fp32 FMAs per instruction, sm_100) advances
// both dot products of the hot loop
__global__ void __launch_bounds__(256)
knn_pack_kernel(const float* __restrict__ in, int ld_in, int d, int64_t n, const int32_t* __restrict__ perm,
                float* __restrict__ out, int W, int dual) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const float* r = in + (int64_t)(perm ? perm[p] : p) * ld_in;
        float* o = out + p * W;
        if (dual) {
            float nc = 0.f, nd = 0.f;
            for (int c = 0; c < 6; ++c) {
                const float v = r[c];
                o[c < 3 ? 2 * c : 2 * (c - 3) + 1] = v;
                if (c < 3) nc = fmaf(v, v, nc);
                else nd = fmaf(v, v, nd);
            }
            o[6] = nc * (1.f - KNN_SLACK);
            o[7] = nd * (1.f - KNN_SLACK);
            continue;
        }
        float acc = 0.f;
        for (int c = 0; c < W - 1; ++c) {
            const float v = c < d ? r[c] : 0.f;
            acc = fmaf(v, v, acc);
            o[c] = v;
        }
        o[W - 1] = acc * (1.f - KNN_SLACK);
    }
}

// exact rows in scan order with row stride S: out[p] = in[perm[p]][0..S) (zero beyond `dim`);
// rows n .. n_pad-1 (the tail a partial tile's bulk copy touches) are zeroed
__global__ void __launch_bounds__(256)
knn_pack_exact_kernel(const float* __restrict__ in, int ld_in, int dim, int64_t n, int64_t n_pad,
                      const int32_t* __restrict__ perm, float* __restrict__ out, int S) {
    const int64_t total = n_pad * S;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / S;
        const int c = (int)(e - p * S);
        float v = 0.f;
        if (p < n && c < dim) v = in[(int64_t)(perm ? perm[p] : p) * ld_in + c];
        out[e] = v;
    }
}

// per-tile bounding boxes of the packed (sorted) rows: boxes[t] = {lo0,lo1,lo2,hi0,hi1,hi2}
__global__ void __launch_bounds__(KNN_TILE)
tile_box_kernel(const float* __restrict__ packed, int W, int nd, int cstride, int64_t n, float* __restrict__ boxes) {
    const int64_t t = blockIdx.x;
    const int64_t r = t * KNN_TILE + threadIdx.x;
    float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    if (r < n)
        for (int c = 0; c < nd; ++c) lo[c] = hi[c] = packed[r * W + c * cstride];   // dual rows: C at columns 0, 2, 4
    __shared__ float s[6][KNN_TILE / 32];
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if ((threadIdx.x & 31) == 0) {
            s[c][threadIdx.x >> 5] = lo[c];
            s[3 + c][threadIdx.x >> 5] = hi[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[threadIdx.x][0];
        for (int w = 1; w < KNN_TILE / 32; ++w)
            v = threadIdx.x < 3 ? fminf(v, s[threadIdx.x][w]) : fmaxf(v, s[threadIdx.x][w]);
        if (threadIdx.x % 3 >= nd) v = 0.f;
        boxes[t * 6 + threadIdx.x] = v;
    }
}

// (Opt-in since the rows follow the Hilbert curve, MMMP_KNN_ORDER=1: the outlier blocks this was written
// for were runs that straddled a jump of the Z curve -- 3033 tiles for one block against a mean of 191;
// along the Hilbert curve the heaviest block loads 333, and launching neighbouring blocks together
// (shared tiles in L2) is worth more than the schedule: 1 M rows 19.0 ms with, 18.4 ms without; half
// the blocks on one rank 10.6 against 10.1 ms.)
// Launch order of the query blocks of a self search, heaviest first.  A block's cost is the number
// of tiles it loads, which grows with the size of its own bounding box (= the box of the tile that
// holds the same rows): a block of outliers loads ten times the average (2682 tiles against 191 at
// 1 M rows) and, scheduled late, is the tail of the launch -- the whole of it when a rank only scans
// an eighth of the blocks.  CTAs start in blockIdx order, so handing out the big boxes first is
// longest-processing-time-first scheduling.  One CTA: bucket the blocks by box size (64 buckets of the
// extent sum, log scale), counting sort, largest bucket first; order inside a bucket is arbitrary
// (it changes when a block runs, never what it computes).
__global__ void __launch_bounds__(1024)
block_order_kernel(const float* __restrict__ boxes, int qblock0, int n_qblocks, int32_t* __restrict__ order) {
    __shared__ int hist[64], start[64];
    if (threadIdx.x < 64) hist[threadIdx.x] = 0;
    __syncthreads();
    auto bucket = [&](int b) {
        const float* x = boxes + (size_t)(qblock0 + b) * 6;
        const float e = (x[3] - x[0]) + (x[4] - x[1]) + (x[5] - x[2]);
        int k = 63 - (int)(4.f * (log2f(fmaxf(e, 1e-9f)) + 8.f));   // 4 buckets per octave, larger extent = earlier bucket
        return min(max(k, 0), 63);
    };
    for (int b = threadIdx.x; b < n_qblocks; b += blockDim.x) atomicAdd(hist + bucket(b), 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < 64; ++k) {
            start[k] = acc;
            acc += hist[k];
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_qblocks; b += blockDim.x) order[atomicAdd(start + bucket(b), 1)] = qblock0 + b;
}

struct KnnArgs {
    const float* pref;     // packed filter rows of the references [n_ref, W], scan order
    const float* pquery;   // packed filter rows of the queries [n_query, W], scan order
    const float* xs;       // exact rows of the references [n_ref rounded up to 4, S], scan order
    const float* xquery;   // exact rows of the queries [n_query, ldx], original order
    const int32_t* perm;   // scan position -> original reference row (NULL: identity)
    const int32_t* qperm;  // scan position -> original query row (NULL: identity)
    const float* boxes;    // per-tile bounding boxes (NULL: no culling)
    int64_t n_ref, n_query, query_id0;
    int ldx, S, kc;
    int flags;             // 1 = exclude self, 2 = causal, 4 = queries ARE the references (same scan order)
    int tiles_per_split, n_splits;
    int qblock0;           // first query block (of LS scan positions) this launch covers
    const int32_t* block_order;  // optional: query block of CTA x (absolute), heaviest first (NULL: qblock0 + x)
    int32_t* out_rows;     // sharded launch: compact output, row r = scan position qblock0 * LS + r;
                           // out_rows[r] = original query row (NULL: output indexed by original row)
    int32_t* out_idx;      // [n_query, n_splits, kc]
    float* out_val;        // candidate-list domain (L2/SQ_SUM: squared ; GROUP: distance)
    unsigned long long* stats;  // optional: [0] exact evaluations, [1] insertions, [2] drains, [3] tiles loaded, [4] most tiles loaded by one CTA
    MetricDev M;
};

// exact fp32 metric in the candidate list's domain: a = query row in a shared-memory column
// (element stride `as`), b = reference row in the shared-memory tile (contiguous)
__device__ __forceinline__ float exact_metric(const MetricDev& M, const float* a, int as, const float* b) {
    float v = 0.f;
    if (M.kind == MMMP_METRIC_GROUP_L2_SUM) {
#pragma unroll
        for (int g = 0; g < MMMP_MAX_GROUPS; ++g) {
            if (g < M.n_groups) {
                const float dx = a[(3 * g) * as] - b[3 * g], dy = a[(3 * g + 1) * as] - b[3 * g + 1],
                            dz = a[(3 * g + 2) * as] - b[3 * g + 2];
                const float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                v = fmaf(M.w[g], s2 * rsqrt_ftz(fmaxf(s2, 1e-30f)), v);  // candidates are re-ranked in fp64
            }
        }
    } else {
#pragma unroll 4
        for (int i = 0; i < M.dim; ++i) {
            const float d = a[i * as] - b[i];
            v = fmaf(d, d, v);
        }
    }
    return v;
}

__device__ __forceinline__ float box_dist2(const float* a, const float* b) {  // {lo[3], hi[3]} each
    float d2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float g = fmaxf(fmaxf(b[c] - a[3 + c], a[c] - b[3 + c]), 0.f);
        d2 = fmaf(g, g, d2);
    }
    return d2;
}

struct ScanCtl {                  // per-CTA control block in shared memory
    float wbox[KNN_CWARPS][6];    // query box of every consumer warp
    volatile float wtau[KNN_CWARPS];  // worst (largest) filter-domain bound of every consumer warp
    volatile int tile_id[KNN_STAGES];
    float tbox[KNN_STAGES][6];    // box of the tile in every stage (written by the producer with tile_id)
};

// shared-memory layout (floats unless noted), LS = KNN_CTHREADS * Q:
//   ring[STAGES][ TILE*W filter | TILE*S exact ]  lv[kc][LS]  li[kc][LS]  xq[S][LS]
//   warp queues (u16) [CWARPS][32*Q*QC(F4)]   bars[2*STAGES] (u64)   ctl
__host__ __device__ constexpr size_t knn_stage_floats(int W, int S) { return (size_t)KNN_TILE * (W + S); }
__host__ __device__ constexpr size_t knn_align(size_t x, size_t a) { return (x + a - 1) / a * a; }

// DUAL (F4 = 2 only): the filter rows are the dual rows of mmmp_fk_features, two tests per pair --
//   |dC|^2 <= tau^2   and   |dC|^2 + |dD|^2 <= 2 tau^2
// -- thr1 / thr hold the two thresholds, tauf stays the C-domain bound the tile culling uses.
template <int F4, int Q, bool DUAL>
__global__ void __launch_bounds__(KNN_THREADS, F4 == 1 ? MMMP_KNN_MINB : (F4 == 2 ? MMMP_KNN_MINB2 :
                                               (F4 == 4 ? MMMP_KNN_MINB4 : MMMP_KNN_MINB8)))
knn_scan_kernel(const __grid_constant__ KnnArgs A) {
    static_assert(!DUAL || F4 == 2, "dual rows are 8 columns wide");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int W = F4 * 4;
    constexpr int KNN_RU = knn_ru(F4);
    constexpr int KNN_QC = knn_qc(F4);
    constexpr int LS = KNN_CTHREADS * Q;  // column stride of the per-query shared arrays
    const int S = A.S;
    const int kc = A.kc;
    const size_t stage_floats = knn_stage_floats(W, S);
    float* ring = reinterpret_cast<float*>(smem_raw);
    float* lv = ring + KNN_STAGES * stage_floats;
    int* li = reinterpret_cast<int*>(lv + (size_t)kc * LS);
    float* xq = reinterpret_cast<float*>(li + (size_t)kc * LS);
    // warp queues: [CWARPS][32 * Q * QC] entries of 16 bits, (query slot, owner lane) << 8 | tile row
    unsigned short* wqueue = reinterpret_cast<unsigned short*>(xq + (size_t)S * LS);
    unsigned char* queue = reinterpret_cast<unsigned char*>(wqueue);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + knn_align((size_t)(queue - smem_raw) +
                                                                      (size_t)Q * KNN_QC * KNN_CTHREADS * 2, 8));
    ScanCtl* ctl = reinterpret_cast<ScanCtl*>(bars + 2 * KNN_STAGES);
    const uint32_t ring_s = smem_u32(ring);
    const uint32_t full_s = smem_u32(bars), empty_s = smem_u32(bars + KNN_STAGES);

    const int tid = threadIdx.x;
    const bool cull = A.boxes != nullptr;
    if (tid == 0) {
        for (int s = 0; s < KNN_STAGES; ++s) {
            mbar_init(full_s + 8 * s, 1);
            mbar_init(empty_s + 8 * s, KNN_CWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- reference tiles of this CTA ---------------------------------------------
    const int64_t q_first = (A.block_order ? (int64_t)__ldg(A.block_order + blockIdx.x) : (int64_t)blockIdx.x + A.qblock0) * LS;
    int64_t ref_end = A.n_ref;
    if (A.flags & 2) {
        const int64_t last_q = A.query_id0 + q_first + LS;  // exclusive bound on this CTA's query ids
        if (last_q < ref_end) ref_end = last_q;
    }
    const int64_t total_tiles = (ref_end + KNN_TILE - 1) / KNN_TILE;
    const int64_t tile_lo = (int64_t)blockIdx.y * A.tiles_per_split;
    int64_t tile_hi = tile_lo + A.tiles_per_split;
    if (tile_hi > total_tiles) tile_hi = total_tiles;

    // ---- queries: filter form in registers (a2 = -2 x, last column of the packed row = A_q),
    //      exact row in this thread's shared-memory column ----
    float a2[Q][W], thr[Q], thr1[Q], Aq[Q], Aq1[Q], tauf[Q];
    float2 nthr2[Q];  // DUAL: (-thr1, -(thr - thr1)), added to (accC, accD) by one packed FADD2
    int64_t qrow[Q];
    if (tid < KNN_CTHREADS) {
        float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
        for (int qq = 0; qq < Q; ++qq) {
            const int64_t qi = q_first + qq * KNN_CTHREADS + tid;
            const bool valid = qi < A.n_query;
            const int col = qq * KNN_CTHREADS + tid;
#pragma unroll
            for (int i = 0; i < W - 1; ++i) {
                const float v = valid ? __ldg(A.pquery + qi * W + i) : 0.f;
                a2[qq][i] = -2.f * v;
                const int ci = DUAL ? (i & 1 ? 3 : i >> 1) : i;   // dual rows: C at columns 0, 2, 4
                if (ci < 3 && valid) {
                    lo[ci] = fminf(lo[ci], v);
                    hi[ci] = fmaxf(hi[ci], v);
                }
            }
            Aq[qq] = valid ? __ldg(A.pquery + qi * W + (W - 1)) : 0.f;
            a2[qq][W - 1] = 1.f;   // multiplier of the reference row's norm column (packed FFMA2 path)
            Aq1[qq] = 0.f;
            if (DUAL) {  // columns 6, 7 of a packed dual row are the two norms, not coordinates
                Aq1[qq] = -0.5f * a2[qq][6];
                Aq[qq] += Aq1[qq];
                a2[qq][6] = 0.f;
            }
            thr[qq] = valid ? CUDART_INF_F : -CUDART_INF_F;
            thr1[qq] = thr[qq];
            nthr2[qq] = make_float2(-thr[qq], -thr[qq]);
            tauf[qq] = valid ? CUDART_INF_F : 0.f;
            qrow[qq] = valid ? (A.qperm ? (int64_t)__ldg(A.qperm + qi) : qi) : 0;
            for (int c = 0; c < S; ++c)
                xq[c * LS + col] = (valid && c < A.M.dim) ? __ldg(A.xquery + qrow[qq] * A.ldx + c) : 0.f;
            for (int c = 0; c < kc; ++c) {
                lv[c * LS + col] = CUDART_INF_F;
                li[c * LS + col] = 0x7fffffff;
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
            }
            if ((tid & 31) == 0) {
                ctl->wbox[tid >> 5][c] = (W - 1 > c) ? lo[c] : 0.f;
                ctl->wbox[tid >> 5][3 + c] = (W - 1 > c) ? hi[c] : 0.f;
            }
        }
        // a warp without a single valid query must not hold the CTA's culling bound at infinity
        float w0 = 0.f;
#pragma unroll
        for (int qq = 0; qq < Q; ++qq) w0 = fmaxf(w0, tauf[qq]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w0 = fmaxf(w0, __shfl_xor_sync(0xffffffffu, w0, o));
        if ((tid & 31) == 0) ctl->wtau[tid >> 5] = w0;
    }
    __syncthreads();

    if (tid >= KNN_CTHREADS) {
        // ================= producer warp: lane 0 drives the TMA ring =================
        const int lane = tid & 31;
        int slot = 0;
        unsigned long long loaded = 0;
        auto push = [&](int64_t t, const float* tb) {  // t < 0: end marker; tb: the tile's box (lane 0) or NULL
            if (lane == 0) {
                const int s = slot % KNN_STAGES;
                mbar_wait_relaxed(empty_s + 8 * s, ((slot / KNN_STAGES) & 1) ^ 1);
                ctl->tile_id[s] = (int)t;
                if (tb)
#pragma unroll
                    for (int c = 0; c < 6; ++c) ctl->tbox[s][c] = tb[c];
                if (t < 0) {
                    mbar_arrive(full_s + 8 * s);
                } else {
                    const int64_t r0 = t * KNN_TILE;
                    int64_t cnt = ref_end - r0;
                    if (cnt > KNN_TILE) cnt = KNN_TILE;
                    const uint32_t fbytes = (uint32_t)(cnt * W * sizeof(float));
                    const uint32_t xbytes = (uint32_t)(((cnt + 3) & ~(int64_t)3) * S * sizeof(float));
                    const uint32_t dst = ring_s + (uint32_t)(s * stage_floats * sizeof(float));
                    mbar_expect_tx(full_s + 8 * s, fbytes + xbytes);
                    tma_load_1d(dst, A.pref + r0 * W, fbytes, full_s + 8 * s);
                    tma_load_1d(dst + KNN_TILE * W * 4, A.xs + r0 * S, xbytes, full_s + 8 * s);
                    ++loaded;
                }
            }
            ++slot;
            __syncwarp();
        };
        if (!cull) {
            for (int64_t t = tile_lo; t < tile_hi; ++t) push(t, nullptr);
        } else {
            // outward walk from the tile that holds this CTA's own rows (queries and references
            // share the spatial order), 16 boxes per side and step; a tile is loaded only if it
            // can still matter when its turn comes
            float cbox[6];
            for (int c = 0; c < 3; ++c) {
                cbox[c] = CUDART_INF_F;
                cbox[3 + c] = -CUDART_INF_F;
                for (int w = 0; w < KNN_CWARPS; ++w) {
                    cbox[c] = fminf(cbox[c], ctl->wbox[w][c]);
                    cbox[3 + c] = fmaxf(cbox[3 + c], ctl->wbox[w][3 + c]);
                }
            }
            int64_t centre = (int64_t)((double)(q_first + LS / 2) / (double)(A.n_query > 0 ? A.n_query : 1) *
                                       (double)total_tiles);
            if (centre < tile_lo) centre = tile_lo;
            if (centre >= tile_hi) centre = tile_hi - 1;
            // KNN_WALK blocks of 16 tiles per side and step: the boxes of a step are loaded together
            // (one L2 latency), the blocks are then handled nearest first.  Far from the CTA's rows
            // whole steps fail the test, and the consumers sit idle until the walk is over
            int64_t up = centre, down = centre - 1;
            while (up < tile_hi || down >= tile_lo) {
                int t[KNN_WALK];
                bool in[KNN_WALK];
                float tb[KNN_WALK][6], d2[KNN_WALK];
#pragma unroll
                for (int i = 0; i < KNN_WALK; ++i) {
                    const int64_t ti = lane < 16 ? up + 16 * i + lane : down - 16 * i - (lane - 16);
                    in[i] = lane < 16 ? ti < tile_hi : ti >= tile_lo;
                    t[i] = (int)ti;
#pragma unroll
                    for (int c = 0; c < 6; ++c) tb[i][c] = in[i] ? __ldg(A.boxes + ti * 6 + c) : 0.f;
                }
                // one test per tile against the CTA's box and its loosest bound; testing every
                // consumer warp's own box here loads fewer tiles but was measured 6 % slower (the
                // producer's push latency is on the consumers' critical path)
#pragma unroll
                for (int i = 0; i < KNN_WALK; ++i) d2[i] = in[i] ? box_dist2(cbox, tb[i]) : CUDART_INF_F;
                up += 16 * KNN_WALK;
                down -= 16 * KNN_WALK;
#pragma unroll
                for (int i = 0; i < KNN_WALK; ++i) {
                    float worst = 0.f;
                    for (int w = 0; w < KNN_CWARPS; ++w) worst = fmaxf(worst, ctl->wtau[w]);
                    const unsigned cand =
                        __ballot_sync(0xffffffffu, in[i] && !(d2[i] > worst * (1.f + 1e-4f) + 1e-12f));
                    if (!cand) continue;
                    // nearest first: up[0], down[0], up[1], ... -- bit r of `order` = lane (r >> 1) + 16 (r & 1)
                    unsigned order = __ballot_sync(0xffffffffu, (cand >> ((lane >> 1) + ((lane & 1) << 4))) & 1u);
                    while (order) {
                        const int r = __ffs(order) - 1;
                        order &= order - 1;
                        const int l = (r >> 1) + ((r & 1) << 4);
                        const int tt = __shfl_sync(0xffffffffu, t[i], l);
                        const float dd = __shfl_sync(0xffffffffu, d2[i], l);
                        float sb[6];  // the consumers' own cull test reads the box from shared memory
#pragma unroll
                        for (int c = 0; c < 6; ++c) sb[c] = __shfl_sync(0xffffffffu, tb[i][c], l);
                        int go = 0;
                        if (lane == 0) {  // the bounds have tightened while earlier tiles were scanned: test again
                            float now = 0.f;
                            for (int w = 0; w < KNN_CWARPS; ++w) now = fmaxf(now, ctl->wtau[w]);
                            go = !(dd > now * (1.f + 1e-4f) + 1e-12f);
                        }
                        go = __shfl_sync(0xffffffffu, go, 0);
                        if (go) push(tt, sb);
                    }
                }
            }
        }
        push(-1, nullptr);
        if (A.stats && lane == 0) {
            atomicAdd(A.stats + 3, loaded);
            atomicMax(A.stats + 4, loaded);
        }
        return;
    }

    // ================= consumer warps =================
    const int warp = tid >> 5;
    const float* tile_x = nullptr;   // exact rows of the tile being scanned
    int64_t tile_r0 = 0;
    constexpr int WQ_CAP = 32 * Q * KNN_QC;
    unsigned short* wq = wqueue + warp * WQ_CAP;
    int wcnt = 0;  // entries in the warp's queue (warp uniform)
    const int lane = tid & 31;
    const unsigned lanes_below = (1u << lane) - 1u;

    // Drain: the queue is a compact list of (owner, row) pairs, 32 per round, whichever lane
    // queued them.  Exact fp32 metric of the owner's query row (shared-memory column) against
    // the tile's exact row (shared memory); the few values that beat the owner's k-th best are
    // inserted into the owner's list by the lane that evaluated them -- lanes that hold values
    // for the same owner take turns (match_any).  Candidate order is (value, scan position): the
    // result does not depend on visit order.
    auto drain = [&]() {
        const int wbase = warp * 32;
        __syncwarp();
        for (int g0 = 0; g0 < wcnt; g0 += 32) {
            const int g = g0 + lane;
            const bool valid = g < wcnt;
            const unsigned ent = wq[valid ? g : 0];
            const int j = ent & 0xff;
            const int oq = ent >> 13, ol = (ent >> 8) & 31;
            const int ocol = oq * KNN_CTHREADS + wbase + ol;
            const int64_t pos = tile_r0 + j;
            bool ok = valid && pos < ref_end;
            if ((A.flags & 5) == 5) ok = ok && pos != q_first + ocol;  // same scan order: self = same position
            if (A.flags & 2) ok = ok && pos < A.query_id0 + q_first + ocol;  // causal searches keep row order
            if ((A.flags & 5) == 1) {  // self, found by original id (queries in their own order)
                int64_t orow = 0;
#pragma unroll
                for (int qq = 0; qq < Q; ++qq) {
                    const int64_t r = __shfl_sync(0xffffffffu, qrow[qq], ol);
                    if (qq == oq) orow = r;
                }
                const int64_t orig = (A.perm && ok) ? (int64_t)__ldg(A.perm + pos) : pos;
                ok = ok && orig != A.query_id0 + orow;
            }
            const float v = exact_metric(A.M, xq + ocol, LS, tile_x + (size_t)j * S);
            float* lvc = lv + ocol;
            int* lic = li + ocol;
            bool pass = ok && v < CUDART_INF_F;
            if (pass) {
                const float tv = lvc[(kc - 1) * LS];
                pass = v < tv || (v == tv && (int)pos < lic[(kc - 1) * LS]);
            }
            // (A warp-cooperative insertion -- lane c holds slot c of one owner's list, a ballot gives the
            //  position, the lanes behind it shift -- was measured in round 2: converged, but it takes the
            //  passing values one at a time, and several lanes inserting into different lists at once
            //  win: FK 24.3 -> 26.1 ms, 7-D L2 53.8 -> 58.3 ms.  The turn-taking loop stays.)
            unsigned pm = __ballot_sync(0xffffffffu, pass);
            while (pm) {
                if (pass) {
                    const unsigned peers = __match_any_sync(pm, ocol);
                    if (__ffs(peers) - 1 == lane) {
                        const float tv = lvc[(kc - 1) * LS];  // the list as it is now
                        if (v < tv || (v == tv && (int)pos < lic[(kc - 1) * LS])) {
                            int p = kc - 1;
                            while (p > 0) {
                                const float u = lvc[(p - 1) * LS];
                                const int ui = lic[(p - 1) * LS];
                                if (!(u > v || (u == v && ui > (int)pos))) break;
                                lvc[p * LS] = u;
                                lic[p * LS] = ui;
                                --p;
                            }
                            lvc[p * LS] = v;
                            lic[p * LS] = (int)pos;
                            if (A.stats) atomicAdd(A.stats + 1, 1ull);
                        }
                        pass = false;
                    }
                }
                __syncwarp();
                pm = __ballot_sync(0xffffffffu, pass);
            }
        }
        if (A.stats && lane == 0) atomicAdd(A.stats, (unsigned long long)wcnt);
        wcnt = 0;
        __syncwarp();
        // every thread reads its own bounds back from its lists
        float worst = 0.f;
#pragma unroll
        for (int qq = 0; qq < Q; ++qq) {
            const float tau = lv[(size_t)(kc - 1) * LS + qq * KNN_CTHREADS + tid];
            if (tau < CUDART_INF_F) {
                const float tf = ((A.M.kind == MMMP_METRIC_GROUP_L2_SUM) ? tau * tau : tau) * (1.f + KNN_SLACK);
                thr[qq] = (DUAL ? 2.f * tf : tf) - Aq[qq];
                thr1[qq] = tf - Aq1[qq];
                nthr2[qq] = make_float2(-thr1[qq], -(thr[qq] - thr1[qq]));
                tauf[qq] = tf;
            }
            worst = fmaxf(worst, tauf[qq]);
        }
        // publish the warp's worst bound for the producer's culling test
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
        if (lane == 0) {
            ctl->wtau[warp] = worst;
            if (A.stats) atomicAdd(A.stats + 2, 1ull);
        }
    };

    for (int slot = 0;; ++slot) {
        const int s = slot % KNN_STAGES;
        mbar_wait(full_s + 8 * s, (slot / KNN_STAGES) & 1);
        const int t = ctl->tile_id[s];
        if (t < 0) break;
        const int64_t r0 = (int64_t)t * KNN_TILE;
        const int cnt = (int)((ref_end - r0) < KNN_TILE ? (ref_end - r0) : KNN_TILE);
        bool skip = false;
        if (cull) {  // warp-level cull: this warp's box against the tile's box
            const float d2 = box_dist2(ctl->wbox[warp], ctl->tbox[s]);
            skip = d2 > ctl->wtau[warp] * (1.f + 1e-4f) + 1e-12f;
        }
        if (!skip) {
            const uint32_t tile_s = ring_s + (uint32_t)(s * stage_floats * sizeof(float));
            tile_x = ring + s * stage_floats + KNN_TILE * W;
            tile_r0 = r0;
            const int cnt4 = (cnt + KNN_RU - 1) & ~(KNN_RU - 1);  // stale rows of a partial tile are rejected by position
            // software pipeline: the rows of step j+1 are loaded while step j is evaluated
#if MMMP_KNN_PREFETCH
            float4 rn[KNN_RU][F4];
#pragma unroll
            for (int u = 0; u < KNN_RU; ++u)
#pragma unroll
                for (int i = 0; i < F4; ++i) rn[u][i] = lds128(tile_s + (uint32_t)(u * W + 4 * i) * 4);
#endif
#pragma unroll 1
            for (int j = 0; j < cnt4; j += KNN_RU) {
                float4 r[KNN_RU][F4];
#if MMMP_KNN_PREFETCH
#pragma unroll
                for (int u = 0; u < KNN_RU; ++u)
#pragma unroll
                    for (int i = 0; i < F4; ++i) r[u][i] = rn[u][i];
                const int jn = (j + KNN_RU < KNN_TILE) ? j + KNN_RU : 0;  // last prefetch wraps: harmless
#pragma unroll
                for (int u = 0; u < KNN_RU; ++u)
#pragma unroll
                    for (int i = 0; i < F4; ++i) rn[u][i] = lds128(tile_s + (uint32_t)((jn + u) * W + 4 * i) * 4);
#else
#pragma unroll
                for (int u = 0; u < KNN_RU; ++u)
#pragma unroll
                    for (int i = 0; i < F4; ++i) r[u][i] = lds128(tile_s + (uint32_t)((j + u) * W + 4 * i) * 4);
#endif
                float sv[Q][KNN_RU];         // filter value - threshold (DUAL: the larger of the two)
                float worst = CUDART_INF_F;  // min over queries and rows
#pragma unroll
                for (int qq = 0; qq < Q; ++qq) {
#pragma unroll
                    for (int u = 0; u < KNN_RU; ++u) {
                        if (DUAL) {
                            // row = (c0 d0 c1 d1 | c2 d2 nC nD), query = (-2 c0, -2 d0, ...): three packed
                            // FFMA2 (sm_100: two fp32 FMAs per instruction) advance (accC, accD) together
                            float2 acc = make_float2(r[u][F4 - 1].z, r[u][F4 - 1].w);
                            acc = __ffma2_rn(make_float2(a2[qq][0], a2[qq][1]), make_float2(r[u][0].x, r[u][0].y), acc);
                            acc = __ffma2_rn(make_float2(a2[qq][2], a2[qq][3]), make_float2(r[u][0].z, r[u][0].w), acc);
                            acc = __ffma2_rn(make_float2(a2[qq][4], a2[qq][5]), make_float2(r[u][F4 - 1].x, r[u][F4 - 1].y), acc);
                            // both tests from one packed add: t = (sC - thr1, sD - (thr - thr1)); the second
                            // test reads t.x + t.y (rounding differs from sC + sD - thr by ulps, far inside
                            // the filter's slack)
                            const float2 t = __fadd2_rn(acc, nthr2[qq]);
                            sv[qq][u] = fmaxf(t.x, t.x + t.y);
                        } else if (F4 >= 2) {
                            // even and odd columns accumulate in the two halves of a packed FFMA2 (the
                            // norm column enters with multiplier 1): half the FMA instructions
                            float2 acc = make_float2(-thr[qq], 0.f);
#pragma unroll
                            for (int i = 0; i < F4; ++i) {
                                acc = __ffma2_rn(make_float2(a2[qq][4 * i + 0], a2[qq][4 * i + 1]),
                                                 make_float2(r[u][i].x, r[u][i].y), acc);
                                acc = __ffma2_rn(make_float2(a2[qq][4 * i + 2], a2[qq][4 * i + 3]),
                                                 make_float2(r[u][i].z, r[u][i].w), acc);
                            }
                            sv[qq][u] = acc.x + acc.y;
                        } else {
                            float acc = r[u][F4 - 1].w;  // B_r
#pragma unroll
                            for (int i = 0; i < F4; ++i) {
                                acc = fmaf(a2[qq][4 * i + 0], r[u][i].x, acc);
                                acc = fmaf(a2[qq][4 * i + 1], r[u][i].y, acc);
                                acc = fmaf(a2[qq][4 * i + 2], r[u][i].z, acc);
                                if (i < F4 - 1) acc = fmaf(a2[qq][4 * i + 3], r[u][i].w, acc);
                            }
                            sv[qq][u] = acc - thr[qq];
                        }
                    }
                    float m = sv[qq][0];
#pragma unroll
                    for (int u = 1; u < KNN_RU; ++u) m = fminf(m, sv[qq][u]);
                    worst = fminf(worst, m);
                }
                // one warp-uniform branch per step
                if (__any_sync(0xffffffffu, worst <= 0.f)) {
#pragma unroll
                    for (int qq = 0; qq < Q; ++qq) {
#pragma unroll
                        for (int u = 0; u < KNN_RU; ++u) {
                            const bool hit = sv[qq][u] <= 0.f;
                            const unsigned m = __ballot_sync(0xffffffffu, hit);
                            if (hit)
                                wq[wcnt + __popc(m & lanes_below)] =
                                    (unsigned short)((((qq << 5) | lane) << 8) | (j + u));
                            wcnt += __popc(m);
                        }
                    }
                    if (wcnt > WQ_CAP - 32 * Q * KNN_RU) drain();  // room for one more step
                }
            }
            // the tile's exact rows leave with the tile: drain before releasing the slot
            if (wcnt) drain();
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty_s + 8 * s);
    }

    // ---- write the candidate columns (rows in ORIGINAL query order) ----------------
#pragma unroll
    for (int qq = 0; qq < Q; ++qq) {
        const int64_t qi = q_first + qq * KNN_CTHREADS + tid;
        if (qi >= A.n_query) continue;
        const int col = qq * KNN_CTHREADS + tid;
        size_t o = ((size_t)qrow[qq] * A.n_splits + blockIdx.y) * kc;
        if (A.out_rows) {
            const int64_t r = qi - (int64_t)A.qblock0 * LS;
            o = ((size_t)r * A.n_splits + blockIdx.y) * kc;
            if (blockIdx.y == 0) A.out_rows[r] = (int32_t)qrow[qq];
        }
        for (int c = 0; c < kc; ++c) {
            const int id = li[c * LS + col];
            A.out_idx[o + c] = id == 0x7fffffff ? -1 : (A.perm ? __ldg(A.perm + id) : id);
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

size_t scan_smem(int F4, int Q, int kc, int S) {
    const size_t w = F4 * 4, ls = (size_t)KNN_CTHREADS * Q;
    size_t b = (size_t)KNN_STAGES * knn_stage_floats((int)w, S) * 4 + (size_t)kc * ls * 8 + (size_t)S * ls * 4;
    b = knn_align(b + (size_t)Q * knn_qc(F4) * KNN_CTHREADS * 2, 8);
    return b + 2 * KNN_STAGES * 8 + sizeof(ScanCtl) + 16;
}

template <int F4, int Q, bool DUAL = false>
int launch_scan(const KnnArgs& A, int n_qblocks, void* stream) {
    const size_t smem = scan_smem(F4, Q, A.kc, A.S);
    // the attribute belongs to the function: set to the device maximum, never to this launch's need -- two
    // host threads searching with different candidate counts must not lower it under each other
    int dev = 0, max_smem = 0;
    MMMP_CUDA_TRY(cudaGetDevice(&dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)max_smem) return MMMP_ERR_LIMIT;
    MMMP_CUDA_TRY(cudaFuncSetAttribute(knn_scan_kernel<F4, Q, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    dim3 grid(n_qblocks, A.n_splits, 1);
    MMMP_LAUNCH((knn_scan_kernel<F4, Q, DUAL>), grid, KNN_THREADS, smem, stream, A);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

unsigned long long* g_knn_stats = nullptr;  // device counters, enabled by mmmp_knn_stats

int grid_for(int64_t n, int sms) {
    int g = mmmp_div_up(n, 256);
    return g > sms * 16 ? sms * 16 : (g < 1 ? 1 : g);
}

// counting sort of rows by grid cell along the Hilbert curve: perm[pos] = original row
int spatial_perm(const float* rows, int ld, int nd, int64_t n, const int* bbox_d, int bits, int32_t* perm,
                 void* stream, int sms) {
    cudaStream_t s = (cudaStream_t)stream;
    const int n_cells = 1 << (3 * bits);
    uint32_t* keys = nullptr;
    int32_t* hist = nullptr;
    int64_t* offs = nullptr;
    cudaError_t ea = cudaMallocAsync(&keys, sizeof(uint32_t) * n, s);
    if (ea == cudaSuccess) ea = cudaMallocAsync(&hist, sizeof(int32_t) * n_cells * 2, s);
    if (ea == cudaSuccess) ea = cudaMallocAsync(&offs, sizeof(int64_t) * (n_cells + 1), s);
    if (ea == cudaSuccess) ea = cudaMemsetAsync(hist, 0, sizeof(int32_t) * n_cells * 2, s);
    if (ea != cudaSuccess) {  // nothing allocated so far may leak
        if (keys) cudaFreeAsync(keys, s);
        if (hist) cudaFreeAsync(hist, s);
        if (offs) cudaFreeAsync(offs, s);
        return MMMP_ERR_CUDA_BASE + (int)ea;
    }
    MMMP_LAUNCH(cell_hist_kernel, grid_for(n, sms), 256, 0, stream, rows, ld, nd, n, bbox_d, bits, keys, hist);
    int rc = mmmp_exclusive_scan_i32(hist, n_cells, offs, stream);
    if (rc == MMMP_OK) {
        int32_t* unordered = reinterpret_cast<int32_t*>(keys);  // the keys are dead once scattered ...
        int32_t* scratch = nullptr;                             // ... but are read while it runs
        cudaError_t e = cudaMallocAsync(&scratch, sizeof(int32_t) * n, s);
        if (e == cudaSuccess) {
            unordered = scratch;
            MMMP_LAUNCH(cell_scatter_kernel, grid_for(n, sms), 256, 0, stream, keys, n, offs, hist + n_cells, unordered);
            int g = mmmp_div_up((int64_t)n_cells * 32, 256);
            if (g > sms * 32) g = sms * 32;
            MMMP_LAUNCH(cell_order_kernel, g, 256, 0, stream, unordered, offs, n_cells, perm);
            e = cudaPeekAtLastError();
            cudaFreeAsync(scratch, s);
        }
        if (e != cudaSuccess) rc = MMMP_ERR_CUDA_BASE + (int)e;
    }
    cudaFreeAsync(keys, s);
    cudaFreeAsync(hist, s);
    cudaFreeAsync(offs, s);
    return rc;
}

}  // namespace

extern "C" int mmmp_scan_order_cell(int x0, int x1, int x2, int bits, uint32_t* index_h) {
    if (!index_h || bits < 1 || bits > 10) return MMMP_ERR_ARG;
    const int hi = 1 << bits;
    if (x0 < 0 || x1 < 0 || x2 < 0 || x0 >= hi || x1 >= hi || x2 >= hi) return MMMP_ERR_ARG;
    *index_h = hilbert3((uint32_t)x0, (uint32_t)x1, (uint32_t)x2, bits);
    return MMMP_OK;
}

extern "C" int mmmp_knn_stats(int enable, unsigned long long* out5_h) {
    constexpr size_t bytes = 5 * sizeof(unsigned long long);
    if (enable && !g_knn_stats) {
        MMMP_CUDA_TRY(cudaMalloc(&g_knn_stats, bytes));
        MMMP_CUDA_TRY(cudaMemset(g_knn_stats, 0, bytes));
    }
    if (out5_h && g_knn_stats) {
        MMMP_CUDA_TRY(cudaDeviceSynchronize());
        MMMP_CUDA_TRY(cudaMemcpy(out5_h, g_knn_stats, bytes, cudaMemcpyDeviceToHost));
        MMMP_CUDA_TRY(cudaMemset(g_knn_stats, 0, bytes));
    }
    if (!enable && g_knn_stats) {
        cudaFree(g_knn_stats);
        g_knn_stats = nullptr;
    }
    return MMMP_OK;
}

// frac_lo >= 0: the queries are the references and only the query blocks
// [floor(B * frac_lo), floor(B * frac_hi)) of the B blocks of the scan order are searched;
// output rows are compact (out_rows names
// the original row of each) and *n_rows_h receives their number
static int knn_impl(const float* fref, const float* xref, int64_t n_ref, const float* fquery, const float* xquery,
                    int64_t n_query, int ldf, int fdim, int ldx, const mmmp_metric* metric, int k, int exclude_self,
                    int causal, int64_t query_id0, double frac_lo, double frac_hi, int32_t* out_rows, int64_t* n_rows_h,
                    int32_t* out_idx, float* out_dist, void* stream) {
    if (!fref || !xref || !fquery || !xquery || !out_idx || n_ref < 0 || n_query < 0 || k < 1 || k > MMMP_MAX_K)
        return MMMP_ERR_ARG;
    if (n_ref >= (1ll << 28)) return MMMP_ERR_LIMIT;
    if (fdim < 1 || fdim > ldf || fdim > 31) return MMMP_ERR_ARG;
    KnnArgs A;
    int rc = fill_metric(metric, ldx, A.M);
    if (rc) return rc;
    const int W = width_for(fdim);
    if (!W) return MMMP_ERR_LIMIT;
    // dual rows of mmmp_fk_features: (C, D, 0, 0), two tests per pair (see the kernel)
    const int dual = (A.M.kind == MMMP_METRIC_GROUP_L2_SUM && fdim == 6 && ldf >= 8) ? 1 : 0;
    const int F4 = W / 4;
    const int S = A.M.dim | 1;  // odd row stride: exact rows of a tile fall into distinct banks
    if (n_query == 0) return MMMP_OK;
    mmmp_pool_keep();
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 148, max_smem = 0;
    MMMP_CUDA_TRY(cudaGetDevice(&dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // queries per thread: measured on B200 in round 1 (1M x 1M, FK and joint-space metrics), one
    // query per thread wins -- smaller CTA boxes cull more tiles and more CTAs fit an SM; the Q = 2
    // instantiations are gone (the kernel keeps its Q parameter)
    const bool ranged = frac_lo >= 0.0;
    const int Q = 1;
    if (scan_smem(F4, Q, k, S) > (size_t)max_smem) return MMMP_ERR_LIMIT;
    bool cull = !causal && n_ref >= KNN_CULL_MIN;
    if (const char* e = getenv("MMMP_KNN_CULL")) cull = cull && atoi(e) != 0;
    const int all_qblocks = mmmp_div_up(n_query, (int64_t)KNN_CTHREADS * Q);
    int qblock0 = 0, n_qblocks = all_qblocks;
    int64_t n_rows = n_query;  // rows this launch answers
    if (ranged) {
        qblock0 = (int)((double)all_qblocks * frac_lo);
        int qblock1 = frac_hi >= 1.0 ? all_qblocks : (int)((double)all_qblocks * frac_hi);
        if (qblock0 > all_qblocks) qblock0 = all_qblocks;
        if (qblock1 > all_qblocks) qblock1 = all_qblocks;
        if (qblock1 < qblock0) qblock1 = qblock0;
        n_qblocks = qblock1 - qblock0;
        const int64_t first = (int64_t)qblock0 * KNN_CTHREADS * Q, end = ((int64_t)qblock0 + n_qblocks) * KNN_CTHREADS * Q;
        n_rows = (end < n_query ? end : n_query) - first;
        if (n_rows < 0) n_rows = 0;
        if (n_rows_h) *n_rows_h = n_rows;
        if (n_rows == 0) return MMMP_OK;
    }
    const int64_t total_tiles = (n_ref + KNN_TILE - 1) / KNN_TILE;
    int n_splits = 1;
    if (!cull && n_qblocks < 2 * sms) {
        n_splits = (2 * sms + n_qblocks - 1) / n_qblocks;
        if (n_splits > 32) n_splits = 32;
        if (n_splits > total_tiles) n_splits = (int)(total_tiles > 0 ? total_tiles : 1);
    }
    const bool same = (fquery == fref && xquery == xref && n_query == n_ref);
    // temporaries (library-owned, stream ordered)
    float *pref = nullptr, *pquery = nullptr, *boxes = nullptr, *xs = nullptr;
    int32_t *perm = nullptr, *qperm = nullptr, *tmp_idx = nullptr, *order = nullptr;
    int* bbox = nullptr;
    float* tmp_val = nullptr;
    const int nd = fdim < 3 ? fdim : 3;
    const size_t n_out = (size_t)n_rows * n_splits * k;
    const int64_t n_ref4 = (n_ref + 3) & ~(int64_t)3;
    auto cleanup = [&]() {
        void* ptrs[] = {pref, same ? nullptr : pquery, boxes, xs, perm, same ? nullptr : qperm, bbox, tmp_idx, tmp_val, order};
        for (void* p : ptrs)
            if (p) cudaFreeAsync(p, s);
    };
#define KNN_TRY(expr)                                       \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) {                            \
            cleanup();                                      \
            return MMMP_ERR_CUDA_BASE + (int)_e;            \
        }                                                   \
    } while (0)
    KNN_TRY(cudaMallocAsync(&pref, (size_t)(n_ref > 0 ? n_ref : 1) * W * sizeof(float), s));
    KNN_TRY(cudaMallocAsync(&xs, (size_t)(n_ref4 > 0 ? n_ref4 : 4) * S * sizeof(float), s));
    if (!same) KNN_TRY(cudaMallocAsync(&pquery, (size_t)n_query * W * sizeof(float), s));
    KNN_TRY(cudaMallocAsync(&tmp_idx, n_out * sizeof(int32_t), s));
    KNN_TRY(cudaMallocAsync(&tmp_val, n_out * sizeof(float), s));
    if (cull) {
        // grid resolution: 2..16 rows per cell, 1..7 bits per axis (measured with the Hilbert order, 1 M rows:
        // 4 / 5 / 6 / 7 bits -> 28.6 / 19.6 / 19.0 / 19.0 ms; rows of one cell are in row order, so a finer
        // grid makes the runs more compact until a cell holds a handful of rows)
        int bits = 1;
        while (bits < 7 && ((int64_t)1 << (3 * (bits + 1))) * 2 <= n_ref) ++bits;
        if (const char* e = getenv("MMMP_KNN_BITS")) bits = atoi(e) < 1 ? 1 : (atoi(e) > 8 ? 8 : atoi(e));
        KNN_TRY(cudaMallocAsync(&bbox, 6 * sizeof(int), s));
        const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
        KNN_TRY(cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, s));
        MMMP_LAUNCH(bbox_kernel, grid_for(n_ref, sms), 256, 0, stream, fref, ldf, nd, n_ref, bbox);
        KNN_TRY(cudaMallocAsync(&perm, sizeof(int32_t) * n_ref, s));
        rc = spatial_perm(fref, ldf, nd, n_ref, bbox, bits, perm, stream, sms);
        if (rc == MMMP_OK && !same) {
            KNN_TRY(cudaMallocAsync(&qperm, sizeof(int32_t) * n_query, s));
            rc = spatial_perm(fquery, ldf, nd, n_query, bbox, bits, qperm, stream, sms);
        }
        if (rc != MMMP_OK) {
            cleanup();
            return rc;
        }
        KNN_TRY(cudaMallocAsync(&boxes, sizeof(float) * 6 * (total_tiles > 0 ? total_tiles : 1), s));
    }
    if (n_ref > 0) {
        MMMP_LAUNCH(knn_pack_kernel, grid_for(n_ref, sms), 256, 0, stream, fref, ldf, fdim, n_ref, perm, pref, W, dual);
        MMMP_LAUNCH(knn_pack_exact_kernel, grid_for(n_ref4 * S, sms), 256, 0, stream, xref, ldx, A.M.dim, n_ref, n_ref4,
                    perm, xs, S);
    }
    if (same) {
        pquery = pref;
        qperm = perm;
    } else {
        MMMP_LAUNCH(knn_pack_kernel, grid_for(n_query, sms), 256, 0, stream, fquery, ldf, fdim, n_query, qperm, pquery, W,
                    dual);
    }
    if (cull && total_tiles > 0)
        MMMP_LAUNCH(tile_box_kernel, (int)total_tiles, KNN_TILE, 0, stream, pref, W, nd, dual ? 2 : 1, n_ref, boxes);
    A.pref = pref;
    A.pquery = pquery;
    A.xs = xs;
    A.xquery = xquery;
    A.perm = perm;
    A.qperm = qperm;
    A.boxes = boxes;
    A.block_order = nullptr;
    if (cull && same && query_id0 == 0 && (int64_t)KNN_CTHREADS * Q == KNN_TILE && n_qblocks > 1 &&
        getenv("MMMP_KNN_ORDER")) {   // the query blocks are the tiles: their boxes are at hand
        cudaError_t eo = cudaMallocAsync(&order, sizeof(int32_t) * n_qblocks, s);
        if (eo != cudaSuccess) {
            cleanup();
            return MMMP_ERR_CUDA_BASE + (int)eo;
        }
        MMMP_LAUNCH(block_order_kernel, 1, 1024, 0, stream, boxes, qblock0, n_qblocks, order);
        A.block_order = order;
    }
    A.n_ref = n_ref;
    A.n_query = n_query;
    A.query_id0 = query_id0;
    A.ldx = ldx;
    A.S = S;
    A.kc = k;
    A.flags = (exclude_self ? 1 : 0) | (causal ? 2 : 0) | ((same && query_id0 == 0) ? 4 : 0);
    A.n_splits = n_splits;
    A.qblock0 = qblock0;
    A.out_rows = ranged ? out_rows : nullptr;
    A.tiles_per_split = (int)((total_tiles + n_splits - 1) / n_splits);
    if (A.tiles_per_split < 1) A.tiles_per_split = 1;
    A.stats = g_knn_stats;
    A.out_idx = tmp_idx;
    A.out_val = tmp_val;
    switch (dual ? 0 : F4 * 10 + Q) {
        case 0: rc = launch_scan<2, 1, true>(A, n_qblocks, stream); break;
        case 11: rc = launch_scan<1, 1>(A, n_qblocks, stream); break;
        case 21: rc = launch_scan<2, 1>(A, n_qblocks, stream); break;
        case 41: rc = launch_scan<4, 1>(A, n_qblocks, stream); break;
        case 81: rc = launch_scan<8, 1>(A, n_qblocks, stream); break;
        default: rc = MMMP_ERR_LIMIT;
    }
    if (rc == MMMP_OK) {
        int g = mmmp_div_up(n_rows, 128);
        if (g > sms * 8) g = sms * 8;
        MMMP_LAUNCH(knn_merge_kernel, g, 128, 0, stream, tmp_idx, tmp_val, n_rows, n_splits, k, A.M.kind, out_idx,
                    out_dist);
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) rc = MMMP_ERR_CUDA_BASE + (int)e;
    }
    cleanup();
#undef KNN_TRY
    return rc;
}

extern "C" int mmmp_knn(const float* fref, const float* xref, int64_t n_ref, const float* fquery,
                        const float* xquery, int64_t n_query, int ldf, int fdim, int ldx, const mmmp_metric* metric,
                        int k, int exclude_self, int causal, int64_t query_id0, int32_t* out_idx, float* out_dist,
                        void* stream) {
    return knn_impl(fref, xref, n_ref, fquery, xquery, n_query, ldf, fdim, ldx, metric, k, exclude_self, causal,
                    query_id0, -1.0, -1.0, nullptr, nullptr, out_idx, out_dist, stream);
}

extern "C" int mmmp_knn_range(const float* fref, const float* xref, int64_t n_ref, int ldf, int fdim, int ldx,
                              const mmmp_metric* metric, int k, int exclude_self, double frac_lo, double frac_hi,
                              int32_t* out_rows, int64_t* n_rows_h, int32_t* out_idx, float* out_dist, void* stream) {
    if (!(frac_lo >= 0.0) || !(frac_hi >= frac_lo) || frac_hi > 1.0 || !out_rows || !n_rows_h) return MMMP_ERR_ARG;
    return knn_impl(fref, xref, n_ref, fref, xref, n_ref, ldf, fdim, ldx, metric, k, exclude_self, 0, 0, frac_lo, frac_hi,
                    out_rows, n_rows_h, out_idx, out_dist, stream);
}

extern "C" int mmmp_knn_shard(const float* fref, const float* xref, int64_t n_ref, int ldf, int fdim, int ldx,
                              const mmmp_metric* metric, int k, int exclude_self, int shard, int n_shards,
                              int32_t* out_rows, int64_t* n_rows_h, int32_t* out_idx, float* out_dist, void* stream) {
    if (shard < 0 || n_shards < 1 || shard >= n_shards) return MMMP_ERR_ARG;
    return mmmp_knn_range(fref, xref, n_ref, ldf, fdim, ldx, metric, k, exclude_self, (double)shard / n_shards,
                          shard + 1 == n_shards ? 1.0 : (double)(shard + 1) / n_shards, out_rows, n_rows_h, out_idx,
                          out_dist, stream);
}
