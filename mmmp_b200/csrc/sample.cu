// Kernel (a): batched joint-space sampling, plus the small data-movement
// kernels around it (fp64->fp32 packing, stable compaction of the collision
// free rows, row gather, exclusive scan).
//
// Reference: np.random.uniform(lo, hi) ; np.round(sample, 2)
//   planners/prm/pybullet/decoupled/decoupled_prm.py:310-318
//   planners/prm/pybullet/coupled/coupled_prm.py:88-94
//   planners/prm/planar/single_arm/prm.py:52-58
// The reference draws from numpy's global Mersenne Twister; a batched sampler
// cannot reproduce that stream, so the generator here is the counter-based
// Philox4x32-10 (sample i / joint j is a pure function of (seed, i, j)), and
// oracle/sampling.py restates it bit for bit.
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// 53-bit uniform in [0,1) from two 32-bit words (numpy's construction)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

struct SampleParams {
    double lo[32];
    double hi[32];
};

__global__ void __launch_bounds__(256)
sample_uniform_kernel(SampleParams P, int D, int64_t N, uint64_t seed, uint64_t first, double scale, int do_round,
                      double* __restrict__ out64, int ld64, float* __restrict__ out32, int ld32) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t ctr = first + (uint64_t)i;
        for (int jp = 0; jp * 2 < D; ++jp) {
            uint32_t r[4];
            philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)jp, 0u, (uint32_t)seed,
                          (uint32_t)(seed >> 32), r);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = jp * 2 + h;
                if (j >= D) break;
                const double u = u53(r[2 * h], r[2 * h + 1]);
                // lo + (hi - lo) * u, as numpy's uniform()
                double x = __dadd_rn(P.lo[j], __dmul_rn(__dsub_rn(P.hi[j], P.lo[j]), u));
                if (do_round) x = rint(__dmul_rn(x, scale)) / scale;  // np.round(x, dp)
                if (out64) out64[i * ld64 + j] = x;
                if (out32) out32[i * ld32 + j] = (float)x;
            }
        }
        if (out32)
            for (int j = D; j < ld32; ++j) out32[i * ld32 + j] = 0.f;
    }
}

__global__ void __launch_bounds__(256)
pack_f32_kernel(const double* __restrict__ in, int ld64, int D, int64_t N, float* __restrict__ out, int ld32) {
    const int64_t total = N * ld32;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / ld32;
        const int j = (int)(t - i * ld32);
        out[t] = j < D ? (float)in[i * ld64 + j] : 0.f;
    }
}

// ---- exclusive scan (int32 -> int64), 3 kernels, 2048 items per CTA --------
constexpr int SCAN_T = 256;
constexpr int SCAN_I = 8;
constexpr int SCAN_TILE = SCAN_T * SCAN_I;

template <class In>
__device__ __forceinline__ int load_item(const In* in, int64_t i);
template <>
__device__ __forceinline__ int load_item<int32_t>(const int32_t* in, int64_t i) { return in[i]; }
template <>
__device__ __forceinline__ int load_item<uint8_t>(const uint8_t* in, int64_t i) { return in[i] == 0 ? 1 : 0; }

template <class In>
__global__ void __launch_bounds__(SCAN_T) scan_tile_sums_kernel(const In* __restrict__ in, int64_t n, int64_t* tile_sums) {
    __shared__ int warp_sums[SCAN_T / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int s = 0;
    for (int k = 0; k < SCAN_I; ++k) {
        const int64_t i = base + k * SCAN_T + threadIdx.x;
        if (i < n) s += load_item<In>(in, i);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < SCAN_T / 32; ++w) t += warp_sums[w];
        tile_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int64_t* tile_sums, int64_t n_tiles, int64_t* total_out) {
    // single CTA: sequential over chunks of 1024 with a block-wide scan per chunk
    __shared__ int64_t sh[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t c = 0; c < n_tiles; c += 1024) {
        const int64_t i = c + threadIdx.x;
        const int64_t v = i < n_tiles ? tile_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int64_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        const int64_t incl = sh[threadIdx.x];
        if (i < n_tiles) tile_sums[i] = carry + incl - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// MODE 0: write exclusive offsets (int64) ; MODE 1: scatter indices of set items (compaction)
template <class In, int MODE>
__global__ void __launch_bounds__(SCAN_T)
scan_apply_kernel(const In* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_sums,
                  int64_t* __restrict__ out_offsets, int32_t* __restrict__ out_idx) {
    __shared__ int warp_sums[SCAN_T / 32];
    __shared__ int chunk_carry;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) chunk_carry = 0;
    __syncthreads();
    const int64_t tile_off = tile_sums[blockIdx.x];
    for (int k = 0; k < SCAN_I; ++k) {
        const int64_t i = base + k * SCAN_T + threadIdx.x;
        const int v = i < n ? load_item<In>(in, i) : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < wid; ++w) woff += warp_sums[w];
        const int excl = chunk_carry + woff + incl - v;
        if (i < n) {
            if (MODE == 0) out_offsets[i] = tile_off + excl;
            if (MODE == 1 && v) out_idx[tile_off + excl] = (int32_t)i;
        }
        __syncthreads();
        if (threadIdx.x == SCAN_T - 1) chunk_carry = excl + v;
        __syncthreads();
    }
}

__global__ void write_total_kernel(const int64_t* total, int64_t* out_offsets_n, int32_t* count) {
    if (out_offsets_n) *out_offsets_n = *total;
    if (count) *count = (int32_t)*total;
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint32_t* __restrict__ in, int row_words, const int32_t* __restrict__ idx, int64_t n,
                   uint32_t* __restrict__ out) {
    const int64_t total = n * row_words;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / row_words;
        const int j = (int)(t - i * row_words);
        out[t] = in[(int64_t)idx[i] * row_words + j];
    }
}

int grid_for(int64_t items, int block) {
    int64_t g = (items + block - 1) / block;
    const int64_t cap = 148 * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <class In, int MODE>
int scan_driver(const In* in, int64_t n, int64_t* out_offsets, int32_t* out_idx, int32_t* count, void* stream) {
    mmmp_pool_keep();
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* tmp = nullptr;
    MMMP_CUDA_TRY(cudaMallocAsync(&tmp, sizeof(int64_t) * (n_tiles + 1), s));
    MMMP_LAUNCH((scan_tile_sums_kernel<In>), (int)n_tiles, SCAN_T, 0, stream, in, n, tmp);
    MMMP_LAUNCH(scan_sums_kernel, 1, 1024, 0, stream, tmp, n_tiles, tmp + n_tiles);
    MMMP_LAUNCH((scan_apply_kernel<In, MODE>), (int)n_tiles, SCAN_T, 0, stream, in, n, tmp, out_offsets, out_idx);
    MMMP_LAUNCH(write_total_kernel, 1, 1, 0, stream, tmp + n_tiles, MODE == 0 ? out_offsets + n : nullptr, count);
    MMMP_LAUNCH_CHECK();
    MMMP_CUDA_TRY(cudaFreeAsync(tmp, s));
    return MMMP_OK;
}

}  // namespace

extern "C" int mmmp_sample_uniform(const double* lo, const double* hi, int D, int64_t N, uint64_t seed,
                                   uint64_t first_index, int round_dp, double* out64, int ld64, float* out32,
                                   int ld32, void* stream) {
    if (!lo || !hi || D < 1 || D > 32 || N < 0 || (!out64 && !out32)) return MMMP_ERR_ARG;
    if ((out64 && ld64 < D) || (out32 && ld32 < D)) return MMMP_ERR_ARG;
    if (N == 0) return MMMP_OK;
    SampleParams P;
    for (int j = 0; j < D; ++j) {
        P.lo[j] = lo[j];
        P.hi[j] = hi[j];
    }
    double scale = 1.0;
    for (int i = 0; i < round_dp; ++i) scale *= 10.0;
    MMMP_LAUNCH(sample_uniform_kernel, grid_for(N, 256), 256, 0, stream, P, D, N, seed, first_index, scale,
                round_dp >= 0 ? 1 : 0, out64, ld64, out32, ld32);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_pack_f32(const double* q64, int ld64, int D, int64_t N, float* out32, int ld32, void* stream) {
    if (!q64 || !out32 || D < 1 || ld64 < D || ld32 < D || N < 0) return MMMP_ERR_ARG;
    if (N == 0) return MMMP_OK;
    MMMP_LAUNCH(pack_f32_kernel, grid_for(N * ld32, 256), 256, 0, stream, q64, ld64, D, N, out32, ld32);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}

extern "C" int mmmp_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, void* stream) {
    if (!in || !out || n < 0) return MMMP_ERR_ARG;
    if (n == 0) {
        MMMP_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(int64_t), (cudaStream_t)stream));
        return MMMP_OK;
    }
    return scan_driver<int32_t, 0>(in, n, out, nullptr, nullptr, stream);
}

extern "C" int mmmp_compact_free(const uint8_t* flags, int64_t N, int32_t* idx_out, int32_t* count, void* stream) {
    if (!flags || !idx_out || !count || N < 0) return MMMP_ERR_ARG;
    if (N == 0) {
        MMMP_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), (cudaStream_t)stream));
        return MMMP_OK;
    }
    return scan_driver<uint8_t, 1>(flags, N, nullptr, idx_out, count, stream);
}

extern "C" int mmmp_gather_rows(const void* in, int row_bytes, const int32_t* idx, int64_t n, void* out,
                                void* stream) {
    if (!in || !idx || !out || row_bytes < 4 || (row_bytes & 3) || n < 0) return MMMP_ERR_ARG;
    if (n == 0) return MMMP_OK;
    const int rw = row_bytes / 4;
    MMMP_LAUNCH(gather_rows_kernel, grid_for(n * rw, 256), 256, 0, stream, (const uint32_t*)in, rw, idx, n,
                (uint32_t*)out);
    MMMP_LAUNCH_CHECK();
    return MMMP_OK;
}
