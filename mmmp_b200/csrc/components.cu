// Connected components of a roadmap held as a padded adjacency table.
//
// Replaces (reference): DecoupledPRM.compute_connected_components
// planners/prm/pybullet/decoupled/decoupled_prm.py:613-636 -- a Python depth-first search
// that looks every neighbour up with a linear scan of the node list (O(N * E)); unusable at
// 1 M nodes.  Same partition of the nodes, found with a lock-free union-find:
//   init    parent[v] = v
//   hook    one thread per table slot (v, c): the representatives of v and adj[v][c] are
//           found with path halving and the LARGER one is hooked under the smaller with
//           atomicCAS, retried from the value the CAS returns (parents only ever decrease, so
//           stale reads are harmless);
//   flatten label[v] = root of v = the SMALLEST node id of v's component, which is also the
//           node the reference's outer loop starts that component from, so components listed
//           by label come in the reference's order.
// An edge may appear in either direction or both; slots holding -1, or whose mask byte is 0,
// are not edges.
#include "common.cuh"

namespace {

__device__ __forceinline__ int cc_find(int* __restrict__ parent, int v) {
    int p = parent[v];
    while (p != v) {  // path halving
        const int g = parent[p];
        if (g != p) parent[v] = g;
        v = p;
        p = g;
    }
    return v;
}

__global__ void __launch_bounds__(256) cc_init_kernel(int* __restrict__ parent, int64_t n) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        parent[v] = (int)v;
}

__global__ void __launch_bounds__(256)
cc_hook_kernel(const int32_t* __restrict__ adj, const uint8_t* __restrict__ mask, int64_t n, int cap,
               int* __restrict__ parent) {
    const int64_t slots = n * cap;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (int64_t)gridDim.x * blockDim.x) {
        const int u = adj[s];
        if (u < 0 || u >= n || (mask && !mask[s])) continue;
        const int v = (int)(s / cap);
        if (u == v) continue;
        int a = cc_find(parent, v), b = cc_find(parent, u);
        while (a != b) {
            if (a < b) {
                const int seen = atomicCAS(parent + b, b, a);
                if (seen == b) break;
                b = seen;  // b was hooked elsewhere meanwhile: continue from its new parent
            } else {
                const int seen = atomicCAS(parent + a, a, b);
                if (seen == a) break;
                a = seen;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
cc_flatten_kernel(int* __restrict__ parent, int64_t n, int32_t* __restrict__ labels, unsigned long long* __restrict__ count) {
    unsigned roots = 0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)v;
        while (parent[r] != r) r = parent[r];
        labels[v] = r;
        roots += r == (int)v;
    }
    if (count) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) roots += __shfl_xor_sync(0xffffffffu, roots, o);
        if ((threadIdx.x & 31) == 0 && roots) atomicAdd(count, (unsigned long long)roots);
    }
}

}  // namespace

extern "C" int mmmp_connected_components(const int32_t* adj, const uint8_t* mask, int64_t n, int cap, int32_t* labels,
                                         int64_t* n_components, void* stream) {
    if (!adj || !labels || n < 0 || cap < 1) return MMMP_ERR_ARG;
    if (n >= (1ll << 31)) return MMMP_ERR_LIMIT;
    if (n == 0) {
        if (n_components) MMMP_CUDA_TRY(cudaMemsetAsync(n_components, 0, sizeof(int64_t), (cudaStream_t)stream));
        return MMMP_OK;
    }
    mmmp_pool_keep();
    cudaStream_t s = (cudaStream_t)stream;
    int* parent = nullptr;
    MMMP_CUDA_TRY(cudaMallocAsync(&parent, sizeof(int) * n, s));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t cap_grid = (int64_t)sms * 16;
    auto grid_of = [&](int64_t items) {
        const int64_t g = (items + 255) / 256;
        return (int)(g > cap_grid ? cap_grid : (g < 1 ? 1 : g));
    };
    if (n_components) cudaMemsetAsync(n_components, 0, sizeof(int64_t), s);
    MMMP_LAUNCH(cc_init_kernel, grid_of(n), 256, 0, stream, parent, n);
    MMMP_LAUNCH(cc_hook_kernel, grid_of(n * cap), 256, 0, stream, adj, mask, n, cap, parent);
    MMMP_LAUNCH(cc_flatten_kernel, grid_of(n), 256, 0, stream, parent, n, labels,
                reinterpret_cast<unsigned long long*>(n_components));
    cudaError_t e = cudaPeekAtLastError();
    cudaFreeAsync(parent, s);
    return e == cudaSuccess ? MMMP_OK : MMMP_ERR_CUDA_BASE + (int)e;
}
