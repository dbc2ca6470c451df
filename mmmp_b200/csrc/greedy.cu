// Greedy degree-capped accept pass of the degree PRM on the GPU -- the step that turns the
// candidate tables (k1 nearest candidates per node + edge verdicts) into the roadmap's
// adjacency.
//
// Replaces (reference): the second loop of DecoupledPRM.connect_nearest_neighbors_degree,
// planners/prm/pybullet/decoupled/decoupled_prm.py:496-503 -- nodes in id order, candidates in
// ascending (dist, id) order, an edge is inserted at both ends when it is collision free, not
// yet present and both ends hold fewer than k2 edges.  The loop is sequential and order
// dependent: whether candidate (i, c) is accepted depends on every earlier decision that
// touched node i or its partner.  mmmp_greedy_degree_connect is the host replay of that loop;
// this file computes the SAME adjacency (same edges, same order inside every list) in a few
// data-parallel rounds.
//
// Events.  Candidate slot s = i * k1 + c is the event "node i considers nbr[s]", processed at
// time s.  Slots that are not free, point nowhere, or repeat an unordered pair that an EARLIER
// free slot already names (the reverse slot of a smaller node, or an earlier slot of the same
// row) never change the state: if the earlier slot was accepted the edge is "already
// present"; if it was rejected, a degree cap had been reached, and degrees never shrink.  They
// are rejected up front.  Every remaining event is accepted iff deg < k2 at both ends at its
// time.
//
// Bounds instead of order.  Every node walks the events that touch it in time order with two
// counters: L = events already known accepted, U = L + events still undecided.  An undecided
// event is SAFE at this end if U < k2 (the cap cannot bind whatever the undecided ones turn
// out to be) and DEAD if L >= k2.  DEAD at either end rejects it, SAFE at both ends accepts
// it.  Decisions only ever tighten the bounds, so the flags are monotone and the result does
// not depend on the order in which threads run; the earliest undecided event always has exact
// bounds, so every round decides something.  On 1 M nodes x 10 candidates the rounds shrink
// geometrically (a dozen rounds; the naive "earliest event of both ends" schedule needs ~300).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

enum : unsigned char { EV_UNDECIDED = 0, EV_ACCEPTED = 1, EV_REJECTED = 2 };
enum : unsigned char { SIDE_UNKNOWN = 0, SIDE_SAFE = 1, SIDE_DEAD = 2 };

struct GreedyArgs {
    const int32_t* nbr;      // [N, k1]
    const uint8_t* free_;    // [N, k1]
    int64_t N;
    int k1, k2, cap;
    unsigned char* state;    // [N * k1]
    unsigned char* side_i;   // [N * k1] flag computed at the owner end
    unsigned char* side_j;   // [N * k1] flag computed at the partner end
    int32_t* in_cnt;         // [N] incoming events per node
    const int64_t* in_off;   // [N + 1]
    int32_t* in_cursor;      // [N]
    int32_t* in_ev;          // [sum in_cnt] slots of the incoming events, time order per node
    unsigned char* done;     // [N]
    unsigned long long* changed;  // [3] change counters, used in rotation (round r adds to changed[r % 3])
    int32_t* rounds_out;     // device scalar
    int32_t* adj;            // [N, cap]
    int32_t* deg;            // [N]
};

// is slot (i, c) -> j a live event?  (see the header comment)
__device__ __forceinline__ bool event_live(const GreedyArgs& A, int64_t i, int c) {
    const int64_t s = i * A.k1 + c;
    const int j = __ldg(A.nbr + s);
    if (j < 0 || j >= A.N || j == i || !__ldg(A.free_ + s)) return false;
    for (int c2 = 0; c2 < c; ++c2)  // an earlier free slot of the same row names the same partner
        if (__ldg(A.nbr + i * A.k1 + c2) == j && __ldg(A.free_ + i * A.k1 + c2)) return false;
    if (j < i) {  // the partner's row is processed first: a free slot j -> i there decides the pair
        const int64_t b = (int64_t)j * A.k1;
        for (int c2 = 0; c2 < A.k1; ++c2)
            if (__ldg(A.nbr + b + c2) == (int)i && __ldg(A.free_ + b + c2)) return false;
    }
    return true;
}

__global__ void __launch_bounds__(256) greedy_init_kernel(GreedyArgs A) {
    const int64_t slots = A.N * A.k1;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = s / A.k1;
        const bool live = event_live(A, i, (int)(s - i * A.k1));
        A.state[s] = live ? EV_UNDECIDED : EV_REJECTED;
        A.side_i[s] = SIDE_UNKNOWN;
        A.side_j[s] = SIDE_UNKNOWN;
        if (live) atomicAdd(A.in_cnt + A.nbr[s], 1);
    }
}

__global__ void __launch_bounds__(256) greedy_fill_kernel(GreedyArgs A) {
    const int64_t slots = A.N * A.k1;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (int64_t)gridDim.x * blockDim.x) {
        if (A.state[s] != EV_UNDECIDED) continue;
        const int j = A.nbr[s];
        A.in_ev[A.in_off[j] + atomicAdd(A.in_cursor + j, 1)] = (int32_t)s;
    }
}

// incoming events of every node into time order (the atomic cursor filled them in any order)
__global__ void __launch_bounds__(256) greedy_sort_kernel(GreedyArgs A) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < A.N; v += (int64_t)gridDim.x * blockDim.x) {
        int32_t* e = A.in_ev + A.in_off[v];
        const int m = (int)(A.in_off[v + 1] - A.in_off[v]);
        for (int a = 1; a < m; ++a) {
            const int32_t x = e[a];
            int b = a - 1;
            while (b >= 0 && e[b] > x) {
                e[b + 1] = e[b];
                --b;
            }
            e[b + 1] = x;
        }
    }
}

// One visit of the events of node v in time order.  F(slot, is_owner_end) is called for every
// event that is not known rejected up front (own slots are visited whatever their state).
template <class F>
__device__ __forceinline__ void for_events_of(const GreedyArgs& A, int64_t v, F&& f) {
    const int32_t* in = A.in_ev + A.in_off[v];
    const int m = (int)(A.in_off[v + 1] - A.in_off[v]);
    const int64_t own0 = v * A.k1;
    int a = 0;
    for (; a < m && in[a] < own0; ++a)
        if (!f((int64_t)in[a], false)) return;
    for (int c = 0; c < A.k1; ++c)
        if (!f(own0 + c, true)) return;
    for (; a < m; ++a)
        if (!f((int64_t)in[a], false)) return;
}

__global__ void __launch_bounds__(256) greedy_rounds_kernel(GreedyArgs A) {
    cg::grid_group grid = cg::this_grid();
    const int k2 = A.k2;
    // Every round decides at least the earliest undecided event, so the loop ends; a chain of events
    // that each wait for the previous one (a path graph with k2 = 1) takes as many rounds as it is long.
    int round = 0;
    for (;; ++round) {
        unsigned changed = 0;
        for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < A.N; v += (int64_t)gridDim.x * blockDim.x) {
            if (A.done[v]) continue;
            int L = 0, U = 0;
            bool open = false;
            for_events_of(A, v, [&](int64_t s, bool owner) {
                const unsigned char st = __ldcg(A.state + s);
                if (st == EV_REJECTED) return true;
                if (st == EV_ACCEPTED) {
                    ++L;
                    ++U;
                    return true;
                }
                unsigned char* mine = (owner ? A.side_i : A.side_j) + s;
                const unsigned char* theirs = (owner ? A.side_j : A.side_i) + s;
                const unsigned char my = U < k2 ? SIDE_SAFE : (L >= k2 ? SIDE_DEAD : SIDE_UNKNOWN);
                if (my != SIDE_UNKNOWN && __ldcg(mine) != my) {
                    __stcg(mine, my);
                    ++changed;
                }
                const unsigned char other = __ldcg(theirs);
                if (my == SIDE_DEAD || other == SIDE_DEAD) {
                    __stcg(A.state + s, (unsigned char)EV_REJECTED);
                    ++changed;
                } else if (my == SIDE_SAFE && other == SIDE_SAFE) {
                    __stcg(A.state + s, (unsigned char)EV_ACCEPTED);
                    ++changed;
                    ++L;
                    ++U;
                } else {
                    ++U;
                    open = true;
                }
                return true;
            });
            if (!open) A.done[v] = 1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
        if ((threadIdx.x & 31) == 0 && changed) atomicAdd(A.changed + round % 3, (unsigned long long)changed);
        grid.sync();
        if (__ldcg(A.changed + round % 3) == 0ull) break;
        // the counter of round r + 2 was last read after the barrier of round r - 1 and is next written
        // after the barrier of round r + 1: safe to clear between this barrier and the next
        if (blockIdx.x == 0 && threadIdx.x == 0) __stcg(A.changed + (round + 2) % 3, 0ull);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && A.rounds_out) *A.rounds_out = round + 1;
}

__global__ void __launch_bounds__(256) greedy_emit_kernel(GreedyArgs A) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < A.N; v += (int64_t)gridDim.x * blockDim.x) {
        int n = 0;
        int32_t* row = A.adj + v * A.cap;
        for_events_of(A, v, [&](int64_t s, bool owner) {
            if (A.state[s] == EV_ACCEPTED && n < A.cap) row[n++] = owner ? A.nbr[s] : (int32_t)(s / A.k1);
            return true;
        });
        A.deg[v] = n;
        for (; n < A.cap; ++n) row[n] = -1;
    }
}

}  // namespace

extern "C" int mmmp_exclusive_scan_i32(const int32_t* in_d, int64_t n, int64_t* out_d, void* stream);

extern "C" int mmmp_greedy_degree_connect_device(const int32_t* nbr_idx_d, const uint8_t* edge_free_d, int64_t N,
                                                 int k1, int k2, int cap, int32_t* adj_d, int32_t* deg_d,
                                                 int32_t* rounds_d, void* stream) {
    if (!nbr_idx_d || !edge_free_d || !adj_d || !deg_d || N < 0 || k1 < 1 || k2 < 1 || cap < k2) return MMMP_ERR_ARG;
    if (N * (int64_t)k1 >= (1ll << 31)) return MMMP_ERR_LIMIT;
    if (N == 0) return MMMP_OK;
    mmmp_pool_keep();
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 148, coop = 0;
    MMMP_CUDA_TRY(cudaGetDevice(&dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MMMP_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return MMMP_ERR_LIMIT;
    const int64_t slots = N * k1;
    constexpr int N_COUNTERS = 3;   // change counters of the rounds, used in rotation
    // one allocation: state | side_i | side_j | done | pad | in_cnt | in_cursor | changed | in_off | in_ev
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_state = 0, o_si = up(o_state + slots), o_sj = up(o_si + slots), o_done = up(o_sj + slots);
    const size_t o_cnt = up(o_done + N), o_cur = up(o_cnt + 4 * N), o_chg = up(o_cur + 4 * N);
    const size_t o_off = up(o_chg + 8 * N_COUNTERS), o_ev = up(o_off + 8 * (N + 1)), total = up(o_ev + 4 * slots);
    unsigned char* buf = nullptr;
    MMMP_CUDA_TRY(cudaMallocAsync(&buf, total, s));
    cudaMemsetAsync(buf + o_done, 0, o_off - o_done, s);   // done, in_cnt, in_cursor, changed
    GreedyArgs A;
    A.nbr = nbr_idx_d;
    A.free_ = edge_free_d;
    A.N = N;
    A.k1 = k1;
    A.k2 = k2;
    A.cap = cap;
    A.state = buf + o_state;
    A.side_i = buf + o_si;
    A.side_j = buf + o_sj;
    A.done = buf + o_done;
    A.in_cnt = reinterpret_cast<int32_t*>(buf + o_cnt);
    A.in_cursor = reinterpret_cast<int32_t*>(buf + o_cur);
    A.changed = reinterpret_cast<unsigned long long*>(buf + o_chg);
    A.in_off = reinterpret_cast<const int64_t*>(buf + o_off);
    A.in_ev = reinterpret_cast<int32_t*>(buf + o_ev);
    A.rounds_out = rounds_d;
    A.adj = adj_d;
    A.deg = deg_d;
    auto grid_of = [&](int64_t items) {
        const int64_t g = (items + 255) / 256, capg = (int64_t)sms * 16;
        return (int)(g > capg ? capg : (g < 1 ? 1 : g));
    };
    int rc = MMMP_OK;
    MMMP_LAUNCH(greedy_init_kernel, grid_of(slots), 256, 0, stream, A);
    rc = mmmp_exclusive_scan_i32(A.in_cnt, N, const_cast<int64_t*>(A.in_off), stream);
    if (rc == MMMP_OK) {
        MMMP_LAUNCH(greedy_fill_kernel, grid_of(slots), 256, 0, stream, A);
        MMMP_LAUNCH(greedy_sort_kernel, grid_of(N), 256, 0, stream, A);
        int per_sm = 1;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_rounds_kernel, 256, 0);
        if (e == cudaSuccess) {
            int64_t g = (int64_t)sms * (per_sm < 1 ? 1 : per_sm);
            const int64_t need = (N + 255) / 256;
            if (g > need) g = need < 1 ? 1 : need;
            void* params[] = {&A};
            e = cudaLaunchCooperativeKernel((const void*)greedy_rounds_kernel, dim3((unsigned)g), dim3(256), params, 0, s);
            g_mmmp_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if (e == cudaSuccess) {
            MMMP_LAUNCH(greedy_emit_kernel, grid_of(N), 256, 0, stream, A);
            e = cudaPeekAtLastError();
        }
        if (e != cudaSuccess) rc = MMMP_ERR_CUDA_BASE + (int)e;
    }
    cudaFreeAsync(buf, s);
    return rc;
}
