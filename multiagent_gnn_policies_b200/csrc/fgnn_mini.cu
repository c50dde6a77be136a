// fgnn_mini.cu -- small flocks (B*N <= 128 agents, the reference's own cfgs): the whole closed-loop step inside ONE CTA.
// Compiled once per (FGNN_K, FGNN_HP) pair, HP <= 64: -DFGNN_K=<1..4> -DFGNN_HP=<16|32|64>.
//
// At N = 100 the general step is eight launches of one block each: ~28 us of launch latency for a few microseconds of
// work.  Here one CTA of 128 threads (= one readout tile = one adjacency block) runs T steps back to back -- hops, tensor-core
// readout + integrator, cell sort, adjacency + features -- with __syncthreads() where the general path has kernel
// boundaries.  The stages ARE the general path's device code (hop_body, final_tc_tiles, adjacency_body, finalize_reward),
// compiled with coherent loads (FGNN_COHERENT_LOADS: a stage reads what an earlier stage of the same kernel wrote), and the
// cell sort is restated for <= 128 agents (rank by counting in shared memory, same canonical (cell, id) order), so the
// results are bit-identical to the general path and the two can be mixed freely (tests/test_gpu_mini.py).
#define FGNN_COHERENT_LOADS
#define FGNN_MINI_TU
#include "fgnn_final_tc.cuh"

#define FGNN_CAT2(a, b, c, d) a##b##c##d
#define FGNN_CAT(a, b, c, d) FGNN_CAT2(a, b, c, d)

namespace fgnn {

#if FGNN_HP <= 64
constexpr int MINI_THREADS = 128;
static_assert(MINI_THREADS == FINAL_THREADS && MINI_THREADS == ADJ_THREADS, "one readout tile = one adjacency block");
static_assert(MINI_MAX_CELLS >= 9 * 128, "any episode split of <= 128 agents on the smallest (3 x 3) grids");

// bin -> scan -> scatter -> canon of the general path for M <= 128 agents (cell_of / cell_count come from the integrator's
// binning): canonical slot = rank of (cell, id); also what the small kernels of the general path do on the side (k_scan:
// advance t, reset the edge cursor; k_scatter: finalize the reward; k_canon: counters back to zero).
static __device__ __forceinline__ void mini_sort(const Params& p, int advance, int* s_cell) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        const int tn = *p.t + (advance ? 1 : 0);
        *p.t = tn;
        p.nnz_cursor[slot_of(tn, p.K)] = 0;
        p.edge_total[slot_of(tn, p.K)] = 0ull;
    }
    finalize_reward(p);
    const int c = tid < p.M ? p.cell_of[tid] : 0x7fffffff;
    s_cell[tid] = c;
    __syncthreads();
    if (tid < p.M) {
        int rank = 0;
        for (int j = 0; j < p.M; ++j) {
            const int cj = s_cell[j];
            rank += (cj < c || (cj == c && j < tid)) ? 1 : 0;
        }
        p.sorted_id[rank] = tid;
        p.sorted_cell[rank] = c;
        p.sorted_state[rank] = p.state[tid];
    }
    for (int cc = tid; cc <= p.C; cc += MINI_THREADS) {
        int n = 0;
        for (int j = 0; j < p.M; ++j) n += s_cell[j] < cc ? 1 : 0;
        p.cell_start[cc] = n;
        p.cell_count[cc] = 0;
    }
    __syncthreads();
}

template <int K>
static __device__ __forceinline__ void mini_hops(const Params& p) {
    const int tid = threadIdx.x;
    if (K == 3) {
        hop_body<2, true>(p, 0, tid);
        __syncthreads();
    } else if (K == 4) {
        hop_body<3, true>(p, 0, tid);
        __syncthreads();
        hop_body<2, false>(p, 1, tid);
        __syncthreads();
    }
    if (K == 2) hop_body<1, true>(p, 0, tid);
    else if (K >= 3) hop_body<1, false>(p, K - 2, tid);
    __syncthreads();
}

// T closed-loop steps (fgnn_rollout / fgnn_step)
template <int K, int HP>
__global__ void __launch_bounds__(MINI_THREADS, 1) k_mini_rollout(Params p, const uint8_t* __restrict__ tcw, int T, int stage_cap,
                                                                  int adj_smem_offset) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int s_cell[MINI_THREADS];
    // the flock's cell-sorted arrays live in shared memory for the whole launch (the adjacency stage scans ~14 candidates per
    // agent one dependent load after the other: 6 us of L2 round trips per step from global memory); written back at the end
    __shared__ __align__(32) double4 s_sorted_state[MINI_THREADS];
    __shared__ int s_sorted_id[MINI_THREADS], s_sorted_cell[MINI_THREADS], s_cell_of[MINI_THREADS];
    __shared__ int s_cell_start[MINI_MAX_CELLS + 1], s_cell_count[MINI_MAX_CELLS + 1];
    const Params pg = p;               // (global arrays)
    for (int i = threadIdx.x; i <= pg.C; i += MINI_THREADS) s_cell_count[i] = 0;
    p.sorted_state = s_sorted_state; p.sorted_id = s_sorted_id; p.sorted_cell = s_sorted_cell; p.cell_of = s_cell_of;
    p.cell_start = s_cell_start; p.cell_count = s_cell_count;
    TcCtx c;
    final_tc_setup<K, HP>(p, tcw, smem_raw, c);
    unsigned char* s_adj = smem_raw + adj_smem_offset;
    Params pf = p;
    pf.last_hop_done = K >= 2 ? 1 : 0;
    pf.write_z_last = 0;
    pf.tile_lo = 0; pf.tile_hi = 0;
    pf.fuse = nullptr;
    long long ck[5] = {0, 0, 0, 0, 0};
    const bool clk = p.mini_clock != nullptr && threadIdx.x == 0;
    for (int step = 0; step < T; ++step) {
        if (clk) ck[0] = clock64();
        mini_hops<K>(p);
        if (clk) ck[1] = clock64();
        final_tc_tiles<K, HP, true>(pf, c);
        __syncthreads();
        if (clk) ck[2] = clock64();
        mini_sort(p, 1, s_cell);
        if (clk) ck[3] = clock64();
        adjacency_body<false>(p, stage_cap, s_adj);
        __syncthreads();
        if (clk) {
            ck[4] = clock64();
            for (int q = 0; q < 4; ++q) p.mini_clock[q] += ck[q + 1] - ck[q];
        }
    }
    for (int i = threadIdx.x; i < pg.M; i += MINI_THREADS) {
        pg.sorted_state[i] = s_sorted_state[i];
        pg.sorted_id[i] = s_sorted_id[i];
        pg.sorted_cell[i] = s_sorted_cell[i];
        pg.cell_of[i] = s_cell_of[i];
    }
    for (int i = threadIdx.x; i <= pg.C; i += MINI_THREADS) pg.cell_start[i] = s_cell_start[i];
    final_tc_teardown<HP>(c);
}

// select_action alone (fgnn_policy): hops + open readout, z_{K-1} kept for fgnn_get_aggregated
template <int K, int HP>
__global__ void __launch_bounds__(MINI_THREADS, 1) k_mini_policy(Params p, const uint8_t* __restrict__ tcw) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TcCtx c;
    final_tc_setup<K, HP>(p, tcw, smem_raw, c);
    mini_hops<K>(p);
    Params pf = p;
    pf.last_hop_done = K >= 2 ? 1 : 0;
    pf.write_z_last = 1;
    pf.tile_lo = 0; pf.tile_hi = 0;
    pf.fuse = nullptr;
    final_tc_tiles<K, HP, false>(pf, c);
    final_tc_teardown<HP>(c);
}

// env.step(u) alone (fgnn_env_step): integrator + binning, cell sort, adjacency + features
template <typename T2>
static __device__ __forceinline__ void mini_envstep_body(const Params& p, const T2* __restrict__ u, int advance, int stage_cap,
                                                         unsigned char* s_adj, int* s_cell) {
    const int i = threadIdx.x;
    double racc[4] = {0, 0, 0, 0};
    if (i < owned_count(p)) {
        const int a = owned_agent(p, i);
        if (a >= 0) {
            const T2 uu = u[i];
            integrate_and_bin(p, a, ldg256(&p.state[a]), (double)uu.x, (double)uu.y, racc);
        }
    }
    reward_block_flush<MINI_THREADS>(p, racc);
    __syncthreads();
    mini_sort(p, advance, s_cell);
    adjacency_body<false>(p, stage_cap, s_adj);
}

template <int K>        // (K only keeps the symbol unique per object file)
__global__ void __launch_bounds__(MINI_THREADS, 1) k_mini_envstep(Params p, const void* __restrict__ u, int f64, int advance, int stage_cap) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ int s_cell[MINI_THREADS];
    if (f64) mini_envstep_body(p, reinterpret_cast<const double2*>(u), advance, stage_cap, smem_raw, s_cell);
    else mini_envstep_body(p, reinterpret_cast<const float2*>(u), advance, stage_cap, smem_raw, s_cell);
}
#endif

typedef void (*mini_rollout_kernel_t)(Params, const uint8_t*, int, int, int);
mini_rollout_kernel_t FGNN_CAT(get_mini_rollout_k, FGNN_K, _hp, FGNN_HP)() {
#if FGNN_HP <= 64
    return k_mini_rollout<FGNN_K, FGNN_HP>;
#else
    return nullptr;
#endif
}
typedef void (*mini_policy_kernel_t)(Params, const uint8_t*);
mini_policy_kernel_t FGNN_CAT(get_mini_policy_k, FGNN_K, _hp, FGNN_HP)() {
#if FGNN_HP <= 64
    return k_mini_policy<FGNN_K, FGNN_HP>;
#else
    return nullptr;
#endif
}
typedef void (*mini_envstep_kernel_t)(Params, const void*, int, int, int);
mini_envstep_kernel_t FGNN_CAT(get_mini_envstep_k, FGNN_K, _hp, FGNN_HP)() {
#if FGNN_HP <= 64
    return k_mini_envstep<FGNN_K * 1000 + FGNN_HP>;
#else
    return nullptr;
#endif
}
}  // namespace fgnn
