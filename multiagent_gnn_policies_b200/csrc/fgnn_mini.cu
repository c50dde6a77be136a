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
        stg256(&p.sorted_state[rank], ldg256(&p.state[tid]));
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
    TcCtx c;
    final_tc_setup<K, HP>(p, tcw, smem_raw, c);
    unsigned char* s_adj = smem_raw + adj_smem_offset;
    Params pf = p;
    pf.last_hop_done = K >= 2 ? 1 : 0;
    pf.write_z_last = 0;
    pf.tile_lo = 0; pf.tile_hi = 0;
    pf.fuse = nullptr;
    for (int step = 0; step < T; ++step) {
        mini_hops<K>(p);
        final_tc_tiles<K, HP, true>(pf, c);
        __syncthreads();
        mini_sort(p, 1, s_cell);
        adjacency_body<false>(p, stage_cap, s_adj);
        __syncthreads();
    }
    final_tc_teardown<HP>(c);
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
}  // namespace fgnn
