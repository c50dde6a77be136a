// fgnn_kernels.cuh -- device code of the flocking-GNN rollout engine (sm_100a).
//
// One rollout step =  [bin] -> scan_sums -> scan -> scatter -> canon -> adjacency+features -> hop(s) [-> last hop] -> final
// (final = [last hop +] readout MLP + double integrator + binning of the new positions).
// See DESIGN.md for the data layout and the per-kernel byte counts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// FGNN_COHERENT_LOADS (the single-CTA small-flock kernels, fgnn_mini.cu): every stage of a step runs inside ONE kernel, so
// data written by an earlier stage of the SAME kernel is read back -- the non-coherent path (ld.global.nc / __ldg) is only
// valid for data written by an earlier kernel.  With the switch every read-only load of the shared device code becomes a
// plain load (coherent within the SM, and the whole kernel is one CTA).
// The 256-bit record accesses become generic there too: the single-CTA kernel keeps the cell-sorted arrays of its flock in
// SHARED memory and hands the shared device code pointers to them.
#ifdef FGNN_COHERENT_LOADS
#define __ldg(ptr) (*(ptr))
#endif

namespace fgnn {

constexpr int F = 6;          // features per agent (n_states)
constexpr int ROW = 8;        // padded feature row: 8 floats = one 32-byte sector
constexpr int SINV_PAD = 7;   // x_{t-1}[m][7] holds the source scale of graph t, 1/max(deg_t(m),1) (written by k_adjacency at step t)
constexpr int KMAX = 4;       // filter taps supported
constexpr int LMAX = 4;       // hidden layers supported
constexpr int FINAL_THREADS = 128;   // block size of the fused final kernel and the dense Actor kernel
constexpr int DENSE_MT = 32;         // m-tile of the dense Actor kernel
constexpr int RSLOTS = 64;          // reward accumulators per episode (spreads same-address atomics)
// ... of which a handle with many episodes uses fewer: an episode of N agents feeds its sums from ~N/128 blocks, and the
// finalisation walks every slot of every episode (with 64 slots and 256 episodes that walk was 38 us on the critical path
// of k_scatter -- config C3)
__host__ __device__ inline int rslots_of(int B) { return B >= 64 ? 4 : B >= 8 ? 16 : RSLOTS; }
constexpr int ELLW = 8;             // neighbours kept inline per agent: one dependent load less per gather
constexpr int MINI_MAX_CELLS = 1280;        // single-CTA path (fgnn_mini.cu): cells of the flock's grid it keeps in shared memory
#ifndef FGNN_HOP_UNROLL
#define FGNN_HOP_UNROLL 2
#endif
constexpr int HOP_UNROLL = FGNN_HOP_UNROLL;        // edges gathered concurrently per thread in the hop loops (measured: 2 beats 1 and 4 -- the gathers
                                                   // are latency-bound and more registers cost more in occupancy than the extra loads in flight gain)

// All device pointers one step needs.  Passed by value to every kernel.
struct Params {
    int M;                    // total agents = B * N
    int N;                    // agents per episode
    int B;                    // episodes
    int K;                    // taps
    int L;                    // hidden layers
    int G;                    // cells along x of the wrapped cell grid
    int Gy;                   // cells along y
    int C;                    // total cells = B*G*Gy
    // sharding (one rank of a multi-GPU flock).  Single GPU: own == null, the rank owns agents [a_lo, a_lo + n_own) = all.
    // Sharded: the OWNED set is a device list (it changes when agents are handed over between ranks), the
    // agents present on the rank (pool) = owned list followed by the ghost list.
    int a_lo;                 // single GPU: first owned agent (0)
    int n_own;                // single GPU: number of owned agents (M)
    int pool_cap;             // capacity of the owned / ghost lists
    const int* own;           // [pool_cap] owned agents (null = contiguous range)
    const int* n_own_d;       // device count of owned agents
    const int* ghost;         // [pool_cap] agents received from other ranks this step
    const int* n_ghost_d;     // device count of ghosts
    const struct ShardFuse* fuse;   // non-null: the closed final kernel also packs the halo records (fgnn_shard_step_begin)
    int* n_ghost_snap;        // ghost count as of the last graph build (written by the adjacency kernel): what the hop kernels walk,
                              //     so that the ghost counter itself may be reset while they run
    int mean_pooling;
    int half_accel;
    int sums_in_bin;          // every binning site also accumulates the per-scan-tile sums (no k_scan_sums launch)
    int write_z_last;         // final kernel also stores z_{K-1} (debug / fgnn_get_aggregated)
    int last_hop_done;        // z_{K-1} was already written by a separate hop launch: the final kernel reads it
    int tile_lo, tile_hi;     // final kernel: range of 128-agent tiles of this launch (tile_hi <= 0: all of them)
    unsigned nnz_cap;         // directed-edge capacity per ring slot
    double inv_cell;          // 1 / cell size  (cell size = R * (1 + 2^-20))
    double R2;                // comm_radius^2
    double dt;
    double gain;              // action_scalar

    int* t;                   // device step counter (index of the current graph)
    double4* state;           // [M] px,py,vx,vy
    int* cell_of;             // [M]
    int* cell_count;          // [C+1]  (last entry stays 0)
    int* cell_start;          // [C+1]
    int* tmp_id;              // [M] scatter order (atomic, not canonical)
    int* sorted_id;           // [M] canonical: by cell, then by agent id
    double4* sorted_state;    // [M]
    int* sorted_cell;         // [M] cell of every sorted slot (written by k_canon; k_pair_adjacency reads it instead of re-deriving it)
    unsigned* tile_status;    // scan look-back words
    int* tile_counter;        // scan dynamic tile id
    int n_tiles;

    float* xhist;             // [K][M][ROW]  features x_{t}, ring slot = t mod K
    float* sinv;              // [K][M]       source scale 1/max(deg,1) (or 1)
    unsigned* row_start;      // [K][M]
    int* deg;                 // [K][M]
    int* cols;                // [K][nnz_cap]
    int* ell;                 // [K][M][ELLW] first ELLW neighbours of every row, -1 padded (copy of the CSR head)
    unsigned* nnz_cursor;     // [K]
    unsigned long long* edge_total;   // [K] directed edges of every ring slot (pair kernels: the CSR cursor counts long rows only)
    int* overflow;            // sticky flag

    float* zbuf;              // [K][M][ROW]  z_k, k >= 1
    float* ybuf;              // [2][K][M][ROW] hop intermediates (ping-pong)
    float* action;            // [M][2]
    const float* weights;     // packed, see WeightLayout
    const unsigned char* amask;   // [M] or null; 0 = leader: the integrator ignores its action (FlockingLeader-v0)

    double* racc;             // [RSLOTS][B][4] sum vx, vy, vx^2, vy^2 (slotted atomics, B > 1)
    double* racc_part;        // [max grid][4] per-block partial sums (B == 1: no atomics, fixed order)
    int* n_partials;          // number of valid rows in racc_part
    double* reward;           // [B]
    int* reward_pending;
    double* reward_log;       // [T][B] or null
    int* log_index;
    long long* mini_clock;    // [4] or null: cycles the single-CTA kernel spent in hops / readout + integrator / cell sort / adjacency
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may become
// resident while its predecessor drains.  Every such kernel first lets ITS successor do the same, then waits until the
// predecessor grid has completed and its memory operations are visible.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_prologue() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__host__ __device__ inline int slot_of(int t, int K) { int s = t % K; return s < 0 ? s + K : s; }

// Packed weights (floats): w0[6K][HP] b0[HP] | (L-1) x { wh[HP][HP] bh[HP] } | wl[HP][2] bl[2] (+pad)
struct WeightLayout {
    int in0, HP, L;
    __host__ __device__ int off_w0() const { return 0; }
    __host__ __device__ int off_b0() const { return in0 * HP; }
    __host__ __device__ int off_wh(int l) const { return in0 * HP + HP + (l - 1) * (HP * HP + HP); }   // l = 1..L-1
    __host__ __device__ int off_bh(int l) const { return off_wh(l) + HP * HP; }
    __host__ __device__ int off_wl() const { return in0 * HP + HP + (L - 1) * (HP * HP + HP); }
    __host__ __device__ int off_bl() const { return off_wl() + HP * 2; }
    __host__ __device__ int total() const { return (off_bl() + 2 + 3) & ~3; }
};

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int wrap(long long v, int G) {
    if (v == (long long)(int)v) {                 // 32-bit remainder: ~5x fewer instructions than the emulated 64-bit one
        const int r = (int)v % G;
        return r < 0 ? r + G : r;
    }
    const int r = (int)(v % G);
    return r < 0 ? r + G : r;
}

// episode of agent a (B == 1: no division)
__device__ __forceinline__ int episode_of(const Params& p, int a) { return p.B == 1 ? 0 : a / p.N; }

__device__ __forceinline__ void cell_coords(const Params& p, double px, double py, long long& ix, long long& iy) {
    ix = (long long)floor(px * p.inv_cell);
    iy = (long long)floor(py * p.inv_cell);
}

__device__ __forceinline__ int cell_index(const Params& p, int ep, long long ix, long long iy) {
    return (ep * p.Gy + wrap(iy, p.Gy)) * p.G + wrap(ix, p.G);
}

// One more agent in cell c: per-cell counter, and (sums_in_bin) the population of the cell's scan tile, which k_scan reads
// instead of a k_scan_sums pass.  The lanes of a warp mostly fall into the same tile: one atomic per distinct tile.
constexpr int SCAN_TILE_LOG2 = 11;
__device__ __forceinline__ int bin_agent(const Params& p, int c) {
    const int rank = atomicAdd(&p.cell_count[c], 1);
    if (p.sums_in_bin) {
        const unsigned ts = (unsigned)c >> SCAN_TILE_LOG2;
        const unsigned peers = __match_any_sync(__activemask(), ts);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&p.tile_status[ts], (unsigned)__popc(peers));
    }
    return rank;
}

// number of agents present on this rank and the i-th of them
// Sharded: the owned list keeps its order; an agent that is handed over leaves a tombstone (-1) and new agents
// are appended, so the kernels keep walking spatially coherent runs.  owned_count = high-water mark,
// owned_agent / pool_agent may return -1 (skip).
__device__ __forceinline__ int owned_count(const Params& p) { return p.own ? *p.n_own_d : p.n_own; }
__device__ __forceinline__ int owned_agent(const Params& p, int i) { return p.own ? p.own[i] : p.a_lo + i; }
// number of agents in the cell-sorted arrays (valid after k_scan): the last scan entry
__device__ __forceinline__ int sorted_count(const Params& p) { return p.own ? p.cell_start[p.C] : p.M; }
__device__ __forceinline__ int pool_size(const Params& p) { return p.own ? *p.n_own_d + *p.n_ghost_d : p.M; }
__device__ __forceinline__ int pool_agent(const Params& p, int i) {
    if (!p.own) return i;
    const int no = *p.n_own_d;
    return i < no ? p.own[i] : p.ghost[i - no];
}
// the same pool as the hop kernels see it: ghost count from the snapshot taken when the graph was built
__device__ __forceinline__ int hop_pool_size(const Params& p) { return p.own ? *p.n_own_d + *p.n_ghost_snap : p.M; }

// r2 exactly as numpy evaluates dx*dx + dy*dy (two roundings of the products, one of the sum; no FMA)
__device__ __forceinline__ double r2_exact(double dx, double dy) {
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// ---- 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256) ----
// Every per-agent record here is one aligned 32-byte sector (feature row, ELL head, double4 state).  The gathers are
// bound by the L1TEX wavefront pipe, which charges per instruction AND per line touched: one 256-bit access per
// record instead of a 128-bit + 64-bit (or two 128-bit) pair halves the wavefronts of every scattered record read.
// `.nc` variants: data written by an EARLIER kernel only.
#ifndef FGNN_COHERENT_LOADS
__device__ __forceinline__ void ldg256_nc(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ldg256_nc(const int* p, int (&v)[8]) {
    asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ double4 ldg256_nc(const double4* p) {
    double4 v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double4 ldg256(const double4* p) {            // coherent: the kernel also writes this array
    double4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg256(double4* p, const double4 v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void stg256(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}
__device__ __forceinline__ void stg256(int* p, const int (&v)[8]) {
    asm volatile("st.global.v8.s32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
#else
// single-CTA kernels: generic, coherent accesses (the pointer may be into shared memory, where 256-bit accesses do not
// exist): two 128-bit halves
__device__ __forceinline__ void ldg256_nc(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ldg256_nc(const int* p, int (&v)[8]) {
    const int4 a = reinterpret_cast<const int4*>(p)[0], b = reinterpret_cast<const int4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ double4 ldg256_nc(const double4* p) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ double4 ldg256(const double4* p) { return ldg256_nc(p); }
__device__ __forceinline__ void stg256(double4* p, const double4 v) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2*>(p)[1] = make_double2(v.z, v.w);
}
__device__ __forceinline__ void stg256(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    reinterpret_cast<float4*>(p)[0] = make_float4(a, b, c, d);
    reinterpret_cast<float4*>(p)[1] = make_float4(e, f, g, h);
}
__device__ __forceinline__ void stg256(int* p, const int (&v)[8]) {
    reinterpret_cast<int4*>(p)[0] = make_int4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<int4*>(p)[1] = make_int4(v[4], v[5], v[6], v[7]);
}
#endif

__device__ __forceinline__ void load_row6(const float* __restrict__ base, int idx, float (&v)[F]) {
    float r[8];
    ldg256_nc(base + (size_t)idx * ROW, r);
#pragma unroll
    for (int f = 0; f < F; ++f) v[f] = r[f];
}

__device__ __forceinline__ void store_row6(float* base, int idx, const float (&v)[F]) {
    stg256(base + (size_t)idx * ROW, v[0], v[1], v[2], v[3], v[4], v[5], 0.f, 0.f);
}

#if defined(FGNN_MAIN_TU) || defined(FGNN_MINI_TU)
#ifdef FGNN_MAIN_TU
// ------------------------------------------------------------------------------------------
// K_A  bin: cell of every agent + per-cell population count
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bin(Params p) {      // owned agents (ghosts are binned by k_shard_unpack)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= owned_count(p)) return;
    const int a = owned_agent(p, i);
    if (a < 0) return;
    double4 s = p.state[a];
    long long ix, iy;
    cell_coords(p, s.x, s.y, ix, iy);
    int c = cell_index(p, episode_of(p, a), ix, iy);
    p.cell_of[a] = c;
    bin_agent(p, c);
}

#endif  // FGNN_MAIN_TU (k_bin)

// reward_b = -(var(vx) + var(vy)) per episode from the accumulated sums; clears the sums.
// B == 1: per-block partials summed in a fixed order (bit-reproducible); B > 1: RSLOTS atomic slots per episode.
static __device__ __forceinline__ void finalize_reward(const Params& p) {
    if (*p.reward_pending == 0) return;
    __syncthreads();
    const int li = p.reward_log ? *p.log_index : 0;
    const double n = (double)p.N;
    if (p.B == 1) {
        __shared__ double s_red[4][256];
        const int np = *p.n_partials;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            a0 += p.racc_part[i * 4 + 0]; a1 += p.racc_part[i * 4 + 1];
            a2 += p.racc_part[i * 4 + 2]; a3 += p.racc_part[i * 4 + 3];
        }
        s_red[0][threadIdx.x] = a0; s_red[1][threadIdx.x] = a1; s_red[2][threadIdx.x] = a2; s_red[3][threadIdx.x] = a3;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
#pragma unroll
                for (int k = 0; k < 4; ++k) s_red[k][threadIdx.x] += s_red[k][threadIdx.x + o];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double mx = s_red[0][0] / n, my = s_red[1][0] / n;
            const double r = -((s_red[2][0] / n - mx * mx) + (s_red[3][0] / n - my * my));
            p.reward[0] = r;
            if (p.reward_log) p.reward_log[(size_t)li] = r;
        }
    } else {
        for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
            double q[4] = {0, 0, 0, 0};
            const int nslots = rslots_of(p.B);
            // loads first (independent: several slots in flight), stores after -- interleaved they form one dependent chain
            // of L2 round trips per slot
#pragma unroll 8
            for (int sidx = 0; sidx < nslots; ++sidx) {
                const double* src = p.racc + ((size_t)sidx * p.B + b) * 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) q[k] += src[k];
            }
            for (int sidx = 0; sidx < nslots; ++sidx) {
                double* src = p.racc + ((size_t)sidx * p.B + b) * 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) src[k] = 0.0;
            }
            const double mx = q[0] / n, my = q[1] / n;
            const double r = -((q[2] / n - mx * mx) + (q[3] / n - my * my));
            p.reward[b] = r;
            if (p.reward_log) p.reward_log[(size_t)li * p.B + b] = r;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *p.reward_pending = 0;
        if (p.reward_log) *p.log_index = li + 1;
    }
}

#ifdef FGNN_MAIN_TU
__global__ void __launch_bounds__(256) k_finalize_reward(Params p) { finalize_reward(p); }

// ------------------------------------------------------------------------------------------
// K_B  scan: exclusive prefix sum of cell_count[0..C] -> cell_start[0..C] (single pass,
//      direct look-back), zeroes cell_count; tile 0 also advances t.  (The reward is finalised by k_scatter.)
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ROUNDS = 2;                                   // int4 vectors per thread
constexpr int SCAN_ITEMS = 4 * SCAN_ROUNDS;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
static_assert(SCAN_TILE == 1 << SCAN_TILE_LOG2, "bin_agent() maps a cell to its scan tile with a shift");
constexpr unsigned FLAG_AGG = 1u << 30, VAL_MASK = (1u << 30) - 1;

// Accesses are COALESCED int4 vectors (vector r of thread t sits at r*256 + t inside the tile): the first version
// gave every thread 16 consecutive counts, so each warp-wide load touched 32 different lines and the three passes
// over the array (load, zero, store) kept the L1TEX wavefront pipe of the few SMs that run this small grid busy for
// ~20 us (ncu: lsu wavefronts 22-27 % of peak for 12 MB of traffic, barrier/lg_throttle stalls).
// Two-pass mode, pass 1: the sum of every tile's counts into tile_status (plain ints, no flags).  Pass 2 (k_scan with
// two_pass = 1) then reads its predecessors' sums without waiting on anybody.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(Params p) {
    pdl_prologue();
    __shared__ int s_w[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.C + 1;
    int sum = 0;
#pragma unroll
    for (int r = 0; r < SCAN_ROUNDS; ++r) {
        const int e = blockIdx.x * SCAN_TILE + (r * SCAN_THREADS + tid) * 4;
        if (e + 3 < n) {
            const int4 v = *reinterpret_cast<const int4*>(p.cell_count + e);
            sum += v.x + v.y + v.z + v.w;
        } else {
            for (int j = 0; j < 3; ++j) sum += e + j < n ? p.cell_count[e + j] : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_w[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s_w[w];
        p.tile_status[blockIdx.x] = (unsigned)t;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(Params p, int advance, int static_ids, int two_pass) {
    pdl_prologue();
    __shared__ int s_tile;
    __shared__ int s_warp[SCAN_ROUNDS][SCAN_THREADS / 32];
    __shared__ int s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = SCAN_THREADS / 32;
    // Tile id: the block index when every tile is resident at once (the look-back below then cannot wait on a block
    // that has not started; saves the serialised same-address atomics), else handed out by an atomic counter.
    int tile = blockIdx.x;
    if (!static_ids) {
        if (tid == 0) s_tile = atomicAdd(p.tile_counter, 1);
        __syncthreads();
        tile = s_tile;
    }
    if (tile == 0 && tid == 0) {
        const int tn = *p.t + (advance ? 1 : 0);
        *p.t = tn;
        p.nnz_cursor[slot_of(tn, p.K)] = 0;          // edge cursor of the slot about to be rebuilt
        p.edge_total[slot_of(tn, p.K)] = 0ull;
    }
    const int n = p.C + 1;
    int4 v[SCAN_ROUNDS];
    int sum[SCAN_ROUNDS], inc[SCAN_ROUNDS];
#pragma unroll
    for (int r = 0; r < SCAN_ROUNDS; ++r) {
        const int e = tile * SCAN_TILE + (r * SCAN_THREADS + tid) * 4;      // first element of this vector
        if (e + 3 < n) {
            v[r] = *reinterpret_cast<const int4*>(p.cell_count + e);
            *reinterpret_cast<int4*>(p.cell_count + e) = make_int4(0, 0, 0, 0);
        } else {
            v[r].x = e < n ? p.cell_count[e] : 0;
            v[r].y = e + 1 < n ? p.cell_count[e + 1] : 0;
            v[r].z = e + 2 < n ? p.cell_count[e + 2] : 0;
            v[r].w = 0;                                                      // e + 3 >= n here
            if (e < n) p.cell_count[e] = 0;
            if (e + 1 < n) p.cell_count[e + 1] = 0;
            if (e + 2 < n) p.cell_count[e + 2] = 0;
        }
        sum[r] = v[r].x + v[r].y + v[r].z + v[r].w;
        inc[r] = sum[r];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, inc[r], o);
            if (lane >= o) inc[r] += y;
        }
        if (lane == 31) s_warp[r][warp] = inc[r];
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan over the (round, warp) totals in element order; lane l <-> round l / NW, warp l % NW
        static_assert(SCAN_ROUNDS * NW <= 32, "one warp scans the partial totals");
        int w = lane < SCAN_ROUNDS * NW ? s_warp[lane / NW][lane % NW] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += y;
        }
        if (lane < SCAN_ROUNDS * NW) s_warp[lane / NW][lane % NW] = winc - w;
        const int total = __shfl_sync(0xffffffffu, winc, SCAN_ROUNDS * NW - 1);
        int part = 0;
        if (two_pass) {                                   // predecessors' sums were written by k_scan_sums
            for (int idx = lane; idx < tile; idx += 32) part += (int)__ldg(&p.tile_status[idx]);
        } else {
            volatile unsigned* status = p.tile_status;
            if (lane == 0) status[tile] = FLAG_AGG | (unsigned)total;      // publish right away: nobody waits for a prefix
            // Look-back without a dependency chain: the tile count is small, so this warp fetches the aggregates of
            // ALL predecessor tiles directly -- one round trip to L2 for the whole prefix instead of one per 32 tiles.
            for (int idx = lane; idx < tile; idx += 32) {
                unsigned w32;
                while (((w32 = status[idx]) >> 30) == 0) __nanosleep(40);
                part += (int)(w32 & VAL_MASK);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_excl = part;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SCAN_ROUNDS; ++r) {
        const int e = tile * SCAN_TILE + (r * SCAN_THREADS + tid) * 4;
        const int run = s_excl + s_warp[r][warp] + (inc[r] - sum[r]);
        const int4 o = make_int4(run, run + v[r].x, run + v[r].x + v[r].y, run + v[r].x + v[r].y + v[r].z);
        if (e + 3 < n) {
            *reinterpret_cast<int4*>(p.cell_start + e) = o;
        } else {
            if (e < n) p.cell_start[e] = o.x;
            if (e + 1 < n) p.cell_start[e + 1] = o.y;
            if (e + 2 < n) p.cell_start[e + 2] = o.z;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K_C  scatter agent ids into their cell's slot range (atomic order, canonicalised by K_C2)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter(Params p) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = i < pool_size(p) ? pool_agent(p, i) : -1;
    if (a >= 0) {
        int c = p.cell_of[a];
        int slot = p.cell_start[c] + atomicAdd(&p.cell_count[c], 1);
        p.tmp_id[slot] = a;
    }
    // Reward of the step just integrated: a chain of dependent global loads (~5 us) that used to sit on the scan
    // kernel's critical path; here it hides behind the other blocks' scatter traffic.
    if (blockIdx.x == 0) finalize_reward(p);
}

// K_C2 canon: order every cell's list by agent id (rank by counting), copy the state next to it.
//      Makes CSR row order -- and with it every fp32 sum downstream -- run-to-run reproducible.
__global__ void __launch_bounds__(256) k_canon(Params p) {
    pdl_prologue();
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = s; i <= p.C; i += gridDim.x * blockDim.x) p.cell_count[i] = 0;   // fill counters -> 0 for the next bin
    if (s >= sorted_count(p)) return;
    int a = p.tmp_id[s];
    int c = p.cell_of[a];
    int q0 = p.cell_start[c], q1 = p.cell_start[c + 1];
    int rank = 0;
    for (int q = q0; q < q1; ++q) rank += (p.tmp_id[q] < a) ? 1 : 0;
    int dst = q0 + rank;
    p.sorted_id[dst] = a;
    p.sorted_cell[dst] = c;
    stg256(&p.sorted_state[dst], ldg256_nc(&p.state[a]));
}

#endif  // FGNN_MAIN_TU (scan / scatter / canon kernels)

// ------------------------------------------------------------------------------------------
// K_D  adjacency + degree + 6-d relative features (gym_flock compute_helpers), CSR/ELL emission.
//      One thread per agent, in cell-sorted order; float64 arithmetic for the radius cut and the
//      feature sums (bit-identical edge set to the float64 oracle).
//      Single pass over the 3x3 cell neighbourhood (the three cells of a grid row are one contiguous
//      slot range): features accumulate in the accept branch, accepted neighbour ids are staged in
//      shared memory.  (A filter-then-compute split was measured slower: more registers, lower occupancy.)
//      The warp reserves one contiguous run of CSR edge slots for its 32 rows; the first ELLW
//      neighbour ids are also stored inline per agent (ELL head).  Rows longer than the stage
//      re-scan their tail.
// ------------------------------------------------------------------------------------------
constexpr int ADJ_THREADS = 128;
constexpr int WS_CAP = 48;            // warp-staged variant: candidate slots per grid row and warp
constexpr int WS_STAGE = 12;          // ... and neighbour ids staged per thread (longer rows re-scan shared memory)
constexpr size_t WS_SMEM = (size_t)(ADJ_THREADS / 32) * 3 * WS_CAP * (sizeof(double4) + sizeof(int));

// 1/x in float64 from the hardware seed (rcp.approx.ftz.f64: 2^-23) and two Newton steps: within ~1 ulp of the
// IEEE quotient in 5 instructions instead of the ~30 (with a slow-path branch) of `1.0 / x`.  The feature sums
// are rounded to fp32 afterwards, so the last float64 bit is irrelevant to parity.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// One accepted pair: feature sums in float64
#define FGNN_ADJ_ACCEPT(o, dx, dy, r2)                                                                \
    {                                                                                                 \
        const double inv = fast_rcp(r2);                                                              \
        const double inv2 = inv * inv;                                                                \
        f0 += me.z - (o).z;                                                                           \
        f1 += (dx) * inv2;                                                                            \
        f2 += (dx) * inv;                                                                             \
        f3 += me.w - (o).w;                                                                           \
        f4 += (dy) * inv2;                                                                            \
        f5 += (dy) * inv;                                                                             \
    }

// WS = true: the 32 agents of a warp are consecutive in cell order, so (when the warp sits inside one grid row)
// the candidates of all its lanes are three contiguous slot ranges.  The warp copies them ONCE into shared memory
// with coalesced loads and every lane scans its own sub-range there: the per-lane dependent global loads (ncu:
// long_scoreboard 36 %, L1 wavefronts 38 % of peak) become shared-memory reads.  Warps that straddle a row, touch
// the grid seam or overflow the tile take the per-lane global path.  Neighbour order is identical in both paths.
template <bool WS>
__device__ __forceinline__ void adjacency_body(const Params& p, int stage_cap, unsigned char* s_adj_raw) {
    int* s_stage = reinterpret_cast<int*>(s_adj_raw);     // [stage_cap][ADJ_THREADS] accepted neighbour ids
    const int tid = threadIdx.x;
    const int s = blockIdx.x * ADJ_THREADS + tid;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // housekeeping for the next scan
    for (int i = s; i < p.n_tiles; i += gridDim.x * ADJ_THREADS) p.tile_status[i] = 0;
    if (s == 0) {
        *p.tile_counter = 0;
        if (p.own) *p.n_ghost_snap = *p.n_ghost_d;
    }

    const int t = *p.t;
    const int g = slot_of(t, p.K);
    const bool valid = s < sorted_count(p);
    int a = 0;
    double4 me = make_double4(0, 0, 0, 0);
    int q0[9], q1[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) q0[j] = q1[j] = 0;
    int count = 0;
    double f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0;
    int ep = 0, cxw = 0, wy = 0;
    if (valid) {
        a = p.sorted_id[s];
        me = ldg256_nc(&p.sorted_state[s]);
        ep = episode_of(p, a);
        long long ix, iy;
        cell_coords(p, me.x, me.y, ix, iy);
        cxw = wrap(ix, p.G);
        wy = wrap(iy, p.Gy);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            // rows wy-1, wy, wy+1 on the wrapped grid without further remainders
            const int wr = r == 0 ? (wy == 0 ? p.Gy - 1 : wy - 1) : r == 1 ? wy : (wy == p.Gy - 1 ? 0 : wy + 1);
            const int rowbase = (ep * p.Gy + wr) * p.G;
            if (cxw >= 1 && cxw <= p.G - 2) {            // the row's three cells are contiguous slots
                q0[3 * r] = __ldg(&p.cell_start[rowbase + cxw - 1]);
                q1[3 * r] = __ldg(&p.cell_start[rowbase + cxw + 2]);
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int wc = c == 0 ? (cxw == 0 ? p.G - 1 : cxw - 1) : c == 1 ? cxw : (cxw == p.G - 1 ? 0 : cxw + 1);
                    const int cell = rowbase + wc;
                    q0[3 * r + c] = __ldg(&p.cell_start[cell]);
                    q1[3 * r + c] = __ldg(&p.cell_start[cell + 1]);
                }
            }
        }
    }
    // ---- warp-staged fast path? (warp-uniform decision) ----
    bool fast = false;
    int lo[3] = {0, 0, 0};
    // candidate tile of this warp, structure-of-arrays: lanes scanning neighbouring slots read consecutive 8-byte
    // words (an array of double4 records put slot q in banks 8q mod 32: 4- to 8-way conflicts on every read)
    double* w_x = nullptr;
    double* w_y = nullptr;
    double* w_vx = nullptr;
    double* w_vy = nullptr;
    int* w_cid = nullptr;
    if (WS) {
        unsigned char* base = s_adj_raw + (size_t)(stage_cap + 1) * ADJ_THREADS * sizeof(int);   // + the dummy stage row
        double* wd = reinterpret_cast<double*>(base) + (size_t)warp * 4 * 3 * WS_CAP;
        w_x = wd; w_y = wd + 3 * WS_CAP; w_vx = wd + 2 * 3 * WS_CAP; w_vy = wd + 3 * 3 * WS_CAP;
        w_cid = reinterpret_cast<int*>(base + (size_t)(ADJ_THREADS / 32) * 3 * WS_CAP * sizeof(double4)) + warp * 3 * WS_CAP;
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        if (vm) {
            const int src = __ffs(vm) - 1;
            const int ep0 = __shfl_sync(0xffffffffu, ep, src);
            const int wy0 = __shfl_sync(0xffffffffu, wy, src);
            const bool ok = !valid || (ep == ep0 && wy == wy0 && cxw >= 1 && cxw <= p.G - 2);
            fast = __all_sync(0xffffffffu, ok);
            if (fast) {
                int n[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    int l = valid ? q0[3 * r] : 0x7fffffff;
                    int h = valid ? q1[3 * r] : -1;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        l = min(l, __shfl_xor_sync(0xffffffffu, l, o));
                        h = max(h, __shfl_xor_sync(0xffffffffu, h, o));
                    }
                    lo[r] = l;
                    n[r] = h - l;
                }
                fast = n[0] <= WS_CAP && n[1] <= WS_CAP && n[2] <= WS_CAP;
                if (fast) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        for (int i = lane; i < n[r]; i += 32) {
                            const double4 c = ldg256_nc(&p.sorted_state[lo[r] + i]);
                            w_x[r * WS_CAP + i] = c.x; w_y[r * WS_CAP + i] = c.y;
                            w_vx[r * WS_CAP + i] = c.z; w_vy[r * WS_CAP + i] = c.w;
                            w_cid[r * WS_CAP + i] = __ldg(&p.sorted_id[lo[r] + i]);
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (valid) {
        if (WS && fast) {
            // Phase 1 -- filter: positions only, ~12 instructions per candidate, no divergent feature code.  The
            // accepted candidates' tile indices go to the stage.  (Fused, the feature code ran in nearly every
            // iteration for a handful of lanes: 19 of 32 lanes active on average.)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double* cx = w_x + r * WS_CAP - lo[r];
                const double* cy = w_y + r * WS_CAP - lo[r];
#pragma unroll 2
                for (int q = q0[3 * r]; q < q1[3 * r]; ++q) {
                    const double r2 = r2_exact(me.x - cx[q], me.y - cy[q]);
                    // branch-free: always store to the next free stage row (row `stage_cap` is a dummy), advance on accept
                    const int slot = count < stage_cap ? count : stage_cap;
                    s_stage[slot * ADJ_THREADS + tid] = r * WS_CAP + q - lo[r];
                    count += (q != s && r2 < p.R2) ? 1 : 0;
                }
            }
            // Phase 2 -- features of the accepted pairs, in the same order
            const int n_st = count < stage_cap ? count : stage_cap;
            for (int e = 0; e < n_st; ++e) {
                const int ci = s_stage[e * ADJ_THREADS + tid];
                const double4 o = make_double4(w_x[ci], w_y[ci], w_vx[ci], w_vy[ci]);
                const double dx = me.x - o.x, dy = me.y - o.y;
                const double r2 = r2_exact(dx, dy);
                FGNN_ADJ_ACCEPT(o, dx, dy, r2)
            }
            if (count > stage_cap) {                      // rare long row: the pairs beyond the stage, same order
                int w = 0;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int off = r * WS_CAP - lo[r];
                    for (int q = q0[3 * r]; q < q1[3 * r]; ++q) {
                        const double4 o = make_double4(w_x[off + q], w_y[off + q], w_vx[off + q], w_vy[off + q]);
                        const double dx = me.x - o.x, dy = me.y - o.y;
                        const double r2 = r2_exact(dx, dy);
                        if (q != s && r2 < p.R2) {
                            if (w >= stage_cap) FGNN_ADJ_ACCEPT(o, dx, dy, r2)
                            ++w;
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                for (int q = q0[j]; q < q1[j]; ++q) {
                    const double4 o = ldg256_nc(&p.sorted_state[q]);
                    const double dx = me.x - o.x, dy = me.y - o.y;
                    const double r2 = r2_exact(dx, dy);
                    if (q != s && r2 < p.R2) {
                        FGNN_ADJ_ACCEPT(o, dx, dy, r2)
                        if (count < stage_cap) s_stage[count * ADJ_THREADS + tid] = __ldg(&p.sorted_id[q]);
                        ++count;
                    }
                }
            }
        }
    }
    // reserve a contiguous run of edge slots for the warp's 32 rows
    int inc = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&p.nnz_cursor[g], (unsigned)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!valid) return;
    unsigned row = base + (unsigned)(inc - count);
    if (row + (unsigned)count > p.nnz_cap || row + (unsigned)count < row) {   // capacity exceeded: drop the row, flag it
        *p.overflow = 1;
        count = 0;
        row = 0;
    }
    int* cols = p.cols + (size_t)g * p.nnz_cap + row;
    int head[ELLW];
#pragma unroll
    for (int e = 0; e < ELLW; ++e) head[e] = -1;
    const int staged = count < stage_cap ? count : stage_cap;
    for (int e = 0; e < staged; ++e) {
        int id = s_stage[e * ADJ_THREADS + tid];
        if (WS && fast) id = w_cid[id];                   // the fast path staged tile indices
        cols[e] = id;
#pragma unroll
        for (int u = 0; u < ELLW; ++u)
            if (u == e) head[u] = id;
    }
    if (count > stage_cap) {                              // long row: re-scan for the ids beyond the stage (same order)
        int w = 0;
        if (WS && fast) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int off = r * WS_CAP - lo[r];
                for (int q = q0[3 * r]; q < q1[3 * r]; ++q) {
                    const double r2 = r2_exact(me.x - w_x[off + q], me.y - w_y[off + q]);
                    if (q != s && r2 < p.R2) {
                        if (w >= stage_cap) cols[w] = w_cid[off + q];
                        ++w;
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                for (int q = q0[j]; q < q1[j]; ++q) {
                    const double2 o = *reinterpret_cast<const double2*>(&p.sorted_state[q]);
                    const double r2 = r2_exact(me.x - o.x, me.y - o.y);
                    if (q != s && r2 < p.R2) {
                        if (w >= stage_cap) cols[w] = __ldg(&p.sorted_id[q]);
                        ++w;
                    }
                }
            }
        }
    }
    const size_t ga = (size_t)g * p.M + a;
    static_assert(ELLW == 8 && ROW == 8, "one 32-byte record per agent");
    stg256(p.ell + ga * ELLW, head);
    stg256(p.xhist + ga * ROW, (float)f0, (float)f1, (float)f2, (float)f3, (float)f4, (float)f5, 0.f, 0.f);
    p.deg[ga] = count;
    p.row_start[ga] = row;
    const float sv = p.mean_pooling ? (float)(1.0 / (double)(count > 0 ? count : 1)) : 1.0f;
    p.sinv[ga] = sv;
    // the first hop through graph t gathers x_{t-1}[m] * sinv_t[m]: keep the scale in the pad of that very row
    if (p.K > 1) p.xhist[((size_t)slot_of(t - 1, p.K) * p.M + a) * ROW + SINV_PAD] = sv;
}
#undef FGNN_ADJ_ACCEPT

#ifdef FGNN_MAIN_TU
template <bool WS>
__global__ void __launch_bounds__(ADJ_THREADS, WS ? 8 : 1) k_adjacency_t(Params p, int stage_cap) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char s_adj_raw[];
    adjacency_body<WS>(p, stage_cap, s_adj_raw);
}
#endif

#endif  // FGNN_MAIN_TU || FGNN_MINI_TU

// ------------------------------------------------------------------------------------------
// Neighbour gather shared by the hop kernels and the fused final kernel:
//   acc[b][:] = sum_{m in N_g(a)} src[b][m][:] * (PRESCALED ? 1 : sinv_g[m])      (in row order)
// Row metadata (deg, row_start) and the first ELLW column indices come from per-agent arrays, so the
// row gathers start after ONE dependent load; rows longer than ELLW continue from the CSR array.
// ------------------------------------------------------------------------------------------
template <int NB, bool PRESCALED>
__device__ __forceinline__ void gather_rows(const Params& p, int g, int a, const float* const (&src)[NB],
                                            float (&acc)[NB][F]) {
    const size_t M = p.M;
    const int d = __ldg(&p.deg[(size_t)g * M + a]);
    const unsigned rs = __ldg(&p.row_start[(size_t)g * M + a]);
    int head[ELLW];
    ldg256_nc(p.ell + ((size_t)g * M + a) * ELLW, head);
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int f = 0; f < F; ++f) acc[b][f] = 0.f;
    // Un-prescaled sources (first hop): src[0] is x_{t-1}, whose pad slot carries the scale of graph t = g.
#pragma unroll
    for (int e0 = 0; e0 < ELLW; e0 += HOP_UNROLL) {
        if (e0 < d) {
            float v[HOP_UNROLL][NB][ROW];
#pragma unroll
            for (int u = 0; u < HOP_UNROLL; ++u) {
                const int m = head[e0 + u];
                if (m >= 0) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) ldg256_nc(src[b] + (size_t)m * ROW, v[u][b]);
                }
            }
#pragma unroll
            for (int u = 0; u < HOP_UNROLL; ++u) {
                if (head[e0 + u] >= 0) {
                    const float sc = v[u][0][SINV_PAD];
#pragma unroll
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int f = 0; f < F; ++f)
                            acc[b][f] = PRESCALED ? acc[b][f] + v[u][b][f] : fmaf(v[u][b][f], sc, acc[b][f]);
                }
            }
        }
    }
    if (d > ELLW) {                                       // long rows: the tail lives in the CSR array
        const int* __restrict__ cols = p.cols + (size_t)g * p.nnz_cap + rs;
        for (int e = ELLW; e < d; e += HOP_UNROLL) {
            int m[HOP_UNROLL];
            float v[HOP_UNROLL][NB][ROW];
#pragma unroll
            for (int u = 0; u < HOP_UNROLL; ++u) m[u] = (e + u < d) ? __ldg(&cols[e + u]) : -1;
#pragma unroll
            for (int u = 0; u < HOP_UNROLL; ++u) {
                if (m[u] >= 0) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) ldg256_nc(src[b] + (size_t)m[u] * ROW, v[u][b]);
                }
            }
#pragma unroll
            for (int u = 0; u < HOP_UNROLL; ++u) {
                if (m[u] >= 0) {
                    const float sc = v[u][0][SINV_PAD];
#pragma unroll
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int f = 0; f < F; ++f)
                            acc[b][f] = PRESCALED ? acc[b][f] + v[u][b][f] : fmaf(v[u][b][f], sc, acc[b][f]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K_E  graph-shift hop j (not the last one): for taps k = j+1 .. K-1
//      Y_k <- Y_k * A_{t-j}   i.e.  out_k[n] = sum_{m in N_{t-j}(n)} in_k[m] * sinv_{t-j}[m]
//      in_k = x_{t-k} (j == 0) or the previous hop's intermediate; tap k = j+1 is finished (z_k).
//      Intermediates are stored PRE-SCALED by the source scale of the next hop's graph,
//      sinv_{t-j-1}[n], so the consumer gathers one 32-byte row per edge and nothing else.
// ------------------------------------------------------------------------------------------
#ifndef FGNN_HOP_MIN_BLOCKS
#define FGNN_HOP_MIN_BLOCKS 4
#endif
struct ShardFuse;
__device__ void shard_prepare_block(const Params& p);      // (defined with the k_shard_* kernels below)

// one thread = pool agent i
template <int NB, bool FIRST>
__device__ __forceinline__ void hop_body(const Params& p, int j, int i) {
    if (i >= hop_pool_size(p)) return;
    const int a = pool_agent(p, i);
    if (a < 0) return;
    const int t = *p.t;
    const int g = slot_of(t - j, p.K);
    const size_t M = p.M;
    const float* src[NB];
    float* dst[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        int k = j + 1 + b;
        src[b] = FIRST ? p.xhist + (size_t)slot_of(t - k, p.K) * M * ROW
                       : p.ybuf + ((size_t)((j - 1) & 1) * p.K + k) * M * ROW;
        dst[b] = (b == 0) ? p.zbuf + (size_t)k * M * ROW : p.ybuf + ((size_t)(j & 1) * p.K + k) * M * ROW;
    }
    float acc[NB][F];
    gather_rows<NB, !FIRST>(p, g, a, src, acc);
    store_row6(dst[0], a, acc[0]);
    if (NB > 1) {
        const float sn = __ldg(&p.sinv[(size_t)slot_of(t - j - 1, p.K) * M + a]);
#pragma unroll
        for (int b = 1; b < NB; ++b) {
#pragma unroll
            for (int f = 0; f < F; ++f) acc[b][f] *= sn;
            store_row6(dst[b], a, acc[b]);
        }
    }
}

template <int NB, bool FIRST>
__global__ void __launch_bounds__(256, NB == 2 ? FGNN_HOP_MIN_BLOCKS : 1) k_hop(Params p, int j, int prepare) {
    pdl_prologue();
    // p2p step of a sharded rank: block 0 of the last hop launch also zeroes the per-step counters, advances the frame and
    // works out the interior interval for the pack that follows (the hops walk the ghost-count SNAPSHOT, so the counter may be
    // reset under them): the one-block k_shard_prepare launch leaves the step's critical path
    if (prepare && blockIdx.x == 0) shard_prepare_block(p);
    hop_body<NB, FIRST>(p, j, blockIdx.x * blockDim.x + threadIdx.x);
}

// ---- multi-GPU halo / hand-over control block (see the k_shard_* kernels at the end of this file) ----
constexpr int SREC = 6;

struct ShardCtl {
    const double* bounds;     // [world + 1] strip boundaries at shift = 0 (bounds[0] = -inf, bounds[world] = +inf)
    double* shift;            // device scalar: frame displacement so far
    double dshift;            // per step
    double depth;             // halo depth for the windows
    double margin;            // hand-over hysteresis
    int world, rank;
    int handover_after;       // first step index at which hand-overs are allowed (histories must be valid)
    int* own;                 // [pool_cap] owned list: order kept, -1 = handed over, appended at n_own
    int* n_own;               // high-water mark
    int* free_slots;          // [pool_cap] stack of tombstoned positions, reused by the next agents received
    int* n_free;
    int* ghost;               // [pool_cap]
    int* n_ghost;
    int* counter;             // records written
    long long* xminmax;       // [pack blocks][2] order-preserving keys of min / max px over the agents kept, per block
    double* safe;             // [2] this step's interior x-interval: an owned agent strictly inside it is wanted by no other rank
                              //     and stays owned (k_shard_prepare): the pack step costs it two comparisons
};

// order-preserving map double <-> signed 64-bit integer (for atomicMin / atomicMax on coordinates)
__device__ __forceinline__ long long dkey(double x) {
    long long b = __double_as_longlong(x);
    return b >= 0 ? b : b ^ 0x7fffffffffffffffll;
}
__device__ __forceinline__ double dunkey(long long k) {
    return __longlong_as_double(k >= 0 ? k : k ^ 0x7fffffffffffffffll);
}

__device__ __forceinline__ int strip_of(const ShardCtl& c, double xs) {
    int s = 0;
    for (int q = 1; q < c.world; ++q) s += (xs >= c.bounds[q]) ? 1 : 0;
    return s;
}


// Arguments of the pack step (device-resident when it is fused into the closed final kernel: the CUDA graph of a step
// stays valid when buffers change).  Two transports:
//   gather (p2p == 0): every record goes into ONE send buffer that the host all-gathers to every rank;
//   p2p    (p2p == 1): a record goes straight into the inbox of each rank that needs it -- plain stores into peer memory
//                      over NVLink (CUDA-IPC mapped) -- slot [parity][sender] of the receiver, parity = t & 1.
struct ShardFuse {
    ShardCtl ctl;
    const double* windows;    // gather: [q * wstride + {0, 1}] = rank q's owned x-interval (one step old)
    long long wstride;
    double* buf;              // gather: send buffer, (cap + 1) records
    int cap;
    int p2p;
    double* const* peer_inbox;    // p2p: [world] base of every rank's inbox  [2][world][cap + 1][SREC]  (own entry: local)
    int* const* peer_flags;       // p2p: [world] base of every rank's flags  [2][world]
    int* dest_count;              // p2p: [world] records written for each destination this step
};

__device__ __forceinline__ size_t inbox_offset(const ShardFuse& f, int parity, int sender) {
    return ((size_t)(parity * f.ctl.world + sender) * (f.cap + 1)) * SREC;
}

// Hand-over decision and halo record of ONE owned agent whose (new) state is `st` (used by k_shard_pack and,
// fused, by the epilogue of the closed final kernel).  i = position in the owned list.
__device__ __forceinline__ void shard_pack_agent(const Params& p, const ShardFuse& f, int i, int a, const double4 st,
                                                 long long& klo, long long& khi, double safe_lo, double safe_hi) {
    const ShardCtl& c = f.ctl;
    if (st.x > safe_lo && st.x < safe_hi) {               // interior of the strip (~99 % of a large shard): nothing to decide
        const long long k = dkey(st.x);
        klo = min(klo, k);
        khi = max(khi, k);
        return;
    }
    const double shift = *c.shift;
    const double xs = st.x - shift;
    const int t = *p.t;
    int new_owner = -1;
    if (t >= c.handover_after) {
        const int sp = strip_of(c, xs);
        if (sp > c.rank && xs - c.bounds[c.rank + 1] > c.margin) new_owner = sp;
        if (sp < c.rank && c.bounds[c.rank] - xs > c.margin) new_owner = sp;
    }
    if (new_owner < 0) {
        const long long k = dkey(st.x);
        klo = min(klo, k);
        khi = max(khi, k);
    } else {                                              // tombstone; stays here as a ghost for this step (already binned)
        c.own[i] = -1;
        c.free_slots[atomicAdd(c.n_free, 1)] = i;         // at most pool_cap tombstones can exist
        const int slot = atomicAdd(c.n_ghost, 1);
        if (slot < p.pool_cap) c.ghost[slot] = a; else *p.overflow = 1;
    }
    // who needs this agent's state?  windows: the x-intervals the ranks announced one step ago
    const double* windows = f.p2p ? f.peer_inbox[c.rank] + inbox_offset(f, (t + 1) & 1, 0) + 1 : f.windows;
    const long long wstride = f.p2p ? (long long)(f.cap + 1) * SREC : f.wstride;
    bool wanted = new_owner >= 0 && !f.p2p;
    const int lane = threadIdx.x & 31;
    for (int q = 0; q < c.world && !wanted; ++q) {
        if (q == c.rank) continue;
        const bool in_strip = xs >= c.bounds[q] - c.depth && xs <= c.bounds[q + 1] + c.depth;
        const bool in_ival = st.x >= windows[q * wstride] - c.depth && st.x <= windows[q * wstride + 1] + c.depth;
        if (!f.p2p) {
            wanted = in_strip || in_ival;
        } else {
            // boundary agents come in runs (the last cells of a grid row): the lanes of a warp that send to q take their
            // slots with ONE atomic (thousands of single-address atomics were the cost of the sharded final kernel)
            const bool sends = in_strip || in_ival || q == new_owner;
            const unsigned act = __activemask();
            const unsigned m = __ballot_sync(act, sends);
            if (sends) {
                const int leader = __ffs(m) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&f.dest_count[q], __popc(m));
                base = __shfl_sync(m, base, leader);
                const int slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < f.cap) {                       // overflow is reported through the header count
                    double2* rec = reinterpret_cast<double2*>(f.peer_inbox[q] + inbox_offset(f, t & 1, c.rank) + (size_t)(slot + 1) * SREC);
                    rec[0] = make_double2((double)a, st.x);   // three 16-byte stores into peer memory; they are complete when
                    rec[1] = make_double2(st.y, st.z);        // this grid is -- k_shard_flag (a later launch) fences and raises
                    rec[2] = make_double2(st.w, (double)new_owner);   // the flag
                }
            }
        }
    }
    if (!f.p2p) {                                          // gather transport: same aggregation on the single send buffer
        const unsigned act = __activemask();
        const unsigned m = __ballot_sync(act, wanted);
        if (wanted) {
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(c.counter, __popc(m));
            base = __shfl_sync(m, base, leader);
            const int slot = base + __popc(m & ((1u << lane) - 1u));
            if (slot < f.cap) {                            // overflow is reported through the header count
                double* rec = f.buf + (size_t)(slot + 1) * SREC;
                rec[0] = (double)a; rec[1] = st.x; rec[2] = st.y; rec[3] = st.z; rec[4] = st.w; rec[5] = (double)new_owner;
            }
        }
    }
}

// per-block reduction of the kept x-interval into slot blockIdx.x (k_shard_header reduces the slots)
template <int THREADS>
__device__ __forceinline__ void shard_interval_flush(const ShardCtl& c, long long klo, long long khi) {
    __shared__ long long s_klo[THREADS / 32], s_khi[THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, o));
        khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_klo[threadIdx.x >> 5] = klo; s_khi[threadIdx.x >> 5] = khi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < THREADS / 32; ++w) { klo = min(klo, s_klo[w]); khi = max(khi, s_khi[w]); }
        c.xminmax[2 * blockIdx.x] = klo;
        c.xminmax[2 * blockIdx.x + 1] = khi;
    }
}

// p2p transport, after the final kernel has stored the records into the peers' inboxes: header [count, x_lo, x_hi] of this
// rank into every peer's inbox (and its own), then -- fenced -- the flag word t + 1 that tells the peer its half t & 1 is
// complete.  One block; thread q serves peer q.
template <int THREADS>
__device__ __forceinline__ void shard_flag_body(const Params& p, const ShardFuse& f, int n_blocks) {
    __shared__ long long s_lo[THREADS / 32], s_hi[THREADS / 32];
    const ShardCtl& c = f.ctl;
    long long klo = 0x7fffffffffffffffll, khi = -0x7fffffffffffffffll - 1;
    for (int i = threadIdx.x; i < n_blocks; i += THREADS) {
        klo = min(klo, c.xminmax[2 * i]);
        khi = max(khi, c.xminmax[2 * i + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, o));
        khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = klo; s_hi[threadIdx.x >> 5] = khi; }
    __syncthreads();
    for (int w = 0; w < THREADS / 32; ++w) { klo = min(klo, s_lo[w]); khi = max(khi, s_hi[w]); }
    const int t = *p.t;
    for (int q = threadIdx.x; q < c.world; q += THREADS) {
        double* hdr = f.peer_inbox[q] + inbox_offset(f, t & 1, c.rank);
        hdr[0] = q == c.rank ? 0.0 : (double)f.dest_count[q];
        hdr[1] = dunkey(klo); hdr[2] = dunkey(khi);
        hdr[3] = 0.0; hdr[4] = 0.0; hdr[5] = 0.0;
        __threadfence_system();
        volatile int* flag = f.peer_flags[q] + (t & 1) * c.world + c.rank;      // (own entry: "my header of this step is written")
        *flag = t + 1;
    }
}

// double integrator exactly in numpy's evaluation order (no FMA contraction), then bin the new position
// (u0, u1): the action as float64 -- exactly (double)fp32 for the policy's own actions (learner/gnn_dagger.py:161), the
// controller's float64 value when the expert drives the env (learner/gnn_dagger.py:156-163)
__device__ __forceinline__ double4 integrate_and_bin(const Params& p, int a, const double4 s, double u0, double u1,
                                                     double (&racc)[4]) {
    double ax = __dmul_rn(u0, p.gain), ay = __dmul_rn(u1, p.gain);
    if (p.amask && p.amask[a] == 0) { ax = 0.0; ay = 0.0; }      // u * mask (leaders keep their velocity)
    double nx = __dadd_rn(s.x, __dmul_rn(s.z, p.dt));
    double ny = __dadd_rn(s.y, __dmul_rn(s.w, p.dt));
    if (p.half_accel) {
        nx = __dadd_rn(nx, __dmul_rn(__dmul_rn(__dmul_rn(ax, p.dt), p.dt), 0.5));
        ny = __dadd_rn(ny, __dmul_rn(__dmul_rn(__dmul_rn(ay, p.dt), p.dt), 0.5));
    }
    const double nvx = __dadd_rn(s.z, __dmul_rn(ax, p.dt));
    const double nvy = __dadd_rn(s.w, __dmul_rn(ay, p.dt));
    stg256(&p.state[a], make_double4(nx, ny, nvx, nvy));
    long long ix, iy;
    cell_coords(p, nx, ny, ix, iy);
    const int ep = episode_of(p, a);
    const int c = cell_index(p, ep, ix, iy);
    p.cell_of[a] = c;
    bin_agent(p, c);
    // velocity-variance reward sums
    if (p.B == 1) {                 // thread-local; reduced per block by reward_block_flush()
        racc[0] += nvx; racc[1] += nvy; racc[2] += nvx * nvx; racc[3] += nvy * nvy;
    } else {                        // slotted atomics, warp-reduced when the warp sits in one episode
        double v0 = nvx, v1 = nvy, v2 = nvx * nvx, v3 = nvy * nvy;
        const unsigned mask = __activemask();
        const int ep0 = __shfl_sync(mask, ep, __ffs(mask) - 1);
        const bool uniform = __all_sync(mask, ep == ep0);
        double* dst = p.racc + ((size_t)(blockIdx.x % rslots_of(p.B)) * p.B + ep) * 4;
        if (uniform && mask == 0xffffffffu) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                v3 += __shfl_xor_sync(0xffffffffu, v3, o);
            }
            if ((threadIdx.x & 31) == 0) {
                atomicAdd(dst + 0, v0); atomicAdd(dst + 1, v1); atomicAdd(dst + 2, v2); atomicAdd(dst + 3, v3);
            }
        } else {
            atomicAdd(dst + 0, v0); atomicAdd(dst + 1, v1); atomicAdd(dst + 2, v2); atomicAdd(dst + 3, v3);
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *p.reward_pending = 1;
    return make_double4(nx, ny, nvx, nvy);
}

// B == 1: sum the block's thread-local reward sums in a fixed order and store them as this block's partial.
// Must be called by every thread of the block (even ones that integrated nothing).
template <int THREADS>
__device__ __forceinline__ void reward_block_flush(const Params& p, const double (&racc)[4]) {
    if (p.B != 1) return;
    __shared__ double s_part[4][THREADS / 32];
    double v[4] = {racc[0], racc[1], racc[2], racc[3]};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s_part[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += s_part[threadIdx.x][w];
        p.racc_part[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *p.n_partials = gridDim.x; *p.reward_pending = 1; }
}

#ifdef FGNN_MAIN_TU
// first half of env.step(u) with an externally supplied action
template <typename T2>      // float2: select_action's fp32 action; double2: the controller's float64 action
__global__ void __launch_bounds__(256) k_integrate(Params p, const T2* __restrict__ u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double racc[4] = {0, 0, 0, 0};
    if (i < owned_count(p)) {
        const int a = owned_agent(p, i);
        if (a >= 0) {
            const T2 uu = u[i];                                            // u is in owned-list order
            integrate_and_bin(p, a, ldg256(&p.state[a]), (double)uu.x, (double)uu.y, racc);
        }
    }
    reward_block_flush<256>(p, racc);
}

// ------------------------------------------------------------------------------------------
// read-back helpers
// ------------------------------------------------------------------------------------------
__global__ void k_pack_rows(const float* __restrict__ rows, float* __restrict__ out, int M) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * F) return;
    int a = i / F, f = i % F;
    out[i] = rows[(size_t)a * ROW + f];
}

__global__ void k_export_dense(Params p, int g, float* __restrict__ out) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= p.M) return;
    const int ep = a / p.N, i = a % p.N;
    const unsigned rs = p.row_start[(size_t)g * p.M + a];
    const int d = p.deg[(size_t)g * p.M + a];
    const float sc = p.sinv[(size_t)g * p.M + a];
    const int* cols = p.cols + (size_t)g * p.nnz_cap + rs;
    const int* head = p.ell + ((size_t)g * p.M + a) * ELLW;      // the first ELLW neighbours (CSR rows may hold long rows only)
    float* row = out + ((size_t)ep * p.N + i) * p.N;
    for (int e = 0; e < d; ++e) row[(e < ELLW ? head[e] : cols[e]) % p.N] = sc;
}


// Complete CSR rows for fgnn_get_csr when the step keeps only the ELL head + the tail of long rows (pair kernels):
// row a = head[0 .. min(d, ELLW)) followed by cols[row_start + ELLW .. row_start + d); rows placed by one atomic per warp.
__global__ void __launch_bounds__(256) k_csr_assemble(Params p, int g, unsigned* __restrict__ cursor, unsigned* __restrict__ row_out,
                                                      int* __restrict__ cols_out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const size_t ga = (size_t)g * p.M + (a < p.M ? a : 0);
    const int d = a < p.M ? p.deg[ga] : 0;
    int inc = d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    unsigned base = 0;
    if (lane == 31 && inc > 0) base = atomicAdd(cursor, (unsigned)inc);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (a >= p.M) return;
    const unsigned row = base + (unsigned)(inc - d);
    row_out[a] = row;
    if (row + (unsigned)d > p.nnz_cap) { *p.overflow = 1; return; }
    const int* head = p.ell + ga * ELLW;
    const int* tail = p.cols + (size_t)g * p.nnz_cap + p.row_start[ga];
    for (int e = 0; e < d; ++e) cols_out[row + e] = e < ELLW ? head[e] : tail[e];
}

// ------------------------------------------------------------------------------------------
// Expert controller (gym_flock FlockingRelativeEnv.controller):
//   u_i = -sum_j [ grad(dp_ij, r2_ij) * 1(r2_ij <= comm_radius) + dv_ij ] over j in the mask,
//   grad(d, r2) = -2 d / r2^2 + 2 d / r2;  mask = radius neighbours (decentralised) or every other
//   agent of the episode (centralised); clipped to +-max_accel*gain and divided by gain.
// Runs on the cell structure of the CURRENT graph (sorted_state / cell_start), float64.
// Centralised: the velocity term over all agents is N v_i - sum_j v_j (per-episode sums, k_vel_sum);
// the potential term has a finite cut-off sqrt(comm_radius) and is found by a (2w+1)^2 cell search.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vel_sum(Params p, double* __restrict__ vsum /* [B][2] */) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= p.M) return;
    const double4 s = p.state[a];
    const int ep = a / p.N;
    atomicAdd(&vsum[ep * 2 + 0], s.z);
    atomicAdd(&vsum[ep * 2 + 1], s.w);
}

template <typename T2>      // float2 / double2 output
__global__ void __launch_bounds__(128) k_controller(Params p, int centralized, int window, double grad_cut /* comm_radius */,
                                                    double max_u, const double* __restrict__ vsum, T2* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= sorted_count(p)) return;
    const int a = p.sorted_id[s];
    const double4 me = p.sorted_state[s];
    const int ep = a / p.N;
    long long ix, iy;
    cell_coords(p, me.x, me.y, ix, iy);
    double gx = 0, gy = 0, dvx = 0, dvy = 0;
    // distinct wrapped cells only: a window wider than the grid would visit cells twice
    const int span = 2 * window + 1 <= p.G ? 2 * window + 1 : p.G;
    const int spany = 2 * window + 1 <= p.Gy ? 2 * window + 1 : p.Gy;
    const long long x0 = 2 * window + 1 <= p.G ? ix - window : 0;
    const long long y0 = 2 * window + 1 <= p.Gy ? iy - window : 0;
    for (int ry = 0; ry < spany; ++ry) {
        for (int rx = 0; rx < span; ++rx) {
            const int c = cell_index(p, ep, x0 + rx, y0 + ry);
            const int q0 = __ldg(&p.cell_start[c]), q1 = __ldg(&p.cell_start[c + 1]);
            for (int q = q0; q < q1; ++q) {
                if (q == s) continue;
                const double4 o = p.sorted_state[q];
                const double dx = me.x - o.x, dy = me.y - o.y;
                const double r2 = r2_exact(dx, dy);
                const bool nb = r2 < p.R2;
                if (!centralized && !nb) continue;
                if (!(r2 > grad_cut)) {
                    const double inv = 1.0 / r2;
                    const double gsc = -2.0 * inv * inv + 2.0 * inv;
                    gx += dx * gsc;
                    gy += dy * gsc;
                }
                if (!centralized) {
                    dvx += me.z - o.z;
                    dvy += me.w - o.w;
                }
            }
        }
    }
    if (centralized) {
        dvx = (double)p.N * me.z - vsum[ep * 2 + 0];
        dvy = (double)p.N * me.w - vsum[ep * 2 + 1];
    }
    double ux = -gx - dvx, uy = -gy - dvy;
    ux = fmin(fmax(ux, -max_u), max_u) / p.gain;
    uy = fmin(fmax(uy, -max_u), max_u) / p.gain;
    T2 o;
    o.x = (decltype(o.x))ux;
    o.y = (decltype(o.y))uy;
    out[a] = o;
}


// ------------------------------------------------------------------------------------------
// Multi-GPU halo exchange with ownership hand-over (every rank keeps full-size arrays, indices are global).
// Record = SREC doubles: [agent id, px, py, vx, vy, new owner or -1]; record 0 of a rank's buffer is the
// header [count, own x_lo, own x_hi, 0, 0, 0].
//
// Territories are x-strips in a frame moving with the flock (bounds[q] <= x - shift < bounds[q+1] is rank q's
// strip; shift advances by dshift per step).  An owner hands an agent over to the rank whose strip it has
// entered by more than `margin`: the receiver already holds the agent as a ghost with a valid K-deep
// history (it recomputes graph / features / hops for everything inside its window), so the hand-over moves
// NO data -- the record just names the new owner.  Owned sets therefore stay spatially compact and the
// halo stays a thin layer however long the rollout runs; an arbitrary initial index order re-partitions
// itself after `handover_after` steps.
//   k_shard_prepare : zero the per-step counters, advance the frame shift
//   k_shard_pack    : hand-over decisions (tombstone in the owned list), records for the other ranks' windows
//   k_shard_unpack  : install received states; new owner -> appended to the owned list, otherwise ghost list; bin
// A rank's window = its strip +- depth, united with the x-interval of what it still owns +- depth.
// ------------------------------------------------------------------------------------------
// Per step, before the pack: zero the counters, advance the frame, and work out the interior interval (safe[0], safe[1]) of
// absolute x inside which an owned agent (a) lies strictly inside this rank's strip, so it is not handed over, and (b) is
// outside every other rank's window (strip +- depth united with its announced owned interval +- depth), so nobody wants
// its record.  Conservative: empty when some other rank's window covers the middle of what this rank owns.
__device__ __forceinline__ void shard_prepare_body(const Params& p, const ShardFuse& f, int advance) {
    const ShardCtl& c = f.ctl;
    if (f.dest_count && threadIdx.x < c.world) f.dest_count[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        *c.n_ghost = 0; *c.counter = 0;
        double shift = *c.shift;
        if (advance) { shift += c.dshift; *c.shift = shift; }
        const int t = *p.t;
        const double* windows = f.p2p ? f.peer_inbox[c.rank] + inbox_offset(f, (t + 1) & 1, 0) + 1 : f.windows;
        const long long wstride = f.p2p ? (long long)(f.cap + 1) * SREC : f.wstride;
        double lo = c.bounds[c.rank] + shift, hi = c.bounds[c.rank + 1] + shift;       // strictly inside the own strip
        // seed: the middle of what this rank owned one step ago; the interval grows from it until it meets a window
        const double centre = 0.5 * (windows[c.rank * wstride] + windows[c.rank * wstride + 1]);
        if (!(centre > lo && centre < hi)) { lo = 1.0; hi = -1.0; }
        for (int q = 0; q < c.world; ++q) {
            if (q == c.rank) continue;
            const double wl = fmin(c.bounds[q] + shift, windows[q * wstride]) - c.depth;
            const double wh = fmax(c.bounds[q + 1] + shift, windows[q * wstride + 1]) + c.depth;
            if (wh < centre) lo = fmax(lo, wh);
            else if (wl > centre) hi = fmin(hi, wl);
            else { lo = 1.0; hi = -1.0; }                  // a window straddles the centre: no interior
        }
        c.safe[0] = lo;
        c.safe[1] = hi;
    }
}

__global__ void k_shard_prepare(Params p, ShardFuse f, int advance) {
    if (blockIdx.x == 0) shard_prepare_body(p, f, advance);
}

__device__ void shard_prepare_block(const Params& p) { shard_prepare_body(p, *p.fuse, 1); }

__global__ void __launch_bounds__(256) k_shard_pack(Params p, ShardFuse f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    long long klo = 0x7fffffffffffffffll, khi = -0x7fffffffffffffffll - 1;
    const int a = i < owned_count(p) ? owned_agent(p, i) : -1;
    if (a >= 0) shard_pack_agent(p, f, i, a, p.state[a], klo, khi, f.ctl.safe[0], f.ctl.safe[1]);
    shard_interval_flush<256>(f.ctl, klo, khi);
}

// header record [count, x_lo, x_hi, 0, 0, 0]: reduce the per-block intervals written by k_shard_pack
__global__ void __launch_bounds__(256) k_shard_header(ShardCtl c, double* __restrict__ buf, int n_blocks) {
    __shared__ long long s_lo[8], s_hi[8];
    long long klo = 0x7fffffffffffffffll, khi = -0x7fffffffffffffffll - 1;
    for (int i = threadIdx.x; i < n_blocks; i += blockDim.x) {
        klo = min(klo, c.xminmax[2 * i]);
        khi = max(khi, c.xminmax[2 * i + 1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, o));
        khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = klo; s_hi[threadIdx.x >> 5] = khi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { klo = min(klo, s_lo[w]); khi = max(khi, s_hi[w]); }
        buf[0] = (double)*c.counter; buf[1] = dunkey(klo); buf[2] = dunkey(khi);
        buf[3] = 0.0; buf[4] = 0.0; buf[5] = 0.0;
    }
}

__global__ void __launch_bounds__(256) k_shard_flag(Params p, ShardFuse f, int n_blocks) { shard_flag_body<256>(p, f, n_blocks); }

// grid = (record chunks, world): a block of sender q leaves at once when q sent fewer records than its first slot.
// parity_stride != 0 (p2p inbox): this step's records sit in half t & 1 of `recv`.
// flag_fuse != null (p2p): block (0, 0) first does k_shard_flag's work -- headers and flags of THIS rank to every peer -- so the
// step needs no separate one-block launch for it; every block then waits for its sender's flag and for this rank's own
// (the own header, which the window test below reads, is written by block (0, 0) of this very grid).
__global__ void __launch_bounds__(256) k_shard_unpack(Params p, ShardCtl c, const double* __restrict__ recv /* [world][cap+1][SREC] */,
                                                      int cap, long long parity_stride, const int* wait_flags,
                                                      const ShardFuse* flag_fuse, int n_pack_blocks) {
    const int q = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    if (flag_fuse && blockIdx.x == 0 && blockIdx.y == 0) {
        shard_flag_body<256>(p, *flag_fuse, n_pack_blocks);
        __syncthreads();
    }
    if (q == c.rank) return;
    if (wait_flags) {                  // p2p: the blocks of sender q wait for q's flag of this step themselves (no separate launch).
        if (threadIdx.x == 0) {        // A peer that never arrives must not hang the GPU: after ~30 s give up, overflow = 2.
                                       // (Not less: the first steps of a many-rank job see seconds of skew -- graph
                                       // instantiation, lazily enabled peer mappings -- and a wait that gives up early
                                       // installs a stale half of the inbox.)
            const int t = *p.t;
            const long long t0 = clock64();
            for (int w = 0; w < 2; ++w) {
                volatile const int* flag = wait_flags + (t & 1) * c.world + (w == 0 ? q : c.rank);
                while (*flag != t + 1) {
                    __nanosleep(100);
                    if (clock64() - t0 > 60000000000ll) { *p.overflow = 2; break; }
                }
            }
            __threadfence_system();
        }
        __syncthreads();
    }
    recv += (size_t)(*p.t & 1) * parity_stride;
    const double* base = recv + (size_t)q * (cap + 1) * SREC;
    const int count = (int)base[0];
    if (r == 0 && count > cap) *p.overflow = 1;
    if (r >= count || r >= cap) return;
    const double* rec = base + (size_t)(r + 1) * SREC;
    const int new_owner = (int)rec[5];
    // this rank's window: its strip and the x-interval of what it owns (own header), +- depth
    const double* own_hdr = recv + (size_t)c.rank * (cap + 1) * SREC;
    const double shift = *c.shift;
    const double xs = rec[1] - shift;
    const bool in_strip = xs >= c.bounds[c.rank] - c.depth && xs <= c.bounds[c.rank + 1] + c.depth;
    const bool in_ival = rec[1] >= own_hdr[1] - c.depth && rec[1] <= own_hdr[2] + c.depth;
    if (new_owner != c.rank && !in_strip && !in_ival) return;  // not near this rank
    const int a = (int)rec[0];
    p.state[a] = make_double4(rec[1], rec[2], rec[3], rec[4]);
    if (new_owner == c.rank) {                                 // into a tombstoned slot if there is one, else appended
        const int k = atomicSub(c.n_free, 1) - 1;
        if (k >= 0) {
            c.own[c.free_slots[k]] = a;
        } else {
            atomicAdd(c.n_free, 1);
            const int slot = atomicAdd(c.n_own, 1);
            if (slot < p.pool_cap) c.own[slot] = a; else { atomicSub(c.n_own, 1); *p.overflow = 1; }
        }
    } else {
        const int slot = atomicAdd(c.n_ghost, 1);
        if (slot < p.pool_cap) c.ghost[slot] = a; else *p.overflow = 1;
    }
    long long ix, iy;
    cell_coords(p, rec[1], rec[2], ix, iy);
    const int cc = cell_index(p, a / p.N, ix, iy);
    p.cell_of[a] = cc;
    bin_agent(p, cc);
}

// owned list = contiguous range [lo, lo + count)  (reset)
__global__ void __launch_bounds__(256) k_own_init(int* __restrict__ own, int* __restrict__ n_own, int* __restrict__ n_ghost,
                                                  int* __restrict__ n_free, int lo, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) own[i] = lo + i;
    if (i == 0) { *n_own = count; *n_ghost = 0; *n_free = 0; }
}

// owned-list order <-> global arrays (policy / integrate on a sharded handle)
__global__ void __launch_bounds__(256) k_gather_owned2(Params p, const float* __restrict__ src, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= owned_count(p)) return;
    const int a = owned_agent(p, i);
    reinterpret_cast<float2*>(dst)[i] = a >= 0 ? reinterpret_cast<const float2*>(src)[a] : make_float2(0.f, 0.f);
}

#endif  // FGNN_MAIN_TU

}  // namespace fgnn
