// fgnn.cu -- host side of libfgnn.so: handle, memory, launches, CUDA-graph rollout, C ABI (include/fgnn.h).
#define FGNN_MAIN_TU 1
#include "fgnn_kernels.cuh"
#include "fgnn_final_tc.cuh"
#include "fgnn_pair.cuh"
#include "../../include/fgnn.h"

#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace fgnn;

static thread_local std::string g_err;
static int fail(const std::string& msg) { g_err = msg; return 1; }
namespace fgnn { void set_error(const char* msg) { g_err = msg ? msg : ""; } }     // for the other translation units

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return fail(buf__);                                                                      \
        }                                                                                            \
    } while (0)

// ---- NCCL inside the step graph ---------------------------------------------------------------------------------
// The library is resolved at run time (dlopen): a process that has imported torch already holds libnccl.so.2, so
// the same NCCL serves both; nothing is linked at build time.  Declarations follow nccl.h (ncclUniqueId is 128
// bytes passed by value, ncclFloat64 = 8, ncclSuccess = 0).
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool tried = false, ok = false;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.tried) return g_nccl.ok ? 0 : fail("NCCL library not available (libnccl.so.2)");
    g_nccl.tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(std::string("dlopen(libnccl.so.2) failed: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.CommDestroy && g_nccl.GetErrorString;
    return g_nccl.ok ? 0 : fail("libnccl.so.2 lacks the expected symbols");
}

int nccl_check(int rc, const char* what) {
    if (rc == 0) return 0;
    return fail(std::string(what) + " -> " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"));
}
}  // namespace

#define FGNN_MAX_CHUNKS 32
#ifndef FGNN_POLICY_CHUNKS_DEFAULT
#define FGNN_POLICY_CHUNKS_DEFAULT 4              // measured: e2e 0.664 -> 0.626 ms/step at N=1M (scripts/policy_chunks_check.py)
#endif
#ifndef FGNN_POLICY_PIPE_DEFAULT
#define FGNN_POLICY_PIPE_DEFAULT 2                // waves of the readout grid per chunk of the pipelined host path (0: equal chunks)
#endif

struct fgnn_handle {
    fgnn_config cfg;
    Params p;
    int HP = 0;
    int sm_count = 148;
    size_t weights_floats = 0;
    std::vector<float> w_host;       // packed weights, host mirror
    float* d_weights = nullptr;
    unsigned char* d_amask = nullptr; // [M] leader mask (allocated by fgnn_set_agent_mask)
    float* d_u_in = nullptr;         // [M][2] staged external action
    float* d_staging = nullptr;      // read-back staging (K*M*6 floats)
    double* d_reward_log = nullptr;
    int reward_log_cap = 0;          // in steps
    bool binned = false;             // cell_of/cell_count describe the current positions
    bool weights_dirty = true;
    int64_t launches = 0;
    int64_t t_host = -1;
    std::vector<void*> allocs;
    // graph of one closed-loop step
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    int graph_kernels = 0;
    int final_grid_closed = 0, final_grid_open = 0;
    size_t final_smem = 0;
    int adj_stage = 16;              // neighbour ids staged per thread in k_adjacency
    bool adj_warp_staged = false;    // k_adjacency_t<true>: candidates staged per warp in shared memory
    bool last_hop_separate = false;  // last hop as its own launch instead of inside the final kernel
    bool scan_two_pass = false;      // tile sums in their own launch: the scan proper never waits on another block
    bool pdl = false;                // programmatic dependent launch between the step kernels
    bool pair_mode = false;          // k_pair_adjacency instead of k_adjacency_t: warp-tiled, TMA-staged, fp32 pre-filter (fgnn_pair.cuh)
    int pair_minb = 6;
    PairGeom geo;                    // fp32 pre-filter thresholds, cell-index dividers, source-scale table
    unsigned* d_csr_rows = nullptr;  // fgnn_get_csr in pair mode: complete rows assembled on demand (ELL head + CSR tail)
    int* d_csr_cols = nullptr;
    unsigned* d_csr_cursor = nullptr;
    int policy_chunks = 1;           // fgnn_policy to a host buffer: readout chunks overlapped with their D2H copies
    int policy_pipe = 0;             // > 0: chunks of this many waves of the readout grid, the LAST hop inside each chunk's readout
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_event[FGNN_MAX_CHUNKS + 1] = {};
    // small flocks (B*N <= 128): the whole closed-loop step in one CTA, T steps per launch (fgnn_mini.cu)
    void (*mini_kernel)(Params, const uint8_t*, int, int, int) = nullptr;
    void (*mini_policy)(Params, const uint8_t*) = nullptr;
    void (*mini_envstep)(Params, const void*, int, int, int) = nullptr;
    size_t mini_smem = 0;
    int mini_adj_off = 0;
    // tensor-core readout (tcgen05, 3xTF32)
    bool use_tc = false;
    std::vector<uint8_t> tc_host;    // TcLayout pack, host mirror
    uint8_t* d_tc_weights = nullptr;
    size_t tc_smem = 0;
    int tc_grid_closed = 0, tc_grid_open = 0;
    std::vector<std::vector<float>> raw_w, raw_b;   // per layer, reference layout
    // sharding
    bool sharded = false;
    bool shard_configured = false;
    bool shard_fold = true;          // p2p step: k_shard_prepare's work in block 0 of the last hop
    ShardCtl ctl;
    int* d_own = nullptr;
    int* d_ghost = nullptr;
    int* d_counts = nullptr;         // [n_own, n_free, n_ghost, record counter]
    int* d_free = nullptr;
    ShardFuse* d_fuse = nullptr;     // device copy of the fused-pack arguments
    ShardFuse fuse_host;             // what d_fuse currently holds
    long long* d_xminmax = nullptr;
    double* d_shift = nullptr;
    double* d_safe = nullptr;         // [2] interior x-interval of the step (k_shard_prepare)
    double* d_bounds = nullptr;      // [world + 1], allocated by fgnn_shard_configure
    int launch_pool = 0;             // grid sizing for kernels over the pool (pool capacity, or M)
    void* nccl_comm = nullptr;       // ncclComm_t of fgnn_comm_init (the halo all-gather inside the step graph)
    bool nccl_warm = false;
    // p2p halo transport (fgnn_p2p_*): records stored straight into the peers' inboxes over NVLink
    double* p2p_inbox = nullptr;     // [2][world][cap + 1][SREC] doubles followed by the flag words [2][world] ints
    size_t p2p_bytes = 0;
    int p2p_cap = 0, p2p_world = 0;
    bool p2p_connected = false;
    double** d_peer_inbox = nullptr; // [world] device array of inbox bases (own entry: p2p_inbox)
    int** d_peer_flags = nullptr;
    int* d_dest_count = nullptr;
    std::vector<void*> p2p_opened;   // cudaIpcOpenMemHandle mappings to close
    void* shard_graph_store = nullptr;
    long long shard_epoch = 0;       // bumped by fgnn_shard_configure: invalidates cached graphs
    // per-kernel profiling of one step (fgnn_profile_step)
    bool profiling = false;
    cudaStream_t prof_stream = nullptr;
    std::vector<cudaEvent_t> prof_events;
    std::vector<std::string> prof_names;
};

template <typename T>
static int dalloc(fgnn_handle* h, T** ptr, size_t count, bool zero = true) {
    void* q = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = sizeof(T);
    CK(cudaMalloc(&q, bytes));
    if (zero) CK(cudaMemset(q, 0, bytes));
    h->allocs.push_back(q);
    *ptr = reinterpret_cast<T*>(q);
    return 0;
}

// kernel-variant defaults (each can be overridden per process by the environment variable named in fgnn_create)
#ifndef FGNN_ADJ_DEFAULT_WS
#define FGNN_ADJ_DEFAULT_WS true                 // measured: 120 -> 112 us at N=1M, d~5 (profiles/r1_bench_history.md)
#endif
#ifndef FGNN_PDL_DEFAULT
#define FGNN_PDL_DEFAULT false
#endif
#ifndef FGNN_SCAN_TWO_PASS_DEFAULT
#define FGNN_SCAN_TWO_PASS_DEFAULT true           // measured: 312 -> 294 us/step; blocks spinning on other blocks' status words are slow here
#endif
#ifndef FGNN_STEP_MODE_DEFAULT
#define FGNN_STEP_MODE_DEFAULT 1                 // 1: k_pair_adjacency (warp tiles, TMA staging, fp32 pre-filter) where the geometry suits it:
#endif                                           // measured at N=1M, d~5: 93 us vs 103 us for k_adjacency_t (profiles/r2_pair_adjacency.md)
#ifndef FGNN_LAST_HOP_SEPARATE_DEFAULT
#define FGNN_LAST_HOP_SEPARATE_DEFAULT true      // measured: 312 -> 307 us/step (high-occupancy gather + streaming readout)
#endif

static int pad_hidden(int H) { return H <= 16 ? 16 : H <= 32 ? 32 : H <= 64 ? 64 : 128; }

// ---- kernel dispatch ------------------------------------------------------------------------
typedef void (*final_kernel_t)(Params);
typedef void (*dense_kernel_t)(const float*, const float*, float*, const float*, int, int);

namespace fgnn {
typedef void (*final_tc_kernel_t)(Params, const uint8_t*);
typedef void (*mini_rollout_kernel_t)(Params, const uint8_t*, int, int, int);
typedef void (*mini_policy_kernel_t)(Params, const uint8_t*);
typedef void (*mini_envstep_kernel_t)(Params, const void*, int, int, int);
#define FGNN_DECL(K, HP) final_kernel_t get_final_k##K##_hp##HP(bool closed); dense_kernel_t get_dense_k##K##_hp##HP(); \
    final_tc_kernel_t get_final_tc_k##K##_hp##HP(bool closed); mini_rollout_kernel_t get_mini_rollout_k##K##_hp##HP(); \
    mini_policy_kernel_t get_mini_policy_k##K##_hp##HP(); mini_envstep_kernel_t get_mini_envstep_k##K##_hp##HP();
#define FGNN_DECL_K(K) FGNN_DECL(K, 16) FGNN_DECL(K, 32) FGNN_DECL(K, 64) FGNN_DECL(K, 128)
FGNN_DECL_K(1) FGNN_DECL_K(2) FGNN_DECL_K(3) FGNN_DECL_K(4)
}

#define FGNN_CASE_HP(K, HP) case HP: return closed_or_dense == 7 ? (void*)get_mini_envstep_k##K##_hp##HP() : closed_or_dense == 6 ? (void*)get_mini_policy_k##K##_hp##HP() \
    : closed_or_dense == 5 ? (void*)get_mini_rollout_k##K##_hp##HP() : closed_or_dense == 2 ? (void*)get_dense_k##K##_hp##HP() \
    : closed_or_dense >= 3 ? (void*)get_final_tc_k##K##_hp##HP(closed_or_dense == 4) : (void*)get_final_k##K##_hp##HP(closed_or_dense == 1);
#define FGNN_CASE_K(K) case K: switch (HP) { FGNN_CASE_HP(K, 16) FGNN_CASE_HP(K, 32) FGNN_CASE_HP(K, 64) default: FGNN_CASE_HP(K, 128) } break;
static void* kernel_lookup(int K, int HP, int closed_or_dense) {
    switch (K) { FGNN_CASE_K(1) FGNN_CASE_K(2) FGNN_CASE_K(3) default: FGNN_CASE_K(4) }
    return nullptr;
}
static final_kernel_t final_kernel(int K, int HP, bool closed) { return (final_kernel_t)kernel_lookup(K, HP, closed ? 1 : 0); }
static dense_kernel_t dense_kernel(int K, int HP) { return (dense_kernel_t)kernel_lookup(K, HP, 2); }
static final_tc_kernel_t final_tc_kernel(int K, int HP, bool closed) { return (final_tc_kernel_t)kernel_lookup(K, HP, closed ? 4 : 3); }
static mini_rollout_kernel_t mini_rollout_kernel(int K, int HP) { return HP <= 64 ? (mini_rollout_kernel_t)kernel_lookup(K, HP, 5) : nullptr; }
static size_t final_smem_bytes(const fgnn_handle* h) {
    WeightLayout wl{F * h->cfg.k, h->HP, h->cfg.n_layers};
    if (h->HP > 64) return (size_t)2 * h->HP * FINAL_THREADS * sizeof(float);
    return ((size_t)wl.total() + (size_t)h->HP * FINAL_THREADS) * sizeof(float);
}

static inline int blocks_for(int n, int threads) { return (n + threads - 1) / threads; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a process-wide property of the FUNCTION: handles with different shapes
// (layers, stage depth) share the kernels, so the limit is only ever raised.
static cudaError_t raise_smem_limit(const void* func, size_t bytes) {
    static std::map<std::pair<int, const void*>, size_t> limit;      // (per device: the attribute lives in the device's context)
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& cur = limit[std::make_pair(dev, func)];
    if (bytes <= cur) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

// Launch one of the closed-loop step kernels, with the programmatic-serialization attribute when the handle asks for
// it (the kernel may then become resident while its predecessor drains; it starts with pdl_prologue()).
template <typename... KArgs, typename... Args>
static void launch_step(const fgnn_handle* h, void (*kernel)(KArgs...), dim3 grid, int block, size_t smem, cudaStream_t st,
                        Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
static void launch_step(const fgnn_handle* h, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st,
                        Args... args) {
    launch_step(h, kernel, dim3((unsigned)grid), block, smem, st, args...);
}

// ---- ABI --------------------------------------------------------------------------------------
extern "C" const char* fgnn_last_error(void) { return g_err.c_str(); }
extern "C" int fgnn_version(void) { return 100; }

extern "C" int fgnn_create(const fgnn_config* cfg, fgnn_handle** out) {
    if (!cfg || !out) return fail("fgnn_create: null argument");
    if (cfg->n_agents < 1 || cfg->n_episodes < 1) return fail("fgnn_create: n_agents and n_episodes must be >= 1");
    if (cfg->k < 1 || cfg->k > KMAX) return fail("fgnn_create: k must be in 1..4");
    if (cfg->n_states != F) return fail("fgnn_create: n_states must be 6 (FlockingRelative features)");
    if (cfg->n_actions != 2) return fail("fgnn_create: n_actions must be 2");
    if (cfg->hidden < 1 || cfg->hidden > 128) return fail("fgnn_create: hidden must be in 1..128");
    if (cfg->n_layers < 1 || cfg->n_layers > LMAX) return fail("fgnn_create: n_layers must be in 1..4");
    if (cfg->readout_mode < 0 || cfg->readout_mode > 2) return fail("fgnn_create: readout_mode must be 0 (auto), 1 (FFMA) or 2 (tensor cores)");
    if (!(cfg->comm_radius > 0.0) || !(cfg->dt > 0.0)) return fail("fgnn_create: comm_radius and dt must be > 0");
    const long long M64 = (long long)cfg->n_agents * cfg->n_episodes;
    if (M64 > (1ll << 30) - 1) return fail("fgnn_create: more than 2^30-1 agents on one device");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail("fgnn_create: no such CUDA device");
    CK(cudaSetDevice(cfg->device));

    fgnn_handle* h = new fgnn_handle();
    h->cfg = *cfg;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    h->HP = pad_hidden(cfg->hidden);
    Params& p = h->p;
    memset(&p, 0, sizeof p);
    p.M = (int)M64; p.N = cfg->n_agents; p.B = cfg->n_episodes; p.K = cfg->k; p.L = cfg->n_layers;
    h->sharded = cfg->shard_count > 0;
    if (h->sharded) {
        if (cfg->n_episodes != 1) return fail("fgnn_create: sharding needs n_episodes == 1");
        if (cfg->shard_lo < 0 || (long long)cfg->shard_lo + cfg->shard_count > M64) return fail("fgnn_create: shard range outside the agents");
        if (cfg->ghost_capacity < 0) return fail("fgnn_create: ghost_capacity < 0");
    }
    const long long n_local = h->sharded ? (long long)cfg->shard_count + cfg->ghost_capacity : (long long)cfg->n_agents;
    int G = cfg->grid_dim > 0 ? cfg->grid_dim : (int)std::ceil(std::sqrt((double)n_local));
    if (G < 3) G = 3;
    int Gy = cfg->grid_dim_y > 0 ? cfg->grid_dim_y : G;
    if (Gy < 3) Gy = 3;
    if ((long long)G * Gy * cfg->n_episodes > (1ll << 30)) return fail("fgnn_create: cell grid too large");
    p.G = G; p.Gy = Gy; p.C = G * Gy * cfg->n_episodes;
    p.a_lo = h->sharded ? cfg->shard_lo : 0;
    p.n_own = h->sharded ? cfg->shard_count : (int)M64;
    p.pool_cap = h->sharded ? (int)n_local : (int)M64;
    h->launch_pool = p.pool_cap;
    p.mean_pooling = cfg->mean_pooling; p.half_accel = cfg->half_accel_term;
    const long long cap_per = cfg->edge_capacity > 0 ? cfg->edge_capacity : 48;
    long long cap = cap_per * M64;
    if (cap > 0xfffffff0ll) cap = 0xfffffff0ll;
    if (cap < 1024) cap = 1024;
    p.nnz_cap = (unsigned)cap;
    h->adj_stage = (int)(cap_per < 8 ? 8 : cap_per > 64 ? 64 : cap_per);
    {   // warp-staged adjacency: pays when a warp's three row ranges fit its tile (moderate degree); FGNN_ADJ_MODE=0/1 overrides
        const char* mode = getenv("FGNN_ADJ_MODE");
        // default: on when the expected degree is moderate (edge-capacity hint <= 48 per agent, which includes the
        // automatic capacity); at larger radii a warp's row ranges outgrow the tile and the per-lane path with its
        // deeper stage is faster (measured, C4 sweep: R >= 1.5 at density 1.6)
        h->adj_warp_staged = mode ? atoi(mode) != 0 : (FGNN_ADJ_DEFAULT_WS && cap_per <= 48);
        const char* pc = getenv("FGNN_POLICY_CHUNKS");
        h->policy_chunks = pc ? atoi(pc) : FGNN_POLICY_CHUNKS_DEFAULT;
        const char* pp = getenv("FGNN_POLICY_PIPE");
        h->policy_pipe = pp ? atoi(pp) : FGNN_POLICY_PIPE_DEFAULT;
        const char* pd = getenv("FGNN_PDL");
        h->pdl = pd ? atoi(pd) != 0 : FGNN_PDL_DEFAULT;
        const char* tp = getenv("FGNN_SCAN_TWO_PASS");
        h->scan_two_pass = tp ? atoi(tp) != 0 : FGNN_SCAN_TWO_PASS_DEFAULT;
        const char* sb = getenv("FGNN_SUMS_IN_BIN");
        p.sums_in_bin = (h->scan_two_pass && sb && atoi(sb) != 0) ? 1 : 0;   // opt-in; measured SLOWER at N=1M (299 vs 288 us/step): the warp-match +
                                                                          // second atomic cost the final kernel more than the k_scan_sums launch, and k_scan then pays the cold read
        const char* lh = getenv("FGNN_LAST_HOP_SEPARATE");
        h->last_hop_separate = lh ? atoi(lh) != 0 : FGNN_LAST_HOP_SEPARATE_DEFAULT;
        CK(raise_smem_limit((const void*)k_adjacency_t<true>, (size_t)(((size_t)(WS_STAGE + 1) * ADJ_THREADS * sizeof(int) + WS_SMEM))));
        CK(raise_smem_limit((const void*)k_adjacency_t<false>, (size_t)(((size_t)64 * ADJ_THREADS * sizeof(int)))));
    }
    p.inv_cell = 1.0 / (cfg->comm_radius * (1.0 + 1.0 / 1048576.0));
    p.R2 = cfg->comm_radius * cfg->comm_radius;
    p.dt = cfg->dt;
    p.gain = cfg->action_scalar;
    p.n_tiles = blocks_for(p.C + 1, SCAN_TILE);
    {   // warp-tiled adjacency (FGNN_STEP_MODE=0 selects k_adjacency_t of round 1)
        const char* sm = getenv("FGNN_STEP_MODE");
        // default: on for large flocks of moderate expected degree (the staged path lists at most PR_LIST neighbours per agent).
        // Its chain of dependent steps per warp (cell table -> TMA -> convert -> filter -> list -> features) is longer than
        // k_adjacency_t's, and the warps that straddle a grid row take the slow per-lane path: below a few waves of blocks that
        // latency is the kernel's time (26 us at N = 10k .. 100k against ~20 us), above it the lower instruction count wins
        // (89.5 against 102 us at N = 1M).  FGNN_STEP_MODE=0/1 overrides.
        h->pair_mode = sm ? atoi(sm) != 0 : (FGNN_STEP_MODE_DEFAULT != 0 && G >= 400 && cap_per <= 48);
        memset(&h->geo, 0, sizeof h->geo);
        PairGeom& ge = h->geo;
        // fp32 pre-filter on warp-relative coordinates (|coordinate| <= E).  With u = 2^-24: every coordinate carries
        // u E, a difference u (2 E + |d|), so for r2 <= 4 R^2 the fp32 r2 is within u (16 R E + 24 R^2) of the float64
        // value; twice that is the margin.  Pairs inside the margin take the float64 test.
        const double cell = 1.0 / p.inv_cell, R = cfg->comm_radius;
        const double E = (PR_EXT + 3) * cell;
        const double margin = std::ldexp(16.0 * R * E + 24.0 * R * R, -23);
        ge.far32 = (float)E;
        ge.me32 = (float)((PR_EXT + 0.5) * cell);
        ge.lo32 = std::nextafterf((float)(p.R2 - margin), -INFINITY);
        ge.hi32 = std::nextafterf(std::nextafterf((float)(p.R2 + margin), INFINITY), INFINITY);
        auto fastdiv = [](unsigned d) {
            FastDiv f;
            f.l = 0;
            while ((1ull << f.l) < d) ++f.l;
            f.m = (unsigned)(((1ull << 32) * ((1ull << f.l) - d)) / d + 1);
            return f;
        };
        ge.divG = fastdiv((unsigned)G);
        ge.divGy = fastdiv((unsigned)Gy);
        for (int d = 0; d < 64; ++d) ge.sinvtab[d] = cfg->mean_pooling ? (float)(1.0 / (double)(d > 0 ? d : 1)) : 1.0f;
        if (h->pair_mode) {
            const char* mb = getenv("FGNN_PR_MINB");
            h->pair_minb = mb ? atoi(mb) : FGNN_PR_MINBLOCKS;
            CK(raise_smem_limit((const void*)k_pair_adjacency<5>, (size_t)(pair_adjacency_smem())));
            CK(raise_smem_limit((const void*)k_pair_adjacency<6>, (size_t)(pair_adjacency_smem())));
            CK(raise_smem_limit((const void*)k_pair_adjacency<8>, (size_t)(pair_adjacency_smem())));
        }
    }

    const size_t M = p.M, K = p.K;
    int rc = 0;
    rc |= dalloc(h, &p.t, 1);
    rc |= dalloc(h, &p.state, M);
    rc |= dalloc(h, &p.cell_of, M);
    rc |= dalloc(h, &p.cell_count, (size_t)p.C + 1);
    rc |= dalloc(h, &p.cell_start, (size_t)p.C + 1);
    rc |= dalloc(h, &p.tmp_id, M);
    rc |= dalloc(h, &p.sorted_id, M);
    rc |= dalloc(h, &p.sorted_state, M);
    rc |= dalloc(h, &p.sorted_cell, M);
    rc |= dalloc(h, &p.tile_status, (size_t)p.n_tiles);
    rc |= dalloc(h, &p.tile_counter, 1);
    rc |= dalloc(h, &p.xhist, K * M * ROW);
    rc |= dalloc(h, &p.sinv, K * M);
    rc |= dalloc(h, &p.row_start, K * M);
    rc |= dalloc(h, &p.deg, K * M);
    rc |= dalloc(h, &p.cols, K * (size_t)p.nnz_cap, false);
    rc |= dalloc(h, &p.ell, K * M * ELLW);
    rc |= dalloc(h, &p.nnz_cursor, K);
    rc |= dalloc(h, &p.edge_total, K);
    rc |= dalloc(h, &p.overflow, 1);
    rc |= dalloc(h, &p.zbuf, K * M * ROW);
    rc |= dalloc(h, &p.ybuf, 2 * K * M * ROW);
    rc |= dalloc(h, &p.action, M * 2);
    rc |= dalloc(h, &p.racc, (size_t)RSLOTS * p.B * 4);
    // one row per block of the largest grid that integrates: the final kernel of a sharded handle walks pool_cap slots
    rc |= dalloc(h, &p.racc_part, (size_t)(blocks_for(p.M > p.pool_cap ? p.M : p.pool_cap, FINAL_THREADS) + 1) * 4);
    rc |= dalloc(h, &p.n_partials, 1);
    rc |= dalloc(h, &p.reward, (size_t)p.B);
    rc |= dalloc(h, &p.reward_pending, 1);
    rc |= dalloc(h, &p.log_index, 1);
    if (getenv("FGNN_MINI_CLOCK")) rc |= dalloc(h, &p.mini_clock, 4);
    if (h->sharded) {
        rc |= dalloc(h, &h->d_own, (size_t)p.pool_cap);
        rc |= dalloc(h, &h->d_ghost, (size_t)p.pool_cap);
        rc |= dalloc(h, &h->d_free, (size_t)p.pool_cap);
        rc |= dalloc(h, &h->d_fuse, 1);
        memset(&h->fuse_host, 0, sizeof h->fuse_host);
        rc |= dalloc(h, &h->d_counts, 4);
        rc |= dalloc(h, &h->d_xminmax, (size_t)2 * (blocks_for(p.pool_cap, FINAL_THREADS) + 1));
        rc |= dalloc(h, &h->d_shift, 1);
        rc |= dalloc(h, &h->d_safe, 2);
        rc |= dalloc(h, &p.n_ghost_snap, 1);
        const char* sf = getenv("FGNN_SHARD_FOLD");
        if (sf) h->shard_fold = atoi(sf) != 0;
        p.own = h->d_own;
        p.n_own_d = h->d_counts + 0;
        p.ghost = h->d_ghost;
        p.n_ghost_d = h->d_counts + 2;
        memset(&h->ctl, 0, sizeof h->ctl);
        h->ctl.own = h->d_own; h->ctl.n_own = h->d_counts + 0;
        h->ctl.free_slots = h->d_free; h->ctl.n_free = h->d_counts + 1;
        h->ctl.ghost = h->d_ghost; h->ctl.n_ghost = h->d_counts + 2;
        h->ctl.counter = h->d_counts + 3;
        h->ctl.xminmax = h->d_xminmax;
        h->ctl.shift = h->d_shift;
        h->ctl.safe = h->d_safe;
    }
    rc |= dalloc(h, &h->d_u_in, (M > (size_t)p.pool_cap ? M : (size_t)p.pool_cap) * 2 * 2);   // fp32 or float64 actions
    rc |= dalloc(h, &h->d_staging, K * M * F > (size_t)M * 4 * 2 ? K * M * F : (size_t)M * 4 * 2);
    WeightLayout wl{F * p.K, h->HP, p.L};
    h->weights_floats = wl.total();
    h->w_host.assign(h->weights_floats, 0.f);
    rc |= dalloc(h, &h->d_weights, h->weights_floats);
    if (rc) { fgnn_destroy(h); return 1; }
    p.weights = h->d_weights;
    p.reward_log = nullptr;

    // launch geometry of the fused final kernel: persistent grid sized to the SM count x occupancy
    h->final_smem = final_smem_bytes(h);
    for (int closed = 0; closed < 2; ++closed) {
        final_kernel_t fk = final_kernel(p.K, h->HP, closed != 0);
        CK(raise_smem_limit((const void*)fk, (size_t)(h->final_smem)));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fk, FINAL_THREADS, h->final_smem));
        if (occ < 1) occ = 1;
        int grid = h->sm_count * occ;
        int tiles = blocks_for(h->sharded ? p.pool_cap : p.n_own, FINAL_THREADS);
        if (grid > tiles) grid = tiles;
        (closed ? h->final_grid_closed : h->final_grid_open) = grid;
    }
    // tensor-core readout: HP <= 64 (operands must fit shared memory); readout_mode 1 forces FFMA
    h->raw_w.resize(p.L + 1);
    h->raw_b.resize(p.L + 1);
    if (cfg->readout_mode == 2 && h->HP > 64) { fgnn_destroy(h); return fail("fgnn_create: tensor-core readout needs hidden <= 64"); }
    h->use_tc = (cfg->readout_mode != 1) && h->HP <= 64;
    if (h->use_tc) {
        TcLayout tl;
        tl.K0 = TcLayout::pad8(F * p.K); tl.HP = h->HP; tl.L = p.L;
        const int KA = tl.K0 > tl.HP ? tl.K0 : tl.HP;
        h->tc_host.assign(tl.total_bytes(), 0);
        if (dalloc(h, &h->d_tc_weights, (size_t)tl.total_bytes())) { fgnn_destroy(h); return 1; }
        h->tc_smem = (size_t)tl.total_bytes() + (size_t)2 * 128 * KA * 4 + 16;
        if (h->tc_smem > 227 * 1024) {
            if (cfg->readout_mode == 2) { fgnn_destroy(h); return fail("fgnn_create: tensor-core readout operands exceed shared memory"); }
            h->use_tc = false;
        }
    }
    if (h->use_tc) {
        for (int closed = 0; closed < 2; ++closed) {
            final_tc_kernel_t fk = final_tc_kernel(p.K, h->HP, closed != 0);
            CK(raise_smem_limit((const void*)fk, (size_t)(h->tc_smem)));
            // resident CTAs per SM from shared memory, registers and TMEM columns (the occupancy API does not
            // account for TMEM and was observed to report 0 for this kernel)
            cudaFuncAttributes fa;
            CK(cudaFuncGetAttributes(&fa, (const void*)fk));
            int occ = (int)((size_t)prop.sharedMemPerMultiprocessor / (h->tc_smem + 1024));
            const int occ_reg = 65536 / ((fa.numRegs > 0 ? ((fa.numRegs + 7) & ~7) : 128) * FINAL_THREADS);
            if (occ > occ_reg) occ = occ_reg;
            if (occ > 16) occ = 16;
            if (occ < 1) occ = 1;
            const int max_by_tmem = 512 / tc_tmem_cols(h->HP);
            if (occ > max_by_tmem) occ = max_by_tmem;
            int grid = h->sm_count * occ;
            int tiles = blocks_for(h->sharded ? p.pool_cap : p.n_own, FINAL_THREADS);
            if (grid > tiles) grid = tiles;
            (closed ? h->tc_grid_closed : h->tc_grid_open) = grid;
            if (getenv("FGNN_DEBUG")) fprintf(stderr, "[fgnn] final_tc closed=%d regs=%d smem=%zu occ=%d grid=%d tiles=%d sharded=%d\n", closed, fa.numRegs, h->tc_smem, occ, grid, tiles, (int)h->sharded);
        }
    }
    {   // single-CTA path for small flocks (FGNN_MINI=0: the general kernels)
        const char* mn = getenv("FGNN_MINI");
        if ((!mn || atoi(mn) != 0) && !h->sharded && p.M <= 128 && p.C <= MINI_MAX_CELLS && h->use_tc) {
            h->mini_kernel = mini_rollout_kernel(p.K, h->HP);
            h->mini_adj_off = (int)((h->tc_smem + 127) & ~(size_t)127);
            h->mini_smem = (size_t)h->mini_adj_off + (size_t)h->adj_stage * ADJ_THREADS * sizeof(int);
            if (h->mini_kernel) {
                // (function attributes are process-wide: size them for the largest stage any handle may ask for, 64 ids per thread)
                const size_t stage_max = (size_t)64 * ADJ_THREADS * sizeof(int);
                CK(raise_smem_limit((const void*)h->mini_kernel, (size_t)((h->mini_adj_off + stage_max))));
                h->mini_policy = (mini_policy_kernel_t)kernel_lookup(p.K, h->HP, 6);
                h->mini_envstep = (mini_envstep_kernel_t)kernel_lookup(p.K, h->HP, 7);
                CK(raise_smem_limit((const void*)h->mini_policy, (size_t)(h->tc_smem)));
                CK(raise_smem_limit((const void*)h->mini_envstep, (size_t)(stage_max)));
            }
        }
    }
    *out = h;
    return 0;
}

static uint32_t f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static float u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
// cvt.rna.tf32.f32 on the host: round to nearest (ties away) to 10 mantissa bits
static float tf32_rna(float x) {
    uint32_t u = f2u(x);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0x1000u;
    u &= 0xffffe000u;
    return u2f(u);
}

// (re)build the tensor-core weight pack from the raw per-layer weights that have been set so far
static void pack_tc_weights(fgnn_handle* h) {
    const int K = h->cfg.k, H = h->cfg.hidden, L = h->cfg.n_layers, HP = h->HP;
    TcLayout tl;
    tl.K0 = TcLayout::pad8(F * K); tl.HP = HP; tl.L = L;
    std::fill(h->tc_host.begin(), h->tc_host.end(), 0);
    uint8_t* base = h->tc_host.data();
    auto put = [&](int off_hi, int off_lo, int row, int col, int Kdim, float w) {
        const float hi = tf32_rna(w);
        const float lo = tf32_rna(w - hi);
        const int o = TcLayout::canon_off(row, col, Kdim);
        memcpy(base + off_hi + o, &hi, 4);
        memcpy(base + off_lo + o, &lo, 4);
    };
    if (!h->raw_w[0].empty()) {
        for (int g = 0; g < H; ++g)
            for (int f = 0; f < F; ++f)
                for (int k = 0; k < K; ++k) put(tl.off_w0(0), tl.off_w0(1), g, k * F + f, tl.K0, h->raw_w[0][((size_t)g * F + f) * K + k]);
        memcpy(base + tl.off_b(0), h->raw_b[0].data(), H * 4);
    }
    for (int l = 1; l < L; ++l) {
        if (h->raw_w[l].empty()) continue;
        for (int g = 0; g < H; ++g)
            for (int i = 0; i < H; ++i) put(tl.off_wh(l, 0), tl.off_wh(l, 1), g, i, HP, h->raw_w[l][(size_t)g * H + i]);
        memcpy(base + tl.off_b(l), h->raw_b[l].data(), H * 4);
    }
    if (!h->raw_w[L].empty()) {
        float* wl = reinterpret_cast<float*>(base + tl.off_wl());
        for (int a = 0; a < 2; ++a)
            for (int i = 0; i < H; ++i) wl[(i / 2) * 4 + a * 2 + (i & 1)] = h->raw_w[L][(size_t)a * H + i];   // pairs of hidden units (FFMA2)
        memcpy(base + tl.off_bl(), h->raw_b[L].data(), 2 * 4);
    }
}

extern "C" int fgnn_destroy(fgnn_handle* h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    // the communicator is deliberately not destroyed here: ncclCommDestroy waits for the peers and was observed to
    // hang at interpreter shutdown when ranks tear down in different orders; process exit reclaims it
    for (void* q : h->p2p_opened) cudaIpcCloseMemHandle(q);
    for (void* q : h->allocs) cudaFree(q);
    if (h->copy_stream) {
        cudaStreamDestroy(h->copy_stream);
        for (cudaEvent_t e : h->chunk_event) if (e) cudaEventDestroy(e);
    }
    if (h->d_reward_log) cudaFree(h->d_reward_log);
    delete h;
    return 0;
}

extern "C" int fgnn_set_weights(fgnn_handle* h, int32_t layer, const float* W, const float* b, void* stream) {
    if (!h || !W || !b) return fail("fgnn_set_weights: null argument");
    const int K = h->cfg.k, H = h->cfg.hidden, L = h->cfg.n_layers, HP = h->HP;
    if (layer < 0 || layer > L) return fail("fgnn_set_weights: layer out of range");
    cudaStream_t st = (cudaStream_t)stream;
    WeightLayout wl{F * K, HP, L};
    const int out_dim = layer == L ? 2 : H;
    const int in_dim = layer == 0 ? F * K : H;
    std::vector<float> w((size_t)out_dim * in_dim), bb(out_dim);
    CK(cudaMemcpyAsync(w.data(), W, w.size() * sizeof(float), cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(bb.data(), b, bb.size() * sizeof(float), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    h->raw_w[layer] = w;
    h->raw_b[layer] = bb;
    if (h->use_tc) {
        pack_tc_weights(h);
        CK(cudaMemcpyAsync(h->d_tc_weights, h->tc_host.data(), h->tc_host.size(), cudaMemcpyHostToDevice, st));
    }
    float* dst = h->w_host.data();
    if (layer == 0 && L >= 1) {
        // W (H, F, K) -> w0[(k*F + f)][g]
        for (int i = 0; i < F * K * HP; ++i) dst[wl.off_w0() + i] = 0.f;
        for (int g = 0; g < HP; ++g) dst[wl.off_b0() + g] = 0.f;
        for (int g = 0; g < H; ++g) {
            for (int f = 0; f < F; ++f)
                for (int k = 0; k < K; ++k) dst[wl.off_w0() + (k * F + f) * HP + g] = w[((size_t)g * F + f) * K + k];
            dst[wl.off_b0() + g] = bb[g];
        }
    } else if (layer < L) {
        for (int i = 0; i < HP * HP; ++i) dst[wl.off_wh(layer) + i] = 0.f;
        for (int g = 0; g < HP; ++g) dst[wl.off_bh(layer) + g] = 0.f;
        for (int g = 0; g < H; ++g) {
            for (int i = 0; i < H; ++i) dst[wl.off_wh(layer) + i * HP + g] = w[(size_t)g * H + i];
            dst[wl.off_bh(layer) + g] = bb[g];
        }
    } else {
        for (int i = 0; i < HP * 2; ++i) dst[wl.off_wl() + i] = 0.f;
        for (int a = 0; a < 2; ++a) {
            for (int i = 0; i < H; ++i) dst[wl.off_wl() + i * 2 + a] = w[(size_t)a * H + i];
            dst[wl.off_bl() + a] = bb[a];
        }
    }
    CK(cudaMemcpyAsync(h->d_weights, h->w_host.data(), h->weights_floats * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int launch_check(fgnn_handle* h, const char* name) {
    h->launches += 1;
    {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(std::string("launch of ") + name + " -> " + cudaGetErrorString(e));
    }
    if (h->profiling) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        CK(cudaEventRecord(e, h->prof_stream));
        h->prof_events.push_back(e);
        h->prof_names.push_back(name);
    }
    return 0;
}

// bin (if needed) -> scan -> scatter -> canon -> adjacency
static int enqueue_build(fgnn_handle* h, int advance, cudaStream_t st) {
    Params& p = h->p;
    const int gb = blocks_for(h->launch_pool, 256);
    if (!h->binned) {
        if (h->sharded) return fail("sharded engine: the graph is built after fgnn_shard_unpack (state not binned)");
        k_bin<<<gb, 256, 0, st>>>(p);
        if (launch_check(h, "bin")) return 1;
    }
    if (h->scan_two_pass && !p.sums_in_bin) {      // (sums_in_bin: the binning sites accumulated the tile sums already)
        launch_step(h, k_scan_sums, p.n_tiles, SCAN_THREADS, 0, st, p);
        if (launch_check(h, "scan_sums")) return 1;
    }
    launch_step(h, k_scan, p.n_tiles, SCAN_THREADS, 0, st, p, (int)advance,
                (h->scan_two_pass || p.n_tiles <= h->sm_count * 4) ? 1 : 0, h->scan_two_pass ? 1 : 0);
    if (launch_check(h, "scan")) return 1;
    launch_step(h, k_scatter, gb, 256, 0, st, p);
    if (launch_check(h, "scatter")) return 1;
    launch_step(h, k_canon, gb, 256, 0, st, p);
    if (launch_check(h, "canon")) return 1;
    if (h->pair_mode) {
        const int gp = blocks_for(h->launch_pool, PR_THREADS);
        if (h->pair_minb >= 8) launch_step(h, k_pair_adjacency<8>, gp, PR_THREADS, pair_adjacency_smem(), st, p, h->geo);
        else if (h->pair_minb == 6) launch_step(h, k_pair_adjacency<6>, gp, PR_THREADS, pair_adjacency_smem(), st, p, h->geo);
        else launch_step(h, k_pair_adjacency<5>, gp, PR_THREADS, pair_adjacency_smem(), st, p, h->geo);
        if (launch_check(h, "pair_adjacency")) return 1;
        h->binned = false;
        if (advance) h->t_host += 1;
        return 0;
    }
    if (h->adj_warp_staged) {
        const int stage = h->adj_stage < WS_STAGE ? h->adj_stage : WS_STAGE;
        launch_step(h, k_adjacency_t<true>, blocks_for(h->launch_pool, ADJ_THREADS), ADJ_THREADS,
                    (size_t)(stage + 1) * ADJ_THREADS * sizeof(int) + WS_SMEM, st, p, stage);
    } else {
        launch_step(h, k_adjacency_t<false>, blocks_for(h->launch_pool, ADJ_THREADS), ADJ_THREADS,
                    (size_t)h->adj_stage * ADJ_THREADS * sizeof(int), st, p, h->adj_stage);
    }
    if (launch_check(h, "adjacency")) return 1;
    h->binned = false;
    if (advance) h->t_host += 1;
    return 0;
}

template <int NB, bool FIRST>
static void launch_hop(fgnn_handle* h, int j, cudaStream_t st, int tail = 0) {
    Params p = h->p;
    if (tail) p.fuse = h->d_fuse;
    launch_step(h, k_hop<NB, FIRST>, blocks_for(h->launch_pool, 256), 256, 0, st, p, j, tail);
}

// prepare_tail: p2p step of a sharded rank -- block 0 of the LAST hop launch also does k_shard_prepare's work
// skip_last: the caller runs the last hop inside the final kernel (last_hop_done = 0)
static int enqueue_hops(fgnn_handle* h, cudaStream_t st, bool prepare_tail = false, bool skip_last = false) {
    Params& p = h->p;
    for (int j = 0; j + 2 < p.K; ++j) {     // hops 0 .. K-3 ; hop K-2 lives in the final kernel (or below)
        const int nb = p.K - 1 - j;
        if (nb == 3) launch_hop<3, true>(h, j, st);                   // K = 4, hop 0
        else if (nb == 2 && j == 0) launch_hop<2, true>(h, j, st);    // K = 3, hop 0
        else if (nb == 2) launch_hop<2, false>(h, j, st);             // K = 4, hop 1
        else return fail("internal: unexpected hop shape");
        if (launch_check(h, j == 0 ? "hop0" : "hop1")) return 1;
    }
    if (h->last_hop_separate && !skip_last && p.K >= 2) {  // tap K-1 through graph t-(K-2), written to zbuf[K-1] for the final kernel
        const int j = p.K - 2;
        if (j == 0) launch_hop<1, true>(h, j, st, prepare_tail ? 1 : 0);
        else launch_hop<1, false>(h, j, st, prepare_tail ? 1 : 0);
        if (launch_check(h, "hop_last")) return 1;
    }
    return 0;
}

static int enqueue_final(fgnn_handle* h, bool closed, int write_z, cudaStream_t st, bool fuse_pack = false, int tile_lo = 0,
                         int tile_hi = 0, bool hop_inside = false) {
    Params p = h->p;
    p.tile_lo = tile_lo;
    p.tile_hi = tile_hi;
    p.write_z_last = write_z;
    p.last_hop_done = (h->last_hop_separate && !hop_inside && p.K >= 2) ? 1 : 0;
    p.fuse = fuse_pack ? h->d_fuse : nullptr;
    if (h->use_tc) {
        final_tc_kernel_t fk = final_tc_kernel(p.K, h->HP, closed);
        int grid = closed ? h->tc_grid_closed : h->tc_grid_open;
        if (tile_hi > tile_lo && grid > tile_hi - tile_lo) grid = tile_hi - tile_lo;
        launch_step(h, fk, grid, FINAL_THREADS, h->tc_smem, st, p, (const uint8_t*)h->d_tc_weights);
        if (launch_check(h, "final")) return 1;
    } else {
        final_kernel_t fk = final_kernel(p.K, h->HP, closed);
        int grid = closed ? h->final_grid_closed : h->final_grid_open;
        if (tile_hi > tile_lo && grid > tile_hi - tile_lo) grid = tile_hi - tile_lo;
        launch_step(h, fk, grid, FINAL_THREADS, h->final_smem, st, p);
        if (launch_check(h, "final")) return 1;
    }
    if (closed) h->binned = true;
    return 0;
}

static int copy_out(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (!dst) return 0;
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
    return 0;
}

extern "C" int fgnn_set_state(fgnn_handle* h, const double* x, void* stream) {
    if (!h || !x) return fail("fgnn_set_state: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(h->p.state, x, (size_t)h->p.M * sizeof(double4), cudaMemcpyDefault, st));
    if (h->binned) {
        CK(cudaMemsetAsync(h->p.cell_count, 0, ((size_t)h->p.C + 1) * sizeof(int), st));
        CK(cudaMemsetAsync(h->p.tile_status, 0, (size_t)h->p.n_tiles * sizeof(unsigned), st));
        CK(cudaMemsetAsync(h->p.racc, 0, (size_t)RSLOTS * h->p.B * 4 * sizeof(double), st));
        CK(cudaMemsetAsync(h->p.reward_pending, 0, sizeof(int), st));
        h->binned = false;
    }
    return 0;
}

extern "C" int fgnn_reset(fgnn_handle* h, const double* x, void* stream) {
    if (!h || !x) return fail("fgnn_reset: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    Params& p = h->p;
    CK(cudaSetDevice(h->cfg.device));
    const size_t M = p.M, K = p.K;
    CK(cudaMemsetAsync(p.xhist, 0, K * M * ROW * sizeof(float), st));
    CK(cudaMemsetAsync(p.zbuf, 0, K * M * ROW * sizeof(float), st));
    CK(cudaMemsetAsync(p.ybuf, 0, 2 * K * M * ROW * sizeof(float), st));
    CK(cudaMemsetAsync(p.sinv, 0, K * M * sizeof(float), st));
    CK(cudaMemsetAsync(p.row_start, 0, K * M * sizeof(unsigned), st));
    CK(cudaMemsetAsync(p.deg, 0, K * M * sizeof(int), st));
    CK(cudaMemsetAsync(p.nnz_cursor, 0, K * sizeof(unsigned), st));
    CK(cudaMemsetAsync(p.edge_total, 0, K * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(p.t, 0, sizeof(int), st));
    CK(cudaMemsetAsync(p.cell_count, 0, ((size_t)p.C + 1) * sizeof(int), st));
    CK(cudaMemsetAsync(p.tile_status, 0, (size_t)p.n_tiles * sizeof(unsigned), st));
    CK(cudaMemsetAsync(p.tile_counter, 0, sizeof(int), st));
    CK(cudaMemsetAsync(p.racc, 0, (size_t)RSLOTS * p.B * 4 * sizeof(double), st));
    CK(cudaMemsetAsync(p.reward, 0, (size_t)p.B * sizeof(double), st));
    CK(cudaMemsetAsync(p.reward_pending, 0, sizeof(int), st));
    CK(cudaMemsetAsync(p.action, 0, M * 2 * sizeof(float), st));
    h->binned = false;
    h->t_host = 0;
    CK(cudaMemcpyAsync(p.state, x, M * sizeof(double4), cudaMemcpyDefault, st));
    if (h->sharded) {
        // owned agents are binned here; ghosts arrive through fgnn_shard_pack/unpack, then fgnn_build_graph(0)
        k_own_init<<<blocks_for(p.n_own, 256), 256, 0, st>>>(h->d_own, h->d_counts + 0, h->d_counts + 2, h->d_counts + 1, p.a_lo,
                                                             p.n_own);
        if (launch_check(h, "own_init")) return 1;
        CK(cudaMemsetAsync(h->d_shift, 0, sizeof(double), st));
        k_bin<<<blocks_for(p.n_own, 256), 256, 0, st>>>(p);
        if (launch_check(h, "bin")) return 1;
        h->binned = true;
        return 0;
    }
    return enqueue_build(h, 0, st);
}

extern "C" int fgnn_build_graph(fgnn_handle* h, int32_t advance, void* stream) {
    if (!h) return fail("fgnn_build_graph: null handle");
    CK(cudaSetDevice(h->cfg.device));
    return enqueue_build(h, advance, (cudaStream_t)stream);
}

// u: (rows, 2) fp32 or float64 (f64 != 0)
static int integrate_impl(fgnn_handle* h, const void* u, int f64, double* reward_b, cudaStream_t st) {
    Params& p = h->p;
    CK(cudaSetDevice(h->cfg.device));
    if (h->binned) {      // positions were binned already (closed-loop kernel or a previous integrate): start over
        CK(cudaMemsetAsync(p.cell_count, 0, ((size_t)p.C + 1) * sizeof(int), st));
        CK(cudaMemsetAsync(p.tile_status, 0, (size_t)p.n_tiles * sizeof(unsigned), st));
        CK(cudaMemsetAsync(p.racc, 0, (size_t)RSLOTS * p.B * 4 * sizeof(double), st));
    }
    // sharded handle: u has pool_cap rows in owned-list order (rows beyond the owned count are ignored)
    const int rows = h->sharded ? p.pool_cap : p.n_own;
    CK(cudaMemcpyAsync(h->d_u_in, u, (size_t)rows * 2 * (f64 ? sizeof(double) : sizeof(float)), cudaMemcpyDefault, st));
    if (f64) k_integrate<double2><<<blocks_for(rows, 256), 256, 0, st>>>(p, reinterpret_cast<const double2*>(h->d_u_in));
    else k_integrate<float2><<<blocks_for(rows, 256), 256, 0, st>>>(p, reinterpret_cast<const float2*>(h->d_u_in));
    if (launch_check(h, "integrate")) return 1;
    h->binned = true;
    if (reward_b) {
        k_finalize_reward<<<1, 256, 0, st>>>(p);
        if (launch_check(h, "finalize_reward")) return 1;
        return copy_out(reward_b, p.reward, (size_t)p.B * sizeof(double), st);
    }
    return 0;
}

extern "C" int fgnn_integrate(fgnn_handle* h, const float* u, double* reward_b, void* stream) {
    if (!h || !u) return fail("fgnn_integrate: null argument");
    return integrate_impl(h, u, 0, reward_b, (cudaStream_t)stream);
}

extern "C" int fgnn_integrate_f64(fgnn_handle* h, const double* u, double* reward_b, void* stream) {
    if (!h || !u) return fail("fgnn_integrate_f64: null argument");
    return integrate_impl(h, u, 1, reward_b, (cudaStream_t)stream);
}

static int env_step_impl(fgnn_handle* h, const void* u, int f64, double* reward_b, void* stream) {
    if (h && h->sharded) return fail("fgnn_env_step: sharded handle -- integrate, exchange the halo, then build the graph");
    if (!h || !u) return fail("fgnn_env_step: null argument");
    if (h->mini_envstep && !h->binned && !h->profiling) {     // small flock: integrator + cell sort + adjacency in one single-CTA launch
        cudaStream_t st = (cudaStream_t)stream;
        Params& p = h->p;
        CK(cudaSetDevice(h->cfg.device));
        CK(cudaMemcpyAsync(h->d_u_in, u, (size_t)p.n_own * 2 * (f64 ? sizeof(double) : sizeof(float)), cudaMemcpyDefault, st));
        h->mini_envstep<<<1, FINAL_THREADS, (size_t)h->adj_stage * ADJ_THREADS * sizeof(int), st>>>(p, h->d_u_in, f64, 1, h->adj_stage);
        if (launch_check(h, "mini_envstep")) return 1;
        h->t_host += 1;
        return copy_out(reward_b, p.reward, (size_t)p.B * sizeof(double), st);
    }
    if (integrate_impl(h, u, f64, nullptr, (cudaStream_t)stream)) return 1;
    if (enqueue_build(h, 1, (cudaStream_t)stream)) return 1;
    return copy_out(reward_b, h->p.reward, (size_t)h->p.B * sizeof(double), (cudaStream_t)stream);
}

extern "C" int fgnn_env_step(fgnn_handle* h, const float* u, double* reward_b, void* stream) {
    return env_step_impl(h, u, 0, reward_b, stream);
}

extern "C" int fgnn_env_step_f64(fgnn_handle* h, const double* u, double* reward_b, void* stream) {
    return env_step_impl(h, u, 1, reward_b, stream);
}

extern "C" int fgnn_policy(fgnn_handle* h, float* action, void* stream) {
    if (!h) return fail("fgnn_policy: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (h->mini_policy && !h->profiling) {            // small flock: hops + readout in one single-CTA launch
        h->mini_policy<<<1, FINAL_THREADS, h->tc_smem, st>>>(h->p, (const uint8_t*)h->d_tc_weights);
        if (launch_check(h, "mini_policy")) return 1;
        return copy_out(action, h->p.action, (size_t)h->p.M * 2 * sizeof(float), st);
    }
    // Large flock, action wanted in HOST memory: run the readout in chunks of tiles and copy every chunk's actions out
    // on a second stream while the next chunk computes (the copy, 8 MB at N = 1M, costs twice the kernel).
    // Pipelined form (policy_pipe > 0): the LAST hop runs inside each chunk's readout (the final kernel's own gather,
    // same row order, same bits) instead of as a launch over the whole flock in front of the first chunk, the first
    // chunk is one wave of the persistent readout grid and the others policy_pipe waves: the first copy starts one
    // wave after hop 0 and the copies then run back to back (they are slower than the chunks that feed them).
    bool host_chunks = false;
    if (!h->sharded && action && h->policy_chunks > 1 && h->p.M >= (1 << 18)) {
        cudaPointerAttributes attr;
        const bool on_device = cudaPointerGetAttributes(&attr, action) == cudaSuccess &&
                               (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
        cudaGetLastError();
        host_chunks = !on_device;
    }
    const bool pipe = host_chunks && h->policy_pipe > 0 && h->p.K >= 2;
    if (enqueue_hops(h, st, false, pipe)) return 1;
    if (host_chunks) {
        if (!h->copy_stream) {
            CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            for (int c = 0; c <= FGNN_MAX_CHUNKS; ++c) CK(cudaEventCreateWithFlags(&h->chunk_event[c], cudaEventDisableTiming));
        }
        const int n_tiles = blocks_for(h->p.M, FINAL_THREADS);
        int first, per;                                   // tiles of chunk 0 / of every later chunk
        if (pipe) {
            const int wave = h->use_tc ? h->tc_grid_open : h->final_grid_open;
            first = wave;
            per = wave * h->policy_pipe;
            while (first + (long long)per * (FGNN_MAX_CHUNKS - 1) < n_tiles) per += wave;
        } else {
            const int C = h->policy_chunks < FGNN_MAX_CHUNKS ? h->policy_chunks : FGNN_MAX_CHUNKS;
            first = per = blocks_for(n_tiles, C);
        }
        int c = 0;
        for (int lo = 0; lo < n_tiles; ++c) {
            int hi = lo + (c == 0 ? first : per);
            if (hi > n_tiles) hi = n_tiles;
            if (enqueue_final(h, false, 1, st, false, lo, hi, pipe)) return 1;
            CK(cudaEventRecord(h->chunk_event[c], st));
            CK(cudaStreamWaitEvent(h->copy_stream, h->chunk_event[c], 0));
            const size_t a0 = (size_t)lo * FINAL_THREADS, a1 = (size_t)hi * FINAL_THREADS < (size_t)h->p.M ? (size_t)hi * FINAL_THREADS : (size_t)h->p.M;
            CK(cudaMemcpyAsync(action + a0 * 2, h->p.action + a0 * 2, (a1 - a0) * 2 * sizeof(float), cudaMemcpyDeviceToHost,
                               h->copy_stream));
            lo = hi;
        }
        CK(cudaEventRecord(h->chunk_event[FGNN_MAX_CHUNKS], h->copy_stream));
        CK(cudaStreamWaitEvent(st, h->chunk_event[FGNN_MAX_CHUNKS], 0));     // the caller's stream covers the copies
        return 0;
    }
    if (enqueue_final(h, false, 1, st)) return 1;
    if (h->sharded) {           // owned-list order, pool_cap rows
        if (!action) return 0;
        float* tmp = h->d_staging;
        k_gather_owned2<<<blocks_for(h->p.pool_cap, 256), 256, 0, st>>>(h->p, h->p.action, tmp);
        if (launch_check(h, "gather_owned")) return 1;
        return copy_out(action, tmp, (size_t)h->p.pool_cap * 2 * sizeof(float), st);
    }
    return copy_out(action, h->p.action, (size_t)h->p.M * 2 * sizeof(float), st);
}

static int enqueue_closed_step(fgnn_handle* h, cudaStream_t st) {
    if (enqueue_hops(h, st)) return 1;
    if (enqueue_final(h, true, 0, st)) return 1;
    return enqueue_build(h, 1, st);
}

extern "C" int fgnn_step(fgnn_handle* h, float* action, double* reward_b, void* stream) {
    if (!h) return fail("fgnn_step: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (h->sharded) return fail("fgnn_step: sharded handle -- use fgnn_shard_local_step / pack / unpack / build_graph");
    if (h->binned) return fail("fgnn_step: state was integrated but the graph not rebuilt (call fgnn_build_graph)");
    if (h->mini_kernel && !h->profiling) {
        h->mini_kernel<<<1, FINAL_THREADS, h->mini_smem, st>>>(h->p, (const uint8_t*)h->d_tc_weights, 1, h->adj_stage, h->mini_adj_off);
        if (launch_check(h, "mini_step")) return 1;
        h->t_host += 1;
    } else if (enqueue_closed_step(h, st)) return 1;
    if (copy_out(action, h->p.action, (size_t)h->p.M * 2 * sizeof(float), st)) return 1;
    return copy_out(reward_b, h->p.reward, (size_t)h->p.B * sizeof(double), st);
}

extern "C" int fgnn_rollout(fgnn_handle* h, int32_t T, double* reward_bt, void* stream) {
    if (!h || T < 0) return fail("fgnn_rollout: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    Params& p = h->p;
    CK(cudaSetDevice(h->cfg.device));
    if (h->sharded) return fail("fgnn_rollout: sharded handle -- the host drives the per-step halo exchange");
    if (h->binned) return fail("fgnn_rollout: state was integrated but the graph not rebuilt");
    if (T == 0) return 0;
    const bool want_log = reward_bt != nullptr;
    if (T > h->reward_log_cap) {
        if (h->d_reward_log) CK(cudaFree(h->d_reward_log));
        h->d_reward_log = nullptr;
        int cap = T < 256 ? 256 : T;
        CK(cudaMalloc(&h->d_reward_log, (size_t)cap * p.B * sizeof(double)));
        h->reward_log_cap = cap;
        if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
        if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
    }
    if (h->mini_kernel) {           // small flock: T steps inside one CTA, one launch
        Params pp = p;
        pp.reward_log = h->d_reward_log;
        CK(cudaMemsetAsync(p.log_index, 0, sizeof(int), st));
        h->mini_kernel<<<1, FINAL_THREADS, h->mini_smem, st>>>(pp, (const uint8_t*)h->d_tc_weights, T, h->adj_stage, h->mini_adj_off);
        if (launch_check(h, "mini_rollout")) return 1;
        h->t_host += T;
        h->binned = false;
        if (want_log) return copy_out(reward_bt, h->d_reward_log, (size_t)T * p.B * sizeof(double), st);
        return 0;
    }
    if (!h->graph_exec) {
        // capture one closed-loop step; the device step counter makes the same graph valid for every t
        p.reward_log = h->d_reward_log;      // only the captured step logs rewards
        cudaStream_t cs;
        CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        const int64_t l0 = h->launches, t0 = h->t_host;
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_closed_step(h, cs);
        cudaError_t e = cudaStreamEndCapture(cs, &h->graph);
        cudaStreamDestroy(cs);
        p.reward_log = nullptr;
        h->graph_kernels = (int)(h->launches - l0);
        h->launches = l0;
        h->t_host = t0;
        if (rc) return 1;
        if (e != cudaSuccess) return fail(std::string("fgnn_rollout: capture failed: ") + cudaGetErrorString(e));
        CK(cudaGraphInstantiate(&h->graph_exec, h->graph, 0));
    }
    if (h->d_reward_log) CK(cudaMemsetAsync(p.log_index, 0, sizeof(int), st));
    for (int i = 0; i < T; ++i) CK(cudaGraphLaunch(h->graph_exec, st));
    h->launches += (int64_t)h->graph_kernels * T;
    h->t_host += T;
    h->binned = false;
    if (want_log) return copy_out(reward_bt, h->d_reward_log, (size_t)T * p.B * sizeof(double), st);
    return 0;
}

extern "C" int fgnn_actor_forward_dense(fgnn_handle* h, int32_t batch, int32_t n2, const float* ds, const float* gso,
                                        float* out, void* stream) {
    if (!h || !ds || !gso || !out) return fail("fgnn_actor_forward_dense: null argument");
    if (batch < 1 || n2 < 1) return fail("fgnn_actor_forward_dense: empty batch");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    const int K = h->cfg.k;
    dense_kernel_t dk = dense_kernel(K, h->HP);
    WeightLayout wl{F * K, h->HP, h->cfg.n_layers};
    size_t smem = (size_t)K * F * DENSE_MT * sizeof(float) +
                  (h->HP > 64 ? (size_t)2 * h->HP * FINAL_THREADS * sizeof(float)
                              : ((size_t)wl.total() + (size_t)h->HP * FINAL_THREADS) * sizeof(float));
    CK(raise_smem_limit((const void*)dk, (size_t)(smem)));
    dim3 grid(blocks_for(n2, FINAL_THREADS), batch);
    dk<<<grid, FINAL_THREADS, smem, st>>>(ds, gso, out, h->d_weights, h->cfg.n_layers, n2);
    return launch_check(h, "actor_dense");
}

// ---- read-back ----------------------------------------------------------------------------
extern "C" int fgnn_get_state(fgnn_handle* h, double* x, void* stream) {
    if (!h || !x) return fail("fgnn_get_state: null argument");
    CK(cudaSetDevice(h->cfg.device));
    return copy_out(x, h->p.state, (size_t)h->p.M * sizeof(double4), (cudaStream_t)stream);
}

static int age_slot(fgnn_handle* h, int age, int* slot) {
    if (age < 0 || age >= h->cfg.k) return fail("age must be in 0..k-1");
    *slot = slot_of((int)(h->t_host - age), h->cfg.k);
    return 0;
}

extern "C" int fgnn_get_features(fgnn_handle* h, int32_t age, float* out, void* stream) {
    if (!h || !out) return fail("fgnn_get_features: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    int g;
    if (age_slot(h, age, &g)) return 1;
    const int M = h->p.M;
    k_pack_rows<<<blocks_for(M * F, 256), 256, 0, st>>>(h->p.xhist + (size_t)g * M * ROW, h->d_staging, M);
    if (launch_check(h, "pack_rows")) return 1;
    return copy_out(out, h->d_staging, (size_t)M * F * sizeof(float), st);
}

extern "C" int fgnn_get_degrees(fgnn_handle* h, int32_t age, int32_t* out, void* stream) {
    if (!h || !out) return fail("fgnn_get_degrees: null argument");
    CK(cudaSetDevice(h->cfg.device));
    int g;
    if (age_slot(h, age, &g)) return 1;
    return copy_out(out, h->p.deg + (size_t)g * h->p.M, (size_t)h->p.M * sizeof(int), (cudaStream_t)stream);
}

extern "C" int fgnn_get_aggregated(fgnn_handle* h, float* out, void* stream) {
    if (!h || !out) return fail("fgnn_get_aggregated: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    const int M = h->p.M, K = h->cfg.k;
    for (int k = 0; k < K; ++k) {
        const float* rows = k == 0 ? h->p.xhist + (size_t)slot_of((int)h->t_host, K) * M * ROW
                                   : h->p.zbuf + (size_t)k * M * ROW;
        k_pack_rows<<<blocks_for(M * F, 256), 256, 0, st>>>(rows, h->d_staging + (size_t)k * M * F, M);
        if (launch_check(h, "pack_rows")) return 1;
    }
    return copy_out(out, h->d_staging, (size_t)K * M * F * sizeof(float), st);
}

extern "C" int fgnn_get_action(fgnn_handle* h, float* out, void* stream) {
    if (!h || !out) return fail("fgnn_get_action: null argument");
    CK(cudaSetDevice(h->cfg.device));
    return copy_out(out, h->p.action, (size_t)h->p.M * 2 * sizeof(float), (cudaStream_t)stream);
}

extern "C" int fgnn_export_network_dense(fgnn_handle* h, int32_t age, float* out, void* stream) {
    if (!h || !out) return fail("fgnn_export_network_dense: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    int g;
    if (age_slot(h, age, &g)) return 1;
    const size_t n = (size_t)h->p.B * h->p.N * h->p.N;
    cudaPointerAttributes attr;
    const bool on_device = cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    float* d_out = out;
    if (!on_device) CK(cudaMallocAsync((void**)&d_out, n * sizeof(float), st));
    CK(cudaMemsetAsync(d_out, 0, n * sizeof(float), st));
    k_export_dense<<<blocks_for(h->p.M, 256), 256, 0, st>>>(h->p, g, d_out);
    if (launch_check(h, "export_dense")) return 1;
    if (!on_device) {
        CK(cudaMemcpyAsync(out, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaFreeAsync(d_out, st));
    }
    return 0;
}

extern "C" int fgnn_get_csr(fgnn_handle* h, int32_t age, const uint32_t** row_start, const int32_t** deg,
                            const int32_t** cols, const float** src_scale) {
    if (!h) return fail("fgnn_get_csr: null handle");
    int g;
    if (age_slot(h, age, &g)) return 1;
    const size_t M = h->p.M;
    if (h->pair_mode) {
        // the step keeps the ELL head of every row and the CSR tail of long rows: assemble complete rows on demand
        CK(cudaSetDevice(h->cfg.device));
        if (!h->d_csr_cols) {
            if (dalloc(h, &h->d_csr_rows, M) || dalloc(h, &h->d_csr_cols, (size_t)h->p.nnz_cap, false) || dalloc(h, &h->d_csr_cursor, 1)) return 1;
        }
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(h->d_csr_cursor, 0, sizeof(unsigned)));
        k_csr_assemble<<<blocks_for((int)M, 256), 256>>>(h->p, g, h->d_csr_cursor, h->d_csr_rows, h->d_csr_cols);
        if (launch_check(h, "csr_assemble")) return 1;
        CK(cudaDeviceSynchronize());
        if (row_start) *row_start = h->d_csr_rows;
        if (deg) *deg = h->p.deg + g * M;
        if (cols) *cols = h->d_csr_cols;
        if (src_scale) *src_scale = h->p.sinv + g * M;
        return 0;
    }
    if (row_start) *row_start = h->p.row_start + g * M;
    if (deg) *deg = h->p.deg + g * M;
    if (cols) *cols = h->p.cols + (size_t)g * h->p.nnz_cap;
    if (src_scale) *src_scale = h->p.sinv + g * M;
    return 0;
}

extern "C" int fgnn_get_stats(fgnn_handle* h, fgnn_stats* out, void* stream) {
    if (!h || !out) return fail("fgnn_get_stats: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    int t = 0, ovf = 0;
    unsigned cur[KMAX] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(&t, h->p.t, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&ovf, h->p.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cur, h->p.nnz_cursor, h->cfg.k * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out->step = t;
    out->n_edges = cur[slot_of(t, h->cfg.k)];
    if (h->pair_mode) {
        unsigned long long tot[KMAX] = {0, 0, 0, 0};
        CK(cudaMemcpyAsync(tot, h->p.edge_total, h->cfg.k * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        out->n_edges = (int64_t)tot[slot_of(t, h->cfg.k)];
    }
    out->overflow = ovf;
    out->grid_dim = h->p.G;
    out->n_cells = h->p.C;
    out->edge_capacity = h->p.nnz_cap;
    out->n_ghosts = 0;
    if (h->p.mini_clock && getenv("FGNN_MINI_CLOCK")) {       // debugging aid: cycles per stage of the single-CTA kernel so far
        long long ck[4];
        CK(cudaMemcpyAsync(ck, h->p.mini_clock, sizeof ck, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        fprintf(stderr, "[fgnn] mini kernel cycles: hops %lld  readout+integrator %lld  sort %lld  adjacency %lld (over %d steps)\n", ck[0], ck[1],
                ck[2], ck[3], t);
    }
    if (h->sharded) {
        int ng = 0;
        CK(cudaMemcpyAsync(&ng, h->d_counts + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        out->n_ghosts = ng;
    }
    return 0;
}

static int controller_impl(fgnn_handle* h, int32_t centralized, double max_accel, void* u, int f64, void* stream) {
    if (!h || !u) return fail("fgnn_controller: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    Params& p = h->p;
    CK(cudaSetDevice(h->cfg.device));
    if (h->binned) return fail("fgnn_controller: state was integrated but the graph not rebuilt");
    double* vsum = reinterpret_cast<double*>(h->d_staging);      // [B][2], staging is at least M*8 floats
    if ((size_t)p.B * 2 * sizeof(double) > (size_t)p.M * 8 * sizeof(float)) return fail("fgnn_controller: staging too small");
    if (centralized) {
        CK(cudaMemsetAsync(vsum, 0, (size_t)p.B * 2 * sizeof(double), st));
        k_vel_sum<<<blocks_for(p.M, 256), 256, 0, st>>>(p, vsum);
        if (launch_check(h, "vel_sum")) return 1;
    }
    const double R = h->cfg.comm_radius;
    const double cell = 1.0 / p.inv_cell;
    // decentralised: neighbours are inside the 3x3 window; centralised: potential cut-off sqrt(R)
    int window = 1;
    if (centralized) window = (int)std::ceil(std::sqrt(R) / cell - 1e-12);
    if (window < 1) window = 1;
    if (f64)
        k_controller<double2><<<blocks_for(h->launch_pool, 128), 128, 0, st>>>(p, centralized ? 1 : 0, window, R, max_accel * p.gain, vsum,
                                                                              reinterpret_cast<double2*>(h->d_u_in));
    else
        k_controller<float2><<<blocks_for(h->launch_pool, 128), 128, 0, st>>>(p, centralized ? 1 : 0, window, R, max_accel * p.gain, vsum,
                                                                             reinterpret_cast<float2*>(h->d_u_in));
    if (launch_check(h, "controller")) return 1;
    return copy_out(u, h->d_u_in, (size_t)p.M * 2 * (f64 ? sizeof(double) : sizeof(float)), st);
}

extern "C" int fgnn_controller(fgnn_handle* h, int32_t centralized, double max_accel, float* u, void* stream) {
    return controller_impl(h, centralized, max_accel, u, 0, stream);
}

extern "C" int fgnn_controller_f64(fgnn_handle* h, int32_t centralized, double max_accel, double* u, void* stream) {
    return controller_impl(h, centralized, max_accel, u, 1, stream);
}

struct ShardGraph;
static int enqueue_shard_pack(fgnn_handle* h, const double* windows, int64_t window_stride, double* send_buf, int cap,
                              int advance, cudaStream_t st) {
    Params& p = h->p;
    ShardFuse f;
    memset(&f, 0, sizeof f);
    f.ctl = h->ctl; f.windows = windows; f.wstride = window_stride; f.buf = send_buf; f.cap = cap;
    k_shard_prepare<<<1, 256, 0, st>>>(p, f, advance);
    if (launch_check(h, "shard_prepare")) return 1;
    k_shard_pack<<<blocks_for(p.pool_cap, 256), 256, 0, st>>>(p, f);
    if (launch_check(h, "shard_pack")) return 1;
    k_shard_header<<<1, 256, 0, st>>>(h->ctl, send_buf, blocks_for(p.pool_cap, 256));
    return launch_check(h, "shard_header");
}

static int enqueue_shard_unpack(fgnn_handle* h, const double* recv_buf, int cap, cudaStream_t st, long long parity_stride = 0,
                                const int* wait_flags = nullptr, const ShardFuse* flag_fuse = nullptr, int n_pack_blocks = 0) {
    Params& p = h->p;
    k_shard_unpack<<<dim3((unsigned)blocks_for(cap, 256), (unsigned)h->ctl.world), 256, 0, st>>>(p, h->ctl, recv_buf, cap, parity_stride,
                                                                                                  wait_flags, flag_fuse, n_pack_blocks);
    if (launch_check(h, "shard_unpack")) return 1;
    h->binned = true;
    return 0;
}

extern "C" int fgnn_shard_configure(fgnn_handle* h, const double* bounds, int32_t world, int32_t rank, double depth,
                                    double margin, double dshift, int32_t handover_after) {
    if (!h || !h->sharded || !bounds) return fail("fgnn_shard_configure: bad argument");
    if (world < 1 || rank < 0 || rank >= world) return fail("fgnn_shard_configure: bad rank / world");
    CK(cudaSetDevice(h->cfg.device));
    if (!h->d_bounds) {
        void* q = nullptr;
        CK(cudaMalloc(&q, (size_t)(world + 1) * sizeof(double)));
        h->allocs.push_back(q);
        h->d_bounds = reinterpret_cast<double*>(q);
    } else if (world != h->ctl.world) {
        return fail("fgnn_shard_configure: world size cannot change");
    }
    CK(cudaMemcpy(h->d_bounds, bounds, (size_t)(world + 1) * sizeof(double), cudaMemcpyHostToDevice));
    h->ctl.bounds = h->d_bounds;
    h->ctl.world = world; h->ctl.rank = rank;
    h->ctl.depth = depth; h->ctl.margin = margin; h->ctl.dshift = dshift; h->ctl.handover_after = handover_after;
    h->shard_configured = true;
    h->shard_epoch += 1;                     // cached graphs bake the control block: they are re-captured
    return 0;
}

extern "C" int fgnn_shard_local_step(fgnn_handle* h, void* stream) {
    if (!h || !h->sharded) return fail("fgnn_shard_local_step: handle is not sharded");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (h->binned) return fail("fgnn_shard_local_step: graph not rebuilt since the last step");
    if (enqueue_hops(h, st)) return 1;
    return enqueue_final(h, true, 0, st);
}

extern "C" int fgnn_shard_pack(fgnn_handle* h, const double* windows, int64_t window_stride, double* send_buf, int32_t cap,
                               int32_t advance, void* stream) {
    if (!h || !h->sharded || !windows || !send_buf) return fail("fgnn_shard_pack: bad argument");
    if (!h->shard_configured) return fail("fgnn_shard_pack: call fgnn_shard_configure first");
    CK(cudaSetDevice(h->cfg.device));
    return enqueue_shard_pack(h, windows, window_stride, send_buf, cap, advance, (cudaStream_t)stream);
}

extern "C" int fgnn_shard_unpack(fgnn_handle* h, const double* recv_buf, int32_t cap, void* stream) {
    if (!h || !h->sharded || !recv_buf) return fail("fgnn_shard_unpack: bad argument");
    if (!h->shard_configured) return fail("fgnn_shard_unpack: call fgnn_shard_configure first");
    CK(cudaSetDevice(h->cfg.device));
    return enqueue_shard_unpack(h, recv_buf, cap, (cudaStream_t)stream);
}

extern "C" int fgnn_shard_owned(fgnn_handle* h, int32_t* ids, int32_t* count, void* stream) {
    if (!h || !h->sharded || !count) return fail("fgnn_shard_owned: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    int n = 0;
    CK(cudaMemcpyAsync(&n, h->d_counts + 0, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *count = n;
    if (ids && n > 0) {          // the list may contain -1 entries (handed-over slots): the caller filters them
        CK(cudaMemcpyAsync(ids, h->d_own, (size_t)n * sizeof(int), cudaMemcpyDefault, st));
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}

// CUDA-graph replay of the two halves of a sharded step around the host's all-gather:
//   begin = hops + final(closed) + pack          end = unpack + scan/scatter/canon/adjacency (advance)
// Graphs are captured on first use and re-captured when an argument (pointer, size, depth) changes.
struct ShardGraph {
    cudaGraphExec_t exec = nullptr;
    int kernels = 0;
    const void* a = nullptr; const void* b = nullptr;
    long long i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    double d = 0;
};
static ShardGraph* shard_graphs(fgnn_handle* h) {     // three entries, lazily allocated (begin, end, whole step)
    if (!h->shard_graph_store) h->shard_graph_store = new ShardGraph[3];
    return reinterpret_cast<ShardGraph*>(h->shard_graph_store);
}

template <typename F>
static int run_cached_graph(fgnn_handle* h, ShardGraph& g, const void* a, const void* b, long long i0, long long i1,
                            long long i2, long long i3, double d, cudaStream_t st, F enqueue) {
    const bool hit = g.exec && g.a == a && g.b == b && g.i0 == i0 && g.i1 == i1 && g.i2 == i2 && g.i3 == i3 && g.d == d;
    if (!hit) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        cudaStream_t cs;
        CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        const int64_t l0 = h->launches, t0 = h->t_host;
        const bool binned0 = h->binned;
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue(cs);
        cudaError_t e = cudaStreamEndCapture(cs, &graph);
        cudaStreamDestroy(cs);
        g.kernels = (int)(h->launches - l0);
        h->launches = l0;
        h->t_host = t0;
        h->binned = binned0;
        if (rc) return 1;
        if (e != cudaSuccess) return fail(std::string("shard graph capture failed: ") + cudaGetErrorString(e));
        CK(cudaGraphInstantiate(&g.exec, graph, 0));
        cudaGraphDestroy(graph);
        g.a = a; g.b = b; g.i0 = i0; g.i1 = i1; g.i2 = i2; g.i3 = i3; g.d = d;
    }
    CK(cudaGraphLaunch(g.exec, st));
    h->launches += g.kernels;
    return 0;
}

extern "C" int fgnn_shard_step_begin(fgnn_handle* h, const double* windows, int64_t window_stride, double* send_buf,
                                     int32_t cap, void* stream) {
    if (!h || !h->sharded || !windows || !send_buf) return fail("fgnn_shard_step_begin: bad argument");
    if (!h->shard_configured) return fail("fgnn_shard_step_begin: call fgnn_shard_configure first");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (h->binned) return fail("fgnn_shard_step_begin: graph not rebuilt since the last step");
    // the pack step runs inside the closed final kernel; its arguments live in device memory, so the cached
    // graph stays valid when the buffers change
    ShardFuse f;
    memset(&f, 0, sizeof f);
    f.ctl = h->ctl; f.windows = windows; f.wstride = window_stride; f.buf = send_buf; f.cap = cap;
    if (memcmp(&f, &h->fuse_host, sizeof f) != 0) {
        h->fuse_host = f;
        CK(cudaMemcpyAsync(h->d_fuse, &h->fuse_host, sizeof f, cudaMemcpyHostToDevice, st));
    }
    const int final_grid = h->use_tc ? h->tc_grid_closed : h->final_grid_closed;
    ShardGraph& g = shard_graphs(h)[0];
    int rc = run_cached_graph(h, g, nullptr, send_buf, h->shard_epoch, final_grid, 0, 0, 0.0, st, [&](cudaStream_t cs) {
        if (enqueue_hops(h, cs)) return 1;
        k_shard_prepare<<<1, 256, 0, cs>>>(h->p, f, 1);
        if (launch_check(h, "shard_prepare")) return 1;
        if (enqueue_final(h, true, 0, cs, true)) return 1;
        k_shard_header<<<1, 256, 0, cs>>>(h->ctl, send_buf, final_grid);
        return launch_check(h, "shard_header");
    });
    if (rc) return 1;
    h->binned = true;
    return 0;
}

extern "C" int fgnn_shard_step_end(fgnn_handle* h, const double* recv_buf, int32_t cap, void* stream) {
    if (!h || !h->sharded || !recv_buf) return fail("fgnn_shard_step_end: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (!h->binned) return fail("fgnn_shard_step_end: call fgnn_shard_step_begin first");
    ShardGraph& g = shard_graphs(h)[1];
    int rc = run_cached_graph(h, g, recv_buf, nullptr, cap, h->shard_epoch, 0, 0, 0.0, st, [&](cudaStream_t cs) {
        if (enqueue_shard_unpack(h, recv_buf, cap, cs)) return 1;
        return enqueue_build(h, 1, cs);
    });
    if (rc) return 1;
    h->binned = false;
    h->t_host += 1;
    return 0;
}

extern "C" int fgnn_comm_unique_id(void* id_128) {
    if (!id_128) return fail("fgnn_comm_unique_id: null argument");
    if (nccl_load()) return 1;
    return nccl_check(g_nccl.GetUniqueId(reinterpret_cast<NcclId*>(id_128)), "ncclGetUniqueId");
}

extern "C" int fgnn_comm_init(fgnn_handle* h, const void* id_128, int32_t rank, int32_t world) {
    if (!h || !id_128) return fail("fgnn_comm_init: null argument");
    if (!h->sharded) return fail("fgnn_comm_init: handle is not sharded");
    if (world < 1 || rank < 0 || rank >= world) return fail("fgnn_comm_init: bad rank / world");
    if (h->nccl_comm) return fail("fgnn_comm_init: communicator already initialised");
    if (nccl_load()) return 1;
    CK(cudaSetDevice(h->cfg.device));
    NcclId id;
    memcpy(&id, id_128, sizeof id);
    return nccl_check(g_nccl.CommInitRank(&h->nccl_comm, world, id, rank), "ncclCommInitRank");
}

// One closed-loop step of a rank as ONE CUDA graph: hops + final (with the fused halo pack) -> ncclAllGather of the
// fixed-capacity record buffers -> unpack + scan/scatter/canon/adjacency.  windows = the headers of `recv_buf` as the
// previous step left them (same buffer every step: the graph bakes its address).
extern "C" int fgnn_shard_step(fgnn_handle* h, double* send_buf, double* recv_buf, int32_t cap, void* stream) {
    if (!h || !h->sharded || !send_buf || !recv_buf) return fail("fgnn_shard_step: bad argument");
    if (!h->shard_configured) return fail("fgnn_shard_step: call fgnn_shard_configure first");
    if (!h->nccl_comm) return fail("fgnn_shard_step: call fgnn_comm_init first");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (h->binned) return fail("fgnn_shard_step: graph not rebuilt since the last step");
    const int64_t stride = (int64_t)(cap + 1) * SREC;
    ShardFuse f;
    memset(&f, 0, sizeof f);
    f.ctl = h->ctl; f.windows = recv_buf + 1; f.wstride = stride; f.buf = send_buf; f.cap = cap;
    if (memcmp(&f, &h->fuse_host, sizeof f) != 0) {
        h->fuse_host = f;
        CK(cudaMemcpyAsync(h->d_fuse, &h->fuse_host, sizeof f, cudaMemcpyHostToDevice, st));
    }
    if (!h->nccl_warm) {
        // first use: one eager all-gather of the same buffers so that NCCL sets up its channels outside any capture.
        // `send_buf` still holds what the reset-time exchange gathered into `recv_buf`, so this rewrites the same bytes.
        if (nccl_check(g_nccl.AllGather(send_buf, recv_buf, (size_t)stride, 8, h->nccl_comm, st), "ncclAllGather (warm-up)")) return 1;
        CK(cudaStreamSynchronize(st));
        h->nccl_warm = true;
    }
    const int final_grid = h->use_tc ? h->tc_grid_closed : h->final_grid_closed;
    ShardGraph& g = shard_graphs(h)[2];
    int rc = run_cached_graph(h, g, recv_buf, send_buf, h->shard_epoch, final_grid, cap, 0, 0.0, st, [&](cudaStream_t cs) {
        if (enqueue_hops(h, cs)) return 1;
        k_shard_prepare<<<1, 256, 0, cs>>>(h->p, f, 1);
        if (launch_check(h, "shard_prepare")) return 1;
        if (enqueue_final(h, true, 0, cs, true)) return 1;
        k_shard_header<<<1, 256, 0, cs>>>(h->ctl, send_buf, final_grid);
        if (launch_check(h, "shard_header")) return 1;
        if (nccl_check(g_nccl.AllGather(send_buf, recv_buf, (size_t)stride, 8 /* ncclFloat64 */, h->nccl_comm, cs), "ncclAllGather"))
            return 1;
        if (enqueue_shard_unpack(h, recv_buf, cap, cs)) return 1;
        return enqueue_build(h, 1, cs);
    });
    if (rc) return 1;
    h->binned = false;
    h->t_host += 1;
    return 0;
}

// ---- p2p halo transport -------------------------------------------------------------------------------------------
// Every rank owns an inbox [2][world][cap + 1][SREC] (+ flag words [2][world]); the closed final kernel of a peer stores
// its records for this rank straight into slot [t & 1][peer] through a CUDA-IPC mapping (NVLink), k_shard_flag
// publishes header + flag, k_shard_wait spins on the flags, k_shard_unpack reads the inbox.  No collective, no host
// round trip: the whole step is ONE CUDA graph per rank.
static size_t p2p_inbox_doubles(int world, int cap) { return (size_t)2 * world * (cap + 1) * SREC; }

extern "C" int fgnn_p2p_alloc(fgnn_handle* h, int32_t world, int32_t rank, int32_t cap, void* ipc_handle_64, void** local_ptr) {
    if (!h || !h->sharded) return fail("fgnn_p2p_alloc: handle is not sharded");
    if (!h->shard_configured || world != h->ctl.world || rank != h->ctl.rank) return fail("fgnn_p2p_alloc: call fgnn_shard_configure with the same world / rank first");
    if (world < 1 || world > 256 || cap < 1) return fail("fgnn_p2p_alloc: bad world / capacity");
    if (h->p2p_inbox) return fail("fgnn_p2p_alloc: inbox already allocated");
    CK(cudaSetDevice(h->cfg.device));
    h->p2p_bytes = p2p_inbox_doubles(world, cap) * sizeof(double) + (size_t)2 * world * sizeof(int);
    void* q = nullptr;
    CK(cudaMalloc(&q, h->p2p_bytes));                    // its own allocation: an IPC handle names a whole allocation
    CK(cudaMemset(q, 0, h->p2p_bytes));
    h->allocs.push_back(q);
    h->p2p_inbox = reinterpret_cast<double*>(q);
    h->p2p_cap = cap; h->p2p_world = world;
    if (dalloc(h, &h->d_peer_inbox, (size_t)world) || dalloc(h, &h->d_peer_flags, (size_t)world) || dalloc(h, &h->d_dest_count, (size_t)world))
        return 1;
    if (ipc_handle_64) {
        cudaIpcMemHandle_t hd;
        static_assert(sizeof hd == 64, "cudaIpcMemHandle_t is 64 bytes");
        CK(cudaIpcGetMemHandle(&hd, q));
        memcpy(ipc_handle_64, &hd, sizeof hd);
    }
    if (local_ptr) *local_ptr = q;
    return 0;
}

// handles: world x 64 bytes (cudaIpcMemHandle_t of every rank's inbox, in rank order) -- ranks are separate processes; or
// direct_ptrs: world device pointers to the inboxes (ranks that live in this process, e.g. the single-process tests).
extern "C" int fgnn_p2p_connect(fgnn_handle* h, const void* handles, void* const* direct_ptrs) {
    if (!h || !h->p2p_inbox) return fail("fgnn_p2p_connect: call fgnn_p2p_alloc first");
    if (!handles && !direct_ptrs) return fail("fgnn_p2p_connect: need IPC handles or direct pointers");
    CK(cudaSetDevice(h->cfg.device));
    const int world = h->p2p_world, rank = h->ctl.rank;
    std::vector<double*> inbox(world);
    std::vector<int*> flags(world);
    for (int q = 0; q < world; ++q) {
        void* base = nullptr;
        if (q == rank) {
            base = h->p2p_inbox;
        } else if (direct_ptrs) {
            base = direct_ptrs[q];
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, base) == cudaSuccess && attr.type == cudaMemoryTypeDevice && attr.device != h->cfg.device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, h->cfg.device, attr.device));
                if (!can) return fail("fgnn_p2p_connect: no peer access between the devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(std::string("cudaDeviceEnablePeerAccess -> ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
        } else {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, reinterpret_cast<const char*>(handles) + (size_t)q * sizeof hd, sizeof hd);
            CK(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
            h->p2p_opened.push_back(base);
        }
        if (!base) return fail("fgnn_p2p_connect: null inbox pointer");
        inbox[q] = reinterpret_cast<double*>(base);
        flags[q] = reinterpret_cast<int*>(inbox[q] + p2p_inbox_doubles(world, h->p2p_cap));
    }
    // touch every peer's inbox once from this device: the lazily enabled peer access of the IPC mappings (hundreds of ms per
    // peer) is paid here, not inside the first step's wait loops
    for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        double probe = 0.0;
        CK(cudaMemcpy(&probe, inbox[q], sizeof(double), cudaMemcpyDeviceToHost));
    }
    CK(cudaMemcpy(h->d_peer_inbox, inbox.data(), world * sizeof(double*), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_peer_flags, flags.data(), world * sizeof(int*), cudaMemcpyHostToDevice));
    h->p2p_connected = true;
    h->shard_epoch += 1;
    return 0;
}

// After the reset-time exchange: install the gathered buffer (world x (cap + 1) records, what fgnn_shard_unpack consumed) as
// BOTH halves of the inbox, so that the first p2p step finds every rank's x-interval in the headers.
extern "C" int fgnn_p2p_seed(fgnn_handle* h, const double* gathered, void* stream) {
    if (!h || !h->p2p_inbox || !gathered) return fail("fgnn_p2p_seed: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    const size_t half = p2p_inbox_doubles(h->p2p_world, h->p2p_cap) / 2;
    CK(cudaMemcpyAsync(h->p2p_inbox, gathered, half * sizeof(double), cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(h->p2p_inbox + half, gathered, half * sizeof(double), cudaMemcpyDefault, st));
    CK(cudaMemsetAsync(h->p2p_inbox + 2 * half, 0, (size_t)2 * h->p2p_world * sizeof(int), st));
    return 0;
}

static int enqueue_p2p_step(fgnn_handle* h, const ShardFuse& f, int final_grid, long long parity_stride, cudaStream_t cs) {
    // k_shard_prepare's work rides on block 0 of the last hop, k_shard_flag's on block (0, 0) of the unpack: two one-block
    // launches less on the critical path of the step (FGNN_SHARD_FOLD=0: the separate launches).  (A "last block done" epilogue
    // of the final kernel for the flag was tried and dropped: the per-block fence + ticket cost the kernel what the launch had
    // cost, and 36 registers.)
    const bool fold = h->shard_fold && h->last_hop_separate && h->p.K >= 2;
    if (enqueue_hops(h, cs, fold)) return 1;
    if (!fold) {
        k_shard_prepare<<<1, 256, 0, cs>>>(h->p, f, 1);
        if (launch_check(h, "shard_prepare")) return 1;
    }
    if (enqueue_final(h, true, 0, cs, true)) return 1;
    if (!fold) {
        k_shard_flag<<<1, 256, 0, cs>>>(h->p, f, final_grid);
        if (launch_check(h, "shard_flag")) return 1;
    }
    // (the unpack blocks of sender q wait for q's flag themselves; folded: block (0, 0) raises this rank's flags first)
    const int* my_flags = reinterpret_cast<const int*>(h->p2p_inbox + p2p_inbox_doubles(h->p2p_world, h->p2p_cap));
    if (enqueue_shard_unpack(h, h->p2p_inbox, h->p2p_cap, cs, parity_stride, my_flags, fold ? h->d_fuse : nullptr, final_grid)) return 1;
    return enqueue_build(h, 1, cs);
}

// One closed-loop step of a rank as ONE CUDA graph, halo over p2p stores:
//   hops -> final (+ fused pack: records into the peers' inboxes) -> flag -> wait -> unpack -> scan/scatter/canon/adjacency
extern "C" int fgnn_shard_step_p2p(fgnn_handle* h, void* stream) {
    if (!h || !h->sharded) return fail("fgnn_shard_step_p2p: handle is not sharded");
    if (!h->p2p_connected) return fail("fgnn_shard_step_p2p: call fgnn_p2p_alloc / fgnn_p2p_connect / fgnn_p2p_seed first");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (h->binned) return fail("fgnn_shard_step_p2p: graph not rebuilt since the last step");
    ShardFuse f;
    memset(&f, 0, sizeof f);
    f.ctl = h->ctl; f.cap = h->p2p_cap; f.p2p = 1;
    f.peer_inbox = h->d_peer_inbox; f.peer_flags = h->d_peer_flags; f.dest_count = h->d_dest_count;
    if (memcmp(&f, &h->fuse_host, sizeof f) != 0) {
        h->fuse_host = f;
        CK(cudaMemcpyAsync(h->d_fuse, &h->fuse_host, sizeof f, cudaMemcpyHostToDevice, st));
    }
    const int final_grid = h->use_tc ? h->tc_grid_closed : h->final_grid_closed;
    const long long parity_stride = (long long)(p2p_inbox_doubles(h->p2p_world, h->p2p_cap) / 2);
    ShardGraph& g = shard_graphs(h)[2];
    int rc = run_cached_graph(h, g, h->p2p_inbox, h->d_peer_inbox, h->shard_epoch, final_grid, h->p2p_cap, 1, 0.0, st,
                              [&](cudaStream_t cs) { return enqueue_p2p_step(h, f, final_grid, parity_stride, cs); });
    if (rc) return 1;
    h->binned = false;
    h->t_host += 1;
    return 0;
}

// The halo exchange alone over the p2p transport, for steps whose integrator ran on its own (fgnn_integrate with a host
// action: the reference-facing loop).  prepare -> pack (records into the peers' inboxes) -> unpack (+ flags, wait).
extern "C" int fgnn_shard_exchange_p2p(fgnn_handle* h, int32_t advance, void* stream) {
    if (!h || !h->sharded) return fail("fgnn_shard_exchange_p2p: handle is not sharded");
    if (!h->p2p_connected) return fail("fgnn_shard_exchange_p2p: call fgnn_p2p_alloc / fgnn_p2p_connect / fgnn_p2p_seed first");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    ShardFuse f;
    memset(&f, 0, sizeof f);
    f.ctl = h->ctl; f.cap = h->p2p_cap; f.p2p = 1;
    f.peer_inbox = h->d_peer_inbox; f.peer_flags = h->d_peer_flags; f.dest_count = h->d_dest_count;
    if (memcmp(&f, &h->fuse_host, sizeof f) != 0) {
        h->fuse_host = f;
        CK(cudaMemcpyAsync(h->d_fuse, &h->fuse_host, sizeof f, cudaMemcpyHostToDevice, st));
    }
    const Params& p = h->p;
    const int n_blocks = blocks_for(p.pool_cap, 256);
    k_shard_prepare<<<1, 256, 0, st>>>(p, f, advance);
    if (launch_check(h, "shard_prepare")) return 1;
    k_shard_pack<<<n_blocks, 256, 0, st>>>(p, f);
    if (launch_check(h, "shard_pack")) return 1;
    const long long parity_stride = (long long)(p2p_inbox_doubles(h->p2p_world, h->p2p_cap) / 2);
    const int* my_flags = reinterpret_cast<const int*>(h->p2p_inbox + p2p_inbox_doubles(h->p2p_world, h->p2p_cap));
    return enqueue_shard_unpack(h, h->p2p_inbox, h->p2p_cap, st, parity_stride, my_flags, h->d_fuse, n_blocks);
}

// Keeps the stream busy for `ns` nanoseconds: fgnn_profile_step puts it in front of its first event so that the step's
// kernels are already queued when the clock starts (on an idle stream the first interval would also hold the host's
// launch latency: hop 0 read ~20 us too long).
__global__ void k_delay(unsigned long long ns) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < ns);
}

extern "C" int fgnn_profile_step(fgnn_handle* h, int32_t max_kernels, float* ms_out, char* names_out, int32_t* n_out,
                                 void* stream) {
    if (!h || !ms_out || !n_out) return fail("fgnn_profile_step: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (h->binned) return fail("fgnn_profile_step: state was integrated but the graph not rebuilt");
    if (h->sharded && !h->p2p_connected) return fail("fgnn_profile_step: a sharded handle is profiled through its p2p step (fgnn_p2p_connect first)");
    h->prof_events.clear();
    h->prof_names.clear();
    cudaEvent_t e0;
    CK(cudaEventCreate(&e0));
    k_delay<<<1, 1, 0, st>>>(100000ull);
    CK(cudaEventRecord(e0, st));
    h->profiling = true;
    h->prof_stream = st;
    int rc;
    if (h->sharded) {                                   // the p2p step of this rank, kernel by kernel (the peers must step too)
        ShardFuse f;
        memset(&f, 0, sizeof f);
        f.ctl = h->ctl; f.cap = h->p2p_cap; f.p2p = 1;
        f.peer_inbox = h->d_peer_inbox; f.peer_flags = h->d_peer_flags; f.dest_count = h->d_dest_count;
        if (memcmp(&f, &h->fuse_host, sizeof f) != 0) {
            h->fuse_host = f;
            CK(cudaMemcpyAsync(h->d_fuse, &h->fuse_host, sizeof f, cudaMemcpyHostToDevice, st));
        }
        rc = enqueue_p2p_step(h, f, h->use_tc ? h->tc_grid_closed : h->final_grid_closed,
                              (long long)(p2p_inbox_doubles(h->p2p_world, h->p2p_cap) / 2), st);
        if (!rc) { h->binned = false; }
    } else {
        rc = enqueue_closed_step(h, st);
    }
    h->profiling = false;
    if (rc) return 1;
    CK(cudaStreamSynchronize(st));
    const int n = (int)h->prof_events.size();
    *n_out = n < max_kernels ? n : max_kernels;
    cudaEvent_t prev = e0;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, prev, h->prof_events[i]));
        if (i < max_kernels) {
            ms_out[i] = ms;
            if (names_out) {
                strncpy(names_out + 16 * i, h->prof_names[i].c_str(), 15);
                names_out[16 * i + 15] = 0;
            }
        }
        prev = h->prof_events[i];
    }
    CK(cudaEventDestroy(e0));
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->prof_events.clear();
    return 0;
}

// cached CUDA graphs bake the kernel arguments (Params by value): drop them when an argument changes
static void invalidate_graphs(fgnn_handle* h) {
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
    h->shard_epoch += 1;
}

extern "C" int fgnn_set_dt(fgnn_handle* h, double dt) {
    if (!h) return fail("fgnn_set_dt: null handle");
    if (!(dt > 0.0)) return fail("fgnn_set_dt: dt must be > 0");
    if (dt != h->p.dt) {
        CK(cudaSetDevice(h->cfg.device));
        h->p.dt = dt;
        h->cfg.dt = dt;
        invalidate_graphs(h);
    }
    return 0;
}

extern "C" int fgnn_set_agent_mask(fgnn_handle* h, const uint8_t* mask, void* stream) {
    if (!h) return fail("fgnn_set_agent_mask: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->cfg.device));
    if (!mask) {
        if (h->p.amask) { h->p.amask = nullptr; invalidate_graphs(h); }
        return 0;
    }
    if (!h->d_amask && dalloc(h, &h->d_amask, (size_t)h->p.M, false)) return 1;
    CK(cudaMemcpyAsync(h->d_amask, mask, (size_t)h->p.M, cudaMemcpyDefault, st));
    if (h->p.amask != h->d_amask) { h->p.amask = h->d_amask; invalidate_graphs(h); }
    return 0;
}

extern "C" int fgnn_memcpy_sync(void* dst, const void* src, uint64_t bytes, void* stream) {
    if (!dst || !src) return fail("fgnn_memcpy_sync: null argument");
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

extern "C" int64_t fgnn_launch_count(fgnn_handle* h) { return h ? h->launches : 0; }
