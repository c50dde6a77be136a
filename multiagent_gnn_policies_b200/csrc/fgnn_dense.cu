// fgnn_dense.cu -- Actor.forward on DENSE tensors for any aggregation index and any layer widths.
//
// Reference: learner/actor.py:9-43 (layer shapes: kernel (K,1) at layer ind_agg, (1,1) elsewhere) and :45-86 (forward).
// DAGGER builds its actor with ind_agg = 0 and equal hidden widths (learner/gnn_dagger.py:42-46): that case runs in the
// rollout engine's templated kernels (k_actor_dense / k_final*).  Everything else the reference's constructor accepts --
// ind_agg > 0 (the DDPG actor, learner/gnn_ddpg.py:126), unequal hidden widths, n_s != 6 -- takes the three plain kernels
// below.  This is the small-N compatibility surface (dense (K,N,N) operators), not the hot path: one thread per output
// element, fp32 FFMA sums, no shared-memory tiling.
//
//   x = delay_state.permute(0,2,1,3)                       (B,F,K,N)                       actor.py:63-64
//   layer i < ind_agg : x[b,:,k,n] -> W_i x + b_i, tanh    per tap k (kernel (1,1))        actor.py:73-77
//   layer i = ind_agg : x[b,c,k,:] <- x[b,c,k,:] @ delay_gso[b,k]                          actor.py:68-71
//                       out[b,o,n] = b_o + sum_{c,k} W[o,c,k] x[b,c,k,n]   (kernel (K,1))  actor.py:32-38
//   layer i > ind_agg : per agent (one row left)
//   no tanh after the last layer; result viewed as (B,1,n_a,N)                             actor.py:75-82
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/fgnn.h"

namespace fgnn { void set_error(const char* msg); }

namespace {

constexpr int DTHREADS = 128;
constexpr int MAX_LAYERS = 16;

// x[b, c, k, n] = p[b * sb + c * sc + k * sk + n]: the first layer reads delay_state (B,K,F,N) in place (the reference's
// permute), later layers read the contiguous (B,C,rows,N) buffer of the layer before
struct View {
    const float* p;
    long long sb, sc, sk;
};

__global__ void __launch_bounds__(DTHREADS) k_dense_pointwise(View x, const float* __restrict__ W, const float* __restrict__ bias,
                                                              float* __restrict__ out, int B, int Cin, int Cout, int rows, int N,
                                                              int act) {
    const long long total = (long long)B * Cout * rows * N;
    const long long idx = (long long)blockIdx.x * DTHREADS + threadIdx.x;
    if (idx >= total) return;
    const int n = (int)(idx % N);
    const int k = (int)((idx / N) % rows);
    const int o = (int)((idx / ((long long)N * rows)) % Cout);
    const int b = (int)(idx / ((long long)N * rows * Cout));
    const float* xp = x.p + b * x.sb + k * x.sk + n;
    const float* w = W + (size_t)o * Cin;
    float acc = 0.f;
    for (int c = 0; c < Cin; ++c) acc = fmaf(__ldg(w + c), xp[c * x.sc], acc);
    acc += __ldg(bias + o);
    out[idx] = act ? tanhf(acc) : acc;
}

// out[b,c,k,n] = sum_m x[b,c,k,m] * gso[b,k,m,n]      (out contiguous (B,C,K,N))
__global__ void __launch_bounds__(DTHREADS) k_dense_aggregate(View x, const float* __restrict__ gso, float* __restrict__ out, int B,
                                                              int C, int K, int N) {
    const long long total = (long long)B * C * K * N;
    const long long idx = (long long)blockIdx.x * DTHREADS + threadIdx.x;
    if (idx >= total) return;
    const int n = (int)(idx % N);
    const int k = (int)((idx / N) % K);
    const int c = (int)((idx / ((long long)N * K)) % C);
    const int b = (int)(idx / ((long long)N * K * C));
    const float* xp = x.p + b * x.sb + c * x.sc + k * x.sk;                 // row over m (warp-uniform apart from block edges)
    const float* g = gso + ((size_t)b * K + k) * N * N + n;                  // column n of delay_gso[b,k]: coalesced over n
    float acc = 0.f;
    for (int m = 0; m < N; ++m) acc = fmaf(xp[m], __ldg(g + (size_t)m * N), acc);
    out[idx] = acc;
}

// out[b,o,n] = bias[o] + sum_{c,k} W[o,c,k] * y[b,c,k,n]      (y contiguous (B,Cin,K,N), out (B,Cout,1,N))
__global__ void __launch_bounds__(DTHREADS) k_dense_collapse(const float* __restrict__ y, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ out, int B, int Cin,
                                                             int Cout, int K, int N, int act) {
    const long long total = (long long)B * Cout * N;
    const long long idx = (long long)blockIdx.x * DTHREADS + threadIdx.x;
    if (idx >= total) return;
    const int n = (int)(idx % N);
    const int o = (int)((idx / N) % Cout);
    const int b = (int)(idx / ((long long)N * Cout));
    const float* yp = y + (size_t)b * Cin * K * N + n;
    const float* w = W + (size_t)o * Cin * K;
    float acc = 0.f;
    for (int ck = 0; ck < Cin * K; ++ck) acc = fmaf(__ldg(w + ck), yp[(size_t)ck * N], acc);
    acc += __ldg(bias + o);
    out[idx] = act ? tanhf(acc) : acc;
}

int fail(const std::string& msg) {
    fgnn::set_error(msg.c_str());
    return 1;
}

int check_shapes(int32_t batch, int32_t n_agents, int32_t k, int32_t n_layers, const int32_t* widths) {
    if (!widths) return fail("fgnn_actor_forward_general: null widths");
    if (batch < 1 || n_agents < 1 || k < 1 || n_layers < 1 || n_layers > MAX_LAYERS)
        return fail("fgnn_actor_forward_general: batch, n_agents, k >= 1 and 1 <= n_layers <= 16 expected");
    for (int l = 0; l <= n_layers; ++l)
        if (widths[l] < 1) return fail("fgnn_actor_forward_general: layer widths must be positive");
    return 0;
}

long long buffer_floats(int32_t batch, int32_t n_agents, int32_t k, int32_t n_layers, const int32_t* widths) {
    int maxc = 1;
    for (int l = 0; l <= n_layers; ++l) maxc = widths[l] > maxc ? widths[l] : maxc;
    return (long long)batch * maxc * k * n_agents;
}

int blocks_for(long long total) { return (int)((total + DTHREADS - 1) / DTHREADS); }

}  // namespace

extern "C" int64_t fgnn_actor_general_workspace(int32_t batch, int32_t n_agents, int32_t k, int32_t n_layers,
                                                const int32_t* widths) {
    if (check_shapes(batch, n_agents, k, n_layers, widths)) return -1;
    return 2 * buffer_floats(batch, n_agents, k, n_layers, widths) * (int64_t)sizeof(float);
}

extern "C" int fgnn_actor_forward_general(int32_t device, int32_t batch, int32_t n_agents, int32_t k, int32_t n_layers,
                                          const int32_t* widths, int32_t ind_agg, const float* const* W, const float* const* b,
                                          const float* delay_state, const float* delay_gso, float* out, float* workspace,
                                          void* stream) {
    if (check_shapes(batch, n_agents, k, n_layers, widths)) return 1;
    if (!W || !b || !delay_state || !delay_gso || !out || !workspace) return fail("fgnn_actor_forward_general: null argument");
    for (int l = 0; l < n_layers; ++l)
        if (!W[l] || !b[l]) return fail("fgnn_actor_forward_general: null layer parameter");
    // actor.py:82 views the result as (B,1,n_a,N): the K taps must have been collapsed by layer ind_agg (or K = 1)
    const bool aggregates = ind_agg >= 0 && ind_agg < n_layers;
    if (!aggregates && k != 1)
        return fail("fgnn_actor_forward_general: ind_agg outside 0..n_layers-1 leaves K rows (the reference's final view fails too)");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(std::string("fgnn_actor_forward_general: cudaSetDevice -> ") + cudaGetErrorString(e));
    const int B = batch, N = n_agents, K = k;
    const long long per = buffer_floats(batch, n_agents, k, n_layers, widths);
    float* bufs[2] = {workspace, workspace + per};
    for (int l = 0; l <= n_layers; ++l)      // one thread per element of the widest activation: the grid must fit 2^31 - 1 blocks
        if ((long long)B * widths[l] * K * N / DTHREADS >= 0x7fffffffll) return fail("fgnn_actor_forward_general: tensor too large");
    View cur{delay_state, (long long)K * widths[0] * N, (long long)N, (long long)widths[0] * N};     // (B,K,F,N) read as [b,c,k,n]
    int cur_buf = -1;                      // -1: delay_state itself
    int rows = K;
    int C = widths[0];
    for (int l = 0; l < n_layers; ++l) {
        const int Cout = widths[l + 1];
        const bool last = l == n_layers - 1;
        const int act = last ? 0 : 1;
        float* dst;
        int dst_buf;
        if (l == ind_agg) {
            const int a = cur_buf < 0 ? 0 : 1 - cur_buf;                   // free buffer for the aggregated input
            k_dense_aggregate<<<blocks_for((long long)B * C * K * N), DTHREADS, 0, st>>>(cur, delay_gso, bufs[a], B, C, K, N);
            dst_buf = 1 - a;                                                // cur's buffer: free once the aggregation has read it
            dst = last ? out : bufs[dst_buf];
            k_dense_collapse<<<blocks_for((long long)B * Cout * N), DTHREADS, 0, st>>>(bufs[a], W[l], b[l], dst, B, C, Cout, K, N, act);
            rows = 1;
        } else {
            dst_buf = cur_buf < 0 ? 0 : 1 - cur_buf;
            dst = last ? out : bufs[dst_buf];
            k_dense_pointwise<<<blocks_for((long long)B * Cout * rows * N), DTHREADS, 0, st>>>(cur, W[l], b[l], dst, B, C, Cout, rows,
                                                                                              N, act);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(std::string("fgnn_actor_forward_general: launch of layer ") + std::to_string(l) + " -> " +
                                          cudaGetErrorString(e));
        cur = View{dst, (long long)Cout * rows * N, (long long)rows * N, (long long)N};
        cur_buf = dst_buf;
        C = Cout;
    }
    return 0;
}
