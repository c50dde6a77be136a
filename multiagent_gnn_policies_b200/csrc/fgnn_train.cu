// fgnn_train.cu -- DAGGER.gradient_step on the device: readout forward on the
// aggregated features of a batch of stored states, MSE loss against the expert actions, backward, Adam.
//
// Reference: learner/gnn_dagger.py:76-96 (gradient_step), :49 (Adam, torch defaults), learner/actor.py:73-82.
// With ind_agg = 0 (gnn_dagger.py:43) the graph aggregation precedes every trainable layer, so the update
// needs only z = delay_state @ delay_gso (actor.py:70) per stored state -- 6K floats per agent instead of the
// reference's dense (K,N,N) operator -- and no gradient flows through the graph.
//
// Two kernels per step:
//   k_train_fwd_bwd : persistent CTAs walk 64-row tiles.  Activations of every layer stay in shared memory
//                     (row-major, odd stride: conflict-free when lanes differ in the row); parameters are read in
//                     the reference's own conv layout straight from the torch tensors (warp-broadcast __ldg, L1
//                     resident).  Each CTA accumulates its share of dW / db in its own row of `gpart`.
//   k_train_adam    : one thread per parameter: sums the CTA partials in a fixed order (bit-reproducible), applies
//                     torch.optim.Adam's update in place to param / exp_avg / exp_avg_sq, finalises the loss.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <string>

#include "../../include/fgnn.h"

namespace {

constexpr int TT = 64;            // rows (agents of stored states) per tile
constexpr int TTHREADS = 256;
constexpr int TLAYERS = 5;        // hidden layers <= 4, plus the output layer
constexpr int TF = 6;             // features per tap
constexpr int TA = 2;             // actions

struct TrainParams {
    int R, N, K, in0, H, L, P;
    int S0, SH;                   // shared-memory row strides of the input tile / hidden activations (odd)
    const float* z;               // (B, K, N, 6)
    const float* y;               // (B, 2, N)
    const float* W[TLAYERS];      // reference layout: W_l (out, in) row-major, layer 0 columns c = f*K + k
    const float* b[TLAYERS];
    int off[TLAYERS + 1];         // packed gradient vector: [W_0][b_0][W_1][b_1]...; off[l] = start of W_l
    float* gpart;                 // [grid][P]
    float* lpart;                 // [grid] sum of squared errors
    float dscale;                 // 2 / (R * 2): d loss / d out = dscale * (out - y)
};

struct AdamParams {
    float* p[2 * TLAYERS];        // W_0, b_0, W_1, b_1, ...
    float* m[2 * TLAYERS];
    float* v[2 * TLAYERS];
    int start[2 * TLAYERS + 1];   // packed offsets of the 2(L+1) tensors
    int n_tensors;
    float one_minus_beta1, beta2, one_minus_beta2, step_size, bc2_sqrt, eps;
    int apply;                    // 0: gradients / loss only
    float inv_count;              // 1 / (R * 2)
    float* grads_out;             // [P] or null
    float* loss_out;              // device scalar
};

__device__ __forceinline__ int in_dim(const TrainParams& tp, int l) { return l == 0 ? tp.in0 : tp.H; }
__device__ __forceinline__ int out_dim(const TrainParams& tp, int l) { return l == tp.L ? TA : tp.H; }

__global__ void __launch_bounds__(TTHREADS) k_train_fwd_bwd(TrainParams tp) {
    extern __shared__ float sm[];
    // [act_0: TT x S0][act_1..act_L: TT x SH each][dA: TT x SH][dB: TT x SH]
    float* act0 = sm;
    float* acth = sm + TT * tp.S0;                         // act_l = acth + (l-1) * TT * SH, l = 1..L
    float* dA = acth + (size_t)tp.L * TT * tp.SH;
    float* dB = dA + TT * tp.SH;
    __shared__ float s_red[TTHREADS / 32];
    const int tid = threadIdx.x;
    const int r = tid & (TT - 1), q = tid / TT;            // row of the tile, output-column phase (0..3)
    constexpr int QN = TTHREADS / TT;

    float* gp = tp.gpart + (size_t)blockIdx.x * tp.P;
    for (int i = tid; i < tp.P; i += TTHREADS) gp[i] = 0.f;
    __syncthreads();

    float sq = 0.f;
    const int n_tiles = (tp.R + TT - 1) / TT;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * TT;
        // ---- input tile: act_0[r][f*K + k] = z[b][k][n][f], row = b*N + n (coalesced over (n, f) per tap) ----
        for (int idx = tid; idx < tp.K * TT * TF; idx += TTHREADS) {
            const int k = idx / (TT * TF), rem = idx - k * (TT * TF);
            const int rr = rem / TF, f = rem - rr * TF;
            const int row = row0 + rr;
            float v = 0.f;
            if (row < tp.R) {
                const int bb = row / tp.N, n = row - bb * tp.N;
                v = __ldg(tp.z + (((size_t)bb * tp.K + k) * tp.N + n) * TF + f);
            }
            act0[rr * tp.S0 + f * tp.K + k] = v;
        }
        __syncthreads();
        // ---- forward: hidden layers ----
        for (int l = 0; l < tp.L; ++l) {
            const int nin = in_dim(tp, l);
            const float* ain = l == 0 ? act0 : acth + (size_t)(l - 1) * TT * tp.SH;
            const int sin = l == 0 ? tp.S0 : tp.SH;
            float* aout = acth + (size_t)l * TT * tp.SH;
            const float* __restrict__ W = tp.W[l];
            const float* __restrict__ bias = tp.b[l];
            const float* arow = ain + r * sin;
            for (int g = q; g < tp.H; g += QN) {
                float acc = __ldg(bias + g);
                const float* wrow = W + (size_t)g * nin;
                for (int c = 0; c < nin; ++c) acc = fmaf(__ldg(wrow + c), arow[c], acc);
                aout[r * tp.SH + g] = tanhf(acc);
            }
            __syncthreads();
        }
        // ---- output layer + loss + d(out): thread (r, q < 2) owns out[r][q] ----
        {
            const int l = tp.L;
            const int nin = in_dim(tp, l);
            const float* ain = l == 0 ? act0 : acth + (size_t)(l - 1) * TT * tp.SH;
            const int sin = l == 0 ? tp.S0 : tp.SH;
            if (q < TA) {
                float acc = __ldg(tp.b[l] + q);
                const float* wrow = tp.W[l] + (size_t)q * nin;
                const float* arow = ain + r * sin;
                for (int c = 0; c < nin; ++c) acc = fmaf(__ldg(wrow + c), arow[c], acc);
                const int row = row0 + r;
                float d = 0.f;
                if (row < tp.R) {
                    const int bb = row / tp.N, n = row - bb * tp.N;
                    const float diff = acc - __ldg(tp.y + ((size_t)bb * TA + q) * tp.N + n);
                    sq = fmaf(diff, diff, sq);
                    d = tp.dscale * diff;
                }
                dA[r * tp.SH + q] = d;
            }
            __syncthreads();
        }
        // ---- backward ----
        float* dcur = dA;
        float* dnext = dB;
        for (int l = tp.L; l >= 0; --l) {
            const int nin = in_dim(tp, l), nout = out_dim(tp, l);
            const float* ain = l == 0 ? act0 : acth + (size_t)(l - 1) * TT * tp.SH;
            const int sin = l == 0 ? tp.S0 : tp.SH;
            // dW_l[g][c] += sum_r d[r][g] * a_l[r][c]
            float* gw = gp + tp.off[l];
            for (int o = tid; o < nout * nin; o += TTHREADS) {
                const int g = o / nin, c = o - g * nin;
                float s = 0.f;
#pragma unroll 8
                for (int rr = 0; rr < TT; ++rr) s = fmaf(dcur[rr * tp.SH + g], ain[rr * sin + c], s);
                gw[o] += s;
            }
            // db_l[g] += sum_r d[r][g]
            float* gb = gw + nout * nin;
            for (int g = tid; g < nout; g += TTHREADS) {
                float s = 0.f;
#pragma unroll 8
                for (int rr = 0; rr < TT; ++rr) s += dcur[rr * tp.SH + g];
                gb[g] += s;
            }
            // d_{l-1}[r][c] = (sum_g W_l[g][c] d[r][g]) * (1 - a_l[r][c]^2)
            if (l > 0) {
                const float* __restrict__ W = tp.W[l];
                const float* drow = dcur + r * tp.SH;
                for (int c = q; c < nin; c += QN) {
                    float s = 0.f;
                    for (int g = 0; g < nout; ++g) s = fmaf(__ldg(W + (size_t)g * nin + c), drow[g], s);
                    const float a = ain[r * sin + c];
                    dnext[r * tp.SH + c] = s * (1.f - a * a);
                }
            }
            __syncthreads();
            float* tmp = dcur; dcur = dnext; dnext = tmp;
        }
    }
    // ---- this CTA's sum of squared errors ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < TTHREADS / 32; ++w) t += s_red[w];
        tp.lpart[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) k_train_adam(const float* __restrict__ gpart, const float* __restrict__ lpart, int nparts,
                                                    int P, AdamParams ap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        float t = 0.f;
        for (int c = 0; c < nparts; ++c) t += lpart[c];
        *ap.loss_out = t * ap.inv_count;
    }
    if (i >= P) return;
    float g = 0.f;
    for (int c = 0; c < nparts; ++c) g += gpart[(size_t)c * P + i];
    if (ap.grads_out) ap.grads_out[i] = g;
    if (!ap.apply) return;
    int tsr = 0;
    while (tsr + 1 < ap.n_tensors && i >= ap.start[tsr + 1]) ++tsr;
    const int j = i - ap.start[tsr];
    // torch.optim.Adam (_single_tensor_adam): exp_avg.lerp_(grad, 1-beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad,
    // value=1-beta2); denom = sqrt(exp_avg_sq)/sqrt(bias_correction2) + eps; param.addcdiv_(exp_avg, denom, value=-lr/bc1)
    float m = ap.m[tsr][j], v = ap.v[tsr][j];
    m = m + (g - m) * ap.one_minus_beta1;
    v = v * ap.beta2 + ap.one_minus_beta2 * g * g;
    const float denom = sqrtf(v) / ap.bc2_sqrt + ap.eps;
    ap.m[tsr][j] = m;
    ap.v[tsr][j] = v;
    ap.p[tsr][j] = ap.p[tsr][j] - ap.step_size * (m / denom);
}

}  // namespace

namespace fgnn { void set_error(const char* msg); }      // fgnn.cu: the text fgnn_last_error() returns
static int tfail(const std::string& msg) { fgnn::set_error(msg.c_str()); return 1; }

#define TCK(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return tfail(buf__);                                                                     \
        }                                                                                            \
    } while (0)

struct fgnn_trainer {
    int K = 0, H = 0, L = 0, device = 0, P = 0, in0 = 0;
    int S0 = 0, SH = 0;
    int off[TLAYERS + 1] = {0};
    int start[2 * TLAYERS + 1] = {0};
    size_t smem = 0;
    int max_grid = 0;
    float* d_gpart = nullptr;
    float* d_lpart = nullptr;
    float* d_loss = nullptr;
    float* d_grads = nullptr;
    int64_t launches = 0;
};

extern "C" int fgnn_trainer_create(int32_t k, int32_t hidden, int32_t n_layers, int32_t device, fgnn_trainer** out) {
    if (!out) return tfail("fgnn_trainer_create: null argument");
    if (k < 1 || k > 4) return tfail("fgnn_trainer_create: k must be in 1..4");
    if (hidden < 1 || hidden > 128) return tfail("fgnn_trainer_create: hidden must be in 1..128");
    if (n_layers < 1 || n_layers > TLAYERS - 1) return tfail("fgnn_trainer_create: n_layers must be in 1..4");
    int ndev = 0;
    TCK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return tfail("fgnn_trainer_create: no such CUDA device");
    TCK(cudaSetDevice(device));
    fgnn_trainer* tr = new fgnn_trainer();
    tr->K = k; tr->H = hidden; tr->L = n_layers; tr->device = device; tr->in0 = TF * k;
    tr->S0 = tr->in0 | 1;
    tr->SH = hidden | 1;
    if (tr->SH < TA + 1) tr->SH = TA + 1;                     // the d buffers also hold the 2-wide output gradient
    int p = 0, ti = 0;
    for (int l = 0; l <= n_layers; ++l) {
        const int nin = l == 0 ? tr->in0 : hidden, nout = l == n_layers ? TA : hidden;
        tr->off[l] = p;
        tr->start[ti++] = p; p += nout * nin;
        tr->start[ti++] = p; p += nout;
    }
    tr->off[n_layers + 1] = p;
    tr->start[ti] = p;
    tr->P = p;
    tr->smem = ((size_t)TT * tr->S0 + (size_t)(n_layers + 2) * TT * tr->SH) * sizeof(float);
    cudaDeviceProp prop;
    TCK(cudaGetDeviceProperties(&prop, device));
    if (tr->smem > (size_t)prop.sharedMemPerBlockOptin) { delete tr; return tfail("fgnn_trainer_create: activations exceed shared memory"); }
    TCK(cudaFuncSetAttribute((const void*)k_train_fwd_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr->smem));
    int occ = 0;
    TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_train_fwd_bwd, TTHREADS, tr->smem));
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    tr->max_grid = prop.multiProcessorCount * occ;
    TCK(cudaMalloc((void**)&tr->d_gpart, (size_t)tr->max_grid * tr->P * sizeof(float)));
    TCK(cudaMalloc((void**)&tr->d_lpart, (size_t)tr->max_grid * sizeof(float)));
    TCK(cudaMalloc((void**)&tr->d_loss, sizeof(float)));
    TCK(cudaMalloc((void**)&tr->d_grads, (size_t)tr->P * sizeof(float)));
    *out = tr;
    return 0;
}

extern "C" int fgnn_trainer_destroy(fgnn_trainer* tr) {
    if (!tr) return 0;
    cudaSetDevice(tr->device);
    cudaFree(tr->d_gpart);
    cudaFree(tr->d_lpart);
    cudaFree(tr->d_loss);
    cudaFree(tr->d_grads);
    delete tr;
    return 0;
}

extern "C" int32_t fgnn_trainer_param_count(fgnn_trainer* tr) { return tr ? tr->P : 0; }
extern "C" int64_t fgnn_trainer_launch_count(fgnn_trainer* tr) { return tr ? tr->launches : 0; }

extern "C" int fgnn_trainer_step(fgnn_trainer* tr, int32_t batch, int32_t n_agents, const float* z, const float* target,
                                 float* const* params, float* const* exp_avg, float* const* exp_avg_sq, int64_t step,
                                 double lr, double beta1, double beta2, double eps, int32_t apply, float* loss_out,
                                 float* grads_out, void* stream) {
    if (!tr || !z || !target || !params) return tfail("fgnn_trainer_step: null argument");
    if (apply && (!exp_avg || !exp_avg_sq)) return tfail("fgnn_trainer_step: Adam state missing");
    if (batch < 1 || n_agents < 1) return tfail("fgnn_trainer_step: empty batch");
    if ((long long)batch * n_agents > 0x7fffffffll / 8) return tfail("fgnn_trainer_step: batch too large");
    if (apply && step < 1) return tfail("fgnn_trainer_step: step counts from 1");
    cudaStream_t st = (cudaStream_t)stream;
    TCK(cudaSetDevice(tr->device));
    const int nt = 2 * (tr->L + 1);
    for (int i = 0; i < nt; ++i) {
        if (!params[i] || (apply && (!exp_avg[i] || !exp_avg_sq[i]))) return tfail("fgnn_trainer_step: null tensor pointer");
    }
    TrainParams tp;
    tp.R = batch * n_agents; tp.N = n_agents; tp.K = tr->K; tp.in0 = tr->in0; tp.H = tr->H; tp.L = tr->L; tp.P = tr->P;
    tp.S0 = tr->S0; tp.SH = tr->SH;
    tp.z = z; tp.y = target;
    for (int l = 0; l <= tr->L; ++l) { tp.W[l] = params[2 * l]; tp.b[l] = params[2 * l + 1]; }
    for (int l = 0; l <= tr->L + 1; ++l) tp.off[l] = tr->off[l];
    tp.gpart = tr->d_gpart; tp.lpart = tr->d_lpart;
    tp.dscale = (float)(2.0 / ((double)tp.R * TA));
    const int n_tiles = (tp.R + TT - 1) / TT;
    const int grid = n_tiles < tr->max_grid ? n_tiles : tr->max_grid;
    k_train_fwd_bwd<<<grid, TTHREADS, tr->smem, st>>>(tp);
    TCK(cudaGetLastError());
    tr->launches += 1;

    AdamParams ap;
    ap.n_tensors = nt;
    for (int i = 0; i < nt; ++i) {
        ap.p[i] = params[i];
        ap.m[i] = apply ? exp_avg[i] : nullptr;
        ap.v[i] = apply ? exp_avg_sq[i] : nullptr;
    }
    for (int i = 0; i <= nt; ++i) ap.start[i] = tr->start[i];
    const double bc1 = 1.0 - std::pow(beta1, (double)(apply ? step : 1));
    const double bc2 = 1.0 - std::pow(beta2, (double)(apply ? step : 1));
    ap.one_minus_beta1 = (float)(1.0 - beta1);
    ap.beta2 = (float)beta2;
    ap.one_minus_beta2 = (float)(1.0 - beta2);
    ap.step_size = (float)(lr / bc1);
    ap.bc2_sqrt = (float)std::sqrt(bc2);
    ap.eps = (float)eps;
    ap.apply = apply ? 1 : 0;
    ap.inv_count = (float)(1.0 / ((double)tp.R * TA));
    ap.grads_out = grads_out ? tr->d_grads : nullptr;
    ap.loss_out = tr->d_loss;
    k_train_adam<<<(tr->P + 255) / 256, 256, 0, st>>>(tr->d_gpart, tr->d_lpart, grid, tr->P, ap);
    TCK(cudaGetLastError());
    tr->launches += 1;
    if (grads_out) TCK(cudaMemcpyAsync(grads_out, tr->d_grads, (size_t)tr->P * sizeof(float), cudaMemcpyDefault, st));
    if (loss_out) TCK(cudaMemcpyAsync(loss_out, tr->d_loss, sizeof(float), cudaMemcpyDefault, st));
    return 0;
}
