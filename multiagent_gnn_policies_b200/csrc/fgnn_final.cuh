// fgnn_final.cuh -- the fused final kernel (last hop + readout + integrator) and the dense Actor.forward
// kernel.  Heavy templates: instantiated once per (K, HP) in fgnn_final.cu so they compile in parallel.
#pragma once
#include "fgnn_kernels.cuh"

namespace fgnn {

// ------------------------------------------------------------------------------------------
// readout MLP on CUDA cores (FFMA): in (6K) -> HP -> ... -> HP -> 2, tanh between layers.
// Weights live in shared memory (HP <= 64) and are read as warp-broadcast float4.
// ------------------------------------------------------------------------------------------
template <int IN, int HP>
__device__ __forceinline__ void mlp_ffma(const float (&in)[IN], const float* __restrict__ sw, const WeightLayout wl,
                                         float& o0, float& o1) {
    float h[HP];
    {
        const float* b0 = sw + wl.off_b0();
#pragma unroll
        for (int g = 0; g < HP; ++g) h[g] = b0[g];
        const float4* w0 = reinterpret_cast<const float4*>(sw + wl.off_w0());
#pragma unroll
        for (int i = 0; i < IN; ++i) {
            const float xi = in[i];
#pragma unroll
            for (int g4 = 0; g4 < HP / 4; ++g4) {
                const float4 w = w0[i * (HP / 4) + g4];
                h[4 * g4 + 0] = fmaf(xi, w.x, h[4 * g4 + 0]);
                h[4 * g4 + 1] = fmaf(xi, w.y, h[4 * g4 + 1]);
                h[4 * g4 + 2] = fmaf(xi, w.z, h[4 * g4 + 2]);
                h[4 * g4 + 3] = fmaf(xi, w.w, h[4 * g4 + 3]);
            }
        }
#pragma unroll
        for (int g = 0; g < HP; ++g) h[g] = tanhf(h[g]);
    }
    for (int l = 1; l < wl.L; ++l) {
        float h2[HP];
        const float* bh = sw + wl.off_bh(l);
#pragma unroll
        for (int g = 0; g < HP; ++g) h2[g] = bh[g];
        const float4* wh = reinterpret_cast<const float4*>(sw + wl.off_wh(l));
#pragma unroll
        for (int i = 0; i < HP; ++i) {
            const float xi = h[i];
#pragma unroll
            for (int g4 = 0; g4 < HP / 4; ++g4) {
                const float4 w = wh[i * (HP / 4) + g4];
                h2[4 * g4 + 0] = fmaf(xi, w.x, h2[4 * g4 + 0]);
                h2[4 * g4 + 1] = fmaf(xi, w.y, h2[4 * g4 + 1]);
                h2[4 * g4 + 2] = fmaf(xi, w.z, h2[4 * g4 + 2]);
                h2[4 * g4 + 3] = fmaf(xi, w.w, h2[4 * g4 + 3]);
            }
        }
#pragma unroll
        for (int g = 0; g < HP; ++g) h[g] = tanhf(h2[g]);
    }
    const float2* wlp = reinterpret_cast<const float2*>(sw + wl.off_wl());
    const float* bl = sw + wl.off_bl();
    float a0 = bl[0], a1 = bl[1];
#pragma unroll
    for (int i = 0; i < HP; ++i) {
        const float2 w = wlp[i];
        a0 = fmaf(h[i], w.x, a0);
        a1 = fmaf(h[i], w.y, a1);
    }
    o0 = a0;
    o1 = a1;
}

// Wide layers (HP = 128): activations staged in shared memory (column per thread), weights read
// through L1 as warp-uniform loads.  Correct for any HP; used where registers would not hold h[].
template <int IN, int HP, int THREADS>
__device__ __forceinline__ void mlp_wide(const float (&in)[IN], const float* __restrict__ gw, const WeightLayout wl,
                                         float* sh /* [2][HP][THREADS] */, float& o0, float& o1) {
    const int tid = threadIdx.x;
    float* cur = sh;
    float* nxt = sh + HP * THREADS;
    for (int g = 0; g < HP; ++g) {
        float acc = __ldg(gw + wl.off_b0() + g);
#pragma unroll
        for (int i = 0; i < IN; ++i) acc = fmaf(in[i], __ldg(gw + wl.off_w0() + i * HP + g), acc);
        cur[g * THREADS + tid] = tanhf(acc);
    }
    for (int l = 1; l < wl.L; ++l) {
        for (int g0 = 0; g0 < HP; g0 += 16) {
            float acc[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) acc[u] = __ldg(gw + wl.off_bh(l) + g0 + u);
            for (int i = 0; i < HP; ++i) {
                const float xi = cur[i * THREADS + tid];
                const float4* w = reinterpret_cast<const float4*>(gw + wl.off_wh(l) + i * HP + g0);
#pragma unroll
                for (int u4 = 0; u4 < 4; ++u4) {
                    const float4 ww = __ldg(w + u4);
                    acc[4 * u4 + 0] = fmaf(xi, ww.x, acc[4 * u4 + 0]);
                    acc[4 * u4 + 1] = fmaf(xi, ww.y, acc[4 * u4 + 1]);
                    acc[4 * u4 + 2] = fmaf(xi, ww.z, acc[4 * u4 + 2]);
                    acc[4 * u4 + 3] = fmaf(xi, ww.w, acc[4 * u4 + 3]);
                }
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) nxt[(g0 + u) * THREADS + tid] = tanhf(acc[u]);
        }
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    float a0 = __ldg(gw + wl.off_bl()), a1 = __ldg(gw + wl.off_bl() + 1);
    for (int i = 0; i < HP; ++i) {
        const float xi = cur[i * THREADS + tid];
        a0 = fmaf(xi, __ldg(gw + wl.off_wl() + 2 * i), a0);
        a1 = fmaf(xi, __ldg(gw + wl.off_wl() + 2 * i + 1), a1);
    }
    o0 = a0;
    o1 = a1;
}

// ------------------------------------------------------------------------------------------
// K_F  final: last hop (tap K-1 through graph t-K+2) + gather of z_0..z_{K-2} + readout MLP
//      (+ integrator + binning when CLOSED): DAGGER.select_action fused with env.step's update.
// ------------------------------------------------------------------------------------------

template <int K, int HP, bool CLOSED>
__global__ void __launch_bounds__(FINAL_THREADS) k_final(Params p) {
    extern __shared__ __align__(16) float smem[];
    WeightLayout wl;
    wl.in0 = F * K; wl.HP = HP; wl.L = p.L;
    constexpr bool WIDE = (HP > 64);
    if (!WIDE) {
        const int nw4 = wl.total() / 4;
        const float4* gw = reinterpret_cast<const float4*>(p.weights);
        float4* sw4 = reinterpret_cast<float4*>(smem);
        for (int i = threadIdx.x; i < nw4; i += FINAL_THREADS) sw4[i] = __ldg(gw + i);
        __syncthreads();
    }
    const int t = *p.t;
    const size_t M = p.M;
    const int n_tiles = (p.M + FINAL_THREADS - 1) / FINAL_THREADS;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int a = tile * FINAL_THREADS + threadIdx.x;
        const bool valid = a < p.M;
        float in[F * K];
#pragma unroll
        for (int i = 0; i < F * K; ++i) in[i] = 0.f;
        if (valid) {
            {   // z_0 = x_t
                float v[F];
                load_row6(p.xhist + (size_t)slot_of(t, K) * M * ROW, a, v);
#pragma unroll
                for (int f = 0; f < F; ++f) in[f] = v[f];
            }
#pragma unroll
            for (int k = 1; k < K - 1; ++k) {   // finished taps from earlier hops
                float v[F];
                load_row6(p.zbuf + (size_t)k * M * ROW, a, v);
#pragma unroll
                for (int f = 0; f < F; ++f) in[k * F + f] = v[f];
            }
            if (K >= 2) {   // last hop, tap K-1, through graph t-(K-2)
                constexpr int j = K - 2;
                const int g = slot_of(t - j, K);
                const float* __restrict__ src = (j == 0) ? p.xhist + (size_t)slot_of(t - (K - 1), K) * M * ROW
                                                         : p.ybuf + ((size_t)((j - 1) & 1) * K + (K - 1)) * M * ROW;
                const unsigned rs = p.row_start[(size_t)g * M + a];
                const int d = p.deg[(size_t)g * M + a];
                const int* __restrict__ cols = p.cols + (size_t)g * p.nnz_cap + rs;
                const float* __restrict__ sinv = p.sinv + (size_t)g * M;
                float acc[F];
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f] = 0.f;
                for (int e = 0; e < d; ++e) {
                    const int m = __ldg(&cols[e]);
                    const float sc = __ldg(&sinv[m]);
                    float v[F];
                    load_row6(src, m, v);
#pragma unroll
                    for (int f = 0; f < F; ++f) acc[f] = fmaf(v[f], sc, acc[f]);
                }
#pragma unroll
                for (int f = 0; f < F; ++f) in[(K - 1) * F + f] = acc[f];
                if (p.write_z_last) store_row6(p.zbuf + (size_t)(K - 1) * M * ROW, a, acc);
            }
        }
        float o0, o1;
        if (WIDE) {
            mlp_wide<F * K, HP, FINAL_THREADS>(in, p.weights, wl, smem, o0, o1);
        } else {
            mlp_ffma<F * K, HP>(in, smem, wl, o0, o1);
        }
        if (valid) {
            reinterpret_cast<float2*>(p.action)[a] = make_float2(o0, o1);
            if (CLOSED) integrate_and_bin(p, a, o0, o1);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Actor.forward on DENSE tensors (learner/actor.py:45-86): out[b,0,:,n] = MLP( sum_m ds[b,k,f,m] gso[b,k,m,n] )
// ------------------------------------------------------------------------------------------

template <int K, int HP>
__global__ void __launch_bounds__(FINAL_THREADS) k_actor_dense(const float* __restrict__ ds, const float* __restrict__ gso,
                                                               float* __restrict__ out, const float* __restrict__ weights,
                                                               int L, int N2) {
    extern __shared__ __align__(16) float smem[];
    WeightLayout wl;
    wl.in0 = F * K; wl.HP = HP; wl.L = L;
    constexpr bool WIDE = (HP > 64);
    float* tile = smem;                                    // [K*F][DENSE_MT]
    float* sw = smem + K * F * DENSE_MT;                   // weights or wide staging
    if (!WIDE) {
        for (int i = threadIdx.x; i < wl.total(); i += FINAL_THREADS) sw[i] = __ldg(weights + i);
    }
    const int b = blockIdx.y;
    const int n = blockIdx.x * FINAL_THREADS + threadIdx.x;
    const bool valid = n < N2;
    float in[F * K];
#pragma unroll
    for (int i = 0; i < F * K; ++i) in[i] = 0.f;
    const float* dsb = ds + (size_t)b * K * F * N2;
    const float* gb = gso + (size_t)b * K * N2 * N2;
    for (int m0 = 0; m0 < N2; m0 += DENSE_MT) {
        __syncthreads();
        for (int i = threadIdx.x; i < K * F * DENSE_MT; i += FINAL_THREADS) {
            int kf = i / DENSE_MT, mm = i % DENSE_MT;
            tile[i] = (m0 + mm < N2) ? __ldg(dsb + (size_t)kf * N2 + m0 + mm) : 0.f;
        }
        __syncthreads();
        if (valid) {
            const int mend = min(DENSE_MT, N2 - m0);
            for (int mm = 0; mm < mend; ++mm) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float gv = __ldg(gb + ((size_t)k * N2 + m0 + mm) * N2 + n);
#pragma unroll
                    for (int f = 0; f < F; ++f) in[k * F + f] = fmaf(tile[(k * F + f) * DENSE_MT + mm], gv, in[k * F + f]);
                }
            }
        }
    }
    __syncthreads();
    float o0, o1;
    if (WIDE) {
        mlp_wide<F * K, HP, FINAL_THREADS>(in, weights, wl, sw, o0, o1);
    } else {
        mlp_ffma<F * K, HP>(in, sw, wl, o0, o1);
    }
    if (valid) {
        out[((size_t)b * 2 + 0) * N2 + n] = o0;
        out[((size_t)b * 2 + 1) * N2 + n] = o1;
    }
}


}  // namespace fgnn
