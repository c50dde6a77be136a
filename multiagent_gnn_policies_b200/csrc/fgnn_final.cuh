// fgnn_final.cuh -- the fused final kernel (last hop + readout + integrator) and the dense Actor.forward
// kernel.  Heavy templates: instantiated once per (K, HP) in fgnn_final.cu so they compile in parallel.
#pragma once
#include "fgnn_kernels.cuh"

namespace fgnn {

// Activation.  tanh.approx.f32 (2^-11) breaks the 1e-5 action tolerance and libdevice tanhf costs ~25
// instructions with a branch.  tanh_act(x) = sign(x) (1-e)/(1+e), e = exp(-2|x|) in (0, 1], via ex2.approx +
// rcp.approx: 7 instructions, two of them MUFU, absolute error < 2.5e-7 everywhere (what the next layer's sums
// see), exact saturation to +-1.  An earlier version added a Taylor branch for |x| < 0.35 to get RELATIVE
// accuracy near zero; it cost 10 more instructions per activation -- 64 activations per agent made it 40 % of
// the fused kernel's issue slots -- and bought nothing measurable in action parity.
__device__ __forceinline__ float tanh_act(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(x) * -2.885390081777927f));   // e^{-2|x|}
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return copysignf((1.0f - e) * r, x);
}

// ------------------------------------------------------------------------------------------
// readout MLP on CUDA cores (FFMA): in (6K) -> HP -> ... -> HP -> 2, tanh between layers.
// One thread per agent.  Hidden activations are staged in shared memory, one column per thread
// (bank-conflict free), so the loops over the input index stay ROLLED: the fully unrolled form is
// ~50 KB of SASS and stalls on instruction fetch (ncu: 41% no_instructions).  Weights are read as
// warp-broadcast float4 from shared memory (HP <= 64) or through L1 (HP = 128).
// ------------------------------------------------------------------------------------------
template <int IN, int HP, int THREADS>
__device__ __forceinline__ void mlp_ffma(const float (&in)[IN], const float* __restrict__ w, const WeightLayout wl,
                                         float* __restrict__ sh /* [HP][THREADS] (x2 when HP > 64) */, float& o0, float& o1) {
    constexpr int GC = HP < 64 ? HP : 64;          // output chunk held in registers
    const int tid = threadIdx.x;
    // layer 0: in[] lives in registers (IN <= 24), fully unrolled over the inputs
#pragma unroll 1
    for (int g0 = 0; g0 < HP; g0 += GC) {
        float acc[GC];
        const float4* b4 = reinterpret_cast<const float4*>(w + wl.off_b0() + g0);
#pragma unroll
        for (int g4 = 0; g4 < GC / 4; ++g4) {
            const float4 bb = b4[g4];
            acc[4 * g4 + 0] = bb.x; acc[4 * g4 + 1] = bb.y; acc[4 * g4 + 2] = bb.z; acc[4 * g4 + 3] = bb.w;
        }
#pragma unroll
        for (int i = 0; i < IN; ++i) {
            const float xi = in[i];
            const float4* w4 = reinterpret_cast<const float4*>(w + wl.off_w0() + i * HP + g0);
#pragma unroll
            for (int g4 = 0; g4 < GC / 4; ++g4) {
                const float4 ww = w4[g4];
                acc[4 * g4 + 0] = fmaf(xi, ww.x, acc[4 * g4 + 0]);
                acc[4 * g4 + 1] = fmaf(xi, ww.y, acc[4 * g4 + 1]);
                acc[4 * g4 + 2] = fmaf(xi, ww.z, acc[4 * g4 + 2]);
                acc[4 * g4 + 3] = fmaf(xi, ww.w, acc[4 * g4 + 3]);
            }
        }
#pragma unroll
        for (int g = 0; g < GC; ++g) sh[(g0 + g) * THREADS + tid] = tanh_act(acc[g]);
    }
    // hidden layers: activations from shared memory, rolled loop over the input index.
    // With more than one output chunk the layer writes to the second half of sh (ping-pong).
    float* cur = sh;
    float* nxt = (HP > GC) ? sh + HP * THREADS : sh;
    for (int l = 1; l < wl.L; ++l) {
#pragma unroll 1
        for (int g0 = 0; g0 < HP; g0 += GC) {
            float acc[GC];
            const float4* b4 = reinterpret_cast<const float4*>(w + wl.off_bh(l) + g0);
#pragma unroll
            for (int g4 = 0; g4 < GC / 4; ++g4) {
                const float4 bb = b4[g4];
                acc[4 * g4 + 0] = bb.x; acc[4 * g4 + 1] = bb.y; acc[4 * g4 + 2] = bb.z; acc[4 * g4 + 3] = bb.w;
            }
#pragma unroll 2
            for (int i = 0; i < HP; ++i) {
                const float xi = cur[i * THREADS + tid];
                const float4* w4 = reinterpret_cast<const float4*>(w + wl.off_wh(l) + i * HP + g0);
#pragma unroll
                for (int g4 = 0; g4 < GC / 4; ++g4) {
                    const float4 ww = w4[g4];
                    acc[4 * g4 + 0] = fmaf(xi, ww.x, acc[4 * g4 + 0]);
                    acc[4 * g4 + 1] = fmaf(xi, ww.y, acc[4 * g4 + 1]);
                    acc[4 * g4 + 2] = fmaf(xi, ww.z, acc[4 * g4 + 2]);
                    acc[4 * g4 + 3] = fmaf(xi, ww.w, acc[4 * g4 + 3]);
                }
            }
#pragma unroll
            for (int g = 0; g < GC; ++g) nxt[(g0 + g) * THREADS + tid] = tanh_act(acc[g]);
        }
        if (HP > GC) { float* tmp = cur; cur = nxt; nxt = tmp; }
    }
    const float2* wlp = reinterpret_cast<const float2*>(w + wl.off_wl());
    float a0 = w[wl.off_bl()], a1 = w[wl.off_bl() + 1];
#pragma unroll 4
    for (int i = 0; i < HP; ++i) {
        const float xi = cur[i * THREADS + tid];
        const float2 ww = wlp[i];
        a0 = fmaf(xi, ww.x, a0);
        a1 = fmaf(xi, ww.y, a1);
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
    pdl_prologue();
    extern __shared__ __align__(16) float smem[];
    WeightLayout wl;
    wl.in0 = F * K; wl.HP = HP; wl.L = p.L;
    constexpr bool WIDE = (HP > 64);
    // smem: [weights (HP <= 64 only)] [activation staging: HP x THREADS (x2 when HP = 128)]
    float* sh_act = smem + (WIDE ? 0 : wl.total());
    if (!WIDE) {
        const int nw4 = wl.total() / 4;
        const float4* gw = reinterpret_cast<const float4*>(p.weights);
        float4* sw4 = reinterpret_cast<float4*>(smem);
        for (int i = threadIdx.x; i < nw4; i += FINAL_THREADS) sw4[i] = __ldg(gw + i);
        __syncthreads();
    }
    const int t = *p.t;
    const size_t M = p.M;
    const int n_owned = owned_count(p);
    const int n_tiles = (n_owned + FINAL_THREADS - 1) / FINAL_THREADS;
    double racc[4] = {0, 0, 0, 0};
    long long klo = 0x7fffffffffffffffll, khi = -0x7fffffffffffffffll - 1;     // kept x-interval (sharded, fused pack)
    double safe_lo = 1.0, safe_hi = -1.0;                                     // interior interval of the step (k_shard_prepare), read once
    if (CLOSED && p.fuse) { safe_lo = p.fuse->ctl.safe[0]; safe_hi = p.fuse->ctl.safe[1]; }
    const int tile_end = (p.tile_hi > 0 && p.tile_hi < n_tiles) ? p.tile_hi : n_tiles;     // chunked launches (fgnn_policy)
    for (int tile = p.tile_lo + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int oi = tile * FINAL_THREADS + threadIdx.x;
        const int a = oi < n_owned ? owned_agent(p, oi) : -1;     // -1: beyond the list or a handed-over slot
        const bool valid = a >= 0;
        float in[F * K];
#pragma unroll
        for (int i = 0; i < F * K; ++i) in[i] = 0.f;
        double4 st_own = make_double4(0, 0, 0, 0);         // integrator input, fetched early: its latency hides behind the gather
        if (CLOSED && valid) st_own = ldg256(&p.state[a]);
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
            if (K >= 2 && p.last_hop_done) {          // tap K-1 finished by a separate hop launch
                float v[F];
                load_row6(p.zbuf + (size_t)(K - 1) * M * ROW, a, v);
#pragma unroll
                for (int f = 0; f < F; ++f) in[(K - 1) * F + f] = v[f];
            } else if (K >= 2) {   // last hop, tap K-1, through graph t-(K-2)
                constexpr int j = K - 2;
                const int g = slot_of(t - j, K);
                const float* __restrict__ src = (j == 0) ? p.xhist + (size_t)slot_of(t - (K - 1), K) * M * ROW
                                                         : p.ybuf + ((size_t)((j - 1) & 1) * K + (K - 1)) * M * ROW;
                const float* const srcs[1] = {src};
                float acc1[1][F];
                gather_rows<1, (j > 0)>(p, g, a, srcs, acc1);
                float acc[F];
#pragma unroll
                for (int f = 0; f < F; ++f) acc[f] = acc1[0][f];
#pragma unroll
                for (int f = 0; f < F; ++f) in[(K - 1) * F + f] = acc[f];
                if (p.write_z_last) store_row6(p.zbuf + (size_t)(K - 1) * M * ROW, a, acc);
            }
        }
        float o0, o1;
        mlp_ffma<F * K, HP, FINAL_THREADS>(in, WIDE ? p.weights : smem, wl, sh_act, o0, o1);
        if (valid) {
            reinterpret_cast<float2*>(p.action)[a] = make_float2(o0, o1);
            if (CLOSED) {
                const double4 st_new = integrate_and_bin(p, a, st_own, o0, o1, racc);
                if (p.fuse) shard_pack_agent(p, *p.fuse, oi, a, st_new, klo, khi, safe_lo, safe_hi);
            }
        }
    }
    if (CLOSED) reward_block_flush<FINAL_THREADS>(p, racc);
    if (CLOSED && p.fuse) shard_interval_flush<FINAL_THREADS>(p.fuse->ctl, klo, khi);
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
    mlp_ffma<F * K, HP, FINAL_THREADS>(in, WIDE ? weights : sw, wl, sw + (WIDE ? 0 : wl.total()), o0, o1);
    if (valid) {
        out[((size_t)b * 2 + 0) * N2 + n] = o0;
        out[((size_t)b * 2 + 1) * N2 + n] = o1;
    }
}


}  // namespace fgnn
