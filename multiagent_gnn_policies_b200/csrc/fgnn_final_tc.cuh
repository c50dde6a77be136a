// fgnn_final_tc.cuh -- fused final kernel with the readout MLP on the 5th-gen tensor cores.
//
//   last hop (CSR gather) -> z (6K inputs per agent) -> [tcgen05.mma kind::tf32, 3xTF32 split] -> tanh ->
//   ... hidden layers ... -> 2-wide output layer (FFMA) -> double integrator + binning.
//
// One CTA = 128 threads = one 128-agent tile per iteration = one UMMA M=128 tile: thread r owns agent
// row r everywhere (gather, TMEM lane r via tcgen05.ld 32x32b, integrator).  Per hidden layer:
//   * every thread splits its activations into tf32 hi + lo and stores its row of the A operand into
//     shared memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices),
//   * fence.proxy.async + __syncthreads, then ONE thread issues 3 MMAs per k-step
//     (A_lo*B_hi, A_hi*B_lo, A_hi*B_hi; fp32 accumulate in TMEM) and tcgen05.commit's to an mbarrier,
//   * all threads wait on the mbarrier and read their accumulator row back with tcgen05.ld.
// The 3xTF32 split keeps the products at fp32-level accuracy (plain TF32 would break the 1e-5 action
// parity).  Weights are pre-split into tf32 hi/lo on the host and sit in shared memory in the same
// canonical layout for the whole kernel.
#pragma once
#include "fgnn_final.cuh"

namespace fgnn {

// ---- layout of the tensor-core weight pack (bytes), shared by host packer and kernel ---------------
struct TcLayout {
    int K0;        // layer-0 reduction length padded to a multiple of 8 (tf32 UMMA_K)
    int HP;        // padded hidden width = UMMA N = layer>=1 reduction length
    int L;         // hidden layers
    __host__ __device__ static int pad8(int v) { return (v + 7) & ~7; }
    __host__ __device__ int w0_bytes() const { return HP * K0 * 4; }
    __host__ __device__ int wh_bytes() const { return HP * HP * 4; }
    // [W0_hi][W0_lo] then per hidden layer l = 1..L-1 [Wh_hi][Wh_lo], then fp32: b0[HP], bh[L-1][HP], wl[HP][2], bl[2]
    __host__ __device__ int off_w0(int lo) const { return lo * w0_bytes(); }
    __host__ __device__ int off_wh(int l, int lo) const { return 2 * w0_bytes() + ((l - 1) * 2 + lo) * wh_bytes(); }
    __host__ __device__ int off_f32() const { return 2 * w0_bytes() + (L - 1) * 2 * wh_bytes(); }
    __host__ __device__ int off_b(int l) const { return off_f32() + l * HP * 4; }          // l = 0..L-1
    __host__ __device__ int off_wl() const { return off_f32() + L * HP * 4; }
    __host__ __device__ int off_bl() const { return off_wl() + HP * 2 * 4; }
    __host__ __device__ int total_bytes() const { return (off_bl() + 8 + 127) & ~127; }
    // canonical K-major no-swizzle: element (row, col) of an [rows x Kdim] tf32 operand
    __host__ __device__ static int canon_off(int row, int col, int Kdim) {
        return (row >> 3) * (Kdim / 4) * 128 + (col >> 2) * 128 + (row & 7) * 16 + (col & 3) * 4;
    }
};

#ifdef __CUDACC__
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE, version 1 (Blackwell):
// [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 16-byte K chunks) | [32,46) SBO>>4 (stride between
// 8-row groups) | [46,48) version = 1 | [61,64) layout type = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}

// instruction descriptor kind::tf32: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}\n"
        :: "r"(mbar), "r"(parity) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_slot), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}

// TMEM -> registers: lane = this thread's row, N consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]) {
    static_assert(N % 16 == 0, "columns in multiples of 16");
#pragma unroll
    for (int c = 0; c < N; c += 16) tmem_ld16(taddr + c, &v[c]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x = hi + lo with hi the round-to-nearest (ties away) tf32 of x; the tensor core ignores lo's low 13 mantissa bits.
// cvt.rna.tf32.f32 compiles to this add-and-mask plus an inf/NaN guard (FSETP + SEL per value): the guard is dropped --
// every value split here is finite (features of distinct agents, tanh outputs).  Two values at a time: the subtraction is
// one packed FFMA2 (hi * -1 + x is exactly x - hi).
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split2_tf32(float2 x, float2& hi, float2& lo) {
    hi = make_float2(tf32_hi(x.x), tf32_hi(x.y));
    lo = __ffma2_rn(hi, make_float2(-1.f, -1.f), x);
}

// tanh_act (fgnn_final.cuh) of two pre-activations h + b at once, same operations and roundings, packed fp32x2 arithmetic
// where the pipe has it (sm_100 FADD2 / FMUL2 / FFMA2): 5.5 issue slots per activation instead of 8.
__device__ __forceinline__ float2 tanh2(float2 h, float2 b) {
    const float2 x = __fadd2_rn(h, b);
    const float2 y = __fmul2_rn(x, make_float2(2.885390081777927f, 2.885390081777927f));
    float e0, e1, r0, r1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-fabsf(y.x)));                     // e^{-2|x|}
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-fabsf(y.y)));
    const float2 e = make_float2(e0, e1);
    const float2 den = __fadd2_rn(e, make_float2(1.f, 1.f));
    const float2 num = __ffma2_rn(e, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));    // 1 - e, one rounding
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(den.y));
    const float2 q = __fmul2_rn(num, make_float2(r0, r1));
    return make_float2(copysignf(q.x, x.x), copysignf(q.y, x.y));
}

}  // namespace tc

__host__ __device__ constexpr int tc_tmem_cols(int HP) { return 2 * HP < 32 ? 32 : 2 * HP; }

struct TcCtx {                 // per-CTA state of the tensor-core readout between its setup / tile / teardown parts
    uint8_t* s_w;              // weight pack (canonical layout) in shared memory
    uint8_t* s_ahi;            // A operand tiles (tf32 hi / lo)
    uint8_t* s_alo;
    uint32_t mbar;             // mbarrier (shared address) the MMAs commit to
    uint32_t tmem_base;        // TMEM allocation
    uint32_t phase;            // mbarrier phase
};

// weights -> smem (already in canonical layout), barrier init, TMEM allocation.  Every thread of the CTA.
template <int K, int HP>
__device__ __forceinline__ void final_tc_setup(const Params& p, const uint8_t* __restrict__ tcw, uint8_t* smem_raw, TcCtx& c) {
    static_assert(HP == 16 || HP == 32 || HP == 64, "tensor-core readout supports HP in {16,32,64}");
    constexpr int K0 = (F * K + 7) & ~7;
    constexpr int KA = K0 > HP ? K0 : HP;                  // widest A operand
    constexpr int TM_COLS = tc_tmem_cols(HP);
    TcLayout tl;
    tl.K0 = K0; tl.HP = HP; tl.L = p.L;
    // smem: [weight pack][A_hi 128xKA][A_lo 128xKA][mbarrier][tmem slot]
    c.s_w = smem_raw;
    c.s_ahi = smem_raw + tl.total_bytes();
    c.s_alo = c.s_ahi + 128 * KA * 4;
    uint64_t* s_mbar = reinterpret_cast<uint64_t*>(c.s_alo + 128 * KA * 4);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_mbar + 1);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int n16 = tl.total_bytes() / 16;
    const uint4* g = reinterpret_cast<const uint4*>(tcw);
    uint4* sw = reinterpret_cast<uint4*>(c.s_w);
    for (int i = tid; i < n16; i += FINAL_THREADS) sw[i] = __ldg(g + i);
    if (tid == 0) tc::mbar_init(tc::smem_u32(s_mbar), 1);
    if (warp == 0) tc::tmem_alloc<TM_COLS>(tc::smem_u32(s_tmem));
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    c.tmem_base = *s_tmem;
    c.mbar = tc::smem_u32(s_mbar);
    c.phase = 0;
}

template <int HP>
__device__ __forceinline__ void final_tc_teardown(const TcCtx& c) {
    tc::fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc<tc_tmem_cols(HP)>(c.tmem_base);
}

// the CTA's tiles of one launch (persistent loop), then the per-block reward / interval partials
template <int K, int HP, bool CLOSED>
__device__ __forceinline__ void final_tc_tiles(const Params& p, TcCtx& c) {
    constexpr int K0 = (F * K + 7) & ~7;
    constexpr uint32_t IDESC = tc::make_idesc(128, HP);
    TcLayout tl;
    tl.K0 = K0; tl.HP = HP; tl.L = p.L;
    uint8_t* const s_w = c.s_w;
    uint8_t* const s_ahi = c.s_ahi;
    uint8_t* const s_alo = c.s_alo;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const uint32_t tmem_base = c.tmem_base;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);      // this warp's lane quadrant
    const uint32_t mbar = c.mbar;
    uint32_t phase = c.phase;
    const float* s_f32 = reinterpret_cast<const float*>(s_w);

    const int t = *p.t;
    const size_t M = p.M;
    const int n_owned = owned_count(p);
    const int n_tiles = (n_owned + FINAL_THREADS - 1) / FINAL_THREADS;
    double racc[4] = {0, 0, 0, 0};
    long long klo = 0x7fffffffffffffffll, khi = -0x7fffffffffffffffll - 1;     // kept x-interval (sharded, fused pack)
    double safe_lo = 1.0, safe_hi = -1.0;                                     // interior interval of the step (k_shard_prepare), read once
    if (CLOSED && p.fuse) { safe_lo = p.fuse->ctl.safe[0]; safe_hi = p.fuse->ctl.safe[1]; }
    const int tile_end = (p.tile_hi > 0 && p.tile_hi < n_tiles) ? p.tile_hi : n_tiles;     // chunked launches (fgnn_policy)
    // inputs of one tile: the agent, its z rows (6K readout inputs) and -- closed loop -- its state for the integrator
    auto load_tile = [&](int tile, int& a, float (&in)[K0], double4& st_own) {
        const int oi = tile * FINAL_THREADS + tid;
        a = (tile < tile_end && oi < n_owned) ? owned_agent(p, oi) : -1;     // -1: beyond the list or a handed-over slot
#pragma unroll
        for (int i = 0; i < K0; ++i) in[i] = 0.f;
        st_own = make_double4(0, 0, 0, 0);
        if (a < 0) return;
        if (CLOSED) st_own = ldg256(&p.state[a]);
        {   // z_0 = x_t
            float v[F];
            load_row6(p.xhist + (size_t)slot_of(t, K) * M * ROW, a, v);
#pragma unroll
            for (int f = 0; f < F; ++f) in[f] = v[f];
        }
#pragma unroll
        for (int k = 1; k < K - 1; ++k) {
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
        } else if (K >= 2) {
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
    };
    // Software pipeline over the CTA's tiles: the NEXT tile's rows are requested before the current tile's MLP starts, so
    // their latency (the kernel's largest stall: every warp of a CTA is in the same phase) hides behind ~1000 instructions.
    int a_cur;
    float in[K0];
    double4 st_own;
    load_tile(p.tile_lo + blockIdx.x, a_cur, in, st_own);
    for (int tile = p.tile_lo + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int oi = tile * FINAL_THREADS + tid;
        const int a = a_cur;
        const bool valid = a >= 0;
        const double4 st_cur = st_own;
        // ---- layer 0: A0 = z (tf32 hi/lo), canonical layout, row = tid ----
        {
            const int rowoff = (tid >> 3) * (K0 / 4) * 128 + (tid & 7) * 16;
#pragma unroll
            for (int c = 0; c < K0 / 4; ++c) {
                float2 h0, l0, h1, l1;
                tc::split2_tf32(make_float2(in[4 * c + 0], in[4 * c + 1]), h0, l0);
                tc::split2_tf32(make_float2(in[4 * c + 2], in[4 * c + 3]), h1, l1);
                *reinterpret_cast<float4*>(s_ahi + rowoff + c * 128) = make_float4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<float4*>(s_alo + rowoff + c * 128) = make_float4(l0.x, l0.y, l1.x, l1.y);
            }
        }
        load_tile(tile + gridDim.x, a_cur, in, st_own);          // prefetch (in[] was consumed above)
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t sbo = (K0 / 4) * 128;
            const uint64_t a_hi = tc::make_desc(tc::smem_u32(s_ahi), 128, sbo);
            const uint64_t a_lo = tc::make_desc(tc::smem_u32(s_alo), 128, sbo);
            const uint64_t b_hi = tc::make_desc(tc::smem_u32(s_w + tl.off_w0(0)), 128, sbo);
            const uint64_t b_lo = tc::make_desc(tc::smem_u32(s_w + tl.off_w0(1)), 128, sbo);
#pragma unroll
            for (int ks = 0; ks < K0 / 8; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 16);          // 2 chunks x 128 B >> 4
                tc::mma_tf32(tmem_base, a_lo + adv, b_hi + adv, IDESC, ks > 0);
                tc::mma_tf32(tmem_base, a_hi + adv, b_lo + adv, IDESC, 1);
                tc::mma_tf32(tmem_base, a_hi + adv, b_hi + adv, IDESC, 1);
            }
            tc::commit(mbar);
        }
        tc::mbar_wait(mbar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        float h[HP];
        tc::tmem_ld<HP>(tmem_row, h);
        // ---- hidden layers l = 1 .. L-1 ----
        for (int l = 1; l < p.L; ++l) {
            const float2* bprev = reinterpret_cast<const float2*>(s_f32 + tl.off_b(l - 1) / 4);
            const int rowoff = (tid >> 3) * (HP / 4) * 128 + (tid & 7) * 16;
#pragma unroll
            for (int c = 0; c < HP / 4; ++c) {
                float2 h0, l0, h1, l1;
                tc::split2_tf32(tc::tanh2(make_float2(h[4 * c + 0], h[4 * c + 1]), bprev[2 * c + 0]), h0, l0);
                tc::split2_tf32(tc::tanh2(make_float2(h[4 * c + 2], h[4 * c + 3]), bprev[2 * c + 1]), h1, l1);
                *reinterpret_cast<float4*>(s_ahi + rowoff + c * 128) = make_float4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<float4*>(s_alo + rowoff + c * 128) = make_float4(l0.x, l0.y, l1.x, l1.y);
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            const uint32_t dcol = tmem_base + (uint32_t)((l & 1) * HP);
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t sbo = (HP / 4) * 128;
                const uint64_t a_hi = tc::make_desc(tc::smem_u32(s_ahi), 128, sbo);
                const uint64_t a_lo = tc::make_desc(tc::smem_u32(s_alo), 128, sbo);
                const uint64_t b_hi = tc::make_desc(tc::smem_u32(s_w + tl.off_wh(l, 0)), 128, sbo);
                const uint64_t b_lo = tc::make_desc(tc::smem_u32(s_w + tl.off_wh(l, 1)), 128, sbo);
#pragma unroll
                for (int ks = 0; ks < HP / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 16);
                    tc::mma_tf32(dcol, a_lo + adv, b_hi + adv, IDESC, ks > 0);
                    tc::mma_tf32(dcol, a_hi + adv, b_lo + adv, IDESC, 1);
                    tc::mma_tf32(dcol, a_hi + adv, b_hi + adv, IDESC, 1);
                }
                tc::commit(mbar);
            }
            tc::mbar_wait(mbar, phase);
            phase ^= 1;
            tc::fence_after_sync();
            tc::tmem_ld<HP>(tmem_row + (uint32_t)((l & 1) * HP), h);
        }
        // ---- output layer on CUDA cores: HP FFMA2 (hidden units two at a time: even and odd partial sums) ----
        const float2* blast = reinterpret_cast<const float2*>(s_f32 + tl.off_b(p.L - 1) / 4);
        const float4* wlp = reinterpret_cast<const float4*>(s_w + tl.off_wl());      // [HP/2]: (w_i^0, w_{i+1}^0, w_i^1, w_{i+1}^1)
        const float* bl = s_f32 + tl.off_bl() / 4;
        float2 acc0 = make_float2(bl[0], 0.f), acc1 = make_float2(bl[1], 0.f);
#pragma unroll
        for (int i = 0; i < HP; i += 2) {
            const float2 hv = tc::tanh2(make_float2(h[i], h[i + 1]), blast[i / 2]);
            const float4 ww = wlp[i / 2];
            acc0 = __ffma2_rn(hv, make_float2(ww.x, ww.y), acc0);
            acc1 = __ffma2_rn(hv, make_float2(ww.z, ww.w), acc1);
        }
        const float o0 = acc0.x + acc0.y, o1 = acc1.x + acc1.y;
        if (valid) {
            reinterpret_cast<float2*>(p.action)[a] = make_float2(o0, o1);
            if (CLOSED) {
                const double4 st_new = integrate_and_bin(p, a, st_cur, o0, o1, racc);
                if (p.fuse) shard_pack_agent(p, *p.fuse, oi, a, st_new, klo, khi, safe_lo, safe_hi);
            }
        }
    }
    if (CLOSED) reward_block_flush<FINAL_THREADS>(p, racc);
    if (CLOSED && p.fuse) shard_interval_flush<FINAL_THREADS>(p.fuse->ctl, klo, khi);
    c.phase = phase;
}

template <int K, int HP, bool CLOSED>
__global__ void __launch_bounds__(FINAL_THREADS, HP <= 32 ? 4 : 2) k_final_tc(Params p, const uint8_t* __restrict__ tcw) {
    pdl_prologue();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TcCtx c;
    final_tc_setup<K, HP>(p, tcw, smem_raw, c);
    final_tc_tiles<K, HP, CLOSED>(p, c);
    final_tc_teardown<HP>(c);
}
#endif  // __CUDACC__

}  // namespace fgnn
