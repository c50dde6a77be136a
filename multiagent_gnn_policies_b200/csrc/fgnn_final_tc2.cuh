// fgnn_final_tc2.cuh -- EXPERIMENTAL (readout_mode = 3, not a default; written at the end of round 1 without GPU
// time left to validate it -- see DESIGN.md section 8 item 2).
//
// The streaming readout of fgnn_final_tc.cuh with TWO warps per TMEM lane quadrant: a CTA is 256 threads for one
// 128-agent UMMA M tile; warps w and w + 4 both own TMEM lanes 32 (w & 3) .. +31 and each thread applies
// bias / tanh / tf32 split / shared-memory stores to HALF of the columns of its row.  Same shared memory per CTA,
// twice the warps per tile: the single-warp-per-quadrant kernel is latency-bound at 4 warps per scheduler (ncu:
// issue slots 48 % busy, ~20 % of the stall samples on the MMA round trip, 14 % on the input loads).
// Requires the last hop to have run as its own launch (the inputs are then plain streaming loads) and HP in {32, 64}.
#pragma once
#include "fgnn_final_tc.cuh"

namespace fgnn {

#ifdef __CUDACC__
constexpr int FINAL2_THREADS = 2 * FINAL_THREADS;

// Columns [LO, HI) of the padded layer-0 input row of agent a: column c = tap (c / 6), feature (c % 6); taps >= K are zero.
template <int K, int LO, int HI>
__device__ __forceinline__ void tc2_load_cols(const Params& p, int t, int a, float (&v)[HI - LO]) {
    const size_t M = p.M;
#pragma unroll
    for (int c = 0; c < HI - LO; ++c) v[c] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (k * F < HI && k * F + F > LO) {                        // tap k overlaps the range (compile time)
            float r[F];
            const float* rows = (k == 0) ? p.xhist + (size_t)slot_of(t, K) * M * ROW : p.zbuf + (size_t)k * M * ROW;
            load_row6(rows, a, r);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int c = k * F + f;
                if (c >= LO && c < HI) v[c - LO] = r[f];
            }
        }
    }
}

// split the N values of v into tf32 hi/lo and store them as N/4 consecutive 16-byte chunks starting at chunk `c0` of the
// thread's row of the canonical A tiles
template <int N>
__device__ __forceinline__ void tc2_store_chunks(uint8_t* s_ahi, uint8_t* s_alo, int rowoff, int c0, const float (&v)[N]) {
#pragma unroll
    for (int c = 0; c < N / 4; ++c) {
        float4 hi, lo;
        tc::split_tf32(v[4 * c + 0], hi.x, lo.x);
        tc::split_tf32(v[4 * c + 1], hi.y, lo.y);
        tc::split_tf32(v[4 * c + 2], hi.z, lo.z);
        tc::split_tf32(v[4 * c + 3], hi.w, lo.w);
        *reinterpret_cast<float4*>(s_ahi + rowoff + (c0 + c) * 128) = hi;
        *reinterpret_cast<float4*>(s_alo + rowoff + (c0 + c) * 128) = lo;
    }
}

template <int K, int HP, bool CLOSED>
__global__ void __launch_bounds__(FINAL2_THREADS, 4) k_final_tc2(Params p, const uint8_t* __restrict__ tcw) {
    static_assert(HP == 32 || HP == 64, "two-warp readout supports HP in {32, 64}");
    pdl_prologue();
    constexpr int K0 = (F * K + 7) & ~7;
    constexpr int KA = K0 > HP ? K0 : HP;
    constexpr int TM_COLS = tc_tmem_cols(HP);
    constexpr uint32_t IDESC = tc::make_idesc(128, HP);
    constexpr int HH = HP / 2;                             // hidden columns per thread
    constexpr int H0 = K0 / 2;                             // layer-0 input columns per thread (4, 8 or 12)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TcLayout tl;
    tl.K0 = K0; tl.HP = HP; tl.L = p.L;
    // smem: [weight pack][A_hi 128xKA][A_lo 128xKA][mbarrier][tmem slot][pad][output partials 128 x 2]
    uint8_t* s_w = smem_raw;
    uint8_t* s_ahi = smem_raw + tl.total_bytes();
    uint8_t* s_alo = s_ahi + 128 * KA * 4;
    uint64_t* s_mbar = reinterpret_cast<uint64_t*>(s_alo + 128 * KA * 4);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_mbar + 1);
    float* s_part = reinterpret_cast<float*>(s_mbar + 2);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3;                             // TMEM lane quadrant of this warp
    const int hsel = warp >> 2;                            // which half of the columns this thread handles
    const int row = quad * 32 + lane;                      // row of the 128-agent tile

    {
        const int n16 = tl.total_bytes() / 16;
        const uint4* g = reinterpret_cast<const uint4*>(tcw);
        uint4* s = reinterpret_cast<uint4*>(s_w);
        for (int i = tid; i < n16; i += FINAL2_THREADS) s[i] = __ldg(g + i);
        if (tid == 0) tc::mbar_init(tc::smem_u32(s_mbar), 1);
        if (warp == 0) tc::tmem_alloc<TM_COLS>(tc::smem_u32(s_tmem));
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    const uint32_t tmem_base = *s_tmem;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(hsel * HH);   // lane quadrant, column half
    const uint32_t mbar = tc::smem_u32(s_mbar);
    uint32_t phase = 0;
    const float* s_f32 = reinterpret_cast<const float*>(s_w);

    const int t = *p.t;
    const int n_owned = owned_count(p);
    const int n_tiles = (n_owned + FINAL_THREADS - 1) / FINAL_THREADS;
    const int tile_end = (p.tile_hi > 0 && p.tile_hi < n_tiles) ? p.tile_hi : n_tiles;
    double racc[4] = {0, 0, 0, 0};
    long long klo = 0x7fffffffffffffffll, khi = -0x7fffffffffffffffll - 1;
    for (int tile = p.tile_lo + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int oi = tile * FINAL_THREADS + row;
        const int a = oi < n_owned ? owned_agent(p, oi) : -1;
        const bool valid = a >= 0;
        double4 st_own = make_double4(0, 0, 0, 0);
        if (CLOSED && valid && hsel == 1) st_own = ldg256(&p.state[a]);      // the second half's threads run the epilogue
        // ---- layer 0: this thread's half of the input row -> tf32 hi/lo -> its chunks of the A tile ----
        {
            const int rowoff = (row >> 3) * (K0 / 4) * 128 + (row & 7) * 16;
            float v[H0];
            if (hsel == 0) {
                if (valid) tc2_load_cols<K, 0, H0>(p, t, a, v);
                else {
#pragma unroll
                    for (int c = 0; c < H0; ++c) v[c] = 0.f;
                }
                tc2_store_chunks<H0>(s_ahi, s_alo, rowoff, 0, v);
            } else {
                if (valid) tc2_load_cols<K, H0, K0>(p, t, a, v);
                else {
#pragma unroll
                    for (int c = 0; c < H0; ++c) v[c] = 0.f;
                }
                tc2_store_chunks<H0>(s_ahi, s_alo, rowoff, H0 / 4, v);
            }
        }
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
                const uint64_t adv = (uint64_t)(ks * 16);
                tc::mma_tf32(tmem_base, a_lo + adv, b_hi + adv, IDESC, ks > 0);
                tc::mma_tf32(tmem_base, a_hi + adv, b_lo + adv, IDESC, 1);
                tc::mma_tf32(tmem_base, a_hi + adv, b_hi + adv, IDESC, 1);
            }
            tc::commit(mbar);
        }
        tc::mbar_wait(mbar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        float h[HH];
        tc::tmem_ld<HH>(tmem_row, h);
        // ---- hidden layers l = 1 .. L-1 ----
        for (int l = 1; l < p.L; ++l) {
            const float* bprev = s_f32 + tl.off_b(l - 1) / 4 + hsel * HH;
            const int rowoff = (row >> 3) * (HP / 4) * 128 + (row & 7) * 16;
#pragma unroll
            for (int i = 0; i < HH; ++i) h[i] = tanh_act(h[i] + bprev[i]);
            tc2_store_chunks<HH>(s_ahi, s_alo, rowoff, hsel * (HH / 4), h);
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
            tc::tmem_ld<HH>(tmem_row + (uint32_t)((l & 1) * HP), h);
        }
        // ---- output layer: the 2 x HP products are summed in column order exactly like the one-warp kernel does:
        //      the first half's threads start from the bias, hand their partial sums over through shared memory, the second
        //      half's threads continue the same FMA chain and run the epilogue (so actions stay bit-identical) ----
        const float* blast = s_f32 + tl.off_b(p.L - 1) / 4 + hsel * HH;
        const float2* wlp = reinterpret_cast<const float2*>(s_w + tl.off_wl()) + hsel * HH;
        const float* bl = s_f32 + tl.off_bl() / 4;
#pragma unroll
        for (int i = 0; i < HH; ++i) h[i] = tanh_act(h[i] + blast[i]);
        float o0 = 0.f, o1 = 0.f;
        if (hsel == 0) {
            o0 = bl[0];
            o1 = bl[1];
#pragma unroll
            for (int i = 0; i < HH; ++i) {
                const float2 ww = wlp[i];
                o0 = fmaf(h[i], ww.x, o0);
                o1 = fmaf(h[i], ww.y, o1);
            }
            s_part[2 * row] = o0;
            s_part[2 * row + 1] = o1;
        }
        __syncthreads();
        if (hsel == 1) {
            o0 = s_part[2 * row];
            o1 = s_part[2 * row + 1];
#pragma unroll
            for (int i = 0; i < HH; ++i) {
                const float2 ww = wlp[i];
                o0 = fmaf(h[i], ww.x, o0);
                o1 = fmaf(h[i], ww.y, o1);
            }
            if (valid) {
                reinterpret_cast<float2*>(p.action)[a] = make_float2(o0, o1);
                if (CLOSED) {
                    const double4 st_new = integrate_and_bin(p, a, st_own, o0, o1, racc);
                    if (p.fuse) shard_pack_agent(p, p.fuse->ctl, p.fuse->windows, p.fuse->wstride, p.fuse->buf, p.fuse->cap, oi, a,
                                                 st_new, klo, khi);
                }
            }
        }
    }
    if (CLOSED) reward_block_flush<FINAL2_THREADS>(p, racc);
    if (CLOSED && p.fuse) shard_interval_flush<FINAL2_THREADS>(p.fuse->ctl, klo, khi);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<TM_COLS>(tmem_base);
}
#endif  // __CUDACC__

}  // namespace fgnn
