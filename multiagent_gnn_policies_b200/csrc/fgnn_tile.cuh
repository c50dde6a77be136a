// fgnn_tile.cuh -- K_T: radius adjacency + 6-d features + the FIRST graph-shift hop, fused per cell tile (sm_100a).
//
// One CTA owns a rectangle of w x h cells of the wrapped cell grid.  Because the cell-sorted arrays are cell-major,
// every grid row of the rectangle plus its two-cell halo is ONE contiguous slot range (two at the x seam), so the
// CTA pulls (h+4) row ranges of float64 states into shared memory with TMA bulk copies (cp.async.bulk, one mbarrier)
// and everything downstream runs out of shared memory:
//
//   stage   states of the (w+4) x (h+4) cell window (TMA bulk), agent ids, and -- gathered ONCE per CTA by cp.async
//           instead of once per edge -- the source rows x_{t-1} .. x_{t-K+1} of every agent in the one-cell ring
//   phase A radius test of every agent of the (w+2) x (h+2) window against its 3x3 cells: fp32 on window-relative
//           coordinates with a rigorous margin; the rare ambiguous pair (|r2 - R2| <= margin) and every pair that
//           involves an aliased far agent falls through to the float64 test numpy evaluates, so the edge set stays
//           bit-exact.  Work item = (agent, cell row): a 32-bit accept mask per item, degree by popcount.
//           (The degrees of the ring agents are what the hop needs: z_1[n] = sum_m x_{t-1}[m] / deg_t(m).)
//   phase B one thread per OWNED agent walks its accept masks in canonical order (row -1, 0, +1; slot order inside):
//           float64 feature sums and the hop sums  z_1 = x_{t-1} A_t,  y_k = x_{t-k} A_t  (k >= 2, pre-scaled by the
//           next hop's source scale) from shared memory; emits x_t, deg, 1/deg, the ELL head and CSR rows.
//
// Neighbour order and every floating-point operation order equal those of k_adjacency_t + k_hop<NB, true>: the two
// paths leave bit-identical state (scripts/ab_variants.py checks it).
//
// Replaces, per agent-step: ~14 candidate tests in float64 with per-edge LDG gathers (k_adjacency_t) and
// 2 x d 32-byte sector gathers from L2 (k_hop) by ~20 fp32 tests + d float64 feature terms on shared memory.
// Reference arithmetic: gym_flock compute_helpers, learner/actor.py:68-71 (first product).
#pragma once
#include "fgnn_kernels.cuh"

namespace fgnn {

constexpr int TL_THREADS = 256;
constexpr int TL_TXMAX = 32;                 // largest tile, in cells
constexpr int TL_TYMAX = 16;
constexpr int TL_WMAX = TL_TXMAX + 4;
constexpr int TL_HMAX = TL_TYMAX + 4;
#ifndef FGNN_TL_CAP
#define FGNN_TL_CAP 512
#endif
constexpr int TL_CAP = FGNN_TL_CAP;          // agents staged per pass (window incl. the two-cell halo)
constexpr int TL_STACK = 24;                 // rectangles pending subdivision

// dynamic shared memory of k_tile<K>
__host__ __device__ constexpr size_t tile_smem_bytes(int K) {
    return (size_t)TL_CAP * (sizeof(double4) /* state */ + sizeof(float2) /* xy32 */ + sizeof(int) /* id */ + sizeof(int) /* deg */ +
                             sizeof(unsigned short) * 3 /* lcell, listA, listB */ + 3 * sizeof(unsigned) /* masks */ +
                             (size_t)(K > 1 ? K - 1 : 0) * ROW * sizeof(float));
}

namespace tl {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "TL_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra TL_DONE;\n\t"
        "bra TL_WAIT;\n\t"
        "TL_DONE:\n\t}\n"
        :: "r"(mbar), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP); bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
}  // namespace tl

// geometry + fp32 pre-filter thresholds of one launch (host-computed)
struct TileGeom {
    int tw, th;               // nominal tile, cells (tw + 4 <= G, th + 4 <= Gy: no window cell is staged twice)
    int ntx, nty;             // tiles per episode along x / y
    float lo32, hi32;         // r2_32 < lo32: inside for sure;  r2_32 > hi32: outside for sure
    float far32;              // |window-relative coordinate| beyond this: aliased agent, always the float64 test
    int csr_tail_only;        // 1: CSR rows only for agents with more than ELLW neighbours
};

#ifdef FGNN_MAIN_TU

template <int K>
__global__ void __launch_bounds__(TL_THREADS, 3) k_tile(Params p, TileGeom geo) {
    pdl_prologue();
    constexpr int NBR = K > 1 ? K - 1 : 0;       // source rows of the first hop: x_{t-1} .. x_{t-K+1}
    extern __shared__ __align__(128) unsigned char s_tile_raw[];
    double4* s_st = reinterpret_cast<double4*>(s_tile_raw);                       // [CAP] px,py,vx,vy
    float* s_rows = reinterpret_cast<float*>(s_st + TL_CAP);                      // [NBR][CAP][ROW]
    float2* s_xy = reinterpret_cast<float2*>(s_rows + (size_t)NBR * TL_CAP * ROW);// [CAP] window-relative fp32 position
    int* s_id = reinterpret_cast<int*>(s_xy + TL_CAP);                            // [CAP]
    int* s_deg = s_id + TL_CAP;                                                   // [CAP]
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_deg + TL_CAP);               // [3][CAP] accept masks of owned agents
    unsigned short* s_lcell = reinterpret_cast<unsigned short*>(s_mask + 3 * TL_CAP);   // [CAP] (ly << 6) | lx
    unsigned short* s_listA = s_lcell + TL_CAP;                                   // [CAP] agents of the one-cell ring window
    unsigned short* s_listB = s_listA + TL_CAP;                                   // [CAP] owned agents

    __shared__ int s_cs[TL_HMAX][TL_WMAX + 1];   // first staged index of every window cell (+ row end)
    __shared__ int s_gq[TL_HMAX][2];             // global slot of the row's segment A / segment B
    __shared__ int s_lenA[TL_HMAX];
    __shared__ int s_rowbase[TL_HMAX + 1];       // staged index of the row's first agent
    __shared__ int s_baseA[TL_HMAX + 1], s_baseB[TL_HMAX + 1];
    __shared__ int s_stack[TL_STACK][4];
    __shared__ int s_sp;
    __shared__ float s_sinvtab[64];
    __shared__ __align__(8) unsigned long long s_mbar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   // housekeeping for the next scan (what k_adjacency_t did)
        const int gs = blockIdx.x * TL_THREADS + tid;
        for (int i = gs; i < p.n_tiles; i += gridDim.x * TL_THREADS) p.tile_status[i] = 0;
        if (gs == 0) *p.tile_counter = 0;
    }
    const int tix = blockIdx.x % geo.ntx;
    const int tiy = (blockIdx.x / geo.ntx) % geo.nty;
    const int ep = blockIdx.x / (geo.ntx * geo.nty);
    if (tid == 0) {
        s_stack[0][0] = tix * geo.tw;
        s_stack[0][1] = tiy * geo.th;
        s_stack[0][2] = min(geo.tw, p.G - tix * geo.tw);
        s_stack[0][3] = min(geo.th, p.Gy - tiy * geo.th);
        s_sp = 1;
        tl::mbar_init(tl::smem_u32(&s_mbar), 1);
    }
    if (tid < 64) s_sinvtab[tid] = p.mean_pooling ? (float)(1.0 / (double)(tid > 0 ? tid : 1)) : 1.0f;
    const int t = *p.t;
    const int g = slot_of(t, K);
    const size_t M = p.M;
    uint32_t phase = 0;
    __syncthreads();

    while (true) {
        __syncthreads();                          // s_sp / stack stable; the previous pass is done with shared memory
        const int sp = s_sp;
        if (sp == 0) break;
        const int cx0 = s_stack[sp - 1][0], cy0 = s_stack[sp - 1][1], w = s_stack[sp - 1][2], h = s_stack[sp - 1][3];
        const int W = w + 4, H = h + 4;
        // ---- row ranges of the window -------------------------------------------------------------
        int cxs = cx0 - 2;
        if (cxs < 0) cxs += p.G;
        const int la = min(W, p.G - cxs);         // window cells before the x seam
        if (tid < H) {
            int wy = cy0 - 2 + tid;
            if (wy < 0) wy += p.Gy;
            if (wy >= p.Gy) wy -= p.Gy;
            const int rowbase = (ep * p.Gy + wy) * p.G;
            const int a0 = __ldg(&p.cell_start[rowbase + cxs]);
            const int a1 = __ldg(&p.cell_start[rowbase + cxs + la]);
            int b0 = 0, b1 = 0;
            if (la < W) {
                b0 = __ldg(&p.cell_start[rowbase]);
                b1 = __ldg(&p.cell_start[rowbase + W - la]);
            }
            s_gq[tid][0] = a0;
            s_gq[tid][1] = b0;
            s_lenA[tid] = a1 - a0;
            s_rowbase[tid + 1] = (a1 - a0) + (b1 - b0);      // length for now
        }
        __syncthreads();
        if (warp == 0) {                          // exclusive scan of the row lengths (H <= 32)
            const int len = lane < H ? s_rowbase[lane + 1] : 0;
            int inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            __syncwarp();
            if (lane < H) s_rowbase[lane + 1] = inc;
            if (lane == 0) s_rowbase[0] = 0;
        }
        __syncthreads();
        const int total = s_rowbase[H];
        // owned rows are window rows 2 .. 2+h-1, owned columns 2 .. 2+w-1
        if (total > TL_CAP) {                     // too crowded for one pass: halve the rectangle (uniform decision)
            __syncthreads();
            if (tid == 0) {
                int spn = sp - 1;
                if (w > 1 || h > 1) {
                    if (spn + 2 > TL_STACK) {
                        *p.overflow = 1;          // cannot happen: depth <= log2(TXMAX * TYMAX) + 1
                    } else if (w >= h) {
                        const int w0 = w / 2;
                        s_stack[spn][0] = cx0; s_stack[spn][1] = cy0; s_stack[spn][2] = w0; s_stack[spn][3] = h;
                        s_stack[spn + 1][0] = cx0 + w0; s_stack[spn + 1][1] = cy0; s_stack[spn + 1][2] = w - w0; s_stack[spn + 1][3] = h;
                        spn += 2;
                    } else {
                        const int h0 = h / 2;
                        s_stack[spn][0] = cx0; s_stack[spn][1] = cy0; s_stack[spn][2] = w; s_stack[spn][3] = h0;
                        s_stack[spn + 1][0] = cx0; s_stack[spn + 1][1] = cy0 + h0; s_stack[spn + 1][2] = w; s_stack[spn + 1][3] = h - h0;
                        spn += 2;
                    }
                } else {
                    *p.overflow = 1;              // a single cell whose 5x5 neighbourhood exceeds the stage: results void
                }
                s_sp = spn;
            }
            continue;
        }
        if (tid == 0) s_sp = sp - 1;              // read again only after the barrier at the loop head
        // ---- TMA: the row ranges of float64 states -> s_st -------------------------------------------
        if (tid == 0 && total > 0) {
            const uint32_t mbar = tl::smem_u32(&s_mbar);
            tl::mbar_expect_tx(mbar, (uint32_t)total * (uint32_t)sizeof(double4));
            for (int r = 0; r < H; ++r) {
                const int lenA = s_lenA[r];
                const int lenB = s_rowbase[r + 1] - s_rowbase[r] - lenA;
                if (lenA > 0)
                    tl::bulk_g2s(tl::smem_u32(s_st + s_rowbase[r]), p.sorted_state + s_gq[r][0], (uint32_t)lenA * 32u, mbar);
                if (lenB > 0)
                    tl::bulk_g2s(tl::smem_u32(s_st + s_rowbase[r] + lenA), p.sorted_state + s_gq[r][1], (uint32_t)lenB * 32u, mbar);
            }
        }
        // ---- first staged index of every window cell ---------------------------------------------------
        for (int e = tid; e < H * (W + 1); e += TL_THREADS) {
            const int r = e / (W + 1), lx = e - r * (W + 1);
            int wy = cy0 - 2 + r;
            if (wy < 0) wy += p.Gy;
            if (wy >= p.Gy) wy -= p.Gy;
            const int rowbase = (ep * p.Gy + wy) * p.G;
            int v;
            if (lx < la || la == W) v = __ldg(&p.cell_start[rowbase + cxs + lx]) - s_gq[r][0];
            else v = s_lenA[r] + __ldg(&p.cell_start[rowbase + lx - la]) - s_gq[r][1];
            s_cs[r][lx] = s_rowbase[r] + v;
        }
        __syncthreads();
        if (warp == 0) {                          // list offsets: ring window rows 1 .. H-2 (cols 1 .. W-2), owned rows 2 .. H-3
            const int nA = (lane >= 1 && lane <= H - 2) ? s_cs[lane][W - 1] - s_cs[lane][1] : 0;
            const int nB = (lane >= 2 && lane <= H - 3) ? s_cs[lane][W - 2] - s_cs[lane][2] : 0;
            int ia = nA, ib = nB;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ya = __shfl_up_sync(0xffffffffu, ia, o);
                const int yb = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) { ia += ya; ib += yb; }
            }
            if (lane < H) { s_baseA[lane + 1] = ia; s_baseB[lane + 1] = ib; }
            if (lane == 0) { s_baseA[0] = 0; s_baseB[0] = 0; }
        }
        __syncthreads();
        const int nA = s_baseA[H], nB = s_baseB[H];
        if (nB == 0) {                            // nothing owned here (the bulk copies must still land before reuse)
            if (total > 0) { tl::mbar_wait(tl::smem_u32(&s_mbar), phase); phase ^= 1; }
            continue;
        }
        // ---- per cell: ids, local cell, work lists, source rows of the first hop (cp.async, once per CTA) ---------
        for (int e = tid; e < H * W; e += TL_THREADS) {
            const int r = e / W, lx = e - r * W;
            const int i0 = s_cs[r][lx], i1 = s_cs[r][lx + 1];
            if (i0 == i1) continue;
            const int rel = i0 - s_rowbase[r];
            const int q0 = rel < s_lenA[r] || la == W ? s_gq[r][0] + rel : s_gq[r][1] + rel - s_lenA[r];
            const bool ring = r >= 1 && r <= H - 2 && lx >= 1 && lx <= W - 2;
            const bool own = r >= 2 && r <= H - 3 && lx >= 2 && lx <= W - 3;
            for (int i = i0; i < i1; ++i) {
                const int id = __ldg(&p.sorted_id[q0 + (i - i0)]);
                s_id[i] = id;
                s_lcell[i] = (unsigned short)((r << 6) | lx);
                s_deg[i] = 0;
                if (ring) {
                    s_listA[s_baseA[r] + i - s_cs[r][1]] = (unsigned short)i;
#pragma unroll
                    for (int b = 0; b < NBR; ++b) {
                        const float* src = p.xhist + ((size_t)slot_of(t - 1 - b, K) * M + id) * ROW;
                        const uint32_t dst = tl::smem_u32(s_rows + ((size_t)b * TL_CAP + i) * ROW);
                        tl::cp_async16(dst, src);
                        tl::cp_async16(dst + 16, src + 4);
                    }
                }
                if (own) s_listB[s_baseB[r] + i - s_cs[r][2]] = (unsigned short)i;
            }
        }
        // ---- states have landed: window-relative fp32 coordinates -------------------------------------------
        tl::mbar_wait(tl::smem_u32(&s_mbar), phase);
        phase ^= 1;
        {
            const double ox = s_st[0].x, oy = s_st[0].y;   // staged agent 0 defines the frame (any agent of the window would do)
            for (int i = tid; i < total; i += TL_THREADS) {
                const double2 pxy = *reinterpret_cast<const double2*>(&s_st[i]);
                float rx = (float)(pxy.x - ox), ry = (float)(pxy.y - oy);
                if (!(fabsf(rx) <= geo.far32) || !(fabsf(ry) <= geo.far32)) rx = __int_as_float(0x7fc00000);   // aliased: NaN
                s_xy[i] = make_float2(rx, ry);
            }
        }
        __syncthreads();
        // ---- phase A: accept masks + degrees, work item = (ring agent, cell row) --------------------------------
        for (int it = tid; it < 3 * nA; it += TL_THREADS) {
            const int dyi = (it >= nA ? 1 : 0) + (it >= 2 * nA ? 1 : 0);
            const int i = s_listA[it - dyi * nA];
            const unsigned lc = s_lcell[i];
            const int lx = lc & 63, ly = lc >> 6;
            const int r = ly + dyi - 1;
            const int c0 = s_cs[r][lx - 1], c1 = s_cs[r][lx + 2];
            const float2 me = s_xy[i];
            int cnt = 0;
            unsigned first = 0;
            for (int base = c0; base < c1; base += 32) {
                const int n = min(32, c1 - base);
                unsigned acc = 0, amb = 0;
                for (int k = 0; k < n; ++k) {
                    const float2 o = s_xy[base + k];
                    const float dx = me.x - o.x, dy = me.y - o.y;
                    const float r2 = fmaf(dx, dx, dy * dy);
                    const bool in = r2 < geo.lo32;
                    acc |= (in ? 1u : 0u) << k;
                    amb |= ((!in && !(r2 > geo.hi32)) ? 1u : 0u) << k;
                }
                if (dyi == 1 && i >= base && i < base + 32) {       // self
                    acc &= ~(1u << (i - base));
                    amb &= ~(1u << (i - base));
                }
                while (amb) {                                        // rare: the float64 test numpy evaluates
                    const int k = __ffs(amb) - 1;
                    amb &= amb - 1;
                    const double2 a2 = *reinterpret_cast<const double2*>(&s_st[i]);
                    const double2 o2 = *reinterpret_cast<const double2*>(&s_st[base + k]);
                    if (r2_exact(a2.x - o2.x, a2.y - o2.y) < p.R2) acc |= 1u << k;
                }
                cnt += __popc(acc);
                if (base == c0) first = acc;
            }
            if (cnt) atomicAdd(&s_deg[i], cnt);
            const bool own = ly >= 2 && ly <= H - 3 && lx >= 2 && lx <= W - 3;
            if (own) s_mask[dyi * TL_CAP + (s_baseB[ly] + i - s_cs[ly][2])] = first;
        }
        tl::cp_async_wait_all();
        __syncthreads();
        // ---- phase B: one thread per owned agent --------------------------------------------------------------
        for (int kb0 = 0; kb0 < nB; kb0 += TL_THREADS) {
            const int kb = kb0 + tid;
            const bool valid = kb < nB;
            int i = 0, lx = 2, ly = 2, count = 0, a = 0;
            double4 me = make_double4(0, 0, 0, 0);
            double f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0;
            float acc[NBR > 0 ? NBR : 1][F];
#pragma unroll
            for (int b = 0; b < (NBR > 0 ? NBR : 1); ++b)
#pragma unroll
                for (int f = 0; f < F; ++f) acc[b][f] = 0.f;
            if (valid) {
                i = s_listB[kb];
                const unsigned lc = s_lcell[i];
                lx = lc & 63; ly = lc >> 6;
                a = s_id[i];
                me = s_st[i];
                count = s_deg[i];
            }
            float sn = 0.f;
            if (valid && NBR > 1) sn = __ldg(&p.sinv[(size_t)slot_of(t - 1, K) * M + a]);
            if (valid) {
                const float2 me32 = s_xy[i];
#pragma unroll 1
                for (int dyi = 0; dyi < 3; ++dyi) {
                    const int r = ly + dyi - 1;
                    const int c0 = s_cs[r][lx - 1], c1 = s_cs[r][lx + 2];
                    for (int base = c0; base < c1; base += 32) {
                        unsigned m;
                        if (c1 - c0 <= 32) {
                            m = s_mask[dyi * TL_CAP + kb];
                        } else {                                     // long cell row: the masks beyond the first are recomputed
                            const int n = min(32, c1 - base);
                            unsigned amb = 0;
                            m = 0;
                            for (int k = 0; k < n; ++k) {
                                const float2 o = s_xy[base + k];
                                const float dx = me32.x - o.x, dy = me32.y - o.y;
                                const float r2 = fmaf(dx, dx, dy * dy);
                                const bool in = r2 < geo.lo32;
                                m |= (in ? 1u : 0u) << k;
                                amb |= ((!in && !(r2 > geo.hi32)) ? 1u : 0u) << k;
                            }
                            if (dyi == 1 && i >= base && i < base + 32) {
                                m &= ~(1u << (i - base));
                                amb &= ~(1u << (i - base));
                            }
                            while (amb) {
                                const int k = __ffs(amb) - 1;
                                amb &= amb - 1;
                                const double2 o2 = *reinterpret_cast<const double2*>(&s_st[base + k]);
                                if (r2_exact(me.x - o2.x, me.y - o2.y) < p.R2) m |= 1u << k;
                            }
                        }
                        while (m) {
                            const int j = base + __ffs(m) - 1;
                            m &= m - 1;
                            const double4 o = s_st[j];
                            const double dx = me.x - o.x, dy = me.y - o.y;
                            const double r2 = r2_exact(dx, dy);
                            const double inv = fast_rcp(r2);
                            const double inv2 = inv * inv;
                            f0 += me.z - o.z;
                            f1 += dx * inv2;
                            f2 += dx * inv;
                            f3 += me.w - o.w;
                            f4 += dy * inv2;
                            f5 += dy * inv;
                            if (NBR > 0) {
                                const int dj = s_deg[j];
                                const float sc = dj < 64 ? s_sinvtab[dj] : (p.mean_pooling ? (float)(1.0 / (double)dj) : 1.0f);
#pragma unroll
                                for (int b = 0; b < NBR; ++b) {
                                    const float* row = s_rows + ((size_t)b * TL_CAP + j) * ROW;
                                    const float4 v0 = *reinterpret_cast<const float4*>(row);
                                    const float2 v1 = *reinterpret_cast<const float2*>(row + 4);
                                    acc[b][0] = fmaf(v0.x, sc, acc[b][0]);
                                    acc[b][1] = fmaf(v0.y, sc, acc[b][1]);
                                    acc[b][2] = fmaf(v0.z, sc, acc[b][2]);
                                    acc[b][3] = fmaf(v0.w, sc, acc[b][3]);
                                    acc[b][4] = fmaf(v1.x, sc, acc[b][4]);
                                    acc[b][5] = fmaf(v1.y, sc, acc[b][5]);
                                }
                            }
                        }
                    }
                }
            }
            // ---- CSR rows: one contiguous run of edge slots per warp (all rows, or only the rows longer than the ELL head)
            const bool wants_row = valid && (geo.csr_tail_only ? count > ELLW : count > 0);
            int inc = wants_row ? count : 0;
            const int mine = inc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            const int wtotal = __shfl_sync(0xffffffffu, inc, 31);
            unsigned rbase = 0;
            if (lane == 31 && wtotal > 0) rbase = atomicAdd(&p.nnz_cursor[g], (unsigned)wtotal);
            rbase = __shfl_sync(0xffffffffu, rbase, 31);
            if (!valid) continue;
            unsigned row = rbase + (unsigned)(inc - mine);
            bool write_row = wants_row;
            if (wants_row && (row + (unsigned)count > p.nnz_cap || row + (unsigned)count < row)) {   // capacity exceeded: drop the row, flag it
                *p.overflow = 1;
                count = 0;
                row = 0;
                write_row = false;
            }
            if (!wants_row) row = 0;
            // second walk over the masks: neighbour ids, in the same order
            int head[ELLW];
#pragma unroll
            for (int e = 0; e < ELLW; ++e) head[e] = -1;
            {
                int* cols = p.cols + (size_t)g * p.nnz_cap + row;
                int e = 0;
                const float2 me32 = s_xy[i];
#pragma unroll 1
                for (int dyi = 0; dyi < 3 && e < count; ++dyi) {
                    const int r = ly + dyi - 1;
                    const int c0 = s_cs[r][lx - 1], c1 = s_cs[r][lx + 2];
                    for (int base = c0; base < c1; base += 32) {
                        unsigned m;
                        if (c1 - c0 <= 32) {
                            m = s_mask[dyi * TL_CAP + kb];
                        } else {
                            const int n = min(32, c1 - base);
                            unsigned amb = 0;
                            m = 0;
                            for (int k = 0; k < n; ++k) {
                                const float2 o = s_xy[base + k];
                                const float dx = me32.x - o.x, dy = me32.y - o.y;
                                const float r2 = fmaf(dx, dx, dy * dy);
                                const bool in = r2 < geo.lo32;
                                m |= (in ? 1u : 0u) << k;
                                amb |= ((!in && !(r2 > geo.hi32)) ? 1u : 0u) << k;
                            }
                            if (dyi == 1 && i >= base && i < base + 32) {
                                m &= ~(1u << (i - base));
                                amb &= ~(1u << (i - base));
                            }
                            while (amb) {
                                const int k = __ffs(amb) - 1;
                                amb &= amb - 1;
                                const double2 o2 = *reinterpret_cast<const double2*>(&s_st[base + k]);
                                if (r2_exact(me.x - o2.x, me.y - o2.y) < p.R2) m |= 1u << k;
                            }
                        }
                        while (m) {
                            const int j = base + __ffs(m) - 1;
                            m &= m - 1;
                            const int id = s_id[j];
                            if (write_row) cols[e] = id;
#pragma unroll
                            for (int u = 0; u < ELLW; ++u)
                                if (u == e) head[u] = id;
                            ++e;
                        }
                    }
                }
            }
            const size_t ga = (size_t)g * M + a;
            stg256(p.ell + ga * ELLW, head);
            stg256(p.xhist + ga * ROW, (float)f0, (float)f1, (float)f2, (float)f3, (float)f4, (float)f5, 0.f, 0.f);
            p.deg[ga] = count;
            p.row_start[ga] = row;
            p.sinv[ga] = count < 64 ? s_sinvtab[count] : (p.mean_pooling ? (float)(1.0 / (double)count) : 1.0f);
            if (NBR > 0) {
                store_row6(p.zbuf + (size_t)1 * M * ROW, a, acc[0]);
#pragma unroll
                for (int b = 1; b < NBR; ++b) {
#pragma unroll
                    for (int f = 0; f < F; ++f) acc[b][f] *= sn;
                    store_row6(p.ybuf + (size_t)(1 + b) * M * ROW, a, acc[b]);
                }
            }
        }
    }
}

#endif  // FGNN_MAIN_TU

}  // namespace fgnn
