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

#ifndef FGNN_TL_THREADS
#define FGNN_TL_THREADS 256
#endif
#ifndef FGNN_TL_MINBLOCKS
#define FGNN_TL_MINBLOCKS 3
#endif
constexpr int TL_THREADS = FGNN_TL_THREADS;
constexpr int TL_TXMAX = 24;                 // largest tile, in cells
constexpr int TL_TYMAX = 12;
constexpr int TL_NBR = 12;                   // accepted neighbours staged per owned agent (longer rows re-run the filter for the rest)
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
                             sizeof(unsigned short) * 3 /* lcell, listA, listB */ +
                             TL_NBR * sizeof(unsigned short) /* staged neighbour lists */ +
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
    float lo32, hi32n;        // r2_32 < lo32: inside for sure;  r2_32 >= hi32n: outside for sure
    float far32;              // |window-relative coordinate| beyond this: aliased agent, always the float64 test
    int csr_tail_only;        // 1: CSR rows only for agents with more than ELLW neighbours
    float sinvtab[64];        // source scale by degree: (float)(1.0 / max(d, 1)) under mean pooling, else 1
};

#ifdef FGNN_MAIN_TU

// fp32 radius test of agent `me` against the staged agents [base, base + n), 1 <= n <= 32: bit k of the result = inside.
// Per candidate: r2 in fp32 (4 instructions) and two funnel shifts that push the SIGN bits of r2 - lo and r2 - hi' into two
// masks (no compare / select / variable shift).  Pairs whose fp32 r2 falls inside the margin around R^2 (and, for an
// aliased `me`, every pair) take the float64 test numpy evaluates.
__device__ __forceinline__ unsigned tile_filter(const float2* __restrict__ s_xy, const double4* __restrict__ s_st, const TileGeom& geo,
                                                double R2, int i, float2 me, bool me_far, int base, int n) {
    unsigned in = 0, le = 0;
    if (!me_far) {
        const float2* c = s_xy + base;
        const float lo = geo.lo32, hi = geo.hi32n;
        int k = 0;
#pragma unroll 1
        for (; k + 2 <= n; k += 2) {              // rows hold ~5 candidates: a short rolled loop beats a deep unroll cascade
            const float2 o0 = c[k], o1 = c[k + 1];
            const float dx0 = me.x - o0.x, dy0 = me.y - o0.y, dx1 = me.x - o1.x, dy1 = me.y - o1.y;
            const float r20 = fmaf(dx0, dx0, dy0 * dy0), r21 = fmaf(dx1, dx1, dy1 * dy1);
            in = __funnelshift_l(__float_as_uint(r20 - lo), in, 1);             // sign set <=> r2 < lo
            le = __funnelshift_l(__float_as_uint(r20 - hi), le, 1);             // sign set <=> r2 <= hi32
            in = __funnelshift_l(__float_as_uint(r21 - lo), in, 1);
            le = __funnelshift_l(__float_as_uint(r21 - hi), le, 1);
        }
        if (k < n) {
            const float2 o0 = c[k];
            const float dx0 = me.x - o0.x, dy0 = me.y - o0.y;
            const float r20 = fmaf(dx0, dx0, dy0 * dy0);
            in = __funnelshift_l(__float_as_uint(r20 - lo), in, 1);
            le = __funnelshift_l(__float_as_uint(r20 - hi), le, 1);
        }
        in = __brev(in) >> (32 - n);              // candidate k was shifted in first: it sits at bit n-1-k
        le = __brev(le) >> (32 - n);
    } else {
        le = 0xffffffffu >> (32 - n);
    }
    unsigned amb = le & ~in;
    if (i >= base && i < base + n) {             // self
        in &= ~(1u << (i - base));
        amb &= ~(1u << (i - base));
    }
    if (amb) {
        const double2 a2 = *reinterpret_cast<const double2*>(&s_st[i]);
        do {
            const int k = __ffs(amb) - 1;
            amb &= amb - 1;
            const double2 o2 = *reinterpret_cast<const double2*>(&s_st[base + k]);
            if (r2_exact(a2.x - o2.x, a2.y - o2.y) < R2) in |= 1u << k;
        } while (amb);
    }
    return in;
}

template <int K>
__global__ void __launch_bounds__(TL_THREADS, FGNN_TL_MINBLOCKS) k_tile(Params p, TileGeom geo) {
    pdl_prologue();
    constexpr int NBR = K > 1 ? K - 1 : 0;       // source rows of the first hop: x_{t-1} .. x_{t-K+1}
    constexpr int CS_PER_THREAD = (TL_HMAX * (TL_WMAX + 1) + TL_THREADS - 1) / TL_THREADS;
    extern __shared__ __align__(128) unsigned char s_tile_raw[];
    double4* s_st = reinterpret_cast<double4*>(s_tile_raw);                       // [CAP] px,py,vx,vy
    float* s_rows = reinterpret_cast<float*>(s_st + TL_CAP);                      // [NBR][CAP][ROW]
    float2* s_xy = reinterpret_cast<float2*>(s_rows + (size_t)NBR * TL_CAP * ROW);// [CAP] window-relative fp32 position
    int* s_id = reinterpret_cast<int*>(s_xy + TL_CAP);                            // [CAP]
    int* s_deg = s_id + TL_CAP;                                                   // [CAP]
    unsigned short* s_lcell = reinterpret_cast<unsigned short*>(s_deg + TL_CAP);  // [CAP] far << 15 | ly << 6 | lx
    unsigned short* s_listA = s_lcell + TL_CAP;                                   // [CAP] agents of the one-cell ring window
    unsigned short* s_listB = s_listA + TL_CAP;                                   // [CAP] owned agents
    unsigned short* s_nbr = s_listB + TL_CAP;                                     // [TL_NBR][CAP] staged index of the e-th neighbour of owned agent kb

    __shared__ int s_cs[TL_HMAX][TL_WMAX + 1];   // first staged index of every window cell (+ row end)
    __shared__ int s_endA[TL_HMAX];              // x seam: global slot where the row's first segment ends
    __shared__ int s_gq[TL_HMAX][2];             // global slot of the row's segment A / segment B
    __shared__ int s_lenA[TL_HMAX];
    __shared__ int s_rowbase[TL_HMAX + 1];       // staged index of the row's first agent
    __shared__ int s_baseA[TL_HMAX + 1], s_baseB[TL_HMAX + 1];
    __shared__ int s_stack[TL_STACK][4];
    __shared__ int s_sp;
    __shared__ float s_sinvtab[64];
    __shared__ __align__(8) unsigned long long s_mbar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tix = blockIdx.x, tiy = blockIdx.y, ep = blockIdx.z;      // grid = (ntx, nty, episodes)
    {   // housekeeping for the next scan (what k_adjacency_t did)
        const int gs = ((ep * gridDim.y + tiy) * gridDim.x + tix) * TL_THREADS + tid;
        for (int i = gs; i < p.n_tiles; i += gridDim.x * gridDim.y * gridDim.z * TL_THREADS) p.tile_status[i] = 0;
        if (gs == 0) *p.tile_counter = 0;
    }
    if (tid == 0) {
        s_stack[0][0] = tix * geo.tw;
        s_stack[0][1] = tiy * geo.th;
        s_stack[0][2] = min(geo.tw, p.G - tix * geo.tw);
        s_stack[0][3] = min(geo.th, p.Gy - tiy * geo.th);
        s_sp = 1;
        tl::mbar_init(tl::smem_u32(&s_mbar), 1);
    }
    if (tid < 64) s_sinvtab[tid] = geo.sinvtab[tid];
    const int t = *p.t;
    const int g = slot_of(t, K);
    const size_t M = p.M;
    uint32_t phase = 0;

    while (true) {
        __syncthreads();                          // s_sp / stack stable; the previous pass is done with shared memory
        const int sp = s_sp;
        if (sp == 0) break;
        const int cx0 = s_stack[sp - 1][0], cy0 = s_stack[sp - 1][1], w = s_stack[sp - 1][2], h = s_stack[sp - 1][3];
        const int W = w + 4, H = h + 4;
        int cxs = cx0 - 2;
        if (cxs < 0) cxs += p.G;
        const int la = min(W, p.G - cxs);         // window cells before the x seam
        const unsigned magic = (unsigned)__fdividef(65536.0f, (float)(W + 1)) + 1u;   // e / (W + 1) = (e * magic) >> 16 for e < 1024, W + 1 <= 37
        // ---- P1: global slot of every window cell (one load per thread for the nominal tile) ----------------
        int craw[CS_PER_THREAD];
#pragma unroll
        for (int u = 0; u < CS_PER_THREAD; ++u) {
            const int e = tid + u * TL_THREADS;
            craw[u] = 0;
            if (e < H * (W + 1)) {
                const int r = (int)(((unsigned)e * magic) >> 16), lx = e - r * (W + 1);
                int wy = cy0 - 2 + r;
                if (wy < 0) wy += p.Gy;
                if (wy >= p.Gy) wy -= p.Gy;
                const int rowbase = (ep * p.Gy + wy) * p.G;
                const bool segA = lx < la || la == W;
                craw[u] = __ldg(&p.cell_start[segA ? rowbase + cxs + lx : rowbase + lx - la]);
                s_cs[r][lx] = craw[u];
                if (lx == 0 && la < W) s_endA[r] = __ldg(&p.cell_start[rowbase + p.G]);
            }
        }
        __syncthreads();
        // ---- P2: row lengths -> staged row bases, list offsets (one warp) ------------------------------------
        if (warp == 0) {
            int lenA = 0, lenB = 0, nA = 0, nB = 0;
            if (lane < H) {
                const int* c = s_cs[lane];
                const int a0 = c[0];
                if (la == W) {
                    lenA = c[W] - a0;
                    if (lane >= 1 && lane <= H - 2) nA = c[W - 1] - c[1];
                    if (lane >= 2 && lane <= H - 3) nB = c[W - 2] - c[2];
                } else {
                    const int endA = s_endA[lane];
                    lenA = endA - a0;
                    lenB = c[W] - c[la];
                    // agents of the cells [x0, x1) of a row that crosses the seam at cell la
                    auto span = [&](int x0, int x1) {          // (c[la] is the START of the second segment, not the end of the first)
                        const int inA = x0 < la ? (x1 < la ? c[x1] : endA) - c[x0] : 0;
                        const int inB = x1 > la ? c[x1] - c[x0 > la ? x0 : la] : 0;
                        return inA + inB;
                    };
                    if (lane >= 1 && lane <= H - 2) nA = span(1, W - 1);
                    if (lane >= 2 && lane <= H - 3) nB = span(2, W - 2);
                }
                s_gq[lane][0] = a0;
                s_gq[lane][1] = la < W ? c[la] : 0;
                s_lenA[lane] = lenA;
            }
            int il = lenA + lenB, ia = nA, ib = nB;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int yl = __shfl_up_sync(0xffffffffu, il, o);
                const int ya = __shfl_up_sync(0xffffffffu, ia, o);
                const int yb = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) { il += yl; ia += ya; ib += yb; }
            }
            if (lane < H) { s_rowbase[lane + 1] = il; s_baseA[lane + 1] = ia; s_baseB[lane + 1] = ib; }
            if (lane == 0) { s_rowbase[0] = 0; s_baseA[0] = 0; s_baseB[0] = 0; }
        }
        __syncthreads();
        const int total = s_rowbase[H];
        const int nA = s_baseA[H], nB = s_baseB[H];
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
        if (nB == 0) continue;                    // nothing owned in this rectangle
        // ---- P3: TMA bulk copies of the row ranges of float64 states; cell table global slots -> staged indices -------
        if (tid == 0) {
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
#pragma unroll
        for (int u = 0; u < CS_PER_THREAD; ++u) {
            const int e = tid + u * TL_THREADS;
            if (e < H * (W + 1)) {
                const int r = (int)(((unsigned)e * magic) >> 16), lx = e - r * (W + 1);
                const bool segA = lx < la || la == W;
                s_cs[r][lx] = s_rowbase[r] + (segA ? craw[u] - s_gq[r][0] : s_lenA[r] + craw[u] - s_gq[r][1]);
            }
        }
        // ---- P4: one warp per window row: ids (coalesced), then -- states landed -- local cell, fp32 coordinates, work
        //      lists and the source rows of the first hop (cp.async, fetched once per CTA instead of once per edge) -------
        constexpr int NWARP = TL_THREADS / 32;
        constexpr int ROWS_PER_WARP = (TL_HMAX + NWARP - 1) / NWARP;
        int idv[ROWS_PER_WARP][2];                // ids of this lane's first two agents of rows warp, warp + NWARP, ...
#pragma unroll
        for (int s_ = 0; s_ < ROWS_PER_WARP; ++s_) {
            const int r = warp + NWARP * s_;
            idv[s_][0] = idv[s_][1] = -1;
            if (r < H) {
                const int n = s_rowbase[r + 1] - s_rowbase[r], lenA = s_lenA[r];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int j = lane + 32 * v;
                    if (j < n) idv[s_][v] = __ldg(&p.sorted_id[j < lenA ? s_gq[r][0] + j : s_gq[r][1] + j - lenA]);
                }
            }
        }
        __syncthreads();                          // cell table final
        tl::mbar_wait(tl::smem_u32(&s_mbar), phase);
        phase ^= 1;
        {
            // staged agent 0 defines the frame: window-relative fp32 coordinates, and -- for every agent of the same wrap of the
            // grid -- the local column from the difference of the UNWRAPPED cell indices (no remainder by G per agent)
            const double ox = s_st[0].x, oy = s_st[0].y;
            long long ix0, iy0;
            cell_coords(p, ox, oy, ix0, iy0);
            int lx0 = wrap(ix0, p.G) - cxs;
            if (lx0 < 0) lx0 += p.G;
            auto stage_agent = [&](int r, int i, int id, int offA, int offB) {
                const double2 pxy = *reinterpret_cast<const double2*>(&s_st[i]);
                const float rx = (float)(pxy.x - ox), ry = (float)(pxy.y - oy);
                const bool far = !(fabsf(rx) <= geo.far32) || !(fabsf(ry) <= geo.far32);   // aliased from another wrap of the grid
                long long ix, iy;
                cell_coords(p, pxy.x, pxy.y, ix, iy);
                int lx = (int)(ix - ix0) + lx0;
                if (far || (unsigned)lx >= (unsigned)W) {          // another wrap of the grid (possible when it is small): full remainder
                    lx = wrap(ix, p.G) - cxs;
                    if (lx < 0) lx += p.G;
                }
                s_xy[i] = make_float2(rx, ry);
                s_id[i] = id;
                s_lcell[i] = (unsigned short)((far ? 0x8000 : 0) | (r << 6) | lx);
                const bool ring = r >= 1 && r <= H - 2 && lx >= 1 && lx <= W - 2;
                if (ring) {
                    s_listA[offA + i] = (unsigned short)i;
#pragma unroll
                    for (int b = 0; b < NBR; ++b) {
                        const float* src = p.xhist + ((size_t)slot_of(t - 1 - b, K) * M + id) * ROW;
                        const uint32_t dst = tl::smem_u32(s_rows + (b * TL_CAP + i) * ROW);
                        tl::cp_async16(dst, src);
                        tl::cp_async16(dst + 16, src + 4);
                    }
                    if (r >= 2 && r <= H - 3 && lx >= 2 && lx <= W - 3) s_listB[offB + i] = (unsigned short)i;
                }
            };
#pragma unroll
            for (int s_ = 0; s_ < ROWS_PER_WARP; ++s_) {
                const int r = warp + NWARP * s_;
                if (r < H) {
                    const int rb = s_rowbase[r], n = s_rowbase[r + 1] - rb, lenA = s_lenA[r];
                    const int offA = s_baseA[r] - s_cs[r][1], offB = s_baseB[r] - s_cs[r][2];   // list position = staged index + offset
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int j = lane + 32 * v;
                        if (j < n) stage_agent(r, rb + j, idv[s_][v], offA, offB);
                    }
                    for (int j = lane + 64; j < n; j += 32)      // crowded row
                        stage_agent(r, rb + j, __ldg(&p.sorted_id[j < lenA ? s_gq[r][0] + j : s_gq[r][1] + j - lenA]), offA, offB);
                }
            }
        }
        __syncthreads();
        // ---- phase A: radius test of every agent of the ring window against its 3x3 cells (three cell rows in turn): degree;
        //      owned agents also keep the staged indices of their first TL_NBR neighbours, in canonical order ----------------
        for (int kk = tid; kk < nA; kk += TL_THREADS) {
            const int i = s_listA[kk];
            const unsigned lc = s_lcell[i];
            const int lx = lc & 63, ly = (lc >> 6) & 31;
            const bool far = (lc & 0x8000) != 0;
            const bool own = ly >= 2 && ly <= H - 3 && lx >= 2 && lx <= W - 3;
            unsigned short* nb = s_nbr + (own ? s_baseB[ly] + i - s_cs[ly][2] : 0);
            const float2 me = s_xy[i];
            const int c00 = s_cs[ly - 1][lx - 1], c10 = s_cs[ly][lx - 1], c20 = s_cs[ly + 1][lx - 1];
            const int n0 = s_cs[ly - 1][lx + 2] - c00, n1 = s_cs[ly][lx + 2] - c10, n2 = s_cs[ly + 1][lx + 2] - c20;
            const int n = n0 + n1 + n2;
            int cnt = 0;
            if (n <= 32 && !far) {
                // common case: ONE pair of sign masks over the three cell rows (candidate b of the concatenation = bit b):
                // per candidate 4 fp32 instructions + 2 subtractions + 2 funnel shifts, nothing per row but the loop itself
                unsigned in = 0, le = 0;
                const float lo = geo.lo32, hi = geo.hi32n;
                auto push = [&](int c0, int nr) {
                    const float2* c = s_xy + c0;
                    int k = 0;
#pragma unroll 1
                    for (; k + 2 <= nr; k += 2) {
                        const float2 o0 = c[k], o1 = c[k + 1];
                        const float dx0 = me.x - o0.x, dy0 = me.y - o0.y, dx1 = me.x - o1.x, dy1 = me.y - o1.y;
                        const float r20 = fmaf(dx0, dx0, dy0 * dy0), r21 = fmaf(dx1, dx1, dy1 * dy1);
                        in = __funnelshift_l(__float_as_uint(r20 - lo), in, 1);
                        le = __funnelshift_l(__float_as_uint(r20 - hi), le, 1);
                        in = __funnelshift_l(__float_as_uint(r21 - lo), in, 1);
                        le = __funnelshift_l(__float_as_uint(r21 - hi), le, 1);
                    }
                    if (k < nr) {
                        const float2 o0 = c[k];
                        const float dx0 = me.x - o0.x, dy0 = me.y - o0.y;
                        const float r20 = fmaf(dx0, dx0, dy0 * dy0);
                        in = __funnelshift_l(__float_as_uint(r20 - lo), in, 1);
                        le = __funnelshift_l(__float_as_uint(r20 - hi), le, 1);
                    }
                };
                push(c00, n0);
                push(c10, n1);
                push(c20, n2);
                if (n > 0) {
                    in = __brev(in) >> (32 - n);             // the first candidate was shifted in first
                    le = __brev(le) >> (32 - n);
                    const unsigned self = 1u << (n0 + i - c10);
                    in &= ~self;
                    unsigned amb = le & ~in & ~self;
                    auto staged = [&](int b) { return b < n0 ? c00 + b : (b < n0 + n1 ? c10 + b - n0 : c20 + b - n0 - n1); };
                    if (amb) {                               // rare: the float64 test numpy evaluates
                        const double2 a2 = *reinterpret_cast<const double2*>(&s_st[i]);
                        do {
                            const int b = __ffs(amb) - 1;
                            amb &= amb - 1;
                            const double2 o2 = *reinterpret_cast<const double2*>(&s_st[staged(b)]);
                            if (r2_exact(a2.x - o2.x, a2.y - o2.y) < p.R2) in |= 1u << b;
                        } while (amb);
                    }
                    cnt = __popc(in);
                    if (own) {
                        int e = 0;
                        while (in && e < TL_NBR) {
                            const int b = __ffs(in) - 1;
                            in &= in - 1;
                            nb[e * TL_CAP] = (unsigned short)staged(b);
                            ++e;
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int dyi = 0; dyi < 3; ++dyi) {
                    const int r = ly + dyi - 1;
                    const int c0 = s_cs[r][lx - 1], c1 = s_cs[r][lx + 2];
                    for (int base = c0; base < c1; base += 32) {
                        unsigned acc = tile_filter(s_xy, s_st, geo, p.R2, i, me, far, base, min(32, c1 - base));
                        if (!own) {
                            cnt += __popc(acc);
                        } else {
                            while (acc) {
                                const int j = base + __ffs(acc) - 1;
                                acc &= acc - 1;
                                if (cnt < TL_NBR) nb[cnt * TL_CAP] = (unsigned short)j;
                                ++cnt;
                            }
                        }
                    }
                }
            }
            s_deg[i] = cnt;
        }
        tl::cp_async_wait_all();
        __syncthreads();
        // ---- phase B: one thread per owned agent: features, first hop, ELL head, CSR row -- one walk over the masks -----
        for (int kb0 = 0; kb0 < nB; kb0 += TL_THREADS) {
            const int kb = kb0 + tid;
            const bool valid = kb < nB;
            int i = 0, count = 0;
            if (valid) {
                i = s_listB[kb];
                count = s_deg[i];
            }
            // CSR rows: one contiguous run of edge slots per warp (all rows, or only the rows longer than the ELL head)
            const bool wants_row = valid && (geo.csr_tail_only ? count > ELLW : count > 0);
            unsigned row = 0;
            bool write_row = wants_row;
            if (__any_sync(0xffffffffu, wants_row)) {
                int inc = wants_row ? count : 0;
                const int mine = inc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                const int wtotal = __shfl_sync(0xffffffffu, inc, 31);
                unsigned rbase = 0;
                if (lane == 31) rbase = atomicAdd(&p.nnz_cursor[g], (unsigned)wtotal);
                rbase = __shfl_sync(0xffffffffu, rbase, 31);
                row = wants_row ? rbase + (unsigned)(inc - mine) : 0u;
                if (wants_row && (row + (unsigned)count > p.nnz_cap || row + (unsigned)count < row)) {   // capacity exceeded: drop the row, flag it
                    *p.overflow = 1;
                    row = 0;
                    write_row = false;
                }
            }
            if (!valid) continue;
            const unsigned lc = s_lcell[i];
            const int lx = lc & 63, ly = (lc >> 6) & 31;
            const bool far = (lc & 0x8000) != 0;
            const int a = s_id[i];
            const double4 me = s_st[i];
            float sn = 0.f;
            if (NBR > 1) sn = __ldg(&p.sinv[(size_t)slot_of(t - 1, K) * M + a]);
            double f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, f5 = 0;
            float acc[NBR > 0 ? NBR : 1][F];
#pragma unroll
            for (int b = 0; b < (NBR > 0 ? NBR : 1); ++b)
#pragma unroll
                for (int f = 0; f < F; ++f) acc[b][f] = 0.f;
            // one accepted pair: float64 feature terms + the first-hop sums from the staged source rows
            auto pair = [&](int j) {
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
                        const float* rowp = s_rows + (b * TL_CAP + j) * ROW;
                        const float4 v0 = *reinterpret_cast<const float4*>(rowp);
                        const float2 v1 = *reinterpret_cast<const float2*>(rowp + 4);
                        acc[b][0] = fmaf(v0.x, sc, acc[b][0]);
                        acc[b][1] = fmaf(v0.y, sc, acc[b][1]);
                        acc[b][2] = fmaf(v0.z, sc, acc[b][2]);
                        acc[b][3] = fmaf(v0.w, sc, acc[b][3]);
                        acc[b][4] = fmaf(v1.x, sc, acc[b][4]);
                        acc[b][5] = fmaf(v1.y, sc, acc[b][5]);
                    }
                }
            };
            const unsigned short* nb = s_nbr + kb;
            int* cols = p.cols + (size_t)g * p.nnz_cap + row;
            const int n_st = count < TL_NBR ? count : TL_NBR;
#pragma unroll 1
            for (int e = 0; e < n_st; ++e) {
                const int j = nb[e * TL_CAP];
                pair(j);
                if (write_row) cols[e] = s_id[j];
            }
            if (count > TL_NBR) {                            // rare long row: the pairs beyond the staged ones, same order
                int e = 0;
#pragma unroll 1
                for (int dyi = 0; dyi < 3; ++dyi) {
                    const int r = ly + dyi - 1;
                    const int c0 = s_cs[r][lx - 1], c1 = s_cs[r][lx + 2];
                    for (int base = c0; base < c1; base += 32) {
                        unsigned m = tile_filter(s_xy, s_st, geo, p.R2, i, s_xy[i], far, base, min(32, c1 - base));
                        while (m) {
                            const int j = base + __ffs(m) - 1;
                            m &= m - 1;
                            if (e >= TL_NBR) {
                                pair(j);
                                if (write_row) cols[e] = s_id[j];
                            }
                            ++e;
                        }
                    }
                }
            }
            if (wants_row && !write_row) count = 0;          // dropped row (capacity): the graph keeps no edge of it
            int head[ELLW];
#pragma unroll
            for (int u = 0; u < ELLW; ++u) head[u] = u < count ? s_id[nb[u * TL_CAP]] : -1;
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
