// rz_tiles.cuh — tile-binned polygon engine (dense output, jobs made of small polygon parts).
//
// The crossing-record pipeline (rz_kernels.cuh) materialises one 8-byte record per scanline crossing
// and sorts them: ~11 passes over 9 GB at BASELINE config 4.  When parts are small compared with the
// raster, it is far cheaper to bin PARTS to 128-column x TILE_R-row tiles (a few million (tile,part)
// records, stably sorted by tile so parts stay in burn order) and let one CTA per tile do the whole
// scanline job in shared memory, part after part:
//   phase 0  (once per call) every ring vertex is transformed to pixel space (edges.rs:94-97);
//   phase 1  threads take the part's ring edges (edges.rs:27-46, 90-110), compute the crossings with
//            the tile's rows (edges.rs:50-55) and XOR one bit per crossing into a TILE_R x 128 bit
//            toggle mask (columns left of the tile clamp to bit 0, columns right of it are dropped);
//   phase 2  prefix-XOR along each mask row = even-odd inside mask (== sorting + pairing the
//            crossings, burners.rs:302-315; an odd row drops its largest column like chunks_exact);
//   phase 3  the part's value is applied to the masked pixels with the reference's pixel-function rule
//            (pixel_functions.rs:56-123), 32 consecutive pixels per warp step.
// Parts are applied strictly in burn order, so every pixel function stays bit-exact, and every output
// byte is written to HBM once.
#pragma once

#include "rz_kernels.cuh"

namespace rz {

constexpr uint32_t TILE_C = 128;  // columns per tile = 4 mask words per row
constexpr int TILE_THREADS = 256;

struct TileParams {
    uint32_t tile_r;           // rows per tile
    uint32_t n_tc, n_tr;       // tile grid of one band in this window
    uint32_t n_tiles;          // n_bands * n_tr * n_tc
    uint32_t part_bits;
    uint32_t win_row_off, out_rows;
    uint32_t vec_ok;
};

struct TileCounters {
    unsigned long long pairs;        // sum over parts of tiles overlapped
    unsigned long long edge_visits;  // sum over parts of tiles * ring vertices
};

// pixel rows / columns a polygon part can fill, from its world extent (one pixel of margin)
__device__ __forceinline__ bool part_pixel_box(const KParams& P, double xlo, double xhi, double ylo, double yhi,
                                               uint32_t& r_lo, uint32_t& r_hi, uint32_t& c_lo, uint32_t& c_hi) {
    // rows whose centre can lie inside: [ceil(y_top - 0.5), ceil(y_bot - 0.5))
    uint32_t a = sat_u32(ceil(__dsub_rn(px_y(P, yhi), 0.5)), P.nrows);
    uint32_t b = sat_u32(ceil(__dsub_rn(px_y(P, ylo), 0.5)), P.nrows);
    r_lo = max(a > 0 ? a - 1 : 0u, P.win_r0);
    r_hi = min(b < P.nrows ? b + 1 : P.nrows, P.win_r1);
    uint32_t cl = sat_u32(floor(__dadd_rn(px_x(P, xlo), 0.5)), P.ncols);
    uint32_t ch = sat_u32(floor(__dadd_rn(px_x(P, xhi), 0.5)), P.ncols);
    c_lo = cl > 0 ? cl - 1 : 0u;
    c_hi = ch < P.ncols ? ch + 1 : P.ncols;
    return r_hi > r_lo && c_hi > c_lo && c_lo < P.ncols;
}

// mode 0: cnt[p] = tiles overlapped by part p (+ totals); mode 1: write its records at off[p]
__global__ void tile_bin_kernel(KParams P, TileParams T, const PartInfo* __restrict__ info,
                                const double* __restrict__ xlo, const double* __restrict__ xhi,
                                const double* __restrict__ ylo, const double* __restrict__ yhi,
                                const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                                uint32_t* __restrict__ cnt, const unsigned long long* __restrict__ off,
                                uint64_t* __restrict__ recs, TileCounters* __restrict__ tc, int mode) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 0, r_lo = 0, r_hi = 0, c_lo = 0, c_hi = 0;
    int32_t band = -1;
    if (p < P.n_parts) {
        band = info[p].band;
        if (band >= 0 && part_pixel_box(P, xlo[p], xhi[p], ylo[p], yhi[p], r_lo, r_hi, c_lo, c_hi)) {
            const uint32_t tr0 = (r_lo - P.win_r0) / T.tile_r, tr1 = (r_hi - 1 - P.win_r0) / T.tile_r;
            const uint32_t tc0 = c_lo / TILE_C, tc1 = (c_hi - 1) / TILE_C;
            n = (tr1 - tr0 + 1) * (tc1 - tc0 + 1);
            if (mode == 1) {
                unsigned long long o = off[p];
                for (uint32_t tr = tr0; tr <= tr1; tr++)
                    for (uint32_t tcol = tc0; tcol <= tc1; tcol++) {
                        const uint64_t tile = ((uint64_t)band * T.n_tr + tr) * T.n_tc + tcol;
                        recs[o++] = (tile << T.part_bits) | p;
                    }
            }
        }
    }
    if (mode == 0) {
        if (p < P.n_parts) cnt[p] = n;
        unsigned long long pairs = n, visits = (unsigned long long)n * (p < P.n_parts ? vend[p] - vbeg[p] : 0u);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pairs += __shfl_down_sync(0xffffffffu, pairs, o);
            visits += __shfl_down_sync(0xffffffffu, visits, o);
        }
        if (lane_id() == 0 && pairs) {
            atomicAdd(&tc->pairs, pairs);
            atomicAdd(&tc->edge_visits, visits);
        }
    }
}

struct InU32 {
    const uint32_t* v;
    __device__ unsigned long long operator()(uint32_t i) const { return v[i]; }
};

// World -> pixel transform of every ring vertex, once per call (edges.rs:94-97): the tile kernel visits
// a part's edges once per overlapped tile and would otherwise repeat these four divides each time.
__global__ void vertex_transform_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                        uint32_t n, double* __restrict__ px, double* __restrict__ py) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    px[i] = px_x(P, x[i]);
    py[i] = px_y(P, y[i]);
}

// One ring edge (pixel-space vertices) against the tile's rows: false when it contributes no crossing there.
struct TileEdge {
    double x_top, y_top, dxdy;
    uint32_t lo, hi;  // active rows [lo, hi) inside the tile (absolute)
};
__device__ __forceinline__ bool tile_edge_setup(const KParams& P, const double* __restrict__ px,
                                                const double* __restrict__ py, uint32_t i, uint32_t r0, uint32_t r1,
                                                TileEdge& e) {
    const double y0 = py[i], y1 = py[i + 1];
    if (!(fabs(__dsub_rn(y0, y1)) >= DBL_EPSILON)) return false;  // edges.rs:100
    const double min_y = fmin(y0, y1), max_y = fmax(y0, y1);
    if (!(min_y < P.nrows_f && max_y >= 0.0)) return false;       // edges.rs:105
    const bool down = y0 < y1;                                    // edges.rs:29
    const double y_top = down ? y0 : y1, y_bot = down ? y1 : y0;
    const uint32_t ystart = sat_u32(ceil(__dsub_rn(y_top, 0.5)), P.nrows);
    const uint32_t yend = sat_u32(ceil(__dsub_rn(y_bot, 0.5)), P.nrows);
    e.lo = max(ystart, r0);
    e.hi = min(yend, r1);
    if (e.hi <= e.lo) return false;
    const double x0 = px[i], x1 = px[i + 1];
    const double x_bot = down ? x1 : x0;
    e.x_top = down ? x0 : x1;
    e.y_top = y_top;
    e.dxdy = __ddiv_rn(__dsub_rn(x_bot, e.x_top), __dsub_rn(y_bot, y_top));
    return true;
}
__device__ __forceinline__ uint32_t tile_edge_col(const KParams& P, const TileEdge& e, uint32_t row) {
    const double cy = __dadd_rn((double)row, 0.5);
    const double xi = __dadd_rn(e.x_top, __dmul_rn(__dsub_rn(cy, e.y_top), e.dxdy));  // edges.rs:50-55
    return sat_u32(floor(__dadd_rn(xi, 0.5)), P.ncols);                               // burners.rs:310-311
}

// A row can only have an odd number of crossings when a non-horizontal edge was skipped for being
// shorter than f64::EPSILON in y (edges.rs:100) while still straddling a pixel centre.  Pixel-centre
// ordinates k+0.5 with k >= 1 are spaced >= EPSILON apart, so that can only happen on raster row 0
// (centre 0.5): only that row's crossing count is tracked.
//
// Warp-specialised CTA, no block barrier in the main loop:
//   producer warps          the 32-edge batches of all parts form one sequence; batch b goes to producer b%P, which bins its edges into toggle
//                          mask slot k%8 (phase 1) and then publishes ready[slot] = k+1;
//   consumer warps          each owns a fixed subset of the tile's 8-row groups and applies parts strictly in
//                          order (phases 2+3) as their masks become ready, clearing the mask words it
//                          read; the last consumer of a part frees the slot (consumed++).
// Producers run up to 8 parts ahead of the consumers, so edge setup (f64 divides, global loads) overlaps
// the pixel work instead of alternating with it across barriers.
constexpr int TILE_SLOTS = 4;
constexpr int TILE_PRODUCERS = 6;
constexpr int TILE_CONSUMERS = 2;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

template <typename N, int FN, int TILE_R>
__global__ void __launch_bounds__(TILE_THREADS, 5)
tile_fill_kernel(KParams P, TileParams T, const uint64_t* __restrict__ recs, const uint32_t* __restrict__ tile_start,
                 const PartInfo* __restrict__ info, const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                 const double* __restrict__ x, const double* __restrict__ y, const uint32_t* __restrict__ tag,
                 uint64_t bg_bits, N* __restrict__ out) {
    static_assert(TILE_THREADS == 32 * (TILE_PRODUCERS + TILE_CONSUMERS), "role split");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    N* tile = reinterpret_cast<N*>(smem_raw);  // [TILE_R][TILE_C]
    __shared__ uint32_t s_mask[TILE_SLOTS][TILE_R][4];
    __shared__ unsigned long long s_val[TILE_SLOTS];
    __shared__ uint32_t s_ready[TILE_SLOTS], s_done[TILE_SLOTS], s_cnt[TILE_SLOTS], s_par0[TILE_SLOTS], s_consumed;
    __shared__ double s_xt[TILE_PRODUCERS][32], s_yt[TILE_PRODUCERS][32], s_dx[TILE_PRODUCERS][32];
    __shared__ uint32_t s_pre[TILE_PRODUCERS][32], s_lo[TILE_PRODUCERS][32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const N bg = value_from_bits<N>(bg_bits);
    const uint64_t part_mask = (1ull << T.part_bits) - 1ull;

    const uint32_t t = blockIdx.x;
    const uint32_t tcol = t % T.n_tc, trow = (t / T.n_tc) % T.n_tr, band = t / (T.n_tc * T.n_tr);
    const uint32_t r0 = P.win_r0 + trow * TILE_R, r1 = min(r0 + TILE_R, P.win_r1);
    const uint32_t c0 = tcol * TILE_C, c1 = min(c0 + TILE_C, P.ncols);
    const uint32_t beg = tile_start[t], n_parts_here = tile_start[t + 1] - beg;

    for (uint32_t i = tid; i < TILE_R * TILE_C; i += TILE_THREADS) tile[i] = bg;  // geo/raster.rs:23-28
    if (n_parts_here) {
        for (uint32_t i = tid; i < TILE_SLOTS * TILE_R * 4; i += TILE_THREADS) (&s_mask[0][0][0])[i] = 0;
        if (tid < TILE_SLOTS) {
            s_ready[tid] = 0;
            s_done[tid] = 0;
            s_cnt[tid] = 0;
            s_par0[tid] = 0;
        }
        if (tid == 0) s_consumed = 0;
    }
    __syncthreads();

    if (n_parts_here && warp < TILE_PRODUCERS) {
        // ================= producers: phase 1, one 32-edge batch at a time =================
        // Batches of all parts form one sequence; batch b goes to producer b % TILE_PRODUCERS, so the
        // edges of one part are binned by several warps at once and parts overlap in a pipeline.
        uint32_t b_first = 0;  // sequence number of the part's first batch
        for (uint32_t k = 0; k < n_parts_here; k++) {
            const uint32_t slot = k % TILE_SLOTS;
            const uint32_t part = (uint32_t)(recs[beg + k] & part_mask);
            const uint32_t vb = vbeg[part], ve = vend[part];
            const uint32_t n_edges = ve > vb ? ve - vb - 1 : 0u;
            const uint32_t nb = max(1u, (n_edges + 31) / 32);
            uint32_t(*mask)[4] = s_mask[slot];
            // first batch of this part that belongs to this warp
            uint32_t j = (warp + TILE_PRODUCERS - b_first % TILE_PRODUCERS) % TILE_PRODUCERS;
            b_first += nb;
            if (j >= nb) continue;
            while (k >= ld_volatile_u32(&s_consumed) + TILE_SLOTS) __nanosleep(64);  // slot still in use
            __syncwarp();
            for (; j < nb; j += TILE_PRODUCERS) {
                const uint32_t i = vb + j * 32 + lane;
                TileEdge e;
                uint32_t cnt = 0;
                if (i + 1 < ve && !(tag[i] & 0x80000000u) && tile_edge_setup(P, x, y, i, r0, r1, e)) cnt = e.hi - e.lo;
                uint32_t inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += up;
                }
                const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
                uint32_t par0 = 0;  // this lane's share of the row-0 crossing count parity
                if (wtot) {
                    s_pre[warp][lane] = inc - cnt;
                    if (cnt) {
                        s_xt[warp][lane] = e.x_top;
                        s_yt[warp][lane] = e.y_top;
                        s_dx[warp][lane] = e.dxdy;
                        s_lo[warp][lane] = e.lo;
                    }
                    __syncwarp();
                    for (uint32_t q = lane; q < wtot; q += 32) {
                        uint32_t lo = 0, hi = 32;  // last edge whose first crossing is <= q
#pragma unroll
                        for (int it = 0; it < 5; it++) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (s_pre[warp][mid] <= q) lo = mid;
                            else hi = mid;
                        }
                        TileEdge eb;
                        eb.x_top = s_xt[warp][lo];
                        eb.y_top = s_yt[warp][lo];
                        eb.dxdy = s_dx[warp][lo];
                        const uint32_t row = s_lo[warp][lo] + (q - s_pre[warp][lo]);
                        const uint32_t col = tile_edge_col(P, eb, row);
                        par0 ^= (row == 0);
                        if (col < c1) {  // right of the tile: no effect on its pixels
                            const uint32_t rel = col <= c0 ? 0u : col - c0;
                            atomicXor(&mask[row - r0][rel >> 5], 1u << (rel & 31));
                        }
                    }
                }
                const uint32_t odd = __popc(__ballot_sync(0xffffffffu, par0 & 1u)) & 1u;
                __threadfence_block();
                uint32_t fin = 0;
                if (lane == 0) {
                    if (odd) atomicXor(&s_par0[slot], 1u);
                    fin = atomicAdd(&s_cnt[slot], 1u) + 1u;
                }
                fin = __shfl_sync(0xffffffffu, fin, 0);
                if (fin == nb) {  // this warp finished the part's last outstanding batch
                    __threadfence_block();
                    // rare: an odd row 0 drops its largest column (chunks_exact(2), burners.rs:305)
                    if (r0 == 0 && ld_volatile_u32(&s_par0[slot])) {
                        uint32_t mx = 0;
                        for (uint32_t ii = vb + lane; ii + 1 < ve; ii += 32) {
                            if (tag[ii] & 0x80000000u) continue;
                            TileEdge e0;
                            if (tile_edge_setup(P, x, y, ii, 0, 1, e0)) mx = max(mx, tile_edge_col(P, e0, 0) + 1u);
                        }
                        mx = __reduce_max_sync(0xffffffffu, mx);
                        if (lane == 0 && mx && mx - 1 < c1) {
                            const uint32_t rel = mx - 1 <= c0 ? 0u : mx - 1 - c0;
                            atomicXor(&mask[0][rel >> 5], 1u << (rel & 31));
                        }
                    }
                    if (lane == 0) {
                        s_par0[slot] = 0;
                        s_cnt[slot] = 0;
                        s_val[slot] = info[part].value_bits;
                        __threadfence_block();
                        st_volatile_u32(&s_ready[slot], k + 1);
                    }
                }
                __syncwarp();
            }
        }
    } else if (n_parts_here) {
        // ================= consumer: phases 2+3 on its own rows =================
        const uint32_t cw = warp - TILE_PRODUCERS;
        for (uint32_t k = 0; k < n_parts_here; k++) {
            const uint32_t slot = k % TILE_SLOTS;
            while (ld_volatile_u32(&s_ready[slot]) != k + 1) __nanosleep(32);
            __syncwarp();
            __threadfence_block();
            const N v = value_from_bits<N>(*reinterpret_cast<volatile unsigned long long*>(&s_val[slot]));
            uint32_t(*mask)[4] = s_mask[slot];
            for (uint32_t g = cw; g < TILE_R / 8; g += TILE_CONSUMERS) {  // this consumer's 8-row groups
                const uint32_t rr = g * 8 + (lane >> 2), wd = lane & 3u;
                const uint32_t tg = *reinterpret_cast<volatile uint32_t*>(&mask[rr][wd]);
                if (__ballot_sync(0xffffffffu, tg != 0) == 0) continue;  // part does not reach these rows
                mask[rr][wd] = 0;
                uint32_t m = tg;
                m ^= m << 1;
                m ^= m << 2;
                m ^= m << 4;
                m ^= m << 8;
                m ^= m << 16;
                const uint32_t odd_words = __ballot_sync(0xffffffffu, __popc(tg) & 1);
                if (__popc((odd_words >> (lane & ~3u)) & ((1u << wd) - 1u)) & 1) m = ~m;  // carry from the left words
                uint32_t nz = __ballot_sync(0xffffffffu, m != 0);
                N* base = tile + g * (8 * TILE_C) + lane;  // word `src` of the group starts at base + src*32
                while (nz) {  // two mask words per step: their shared-memory round trips overlap
                    const int src0 = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const int src1 = nz ? __ffs(nz) - 1 : src0;
                    const bool two = nz != 0;
                    nz &= nz - 1;
                    const uint32_t mw0 = __shfl_sync(0xffffffffu, m, src0);
                    const uint32_t mw1 = __shfl_sync(0xffffffffu, m, src1);
                    N* p0 = base + src0 * 32;
                    N* p1 = base + src1 * 32;
                    const N cur0 = *p0;
                    const N cur1 = *p1;
                    const N nv0 = apply_px<N, FN>(cur0, v, bg);
                    const N nv1 = apply_px<N, FN>(cur1, v, bg);
                    *p0 = ((mw0 >> lane) & 1u) ? nv0 : cur0;
                    if (two) *p1 = ((mw1 >> lane) & 1u) ? nv1 : cur1;
                }
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0 && atomicAdd(&s_done[slot], 1u) == TILE_CONSUMERS - 1) {
                s_done[slot] = 0;
                __threadfence_block();
                atomicAdd(&s_consumed, 1u);  // the slot may be reused by part k + TILE_SLOTS
            }
        }
    }
    __syncthreads();

    // ---- flush: every output byte is written exactly once ---------------------------------------------
    const uint32_t cols = c1 - c0;
    for (uint32_t rr = warp; rr < r1 - r0; rr += TILE_THREADS / 32) {
        N* dst = out + ((size_t)band * T.out_rows + T.win_row_off + (r0 - P.win_r0) + rr) * P.ncols + c0;
        const N* src = tile + rr * TILE_C;
        if (T.vec_ok && cols == TILE_C) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            for (uint32_t i = lane; i < TILE_C * sizeof(N) / 16; i += 32) __stcs(d4 + i, s4[i]);
        } else {
            for (uint32_t i = lane; i < cols; i += 32) dst[i] = src[i];
        }
    }
}

}  // namespace rz
