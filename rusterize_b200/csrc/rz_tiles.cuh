// rz_tiles.cuh — tile-binned polygon engine (dense output, jobs made of small polygon parts).
//
// The crossing-record pipeline (rz_kernels.cuh) materialises one 8-byte record per scanline crossing
// and sorts them: ~11 passes over 9 GB at BASELINE config 4.  When parts are small compared with the
// raster, it is far cheaper to bin PARTS to 128-column x TILE_R-row tiles (a few million (tile,part)
// records, stably sorted by tile so parts stay in burn order) and let one CTA per tile do the whole
// scanline job in shared memory, part after part:
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

// One ring edge against the tile's rows: false when it contributes no crossing there.
struct TileEdge {
    double x_top, y_top, dxdy;
    uint32_t lo, hi;  // active rows [lo, hi) inside the tile (absolute)
};
__device__ __forceinline__ bool tile_edge_setup(const KParams& P, const double* __restrict__ x,
                                                const double* __restrict__ y, uint32_t i, uint32_t r0, uint32_t r1,
                                                TileEdge& e) {
    const double y0 = px_y(P, y[i]), y1 = px_y(P, y[i + 1]);
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
    const double x0 = px_x(P, x[i]), x1 = px_x(P, x[i + 1]);
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

template <typename N, int FN, int TILE_R>
__global__ void __launch_bounds__(TILE_THREADS)
tile_fill_kernel(KParams P, TileParams T, const uint64_t* __restrict__ recs, const uint32_t* __restrict__ tile_start,
                 const PartInfo* __restrict__ info, const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                 const double* __restrict__ x, const double* __restrict__ y, const uint32_t* __restrict__ tag,
                 uint64_t bg_bits, N* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    N* tile = reinterpret_cast<N*>(smem_raw);  // [TILE_R][TILE_C]
    __shared__ uint32_t s_mask[TILE_R][4];
    __shared__ uint32_t s_cpar[TILE_R / 32];   // parity of the number of crossings per row (all columns)
    __shared__ uint32_t s_max;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const N bg = value_from_bits<N>(bg_bits);
    const uint64_t part_mask = (1ull << T.part_bits) - 1ull;

    const uint32_t t = blockIdx.x;
    const uint32_t tcol = t % T.n_tc, trow = (t / T.n_tc) % T.n_tr, band = t / (T.n_tc * T.n_tr);
    const uint32_t r0 = P.win_r0 + trow * TILE_R, r1 = min(r0 + TILE_R, P.win_r1);
    const uint32_t c0 = tcol * TILE_C, c1 = min(c0 + TILE_C, P.ncols);

    for (uint32_t i = tid; i < TILE_R * TILE_C; i += TILE_THREADS) tile[i] = bg;  // geo/raster.rs:23-28
    for (uint32_t i = tid; i < TILE_R * 4; i += TILE_THREADS) (&s_mask[0][0])[i] = 0;
    if (tid < TILE_R / 32) s_cpar[tid] = 0;
    __syncthreads();

    const uint32_t beg = tile_start[t], end = tile_start[t + 1];
    for (uint32_t rec = beg; rec < end; rec++) {
        const uint32_t part = (uint32_t)(recs[rec] & part_mask);
        const uint32_t vb = vbeg[part], ve = vend[part];
        const N v = value_from_bits<N>(info[part].value_bits);

        // ---- phase 1: crossings of the part's edges with the tile's rows -> toggle bits ------------
        for (uint32_t i = vb + tid; i + 1 < ve; i += TILE_THREADS) {
            if (tag[i] & 0x80000000u) continue;  // last vertex of its ring
            TileEdge e;
            if (!tile_edge_setup(P, x, y, i, r0, r1, e)) continue;
            for (uint32_t row = e.lo; row < e.hi; row++) {
                const uint32_t col = tile_edge_col(P, e, row);
                const uint32_t rr = row - r0;
                atomicXor(&s_cpar[rr >> 5], 1u << (rr & 31));
                if (col >= c1) continue;                 // right of the tile: no effect on its pixels
                const uint32_t rel = col <= c0 ? 0u : col - c0;
                atomicXor(&s_mask[rr][rel >> 5], 1u << (rel & 31));
            }
        }
        __syncthreads();

        // ---- rare: rows with an odd number of crossings drop their largest column (burners.rs:305) ---
        uint32_t odd = 0;
#pragma unroll
        for (int k = 0; k < TILE_R / 32; k++) odd |= s_cpar[k];
        if (odd) {
            for (int k = 0; k < TILE_R / 32; k++) {
                uint32_t bits = s_cpar[k];
                while (bits) {
                    const uint32_t rr = k * 32 + (__ffs(bits) - 1);
                    bits &= bits - 1;
                    const uint32_t row = r0 + rr;
                    if (tid == 0) s_max = 0;
                    __syncthreads();
                    for (uint32_t i = vb + tid; i + 1 < ve; i += TILE_THREADS) {
                        if (tag[i] & 0x80000000u) continue;
                        TileEdge e;
                        if (!tile_edge_setup(P, x, y, i, row, row + 1, e)) continue;
                        atomicMax(&s_max, tile_edge_col(P, e, row) + 1u);
                    }
                    __syncthreads();
                    if (tid == 0 && s_max) {
                        const uint32_t col = s_max - 1;
                        if (col < c1) {
                            const uint32_t rel = col <= c0 ? 0u : col - c0;
                            s_mask[rr][rel >> 5] ^= 1u << (rel & 31);
                        }
                    }
                    __syncthreads();
                }
            }
            if (tid < TILE_R / 32) s_cpar[tid] = 0;
            __syncthreads();
        }

        // ---- phases 2+3: 8 rows per warp step; lane = (row in group, mask word) --------------------
        for (uint32_t g = warp; g < TILE_R / 8; g += TILE_THREADS / 32) {
            const uint32_t rr = g * 8 + (lane >> 2), wd = lane & 3u;
            const uint32_t tg = s_mask[rr][wd];
            if (__ballot_sync(0xffffffffu, tg != 0) == 0) continue;  // part does not reach these rows
            s_mask[rr][wd] = 0;
            uint32_t m = tg;
            m ^= m << 1;
            m ^= m << 2;
            m ^= m << 4;
            m ^= m << 8;
            m ^= m << 16;
            const uint32_t odd_words = __ballot_sync(0xffffffffu, __popc(tg) & 1);
            if (__popc((odd_words >> (lane & ~3u)) & ((1u << wd) - 1u)) & 1) m = ~m;  // carry from the words to the left
            uint32_t nz = __ballot_sync(0xffffffffu, m != 0);
            while (nz) {
                const int src = __ffs(nz) - 1;
                nz &= nz - 1;
                const uint32_t mw = __shfl_sync(0xffffffffu, m, src);
                N* p = tile + (g * 8 + (src >> 2)) * TILE_C + (src & 3) * 32 + lane;
                const N cur = *p;
                const N nv = apply_px<N, FN>(cur, v, bg);
                *p = ((mw >> lane) & 1u) ? nv : cur;
            }
        }
        __syncthreads();
    }

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
