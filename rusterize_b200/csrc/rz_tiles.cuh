// rz_tiles.cuh — tile-binned polygon engine (dense output, jobs made of small polygon parts).
//
// The crossing-record pipeline (rz_kernels.cuh) materialises one 8-byte record per scanline crossing
// and sorts them: ~11 passes over 9 GB at BASELINE config 4.  When parts are small compared with the
// raster it is far cheaper never to materialise crossings.  Parts are binned to tiles of 128 columns x
// TILE_R rows and the scanline job is split into two embarrassingly parallel kernels:
//
//   vertex_transform   every ring vertex to pixel space, once per call (edges.rs:94-97).
//   tile_bin (x2)      per part: the tile rows / tile columns its bounding box overlaps; emits the list
//                      of (part, tile-row) pairs and the (tile, part) records (stably sorted by tile, so
//                      every tile sees its parts in burn order).
//   tile_mask          one WARP per (part, tile-row): takes the part's ring edges 32 at a time
//                      (edges.rs:27-46, 90-110), computes their crossings with the tile-row's rows
//                      (edges.rs:50-55, warp-flattened so all lanes stay busy) and XORs one bit per
//                      crossing into a shared-memory toggle mask spanning the part's tile columns; a
//                      prefix-XOR along each row turns it into the even-odd INSIDE mask (== sorting and
//                      pairing the crossings, burners.rs:302-315), written to global memory as one
//                      TILE_R x 128-bit block per (part, tile).
//   tile_apply         one CTA per tile, tile pixels in shared memory (initialised to the background,
//                      flushed once).  Each warp owns 8 rows and, with NO synchronisation with the other
//                      warps, walks the tile's parts in burn order: one coalesced 128-byte load fetches
//                      its 8 rows x 4 words of the part's inside mask, and the part's value is applied to
//                      the masked pixels with the reference's pixel-function rule
//                      (pixel_functions.rs:56-123), 32 consecutive pixels per step.
//
// Parts are applied strictly in burn order per pixel, so every pixel function stays bit-exact, and every
// output byte is written to HBM once.
#pragma once

#include "rz_kernels.cuh"

namespace rz {

constexpr uint32_t TILE_C = 128;        // columns per tile = 4 mask words per row
constexpr uint32_t MASK_MAX_WORDS = 12; // tile_mask: widest toggle-mask chunk kept in shared memory (384 columns)
constexpr int MASK_WARPS = 4;

struct TileParams {
    uint32_t tile_r;           // rows per tile (64, or 32 for 8-byte dtypes)
    uint32_t n_tc, n_tr;       // tile grid of one band in this window
    uint32_t n_tiles;          // n_bands * n_tr * n_tc
    uint32_t part_bits;
    uint32_t win_row_off, out_rows;
    uint32_t vec_ok;
};

struct TileCounters {
    unsigned long long pairs;        // (part, tile) pairs
    unsigned long long row_pairs;    // (part, tile-row) pairs
    unsigned long long edge_visits;  // sum over parts of tile-rows x column chunks x ring vertices
};

// where a part's inside-mask blocks live: block(tr, tc) = first_block + (tr - tr0) * ntc + (tc - tc0)
struct PartTile {
    unsigned long long first_block;
    uint32_t tr0, tc0;
    uint32_t ntr, ntc;
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

// mode 0: per part the number of (part,tile) pairs and of (part,tile-row) pairs (+ totals);
// mode 1: with the scanned offsets, fill PartTile and emit the row-pair list and the tile records
// [tile | block], block = index of the (part,tile) pair's inside-mask block.
__global__ void tile_bin_kernel(KParams P, TileParams T, const PartInfo* __restrict__ info,
                                const double* __restrict__ xlo, const double* __restrict__ xhi,
                                const double* __restrict__ ylo, const double* __restrict__ yhi,
                                const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                                uint32_t* __restrict__ cnt_tiles, uint32_t* __restrict__ cnt_rows,
                                const unsigned long long* __restrict__ off_tiles,
                                const unsigned long long* __restrict__ off_rows, PartTile* __restrict__ pt,
                                uint64_t* __restrict__ row_pairs, uint64_t* __restrict__ recs,
                                unsigned long long* __restrict__ block_value, uint32_t block_bits,
                                TileCounters* __restrict__ tc, int mode) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ntr = 0, ntc = 0, r_lo = 0, r_hi = 0, c_lo = 0, c_hi = 0;
    if (p < P.n_parts) {
        const int32_t band = info[p].band;
        if (band >= 0 && vend[p] > vbeg[p] + 1 &&
            part_pixel_box(P, xlo[p], xhi[p], ylo[p], yhi[p], r_lo, r_hi, c_lo, c_hi)) {
            const uint32_t tr0 = (r_lo - P.win_r0) / T.tile_r, tr1 = (r_hi - 1 - P.win_r0) / T.tile_r;
            const uint32_t tc0 = c_lo / TILE_C, tc1 = (c_hi - 1) / TILE_C;
            ntr = tr1 - tr0 + 1;
            ntc = tc1 - tc0 + 1;
            if (mode == 1) {
                PartTile q;
                q.first_block = off_tiles[p];
                q.tr0 = tr0;
                q.tc0 = tc0;
                q.ntr = ntr;
                q.ntc = ntc;
                pt[p] = q;
                unsigned long long o = off_tiles[p], orow = off_rows[p];
                for (uint32_t tr = tr0; tr <= tr1; tr++) {
                    row_pairs[orow++] = ((uint64_t)tr << 32) | p;
                    for (uint32_t tcol = tc0; tcol <= tc1; tcol++) {
                        const uint64_t tile = ((uint64_t)band * T.n_tr + tr) * T.n_tc + tcol;
                        block_value[o] = info[p].value_bits;  // tile_apply reads the value by block, not by part
                        recs[o] = (tile << block_bits) | o;   // block index ascends with the part id: burn order
                        o++;
                    }
                }
            }
        }
    }
    if (mode == 0) {
        if (p < P.n_parts) {
            cnt_tiles[p] = ntr * ntc;
            cnt_rows[p] = ntr;
        }
        const uint32_t chunks = (ntc * 4 + MASK_MAX_WORDS - 1) / MASK_MAX_WORDS;
        unsigned long long pairs = (unsigned long long)ntr * ntc, rows = ntr,
                           visits = (unsigned long long)ntr * chunks * (p < P.n_parts ? vend[p] - vbeg[p] : 0u);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pairs += __shfl_down_sync(0xffffffffu, pairs, o);
            rows += __shfl_down_sync(0xffffffffu, rows, o);
            visits += __shfl_down_sync(0xffffffffu, visits, o);
        }
        if (lane_id() == 0 && pairs) {
            atomicAdd(&tc->pairs, pairs);
            atomicAdd(&tc->row_pairs, rows);
            atomicAdd(&tc->edge_visits, visits);
        }
    }
}

struct InU32 {
    const uint32_t* v;
    __device__ unsigned long long operator()(uint32_t i) const { return v[i]; }
};

// World -> pixel transform of every ring vertex, once per call (edges.rs:94-97): tile_mask visits a part's
// edges once per tile-row and would otherwise repeat these four divides each time.
__global__ void vertex_transform_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                        uint32_t n, double* __restrict__ px, double* __restrict__ py) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    px[i] = px_x(P, x[i]);
    py[i] = px_y(P, y[i]);
}

// One ring edge (pixel-space vertices) against rows [r0, r1): false when it has no crossing there.
struct TileEdge {
    double x_top, y_top, dxdy;
    uint32_t lo, hi;  // active rows [lo, hi) (absolute)
};
// y part of the edge test on already loaded pixel-space ordinates; *down = edge runs top to bottom
__device__ __forceinline__ bool tile_edge_rows(const KParams& P, double y0, double y1, uint32_t r0, uint32_t r1,
                                               TileEdge& e, bool* down, double* y_bot) {
    if (!(fabs(__dsub_rn(y0, y1)) >= DBL_EPSILON)) return false;  // edges.rs:100
    const double min_y = fmin(y0, y1), max_y = fmax(y0, y1);
    if (!(min_y < P.nrows_f && max_y >= 0.0)) return false;       // edges.rs:105
    *down = y0 < y1;                                              // edges.rs:29
    e.y_top = *down ? y0 : y1;
    *y_bot = *down ? y1 : y0;
    const uint32_t ystart = sat_u32(ceil(__dsub_rn(e.y_top, 0.5)), P.nrows);
    const uint32_t yend = sat_u32(ceil(__dsub_rn(*y_bot, 0.5)), P.nrows);
    e.lo = max(ystart, r0);
    e.hi = min(yend, r1);
    return e.hi > e.lo;
}
__device__ __forceinline__ void tile_edge_slope(double x0, double x1, bool down, double y_bot, TileEdge& e) {
    const double x_bot = down ? x1 : x0;
    e.x_top = down ? x0 : x1;
    e.dxdy = __ddiv_rn(__dsub_rn(x_bot, e.x_top), __dsub_rn(y_bot, e.y_top));  // edges.rs:36
}
__device__ __forceinline__ bool tile_edge_setup(const KParams& P, const double* __restrict__ px,
                                                const double* __restrict__ py, uint32_t i, uint32_t r0, uint32_t r1,
                                                TileEdge& e) {
    bool down;
    double y_bot;
    if (!tile_edge_rows(P, py[i], py[i + 1], r0, r1, e, &down, &y_bot)) return false;
    tile_edge_slope(px[i], px[i + 1], down, y_bot, e);
    return true;
}
__device__ __forceinline__ uint32_t tile_edge_col(const KParams& P, const TileEdge& e, uint32_t row) {
    const double cy = __dadd_rn((double)row, 0.5);
    const double xi = __dadd_rn(e.x_top, __dmul_rn(__dsub_rn(cy, e.y_top), e.dxdy));  // edges.rs:50-55
    return sat_u32(floor(__dadd_rn(xi, 0.5)), P.ncols);                               // burners.rs:310-311
}

// ---------------------------------------------------------------------------------------------
// tile_mask: one warp per (part, tile-row)
// ---------------------------------------------------------------------------------------------
// A row can only have an odd number of crossings when a non-horizontal edge was skipped for being
// shorter than f64::EPSILON in y (edges.rs:100) while still straddling a pixel centre.  Pixel-centre
// ordinates k+0.5 with k >= 1 are spaced >= EPSILON apart, so that can only happen on raster row 0
// (centre 0.5): only that row's crossing count is tracked, and an odd row 0 drops its largest column like
// chunks_exact(2) drops the unpaired tail (burners.rs:305).
template <int TILE_R>
__global__ void __launch_bounds__(MASK_WARPS * 32, 10)
tile_mask_kernel(KParams P, TileParams T, const uint64_t* __restrict__ row_pairs, uint32_t n_row_pairs,
                 const PartTile* __restrict__ pt, const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                 const double* __restrict__ px, const double* __restrict__ py, const uint32_t* __restrict__ tag,
                 uint32_t* __restrict__ masks) {
    __shared__ uint32_t s_mask[MASK_WARPS][TILE_R][MASK_MAX_WORDS];
    __shared__ double s_xt[MASK_WARPS][32], s_yt[MASK_WARPS][32], s_dx[MASK_WARPS][32];
    __shared__ uint32_t s_pre[MASK_WARPS][32], s_lo[MASK_WARPS][32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t pair = blockIdx.x * MASK_WARPS + warp;
    if (pair >= n_row_pairs) return;
    const uint64_t rp = row_pairs[pair];
    const uint32_t part = (uint32_t)rp, tr = (uint32_t)(rp >> 32);
    const PartTile q = pt[part];
    const uint32_t vb = vbeg[part], ve = vend[part];
    const uint32_t r0 = P.win_r0 + tr * TILE_R, r1 = min(r0 + TILE_R, P.win_r1);
    uint32_t(*mask)[MASK_MAX_WORDS] = s_mask[warp];
    const uint32_t words_total = q.ntc * 4;

    for (uint32_t w0 = 0; w0 < words_total; w0 += MASK_MAX_WORDS) {  // one pass per 384-column chunk
        const uint32_t nw = min(MASK_MAX_WORDS, words_total - w0);
        const uint32_t c0 = q.tc0 * TILE_C + w0 * 32;                 // first pixel column of the chunk
        const uint32_t c1 = min(c0 + nw * 32, P.ncols);
        for (uint32_t i = lane; i < TILE_R * MASK_MAX_WORDS; i += 32) (&mask[0][0])[i] = 0;
        __syncwarp();
        uint32_t par0 = 0;  // this lane's share of the row-0 crossing count parity
        // software pipeline: the next batch's tag / ordinates are in flight while this one is processed
        uint32_t tg_n = 0x80000000u;
        double y0_n = 0.0, y1_n = 0.0;
        if (vb + lane + 1 < ve) {
            tg_n = tag[vb + lane];
            y0_n = py[vb + lane];
            y1_n = py[vb + lane + 1];
        }
        for (uint32_t i0 = vb; i0 + 1 < ve; i0 += 32) {
            const uint32_t i = i0 + lane;
            const uint32_t tg = tg_n;
            const double y0 = y0_n, y1 = y1_n;
            tg_n = 0x80000000u;
            if (i + 33 < ve) {
                tg_n = tag[i + 32];
                y0_n = py[i + 32];
                y1_n = py[i + 33];
            }
            TileEdge e;
            uint32_t cnt = 0;
            bool down;
            double y_bot;
            if (!(tg & 0x80000000u) && tile_edge_rows(P, y0, y1, r0, r1, e, &down, &y_bot)) {
                tile_edge_slope(px[i], px[i + 1], down, y_bot, e);
                cnt = e.hi - e.lo;
            }
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += up;
            }
            const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
            if (wtot == 0) continue;
            s_pre[warp][lane] = inc - cnt;
            if (cnt) {
                s_xt[warp][lane] = e.x_top;
                s_yt[warp][lane] = e.y_top;
                s_dx[warp][lane] = e.dxdy;
                s_lo[warp][lane] = e.lo;
            }
            __syncwarp();
            for (uint32_t k = lane; k < wtot; k += 32) {  // all lanes share the 32 edges' crossings evenly
                uint32_t lo = 0, hi = 32;                 // last edge whose first crossing is <= k
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_pre[warp][mid] <= k) lo = mid;
                    else hi = mid;
                }
                TileEdge b;
                b.x_top = s_xt[warp][lo];
                b.y_top = s_yt[warp][lo];
                b.dxdy = s_dx[warp][lo];
                const uint32_t row = s_lo[warp][lo] + (k - s_pre[warp][lo]);
                const uint32_t col = tile_edge_col(P, b, row);
                par0 ^= (row == 0);
                if (col < c1) {  // a crossing right of the chunk has no effect on its pixels
                    const uint32_t rel = col <= c0 ? 0u : col - c0;  // left of it: parity carry-in at bit 0
                    atomicXor(&mask[row - r0][rel >> 5], 1u << (rel & 31));
                }
            }
            __syncwarp();
        }
        if (r0 == 0 && (__popc(__ballot_sync(0xffffffffu, par0 & 1u)) & 1)) {  // rare: odd row 0
            uint32_t mx = 0;
            for (uint32_t i = vb + lane; i + 1 < ve; i += 32) {
                if (tag[i] & 0x80000000u) continue;
                TileEdge e;
                if (tile_edge_setup(P, px, py, i, 0, 1, e)) mx = max(mx, tile_edge_col(P, e, 0) + 1u);
            }
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0 && mx && mx - 1 < c1) {
                const uint32_t rel = mx - 1 <= c0 ? 0u : mx - 1 - c0;
                mask[0][rel >> 5] ^= 1u << (rel & 31);
            }
        }
        __syncwarp();
        // toggle mask -> inside mask, row by row (one lane per row), in place
        for (uint32_t rr = lane; rr < TILE_R; rr += 32) {
            uint32_t carry = 0;
            for (uint32_t wd = 0; wd < nw; wd++) {
                uint32_t m = mask[rr][wd];
                m ^= m << 1;
                m ^= m << 2;
                m ^= m << 4;
                m ^= m << 8;
                m ^= m << 16;
                m ^= carry;
                carry = (m >> 31) ? 0xffffffffu : 0u;
                mask[rr][wd] = m;
            }
        }
        __syncwarp();
        // one TILE_R x 4-word block per (part, tile): consecutive lanes write consecutive words
        const unsigned long long row_base = q.first_block + (unsigned long long)(tr - q.tr0) * q.ntc;
        for (uint32_t tcl = 0; tcl * 4 < nw; tcl++) {
            uint32_t* dst = masks + (row_base + w0 / 4 + tcl) * (TILE_R * 4);
            for (uint32_t i = lane; i < TILE_R * 4; i += 32) dst[i] = mask[i >> 2][tcl * 4 + (i & 3)];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// tile_apply: one CTA per tile, each warp owns 8 rows, no synchronisation between warps
// ---------------------------------------------------------------------------------------------
template <typename N, int FN, bool BGNAN>
__device__ __forceinline__ void apply_group_mask(N* __restrict__ base, uint32_t m, uint32_t lane, N v, N bg) {
    uint32_t nz = __ballot_sync(0xffffffffu, m != 0);
    while (nz) {  // two mask words per step: their shared-memory round trips overlap
        const int src0 = __ffs(nz) - 1;
        nz &= nz - 1;
        const int src1 = nz ? __ffs(nz) - 1 : src0;
        const bool two = nz != 0;
        nz &= nz - 1;
        const uint32_t mw0 = __shfl_sync(0xffffffffu, m, src0);
        const uint32_t mw1 = __shfl_sync(0xffffffffu, m, src1);
        N* p0 = base + src0 * 32;  // word `src` of the group covers pixels src*32 .. src*32+31 of the 8 rows
        N* p1 = base + src1 * 32;
        const N cur0 = *p0;
        const N cur1 = *p1;
        const N nv0 = apply_px<N, FN, BGNAN>(cur0, v, bg);
        const N nv1 = apply_px<N, FN, BGNAN>(cur1, v, bg);
        *p0 = ((mw0 >> lane) & 1u) ? nv0 : cur0;
        if (two) *p1 = ((mw1 >> lane) & 1u) ? nv1 : cur1;
    }
}

template <typename N, int FN, int TILE_R, bool BGNAN>
__global__ void __launch_bounds__(TILE_R * 4)
tile_apply_kernel(KParams P, TileParams T, const uint64_t* __restrict__ recs, const uint32_t* __restrict__ tile_start,
                  const unsigned long long* __restrict__ block_value, uint32_t block_bits,
                  const uint32_t* __restrict__ masks, uint64_t bg_bits, N* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    N* rows8 = reinterpret_cast<N*>(smem_raw) + (size_t)warp * 8 * TILE_C;  // this warp's 8 rows x 128 columns
    const N bg = value_from_bits<N>(bg_bits);
    const uint64_t block_mask = (1ull << block_bits) - 1ull;

    const uint32_t t = blockIdx.x;
    const uint32_t tcol = t % T.n_tc, trow = (t / T.n_tc) % T.n_tr, band = t / (T.n_tc * T.n_tr);
    const uint32_t r0 = P.win_r0 + trow * TILE_R + warp * 8;
    if (r0 >= P.win_r1) return;
    const uint32_t r1 = min(r0 + 8, P.win_r1);
    const uint32_t c0 = tcol * TILE_C, c1 = min(c0 + TILE_C, P.ncols);

    for (uint32_t i = lane; i < 8 * TILE_C; i += 32) rows8[i] = bg;  // geo/raster.rs:23-28
    __syncwarp();

    const uint32_t beg = tile_start[t], end = tile_start[t + 1];
    N* base = rows8 + lane;
    const uint32_t* my_masks = masks + warp * 32 + lane;  // lane = (row in this warp's group, mask word)
    for (uint32_t chunk = beg; chunk < end; chunk += 32) {
        // 32 records with one coalesced load; their mask words are fetched four parts ahead of the apply
        const uint32_t n = min(32u, end - chunk);
        const unsigned long long my_blk = lane < n ? (recs[chunk + lane] & block_mask) : 0ull;
        for (uint32_t j0 = 0; j0 < n; j0 += 4) {
            uint32_t m[4];
            unsigned long long vbits[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const unsigned long long blk = __shfl_sync(0xffffffffu, my_blk, (j0 + u) & 31);
                const bool live = j0 + u < n;
                m[u] = live ? my_masks[blk * (TILE_R * 4)] : 0u;  // one coalesced 128-byte load
                vbits[u] = live ? block_value[blk] : 0ull;        // broadcast load, in flight with the mask
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (__ballot_sync(0xffffffffu, m[u] != 0) == 0) continue;  // the part does not reach these 8 rows
                apply_group_mask<N, FN, BGNAN>(base, m[u], lane, value_from_bits<N>(vbits[u]), bg);
                __syncwarp();
            }
        }
    }

    // ---- flush: every output byte is written exactly once ---------------------------------------------
    const uint32_t cols = c1 - c0;
    for (uint32_t rr = 0; rr < r1 - r0; rr++) {
        N* dst = out + ((size_t)band * T.out_rows + T.win_row_off + (r0 - P.win_r0) + rr) * P.ncols + c0;
        const N* src = rows8 + rr * TILE_C;
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
