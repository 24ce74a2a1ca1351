// rz_tiles.cuh — tile-binned polygon engine (dense output, jobs made of small polygon parts).
//
// The crossing-record pipeline (rz_kernels.cuh) materialises one 8-byte record per scanline crossing
// and sorts them: ~11 passes over 9 GB at BASELINE config 4.  When parts are small compared with the
// raster it is far cheaper never to materialise crossings.  Parts are binned to tiles of 128 columns x
// TILE_R rows and the scanline job is split into two embarrassingly parallel kernels:
//
//   tile_bin (x2)      per part: the tiles its bounding box overlaps; emits the mask units - (part, run of
//                      tile rows) handled by one warp - and the (tile, block) records, stably sorted by tile so
//                      every tile sees its parts in burn order; block_pos then gives every (part, tile) block
//                      its position in tile order.
//   tile_mask          one WARP per mask unit: reads the part's ring vertices from the world-coordinate pool,
//                      31 edges at a time, transforms them (edges.rs:94-97), finds with integer row compares
//                      the edges that cross the unit's rows (edges.rs:27-46, 90-110), computes their crossings
//                      (edges.rs:50-55, warp-flattened so all lanes stay busy) and XORs one bit per crossing
//                      into a shared-memory toggle mask spanning the part's tile columns; a prefix-XOR along
//                      each row turns it into the even-odd INSIDE mask (== sorting and pairing the crossings,
//                      burners.rs:302-315), written as one TILE_R x 128-bit block per (part, tile).
//   tile_apply         one CTA per 8 tiles of a tile row, pixels in REGISTERS (initialised to the background,
//                      flushed once).  Each warp owns 8 rows and, with no synchronisation with the other
//                      warps, walks the tile's parts in burn order: one coalesced 128-byte load brings the
//                      part's inside mask, lane = (row, 32-column word), and the lane applies the part's value
//                      to its own 32 pixels with the reference's pixel-function rule
//                      (pixel_functions.rs:56-123).
//
// Parts are applied strictly in burn order per pixel, so every pixel function stays bit-exact, and every
// output byte is written to HBM once.
#pragma once

#include <type_traits>

#include "rz_kernels.cuh"

namespace rz {

constexpr uint32_t TILE_C = 128;           // columns per tile = 4 mask words per row
constexpr uint32_t MASK_MAX_WORDS = 16;    // tile_mask: widest toggle-mask chunk kept in shared memory (512 columns)
constexpr uint32_t MASK_SMEM_WORDS = 1280; // tile_mask: toggle-mask words per warp (rows x (chunk words + 1 pad))
constexpr int MASK_WARPS = 4;
constexpr uint32_t MASK_UNITS = 4;  // consecutive mask units built by one warp
constexpr uint32_t VROW_RING_END = 0x80000000u;  // vertex tag: "last vertex of its ring"

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
    unsigned long long row_pairs;    // mask units: (part, run of tile rows) handled by one tile_mask warp
    unsigned long long edge_visits;  // sum over parts of units x column chunks x ring vertices
    unsigned long long cross_lb;     // lower bound of the crossing count: 2 per part row (a closed ring crosses a
                                     // row's centre line an even number of times, at least twice)
    unsigned int nonfinite;          // some part burns a NaN / infinite value (float dtypes)
    unsigned int eq_bg;              // some part burns a value whose bits equal the background's
};

// where a part's inside-mask blocks live: block(tr, tc) = first_block + (tr - tr0) * ntc + (tc - tc0)
struct PartTile {
    unsigned long long first_block;
    uint32_t tr0, tc0;
    uint32_t ntr, ntc;
    uint32_t r_lo, r_hi;  // rows the part can fill (absolute, clamped to the window)
};

// A mask unit = the tile rows [tr, tr + k) of one part, built by one tile_mask warp: as many tile rows as
// the part's rows in them fit the warp's shared-memory toggle mask.  Packed [tr:26 | k:6 | part:32].
__device__ __forceinline__ uint32_t mask_row_capacity(uint32_t ntc) {
    return MASK_SMEM_WORDS / (min(ntc * 4u, MASK_MAX_WORDS) + 1u);
}
template <typename F>
__device__ __forceinline__ uint32_t for_each_mask_unit(const KParams& P, uint32_t tile_r, uint32_t tr0, uint32_t ntr,
                                                       uint32_t ntc, uint32_t r_lo, uint32_t r_hi, F&& emit) {
    const uint32_t cap = mask_row_capacity(ntc);
    uint32_t n = 0, tr = tr0;
    const uint32_t tr_end = tr0 + ntr;
    while (tr < tr_end) {
        const uint32_t row_start = max(r_lo, P.win_r0 + tr * tile_r);
        uint32_t k = 1;
        while (tr + k < tr_end && k < 63u && min(r_hi, P.win_r0 + (tr + k + 1) * tile_r) - row_start <= cap) k++;
        emit(tr, k);
        tr += k;
        n++;
    }
    return n;
}

// Rust `f64 as usize` after floor / ceil, then min(., lim): cvt.rmi / cvt.rpi saturate (negatives and NaN
// give 0, huge values 2^32-1), which is exactly the cast's rule (edges.rs:32-33, burners.rs:310-311).
__device__ __forceinline__ uint32_t floor_sat_u32(double v, uint32_t lim) { return min(__double2uint_rd(v), lim); }
__device__ __forceinline__ uint32_t ceil_sat_u32(double v, uint32_t lim) { return min(__double2uint_ru(v), lim); }
// first row whose centre lies at or below pixel ordinate y: ystart / yend of edges.rs:32-33, clamped to nrows
__device__ __forceinline__ uint32_t vertex_row(const KParams& P, double y) {
    return ceil_sat_u32(__dsub_rn(y, 0.5), P.nrows);
}

// pixel rows / columns a polygon part can fill, from its world extent (one pixel of margin on columns)
__device__ __forceinline__ bool part_pixel_box(const KParams& P, double xlo, double xhi, double ylo, double yhi,
                                               uint32_t& r_lo, uint32_t& r_hi, uint32_t& c_lo, uint32_t& c_hi) {
    // rows whose centre can lie inside: [ceil(y_top - 0.5), ceil(y_bot - 0.5))
    // (world -> pixel and ceil(. - 0.5) are monotone, so every edge's rows lie inside [a, b) exactly)
    const uint32_t a = vertex_row(P, px_y(P, yhi)), b = vertex_row(P, px_y(P, ylo));
    r_lo = max(a, P.win_r0);
    r_hi = min(b, P.win_r1);
    uint32_t cl = sat_u32(floor(__dadd_rn(px_x(P, xlo), 0.5)), P.ncols);
    uint32_t ch = sat_u32(floor(__dadd_rn(px_x(P, xhi), 0.5)), P.ncols);
    c_lo = cl > 0 ? cl - 1 : 0u;
    c_hi = ch < P.ncols ? ch + 1 : P.ncols;
    return r_hi > r_lo && c_hi > c_lo && c_lo < P.ncols;
}

// mode 0: per part the number of (part,tile) pairs and of (part,tile-row) pairs (+ totals);
// mode 1: with the scanned offsets, fill PartTile and emit the row-pair list and the tile records
// [tile | block], block = index of the (part,tile) pair's inside-mask block.
static __global__ void tile_bin_kernel(KParams P, TileParams T, const PartInfo* __restrict__ info,
                                const double* __restrict__ xlo, const double* __restrict__ xhi,
                                const double* __restrict__ ylo, const double* __restrict__ yhi,
                                const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                                uint32_t* __restrict__ cnt_tiles, uint32_t* __restrict__ cnt_rows,
                                const unsigned long long* __restrict__ off_tiles,
                                const unsigned long long* __restrict__ off_rows, PartTile* __restrict__ pt,
                                uint64_t* __restrict__ row_pairs, uint64_t* __restrict__ recs,
                                unsigned long long* __restrict__ block_value, uint32_t block_bits,
                                TileCounters* __restrict__ tc, int mode, int float_bytes, unsigned long long bg_bits,
                                unsigned long long value_mask) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ntr = 0, ntc = 0, n_units = 0, r_lo = 0, r_hi = 0, c_lo = 0, c_hi = 0;
    if (p < P.n_parts) {
        const int32_t band = info[p].band;
        if (band >= 0 && vend[p] > vbeg[p] + 1 &&
            part_pixel_box(P, xlo[p], xhi[p], ylo[p], yhi[p], r_lo, r_hi, c_lo, c_hi)) {
            const uint32_t tr0 = (r_lo - P.win_r0) / T.tile_r, tr1 = (r_hi - 1 - P.win_r0) / T.tile_r;
            const uint32_t tc0 = c_lo / TILE_C, tc1 = (c_hi - 1) / TILE_C;
            ntr = tr1 - tr0 + 1;
            ntc = tc1 - tc0 + 1;
            if (mode == 0) {
                n_units = for_each_mask_unit(P, T.tile_r, tr0, ntr, ntc, r_lo, r_hi, [](uint32_t, uint32_t) {});
            } else {
                PartTile q;
                q.first_block = off_tiles[p];
                q.tr0 = tr0;
                q.tc0 = tc0;
                q.ntr = ntr;
                q.ntc = ntc;
                q.r_lo = r_lo;
                q.r_hi = r_hi;
                pt[p] = q;
                unsigned long long o = off_tiles[p], orow = off_rows[p];
                for_each_mask_unit(P, T.tile_r, tr0, ntr, ntc, r_lo, r_hi, [&](uint32_t tr, uint32_t k) {
                    row_pairs[orow++] = ((uint64_t)tr << 38) | ((uint64_t)k << 32) | p;
                });
                for (uint32_t tr = tr0; tr <= tr1; tr++) {
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
            cnt_rows[p] = n_units;
            if (float_bytes) {  // exponent all ones: NaN or infinity (tile_apply's additive mode needs finite values)
                const unsigned long long vb = info[p].value_bits;
                const bool bad = float_bytes == 4 ? ((vb >> 23) & 0xffu) == 0xffu : ((vb >> 52) & 0x7ffu) == 0x7ffu;
                if (bad && info[p].band >= 0) atomicOr(&tc->nonfinite, 1u);
            }
            if (info[p].band >= 0 && ((info[p].value_bits ^ bg_bits) & value_mask) == 0) atomicOr(&tc->eq_bg, 1u);
        }
        const uint32_t chunks = (ntc * 4 + MASK_MAX_WORDS - 1) / MASK_MAX_WORDS;
        unsigned long long pairs = (unsigned long long)ntr * ntc, rows = n_units,
                           visits = (unsigned long long)n_units * chunks * (p < P.n_parts ? vend[p] - vbeg[p] : 0u),
                           cross = ntr ? 2ull * (r_hi - r_lo) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pairs += __shfl_down_sync(0xffffffffu, pairs, o);
            rows += __shfl_down_sync(0xffffffffu, rows, o);
            visits += __shfl_down_sync(0xffffffffu, visits, o);
            cross += __shfl_down_sync(0xffffffffu, cross, o);
        }
        if (lane_id() == 0 && pairs) {
            atomicAdd(&tc->pairs, pairs);
            atomicAdd(&tc->row_pairs, rows);
            atomicAdd(&tc->edge_visits, visits);
            atomicAdd(&tc->cross_lb, cross);
        }
    }
}

struct InU32 {
    const uint32_t* v;
    __device__ unsigned long long operator()(uint32_t i) const { return v[i]; }
};

// One ring edge (pixel-space vertices): its crossing with a row's centre line.
struct TileEdge {
    double x_top, y_top, dxdy;
};
// ... as tile_mask keeps it in shared memory: plus its first row in the unit and the index of its first crossing
struct alignas(16) MaskEdge {
    double x_top, y_top, dxdy;
    uint32_t lo, pre;
};
// false when the reference skips the edge as horizontal (edges.rs:100)
__device__ __forceinline__ bool tile_edge_slope(double x0, double y0, double x1, double y1, TileEdge& e) {
    if (!(fabs(__dsub_rn(y0, y1)) >= DBL_EPSILON)) return false;
    const bool down = y0 < y1;  // edges.rs:29
    const double x_bot = down ? x1 : x0, y_bot = down ? y1 : y0;
    e.x_top = down ? x0 : x1;
    e.y_top = down ? y0 : y1;
    e.dxdy = __ddiv_rn(__dsub_rn(x_bot, e.x_top), __dsub_rn(y_bot, e.y_top));  // edges.rs:36
    return true;
}
__device__ __forceinline__ uint32_t tile_edge_col(const KParams& P, const TileEdge& e, uint32_t row) {
    const double cy = __dadd_rn((double)row, 0.5);
    const double xi = __dadd_rn(e.x_top, __dmul_rn(__dsub_rn(cy, e.y_top), e.dxdy));  // edges.rs:50-55
    return floor_sat_u32(__dadd_rn(xi, 0.5), P.ncols);                                // burners.rs:310-311
}

// ---------------------------------------------------------------------------------------------
// streamed upload: the polygon pool is pulled from page-locked host memory window by window
// ---------------------------------------------------------------------------------------------
// An end-to-end call is bound by PCIe: geometry host->device, then the raster device->host.  The two
// directions are independent, so instead of uploading the whole pool first, every polygon part is assigned
// to the first row window of the call that its rows touch ("bucket"; parts touching none go to a last one),
// and the parts of bucket b are copied by a kernel that reads the mapped host arrays directly, right before
// window b is computed - while the copy stream is still sending window b-1 to the host.
static __global__ void part_bucket_kernel(KParams P, const uint8_t* __restrict__ part_kind, const double* __restrict__ ylo,
                                   const double* __restrict__ yhi, uint32_t shard_r0, uint32_t shard_r1,
                                   uint32_t win_rows, uint32_t n_buckets, uint32_t* __restrict__ bucket_of,
                                   unsigned int* __restrict__ cnt) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_parts) return;
    uint32_t b = 0xffffffffu;  // not a polygon part
    if (part_kind[p] == 0) {
        const uint32_t lo = max(vertex_row(P, px_y(P, yhi[p])), shard_r0);
        const uint32_t hi = min(vertex_row(P, px_y(P, ylo[p])), shard_r1);
        b = hi > lo ? min((lo - shard_r0) / win_rows, n_buckets - 1) : n_buckets;
        atomicAdd(&cnt[b], 1u);
    }
    bucket_of[p] = b;
}
// cnt[0 .. n_buckets] -> off[0 .. n_buckets + 1] (exclusive), cursors reset; a handful of buckets: one thread
static __global__ void bucket_scan_kernel(unsigned int* __restrict__ cnt, unsigned int* __restrict__ off, uint32_t n) {
    unsigned int run = 0;
    for (uint32_t i = 0; i < n; i++) {
        off[i] = run;
        run += cnt[i];
        cnt[i] = 0;
    }
    off[n] = run;
}
static __global__ void bucket_scatter_kernel(uint32_t n_parts, const uint32_t* __restrict__ bucket_of,
                                      const unsigned int* __restrict__ off, unsigned int* __restrict__ cursor,
                                      uint32_t* __restrict__ order) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_parts) return;
    const uint32_t b = bucket_of[p];
    if (b != 0xffffffffu) order[off[b] + atomicAdd(&cursor[b], 1u)] = p;
}
// one CTA per part: its vertex range from the mapped host arrays (PCIe reads) to the device pool
static __global__ void __launch_bounds__(128)
pull_parts_kernel(const uint32_t* __restrict__ order, uint32_t n, const uint32_t* __restrict__ vbeg,
                  const uint32_t* __restrict__ vend, const double* __restrict__ hx, const double* __restrict__ hy,
                  const uint32_t* __restrict__ htag, double* __restrict__ dx, double* __restrict__ dy,
                  uint32_t* __restrict__ dtag) {
    if (blockIdx.x >= n) return;
    const uint32_t p = order[blockIdx.x];
    const uint32_t e = vend[p];
    for (uint32_t i = vbeg[p] + threadIdx.x; i < e; i += 4 * 128) {  // four independent loads per array in flight
        double a[4], b[4];
        uint32_t t[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = i + k * 128;
            if (j < e) {
                a[k] = __ldcs(hx + j);
                b[k] = __ldcs(hy + j);
                t[k] = __ldcs(htag + j);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t j = i + k * 128;
            if (j < e) {
                dx[j] = a[k];
                dy[j] = b[k];
                dtag[j] = t[k];
            }
        }
    }
}

// After the stable sort of the [tile | block] records: where every inside-mask block lives.  tile_mask writes
// block b at position pos[b] of the mask array, i.e. in tile order, and the value of its part goes to the
// same position, so that tile_apply streams a tile's blocks from consecutive memory with no indirection.
static __global__ void block_pos_kernel(const uint64_t* __restrict__ recs, uint32_t n, uint32_t block_bits,
                                 const unsigned long long* __restrict__ block_value, uint32_t* __restrict__ pos,
                                 unsigned long long* __restrict__ value_sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t blk = (uint32_t)(recs[i] & ((1ull << block_bits) - 1ull));
    pos[blk] = i;
    value_sorted[i] = block_value[blk];
}

// ---------------------------------------------------------------------------------------------
// tile_mask: one warp per mask unit = (part, run of tile rows)
// ---------------------------------------------------------------------------------------------
// A row can only have an odd number of crossings when a non-horizontal edge was skipped for being
// shorter than f64::EPSILON in y (edges.rs:100) while still straddling a pixel centre.  Pixel-centre
// ordinates k+0.5 with k >= 1 are spaced >= EPSILON apart, so that can only happen on raster row 0
// (centre 0.5): only that row's crossing count is tracked, and an odd row 0 drops its largest column like
// chunks_exact(2) drops the unpaired tail (burners.rs:305).
template <int TILE_R>
static __global__ void __launch_bounds__(MASK_WARPS * 32, 8)
tile_mask_kernel(KParams P, TileParams T, const uint64_t* __restrict__ units, uint32_t n_units,
                 const PartTile* __restrict__ pt, const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                 const double* __restrict__ wx, const double* __restrict__ wy, const uint32_t* __restrict__ tag,
                 const uint32_t* __restrict__ pos, uint32_t* __restrict__ masks) {
    __shared__ uint32_t s_mask[MASK_WARPS][MASK_SMEM_WORDS];
    __shared__ MaskEdge s_edge[MASK_WARPS][32];  // the batch's active edges, compacted
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    // A warp builds MASK_UNITS consecutive units (units are in part order, so are their vertices).  While it works
    // on one it asks the L2 for the next one's part record and vertex range: a unit lives for ~20 us and would
    // otherwise start with three dependent global loads.
    const uint32_t unit0 = (blockIdx.x * MASK_WARPS + warp) * MASK_UNITS;
    for (uint32_t unit = unit0; unit < min(unit0 + MASK_UNITS, n_units); unit++) {
    const uint64_t un = units[unit];
    const uint32_t part = (uint32_t)un, k_tr = (uint32_t)(un >> 32) & 63u, tr = (uint32_t)(un >> 38);
    if (unit + 1 < n_units && lane < 3) {
        const uint32_t pn = (uint32_t)units[unit + 1];
        const void* a = lane == 0 ? (const void*)(pt + pn) : lane == 1 ? (const void*)(vbeg + pn) : (const void*)(vend + pn);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
    const PartTile q = pt[part];
    const uint32_t vb = vbeg[part], ve = vend[part];
    if (lane < 12) {  // the vertices after this part's are (most often) the next unit's first batch
        const uint32_t k = lane >> 2;  // 0: x, 1: y, 2: tag; four 128-byte lines each (tags: two)
        const char* base = k == 0 ? (const char*)(wx + ve) : k == 1 ? (const char*)(wy + ve) : (const char*)(tag + ve);
        if (k < 2 || (lane & 3u) < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (lane & 3u) * 128));
    }
    const uint32_t t0 = P.win_r0 + tr * TILE_R;  // first row of the unit's tile rows
    const uint32_t row_start = max(q.r_lo, t0), row_end = min(q.r_hi, t0 + k_tr * TILE_R);
    const uint32_t n_rows = row_end - row_start;
    uint32_t* mask = s_mask[warp];
    const uint32_t words_total = q.ntc * 4;
    const uint32_t lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);

    for (uint32_t w0 = 0; w0 < words_total; w0 += MASK_MAX_WORDS) {  // one pass per 512-column chunk
        const uint32_t nw = min(MASK_MAX_WORDS, words_total - w0);
        const uint32_t stride = nw + 1;                               // odd or padded: rows spread over banks
        const uint32_t c0 = q.tc0 * TILE_C + w0 * 32;                 // first pixel column of the chunk
        const uint32_t c1 = min(c0 + nw * 32, P.ncols);
        for (uint32_t i = lane; i < n_rows * stride; i += 32) mask[i] = 0;
        __syncwarp();
        uint32_t par0 = 0;  // this lane's share of the row-0 crossing count parity
        // Batches of 31 edges: lane l holds ring vertex i0 + l (world coordinates + tag, prefetched one batch
        // ahead) and, for l < 31, edge (i0 + l, i0 + l + 1) whose second vertex comes from lane l + 1.  The
        // vertex goes to pixel space here (edges.rs:94-97) and its row index ceil(py - 0.5), clamped to
        // [0, nrows] (edges.rs:32-33), decides with integer compares which edges cross this unit's rows: an
        // edge is active on rows [min(row_i, row_i+1), max(..)), which also implies the reference's culling
        // test (edges.rs:105).  Only those edges pay for the x divides and the slope.
        double xw_n = 0.0, yw_n = 0.0;
        uint32_t tg_n = VROW_RING_END;
        if (vb + lane < ve) {
            xw_n = wx[vb + lane];
            yw_n = wy[vb + lane];
            tg_n = tag[vb + lane];
        }
        for (uint32_t i0 = vb; i0 + 1 < ve; i0 += 31) {
            const double xw = xw_n, yw = yw_n;
            const uint32_t tg = tg_n;
            tg_n = VROW_RING_END;
            if (i0 + 31 + lane < ve) {
                xw_n = wx[i0 + 31 + lane];
                yw_n = wy[i0 + 31 + lane];
                tg_n = tag[i0 + 31 + lane];
            }
            const double y0 = px_y(P, yw), x0 = px_x(P, xw);  // one transform per vertex, shared by its two edges
            const uint32_t row0 = vertex_row(P, y0);
            const double y1 = __shfl_down_sync(0xffffffffu, y0, 1);
            const uint32_t row1 = __shfl_down_sync(0xffffffffu, row0, 1);
            const double x1 = __shfl_down_sync(0xffffffffu, x0, 1);
            const uint32_t lo = max(min(row0, row1), row_start), hi = min(max(row0, row1), row_end);
            uint32_t cnt = (lane < 31 && i0 + lane + 1 < ve && !(tg & VROW_RING_END) && hi > lo) ? hi - lo : 0u;
            TileEdge e;
            if (cnt && !tile_edge_slope(x0, y0, x1, y1, e)) cnt = 0;
            par0 ^= (cnt != 0 && lo == 0);  // a kept edge starting at row 0 has exactly one crossing on it
            const uint32_t act = __ballot_sync(0xffffffffu, cnt != 0);
            if (act == 0) continue;
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += up;
            }
            const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t pre = inc - cnt;
            if (cnt) {  // active edges are compacted: slot = rank among the batch's active edges
                const uint32_t slot = __popc(act & lt_mask);
                MaskEdge me;
                me.x_top = e.x_top;
                me.y_top = e.y_top;
                me.dxdy = e.dxdy;
                me.lo = lo;
                me.pre = pre;
                s_edge[warp][slot] = me;
            }
            __syncwarp();
            // all lanes share the batch's crossings evenly: crossing kk belongs to the last active edge whose
            // first crossing is <= kk, found from the bit pattern of the first-crossing positions
            // Two chunks of 32 crossings per iteration, evaluated without branches up to the atomic, so that the
            // two dependent f64 chains (shared-memory load -> 5 double ops -> convert) overlap.
            uint32_t cum = 0;
            auto crossing = [&](uint32_t slot, uint32_t kk, uint32_t& word, uint32_t& bit) -> bool {
                const MaskEdge me = s_edge[warp][min(slot, 31u)];  // two 16-byte shared-memory loads
                TileEdge b;
                b.x_top = me.x_top;
                b.y_top = me.y_top;
                b.dxdy = me.dxdy;
                const uint32_t row = me.lo + (kk - me.pre);
                // floor(x + 0.5) as usize (burners.rs:310-311); the clamp to ncols is implied by `col < c1` below
                const double cy = __dadd_rn((double)row, 0.5);
                const uint32_t col = __double2uint_rd(__dadd_rn(__dadd_rn(b.x_top, __dmul_rn(__dsub_rn(cy, b.y_top), b.dxdy)), 0.5));
                const uint32_t rel = col <= c0 ? 0u : col - c0;  // left of the chunk: parity carry-in at bit 0
                word = (row - row_start) * stride + (rel >> 5);
                bit = 1u << (rel & 31);
                return kk < wtot && col < c1;  // a crossing right of the chunk has no effect on its pixels
            };
            for (uint32_t k0 = 0; k0 < wtot; k0 += 64) {
                const uint32_t da = pre - k0, db = da - 32u;
                const uint32_t sa = __reduce_or_sync(0xffffffffu, (cnt && da < 32u) ? 1u << da : 0u);
                const uint32_t sb = __reduce_or_sync(0xffffffffu, (cnt && db < 32u) ? 1u << db : 0u);
                const uint32_t slot_a = cum + __popc(sa & le_mask) - 1u;
                cum += __popc(sa);
                const uint32_t slot_b = cum + __popc(sb & le_mask) - 1u;
                cum += __popc(sb);
                uint32_t wa, ba, wb, bb;
                const bool oka = crossing(slot_a, k0 + lane, wa, ba);
                const bool okb = crossing(slot_b, k0 + 32 + lane, wb, bb);
                if (oka) atomicXor(&mask[wa], ba);
                if (okb) atomicXor(&mask[wb], bb);
            }
            __syncwarp();
        }
        if (row_start == 0 && (__popc(__ballot_sync(0xffffffffu, par0 & 1u)) & 1)) {  // rare: odd row 0
            uint32_t mx = 0;
            for (uint32_t i = vb + lane; i + 1 < ve; i += 32) {
                if (tag[i] & VROW_RING_END) continue;
                const double y0 = px_y(P, wy[i]), y1 = px_y(P, wy[i + 1]);
                const uint32_t va = vertex_row(P, y0), vn = vertex_row(P, y1);
                if (!(min(va, vn) == 0 && max(va, vn) > 0)) continue;
                TileEdge e;
                if (tile_edge_slope(px_x(P, wx[i]), y0, px_x(P, wx[i + 1]), y1, e)) mx = max(mx, tile_edge_col(P, e, 0) + 1u);
            }
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0 && mx && mx - 1 < c1) {
                const uint32_t rel = mx - 1 <= c0 ? 0u : mx - 1 - c0;
                mask[rel >> 5] ^= 1u << (rel & 31);
            }
        }
        __syncwarp();
        // toggle mask -> inside mask, row by row (one lane per row), in place
        for (uint32_t rr = lane; rr < n_rows; rr += 32) {
            uint32_t carry = 0;
            uint32_t* mrow = mask + rr * stride;
            for (uint32_t wd = 0; wd < nw; wd++) {
                uint32_t m = mrow[wd];
                m ^= m << 1;
                m ^= m << 2;
                m ^= m << 4;
                m ^= m << 8;
                m ^= m << 16;
                m ^= carry;
                carry = (m >> 31) ? 0xffffffffu : 0u;
                mrow[wd] = m;
            }
        }
        __syncwarp();
        // one TILE_R x 4-word block per (part, tile): consecutive lanes write consecutive words; rows of the
        // tile outside the part's rows are empty
        for (uint32_t j = 0; j < k_tr; j++) {
            const unsigned long long row_base = q.first_block + (unsigned long long)(tr + j - q.tr0) * q.ntc + w0 / 4;
            const uint32_t tj = t0 + j * TILE_R;
            for (uint32_t tcl = 0; tcl * 4 < nw; tcl++) {
                uint32_t* dst = masks + (size_t)pos[row_base + tcl] * (TILE_R * 4);  // tile order
                for (uint32_t i = lane; i < TILE_R * 4; i += 32) {
                    const uint32_t rel = tj + (i >> 2) - row_start;  // wraps above n_rows for rows before row_start
                    dst[i] = rel < n_rows ? mask[rel * stride + tcl * 4 + (i & 3)] : 0u;
                }
            }
        }
        __syncwarp();
    }
    }  // units of this warp
}

// ---------------------------------------------------------------------------------------------
// tile_apply: one CTA per tile, each warp owns 8 rows held in REGISTERS, no synchronisation between warps
// ---------------------------------------------------------------------------------------------
// A part's inside mask arrives as one coalesced 128-byte load per warp: lane 4r + w holds the 32-bit word of
// (row r, columns 32w .. 32w+31) of the warp's 8 x 128 pixels - and that lane OWNS those 32 pixels, in
// registers px[0..31] (initialised to the background: geo/raster.rs:23-28).  Applying a part is therefore
// lane-local: no shuffle, no shared memory; bit b of the lane's own mask word decides whether pixel b takes the
// part's value through the reference's pixel-function rule (pixel_functions.rs:56-123), parts in burn order.
//
// MODE selects how the rule is evaluated (all bit-exact with the reference):
//   0  generic: apply_px on every pixel slot, selected by the mask bit.
//   1  additive with a touched mask - `sum` / `count` on a float dtype with a NaN background when every burn value
//      is finite (checked on the device, TileCounters::nonfinite).  The running value can then never become NaN,
//      so "untouched" is just "never written": pixels start at -0.0 (the additive identity: -0.0 + v == v for
//      every v, signed zeros included), a masked add is ONE predicated instruction, the lane ORs the mask word
//      into its touched word, and untouched pixels become the background at the flush.
//   2  additive, plain - `sum` / `count` on an integer dtype with background 0: `cur == bg ? v : cur + v` is
//      `cur + v` for every cur (wrapping), no touched mask needed.
//   3  `first` / `min` / `max` with a touched mask, when no write can ever make a pixel look untouched again: no
//      burn value equals the background (integer dtypes; TileCounters::eq_bg) or, for a NaN background, no value
//      is NaN.  `min` / `max` then start from the type's largest / smallest value and are one compare-select,
//      `first` writes where the mask bit is set and the touched bit is not.
// (Measured alternatives on config 4: pixels in shared memory 6.65 ms; registers with one column per lane and
// shuffled mask words 5.30 ms; four consecutive pixels per lane 6.09 ms.)
// the largest (+inf) / smallest (-inf) value of a dtype, from its bit pattern
template <typename N> __device__ __forceinline__ N type_extreme(bool largest) {
    constexpr int bits = 8 * sizeof(N);
    uint64_t b;
    if (std::is_floating_point<N>::value) b = sizeof(N) == 8 ? (largest ? 0x7ff0000000000000ull : 0xfff0000000000000ull)
                                                             : (largest ? 0x7f800000ull : 0xff800000ull);
    else if (std::is_signed<N>::value) b = largest ? ((1ull << (bits - 1)) - 1ull) : (1ull << (bits - 1));
    else b = largest ? ~0ull : 0ull;
    return value_from_bits<N>(b);
}

template <typename N, int FN, int MODE, bool BGNAN>
__device__ __forceinline__ void apply_part_word(N (&px)[32], uint32_t mw, N v, N bg) {
#pragma unroll
    for (int b = 0; b < 32; b++) {
        if (MODE == 0) {
            const N nv = apply_px<N, FN, BGNAN, true>(px[b], v, bg);
            if (mw & (1u << b)) px[b] = nv;
        } else if (MODE == 3) {  // mw has the already-written pixels removed for `first`
            if (mw & (1u << b)) px[b] = FN == RZ_FIRST ? v : (FN == RZ_MIN ? (px[b] > v ? v : px[b]) : (px[b] < v ? v : px[b]));
        } else {
            if (mw & (1u << b)) px[b] = add_v(px[b], FN == RZ_COUNT ? (N)1 : v);
        }
    }
}

constexpr int APPLY_TILES = 8;  // consecutive tiles of one tile row handled by one CTA

template <typename N, int FN, int TILE_R, int MODE, bool BGNAN>
static __global__ void __launch_bounds__(TILE_R * 4, 4)
tile_apply_kernel(KParams P, TileParams T, const uint32_t* __restrict__ tile_start,
                  const unsigned long long* __restrict__ value_sorted, const uint32_t* __restrict__ masks,
                  uint64_t bg_bits, N* __restrict__ out) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const N bg = value_from_bits<N>(bg_bits);

    // grid = (groups of APPLY_TILES tile columns, tile rows x bands), or 1-D when that does not fit the grid limits
    const uint32_t groups = (T.n_tc + APPLY_TILES - 1) / APPLY_TILES;
    uint32_t tgrp, trow, band;
    if (gridDim.y > 1 || T.n_tr * P.n_bands == 1) {
        tgrp = blockIdx.x;
        trow = blockIdx.y;
        band = 0;
        if (P.n_bands > 1) {
            band = trow / T.n_tr;
            trow -= band * T.n_tr;
        }
    } else {
        const uint32_t tt = blockIdx.x;
        tgrp = tt % groups;
        trow = (tt / groups) % T.n_tr;
        band = tt / (groups * T.n_tr);
    }
    const uint32_t tcol0 = tgrp * APPLY_TILES, n_here = min((uint32_t)APPLY_TILES, T.n_tc - tcol0);
    const uint32_t t0 = (band * T.n_tr + trow) * T.n_tc + tcol0;
    const uint32_t r0 = P.win_r0 + trow * TILE_R + warp * 8;
    if (r0 >= P.win_r1) return;

    N ident;
    if (MODE == 1) {  // -0.0: the additive identity of IEEE addition
        const uint64_t neg0 = sizeof(N) == 8 ? 0x8000000000000000ull : 0x80000000ull;
        ident = value_from_bits<N>(neg0);
    } else if (MODE == 3 && FN == RZ_MIN) {
        ident = type_extreme<N>(true);
    } else if (MODE == 3 && FN == RZ_MAX) {
        ident = type_extreme<N>(false);
    } else {
        ident = bg;  // MODE 0: the background; MODE 2: bg == 0; MODE 3 first: never read before written
    }

    // A tile's blocks [beg, end) are consecutive in memory, parts in burn order, and so are the tiles of a tile
    // row.  A ring of APPLY_DEPTH blocks is kept in flight (one coalesced 128-byte load plus one broadcast load of
    // the part's value per block); when a tile's blocks are used up the ring is primed with the NEXT tile's
    // first blocks before this tile is flushed, so only the first tile of a CTA waits for its first loads.
    constexpr int APPLY_DEPTH = 4;
    const uint32_t* my_masks = masks + warp * 32 + lane;
    uint32_t m[APPLY_DEPTH];
    N val[APPLY_DEPTH];  // the value occupies the low bytes of its 8-byte slot
    uint32_t beg = tile_start[t0], end = tile_start[t0 + 1];
#pragma unroll
    for (int u = 0; u < APPLY_DEPTH; u++) {
        const bool live = beg + u < end;
        m[u] = live ? my_masks[(size_t)(beg + u) * (TILE_R * 4)] : 0u;
        val[u] = live ? *reinterpret_cast<const N*>(value_sorted + beg + u) : bg;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    for (uint32_t ti = 0; ti < n_here; ti++) {
    const uint32_t end_next = ti + 1 < n_here ? tile_start[t0 + ti + 2] : end;  // needed after this tile's blocks
    const uint32_t c0 = (tcol0 + ti) * TILE_C;
    N px[32];  // pixel b = (row r0 + (lane >> 2), column c0 + 32 * (lane & 3) + b)
    uint32_t touched = 0;
#pragma unroll
    for (int b = 0; b < 32; b++) px[b] = ident;

    for (uint32_t j0 = beg; j0 < end; j0 += APPLY_DEPTH) {
#pragma unroll
        for (int u = 0; u < APPLY_DEPTH; u++) {
            const uint32_t mu = m[u];
            const N v = val[u];
            const uint32_t nxt = j0 + u + APPLY_DEPTH;  // refill this slot of the ring
            const bool live = nxt < end;
            m[u] = live ? my_masks[(size_t)nxt * (TILE_R * 4)] : 0u;
            val[u] = live ? *reinterpret_cast<const N*>(value_sorted + nxt) : bg;
            if (__ballot_sync(0xffffffffu, mu != 0) == 0) continue;  // the part does not reach these 8 rows
            if (MODE == 3) {
                apply_part_word<N, FN, 3, BGNAN>(px, FN == RZ_FIRST ? (mu & ~touched) : mu, v, bg);
                touched |= mu;
            } else if (MODE != 0) {
                apply_part_word<N, FN, MODE, BGNAN>(px, mu, v, bg);
                touched |= mu;
            } else if (FN == RZ_SUM && is_nan_v(v)) {
                // sum: a NaN value replaces the pixel (pixel_functions.rs:56-65), i.e. behaves like `last`
                apply_part_word<N, RZ_LAST, 0, BGNAN>(px, mu, v, bg);
            } else {
                apply_part_word<N, FN, 0, BGNAN>(px, mu, v, bg);
            }
        }
    }
    // the ring is drained: prime it with the next tile's first blocks, in flight during the flush below
#pragma unroll
    for (int u = 0; u < APPLY_DEPTH; u++) {
        const bool live = end + u < end_next;
        m[u] = live ? my_masks[(size_t)(end + u) * (TILE_R * 4)] : 0u;
        val[u] = live ? *reinterpret_cast<const N*>(value_sorted + end + u) : bg;
    }
    if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int b = 0; b < 32; b++)
            if (!(touched & (1u << b))) px[b] = bg;
    }

    // ---- flush: every output byte is written exactly once ---------------------------------------------
    // A lane's 32 pixels are contiguous in one output row; storing them directly would make every store
    // instruction touch 32 different 128-byte lines.  The warp transposes through shared memory instead (16-byte
    // chunks, segments padded by 16 bytes so that both directions are bank-conflict free) and writes whole rows
    // with 16-byte streaming stores, 512 contiguous bytes per instruction.
    constexpr uint32_t SEGB = 32 * sizeof(N) + 16, ROWB = 4 * SEGB;  // bytes of one 32-pixel segment / one row
    constexpr int K16 = 2 * sizeof(N);                                // 16-byte chunks per segment
    const uint32_t rows_here = min(8u, P.win_r1 - r0);
    N* gbase = out + ((size_t)band * T.out_rows + T.win_row_off + (r0 - P.win_r0)) * P.ncols + c0;
    if (T.vec_ok && c0 + TILE_C <= P.ncols) {
        unsigned char* stage = smem_raw + (size_t)warp * 8 * ROWB;
        unsigned char* mine = stage + (lane >> 2) * ROWB + (lane & 3u) * SEGB;
#pragma unroll
        for (int k = 0; k < K16; k++) {
            alignas(16) N q[16 / sizeof(N)];
#pragma unroll
            for (int j = 0; j < (int)(16 / sizeof(N)); j++) q[j] = px[k * (16 / sizeof(N)) + j];
            *reinterpret_cast<uint4*>(mine + 16 * k) = *reinterpret_cast<const uint4*>(q);
        }
        __syncwarp();
        constexpr int CH = 4 * K16;  // 16-byte chunks per 128-pixel row
        for (uint32_t rr = 0; rr < rows_here; rr++) {
            uint4* drow = reinterpret_cast<uint4*>(gbase + (size_t)rr * P.ncols);
#pragma unroll
            for (int j0 = 0; j0 < CH; j0 += 32) {
                const int j = j0 + (int)lane;
                if (CH >= 32 || j < CH)
                    __stcs(drow + j, *reinterpret_cast<const uint4*>(stage + rr * ROWB + (j / K16) * SEGB + 16 * (j % K16)));
            }
        }
    } else {
        const uint32_t row = r0 + (lane >> 2), col = c0 + (lane & 3u) * 32u;
        if (row < P.win_r1) {
            N* dst = gbase + (size_t)(lane >> 2) * P.ncols + (lane & 3u) * 32u;
#pragma unroll
            for (int b = 0; b < 32; b++)
                if (col + b < P.ncols) dst[b] = px[b];
        }
    }
    __syncwarp();  // the staging rows are reused by the next tile
    beg = end;
    end = end_next;
    }  // tiles of this CTA
}

}  // namespace rz
