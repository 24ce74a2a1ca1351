// rz_tiles.cuh — tile-binned polygon engine (dense output, jobs made of polygon parts).
//
// The crossing-record pipeline (rz_kernels.cuh) materialises one 8-byte record per scanline crossing
// and sorts them: ~11 passes over 9 GB at BASELINE config 4.  When parts are small compared with the
// raster it is far cheaper never to materialise crossings.  Parts are binned to tiles of 128 columns x
// TILE_R rows and the scanline job is split into two embarrassingly parallel kernels:
//
//   tile_bin (x2)      per part: the tiles its bounding box overlaps.  Pass 0 counts (part,tile) pairs, mask
//                      units and mask words per part; one look-back scan turns the three counts into offsets;
//                      pass 1 emits the mask units - (part, run of tile rows) handled by one warp - one
//                      descriptor per (part,tile) BLOCK and the [tile | block] records, which are then stably
//                      sorted by tile so that every tile sees its parts in burn order.
//   tile_mask          one WARP per mask unit: reads the part's ring vertices from the world-coordinate pool,
//                      31 edges at a time, transforms them (edges.rs:94-97), finds with integer row compares
//                      the edges that cross the unit's rows (edges.rs:27-46, 90-110), computes their crossings
//                      (edges.rs:50-55, warp-flattened so all lanes stay busy) and XORs one bit per crossing
//                      into a shared-memory toggle mask spanning the part's 32-column words; a prefix-XOR
//                      along each row turns it into the even-odd INSIDE mask (== sorting and pairing the
//                      crossings, burners.rs:302-315).  The mask is stored COMPACT: a block holds only the
//                      rows and the 32-column words of its tile that the part's box covers, and the blocks of
//                      a part are consecutive in memory (config 4: ~1 GB instead of 4.3 GB of full
//                      64 x 128-bit blocks).
//   tile_apply         one CTA per 8 tiles of a tile row, pixels in REGISTERS (initialised to the background,
//                      flushed once).  Each warp owns 8 rows and, with no synchronisation with the other
//                      warps, walks the tile's block descriptors in burn order: lane = (row, 32-column word)
//                      fetches its word of the part's inside mask (if the block covers it) and applies the
//                      part's value to its own 32 pixels with the reference's pixel-function rule
//                      (pixel_functions.rs:56-123).
//
// Parts are applied strictly in burn order per pixel, so every pixel function stays bit-exact, and every
// output byte is written to HBM once.  Every count the host would need in the middle of a call (pairs, units,
// mask words, "some value is NaN") stays on the device: buffers are sized from an upper bound cached per
// (geometry set, grid) and the kernels read the actual counts from TileCounters, so a steady-state call has
// no host synchronisation between its first and its last kernel.
#pragma once

#include <type_traits>

#include "rz_kernels.cuh"

namespace rz {

constexpr uint32_t TILE_C = 128;           // columns per tile = 4 mask words per row
constexpr uint32_t MASK_MAX_WORDS = 16;    // tile_mask: widest toggle-mask chunk kept in shared memory (512 columns)
constexpr uint32_t MASK_SMEM_WORDS = 1280; // tile_mask: toggle-mask words per warp (rows x chunk stride)
constexpr int MASK_WARPS = 4;
constexpr uint32_t MASK_UNITS = 4;  // consecutive mask units built by one warp
constexpr uint32_t VROW_RING_END = 0x80000000u;  // vertex tag: "last vertex of its ring"

struct TileParams {
    uint32_t tile_r;           // rows per tile (64, or 32 for 8-byte dtypes)
    uint32_t n_tc, n_tr;       // tile grid of one band in this window
    uint32_t n_tiles;          // n_bands * n_tr * n_tc
    uint32_t block_bits;       // record = [tile | block index]
    uint32_t win_row_off, out_rows;
    uint32_t vec_ok;
    uint32_t cap_pairs, cap_units;  // capacity of the record / unit arrays (upper bounds cached on the host)
    uint32_t apply_tiles;           // consecutive tiles of a tile row per tile_apply CTA (1 .. APPLY_TILES)
};

struct TileCounters {
    unsigned long long pairs;        // (part, tile) pairs = mask blocks
    unsigned long long units;        // mask units: (part, run of tile rows) handled by one tile_mask warp
    unsigned long long words;        // 32-bit words of all compact mask blocks (each block padded to 8 words)
    unsigned long long edge_visits;  // sum over parts of units x column chunks x ring vertices
    unsigned long long cross_lb;     // lower bound of the crossing count: 2 per part row (a closed ring crosses a
                                     // row's centre line an even number of times, at least twice)
    unsigned int nonfinite;          // some part burns a NaN / infinite value (float dtypes)
    unsigned int eq_bg;              // some part burns a value whose bits equal the background's
    unsigned int overflow;           // a count exceeded its cached upper bound (never expected; checked by the host
                                     // whenever it synchronises anyway)
    unsigned int scan_ticket;        // block ticket of the look-back scan
    unsigned int mask_ticket;        // next unassigned mask unit (tile_mask's warps take MASK_UNITS at a time)
    unsigned int pad_;
};

// One (part, tile) block of the inside mask: rows [row_off, row_off + nrows) x words [w_off, w_off + nw) of the
// tile, row-major, at word offset `woff` of the mask array; plus the value its part burns.
struct alignas(16) BlockDesc {
    unsigned long long value_bits;
    uint32_t woff;
    uint32_t geom;  // row_off | nrows << 8 | w_off << 16 | nw << 20
};
__device__ __forceinline__ uint32_t block_nw(uint32_t geom) { return geom >> 20; }
// (Tried: tile_mask flagging blocks whose words are all ones - tiles inside a large polygon - so that tile_apply
// skips their copy.  Config 3: ~15 % of the blocks, tile_apply unchanged at 4.75 / 3.4 ms, tile_mask 2 % slower.)
__device__ __forceinline__ uint32_t pack_geom(uint32_t row_off, uint32_t nrows, uint32_t w_off, uint32_t nw) {
    return row_off | (nrows << 8) | (w_off << 16) | (nw << 20);
}

// where a part's blocks live: block(tr, tc) = first_block + (tr - tr0) * ntc + (tc - tc0)
struct alignas(8) PartTile {
    uint32_t first_block;
    uint32_t tr0, tc0;
    uint32_t ntr, ntc;
    uint32_t r_lo, r_hi;  // rows the part can fill (absolute, clamped to the window)
    uint32_t w_lo, w_hi;  // 32-column words the part can fill (absolute)
    uint32_t pad;
};

// A mask unit = the tile rows [tr, tr + k) of one part, built by one tile_mask warp: as many tile rows as
// the part's rows in them fit the warp's shared-memory toggle mask.  Packed [tr:26 | k:6 | part:32].
__host__ __device__ __forceinline__ uint32_t mask_stride(uint32_t nw) { return nw | 1u; }  // odd: rows spread over banks
__device__ __forceinline__ uint32_t mask_row_capacity(uint32_t n_words) {
    return MASK_SMEM_WORDS / mask_stride(min(n_words, MASK_MAX_WORDS));
}
template <typename F>
__device__ __forceinline__ uint32_t for_each_mask_unit(const KParams& P, uint32_t tile_r, uint32_t tr0, uint32_t ntr,
                                                       uint32_t n_words, uint32_t r_lo, uint32_t r_hi, F&& emit) {
    const uint32_t cap = mask_row_capacity(n_words);
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
// Column chunks of a part's word range [w_lo, w_hi): the first starts at w_lo, the following ones at multiples
// of 16 words from w_lo's tile, so that a tile column (4 words) never straddles two chunks.
__device__ __forceinline__ uint32_t next_chunk(uint32_t a) { return (a & ~3u) + MASK_MAX_WORDS; }
__device__ __forceinline__ uint32_t n_chunks(uint32_t w_lo, uint32_t w_hi) {
    return w_hi > w_lo ? (w_hi - (w_lo & ~3u) + MASK_MAX_WORDS - 1) / MASK_MAX_WORDS : 0u;
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

// rows / words of tile (tr, tc) covered by a part's box; returns the block's size in words, padded to 8
__device__ __forceinline__ uint32_t block_geom(const KParams& P, uint32_t tile_r, uint32_t tr, uint32_t tc, uint32_t r_lo,
                                               uint32_t r_hi, uint32_t w_lo, uint32_t w_hi, uint32_t& geom) {
    const uint32_t t0 = P.win_r0 + tr * tile_r;
    const uint32_t ra = max(r_lo, t0), rb = min(r_hi, t0 + tile_r);
    const uint32_t wa = max(w_lo, tc * 4u), wb = min(w_hi, tc * 4u + 4u);
    geom = pack_geom(ra - t0, rb - ra, wa - tc * 4u, wb - wa);
    return ((rb - ra) * (wb - wa) + 7u) & ~7u;
}

// mode 0: per part the number of blocks, of mask units and of mask words (+ totals into TileCounters);
//         with ignore_band every polygon part counts (upper bounds for the host's plan cache);
// mode 1: with the scanned offsets, fill PartTile and emit the unit list, the block descriptors (part order)
//         and the tile records [tile | block].
static __global__ void tile_bin_kernel(KParams P, TileParams T, const PartInfo* __restrict__ info,
                                const uint8_t* __restrict__ part_kind,
                                const double* __restrict__ xlo, const double* __restrict__ xhi,
                                const double* __restrict__ ylo, const double* __restrict__ yhi,
                                const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                                uint32_t* __restrict__ cnt, /* [3][n_parts]: blocks, units, words */
                                const uint32_t* __restrict__ off, /* [3][n_parts] exclusive prefixes */
                                PartTile* __restrict__ pt, uint64_t* __restrict__ units, uint64_t* __restrict__ recs,
                                BlockDesc* __restrict__ desc, TileCounters* __restrict__ tc, int mode, int ignore_band,
                                int float_bytes, unsigned long long bg_bits, unsigned long long value_mask) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ntr = 0, ntc = 0, n_units = 0, n_words = 0, r_lo = 0, r_hi = 0, c_lo = 0, c_hi = 0, w_lo = 0, w_hi = 0;
    if (p < P.n_parts) {
        const int32_t band = ignore_band ? 0 : info[p].band;
        if (band >= 0 && part_kind[p] == 0 && vend[p] > vbeg[p] + 1 &&
            part_pixel_box(P, xlo[p], xhi[p], ylo[p], yhi[p], r_lo, r_hi, c_lo, c_hi)) {
            const uint32_t tr0 = (r_lo - P.win_r0) / T.tile_r, tr1 = (r_hi - 1 - P.win_r0) / T.tile_r;
            w_lo = c_lo >> 5;
            w_hi = ((c_hi - 1) >> 5) + 1;
            const uint32_t tc0 = w_lo >> 2, tc1 = (w_hi - 1) >> 2;
            ntr = tr1 - tr0 + 1;
            ntc = tc1 - tc0 + 1;
            if (mode == 0) {
                n_units = for_each_mask_unit(P, T.tile_r, tr0, ntr, w_hi - w_lo, r_lo, r_hi, [](uint32_t, uint32_t) {});
                // words: a tile row's blocks hold rows(tr) x words(tc), each padded to 8 words
                for (uint32_t tr = tr0; tr <= tr1; tr++) {
                    const uint32_t t0 = P.win_r0 + tr * T.tile_r;
                    const uint32_t rows = min(r_hi, t0 + T.tile_r) - max(r_lo, t0);
                    if (ntc == 1) {
                        n_words += (rows * (w_hi - w_lo) + 7u) & ~7u;
                    } else {
                        n_words += (rows * ((tc0 + 1) * 4u - w_lo) + 7u) & ~7u;
                        n_words += (rows * (w_hi - tc1 * 4u) + 7u) & ~7u;
                        n_words += (ntc - 2) * ((rows * 4u + 7u) & ~7u);
                    }
                }
            } else {
                const uint32_t o_blk = off[p], o_unit = off[P.n_parts + p];
                uint32_t woff = off[2 * (size_t)P.n_parts + p];
                if ((unsigned long long)o_blk + (unsigned long long)ntr * ntc > T.cap_pairs) {
                    atomicOr(&tc->overflow, 1u);
                } else {
                    PartTile q;
                    q.first_block = o_blk;
                    q.tr0 = tr0;
                    q.tc0 = tc0;
                    q.ntr = ntr;
                    q.ntc = ntc;
                    q.r_lo = r_lo;
                    q.r_hi = r_hi;
                    q.w_lo = w_lo;
                    q.w_hi = w_hi;
                    q.pad = 0;
                    pt[p] = q;
                    uint32_t ou = o_unit;
                    for_each_mask_unit(P, T.tile_r, tr0, ntr, w_hi - w_lo, r_lo, r_hi, [&](uint32_t tr, uint32_t k) {
                        if (ou < T.cap_units) units[ou] = ((uint64_t)tr << 38) | ((uint64_t)k << 32) | p;
                        else atomicOr(&tc->overflow, 1u);
                        ou++;
                    });
                    const unsigned long long vb = info[p].value_bits;
                    uint32_t o = o_blk;
                    for (uint32_t tr = tr0; tr <= tr1; tr++) {
                        for (uint32_t tcol = tc0; tcol <= tc1; tcol++) {
                            const uint64_t tile = ((uint64_t)band * T.n_tr + tr) * T.n_tc + tcol;
                            BlockDesc d;
                            d.value_bits = vb;  // tile_apply reads the value with the block, not by part
                            d.woff = woff;
                            woff += block_geom(P, T.tile_r, tr, tcol, r_lo, r_hi, w_lo, w_hi, d.geom);
                            desc[o] = d;
                            recs[o] = (tile << T.block_bits) | o;  // block index ascends with the part id: burn order
                            o++;
                        }
                    }
                }
            }
        }
    }
    if (mode == 0) {
        if (p < P.n_parts) {
            cnt[p] = ntr * ntc;
            cnt[P.n_parts + p] = n_units;
            cnt[2 * (size_t)P.n_parts + p] = n_words;
            if (!ignore_band && ntr) {
                const unsigned long long vb = info[p].value_bits;
                if (float_bytes) {  // exponent all ones: NaN or infinity (tile_apply's additive mode needs finite values)
                    const bool bad = float_bytes == 4 ? ((vb >> 23) & 0xffu) == 0xffu : ((vb >> 52) & 0x7ffu) == 0x7ffu;
                    if (bad) atomicOr(&tc->nonfinite, 1u);
                }
                if (((vb ^ bg_bits) & value_mask) == 0) atomicOr(&tc->eq_bg, 1u);
            }
        }
        unsigned long long pairs = (unsigned long long)ntr * ntc, un = n_units, words = n_words,
                           visits = (unsigned long long)n_units * n_chunks(w_lo, w_hi) * (ntr ? vend[p] - vbeg[p] : 0u),
                           cross = ntr ? 2ull * (r_hi - r_lo) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pairs += __shfl_down_sync(0xffffffffu, pairs, o);
            un += __shfl_down_sync(0xffffffffu, un, o);
            words += __shfl_down_sync(0xffffffffu, words, o);
            visits += __shfl_down_sync(0xffffffffu, visits, o);
            cross += __shfl_down_sync(0xffffffffu, cross, o);
        }
        // one set of atomics per CTA, not per warp: all of them hit the same five words (C4: 31 k warps)
        __shared__ unsigned long long s_tot[5];
        if (threadIdx.x < 5) s_tot[threadIdx.x] = 0;
        __syncthreads();
        if (lane_id() == 0 && pairs) {
            atomicAdd(&s_tot[0], pairs);
            atomicAdd(&s_tot[1], un);
            atomicAdd(&s_tot[2], words);
            atomicAdd(&s_tot[3], visits);
            atomicAdd(&s_tot[4], cross);
        }
        __syncthreads();
        if (threadIdx.x < 5 && s_tot[threadIdx.x]) {
            unsigned long long* dst = threadIdx.x == 0 ? &tc->pairs : threadIdx.x == 1 ? &tc->units : threadIdx.x == 2 ? &tc->words
                                      : threadIdx.x == 3 ? &tc->edge_visits : &tc->cross_lb;
            atomicAdd(dst, s_tot[threadIdx.x]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// single-launch exclusive scan of K u32 arrays (decoupled look-back)
// ---------------------------------------------------------------------------------------------
// in / out are [K][n]; status holds K 64-bit words per block: flag (2 top bits: 1 = block aggregate, 2 = inclusive
// prefix) | value, zeroed before the launch together with the ticket.  Blocks take their index from a ticket so
// that a block only ever waits for blocks that already run.  Totals must stay below 2^32 (checked by the host
// against its cached upper bounds).
constexpr int LB_THREADS = 256;
constexpr int LB_ITEMS = 8;
constexpr int LB_TILE = LB_THREADS * LB_ITEMS;
constexpr unsigned long long LB_FLAG_A = 1ull << 62, LB_FLAG_P = 2ull << 62, LB_VALUE = (1ull << 62) - 1ull;

template <int K>
static __global__ void __launch_bounds__(LB_THREADS)
scan_lookback_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                     unsigned long long* __restrict__ status, unsigned int* __restrict__ ticket) {
    __shared__ uint32_t s_bid;
    __shared__ uint32_t s_warp[K][LB_THREADS / 32];
    __shared__ uint32_t s_base[K];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t i0 = bid * LB_TILE + tid * LB_ITEMS;
    uint32_t v[K][LB_ITEMS], tsum[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        tsum[k] = 0;
#pragma unroll
        for (int j = 0; j < LB_ITEMS; j++) {
            v[k][j] = i0 + j < n ? in[(size_t)k * n + i0 + j] : 0u;
            tsum[k] += v[k][j];
        }
    }
    // exclusive scan of the thread sums inside the block
    uint32_t texc[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        uint32_t inc = tsum[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += up;
        }
        if (lane == 31) s_warp[k][warp] = inc;
        texc[k] = inc - tsum[k];
    }
    __syncthreads();
    if (warp < K) {  // warp k: block aggregate of array k, then the look-back
        const int k = warp;
        const uint32_t wv = lane < LB_THREADS / 32 ? s_warp[k][lane] : 0u;
        uint32_t inc = wv;
#pragma unroll
        for (int o = 1; o < LB_THREADS / 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += up;
        }
        __syncwarp();
        if (lane < LB_THREADS / 32) s_warp[k][lane] = inc - wv;  // exclusive warp bases
        const uint32_t agg = __shfl_sync(0xffffffffu, inc, LB_THREADS / 32 - 1);
        volatile unsigned long long* st = status + k;
        if (lane == 0) {
            st[(size_t)bid * K] = (bid == 0 ? LB_FLAG_P : LB_FLAG_A) | agg;
            __threadfence();
        }
        unsigned long long prefix = 0;
        if (bid > 0) {
            long long j = (long long)bid - 1 - (long long)lane;  // this lane's predecessor in the current window of 32
            while (true) {
                unsigned long long s = LB_FLAG_P;  // before block 0: an empty inclusive prefix
                if (j >= 0) {
                    do {
                        s = st[(size_t)j * K];
                    } while ((s >> 62) == 0);
                }
                const uint32_t is_p = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                // sum the values of the lanes up to and including the nearest inclusive prefix
                const uint32_t first_p = is_p ? (uint32_t)__ffs(is_p) - 1u : 31u;
                unsigned long long val = lane <= first_p ? (s & LB_VALUE) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
                prefix += __shfl_sync(0xffffffffu, val, 0);
                if (is_p) break;
                j -= 32;
            }
            if (lane == 0) {
                st[(size_t)bid * K] = LB_FLAG_P | ((prefix + agg) & LB_VALUE);
                __threadfence();
            }
        }
        if (lane == 0) s_base[k] = (uint32_t)prefix;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        uint32_t run = s_base[k] + s_warp[k][warp] + texc[k];
#pragma unroll
        for (int j = 0; j < LB_ITEMS; j++) {
            if (i0 + j < n) out[(size_t)k * n + i0 + j] = run;
            run += v[k][j];
        }
    }
}

// One ring edge (pixel-space vertices): its crossing with a row's centre line.
struct TileEdge {
    double x_top, y_top, dxdy;
};
// ... as tile_mask keeps it in shared memory.  Crossing kk of the batch belongs to this edge when
// pre <= kk < pre + cnt; it lies on row (lo + kk - pre), whose centre ordinate row + 0.5 is cyb + kk (exact: small
// integers and halves) and whose index relative to the unit's first row is ib + kk.
struct alignas(16) MaskEdge {
    double x_top, y_top, dxdy, cyb;
    int32_t ib;
    uint32_t pad[3];
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

// After the stable sort of the [tile | block] records: the block descriptors in tile order, so that tile_apply
// streams a tile's descriptors from consecutive memory.  Records beyond the actual count are 0xff.. fillers.
static __global__ void block_pos_kernel(const uint64_t* __restrict__ recs, const TileCounters* __restrict__ tc,
                                 uint32_t block_bits, const BlockDesc* __restrict__ desc,
                                 BlockDesc* __restrict__ desc_sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t)tc->pairs) return;
    const uint32_t blk = (uint32_t)(recs[i] & ((1ull << block_bits) - 1ull));
    const uint4 d = *reinterpret_cast<const uint4*>(desc + blk);
    *reinterpret_cast<uint4*>(desc_sorted + i) = d;
}

// ---------------------------------------------------------------------------------------------
// tile_mask: one warp per mask unit = (part, run of tile rows)
// ---------------------------------------------------------------------------------------------
// A row can only have an odd number of crossings when a non-horizontal edge was skipped for being
// shorter than f64::EPSILON in y (edges.rs:100) while still straddling a pixel centre.  Pixel-centre
// ordinates k+0.5 with k >= 1 are spaced >= EPSILON apart, so that can only happen on raster row 0
// (centre 0.5): only that row's crossing count is tracked, and an odd row 0 drops its largest column like
// chunks_exact(2) drops the unpaired tail (burners.rs:305).  (Non-finite coordinates would break this argument:
// geometry sets holding any are never given to this engine.)
template <int TILE_R>
static __global__ void __launch_bounds__(MASK_WARPS * 32, 8)
tile_mask_kernel(KParams P, TileParams T, const uint64_t* __restrict__ units, TileCounters* tcnt,
                 const PartTile* __restrict__ pt, const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                 const double* __restrict__ wx, const double* __restrict__ wy, const uint32_t* __restrict__ tag,
                 const BlockDesc* __restrict__ desc, uint32_t* __restrict__ masks) {
    __shared__ uint32_t s_mask[MASK_WARPS][MASK_SMEM_WORDS];
    __shared__ MaskEdge s_edge[MASK_WARPS][32];  // the batch's active edges, compacted
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t n_units = min((uint32_t)tcnt->units, T.cap_units);
    // A warp builds MASK_UNITS consecutive units (units are in part order, so are their vertices).  While it works
    // on one it asks the L2 for the next one's part record and vertex range: a unit lives for ~20 us and would
    // otherwise start with three dependent global loads.
    // The grid is persistent (one CTA per resident slot) and the warps draw their units from a ticket counter: no
    // partial last wave, and long parts do not hold up a fixed share of the work (8 GPUs: 0.55 -> see DESIGN.md).
    for (;;) {
    uint32_t unit0 = 0;
    if (lane == 0) unit0 = atomicAdd(&tcnt->mask_ticket, MASK_UNITS);
    unit0 = __shfl_sync(0xffffffffu, unit0, 0);
    if (unit0 >= n_units) break;
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
    const uint32_t lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);

    for (uint32_t wa = q.w_lo; wa < q.w_hi; wa = next_chunk(wa)) {  // one pass per chunk of <= 512 columns
        const uint32_t nw = min(next_chunk(wa), q.w_hi) - wa;
        const uint32_t stride = mask_stride(nw);
        const uint32_t c0 = wa * 32;  // first pixel column of the chunk
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
                MaskEdge* me = &s_edge[warp][__popc(act & lt_mask)];
                // row + 0.5 of crossing kk is (lo + 0.5 - pre) + kk: all terms are small integers or halves, exact
                const double cyb = __dsub_rn(__dadd_rn((double)lo, 0.5), (double)pre);
                *reinterpret_cast<double2*>(&me->x_top) = make_double2(e.x_top, e.y_top);
                *reinterpret_cast<double2*>(&me->dxdy) = make_double2(e.dxdy, cyb);
                me->ib = (int32_t)(lo - row_start) - (int32_t)pre;
            }
            __syncwarp();
            // all lanes share the batch's crossings evenly: crossing kk belongs to the last active edge whose
            // first crossing is <= kk, found from the bit pattern of the first-crossing positions
            // Two chunks of 32 crossings per iteration, evaluated without branches up to the atomic, so that the
            // two dependent f64 chains (shared-memory load -> 5 double ops -> convert) overlap.
            uint32_t cum = 0;
            double kd = (double)lane;  // crossing index of the lane's first chunk as a double (no conversion per crossing)
            auto crossing = [&](uint32_t slot, uint32_t kk, double kkd, uint32_t& word, uint32_t& bit) -> bool {
                const MaskEdge* me = &s_edge[warp][min(slot, 31u)];
                const double2 a = *reinterpret_cast<const double2*>(&me->x_top);  // x_top, y_top
                const double2 b = *reinterpret_cast<const double2*>(&me->dxdy);   // dxdy, cyb
                const uint32_t rrow = (uint32_t)(me->ib + (int32_t)kk);           // row - row_start
                const double cy = __dadd_rn(b.y, kkd);                            // row + 0.5
                // floor(x + 0.5) as usize (burners.rs:310-311); the clamp to ncols is implied by `col < c1` below
                const uint32_t col = __double2uint_rd(__dadd_rn(__dadd_rn(a.x, __dmul_rn(__dsub_rn(cy, a.y), b.x)), 0.5));
                const uint32_t rel = col <= c0 ? 0u : col - c0;  // left of the chunk: parity carry-in at bit 0
                word = rrow * stride + (rel >> 5);
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
                uint32_t wd_a, bt_a, wd_b, bt_b;
                const bool oka = crossing(slot_a, k0 + lane, kd, wd_a, bt_a);
                const bool okb = crossing(slot_b, k0 + 32 + lane, __dadd_rn(kd, 32.0), wd_b, bt_b);
                kd = __dadd_rn(kd, 64.0);
                if (oka) atomicXor(&mask[wd_a], bt_a);
                if (okb) atomicXor(&mask[wd_b], bt_b);
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
        // compact blocks: rows x words of every (tile row of the unit, tile column of the chunk); consecutive lanes
        // write consecutive words
        const uint32_t tca = wa >> 2, tcb = (wa + nw - 1) >> 2;
        for (uint32_t j = 0; j < k_tr; j++) {
            const uint32_t tj = t0 + j * TILE_R;
            const uint32_t blk_row = q.first_block + (tr + j - q.tr0) * q.ntc - q.tc0;
            for (uint32_t tcl = tca; tcl <= tcb; tcl++) {
                const uint2 wg = *reinterpret_cast<const uint2*>(&desc[blk_row + tcl].woff);
                const uint32_t geom = wg.y;
                const uint32_t b_row = geom & 0xffu, b_nr = (geom >> 8) & 0xffu, b_w = (geom >> 16) & 0xfu, b_nw = block_nw(geom);
                const uint32_t inv = b_nw == 1 ? 65536u : b_nw == 2 ? 32768u : b_nw == 3 ? 21846u : 16384u;
                const uint32_t src0 = (tj + b_row - row_start) * stride + (tcl * 4u + b_w - wa);
                uint32_t* dst = masks + wg.x;
                for (uint32_t i = lane; i < b_nr * b_nw; i += 32) {
                    const uint32_t r = (i * inv) >> 16, w = i - r * b_nw;
                    dst[i] = mask[src0 + r * stride + w];
                }
            }
        }
        __syncwarp();
    }
    }  // units of this draw
    }  // draws
}

// ---------------------------------------------------------------------------------------------
// tile_apply: one CTA per 8 tiles of a tile row, each warp owns 8 rows held in REGISTERS, no synchronisation
// between warps
// ---------------------------------------------------------------------------------------------
// Lane 4r + w of a warp OWNS the 32 pixels of (row r of the warp's 8, columns 32w .. 32w+31) of the tile, in
// registers px[0..31] (initialised to the background: geo/raster.rs:23-28).  For every block of the tile, in burn
// order, the lane fetches its own word of the part's inside mask - if the block's rows and words cover it - and
// applies the part's value to its pixels: no shuffle, no shared memory; bit b of the word decides whether pixel b
// takes the value through the reference's pixel-function rule (pixel_functions.rs:56-123).  Warps whose rows the
// block does not reach skip it without a load.
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
// Modes 1 and 3 depend on a property of the burn values that only the device knows (TileCounters): the kernel
// holds both the fast body and the generic one and picks at run time (one uniform branch per CTA).
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

constexpr int APPLY_TILES = 32;  // most consecutive tiles of one tile row handled by one CTA (TileParams::apply_tiles):
                                 // long runs amortise the pipeline start-up (C4: 8 -> 3.61 ms, 16 -> 3.34, 32 -> 3.29)
constexpr int AP_STAGES = 3;               // batches of mask blocks in flight per CTA
constexpr uint32_t AP_STAGE_WORDS = 2048;  // mask words of one batch (a block holds at most 64 x 4 = 256)
constexpr uint32_t AP_STAGE_BLOCKS = 32;   // blocks of one batch (one producer lane each)

// A block as the consumer warps see it: its value, where its words start inside the batch's stage, its geometry.
struct alignas(16) StagedDesc {
    unsigned long long value_bits;
    uint32_t soff;
    uint32_t geom;
};

// ---- mbarrier / bulk-copy (TMA) primitives -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (the TMA engine moves `bytes`, a multiple of 16, and signals `bar` with them)
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// Shared memory of one tile_apply CTA, after the flush staging rows of its consumer warps.
struct ApplyShared {
    uint32_t words[AP_STAGES][AP_STAGE_WORDS];
    StagedDesc desc[AP_STAGES][AP_STAGE_BLOCKS];
    uint64_t full[AP_STAGES], empty[AP_STAGES];
    uint32_t count[AP_STAGES];
};
template <typename N, int TILE_R> __host__ __device__ constexpr size_t apply_flush_bytes() { return (size_t)(TILE_R / 8) * 8 * 4 * (32 * sizeof(N) + 16); }
template <typename N, int TILE_R> __host__ __device__ constexpr size_t apply_smem_bytes() { return apply_flush_bytes<N, TILE_R>() + sizeof(ApplyShared); }

// One CTA = TILE_R / 8 consumer warps + ONE producer warp.  The descriptors of a CTA's tiles are consecutive in
// memory, tile after tile, parts in burn order inside a tile.  The producer walks them in batches of up to 32 blocks
// / 2048 mask words: lane i reads descriptor i of the batch (coalesced), a warp scan places the blocks' compact mask
// words back to back in a shared-memory stage, and every lane issues ONE bulk copy (cp.async.bulk, the TMA engine)
// for its block; the copies signal the stage's `full` mbarrier with their byte counts.  Three stages are in flight,
// recycled through `empty` mbarriers the consumer warps arrive on.  The consumers never touch global memory for
// masks: every block costs them a broadcast descriptor read, a row test and - if the block reaches their 8 rows -
// one shared-memory word per lane.
template <typename N, int FN, int TILE_R, int MODE, bool BGNAN>
__device__ __forceinline__ void tile_apply_body(const KParams& P, const TileParams& T, const uint32_t* __restrict__ tile_start,
                                                const BlockDesc* __restrict__ desc, const uint32_t* __restrict__ masks,
                                                uint64_t bg_bits, N* __restrict__ out, unsigned char* smem_raw) {
    constexpr uint32_t NW = TILE_R / 8;  // consumer warps
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const N bg = value_from_bits<N>(bg_bits);
    ApplyShared& sh = *reinterpret_cast<ApplyShared*>(smem_raw + apply_flush_bytes<N, TILE_R>());

    // grid = (groups of apply_tiles tile columns, tile rows x bands), or 1-D when that does not fit the grid limits
    const uint32_t groups = (T.n_tc + T.apply_tiles - 1) / T.apply_tiles;
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
    const uint32_t tcol0 = tgrp * T.apply_tiles, n_here = min(T.apply_tiles, T.n_tc - tcol0);
    const uint32_t t0 = (band * T.n_tr + trow) * T.n_tc + tcol0;
    const uint32_t tile_row0 = P.win_r0 + trow * TILE_R;
    const uint32_t n_active = min(NW, (P.win_r1 - tile_row0 + 7u) / 8u);  // consumer warps that own raster rows
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < AP_STAGES; k++) {
            mbar_init(&sh.full[k], 1);
            mbar_init(&sh.empty[k], n_active);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint32_t cta_beg = tile_start[t0], cta_end = tile_start[t0 + n_here];  // blocks of this CTA's tiles

    if (warp == NW) {
        // ---- producer --------------------------------------------------------------------------------------
        uint32_t jb = cta_beg;
        for (uint32_t k = 0; jb < cta_end; k++) {
            const uint32_t st = k % AP_STAGES;
            if (k >= (uint32_t)AP_STAGES) mbar_wait(&sh.empty[st], ((k / AP_STAGES) - 1u) & 1u);
            const uint32_t j = jb + lane;
            uint4 d = make_uint4(0, 0, 0, 0);
            if (j < cta_end) d = __ldg(reinterpret_cast<const uint4*>(desc + j));  // value lo, value hi, woff, geom
            const uint32_t size = j < cta_end ? ((((d.w >> 8) & 0xffu) * block_nw(d.w) + 7u) & ~7u) : 0u;
            uint32_t inc = size;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += up;
            }
            const uint32_t fits = __ballot_sync(0xffffffffu, j < cta_end && inc <= AP_STAGE_WORDS);  // a prefix of the lanes
            const uint32_t cnt = __popc(fits);                                                        // >= 8
            const uint32_t total = __shfl_sync(0xffffffffu, inc, cnt - 1);
            if (lane < cnt) {
                StagedDesc sd;
                sd.value_bits = (unsigned long long)d.x | ((unsigned long long)d.y << 32);
                sd.soff = inc - size;
                sd.geom = d.w;
                sh.desc[st][lane] = sd;
            }
            if (lane == 0) sh.count[st] = cnt;
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&sh.full[st], total * 4u);
            __syncwarp();
            if (lane < cnt) bulk_copy_g2s(&sh.words[st][inc - size], masks + d.z, size * 4u, &sh.full[st]);
            jb += cnt;
        }
        return;
    }

    // ---- consumers ---------------------------------------------------------------------------------------------
    const uint32_t wr0 = warp * 8;  // first of this warp's rows inside the tile
    const uint32_t r0 = tile_row0 + wr0;
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
    const uint32_t my_r = wr0 + (lane >> 2), my_w = lane & 3u;
    uint32_t j = cta_beg, batch_beg = cta_beg, batch_end = cta_beg, n_batches = 0, st = 0;
    for (uint32_t ti = 0; ti < n_here; ti++) {
    const uint32_t end = tile_start[t0 + ti + 1];
    const uint32_t c0 = (tcol0 + ti) * TILE_C;
    N px[32];  // pixel b = (row r0 + (lane >> 2), column c0 + 32 * (lane & 3) + b)
    uint32_t touched = 0;
#pragma unroll
    for (int b = 0; b < 32; b++) px[b] = ident;

    while (j < end) {
        if (j >= batch_end) {  // next batch: hand the finished stage back, wait for the next one to land
            if (n_batches) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.empty[(n_batches - 1u) % AP_STAGES]);
            }
            st = n_batches % AP_STAGES;
            mbar_wait(&sh.full[st], (n_batches / AP_STAGES) & 1u);
            batch_beg = batch_end;
            batch_end += sh.count[st];
            n_batches++;
        }
        const uint32_t lim = min(end, batch_end);
        for (; j < lim; j++) {
            const StagedDesc d = sh.desc[st][j - batch_beg];
            const uint32_t b_row = d.geom & 0xffu, b_nr = (d.geom >> 8) & 0xffu;
            if (!(b_row < wr0 + 8u && b_row + b_nr > wr0)) continue;  // the block does not reach these 8 rows
            const uint32_t rr = my_r - b_row, ww = my_w - ((d.geom >> 16) & 0xfu), nw = block_nw(d.geom);
            const uint32_t mu = (rr < b_nr && ww < nw) ? sh.words[st][d.soff + rr * nw + ww] : 0u;
            const N v = value_from_bits<N>(d.value_bits);
            if (MODE == 3) {
                // `first` = `last` under the mask of the pixels nobody has written yet
                if (FN == RZ_FIRST) {
                    // The mask goes through a (no-op) shuffle so that ptxas keeps it in ONE register and expands it with
                    // R2P: it otherwise folds `mu & ~touched & bit` into a 3-input LOP3 per pixel - 64 instead of 36
                    // instructions per word (config 3 `first`: 4.75 ms against 3.4 ms for `last`).
                    const uint32_t fresh = __shfl_sync(0xffffffffu, mu & ~touched, lane);
                    apply_part_word<N, RZ_LAST, 0, BGNAN>(px, fresh, v, bg);
                } else {
                    apply_part_word<N, FN, 3, BGNAN>(px, mu, v, bg);
                }
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
    }  // tiles of this CTA
}

// FAST_MODE runs when the burn values allow it (see MODE above), else the generic MODE 0 body.
template <typename N, int FN, int TILE_R, int FAST_MODE, bool BGNAN>
static __global__ void __launch_bounds__(TILE_R * 4 + 32, TILE_R == 64 ? 3 : 4)
tile_apply_kernel(KParams P, TileParams T, const uint32_t* __restrict__ tile_start, const BlockDesc* __restrict__ desc,
                  const uint32_t* __restrict__ masks, const TileCounters* __restrict__ tcnt, uint64_t bg_bits,
                  N* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (FAST_MODE == 1 || FAST_MODE == 3) {
        const bool slow = std::is_floating_point<N>::value ? tcnt->nonfinite != 0 : tcnt->eq_bg != 0;
        if (slow) {
            tile_apply_body<N, FN, TILE_R, 0, BGNAN>(P, T, tile_start, desc, masks, bg_bits, out, smem_raw);
            return;
        }
    }
    tile_apply_body<N, FN, TILE_R, FAST_MODE, BGNAN>(P, T, tile_start, desc, masks, bg_bits, out, smem_raw);
}

}  // namespace rz
