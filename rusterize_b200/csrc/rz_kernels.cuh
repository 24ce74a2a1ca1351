// rz_kernels.cuh — device code of the burn path (sm_100a).
//
// Pipeline (one "window" = a contiguous range of raster rows of every band):
//   part_prepare      per part: band, value, column-tile range                         (rasterize.rs:162-196)
//   *_count / *_emit  edge setup: polygon scanline crossings, Bresenham line pixels,
//                     point cells -> 64-bit records  [task | part | col]                (edges.rs, burners.rs)
//   radix sort        LSD, 8-bit digits, keys only                                      (burners.rs:276,302)
//   task_index        record range of every (band,row,column-tile) task
//   fill              one warp per task: row tile in shared memory, records replayed in
//                     part order with the reference's pixel-function rules, tile flushed
//                     once with coalesced stores                                        (pixel_functions.rs)
//
// f64 discipline: every geometry operation uses the explicit round-to-nearest intrinsics
// (__dadd_rn/__dsub_rn/__dmul_rn/__ddiv_rn), which the compiler never contracts into FMA, because
// the reference computes x0 + (cy - y0) * dxdy with two roundings (edges.rs:50-55).
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/rz_b200.h"

namespace rz {

struct KParams {
    double xmin, ymax, xres, yres;  // world -> pixel
    double inv_xres, inv_yres;      // 1/res when res is a power of two (then d * inv == d / res bit for bit), else 0
    double nrows_f, ncols_f;
    uint32_t nrows, ncols;          // full raster
    uint32_t win_r0, win_r1;        // rows of this window (absolute)
    uint32_t tile_w, tile_shift;    // column-tile width (power of two) and its log2
    uint32_t n_tiles;               // ceil(ncols / tile_w)
    uint32_t col_bits, part_bits;   // key layout: [task | part | col]
    uint32_t part_shift, task_shift;
    uint32_t n_bands;
    uint32_t dedup_lines;           // xres != yres (burn_geometry.rs:179,202)
    uint32_t n_parts;
};

// Per-part data resolved once per call.
struct PartInfo {
    uint64_t value_bits;  // field value of the owning geometry, in the output dtype
    int32_t band;         // -1: skipped (null field)
    uint16_t t_lo, t_hi;  // column tiles a polygon part can touch (t_lo > t_hi: none)
};

struct Counters {
    unsigned long long records;     // total records counted (count pass)
    unsigned long long crossings;   // polygon crossings without tile replication
    unsigned long long cursor;      // emit allocation cursor (line / point records; polygons use scanned block bases)
    unsigned long long poly_records; // records of polygon parts (they occupy the front of the buffer)
    unsigned int bad_line;          // a line segment left the supported ±2^29 pixel domain
    unsigned int pad;
};

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
// (X - xmin) / xres, (ymax - Y) / yres (edges.rs:81-82, 94-97).  Dividing by a power of two is an exact scaling,
// so is multiplying by its reciprocal: both give the correctly rounded value of the same real number, and the
// 14-instruction f64 divide becomes one multiply for resolutions such as 1, 0.5 or 0.25 (warp-uniform branch).
__device__ __forceinline__ double px_x(const KParams& P, double X) {
    const double d = __dsub_rn(X, P.xmin);
    return P.inv_xres != 0.0 ? __dmul_rn(d, P.inv_xres) : __ddiv_rn(d, P.xres);
}
__device__ __forceinline__ double px_y(const KParams& P, double Y) {
    const double d = __dsub_rn(P.ymax, Y);
    return P.inv_yres != 0.0 ? __dmul_rn(d, P.inv_yres) : __ddiv_rn(d, P.yres);
}

// Rust `f64 as usize` followed by min(., lim): NaN / negatives -> 0, saturating.
__device__ __forceinline__ uint32_t sat_u32(double v, uint32_t lim) {
    if (!(v > 0.0)) return 0u;
    if (v >= (double)lim) return lim;
    return (uint32_t)v;
}
// Rust `f64 as isize` restricted to ±2^40 (callers flag anything beyond ±2^29 as unsupported).
__device__ __forceinline__ long long sat_i64(double v) {
    if (!(v == v)) return 0;
    if (v <= -1099511627776.0) return -1099511627776LL;
    if (v >= 1099511627776.0) return 1099511627776LL;
    return (long long)v;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---------------------------------------------------------------------------------------------
// part_prepare
// ---------------------------------------------------------------------------------------------
// One thread per part.  Resolves the burn value (field[i] or the scalar), the band and, for polygon
// parts, the range of column tiles whose pixels the part can fill.
static __global__ void part_prepare_kernel(KParams P, const uint8_t* __restrict__ part_kind,
                                    const uint32_t* __restrict__ part_geom, const double* __restrict__ part_xlo,
                                    const double* __restrict__ part_xhi, const uint8_t* __restrict__ field,
                                    uint32_t itemsize, int field_is_scalar, const uint8_t* __restrict__ field_valid,
                                    const int32_t* __restrict__ band_of_geom, PartInfo* __restrict__ info) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_parts) return;
    uint32_t g = part_geom[p];
    PartInfo pi;
    pi.band = band_of_geom ? band_of_geom[g] : 0;
    if (field_valid && !field_valid[g]) pi.band = -1;
    uint64_t v = 0;
    const uint8_t* src = field + (field_is_scalar ? 0 : (size_t)g * itemsize);
    for (uint32_t k = 0; k < itemsize; k++) v |= (uint64_t)src[k] << (8 * k);
    pi.value_bits = v;
    pi.t_lo = 1;
    pi.t_hi = 0;
    if (part_kind[p] == 0) {
        // Crossing columns lie within one pixel of the columns of the part's x-extent (x on an edge
        // is an interpolation between its end points, up to rounding).
        double fl = floor(__dadd_rn(px_x(P, part_xlo[p]), 0.5)), fh = floor(__dadd_rn(px_x(P, part_xhi[p]), 0.5));
        uint32_t cl = sat_u32(fl, P.ncols), ch = sat_u32(fh, P.ncols);
        cl = cl > 0 ? cl - 1 : 0;
        ch = ch < P.ncols ? ch + 1 : P.ncols;
        if (ch > cl && cl < P.ncols) {  // fillable pixels [cl, ch)
            pi.t_lo = (uint16_t)(cl >> P.tile_shift);
            pi.t_hi = (uint16_t)((ch - 1) >> P.tile_shift);
        }
    }
    info[p] = pi;
}

// ---------------------------------------------------------------------------------------------
// small device -> host readbacks without a copy engine
// ---------------------------------------------------------------------------------------------
// Counters the host needs in the middle of a call (record counts, cost-model sums) are stored by this kernel
// straight into page-locked, mapped host memory.  A cudaMemcpy of a few bytes would queue on the copy engine
// behind the 2 GiB raster window that is on its way to the host, and stall the next window for its duration.
static __global__ void readback_kernel(const unsigned long long* __restrict__ src, volatile unsigned long long* dst, uint32_t n) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// vertex tags rebuilt on the device (the host's tag[] is not uploaded)
// ---------------------------------------------------------------------------------------------
// tag = part id | 0x80000000 on the last vertex of a ring / line string | 0x40000000 on every vertex of a closed
// line string (rz::Flattener::end_seq).  One warp per part writes the part id over the part's vertex range...
static __global__ void tag_parts_kernel(uint32_t n_parts, uint8_t kind, const uint8_t* __restrict__ part_kind,
                                 const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                                 uint32_t* __restrict__ tag) {
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_parts || part_kind[p] != kind) return;
    for (uint32_t i = vbeg[p] + lane_id(); i < vend[p]; i += 32) tag[i] = p;
}
// ... then one thread per sequence sets the flags (sequences are contiguous in their pool, in order)
static __global__ void tag_seqs_kernel(uint32_t n_seq, const uint32_t* __restrict__ seq_end, const uint8_t* __restrict__ closed,
                                uint32_t* __restrict__ tag) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seq) return;
    const uint32_t e = seq_end[s];
    if (closed[s])
        for (uint32_t i = s ? seq_end[s - 1] + 1 : 0u; i <= e; i++) tag[i] |= 0x40000000u;
    tag[e] |= 0x80000000u;
}

// ---------------------------------------------------------------------------------------------
// polygon edge setup — rust/src/geo/edges.rs:27-55, 90-110; burners.rs:279-315
// ---------------------------------------------------------------------------------------------
struct PolyEdgeRec {
    double x_top, y_top, dxdy;
    uint32_t row_lo;   // first window row the edge is active on
    uint32_t n_rows;   // rows in [row_lo, row_lo+n_rows) are active
    uint32_t part;
    int32_t band;
    uint32_t t_lo, n_t;
};

__device__ __forceinline__ bool poly_edge_setup(const KParams& P, const double* __restrict__ x,
                                                const double* __restrict__ y, const uint32_t* __restrict__ tag,
                                                const PartInfo* __restrict__ info, uint32_t i, uint32_t n,
                                                PolyEdgeRec& e) {
    e.n_rows = 0;
    e.n_t = 0;
    if (i >= n) return false;
    uint32_t t = tag[i];
    if (t & 0x80000000u) return false;  // last vertex of its ring
    e.part = t & 0x3fffffffu;
    PartInfo pi = info[e.part];
    if (pi.band < 0 || pi.t_lo > pi.t_hi) return false;
    e.band = pi.band;
    e.t_lo = pi.t_lo;
    e.n_t = (uint32_t)pi.t_hi - pi.t_lo + 1;
    double x0 = px_x(P, x[i]), y0 = px_y(P, y[i]);
    double x1 = px_x(P, x[i + 1]), y1 = px_y(P, y[i + 1]);
    if (!(fabs(__dsub_rn(y0, y1)) >= DBL_EPSILON)) return false;  // skip horizontal (edges.rs:100)
    double min_y = fmin(y0, y1), max_y = fmax(y0, y1);
    if (!(min_y < P.nrows_f && max_y >= 0.0)) return false;  // edges.rs:105
    double x_bot, y_bot;
    if (y0 < y1) { e.x_top = x0; e.y_top = y0; x_bot = x1; y_bot = y1; }
    else { e.x_top = x1; e.y_top = y1; x_bot = x0; y_bot = y0; }
    uint32_t ystart = sat_u32(ceil(__dsub_rn(e.y_top, 0.5)), P.nrows);  // edges.rs:32-33 (+ rows < nrows)
    uint32_t yend = sat_u32(ceil(__dsub_rn(y_bot, 0.5)), P.nrows);
    e.dxdy = __ddiv_rn(__dsub_rn(x_bot, e.x_top), __dsub_rn(y_bot, e.y_top));
    uint32_t lo = max(ystart, P.win_r0), hi = min(yend, P.win_r1);
    if (hi <= lo) return false;
    e.row_lo = lo;
    e.n_rows = hi - lo;
    return true;
}

// key of the crossing of edge e with row `row`, replicated into column tile `tile`
__device__ __forceinline__ uint64_t poly_key(const KParams& P, const PolyEdgeRec& e, uint32_t row, uint32_t tile) {
    double cy = __dadd_rn((double)row, 0.5);
    double xi = __dadd_rn(e.x_top, __dmul_rn(__dsub_rn(cy, e.y_top), e.dxdy));  // edges.rs:50-55
    uint32_t col = sat_u32(floor(__dadd_rn(xi, 0.5)), P.ncols);                 // burners.rs:310-311
    uint32_t ts = tile << P.tile_shift;
    uint32_t te = min(ts + P.tile_w, P.ncols);
    uint32_t rel = min(max(col, ts), te) - ts;
    uint64_t task = ((uint64_t)e.band * (P.win_r1 - P.win_r0) + (row - P.win_r0)) * P.n_tiles + tile;
    return (task << P.task_shift) | ((uint64_t)e.part << P.part_shift) | rel;
}

constexpr int SETUP_THREADS = 256;
constexpr int SETUP_ITEMS = 4;      // line kernels: vertices per thread
constexpr uint32_t LONG_EDGE = 64;  // line kernels: records above which a warp cooperates on one segment

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp, uint32_t* total) {
    uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (blockDim.x >> 5) ? smem_warp[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= (uint32_t)o) winc += t;
        }
        smem_warp[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem_warp[32] = winc;
    }
    __syncthreads();
    *total = smem_warp[32];
    return smem_warp[warp] + inc - v;
}

// Count pass: one thread per ring vertex.  Besides the global totals it stores the number of
// records of every block so that the emit pass can place blocks in vertex order (= part order),
// which lets a polygon-only job sort on the task bits alone.
static __global__ void __launch_bounds__(SETUP_THREADS)
poly_count_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                  const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                  uint32_t* __restrict__ block_total, Counters* __restrict__ ctr) {
    __shared__ uint32_t s_rec[SETUP_THREADS / 32], s_crs[SETUP_THREADS / 32];
    PolyEdgeRec e;
    uint32_t rec = 0, crs = 0;
    if (poly_edge_setup(P, x, y, tag, info, blockIdx.x * SETUP_THREADS + threadIdx.x, n, e)) {
        rec = e.n_rows * e.n_t;
        crs = e.n_rows;
    }
    rec = __reduce_add_sync(0xffffffffu, rec);
    crs = __reduce_add_sync(0xffffffffu, crs);
    if (lane_id() == 0) { s_rec[threadIdx.x >> 5] = rec; s_crs[threadIdx.x >> 5] = crs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SETUP_THREADS / 32; w++) { rec += s_rec[w]; crs += s_crs[w]; }
        block_total[blockIdx.x] = rec;
        if (rec) {
            atomicAdd(&ctr->records, (unsigned long long)rec);
            atomicAdd(&ctr->poly_records, (unsigned long long)rec);
        }
        if (crs) atomicAdd(&ctr->crossings, (unsigned long long)crs);
    }
}

// In-place exclusive scan of the per-block totals (single block; n is a few hundred thousand).
static __global__ void __launch_bounds__(1024) scan_u32_kernel(uint32_t* __restrict__ v, uint32_t n) {
    __shared__ uint32_t s_warp[33];
    uint32_t carry = 0;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 < n; b0 += 1024) {
        uint32_t i = b0 + threadIdx.x;
        uint32_t val = i < n ? v[i] : 0, inc = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (uint32_t)o) winc += t;
            }
            s_warp[lane] = winc - w;
            if (lane == 31) s_warp[32] = winc;
        }
        __syncthreads();
        if (i < n) v[i] = carry + s_warp[warp] + inc - val;
        carry += s_warp[32];
        __syncthreads();
    }
}

// Emit pass: blocks write at their scanned base, warps flatten the records of their 32 edges so
// that consecutive lanes store consecutive records (coalesced 256-byte stores).
static __global__ void __launch_bounds__(SETUP_THREADS)
poly_emit_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                 const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                 const uint32_t* __restrict__ block_base, uint64_t base_offset, uint64_t* __restrict__ keys) {
    __shared__ uint32_t s_wtot[SETUP_THREADS / 32];
    __shared__ double s_xtop[SETUP_THREADS], s_ytop[SETUP_THREADS], s_dxdy[SETUP_THREADS];
    __shared__ uint32_t s_pre[SETUP_THREADS], s_rowlo[SETUP_THREADS], s_part[SETUP_THREADS];
    __shared__ uint32_t s_tl[SETUP_THREADS];  // t_lo | n_t << 16
    __shared__ int32_t s_band[SETUP_THREADS];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, w0 = warp * 32;
    PolyEdgeRec e;
    poly_edge_setup(P, x, y, tag, info, blockIdx.x * SETUP_THREADS + threadIdx.x, n, e);
    const uint32_t cnt = e.n_rows * e.n_t;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 0) s_wtot[warp] = wtot;
    s_pre[threadIdx.x] = inc - cnt;
    s_xtop[threadIdx.x] = e.x_top;
    s_ytop[threadIdx.x] = e.y_top;
    s_dxdy[threadIdx.x] = e.dxdy;
    s_rowlo[threadIdx.x] = e.row_lo;
    s_part[threadIdx.x] = e.part;
    s_tl[threadIdx.x] = e.t_lo | (e.n_t << 16);
    s_band[threadIdx.x] = e.band;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < warp; w++) wbase += s_wtot[w];
    uint64_t* out = keys + base_offset + block_base[blockIdx.x] + wbase;
    for (uint32_t t = lane; t < wtot; t += 32) {
        // last edge of this warp whose first record is <= t
        uint32_t lo = 0, hi = 32;
#pragma unroll
        for (int it = 0; it < 5; it++) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_pre[w0 + mid] <= t) lo = mid;
            else hi = mid;
        }
        const uint32_t j = w0 + lo;
        const uint32_t local = t - s_pre[j];
        const uint32_t tl = s_tl[j], n_t = tl >> 16;
        uint32_t row_off = local, tt = 0;
        if (n_t > 1) { row_off = local / n_t; tt = local - row_off * n_t; }
        PolyEdgeRec b;
        b.x_top = s_xtop[j];
        b.y_top = s_ytop[j];
        b.dxdy = s_dxdy[j];
        b.part = s_part[j];
        b.band = s_band[j];
        out[t] = poly_key(P, b, s_rowlo[j] + row_off, (tl & 0xffffu) + tt);
    }
}

// ---------------------------------------------------------------------------------------------
// line segments — rust/src/geo/edges.rs:112-134; burners.rs:35-92
// ---------------------------------------------------------------------------------------------
// The reference walks an integer Bresenham loop.  Iteration k (k = 0 .. L-1, L = max(dx,|dy|))
// visits   major = m0 + s_major*k,   minor = n0 + s_minor * floor((2*d_minor*k + d_major) / (2*d_major))
// (derived from the error recurrence at burners.rs:60-84; tests/test_bresenham.py checks the closed
// form against the loop exhaustively).  That lets us clip to the window in O(1) and emit any k.
struct LineRec {
    long long ix0, iy0;
    long long dmaj, dmin;  // |delta| along major / minor axis
    int sx, sy;
    int xmajor;
    uint32_t k_lo, n;      // iterations [k_lo, k_lo+n) are inside the window
    uint32_t part;
    int32_t band;
};

constexpr long long LINE_DOMAIN = 1LL << 29;

__device__ __forceinline__ long long ceil_div(long long a, long long b) {  // b > 0
    long long q = a / b, r = a % b;
    return q + ((r != 0) && (r > 0));
}
__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
    long long q = a / b, r = a % b;
    return q - ((r != 0) && (r < 0));
}

// pixel visited by iteration k
__device__ __forceinline__ void line_pixel(const LineRec& l, long long k, long long& px, long long& py) {
    long long q = l.dmaj > 0 ? (2 * l.dmin * k + l.dmaj) / (2 * l.dmaj) : 0;
    if (l.xmajor) { px = l.ix0 + l.sx * k; py = l.iy0 + l.sy * q; }
    else { py = l.iy0 + l.sy * k; px = l.ix0 + l.sx * q; }
}

// intersect [klo,khi] with { k : lo <= c0 + s*k < hi }
__device__ __forceinline__ void clip_linear(long long c0, int s, long long lo, long long hi, long long& klo, long long& khi) {
    if (s > 0) { klo = max(klo, lo - c0); khi = min(khi, hi - 1 - c0); }
    else { klo = max(klo, c0 - hi + 1); khi = min(khi, c0 - lo); }
}
// intersect [klo,khi] with { k : lo <= c0 + s*q(k) < hi }, q(k) = floor((2*dmin*k + dmaj)/(2*dmaj))
__device__ __forceinline__ void clip_stepped(long long c0, int s, long long lo, long long hi, long long dmin, long long dmaj,
                                             long long& klo, long long& khi) {
    long long qa, qb;  // need qa <= q <= qb
    if (s > 0) { qa = lo - c0; qb = hi - 1 - c0; }
    else { qa = c0 - hi + 1; qb = c0 - lo; }
    if (dmin == 0) {  // q == 0 for every k
        if (qa > 0 || qb < 0) khi = klo - 1;
        return;
    }
    // q >= qa  <=>  k >= ceil((2*dmaj*qa - dmaj) / (2*dmin));   q <= qb  <=>  k <= ceil((2*dmaj*(qb+1) - dmaj)/(2*dmin)) - 1
    if (qa > 0) klo = max(klo, ceil_div(2 * dmaj * qa - dmaj, 2 * dmin));
    if (qb < 0) { khi = klo - 1; return; }
    if (qb < dmin) khi = min(khi, ceil_div(2 * dmaj * (qb + 1) - dmaj, 2 * dmin) - 1);
}

__device__ __forceinline__ bool line_setup(const KParams& P, const double* __restrict__ x, const double* __restrict__ y,
                                           const uint32_t* __restrict__ tag, const PartInfo* __restrict__ info,
                                           uint32_t i, uint32_t n, LineRec& l, bool* kept, Counters* ctr) {
    l.n = 0;
    *kept = false;
    if (i >= n) return false;
    uint32_t t = tag[i];
    if (t & 0x80000000u) return false;
    l.part = t & 0x3fffffffu;
    PartInfo pi = info[l.part];
    if (pi.band < 0) return false;
    l.band = pi.band;
    double x0 = px_x(P, x[i]), y0 = px_y(P, y[i]);
    double x1 = px_x(P, x[i + 1]), y1 = px_y(P, y[i + 1]);
    double min_x = fmin(x0, x1), max_x = fmax(x0, x1), min_y = fmin(y0, y1), max_y = fmax(y0, y1);
    if (!(min_x < P.ncols_f && max_x >= 0.0 && min_y < P.nrows_f && max_y >= 0.0)) return false;  // edges.rs:130
    *kept = true;
    long long ix0 = sat_i64(floor(x0)), ix1 = sat_i64(floor(x1));
    long long iy0 = sat_i64(floor(y0)), iy1 = sat_i64(floor(y1));
    if (llabs(ix0) > LINE_DOMAIN || llabs(ix1) > LINE_DOMAIN || llabs(iy0) > LINE_DOMAIN || llabs(iy1) > LINE_DOMAIN) {
        atomicOr(&ctr->bad_line, 1u);
        return false;
    }
    long long dx = llabs(ix1 - ix0), dy = llabs(iy1 - iy0);
    l.ix0 = ix0;
    l.iy0 = iy0;
    l.sx = ix0 < ix1 ? 1 : -1;
    l.sy = iy0 < iy1 ? 1 : -1;
    l.xmajor = dx >= dy;
    l.dmaj = l.xmajor ? dx : dy;
    l.dmin = l.xmajor ? dy : dx;
    long long klo = 0, khi = l.dmaj - 1;  // the segment's last pixel is not written by the loop
    if (l.xmajor) {
        clip_linear(ix0, l.sx, 0, (long long)P.ncols, klo, khi);
        clip_stepped(iy0, l.sy, (long long)P.win_r0, (long long)P.win_r1, l.dmin, l.dmaj, klo, khi);
    } else {
        clip_linear(iy0, l.sy, (long long)P.win_r0, (long long)P.win_r1, klo, khi);
        clip_stepped(ix0, l.sx, 0, (long long)P.ncols, l.dmin, l.dmaj, klo, khi);
    }
    if (khi < klo) return true;
    l.k_lo = (uint32_t)klo;
    l.n = (uint32_t)(khi - klo + 1);
    return true;
}

__device__ __forceinline__ uint64_t pixel_key(const KParams& P, int32_t band, uint32_t part, uint32_t row, uint32_t col) {
    uint32_t tile = col >> P.tile_shift;
    uint64_t task = ((uint64_t)band * (P.win_r1 - P.win_r0) + (row - P.win_r0)) * P.n_tiles + tile;
    return (task << P.task_shift) | ((uint64_t)part << P.part_shift) | (col - (tile << P.tile_shift));
}

__device__ __forceinline__ uint64_t line_key(const KParams& P, const LineRec& l, uint32_t k) {
    long long px, py;
    line_pixel(l, (long long)k, px, py);
    return pixel_key(P, l.band, l.part, (uint32_t)py, (uint32_t)px);
}

// count pass; also records, per line part, the highest kept segment (the one that may write its end pixel)
static __global__ void __launch_bounds__(SETUP_THREADS)
line_count_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                  const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                  uint32_t* __restrict__ last_kept, Counters* __restrict__ ctr) {
    unsigned long long rec = 0;
    uint32_t base = blockIdx.x * (SETUP_THREADS * SETUP_ITEMS);
#pragma unroll
    for (int r = 0; r < SETUP_ITEMS; r++) {
        LineRec l;
        bool kept;
        uint32_t i = base + r * SETUP_THREADS + threadIdx.x;
        line_setup(P, x, y, tag, info, i, n, l, &kept, ctr);
        rec += l.n;
        if (kept) atomicMax(&last_kept[l.part], i + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rec += __shfl_down_sync(0xffffffffu, rec, o);
    if (lane_id() == 0 && rec) atomicAdd(&ctr->records, rec);
}

static __global__ void __launch_bounds__(SETUP_THREADS)
line_emit_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                 const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                 Counters* __restrict__ ctr, uint64_t* __restrict__ keys) {
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_base;
    uint32_t base = blockIdx.x * (SETUP_THREADS * SETUP_ITEMS);
    LineRec l[SETUP_ITEMS];
    uint32_t mine = 0;
#pragma unroll
    for (int r = 0; r < SETUP_ITEMS; r++) {
        bool kept;
        line_setup(P, x, y, tag, info, base + r * SETUP_THREADS + threadIdx.x, n, l[r], &kept, ctr);
        mine += l[r].n;
    }
    uint32_t total;
    uint32_t off = block_exclusive_scan(mine, s_warp, &total);
    if (total == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(&ctr->cursor, (unsigned long long)total);
    __syncthreads();
    uint64_t* out = keys + s_base + off;
    uint32_t lane = lane_id();
#pragma unroll
    for (int r = 0; r < SETUP_ITEMS; r++) {
        bool is_long = l[r].n > LONG_EDGE;
        if (!is_long)
            for (uint32_t k = 0; k < l[r].n; k++) out[k] = line_key(P, l[r], l[r].k_lo + k);
        uint32_t m = __ballot_sync(0xffffffffu, is_long);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            LineRec b;
            b.ix0 = __shfl_sync(0xffffffffu, l[r].ix0, src);
            b.iy0 = __shfl_sync(0xffffffffu, l[r].iy0, src);
            b.dmaj = __shfl_sync(0xffffffffu, l[r].dmaj, src);
            b.dmin = __shfl_sync(0xffffffffu, l[r].dmin, src);
            b.sx = __shfl_sync(0xffffffffu, l[r].sx, src);
            b.sy = __shfl_sync(0xffffffffu, l[r].sy, src);
            b.xmajor = __shfl_sync(0xffffffffu, l[r].xmajor, src);
            b.k_lo = __shfl_sync(0xffffffffu, l[r].k_lo, src);
            b.n = __shfl_sync(0xffffffffu, l[r].n, src);
            b.part = __shfl_sync(0xffffffffu, l[r].part, src);
            b.band = __shfl_sync(0xffffffffu, l[r].band, src);
            unsigned long long p = (unsigned long long)(uintptr_t)out;
            uint64_t* bout = (uint64_t*)(uintptr_t)__shfl_sync(0xffffffffu, p, src);
            for (uint32_t k = lane; k < b.n; k += 32) bout[k] = line_key(P, b, b.k_lo + k);
        }
        out += l[r].n;
    }
}

// The end pixel of the last kept segment of a pooled line part is written iff that segment's own
// line string is not closed (burners.rs:87-89).  One thread per part; mode 0 counts, mode 1 emits.
static __global__ void line_final_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                  const uint32_t* __restrict__ tag, const uint8_t* __restrict__ part_kind,
                                  const PartInfo* __restrict__ info, const uint32_t* __restrict__ last_kept,
                                  Counters* __restrict__ ctr, uint64_t* __restrict__ keys, int mode) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_parts || part_kind[p] != 1) return;
    uint32_t lk = last_kept[p];
    if (lk == 0) return;
    uint32_t i = lk - 1;
    if (tag[i] & 0x40000000u) return;  // closed line string
    PartInfo pi = info[p];
    if (pi.band < 0) return;
    long long ix1 = sat_i64(floor(px_x(P, x[i + 1]))), iy1 = sat_i64(floor(px_y(P, y[i + 1])));
    if (ix1 < 0 || ix1 >= (long long)P.ncols || iy1 < (long long)P.win_r0 || iy1 >= (long long)P.win_r1) return;
    if (mode == 0) atomicAdd(&ctr->records, 1ull);
    else keys[atomicAdd(&ctr->cursor, 1ull)] = pixel_key(P, pi.band, p, (uint32_t)iy1, (uint32_t)ix1);
}

// ---------------------------------------------------------------------------------------------
// points — rust/src/geo/edges.rs:79-88; burners.rs:250-258
// ---------------------------------------------------------------------------------------------
static __global__ void point_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                             const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                             Counters* __restrict__ ctr, uint64_t* __restrict__ keys, int mode) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    uint64_t key = 0;
    if (i < n) {
        uint32_t part = tag[i] & 0x3fffffffu;
        PartInfo pi = info[part];
        double px = px_x(P, x[i]), py = px_y(P, y[i]);
        if (pi.band >= 0 && px >= 0.0 && px < P.ncols_f && py >= 0.0 && py < P.nrows_f) {
            uint32_t col = (uint32_t)px, row = (uint32_t)py;  // `as usize` of an in-range value truncates
            if (row >= P.win_r0 && row < P.win_r1) {
                hit = true;
                if (mode) key = pixel_key(P, pi.band, part, row, col);
            }
        }
    }
    uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (m == 0) return;
    uint32_t lane = lane_id();
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(mode ? &ctr->cursor : &ctr->records, (unsigned long long)__popc(m));
    if (mode) {
        base = __shfl_sync(0xffffffffu, base, 0);
        if (hit) keys[base + __popc(m & ((1u << lane) - 1))] = key;
    }
}

// ---------------------------------------------------------------------------------------------
// all_touched line walk — rust/src/rasterization/burners.rs:94-247 (GDAL-derived)
// ---------------------------------------------------------------------------------------------
// A data-dependent f64 walk per segment; one thread walks one segment and calls f(row, col) for every
// in-raster pixel, in the reference's order.  Used for line parts and, with all_touched, for every
// polygon ring (burn_geometry.rs:89-106, 225-238).
template <typename F>
__device__ __forceinline__ void all_touched_walk(const KParams& P, double x, double y, double xe, double ye, F f) {
    const double EPS_INTERSECT = 1e-4, TOL = 1e-9;
    const long long nrows = P.nrows, ncols = P.ncols;
    if (x > xe) { double t = x; x = xe; xe = t; t = y; y = ye; ye = t; }
    if (fabs(__dsub_rn(x, xe)) < 0.01) {  // vertical
        if (ye < y) { double t = y; y = ye; ye = t; }
        long long ix = sat_i64(floor(xe)), iy = sat_i64(floor(y));
        long long iy_end = sat_i64(floor(__dsub_rn(ye, EPS_INTERSECT)));
        if (ix < 0 || ix >= ncols) return;
        iy = max(iy, 0LL);
        iy_end = min(iy_end, nrows - 1);
        for (long long r = iy; r <= iy_end; r++) f((uint32_t)r, (uint32_t)ix);
        return;
    }
    if (fabs(__dsub_rn(y, ye)) < 0.01) {  // horizontal
        if (xe < x) { double t = x; x = xe; xe = t; }
        long long ix = sat_i64(floor(x)), iy = sat_i64(floor(y));
        long long ix_end = sat_i64(floor(__dsub_rn(xe, EPS_INTERSECT)));
        if (iy < 0 || iy >= nrows) return;
        ix = max(ix, 0LL);
        ix_end = min(ix_end, ncols - 1);
        for (long long c = ix; c <= ix_end; c++) f((uint32_t)iy, (uint32_t)c);
        return;
    }
    const double slope = __ddiv_rn(__dsub_rn(ye, y), __dsub_rn(xe, x));
    const double inv_slope = __ddiv_rn(1.0, slope);
    if (x < 0.0) { y = __dadd_rn(y, __dmul_rn(__dsub_rn(0.0, x), slope)); x = 0.0; }
    if (xe > P.ncols_f) { ye = __dadd_rn(ye, __dmul_rn(__dsub_rn(P.ncols_f, xe), slope)); xe = P.ncols_f; }
    if (y < 0.0) { x = __dadd_rn(x, __dmul_rn(__dsub_rn(0.0, y), inv_slope)); y = 0.0; }
    else if (y > P.nrows_f) { x = __dadd_rn(x, __dmul_rn(__dsub_rn(P.nrows_f, y), inv_slope)); y = P.nrows_f; }
    if (ye < 0.0) xe = __dadd_rn(xe, __dmul_rn(__dsub_rn(0.0, ye), inv_slope));
    else if (ye > P.nrows_f) xe = __dadd_rn(xe, __dmul_rn(__dsub_rn(P.nrows_f, ye), inv_slope));
    // f64::clamp keeps NaN
    x = x < 0.0 ? 0.0 : (x > P.ncols_f ? P.ncols_f : x);
    xe = xe < 0.0 ? 0.0 : (xe > P.ncols_f ? P.ncols_f : xe);
    // every iteration advances x by at least TOL/|slope| > 0; the bound only guards against NaN/denormal stalls
    for (uint32_t guard = 0; x >= 0.0 && x < xe && guard < 0x7fffffffu; guard++) {
        const long long ix = sat_i64(floor(x)), iy = sat_i64(floor(y));
        if (ix >= 0 && ix < ncols && iy >= 0 && iy < nrows) f((uint32_t)iy, (uint32_t)ix);
        double sx = __dsub_rn(floor(__dadd_rn(x, 1.0)), x);
        double sy = __dmul_rn(sx, slope);
        if (sat_i64(floor(__dadd_rn(y, sy))) == iy) {
            x = __dadd_rn(x, sx);
            y = __dadd_rn(y, sy);
        } else if (slope < 0.0) {
            sy = __dsub_rn((double)iy, y);
            if (sy > -TOL) sy = -TOL;
            sx = __ddiv_rn(sy, slope);
            x = __dadd_rn(x, sx);
            y = __dadd_rn(y, sy);
        } else {
            sy = __dsub_rn((double)(iy + 1), y);
            if (sy < TOL) sy = TOL;
            sx = __ddiv_rn(sy, slope);
            x = __dadd_rn(x, sx);
            y = __dadd_rn(y, sy);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// PixelCache — rust/src/rasterization/pixel_cache.rs, writers.rs:15-60
// ---------------------------------------------------------------------------------------------
// Open-addressing hash set keyed by (part,row,col) holding the smallest burn index that wrote the
// pixel: a write is kept iff it is that first visit (LineWriter::write, writers.rs:25-29).
struct VisitSet {
    unsigned long long* keys;  // ~0 = empty
    unsigned long long* first; // smallest burn index
    unsigned long long mask;   // capacity - 1 (power of two)
    uint32_t col_bits, row_bits;
    __device__ __forceinline__ unsigned long long key_of(uint32_t part, uint32_t row, uint32_t col) const {
        return ((((unsigned long long)part << row_bits) | row) << col_bits) | col;
    }
    __device__ __forceinline__ unsigned long long slot_of(unsigned long long k) const {
        k ^= k >> 33;
        k *= 0xff51afd7ed558ccdull;
        k ^= k >> 33;
        k *= 0xc4ceb9fe1a85ec53ull;
        k ^= k >> 33;
        return k & mask;
    }
    __device__ __forceinline__ void insert(unsigned long long k, unsigned long long burn) const {
        for (unsigned long long h = slot_of(k);; h = (h + 1) & mask) {
            const unsigned long long old = atomicCAS(&keys[h], ~0ull, k);
            if (old == ~0ull || old == k) {
                atomicMin(&first[h], burn);
                return;
            }
        }
    }
    __device__ __forceinline__ bool contains(unsigned long long k) const {
        for (unsigned long long h = slot_of(k);; h = (h + 1) & mask) {
            const unsigned long long cur = keys[h];
            if (cur == k) return true;
            if (cur == ~0ull) return false;
        }
    }
    __device__ __forceinline__ bool is_first(unsigned long long k, unsigned long long burn) const {
        for (unsigned long long h = slot_of(k);; h = (h + 1) & mask)
            if (keys[h] == k) return first[h] == burn;
    }
};

// The reference's PixelCache is a bitset over the bounding box of the line segments extract_line kept
// (pixel_cache.rs:15-37).  CacheBox is that box per polygon part; cache_contains() is PixelCache::contains
// INCLUDING its behaviour for pixels outside the box, which FillWriter does ask about when a ring segment lies
// entirely outside the raster and was dropped (edges.rs:124-132): unravel_index (pixel_cache.rs:39-44)
// wraps, so such a pixel aliases onto another cell of the box - or falls off the bitset, which reads false.
struct CacheBox {
    long long xmin, ymin;             // (x_lo as isize, y_lo as isize): truncation, not floor
    unsigned long long width, length; // floor(hi) - floor(lo) + 1
};
struct CacheAcc {  // per part: order-preserving u64 encodings of the kept segments' min / max ordinates
    unsigned long long xlo, ylo, xhi, yhi;
    unsigned int dropped, pad;  // a ring segment failed extract_line's test
};
__device__ __forceinline__ unsigned long long f64_ordered(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_ordered(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}
__device__ __forceinline__ bool cache_contains(uint32_t nrows, uint32_t ncols, const CacheBox& b, const VisitSet& vs,
                                               uint32_t part, uint32_t row, uint32_t col) {
    const unsigned long long lx = (unsigned long long)((long long)col - b.xmin);
    const unsigned long long ly = (unsigned long long)((long long)row - b.ymin);
    const unsigned long long idx = ly * b.width + lx;  // wrapping, like the release build
    if (idx >= b.width * b.length) return false;        // FixedBitSet::contains past the end
    const unsigned long long cy = idx / b.width, cx = idx - cy * b.width;
    const long long ax = (long long)cx + b.xmin, ay = (long long)cy + b.ymin;
    if (ax < 0 || ay < 0 || ax >= (long long)ncols || ay >= (long long)nrows) return false;  // never walked
    return vs.contains(vs.key_of(part, (uint32_t)ay, (uint32_t)ax));
}

// one thread per ring vertex: the segment (i, i+1) either extends its part's box or marks the part
static __global__ void cache_box_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                 const uint32_t* __restrict__ tag, uint32_t n, CacheAcc* __restrict__ acc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (tag[i] & 0x80000000u)) return;
    const uint32_t part = tag[i] & 0x3fffffffu;
    const double x0 = px_x(P, x[i]), y0 = px_y(P, y[i]), x1 = px_x(P, x[i + 1]), y1 = px_y(P, y[i + 1]);
    const double min_x = fmin(x0, x1), max_x = fmax(x0, x1), min_y = fmin(y0, y1), max_y = fmax(y0, y1);
    if (!(min_x < P.ncols_f && max_x >= 0.0 && min_y < P.nrows_f && max_y >= 0.0)) {  // edges.rs:130
        atomicOr(&acc[part].dropped, 1u);
        return;
    }
    atomicMin(&acc[part].xlo, f64_ordered(min_x));
    atomicMin(&acc[part].ylo, f64_ordered(min_y));
    atomicMax(&acc[part].xhi, f64_ordered(max_x));
    atomicMax(&acc[part].yhi, f64_ordered(max_y));
}
static __global__ void cache_box_init_kernel(uint32_t n_parts, CacheAcc* __restrict__ acc) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_parts) return;
    CacheAcc a;
    a.xlo = a.ylo = f64_ordered(DBL_MAX);  // PixelCache::new's fold seeds (pixel_cache.rs:16-17)
    a.xhi = a.yhi = f64_ordered(-DBL_MAX);
    a.dropped = a.pad = 0;
    acc[p] = a;
}
static __global__ void cache_box_finish_kernel(uint32_t n_parts, const CacheAcc* __restrict__ acc, CacheBox* __restrict__ box) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_parts) return;
    const double xlo = f64_from_ordered(acc[p].xlo), ylo = f64_from_ordered(acc[p].ylo);
    const double xhi = f64_from_ordered(acc[p].xhi), yhi = f64_from_ordered(acc[p].yhi);
    auto as_usize = [](double v) -> unsigned long long {  // Rust `f64 as usize`
        if (!(v > 0.0)) return 0ull;
        if (v >= 18446744073709551615.0) return ~0ull;
        return (unsigned long long)v;
    };
    auto as_isize = [](double v) -> long long {  // Rust `f64 as isize`
        if (!(v == v)) return 0;
        if (v <= -9223372036854775808.0) return (long long)0x8000000000000000ull;
        if (v >= 9223372036854775807.0) return 0x7fffffffffffffffll;
        return (long long)v;
    };
    CacheBox b;
    b.width = as_usize(__dsub_rn(floor(xhi), floor(xlo))) + 1ull;
    b.length = as_usize(__dsub_rn(floor(yhi), floor(ylo))) + 1ull;
    b.xmin = as_isize(xlo);
    b.ymin = as_isize(ylo);
    box[p] = b;
}

// One thread per vertex of a ring / line-string pool.  mode 0 counts the in-window pixels, mode 1
// emits one record per pixel (flagged as "pixel of the part's boundary walk").
static __global__ void __launch_bounds__(256)
touched_walk_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                    const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                    Counters* __restrict__ ctr, uint64_t* __restrict__ keys, int mode,
                    const CacheAcc* __restrict__ acc = nullptr, VisitSet vs = VisitSet{}) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cnt = 0;
    if (mode >= 2) {
        // modes 2 / 3 (sum / count only): count / remember every walked pixel - over the whole raster, not the
        // window - of the polygon parts that had a ring segment dropped; their fill may ask the PixelCache
        // about pixels outside its box, which alias onto these cells (cache_contains)
        if (i < n && !(tag[i] & 0x80000000u)) {
            const uint32_t part = tag[i] & 0x3fffffffu;
            if (info[part].band >= 0 && acc[part].dropped) {
                const double x0 = px_x(P, x[i]), y0 = px_y(P, y[i]), x1 = px_x(P, x[i + 1]), y1 = px_y(P, y[i + 1]);
                const double min_x = fmin(x0, x1), max_x = fmax(x0, x1), min_y = fmin(y0, y1), max_y = fmax(y0, y1);
                if (min_x < P.ncols_f && max_x >= 0.0 && min_y < P.nrows_f && max_y >= 0.0)
                    all_touched_walk(P, x0, y0, x1, y1, [&](uint32_t row, uint32_t col) {
                        if (mode == 2) cnt++;
                        else vs.insert(vs.key_of(part, row, col), 0ull);
                    });
            }
        }
        if (mode == 2) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
            if (lane_id() == 0 && cnt) atomicAdd(&ctr->cursor, cnt);
        }
        return;
    }
    if (i < n && !(tag[i] & 0x80000000u)) {
        const uint32_t part = tag[i] & 0x3fffffffu;
        const PartInfo pi = info[part];
        const double x0 = px_x(P, x[i]), y0 = px_y(P, y[i]), x1 = px_x(P, x[i + 1]), y1 = px_y(P, y[i + 1]);
        const double min_x = fmin(x0, x1), max_x = fmax(x0, x1), min_y = fmin(y0, y1), max_y = fmax(y0, y1);
        if (pi.band >= 0 && min_x < P.ncols_f && max_x >= 0.0 && min_y < P.nrows_f && max_y >= 0.0) {  // edges.rs:130
            const uint64_t flag = 1ull << (P.col_bits - 1);
            all_touched_walk(P, x0, y0, x1, y1, [&](uint32_t row, uint32_t col) {
                if (row < P.win_r0 || row >= P.win_r1) return;
                if (mode == 0) {
                    cnt++;
                    return;
                }
                // warp-aggregated append
                const uint32_t active = __activemask();
                const int leader = __ffs(active) - 1;
                unsigned long long base = 0;
                if ((int)lane_id() == leader) base = atomicAdd(&ctr->cursor, (unsigned long long)__popc(active));
                base = __shfl_sync(active, base, leader);
                keys[base + __popc(active & ((1u << lane_id()) - 1u))] = pixel_key(P, pi.band, part, row, col) | flag;
            });
        }
    }
    if (mode == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        if (lane_id() == 0 && cnt) atomicAdd(&ctr->records, cnt);
    }
}

// ---------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, 64-bit keys (keys only)
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per block
constexpr int RS_RADIX = 256;

// per-block digit histogram, stored digit-major: hist[d * n_blocks + b]
static __global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t shift, uint32_t n_blocks,
                  uint32_t* __restrict__ hist) {
    __shared__ uint32_t s[RS_RADIX];
    s[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s[(uint32_t)(keys[i] >> shift) & (RS_RADIX - 1)], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * n_blocks + blockIdx.x] = s[threadIdx.x];
}

// one block per digit: exclusive scan of its row of block counts (in place) + digit total
static __global__ void __launch_bounds__(1024)
radix_scan_rows_kernel(uint32_t* __restrict__ hist, uint32_t n_blocks, uint32_t* __restrict__ digit_total) {
    __shared__ uint32_t s_warp[33];
    uint32_t* row = hist + (size_t)blockIdx.x * n_blocks;
    uint32_t carry = 0;
    uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        uint32_t i = b0 + threadIdx.x;
        uint32_t v = i < n_blocks ? row[i] : 0, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= (uint32_t)o) winc += t;
            }
            s_warp[lane] = winc - w;
            if (lane == 31) s_warp[32] = winc;
        }
        __syncthreads();
        if (i < n_blocks) row[i] = carry + s_warp[warp] + inc - v;
        carry += s_warp[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

// Stable scatter.  Keys are ranked warp by warp, 32 consecutive keys per round, with
// __match_any_sync; the tile is then staged in shared memory in sorted order so that global stores
// go out in runs of equal digits.
static __global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out, uint32_t n, uint32_t shift,
                     uint32_t n_blocks, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ digit_total) {
    __shared__ uint64_t s_keys[RS_TILE];
    __shared__ uint32_t s_cnt[RS_THREADS / 32][RS_RADIX];
    __shared__ uint32_t s_dbase[RS_RADIX];   // first block-sorted position of each digit
    __shared__ uint32_t s_goff[RS_RADIX];    // global offset minus block-sorted position
    __shared__ uint32_t s_warp[33];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (int w = 0; w < RS_THREADS / 32; w++) s_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile0 = blockIdx.x * RS_TILE;
    const uint32_t wbase = tile0 + warp * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint16_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint32_t i = wbase + r * 32 + lane;
        key[r] = i < n ? keys_in[i] : ~0ull;  // padding sorts after every real key of the tile
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint32_t d = (uint32_t)(key[r] >> shift) & (RS_RADIX - 1);
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t c = 0;
        if ((int)lane == leader) {
            c = s_cnt[warp][d];
            s_cnt[warp][d] = c + __popc(peers);
        }
        c = __shfl_sync(0xffffffffu, c, leader);
        rank[r] = (uint16_t)(c + __popc(peers & ((1u << lane) - 1)));
        __syncwarp();
    }
    __syncthreads();
    // thread d: turn per-warp counts of digit d into exclusive warp bases, get the block count
    uint32_t d = threadIdx.x, run = 0;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; w++) {
        uint32_t t = s_cnt[w][d];
        s_cnt[w][d] = run;
        run += t;
    }
    uint32_t tot;
    uint32_t dbase = block_exclusive_scan(run, s_warp, &tot);
    // global base of digit d for this block = (keys of smaller digits) + (digit d in earlier blocks)
    uint32_t gd = 0;
    {
        // exclusive scan of digit totals (256 values) — reuse the same block scan
        uint32_t t2;
        uint32_t dt = digit_total[d];
        gd = block_exclusive_scan(dt, s_warp, &t2);
    }
    s_dbase[d] = dbase;
    s_goff[d] = gd + hist[d * n_blocks + blockIdx.x] - dbase;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint32_t dd = (uint32_t)(key[r] >> shift) & (RS_RADIX - 1);
        s_keys[s_dbase[dd] + s_cnt[warp][dd] + rank[r]] = key[r];
    }
    __syncthreads();
    uint32_t n_valid = min((uint32_t)RS_TILE, n - tile0);
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        uint32_t pos = r * RS_THREADS + threadIdx.x;
        if (pos < n_valid) {
            uint64_t k = s_keys[pos];
            uint32_t dd = (uint32_t)(k >> shift) & (RS_RADIX - 1);
            keys_out[s_goff[dd] + pos] = k;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// task index: first record of every task (lower bound on the sorted keys)
// ---------------------------------------------------------------------------------------------
static __global__ void task_index_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t task_shift, uint32_t n_tasks,
                                  uint32_t* __restrict__ task_start) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tasks) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if ((keys[mid] >> task_shift) < (uint64_t)t) lo = mid + 1;
        else hi = mid;
    }
    task_start[t] = lo;
}

// ---------------------------------------------------------------------------------------------
// pixel functions — rust/src/rasterization/pixel_functions.rs:56-123
// ---------------------------------------------------------------------------------------------
template <typename N> __device__ __forceinline__ bool is_nan_v(N) { return false; }
template <> __device__ __forceinline__ bool is_nan_v<float>(float v) { return v != v; }
template <> __device__ __forceinline__ bool is_nan_v<double>(double v) { return v != v; }

template <typename N> struct Wrap { typedef N U; };
template <> struct Wrap<int8_t> { typedef uint8_t U; };
template <> struct Wrap<int16_t> { typedef uint16_t U; };
template <> struct Wrap<int32_t> { typedef uint32_t U; };
template <> struct Wrap<int64_t> { typedef uint64_t U; };
// Rust release-mode `+=` (wrapping for integers; one rounding for floats)
template <typename N> __device__ __forceinline__ N add_v(N a, N b) {
    typedef typename Wrap<N>::U U;
    return (N)(U)((U)a + (U)b);
}
template <> __device__ __forceinline__ float add_v<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_v<double>(double a, double b) { return __dadd_rn(a, b); }

// BGNAN: the background is NaN, so `cur == bg` can never hold and the test reduces to is_nan(cur)
// VOK: the caller has checked that v is not NaN (sum's `isnan(v)` test is then dropped)
template <typename N, int FN, bool BGNAN = false, bool VOK = false>
__device__ __forceinline__ N apply_px(N cur, N v, N bg) {
    bool untouched = BGNAN ? is_nan_v(cur) : ((cur == bg) || is_nan_v(cur));
    if (FN == RZ_SUM) return (untouched || (!VOK && is_nan_v(v))) ? v : add_v(cur, v);
    if (FN == RZ_FIRST) return untouched ? v : cur;
    if (FN == RZ_LAST) return v;
    if (FN == RZ_MIN) return (untouched || cur > v) ? v : cur;
    if (FN == RZ_MAX) return (untouched || cur < v) ? v : cur;
    if (FN == RZ_COUNT) return untouched ? (N)1 : add_v(cur, (N)1);
    return (N)1;  // any
}

template <typename N> __device__ __forceinline__ N value_from_bits(uint64_t b) {
    N v;
    memcpy(&v, &b, sizeof(N));
    return v;
}

// ---------------------------------------------------------------------------------------------
// fill — one warp per (band, row, column tile)
// ---------------------------------------------------------------------------------------------
// The task's records arrive grouped by part, parts in burn order (stable sort).  A polygon part's
// records of this row ("run") are its scanline crossings in arbitrary order: the warp XORs one bit
// per crossing into a 1024-bit toggle mask (one 32-bit word per lane); a prefix-XOR over the mask
// is the even-odd inside mask — exactly the spans burners.rs:302-315 gets from sorting and pairing,
// because pairing consecutive sorted crossings == parity of the number of crossings at or left of a
// pixel.  An unpaired last crossing (odd run) is dropped like chunks_exact(2) does, by removing the
// largest column.  The part's value is then applied to the masked pixels with the reference's
// pixel-function rule, 32 consecutive pixels per step, before the next part is looked at: writes
// hit every pixel in burn order, so even floating-point `sum` is bit-exact.
struct FillParams {
    uint32_t n_tasks, n_tiles, tile_w, ncols;
    uint32_t win_rows;         // rows in this window
    uint32_t win_row_off;      // first window row relative to the output's first row
    uint32_t out_rows;         // rows per band in `out`
    uint32_t col_bits, part_shift, part_bits;
    uint32_t dedup_lines;
    uint32_t vec_ok;           // rows of `out` are 16-byte aligned
    uint32_t all_poly;         // no line / point parts: skip the kind lookup
    uint32_t all_touched;      // polygon runs also carry boundary-walk pixels (flagged records)
    uint32_t nrows;            // raster rows (full raster)
    uint32_t win_r0;           // first raster row of this window
};

// all_touched with sum / count: what FillWriter needs for parts whose PixelCache box does not cover their fill
struct AliasCtx {
    const CacheAcc* acc;  // nullptr: not in that mode
    const CacheBox* box;
    VisitSet vs;
};

constexpr int FILL_WARPS = 4;
constexpr uint32_t FILL_MAX_TILE_W = 1024;  // 32 lanes x 32 toggle bits

// Rare path: a polygon run with an odd number of crossings.  burners.rs:305 pairs the sorted
// crossings with chunks_exact(2), i.e. the largest column is ignored: cancel one toggle there.
static __device__ __noinline__ void drop_last_crossing(uint32_t* tog, const uint64_t* __restrict__ keys, uint32_t beg,
                                                uint32_t end, uint32_t col_mask, uint32_t flag_bit, uint32_t w,
                                                uint32_t lane) {
    uint32_t mx = 0;
    for (uint32_t i = beg + lane; i < end; i += 32) {
        const uint32_t k = (uint32_t)keys[i];
        if (!(k & flag_bit)) mx = max(mx, k & col_mask);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (mx < w && lane == (mx >> 5)) tog[lane] ^= 1u << (mx & 31);
    __syncwarp();
}

// Apply value v to the pixels whose bit is set in the warp's inside mask (word `lane` of m covers
// pixels lane*32 .. lane*32+31).  One 32-pixel word per step, conflict-free shared-memory access.
template <typename N, int FN>
__device__ __forceinline__ void apply_mask(N* __restrict__ row_lane, uint32_t m, uint32_t lane, N v, N bg) {
    uint32_t nz = __ballot_sync(0xffffffffu, m != 0);
    while (nz) {
        const int src = __ffs(nz) - 1;
        nz &= nz - 1;
        const uint32_t mw = __shfl_sync(0xffffffffu, m, src);
        N* p = row_lane + src * 32;
        const N cur = *p;
        const N nv = apply_px<N, FN>(cur, v, bg);
        *p = ((mw >> lane) & 1u) ? nv : cur;
    }
}

// Close a polygon run: turn the toggle mask into the even-odd inside mask and burn the value.
// Bits at or beyond the tile's width may end up set; they only touch shared-memory pixels that are
// never flushed.
// FillWriter for a part with a dropped ring segment: remove from the inside mask every pixel the reference's
// PixelCache claims to contain (cache_contains: aliased cells for pixels outside the cache's box)
static __device__ __noinline__ uint32_t drop_cached_fill(uint32_t m, uint32_t walked, const FillParams& F, const AliasCtx& A,
                                                  uint32_t part, uint32_t row, uint32_t c0, uint32_t w, uint32_t lane) {
    const CacheBox b = A.box[part];
    uint32_t cand = m & ~walked;
    while (cand) {
        const uint32_t bit = (uint32_t)__ffs(cand) - 1u;
        cand &= cand - 1;
        const uint32_t rel = lane * 32 + bit;
        if (rel < w && cache_contains(F.nrows, F.ncols, b, A.vs, part, row, c0 + rel)) m &= ~(1u << bit);
    }
    return m;
}

template <typename N, int FN>
__device__ __forceinline__ void finish_poly_run(uint32_t* tog, uint32_t* orm, N* row_lane, uint32_t lane,
                                                uint32_t lt_mask, N v, N bg, const FillParams& F, const AliasCtx& A,
                                                uint32_t part, uint32_t row, uint32_t c0, uint32_t w) {
    const uint32_t t = tog[lane];
    tog[lane] = 0;
    const uint32_t walked = orm[lane];  // all_touched: pixels of the ring walk (burn_geometry.rs:225-238)
    orm[lane] = 0;
    uint32_t m = t;
    m ^= m << 1;
    m ^= m << 2;
    m ^= m << 4;
    m ^= m << 8;
    m ^= m << 16;
    const uint32_t odd_words = __ballot_sync(0xffffffffu, __popc(t) & 1);
    if (__popc(odd_words & lt_mask) & 1) m = ~m;
    if (A.acc && A.acc[part].dropped) m = drop_cached_fill(m, walked, F, A, part, row, c0, w, lane);
    m |= walked;
    apply_mask<N, FN>(row_lane, m, lane, v, bg);
    __syncwarp();
}

template <typename N, int FN, bool ALL_POLY>
static __global__ void __launch_bounds__(FILL_WARPS * 32)
fill_kernel(FillParams F, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ task_start,
            const PartInfo* __restrict__ info, const uint8_t* __restrict__ part_kind, uint64_t bg_bits,
            N* __restrict__ out, AliasCtx A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_tog[FILL_WARPS][32], s_orm[FILL_WARPS][32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    N* row = reinterpret_cast<N*>(smem_raw) + (size_t)warp * FILL_MAX_TILE_W;
    N* row_lane = row + lane;
    uint32_t* tog = s_tog[warp];
    uint32_t* orm = s_orm[warp];
    orm[lane] = 0;
    const N bg = value_from_bits<N>(bg_bits);
    const uint32_t col_mask = (1u << (F.col_bits - 1)) - 1u;  // the top bit of the column field is the walk flag
    const uint32_t flag_bit = 1u << (F.col_bits - 1);
    const uint64_t part_mask = (1ull << F.part_bits) - 1ull;
    const uint32_t lt_mask = (1u << lane) - 1u;
    tog[lane] = 0;

    for (uint32_t task = blockIdx.x * FILL_WARPS + warp; task < F.n_tasks; task += gridDim.x * FILL_WARPS) {
        const uint32_t tile = task % F.n_tiles;
        const uint32_t br = task / F.n_tiles;  // band * win_rows + row
        const uint32_t band = br / F.win_rows, r = br - band * F.win_rows;
        const uint32_t c0 = tile * F.tile_w;
        const uint32_t w = min(F.tile_w, F.ncols - c0);
        N* dst = out + ((size_t)band * F.out_rows + F.win_row_off + r) * F.ncols + c0;
        const uint32_t beg = task_start[task], end = task_start[task + 1];
        uint64_t key_next = beg + lane < end ? keys[beg + lane] : ~0ull;

        // every pixel of the warp's tile starts as background (geo/raster.rs:23-28)
        for (uint32_t i = lane; i < FILL_MAX_TILE_W; i += 32) row[i] = bg;
        __syncwarp();

        uint64_t carry_hi = ~0ull;   // (task|part) of the record before this chunk
        uint32_t run_cnt = 0;        // records of the currently open polygon run
        uint32_t run_beg = beg;      // its first record
        uint32_t open_kind = 3;      // kind of the run left open by the previous chunk (3 = none)
        uint32_t open_part = 0;
        N open_v = bg;
        const uint32_t abs_row = F.win_r0 + r;
        for (uint32_t base = beg; base < end; base += 32) {
            const uint32_t nvalid = min(32u, end - base);
            const bool valid = lane < nvalid;
            const uint64_t key = key_next;
            key_next = base + 32 + lane < end ? keys[base + 32 + lane] : ~0ull;  // prefetch
            const uint64_t hi = key >> F.col_bits;
            uint64_t prev_hi = __shfl_up_sync(0xffffffffu, hi, 1);
            if (lane == 0) prev_hi = carry_hi;
            const uint32_t heads = __ballot_sync(0xffffffffu, valid && hi != prev_hi);
            const uint32_t part = (uint32_t)(hi & part_mask);
            const uint32_t col = (uint32_t)key & col_mask;
            const bool walk_px = ((uint32_t)key & flag_bit) != 0;
            const N my_v = valid ? value_from_bits<N>(info[part].value_bits) : bg;
            const bool last_chunk = base + 32 >= end;

            if ((heads & 1u) && open_kind != 3) {  // the open run ended exactly at the chunk boundary
                if (open_kind == 0) {
                    if (run_cnt & 1u) drop_last_crossing(tog, keys, run_beg, base, col_mask, flag_bit, w, lane);
                    finish_poly_run<N, FN>(tog, orm, row_lane, lane, lt_mask, open_v, bg, F, A, open_part, abs_row, c0, w);
                } else if (open_kind == 1) {
                    tog[lane] = 0;
                    __syncwarp();
                }
                run_cnt = 0;
            }
            open_kind = 3;

            uint32_t start = 0;
            while (start < nvalid) {  // one iteration per run present in this chunk (warp-uniform)
                const uint32_t rest = start < 31 ? (heads & (0xfffffffeu << start)) : 0u;
                const uint32_t stop = rest ? (uint32_t)(__ffs(rest) - 1) : nvalid;
                const bool in_run = lane >= start && lane < stop;
                const bool run_ends = stop < nvalid || last_chunk;
                const N v = __shfl_sync(0xffffffffu, my_v, start);
                uint32_t kind = 0;
                const uint32_t run_part = __shfl_sync(0xffffffffu, part, start);
                if (!ALL_POLY) kind = part_kind[run_part];
                if ((heads >> start) & 1u) run_beg = base + start;
                if (kind == 0) {
                    if (F.all_touched) {
                        if (in_run && col < w) {
                            if (walk_px) atomicOr(&orm[col >> 5], 1u << (col & 31));
                            else atomicXor(&tog[col >> 5], 1u << (col & 31));
                        }
                        run_cnt += __popc(__ballot_sync(0xffffffffu, in_run && !walk_px));
                    } else {
                        if (in_run && col < w) atomicXor(&tog[col >> 5], 1u << (col & 31));
                        run_cnt += stop - start;
                    }
                    __syncwarp();
                    if (run_ends) {
                        if (run_cnt & 1u) drop_last_crossing(tog, keys, run_beg, base + stop, col_mask, flag_bit, w, lane);
                        finish_poly_run<N, FN>(tog, orm, row_lane, lane, lt_mask, v, bg, F, A, run_part, abs_row, c0, w);
                        run_cnt = 0;
                    }
                } else {
                    // line / point pixels: one write per record (burners.rs:69-89, 250-258)
                    const bool dedup = kind == 1 && F.dedup_lines;  // LineWriter + PixelCache (writers.rs:25-29)
                    if (dedup) {
                        if (in_run) {
                            const uint32_t bit = 1u << (col & 31);
                            if (!(atomicOr(&tog[col >> 5], bit) & bit)) row[col] = apply_px<N, FN>(row[col], v, bg);
                        }
                        __syncwarp();
                        if (run_ends) {
                            tog[lane] = 0;
                            __syncwarp();
                        }
                    } else {
                        // revisits are written again: apply once per record, same-pixel records serially
                        const uint32_t run_mask = __ballot_sync(0xffffffffu, in_run);
                        if (in_run) {
                            const uint32_t peers = __match_any_sync(run_mask, col);
                            if ((int)lane == __ffs(peers) - 1) {
                                N cur = row[col];
                                for (int k = __popc(peers); k > 0; k--) cur = apply_px<N, FN>(cur, v, bg);
                                row[col] = cur;
                            }
                        }
                        __syncwarp();
                    }
                }
                if (!run_ends) {
                    open_kind = kind == 0 ? 0u : ((kind == 1 && F.dedup_lines) ? 1u : 2u);
                    open_v = v;
                    open_part = run_part;
                }
                start = stop;
            }
            carry_hi = __shfl_sync(0xffffffffu, hi, 31);
        }

        // flush the tile: every output byte is written exactly once
        if (F.vec_ok && (w * sizeof(N)) % 16 == 0) {
            const uint4* s4 = reinterpret_cast<const uint4*>(row);
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            const uint32_t n4 = (uint32_t)(w * sizeof(N) / 16);
            for (uint32_t i = lane; i < n4; i += 32) __stcs(d4 + i, s4[i]);
        } else {
            for (uint32_t i = lane; i < w; i += 32) dst[i] = row[i];
        }
        __syncwarp();
    }
}

}  // namespace rz
