// rz_sparse.cuh — device code of the sparse (COO) output path.
//
// The reference's SparseArrayWriter pushes one (row, col, value) triplet per pixel WRITE, per band,
// in burn order (rust/src/encoding/writers.rs:86-131):  band -> geometry -> part -> burn order, where
// a polygon part burns rows top to bottom and, inside a row, spans left to right
// (burners.rs:284-319), a line part burns its kept segments in order, each Bresenham step in order
// (burners.rs:54-89), and a point part its points in order (burners.rs:250-258).
//
// GPU formulation: every write unit (polygon span / line segment run / point) gets
//   position = part_base[part] + (prefix of unit lengths inside the part)
// from device-wide scans, and an expand kernel writes the triplets at their final offsets.
#pragma once

#include "rz_kernels.cuh"

namespace rz {

// ---------------------------------------------------------------------------------------------
// generic device-wide scan:  out(i, exclusive_prefix(i))  for values in(i), i < n
// ---------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

struct OpAdd {
    __device__ static unsigned long long identity() { return 0ull; }
    __device__ static unsigned long long combine(unsigned long long a, unsigned long long b) { return a + b; }
};
struct OpMax {
    __device__ static unsigned long long identity() { return 0ull; }
    __device__ static unsigned long long combine(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
};

// inclusive block scan of one u64 per thread; returns the inclusive value, *total = block aggregate
template <typename Op>
__device__ __forceinline__ unsigned long long block_inclusive_scan64(unsigned long long v, unsigned long long* s_warp,
                                                                     unsigned long long* total) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v = Op::combine(t, v);
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < (blockDim.x >> 5) ? s_warp[lane] : Op::identity();
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w = Op::combine(t, w);
        }
        s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    if (warp > 0) v = Op::combine(s_warp[warp - 1], v);
    *total = s_warp[(blockDim.x >> 5) - 1];
    __syncthreads();
    return v;
}

template <typename Op, typename In>
static __global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(In in, uint32_t n, unsigned long long* __restrict__ partial) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long acc = Op::identity();
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++)
        if (base + k < n) acc = Op::combine(acc, in(base + k));
    unsigned long long total;
    block_inclusive_scan64<Op>(acc, s_warp, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// single block: exclusive scan of the block aggregates in place; partial[nb] = grand total
template <typename Op>
static __global__ void __launch_bounds__(1024) scan_partials_kernel(unsigned long long* __restrict__ partial, uint32_t nb) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long carry = Op::identity();
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const unsigned long long v = i < nb ? partial[i] : Op::identity();
        unsigned long long total;
        const unsigned long long inc = block_inclusive_scan64<Op>(v, s_warp, &total);
        // exclusive = combine(carry, inclusive of the previous element)
        unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane_id() == 0) prev = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : Op::identity();
        if (i < nb) partial[i] = Op::combine(carry, threadIdx.x == 0 ? Op::identity() : prev);
        carry = Op::combine(carry, total);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[nb] = carry;
}

template <typename Op, typename In, typename Out>
static __global__ void __launch_bounds__(SC_THREADS)
scan_apply_kernel(In in, uint32_t n, const unsigned long long* __restrict__ partial, Out out) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long v[SC_ITEMS];
    unsigned long long acc = Op::identity();
    const uint32_t base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        v[k] = base + k < n ? in(base + k) : Op::identity();
        acc = Op::combine(acc, v[k]);
    }
    unsigned long long total;
    const unsigned long long inc = block_inclusive_scan64<Op>(acc, s_warp, &total);
    unsigned long long prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane_id() == 0) prev = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : Op::identity();
    // NOTE: s_warp was re-synchronised inside block_inclusive_scan64 before returning; the read above is
    // of values written before that barrier and not modified afterwards.
    unsigned long long run = Op::combine(partial[blockIdx.x], threadIdx.x == 0 ? Op::identity() : prev);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        if (base + k < n) out(base + k, run, Op::combine(run, v[k]));
        run = Op::combine(run, v[k]);
    }
}

// ---------------------------------------------------------------------------------------------
// key layout of sparse polygon crossings:  [part | row | col]   (col in 0..ncols)
// ---------------------------------------------------------------------------------------------
struct SparseLayout {
    uint32_t col_bits, row_bits;
};

__device__ __forceinline__ uint64_t sparse_poly_key(const KParams& P, const SparseLayout& L, const PolyEdgeRec& e,
                                                    uint32_t row) {
    double cy = __dadd_rn((double)row, 0.5);
    double xi = __dadd_rn(e.x_top, __dmul_rn(__dsub_rn(cy, e.y_top), e.dxdy));  // edges.rs:50-55
    uint32_t col = sat_u32(floor(__dadd_rn(xi, 0.5)), P.ncols);                 // burners.rs:310-311
    return ((((uint64_t)e.part << L.row_bits) | row) << L.col_bits) | col;
}

// same structure as poly_emit_kernel, one record per (edge, row), sparse key
static __global__ void __launch_bounds__(SETUP_THREADS)
poly_emit_sparse_kernel(KParams P, SparseLayout L, const double* __restrict__ x, const double* __restrict__ y,
                        const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                        const uint32_t* __restrict__ block_base, uint64_t* __restrict__ keys) {
    __shared__ uint32_t s_wtot[SETUP_THREADS / 32];
    __shared__ double s_xtop[SETUP_THREADS], s_ytop[SETUP_THREADS], s_dxdy[SETUP_THREADS];
    __shared__ uint32_t s_pre[SETUP_THREADS], s_rowlo[SETUP_THREADS], s_part[SETUP_THREADS];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, w0 = warp * 32;
    PolyEdgeRec e;
    poly_edge_setup(P, x, y, tag, info, blockIdx.x * SETUP_THREADS + threadIdx.x, n, e);
    const uint32_t cnt = e.n_rows;  // n_t == 1 in sparse mode
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
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < warp; w++) wbase += s_wtot[w];
    uint64_t* out = keys + block_base[blockIdx.x] + wbase;
    for (uint32_t t = lane; t < wtot; t += 32) {
        uint32_t lo = 0, hi = 32;
#pragma unroll
        for (int it = 0; it < 5; it++) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_pre[w0 + mid] <= t) lo = mid;
            else hi = mid;
        }
        const uint32_t j = w0 + lo;
        PolyEdgeRec b;
        b.x_top = s_xtop[j];
        b.y_top = s_ytop[j];
        b.dxdy = s_dxdy[j];
        b.part = s_part[j];
        out[t] = sparse_poly_key(P, L, b, s_rowlo[j] + (t - s_pre[j]));
    }
}

// ---------------------------------------------------------------------------------------------
// line pixels in burn order + first-visit filter (PixelCache, pixel_cache.rs / writers.rs:15-36)
// ---------------------------------------------------------------------------------------------
// Calls f(part, row, col, j) for the j-th pixel write of segment i (clipped Bresenham run, then the
// end pixel of the part's last kept segment when its line string is open, burners.rs:87-89).
// TOUCHED: the pixels of the all_touched walk of segment i instead (burners.rs:94-247); the pool may then be
// the polygon rings (pass 1 of burn_geometry.rs:225-238) as well as the line strings.
template <bool TOUCHED, typename F>
__device__ __forceinline__ void for_each_line_pixel(const KParams& P, const double* __restrict__ x,
                                                    const double* __restrict__ y, const uint32_t* __restrict__ tag,
                                                    const PartInfo* __restrict__ info,
                                                    const uint32_t* __restrict__ last_kept, Counters* ctr, uint32_t i,
                                                    uint32_t n, F f) {
    if (TOUCHED) {
        if (i >= n || (tag[i] & 0x80000000u)) return;
        const uint32_t part = tag[i] & 0x3fffffffu;
        if (info[part].band < 0) return;
        const double x0 = px_x(P, x[i]), y0 = px_y(P, y[i]), x1 = px_x(P, x[i + 1]), y1 = px_y(P, y[i + 1]);
        const double min_x = fmin(x0, x1), max_x = fmax(x0, x1), min_y = fmin(y0, y1), max_y = fmax(y0, y1);
        if (!(min_x < P.ncols_f && max_x >= 0.0 && min_y < P.nrows_f && max_y >= 0.0)) return;  // edges.rs:130
        uint32_t j = 0;
        all_touched_walk(P, x0, y0, x1, y1, [&](uint32_t row, uint32_t col) { f(part, row, col, j++); });
        return;
    }
    LineRec l;
    bool kept;
    line_setup(P, x, y, tag, info, i, n, l, &kept, ctr);
    if (!kept) return;
    uint32_t j = 0;
    for (uint32_t k = 0; k < l.n; k++, j++) {
        long long px, py;
        line_pixel(l, (long long)l.k_lo + k, px, py);
        f(l.part, (uint32_t)py, (uint32_t)px, j);
    }
    if (last_kept[l.part] == i + 1 && !(tag[i] & 0x40000000u)) {
        long long ix1 = sat_i64(floor(px_x(P, x[i + 1]))), iy1 = sat_i64(floor(px_y(P, y[i + 1])));
        if (ix1 >= 0 && ix1 < (long long)P.ncols && iy1 >= 0 && iy1 < (long long)P.nrows)
            f(l.part, (uint32_t)iy1, (uint32_t)ix1, j);
    }
}

template <bool TOUCHED>
static __global__ void line_visit_insert_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                         const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                                         const uint32_t* __restrict__ last_kept, Counters* __restrict__ ctr,
                                         const unsigned long long* __restrict__ raw_off, VisitSet vs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long b0 = raw_off[i];
    for_each_line_pixel<TOUCHED>(P, x, y, tag, info, last_kept, ctr, i, n,
                                 [&](uint32_t part, uint32_t row, uint32_t col, uint32_t j) {
                                     vs.insert(vs.key_of(part, row, col), b0 + j);
                                 });
}

// ---------------------------------------------------------------------------------------------
// functors for the scans
// ---------------------------------------------------------------------------------------------
// head position of the (part,row) segment a sorted crossing belongs to -> max-scan gives seg start
struct InSegHead {
    const uint64_t* keys;
    uint32_t col_bits;
    __device__ unsigned long long operator()(uint32_t i) const {
        if (i == 0) return 0ull;
        return (keys[i] >> col_bits) != (keys[i - 1] >> col_bits) ? (unsigned long long)i : 0ull;
    }
};
struct OutSegStart {
    uint32_t* seg_start;
    __device__ void operator()(uint32_t i, unsigned long long, unsigned long long inclusive) const {
        seg_start[i] = (uint32_t)inclusive;
    }
};
// length of the span that starts at sorted crossing i (0 when i does not start a span):
// pairs (0,1),(2,3).. of a segment; the unpaired last crossing is dropped (burners.rs:305-315)
struct InSpanLen {
    const uint64_t* keys;
    const uint32_t* seg_start;
    uint32_t n, col_bits;
    __device__ unsigned long long operator()(uint32_t i) const {
        if (((i - seg_start[i]) & 1u) || i + 1 >= n) return 0ull;
        const uint64_t a = keys[i], b = keys[i + 1];
        if ((a >> col_bits) != (b >> col_bits)) return 0ull;
        const uint32_t m = (1u << col_bits) - 1u;
        const uint32_t ca = (uint32_t)a & m, cb = (uint32_t)b & m;
        return cb > ca ? (unsigned long long)(cb - ca) : 0ull;
    }
};
struct OutPrefix64 {
    unsigned long long* off;
    __device__ void operator()(uint32_t i, unsigned long long exclusive, unsigned long long) const { off[i] = exclusive; }
};
// ... minus the pixels the part's boundary walk already wrote (FillWriter, writers.rs:39-60)
struct InSpanKept {
    InSpanLen base;
    VisitSet vs;
    KParams P;
    const CacheBox* box;
    uint32_t row_bits;
    __device__ unsigned long long operator()(uint32_t i) const {
        const uint32_t len = (uint32_t)base(i);
        if (!len) return 0ull;
        const uint64_t k = base.keys[i];  // [part | row | col]
        const uint32_t part = (uint32_t)(k >> (base.col_bits + row_bits));
        const uint32_t row = (uint32_t)(k >> base.col_bits) & ((1u << row_bits) - 1u);
        const uint32_t col = (uint32_t)k & ((1u << base.col_bits) - 1u);
        const CacheBox b = box[part];
        unsigned long long c = 0;
        for (uint32_t j = 0; j < len; j++) c += !cache_contains(P.nrows, P.ncols, b, vs, part, row, col + j);
        return c;
    }
};
// pixel writes of line segment i
template <bool TOUCHED>
struct InLineLen {
    KParams P;
    const double* x;
    const double* y;
    const uint32_t* tag;
    const PartInfo* info;
    const uint32_t* last_kept;
    Counters* ctr;
    uint32_t n;
    __device__ unsigned long long operator()(uint32_t i) const {
        unsigned long long c = 0;
        for_each_line_pixel<TOUCHED>(P, x, y, tag, info, last_kept, ctr, i, n,
                                     [&](uint32_t, uint32_t, uint32_t, uint32_t) { c++; });
        return c;
    }
};
// ... of which first visits of their pixel inside the part (non-square pixels)
template <bool TOUCHED>
struct InLineKept {
    InLineLen<TOUCHED> base;
    const unsigned long long* raw_off;
    VisitSet vs;
    __device__ unsigned long long operator()(uint32_t i) const {
        unsigned long long c = 0;
        const unsigned long long b0 = raw_off[i];
        for_each_line_pixel<TOUCHED>(base.P, base.x, base.y, base.tag, base.info, base.last_kept, base.ctr, i, base.n,
                                     [&](uint32_t part, uint32_t row, uint32_t col, uint32_t j) {
                                         c += vs.is_first(vs.key_of(part, row, col), b0 + j);
                                     });
        return c;
    }
};
struct InPointHit {
    KParams P;
    const double* x;
    const double* y;
    const uint32_t* tag;
    const PartInfo* info;
    __device__ unsigned long long operator()(uint32_t i) const {
        if (info[tag[i] & 0x3fffffffu].band < 0) return 0ull;
        double px = px_x(P, x[i]), py = px_y(P, y[i]);
        return (px >= 0.0 && px < P.ncols_f && py >= 0.0 && py < P.nrows_f) ? 1ull : 0ull;
    }
};

static __global__ void line_last_kept_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                      const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                                      uint32_t* __restrict__ last_kept, Counters* __restrict__ ctr) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    LineRec l;
    bool kept;
    line_setup(P, x, y, tag, info, i, n, l, &kept, ctr);
    if (kept) atomicMax(&last_kept[l.part], i + 1);
}

// ---------------------------------------------------------------------------------------------
// per-part write counts and bases
// ---------------------------------------------------------------------------------------------
// rec_beg[p] = first sorted crossing of polygon part p (lower bound on the part field); p in 0..n_parts
static __global__ void part_rec_range_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t part_shift,
                                      uint32_t n_parts, uint32_t* __restrict__ rec_beg) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_parts) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if ((keys[mid] >> part_shift) < (uint64_t)p) lo = mid + 1;
        else hi = mid;
    }
    rec_beg[p] = lo;
}

// count[p] = number of triplets part p writes; start[p] = value of its stream's prefix at the part's
// first unit (so that unit offset inside the part = prefix - start[p])
static __global__ void part_count_kernel(uint32_t n_parts, const uint8_t* __restrict__ part_kind,
                                  const uint32_t* __restrict__ vbeg, const uint32_t* __restrict__ vend,
                                  const uint32_t* __restrict__ rec_beg, const unsigned long long* __restrict__ poly_off,
                                  uint32_t n_rec, unsigned long long poly_total,
                                  const unsigned long long* __restrict__ line_off, uint32_t n_line,
                                  unsigned long long line_total, const unsigned long long* __restrict__ pt_off,
                                  uint32_t n_pt, unsigned long long pt_total,
                                  const unsigned long long* __restrict__ walk_off, uint32_t n_poly_v,
                                  unsigned long long walk_total, unsigned long long* __restrict__ count,
                                  unsigned long long* __restrict__ start, unsigned long long* __restrict__ walk_start) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_parts) return;
    unsigned long long a = 0, b = 0;
    const uint32_t k = part_kind[p];
    if (k == 0) {
        const uint32_t i0 = rec_beg[p], i1 = rec_beg[p + 1];
        a = i0 < n_rec ? poly_off[i0] : poly_total;
        b = i1 < n_rec ? poly_off[i1] : poly_total;
        if (walk_off) {  // all_touched: the part's boundary walk is written before its fill
            const unsigned long long wa = vbeg[p] < n_poly_v ? walk_off[vbeg[p]] : walk_total;
            const unsigned long long wb = vend[p] < n_poly_v ? walk_off[vend[p]] : walk_total;
            walk_start[p] = wa;
            count[p] = (b - a) + (wb - wa);
            start[p] = a - (wb - wa);  // modulo 2^64: fill offset inside the part = walk count + (prefix - a)
            return;
        }
    } else if (k == 1) {
        a = vbeg[p] < n_line ? line_off[vbeg[p]] : line_total;
        b = vend[p] < n_line ? line_off[vend[p]] : line_total;
    } else {
        a = vbeg[p] < n_pt ? pt_off[vbeg[p]] : pt_total;
        b = vend[p] < n_pt ? pt_off[vend[p]] : pt_total;
    }
    count[p] = b - a;
    start[p] = a;
}

// value of part p for the band-major ordering scan of band `band`
struct InBandCount {
    const unsigned long long* count;
    const PartInfo* info;
    int32_t band;
    __device__ unsigned long long operator()(uint32_t p) const { return info[p].band == band ? count[p] : 0ull; }
};
struct OutBandBase {
    unsigned long long* base;
    const PartInfo* info;
    int32_t band;
    unsigned long long band_base;
    __device__ void operator()(uint32_t p, unsigned long long exclusive, unsigned long long) const {
        if (info[p].band == band) base[p] = band_base + exclusive;
    }
};

// ---------------------------------------------------------------------------------------------
// expand: write the triplets at their final positions
// ---------------------------------------------------------------------------------------------
template <typename N>
static __global__ void __launch_bounds__(256)
poly_expand_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ seg_start,
                   const unsigned long long* __restrict__ poly_off, uint32_t n, SparseLayout L,
                   const PartInfo* __restrict__ info, const unsigned long long* __restrict__ part_base,
                   const unsigned long long* __restrict__ part_start, unsigned long long* __restrict__ rows,
                   unsigned long long* __restrict__ cols, N* __restrict__ data) {
    __shared__ uint32_t s_pre[256], s_row[256], s_col[256];
    __shared__ unsigned long long s_dst[256], s_val[256];
    const uint32_t lane = lane_id(), w0 = threadIdx.x & ~31u;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    uint32_t len = 0;
    if (i < n) {
        InSpanLen f{keys, seg_start, n, L.col_bits};
        len = (uint32_t)f(i);
        if (len) {
            const uint64_t k = keys[i];
            const uint32_t part = (uint32_t)(k >> (L.col_bits + L.row_bits));
            s_row[threadIdx.x] = (uint32_t)(k >> L.col_bits) & ((1u << L.row_bits) - 1u);
            s_col[threadIdx.x] = (uint32_t)k & ((1u << L.col_bits) - 1u);
            s_dst[threadIdx.x] = part_base[part] + (poly_off[i] - part_start[part]);
            s_val[threadIdx.x] = info[part].value_bits;
        }
    }
    uint32_t inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += t;
    }
    const uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
    s_pre[threadIdx.x] = inc - len;
    __syncwarp();
    for (uint32_t t = lane; t < wtot; t += 32) {
        uint32_t lo = 0, hi = 32;
#pragma unroll
        for (int it = 0; it < 5; it++) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_pre[w0 + mid] <= t) lo = mid;
            else hi = mid;
        }
        const uint32_t j = w0 + lo, k = t - s_pre[j];
        const unsigned long long d = s_dst[j] + k;
        rows[d] = s_row[j];
        cols[d] = s_col[j] + k;
        data[d] = value_from_bits<N>(s_val[j]);
    }
}

// all_touched with sum / count: a fill pixel is written unless the part's boundary walk visited it; one
// thread per span writes the kept pixels one after another
template <typename N>
static __global__ void __launch_bounds__(256)
poly_expand_dedup_kernel(KParams P, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ seg_start,
                         const unsigned long long* __restrict__ poly_off, uint32_t n, SparseLayout L, VisitSet vs,
                         const CacheBox* __restrict__ box, const PartInfo* __restrict__ info, const unsigned long long* __restrict__ part_base,
                         const unsigned long long* __restrict__ part_start, unsigned long long* __restrict__ rows,
                         unsigned long long* __restrict__ cols, N* __restrict__ data) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    InSpanLen f{keys, seg_start, n, L.col_bits};
    const uint32_t len = (uint32_t)f(i);
    if (!len) return;
    const uint64_t k = keys[i];
    const uint32_t part = (uint32_t)(k >> (L.col_bits + L.row_bits));
    const uint32_t row = (uint32_t)(k >> L.col_bits) & ((1u << L.row_bits) - 1u);
    const uint32_t col = (uint32_t)k & ((1u << L.col_bits) - 1u);
    const N v = value_from_bits<N>(info[part].value_bits);
    unsigned long long d = part_base[part] + (poly_off[i] - part_start[part]);
    const CacheBox b = box[part];
    for (uint32_t j = 0; j < len; j++) {
        if (cache_contains(P.nrows, P.ncols, b, vs, part, row, col + j)) continue;  // FillWriter (writers.rs:49-54)
        rows[d] = row;
        cols[d] = col + j;
        data[d] = v;
        d++;
    }
}

template <typename N, bool TOUCHED>
static __global__ void __launch_bounds__(256)
line_expand_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                   const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                   const uint32_t* __restrict__ last_kept, Counters* __restrict__ ctr,
                   const unsigned long long* __restrict__ line_off, const unsigned long long* __restrict__ raw_off,
                   VisitSet vs, int dedup, const unsigned long long* __restrict__ part_base,
                   const unsigned long long* __restrict__ part_start, unsigned long long* __restrict__ rows,
                   unsigned long long* __restrict__ cols, N* __restrict__ data) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    unsigned long long d = 0;
    bool have_d = false;
    const unsigned long long b0 = dedup ? raw_off[i] : 0ull;
    for_each_line_pixel<TOUCHED>(P, x, y, tag, info, last_kept, ctr, i, n, [&](uint32_t part, uint32_t row, uint32_t col, uint32_t j) {
        if (!have_d) {
            d = part_base[part] + (line_off[i] - part_start[part]);
            have_d = true;
        }
        if (dedup && !vs.is_first(vs.key_of(part, row, col), b0 + j)) return;
        rows[d] = row;
        cols[d] = col;
        data[d] = value_from_bits<N>(info[part].value_bits);
        d++;
    });
}

template <typename N>
static __global__ void point_expand_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                    const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                                    const unsigned long long* __restrict__ pt_off,
                                    const unsigned long long* __restrict__ part_base,
                                    const unsigned long long* __restrict__ part_start,
                                    unsigned long long* __restrict__ rows, unsigned long long* __restrict__ cols,
                                    N* __restrict__ data) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t part = tag[i] & 0x3fffffffu;
    const PartInfo pi = info[part];
    if (pi.band < 0) return;
    double px = px_x(P, x[i]), py = px_y(P, y[i]);
    if (!(px >= 0.0 && px < P.ncols_f && py >= 0.0 && py < P.nrows_f)) return;
    const unsigned long long d = part_base[part] + (pt_off[i] - part_start[part]);
    rows[d] = (unsigned long long)(uint32_t)py;
    cols[d] = (unsigned long long)(uint32_t)px;
    data[d] = value_from_bits<N>(pi.value_bits);
}

}  // namespace rz

// ---------------------------------------------------------------------------------------------
// replay: SparseArray::build_array (rust/src/encoding/arrays.rs:103-143)
// ---------------------------------------------------------------------------------------------
// Every triplet becomes a record [task | index]; a stable sort on the task bits keeps the triplets of
// a (band,row,tile) task in their original (burn) order, and one warp replays them through the
// pixel function on a shared-memory row tile.
namespace rz {

static __global__ void replay_emit_kernel(const unsigned long long* __restrict__ rows, const unsigned long long* __restrict__ cols,
                                   uint32_t n, const unsigned long long* __restrict__ band_off, uint32_t n_bands,
                                   uint32_t nrows, uint32_t ncols, uint32_t n_tiles, uint32_t tile_shift,
                                   uint32_t idx_bits, uint32_t n_tasks, uint64_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t band = 0;
    while (band + 1 < n_bands && (unsigned long long)i >= band_off[band + 1]) band++;
    const unsigned long long r = rows[i], c = cols[i];
    // out-of-range triplets (the reference would panic on the index) are parked in a task nobody replays
    uint64_t task = n_tasks;
    if (r < nrows && c < ncols) task = ((uint64_t)band * nrows + r) * n_tiles + (uint32_t)(c >> tile_shift);
    keys[i] = (task << idx_bits) | i;
}

template <typename N, int FN>
static __global__ void __launch_bounds__(FILL_WARPS * 32)
replay_fill_kernel(FillParams F, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ task_start,
                   const unsigned long long* __restrict__ cols, const N* __restrict__ data, uint32_t idx_bits,
                   uint64_t bg_bits, N* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    N* row = reinterpret_cast<N*>(smem_raw) + (size_t)warp * FILL_MAX_TILE_W;
    const N bg = value_from_bits<N>(bg_bits);
    const uint64_t idx_mask = (1ull << idx_bits) - 1ull;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t task = blockIdx.x * FILL_WARPS + warp; task < F.n_tasks; task += gridDim.x * FILL_WARPS) {
        const uint32_t tile = task % F.n_tiles;
        const uint32_t br = task / F.n_tiles;
        const uint32_t c0 = tile * F.tile_w;
        const uint32_t w = min(F.tile_w, F.ncols - c0);
        N* dst = out + (size_t)br * F.ncols + c0;
        const uint32_t beg = task_start[task], end = task_start[task + 1];
        for (uint32_t i = lane; i < w; i += 32) row[i] = bg;
        __syncwarp();
        for (uint32_t base = beg; base < end; base += 32) {
            const bool valid = base + lane < end;
            uint32_t col = 0xffffffffu - lane;  // distinct dummies for idle lanes
            N v = bg;
            if (valid) {
                const uint32_t idx = (uint32_t)(keys[base + lane] & idx_mask);
                col = (uint32_t)cols[idx] - c0;
                v = data[idx];
            }
            // records hitting the same pixel are applied in record order, one per round
            const uint32_t peers = __match_any_sync(0xffffffffu, col);
            const uint32_t rank = __popc(peers & lt_mask);
            const uint32_t rounds = __reduce_max_sync(0xffffffffu, valid ? rank : 0u);
            for (uint32_t r = 0; r <= rounds; r++) {
                if (valid && rank == r) row[col] = apply_px<N, FN>(row[col], v, bg);
                __syncwarp();
            }
        }
        for (uint32_t i = lane; i < w; i += 32) dst[i] = row[i];
        __syncwarp();
    }
}

}  // namespace rz
