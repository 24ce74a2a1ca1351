// rz_burn.cuh — line and point pixels written straight onto the raster (order-free pixel functions).
//
// A mixed job (polygons + lines + points) normally takes the crossing-record pipeline, because the reference
// applies geometries in input order and a line may be burned between two polygons (rasterize.rs:162-196).  For the
// pixel functions whose result does not depend on that order the job is split instead: polygon parts go through the
// tile-binned engine (rz_tiles.cuh), then every line / point pixel is applied to the finished raster by one atomic
// (or plain store) per write - no records, no sort, no row-tile fill:
//
//   any                               cur = 1 whatever came before (pixel_functions.rs:118-123): a plain store;
//   count, integer dtype, bg == 0     `cur == bg ? 1 : cur + 1` is cur + 1 (wrapping) for every cur: atomic add 1;
//   sum,   integer dtype, bg == 0     `cur == bg ? v : cur + v` is cur + v (wrapping): atomic add v.
//
// Wrapping integer addition is commutative and associative, so any interleaving gives the reference's value bit for
// bit; revisited pixels of one line part are added once per visit, exactly like the Bresenham loop writes them again
// (burners.rs:60-89).  Not eligible: float dtypes (rounding depends on the order), other functions, a non-zero
// background, non-square pixels or all_touched (per-part PixelCache, writers.rs:25-29).
#pragma once

#include "rz_kernels.cuh"

namespace rz {

struct BurnTarget {
    void* out;               // [band][out_rows][ncols] of the item size
    uint32_t out_rows;       // rows per band in `out`
    uint32_t win_row_off;    // first window row relative to the output's first row
    unsigned long long one;  // bit pattern of the dtype's 1 (any / count)
    int use_part_value;      // sum: add the part's value instead of `one`
};

// SZ-byte wrapping add / store at element index `idx`
template <int SZ, bool ADD> __device__ __forceinline__ void burn_apply(void* base, size_t idx, unsigned long long v) {
    if (SZ == 8) {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(base) + idx;
        if (ADD) atomicAdd(p, v);
        else *p = v;
    } else if (SZ == 4) {
        unsigned int* p = reinterpret_cast<unsigned int*>(base) + idx;
        if (ADD) atomicAdd(p, (unsigned int)v);
        else *p = (unsigned int)v;
    } else if (!ADD) {
        if (SZ == 2) reinterpret_cast<unsigned short*>(base)[idx] = (unsigned short)v;
        else reinterpret_cast<unsigned char*>(base)[idx] = (unsigned char)v;
    } else {  // 1- and 2-byte adds: compare-and-swap on the aligned 32-bit word holding the element
        const uintptr_t a = (uintptr_t)base + idx * SZ;
        unsigned int* w = reinterpret_cast<unsigned int*>(a & ~(uintptr_t)3);
        const unsigned int shift = (unsigned int)(a & 3u) * 8u;
        const unsigned int field = (SZ == 2 ? 0xffffu : 0xffu) << shift;
        unsigned int old = *w, assumed;
        do {
            assumed = old;
            const unsigned int sum = ((assumed >> shift) + (unsigned int)v) << shift;
            old = atomicCAS(w, assumed, (assumed & ~field) | (sum & field));
        } while (old != assumed);
    }
}

template <int SZ, bool ADD>
__device__ __forceinline__ void burn_pixel(const KParams& P, const BurnTarget& B, int32_t band, long long row, long long col,
                                           unsigned long long v) {
    const size_t idx = ((size_t)band * B.out_rows + B.win_row_off + ((uint32_t)row - P.win_r0)) * P.ncols + (uint32_t)col;
    burn_apply<SZ, ADD>(B.out, idx, v);
}

// one thread per line-pool vertex = segment (i, i+1); segments longer than LONG_EDGE pixels are shared by the warp
template <int SZ, bool ADD>
static __global__ void __launch_bounds__(256)
line_burn_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y, const uint32_t* __restrict__ tag,
                 uint32_t n, const PartInfo* __restrict__ info, Counters* __restrict__ ctr, BurnTarget B) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    LineRec l;
    bool kept;
    line_setup(P, x, y, tag, info, i, n, l, &kept, ctr);
    const unsigned long long v = (l.n && B.use_part_value) ? info[l.part].value_bits : B.one;
    const bool is_long = l.n > LONG_EDGE;
    if (!is_long && l.n) {
        // sequential walk: the minor coordinate q(k) = floor((2 dmin k + dmaj) / (2 dmaj)) of line_pixel() is carried
        // along with its remainder instead of being divided out per pixel (a 64-bit division costs more than the
        // atomic it feeds)
        long long num = 2 * l.dmin * (long long)l.k_lo + l.dmaj;
        const long long den = 2 * l.dmaj;
        long long q = den > 0 ? num / den : 0;
        long long rem = den > 0 ? num - q * den : 0;
        const long long step = 2 * l.dmin;  // step <= den, so q advances by at most one per pixel
        long long maj = (l.xmajor ? l.ix0 : l.iy0) + (l.xmajor ? l.sx : l.sy) * (long long)l.k_lo;
        const long long min0 = l.xmajor ? l.iy0 : l.ix0;
        const int smaj = l.xmajor ? l.sx : l.sy, smin = l.xmajor ? l.sy : l.sx;
        for (uint32_t k = 0; k < l.n; k++) {
            const long long mn = min0 + smin * q;
            burn_pixel<SZ, ADD>(P, B, l.band, l.xmajor ? mn : maj, l.xmajor ? maj : mn, v);
            maj += smaj;
            rem += step;
            if (rem >= den) {
                rem -= den;
                q++;
            }
        }
    }
    uint32_t m = __ballot_sync(0xffffffffu, is_long);
    const uint32_t lane = lane_id();
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        LineRec b;
        b.ix0 = __shfl_sync(0xffffffffu, l.ix0, src);
        b.iy0 = __shfl_sync(0xffffffffu, l.iy0, src);
        b.dmaj = __shfl_sync(0xffffffffu, l.dmaj, src);
        b.dmin = __shfl_sync(0xffffffffu, l.dmin, src);
        b.sx = __shfl_sync(0xffffffffu, l.sx, src);
        b.sy = __shfl_sync(0xffffffffu, l.sy, src);
        b.xmajor = __shfl_sync(0xffffffffu, l.xmajor, src);
        b.k_lo = __shfl_sync(0xffffffffu, l.k_lo, src);
        b.n = __shfl_sync(0xffffffffu, l.n, src);
        b.band = __shfl_sync(0xffffffffu, l.band, src);
        const unsigned long long bv = __shfl_sync(0xffffffffu, v, src);
        for (uint32_t k = lane; k < b.n; k += 32) {
            long long px, py;
            line_pixel(b, (long long)(b.k_lo + k), px, py);
            burn_pixel<SZ, ADD>(P, B, b.band, py, px, bv);
        }
    }
}

// the end pixel of the last kept segment of every line part whose line string is open (burners.rs:87-89)
template <int SZ, bool ADD>
static __global__ void line_final_burn_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                              const uint32_t* __restrict__ tag, const uint8_t* __restrict__ part_kind,
                                              const PartInfo* __restrict__ info, const uint32_t* __restrict__ last_kept,
                                              BurnTarget B) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_parts || part_kind[p] != 1) return;
    const uint32_t lk = last_kept[p];
    if (lk == 0) return;
    const uint32_t i = lk - 1;
    if (tag[i] & 0x40000000u) return;  // closed line string
    const PartInfo pi = info[p];
    if (pi.band < 0) return;
    const long long ix1 = sat_i64(floor(px_x(P, x[i + 1]))), iy1 = sat_i64(floor(px_y(P, y[i + 1])));
    if (ix1 < 0 || ix1 >= (long long)P.ncols || iy1 < (long long)P.win_r0 || iy1 >= (long long)P.win_r1) return;
    burn_pixel<SZ, ADD>(P, B, pi.band, iy1, ix1, B.use_part_value ? pi.value_bits : B.one);
}

// points: edges.rs:79-88, burners.rs:250-258 (every member point is written, duplicates included)
template <int SZ, bool ADD>
static __global__ void point_burn_kernel(KParams P, const double* __restrict__ x, const double* __restrict__ y,
                                         const uint32_t* __restrict__ tag, uint32_t n, const PartInfo* __restrict__ info,
                                         BurnTarget B) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t part = tag[i] & 0x3fffffffu;
    const PartInfo pi = info[part];
    const double px = px_x(P, x[i]), py = px_y(P, y[i]);
    if (!(pi.band >= 0 && px >= 0.0 && px < P.ncols_f && py >= 0.0 && py < P.nrows_f)) return;
    const uint32_t col = (uint32_t)px, row = (uint32_t)py;  // `as usize` of an in-range value truncates
    if (row < P.win_r0 || row >= P.win_r1) return;
    burn_pixel<SZ, ADD>(P, B, pi.band, row, col, B.use_part_value ? pi.value_bits : B.one);
}

}  // namespace rz
