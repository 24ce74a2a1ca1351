// rz_oracle.cpp — CPU restatement of rusterize's burn path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for the CUDA path in rusterize_b200/.  It may be loaded only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
// (rusterize_b200/*) never links, imports or calls it.
//
// It restates, function by function, the algorithm of the reference (paths relative to
// /root/reference, which cannot be compiled here: no cargo/rustc in the image):
//   rust/src/geo/edges.rs                       -> world_to_pixel, extract_point/ring/line, PolyEdge
//   rust/src/rasterization/burners.rs           -> burn_line_standard, burn_line_all_touched,
//                                                  burn_point, burn_polygon
//   rust/src/rasterization/burn_geometry.rs     -> burn_geometry (type dispatch, pooling, GC recursion)
//   rust/src/rasterization/pixel_functions.rs   -> px_sum/first/last/min/max/count/any
//   rust/src/rasterization/pixel_cache.rs       -> PixelCache
//   rust/src/encoding/writers.rs, arrays.rs     -> DenseWriter, SparseWriter, LineWriter, FillWriter,
//                                                  sparse replay (build_array)
//   rust/src/rasterize.rs                       -> process order, group_keys, length checks
//   rust/src/geo/raster.rs                      -> raster-info builder (finalize)
// Pinning: tests/test_oracle_golden.py checks it against the reference's own golden GeoTIFFs
// (python/test/data/*.tif), the documented sparse frame (python/docs/python.md:106-136), the R
// known-answer matrices (R/rusterize/tests/testthat/*.R) and the Rust unit known answers.
//
// Build: g++ -O3 -march=native -ffp-contract=off -shared -fPIC (see oracle/Makefile).
// -ffp-contract=off matters: Rust never fuses a*b+c, so neither may we (edges.rs:50-55).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------------------------
// Rust cast semantics
// ----------------------------------------------------------------------------------------------
// `f64 as usize`: truncate toward zero, saturate, NaN -> 0.
inline uint64_t as_usize(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)v;
}
// `f64 as isize`
inline int64_t as_isize(double v) {
    if (!(v == v)) return 0;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    if (v >= 9223372036854775807.0) return INT64_MAX;
    return (int64_t)v;
}
// f64::total_cmp as a strict-weak "less"
inline bool total_less(double a, double b) {
    int64_t x, y;
    std::memcpy(&x, &a, 8);
    std::memcpy(&y, &b, 8);
    x ^= (int64_t)(((uint64_t)(x >> 63)) >> 1);
    y ^= (int64_t)(((uint64_t)(y >> 63)) >> 1);
    return x < y;
}
// f64::clamp (NaN stays NaN)
inline double rs_clamp(double v, double lo, double hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
// f64::min / f64::max (ignore NaN operand)
inline double rs_min(double a, double b) { return std::fmin(a, b); }
inline double rs_max(double a, double b) { return std::fmax(a, b); }

// ----------------------------------------------------------------------------------------------
// Geometry tree (what geo_types::Geometry<f64> holds after WKB decoding)
// ----------------------------------------------------------------------------------------------
struct Coord {
    double x, y;
};
typedef std::vector<Coord> LineStr;
struct Poly {
    std::vector<LineStr> rings;  // rings[0] exterior; all closed on construction (geo_types Polygon::new)
};
enum GType { G_POINT = 1, G_LINE = 2, G_POLY = 3, G_MPOINT = 4, G_MLINE = 5, G_MPOLY = 6, G_COLL = 7 };
struct Geom {
    int type = 0;
    std::vector<Coord> points;   // POINT (1) / MULTIPOINT
    std::vector<LineStr> lines;  // LINESTRING (1) / MULTILINESTRING
    std::vector<Poly> polys;     // POLYGON (1) / MULTIPOLYGON
    std::vector<Geom> members;   // GEOMETRYCOLLECTION
};

// geo_types LineString::close(): push first coord when first != last.
void close_ring(LineStr& r) {
    if (r.empty()) return;
    const Coord &a = r.front(), &b = r.back();
    if (!(a.x == b.x && a.y == b.y)) r.push_back(a);
}

struct WkbCursor {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    uint8_t u8() {
        if (p + 1 > end) { ok = false; return 0; }
        return *p++;
    }
    uint32_t u32(bool le) {
        if (p + 4 > end) { ok = false; return 0; }
        uint32_t v;
        if (le) v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        else v = (uint32_t)p[3] | ((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24);
        p += 4;
        return v;
    }
    double f64(bool le) {
        if (p + 8 > end) { ok = false; return 0; }
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v |= (uint64_t)p[le ? i : 7 - i] << (8 * i);
        p += 8;
        double d;
        std::memcpy(&d, &v, 8);
        return d;
    }
};

// Decode one (possibly nested) WKB geometry.  Returns false when the geometry has no geo_types
// equivalent (empty point), mirroring `try_to_geometry() -> None` (python/src/geo/parse_geometry.rs:77-84).
bool parse_wkb(WkbCursor& c, Geom& g) {
    bool le = c.u8() == 1;
    uint32_t t = c.u32(le);
    int dims = 2;
    if (t & 0x80000000u) dims++;               // EWKB Z
    if (t & 0x40000000u) dims++;               // EWKB M
    bool srid = (t & 0x20000000u) != 0;        // EWKB SRID
    t &= 0x0fffffffu;
    uint32_t iso = t / 1000;                   // ISO: 1000 Z, 2000 M, 3000 ZM
    if (iso == 1 || iso == 2) dims = 3;
    if (iso == 3) dims = 4;
    t %= 1000;
    if (srid) c.u32(le);
    auto coord = [&](Coord& o) {
        o.x = c.f64(le);
        o.y = c.f64(le);
        for (int k = 2; k < dims; k++) c.f64(le);
    };
    auto line = [&](LineStr& l) {
        uint32_t n = c.u32(le);
        for (uint32_t i = 0; i < n && c.ok; i++) {
            Coord q;
            coord(q);
            l.push_back(q);
        }
    };
    auto poly = [&](Poly& p) {
        uint32_t nr = c.u32(le);
        for (uint32_t i = 0; i < nr && c.ok; i++) {
            p.rings.emplace_back();
            line(p.rings.back());
            close_ring(p.rings.back());
        }
        if (p.rings.empty()) p.rings.emplace_back();  // empty exterior
    };
    g.type = (int)t;
    switch (t) {
        case G_POINT: {
            Coord q;
            coord(q);
            if (q.x != q.x && q.y != q.y) return false;  // POINT EMPTY
            g.points.push_back(q);
            return c.ok;
        }
        case G_LINE:
            g.lines.emplace_back();
            line(g.lines.back());
            return c.ok;
        case G_POLY:
            g.polys.emplace_back();
            poly(g.polys.back());
            return c.ok;
        case G_MPOINT:
        case G_MLINE:
        case G_MPOLY:
        case G_COLL: {
            uint32_t n = c.u32(le);
            for (uint32_t i = 0; i < n && c.ok; i++) {
                Geom m;
                bool keep = parse_wkb(c, m);
                if (!c.ok) return false;
                if (t == G_COLL) {
                    if (keep) g.members.push_back(std::move(m));
                } else if (t == G_MPOINT) {
                    if (keep) g.points.push_back(m.points[0]);
                } else if (t == G_MLINE) {
                    g.lines.push_back(std::move(m.lines[0]));
                } else {
                    g.polys.push_back(std::move(m.polys[0]));
                }
            }
            return c.ok;
        }
        default:
            c.ok = false;
            return false;
    }
}

// ----------------------------------------------------------------------------------------------
// Raster info — rust/src/geo/raster.rs:10-20
// ----------------------------------------------------------------------------------------------
struct RasterInfo {
    uint64_t ncols, nrows;
    double xmin, xmax, ymin, ymax, xres, yres;
};

// ----------------------------------------------------------------------------------------------
// Edges — rust/src/geo/edges.rs
// ----------------------------------------------------------------------------------------------
struct PointEdge {
    uint64_t x, y;
};
struct PolyEdge {  // edges.rs:17-46
    uint64_t ystart, yend;
    double x0, y0, dxdy, x_at_yline;
    PolyEdge(double ax, double ay, double bx, double by) {
        double x_top, y_top, x_bot, y_bot;
        if (ay < by) { x_top = ax; y_top = ay; x_bot = bx; y_bot = by; }
        else { x_top = bx; y_top = by; x_bot = ax; y_bot = ay; }
        ystart = as_usize(std::ceil(y_top - 0.5));
        yend = as_usize(std::ceil(y_bot - 0.5));
        dxdy = (x_bot - x_top) / (y_bot - y_top);
        x0 = x_top;
        y0 = y_top;
        x_at_yline = std::numeric_limits<double>::infinity();
    }
    // edges.rs:50-55 — one subtract, one multiply, one add; never fused.
    double intersect_at(uint64_t yline) const {
        double center_y = (double)yline + 0.5;
        return x0 + (center_y - y0) * dxdy;
    }
};
struct LineEdge {
    double x0, y0, x1, y1;
    bool is_closed;
};

// edges.rs:79-88
void extract_point(std::vector<PointEdge>& out, const Coord& p, const RasterInfo& ri) {
    double x = (p.x - ri.xmin) / ri.xres;
    double y = (ri.ymax - p.y) / ri.yres;
    if (x >= 0.0 && x < (double)ri.ncols && y >= 0.0 && y < (double)ri.nrows) out.push_back({as_usize(x), as_usize(y)});
}
// edges.rs:90-110
void extract_ring(std::vector<PolyEdge>& out, const LineStr& l, const RasterInfo& ri) {
    double rows = (double)ri.nrows;
    for (size_t i = 0; i + 1 < l.size(); i++) {
        double x0 = (l[i].x - ri.xmin) / ri.xres;
        double y0 = (ri.ymax - l[i].y) / ri.yres;
        double x1 = (l[i + 1].x - ri.xmin) / ri.xres;
        double y1 = (ri.ymax - l[i + 1].y) / ri.yres;
        if (std::fabs(y0 - y1) >= std::numeric_limits<double>::epsilon()) {
            double min_y = rs_min(y0, y1), max_y = rs_max(y0, y1);
            if (min_y < rows && max_y >= 0.0) out.emplace_back(x0, y0, x1, y1);
        }
    }
}
// edges.rs:112-134
void extract_line(std::vector<LineEdge>& out, const LineStr& l, const RasterInfo& ri) {
    double rows = (double)ri.nrows, cols = (double)ri.ncols;
    bool closed = l.empty() || (l.front().x == l.back().x && l.front().y == l.back().y);
    for (size_t i = 0; i + 1 < l.size(); i++) {
        double x0 = (l[i].x - ri.xmin) / ri.xres;
        double y0 = (ri.ymax - l[i].y) / ri.yres;
        double x1 = (l[i + 1].x - ri.xmin) / ri.xres;
        double y1 = (ri.ymax - l[i + 1].y) / ri.yres;
        double min_x = rs_min(x0, x1), max_x = rs_max(x0, x1);
        double min_y = rs_min(y0, y1), max_y = rs_max(y0, y1);
        if (min_x < cols && max_x >= 0.0 && min_y < rows && max_y >= 0.0) out.push_back({x0, y0, x1, y1, closed});
    }
}

// ----------------------------------------------------------------------------------------------
// Pixel cache — rust/src/rasterization/pixel_cache.rs
// ----------------------------------------------------------------------------------------------
struct PixelCache {
    std::vector<uint64_t> bits;
    uint64_t width;
    int64_t xmin, ymin;
    explicit PixelCache(const std::vector<LineEdge>& edges) {
        double x_lo = std::numeric_limits<double>::max(), y_lo = x_lo;
        double x_hi = std::numeric_limits<double>::lowest(), y_hi = x_hi;
        for (const auto& e : edges) {
            x_lo = rs_min(rs_min(x_lo, e.x0), e.x1);
            y_lo = rs_min(rs_min(y_lo, e.y0), e.y1);
            x_hi = rs_max(rs_max(x_hi, e.x0), e.x1);
            y_hi = rs_max(rs_max(y_hi, e.y0), e.y1);
        }
        width = as_usize(std::floor(x_hi) - std::floor(x_lo)) + 1;
        uint64_t length = as_usize(std::floor(y_hi) - std::floor(y_lo)) + 1;
        bits.assign((width * length + 63) / 64, 0);
        xmin = as_isize(x_lo);
        ymin = as_isize(y_lo);
    }
    uint64_t index(uint64_t x, uint64_t y) const {
        uint64_t lx = (uint64_t)((int64_t)x - xmin), ly = (uint64_t)((int64_t)y - ymin);
        return ly * width + lx;
    }
    bool contains(uint64_t x, uint64_t y) const {
        uint64_t i = index(x, y);
        return i / 64 < bits.size() && ((bits[i / 64] >> (i % 64)) & 1);
    }
    bool insert(uint64_t x, uint64_t y) {
        uint64_t i = index(x, y);
        if (i / 64 >= bits.size()) bits.resize(i / 64 + 1, 0);  // fixedbitset would panic; never hit for in-raster pixels
        if ((bits[i / 64] >> (i % 64)) & 1) return false;
        bits[i / 64] |= 1ull << (i % 64);
        return true;
    }
};

// ----------------------------------------------------------------------------------------------
// Pixel functions — rust/src/rasterization/pixel_functions.rs:56-123
// ----------------------------------------------------------------------------------------------
template <typename N> inline bool is_nan(N) { return false; }
template <> inline bool is_nan<float>(float v) { return v != v; }
template <> inline bool is_nan<double>(double v) { return v != v; }

// Rust release `+=` wraps for integers.
template <typename N> inline N wrap_add(N a, N b) { return a + b; }
#define RZO_WRAP(T, U) \
    template <> inline T wrap_add<T>(T a, T b) { return (T)((U)a + (U)b); }
RZO_WRAP(int8_t, uint8_t)
RZO_WRAP(int16_t, uint16_t)
RZO_WRAP(int32_t, uint32_t)
RZO_WRAP(int64_t, uint64_t)
RZO_WRAP(uint8_t, uint8_t)
RZO_WRAP(uint16_t, uint16_t)
#undef RZO_WRAP

template <typename N> struct Band {
    N* data;
    uint64_t ncols;
    N& at(uint64_t y, uint64_t x) { return data[y * ncols + x]; }
};
template <typename N> using PixelFn = void (*)(Band<N>&, uint64_t, uint64_t, N, N);

template <typename N> void px_sum(Band<N>& a, uint64_t y, uint64_t x, N v, N bg) {
    N& c = a.at(y, x);
    if (c == bg || is_nan(c) || is_nan(v)) c = v;
    else c = wrap_add(c, v);
}
template <typename N> void px_first(Band<N>& a, uint64_t y, uint64_t x, N v, N bg) {
    N& c = a.at(y, x);
    if (c == bg || is_nan(c)) c = v;
}
template <typename N> void px_last(Band<N>& a, uint64_t y, uint64_t x, N v, N) { a.at(y, x) = v; }
template <typename N> void px_min(Band<N>& a, uint64_t y, uint64_t x, N v, N bg) {
    N& c = a.at(y, x);
    if (c == bg || is_nan(c) || c > v) c = v;
}
template <typename N> void px_max(Band<N>& a, uint64_t y, uint64_t x, N v, N bg) {
    N& c = a.at(y, x);
    if (c == bg || is_nan(c) || c < v) c = v;
}
template <typename N> void px_count(Band<N>& a, uint64_t y, uint64_t x, N, N bg) {
    N& c = a.at(y, x);
    if (c == bg || is_nan(c)) c = (N)1;
    else c = wrap_add(c, (N)1);
}
template <typename N> void px_any(Band<N>& a, uint64_t y, uint64_t x, N, N) { a.at(y, x) = (N)1; }

enum { FN_SUM = 0, FN_FIRST, FN_LAST, FN_MIN, FN_MAX, FN_COUNT, FN_ANY };
template <typename N> PixelFn<N> to_function(int fn) {
    switch (fn) {
        case FN_SUM: return px_sum<N>;
        case FN_FIRST: return px_first<N>;
        case FN_LAST: return px_last<N>;
        case FN_MIN: return px_min<N>;
        case FN_MAX: return px_max<N>;
        case FN_COUNT: return px_count<N>;
        default: return px_any<N>;
    }
}

// ----------------------------------------------------------------------------------------------
// Writers — rust/src/encoding/writers.rs
// ----------------------------------------------------------------------------------------------
template <typename N> struct DenseWriter {  // writers.rs:63-78
    Band<N> band;
    PixelFn<N> fn;
    void write(uint64_t y, uint64_t x, N v, N bg) { fn(band, y, x, v, bg); }
};
template <typename N> struct SparseWriter {  // writers.rs:86-99
    std::vector<uint64_t> rows, cols;
    std::vector<N> values;
    void write(uint64_t y, uint64_t x, N v, N) {
        rows.push_back(y);
        cols.push_back(x);
        values.push_back(v);
    }
};
template <typename N, typename W> struct LineWriter {  // writers.rs:15-36
    W& inner;
    PixelCache& cache;
    void write(uint64_t y, uint64_t x, N v, N bg) {
        if (cache.insert(x, y)) inner.write(y, x, v, bg);
    }
};
template <typename N, typename W> struct FillWriter {  // writers.rs:39-60
    W& inner;
    PixelCache& cache;
    void write(uint64_t y, uint64_t x, N v, N bg) {
        if (!cache.contains(x, y)) inner.write(y, x, v, bg);
    }
};

// ----------------------------------------------------------------------------------------------
// Burners — rust/src/rasterization/burners.rs
// ----------------------------------------------------------------------------------------------
// burners.rs:35-92
template <typename N, typename W>
void burn_line_standard(const std::vector<LineEdge>& edges, const RasterInfo& ri, N v, W& w, N bg) {
    if (edges.empty()) return;
    int64_t nrows = (int64_t)ri.nrows, ncols = (int64_t)ri.ncols;
    size_t last = edges.size() - 1;
    for (size_t idx = 0; idx < edges.size(); idx++) {
        const LineEdge& e = edges[idx];
        int64_t ix0 = as_isize(std::floor(e.x0)), ix1 = as_isize(std::floor(e.x1));
        int64_t iy0 = as_isize(std::floor(e.y0)), iy1 = as_isize(std::floor(e.y1));
        int64_t dx = std::llabs(ix1 - ix0), dy = -std::llabs(iy1 - iy0);
        int64_t sx = ix0 < ix1 ? 1 : -1, sy = iy0 < iy1 ? 1 : -1;
        int64_t err = dx + dy;
        while (ix0 != ix1 || iy0 != iy1) {
            if (ix0 >= 0 && ix0 < ncols && iy0 >= 0 && iy0 < nrows) w.write((uint64_t)iy0, (uint64_t)ix0, v, bg);
            int64_t e2 = 2 * err;
            if (e2 >= dy) { err += dy; ix0 += sx; }
            if (e2 <= dx) { err += dx; iy0 += sy; }
        }
        if (idx == last && !e.is_closed && ix0 >= 0 && ix0 < ncols && iy0 >= 0 && iy0 < nrows)
            w.write((uint64_t)iy0, (uint64_t)ix0, v, bg);
    }
}

// burners.rs:94-247 (GDAL-derived all-touched walk)
template <typename N, typename W>
void burn_line_all_touched(const std::vector<LineEdge>& edges, const RasterInfo& ri, N v, W& w, N bg) {
    const double EPS_INTERSECT = 1e-4, TOL = 1e-9;
    if (edges.empty()) return;
    int64_t nrows = (int64_t)ri.nrows, ncols = (int64_t)ri.ncols;
    double nrows_f = (double)ri.nrows, ncols_f = (double)ri.ncols;
    for (const LineEdge& e : edges) {
        double x = e.x0, y = e.y0, xe = e.x1, ye = e.y1;
        if (x > xe) { std::swap(x, xe); std::swap(y, ye); }
        if (std::fabs(x - xe) < 0.01) {  // vertical
            if (ye < y) std::swap(y, ye);
            int64_t ix = as_isize(std::floor(xe));
            int64_t iy = as_isize(std::floor(y));
            int64_t iy_end = as_isize(std::floor(ye - EPS_INTERSECT));
            if (ix < 0 || ix >= ncols) continue;
            iy = std::max<int64_t>(iy, 0);
            iy_end = std::min<int64_t>(iy_end, nrows - 1);
            for (int64_t yy = iy; yy <= iy_end; yy++) w.write((uint64_t)yy, (uint64_t)ix, v, bg);
            continue;
        }
        if (std::fabs(y - ye) < 0.01) {  // horizontal
            if (xe < x) std::swap(x, xe);
            int64_t ix = as_isize(std::floor(x));
            int64_t iy = as_isize(std::floor(y));
            int64_t ix_end = as_isize(std::floor(xe - EPS_INTERSECT));
            if (iy < 0 || iy >= nrows) continue;
            ix = std::max<int64_t>(ix, 0);
            ix_end = std::min<int64_t>(ix_end, ncols - 1);
            for (int64_t xx = ix; xx <= ix_end; xx++) w.write((uint64_t)iy, (uint64_t)xx, v, bg);
            continue;
        }
        double slope = (ye - y) / (xe - x);
        double inv_slope = 1.0 / slope;
        if (x < 0.0) { y += (0.0 - x) * slope; x = 0.0; }
        if (xe > ncols_f) { ye += (ncols_f - xe) * slope; xe = ncols_f; }
        if (y < 0.0) { x += (0.0 - y) * inv_slope; y = 0.0; }
        else if (y > nrows_f) { x += (nrows_f - y) * inv_slope; y = nrows_f; }
        if (ye < 0.0) xe += (0.0 - ye) * inv_slope;
        else if (ye > nrows_f) xe += (nrows_f - ye) * inv_slope;
        x = rs_clamp(x, 0.0, ncols_f);
        xe = rs_clamp(xe, 0.0, ncols_f);
        while (x >= 0.0 && x < xe) {
            int64_t ix = as_isize(std::floor(x)), iy = as_isize(std::floor(y));
            if (ix >= 0 && ix < ncols && iy >= 0 && iy < nrows) w.write((uint64_t)iy, (uint64_t)ix, v, bg);
            double sx = std::floor(x + 1.0) - x;
            double sy = sx * slope;
            if (as_isize(std::floor(y + sy)) == iy) {
                x += sx;
                y += sy;
            } else if (slope < 0.0) {
                sy = (double)iy - y;
                if (sy > -TOL) sy = -TOL;
                sx = sy / slope;
                x += sx;
                y += sy;
            } else {
                sy = (double)(iy + 1) - y;
                if (sy < TOL) sy = TOL;
                sx = sy / slope;
                x += sx;
                y += sy;
            }
        }
    }
}

// burners.rs:250-258
template <typename N, typename W> void burn_point(const std::vector<PointEdge>& pts, N v, W& w, N bg) {
    for (const auto& p : pts) w.write(p.y, p.x, v, bg);
}

// burners.rs:261-320 — active-edge-table scanline, even-odd, pixel-centre sampling.
template <typename N, typename W>
void burn_polygon(std::vector<PolyEdge>& pe, const RasterInfo& ri, N v, W& w, N bg) {
    if (pe.empty()) return;
    std::sort(pe.begin(), pe.end(), [](const PolyEdge& a, const PolyEdge& b) { return a.ystart < b.ystart; });
    uint64_t yline = pe[0].ystart;
    size_t next = 0;  // pe[next..] are the not-yet-activated edges ("drain" cursor)
    std::vector<PolyEdge> active;
    double ncols = (double)ri.ncols;
    while (yline < ri.nrows && (!active.empty() || next < pe.size())) {
        while (next < pe.size() && pe[next].ystart <= yline) active.push_back(pe[next++]);
        active.erase(std::remove_if(active.begin(), active.end(), [&](const PolyEdge& e) { return !(e.yend > yline); }),
                     active.end());
        if (active.empty()) {
            yline++;
            continue;
        }
        for (auto& e : active) e.x_at_yline = e.intersect_at(yline);
        std::sort(active.begin(), active.end(),
                  [](const PolyEdge& a, const PolyEdge& b) { return total_less(a.x_at_yline, b.x_at_yline); });
        for (size_t k = 0; k + 1 < active.size(); k += 2) {
            double x1 = active[k].x_at_yline, x2 = active[k + 1].x_at_yline;
            uint64_t xstart = as_usize(rs_clamp(std::floor(x1 + 0.5), 0.0, ncols));
            uint64_t xend = as_usize(rs_clamp(std::floor(x2 + 0.5), 0.0, ncols));
            for (uint64_t xp = xstart; xp < xend; xp++) w.write(yline, xp, v, bg);
        }
        yline++;
    }
}

// ----------------------------------------------------------------------------------------------
// Geometry dispatch — rust/src/rasterization/burn_geometry.rs
// ----------------------------------------------------------------------------------------------
struct Strategy {
    bool all_touched;     // S::IS_ALL_TOUCHED
    bool requires_dedup;  // S::REQUIRES_DEDUP  (prelude.rs:116-118: all_touched && fn in {sum,count})
};

template <typename N, typename W>
void burn_line_any(const Strategy& s, const std::vector<LineEdge>& e, const RasterInfo& ri, N v, W& w, N bg) {
    if (s.all_touched) burn_line_all_touched<N, W>(e, ri, v, w, bg);
    else burn_line_standard<N, W>(e, ri, v, w, bg);
}

// burn_geometry.rs:76-166 + handle_polygon :212-243 (rings of all member polygons pooled)
template <typename N, typename W>
void burn_polys(const std::vector<Poly>& polys, const Strategy& s, const RasterInfo& ri, N v, W& w, N bg) {
    std::vector<PolyEdge> pe;
    for (const auto& p : polys)
        for (const auto& r : p.rings) extract_ring(pe, r, ri);
    if (s.all_touched) {
        std::vector<LineEdge> le;
        for (const auto& p : polys)
            for (const auto& r : p.rings) extract_line(le, r, ri);
        if (s.requires_dedup) {
            PixelCache cache(le);
            LineWriter<N, W> lw{w, cache};
            burn_line_any<N>(s, le, ri, v, lw, bg);
            FillWriter<N, W> fw{w, cache};
            burn_polygon<N>(pe, ri, v, fw, bg);
        } else {
            burn_line_any<N>(s, le, ri, v, w, bg);
            burn_polygon<N>(pe, ri, v, w, bg);
        }
    } else {
        burn_polygon<N>(pe, ri, v, w, bg);
    }
}

// burn_geometry.rs:168-210 (segments of all member lines pooled; cache when pixels are not square)
template <typename N, typename W>
void burn_lines(const std::vector<LineStr>& lines, const Strategy& s, const RasterInfo& ri, N v, W& w, N bg) {
    std::vector<LineEdge> le;
    for (const auto& l : lines) extract_line(le, l, ri);
    if (ri.xres != ri.yres || s.requires_dedup) {
        PixelCache cache(le);
        LineWriter<N, W> lw{w, cache};
        burn_line_any<N>(s, le, ri, v, lw, bg);
    } else {
        burn_line_any<N>(s, le, ri, v, w, bg);
    }
}

// burn_geometry.rs:24-74
template <typename N, typename W>
void burn_geometry(const Geom& g, const Strategy& s, const RasterInfo& ri, N v, W& w, N bg) {
    switch (g.type) {
        case G_POINT:
        case G_MPOINT: {
            std::vector<PointEdge> pts;
            for (const auto& p : g.points) extract_point(pts, p, ri);
            burn_point<N>(pts, v, w, bg);
            break;
        }
        case G_POLY:
        case G_MPOLY:
            burn_polys<N>(g.polys, s, ri, v, w, bg);
            break;
        case G_LINE:
        case G_MLINE:
            burn_lines<N>(g.lines, s, ri, v, w, bg);
            break;
        case G_COLL:
            for (const auto& m : g.members) burn_geometry<N>(m, s, ri, v, w, bg);
            break;
    }
}

// ----------------------------------------------------------------------------------------------
// Orchestration — rust/src/rasterize.rs
// ----------------------------------------------------------------------------------------------
struct Job {
    const std::vector<Geom>* geoms;
    RasterInfo ri;
    int fn;
    bool all_touched;
    const void* field;          // scalar (field_is_scalar) or [G]
    bool field_is_scalar;
    const uint8_t* field_valid;  // nullable; 0 => geometry skipped (rasterize.rs:187-192)
    const int32_t* band_of_geom;  // nullable => single band
    int n_bands;
    const void* background;
    int threads;
};

// rasterize.rs:162-196 — geometries of one band in ascending original index.
template <typename N, typename W> void process_band(const Job& j, int band, W& w) {
    Strategy s{j.all_touched, j.all_touched && (j.fn == FN_SUM || j.fn == FN_COUNT)};
    N bg = *(const N*)j.background;
    const N* f = (const N*)j.field;
    for (size_t i = 0; i < j.geoms->size(); i++) {
        if (j.band_of_geom && j.band_of_geom[i] != band) continue;
        if (j.field_valid && !j.field_valid[i]) continue;
        N v = j.field_is_scalar ? f[0] : f[i];
        burn_geometry<N>((*j.geoms)[i], s, j.ri, v, w, bg);
    }
}

// rayon over bands only (rasterize.rs:89-101): one task per band, `threads` workers.
template <typename F> void for_bands(int n_bands, int threads, F f) {
    if (threads <= 1 || n_bands <= 1) {
        for (int b = 0; b < n_bands; b++) f(b);
        return;
    }
    std::vector<std::thread> pool;
    int nt = std::min(threads, n_bands);
    for (int t = 0; t < nt; t++)
        pool.emplace_back([=]() {
            for (int b = t; b < n_bands; b += nt) f(b);
        });
    for (auto& th : pool) th.join();
}

// rasterize.rs:71-116 (+ geo/raster.rs:23-28 background fill)
template <typename N> void run_dense(const Job& j, void* out_v) {
    N* out = (N*)out_v;
    uint64_t band_px = j.ri.nrows * j.ri.ncols;
    N bg = *(const N*)j.background;
    std::fill(out, out + band_px * (uint64_t)j.n_bands, bg);
    for_bands(j.n_bands, j.threads, [&](int b) {
        DenseWriter<N> w{Band<N>{out + band_px * (uint64_t)b, j.ri.ncols}, to_function<N>(j.fn)};
        process_band<N>(j, b, w);
    });
}

struct SparseOut {
    std::vector<uint64_t> rows, cols, counts;
    std::vector<uint8_t> data;  // raw bytes of N
    int itemsize;
};

// rasterize.rs:118-157 + writers.rs:101-131
template <typename N> void run_sparse(const Job& j, SparseOut& out) {
    std::vector<SparseWriter<N>> ws((size_t)j.n_bands);
    for_bands(j.n_bands, j.threads, [&](int b) { process_band<N>(j, b, ws[(size_t)b]); });
    out.itemsize = (int)sizeof(N);
    for (auto& w : ws) {
        out.counts.push_back(w.values.size());
        out.rows.insert(out.rows.end(), w.rows.begin(), w.rows.end());
        out.cols.insert(out.cols.end(), w.cols.begin(), w.cols.end());
        const uint8_t* p = (const uint8_t*)w.values.data();
        out.data.insert(out.data.end(), p, p + w.values.size() * sizeof(N));
    }
}

// arrays.rs:103-143 — replay triplets through the pixel function.
template <typename N>
void replay_sparse(const RasterInfo& ri, int fn, const void* bg_v, int n_bands, const uint64_t* counts,
                   const uint64_t* rows, const uint64_t* cols, const void* data_v, void* out_v) {
    N* out = (N*)out_v;
    const N* data = (const N*)data_v;
    N bg = *(const N*)bg_v;
    uint64_t band_px = ri.nrows * ri.ncols;
    std::fill(out, out + band_px * (uint64_t)n_bands, bg);
    uint64_t off = 0;
    PixelFn<N> f = to_function<N>(fn);
    for (int b = 0; b < n_bands; b++) {
        Band<N> band{out + band_px * (uint64_t)b, ri.ncols};
        for (uint64_t k = off; k < off + counts[b]; k++) f(band, rows[k], cols[k], data[k], bg);
        off += counts[b];
    }
}

enum { DT_U8 = 0, DT_U16, DT_U32, DT_U64, DT_I8, DT_I16, DT_I32, DT_I64, DT_F32, DT_F64 };
#define RZO_DISPATCH(dt, CALL)                         \
    switch (dt) {                                      \
        case DT_U8: { typedef uint8_t N; CALL; break; }   \
        case DT_U16: { typedef uint16_t N; CALL; break; } \
        case DT_U32: { typedef uint32_t N; CALL; break; } \
        case DT_U64: { typedef uint64_t N; CALL; break; } \
        case DT_I8: { typedef int8_t N; CALL; break; }    \
        case DT_I16: { typedef int16_t N; CALL; break; }  \
        case DT_I32: { typedef int32_t N; CALL; break; }  \
        case DT_I64: { typedef int64_t N; CALL; break; }  \
        case DT_F32: { typedef float N; CALL; break; }    \
        case DT_F64: { typedef double N; CALL; break; }   \
        default: return 3;                             \
    }

void set_err(char* err, size_t n, const char* msg) {
    if (err && n) {
        std::strncpy(err, msg, n - 1);
        err[n - 1] = 0;
    }
}

// geo::BoundingRect over the tree (geo/raster.rs:75-86); returns false when there is no coordinate.
bool bounds_of(const Geom& g, double b[4]) {
    bool any = false;
    auto add = [&](const Coord& c) {
        if (!any) { b[0] = b[2] = c.x; b[1] = b[3] = c.y; any = true; }
        else {
            b[0] = rs_min(b[0], c.x); b[1] = rs_min(b[1], c.y);
            b[2] = rs_max(b[2], c.x); b[3] = rs_max(b[3], c.y);
        }
    };
    for (const auto& p : g.points) add(p);
    for (const auto& l : g.lines) for (const auto& c : l) add(c);
    for (const auto& p : g.polys) if (!p.rings.empty()) for (const auto& c : p.rings[0]) add(c);  // exterior only
    for (const auto& m : g.members) {
        double mb[4];
        if (bounds_of(m, mb)) { add({mb[0], mb[1]}); add({mb[2], mb[3]}); }
    }
    return any;
}

}  // namespace

// ================================================================================================
// C interface (ctypes) — used by tests/ and bench.py only
// ================================================================================================
extern "C" {

struct rzo_raster_info {
    uint64_t nrows, ncols;
    double xmin, ymin, xmax, ymax, xres, yres;
};

struct rzo_geoms {
    std::vector<Geom> geoms;
};

// python/src/geo/parse_geometry.rs:109-121 — undecodable-to-geo geometries are dropped.
rzo_geoms* rzo_geoms_from_wkb(const uint8_t* const* bufs, const uint64_t* lens, uint64_t n, char* err, uint64_t errlen) {
    rzo_geoms* h = new rzo_geoms();
    for (uint64_t i = 0; i < n; i++) {
        WkbCursor c{bufs[i], bufs[i] + lens[i]};
        Geom g;
        bool keep = parse_wkb(c, g);
        if (!c.ok) {
            set_err(err, errlen, "Cannot parse geometry. Check that the WKB bytes are valid.");
            delete h;
            return nullptr;
        }
        if (keep) h->geoms.push_back(std::move(g));
    }
    return h;
}

// Bench helper: G simple polygons from one coordinate pool (ring i = coords[off[i]..off[i+1]) ).
rzo_geoms* rzo_geoms_from_rings(const double* x, const double* y, const uint64_t* off, uint64_t n_polys) {
    rzo_geoms* h = new rzo_geoms();
    h->geoms.resize(n_polys);
    for (uint64_t i = 0; i < n_polys; i++) {
        Geom& g = h->geoms[i];
        g.type = G_POLY;
        g.polys.emplace_back();
        g.polys[0].rings.emplace_back();
        LineStr& r = g.polys[0].rings[0];
        for (uint64_t k = off[i]; k < off[i + 1]; k++) r.push_back({x[k], y[k]});
        close_ring(r);
    }
    return h;
}

uint64_t rzo_geoms_len(const rzo_geoms* h) { return h->geoms.size(); }
void rzo_geoms_free(rzo_geoms* h) { delete h; }

int rzo_geoms_bounds(const rzo_geoms* h, double out[4]) {
    bool any = false;
    for (const auto& g : h->geoms) {
        double b[4];
        if (!bounds_of(g, b)) continue;
        if (!any) { std::memcpy(out, b, sizeof b); any = true; }
        else {
            out[0] = rs_min(out[0], b[0]); out[1] = rs_min(out[1], b[1]);
            out[2] = rs_max(out[2], b[2]); out[3] = rs_max(out[3], b[3]);
        }
    }
    return any ? 0 : 2;
}

// rust/src/geo/raster.rs:50-156.  Returns 0 ok, 1 ValueError, 2 RuntimeError (message in err).
int rzo_raster_info_build(int has_shape, uint64_t nrows, uint64_t ncols, int has_extent, const double* extent,
                          int has_res, double xres, double yres, int tap, const rzo_geoms* geoms,
                          rzo_raster_info* out, char* err, uint64_t errlen) {
    double xmin, ymin, xmax, ymax;
    bool inferred;
    if (has_extent) {
        if (extent[0] == 0.0 && extent[1] == 0.0 && extent[2] == 0.0 && extent[3] == 0.0 &&
            !std::signbit(extent[0]) && !std::signbit(extent[1]) && !std::signbit(extent[2]) && !std::signbit(extent[3])) {
            set_err(err, errlen, "Unspecified extent (all zeros).");
            return 1;
        }
        xmin = extent[0]; ymin = extent[1]; xmax = extent[2]; ymax = extent[3];
        inferred = false;
    } else {
        double b[4];
        if (!geoms || rzo_geoms_bounds(geoms, b) != 0) {
            set_err(err, errlen, "Cannot infer bounding box from geometry.");
            return 2;
        }
        xmin = b[0]; ymin = b[1]; xmax = b[2]; ymax = b[3];
        inferred = true;
    }
    if (!has_shape && !has_res) { set_err(err, errlen, "Must set at least one of `shape` or `resolution`"); return 1; }
    if (has_shape && has_res) { set_err(err, errlen, "Shape and resolution are mutually exclusive; provide only one"); return 1; }
    if (!has_shape) nrows = ncols = 0;
    if (!has_res) xres = yres = 0.0;
    if (has_shape && (nrows == 0 || ncols == 0)) { set_err(err, errlen, "Shape values must be > 0."); return 1; }
    if (has_res && (xres <= 0.0 || yres <= 0.0)) { set_err(err, errlen, "Resolution values must be > 0."); return 1; }
    if (inferred && !tap && has_res) {
        xmin -= xres / 2.0; xmax += xres / 2.0;
        ymin -= yres / 2.0; ymax += yres / 2.0;
    }
    if (!has_res) {
        xres = (xmax - xmin) / (double)ncols;
        yres = (ymax - ymin) / (double)nrows;
    } else if (tap) {
        xmin = std::floor(xmin / xres) * xres; xmax = std::ceil(xmax / xres) * xres;
        ymin = std::floor(ymin / yres) * yres; ymax = std::ceil(ymax / yres) * yres;
    }
    if (!has_shape) {
        nrows = as_usize(0.5 + (ymax - ymin) / yres);
        ncols = as_usize(0.5 + (xmax - xmin) / xres);
    }
    out->nrows = nrows; out->ncols = ncols;
    out->xmin = xmin; out->ymin = ymin; out->xmax = xmax; out->ymax = ymax;
    out->xres = xres; out->yres = yres;
    return 0;
}

// rust/src/rasterize.rs:199-205 — BTreeMap<&String,_>: bands in byte-lexicographic key order.
// Writes band_of_geom[n]; returns the number of bands; band_first[b] = index of the first geometry
// carrying band b's key (so the caller can recover the names).
int64_t rzo_group_keys(const char* const* keys, uint64_t n, int32_t* band_of_geom, uint64_t* band_first) {
    std::map<std::string, std::vector<uint64_t>> groups;
    for (uint64_t i = 0; i < n; i++) groups[std::string(keys[i])].push_back(i);
    int32_t b = 0;
    for (auto& kv : groups) {
        band_first[b] = kv.second[0];
        for (uint64_t i : kv.second) band_of_geom[i] = b;
        b++;
    }
    return b;
}

static int make_job(const rzo_geoms* g, const rzo_raster_info* ri, int fn, int all_touched, const void* field,
                    int field_is_scalar, uint64_t field_len, const uint8_t* field_valid, const int32_t* band_of_geom,
                    uint64_t by_len, int n_bands, const void* bg, int threads, Job& j, char* err, uint64_t errlen) {
    // rasterize.rs:208-229
    if (!field_is_scalar && field_len != g->geoms.size()) {
        set_err(err, errlen, "Geometry and field lengths must match");
        return 1;
    }
    if (band_of_geom && by_len != g->geoms.size()) {
        set_err(err, errlen, "Geometry and by lengths must match");
        return 1;
    }
    j.geoms = &g->geoms;
    j.ri = RasterInfo{ri->ncols, ri->nrows, ri->xmin, ri->xmax, ri->ymin, ri->ymax, ri->xres, ri->yres};
    j.fn = fn;
    j.all_touched = all_touched != 0;
    j.field = field;
    j.field_is_scalar = field_is_scalar != 0;
    j.field_valid = field_valid;
    j.band_of_geom = band_of_geom;
    j.n_bands = band_of_geom ? n_bands : 1;
    j.background = bg;
    j.threads = threads;
    return 0;
}

int rzo_rasterize_dense(const rzo_geoms* g, const rzo_raster_info* ri, int dtype, int fn, int all_touched,
                        const void* field, int field_is_scalar, uint64_t field_len, const uint8_t* field_valid,
                        const int32_t* band_of_geom, uint64_t by_len, int n_bands, const void* bg, int threads,
                        void* out, char* err, uint64_t errlen) {
    Job j;
    int rc = make_job(g, ri, fn, all_touched, field, field_is_scalar, field_len, field_valid, band_of_geom, by_len,
                      n_bands, bg, threads, j, err, errlen);
    if (rc) return rc;
    RZO_DISPATCH(dtype, run_dense<N>(j, out));
    return 0;
}

struct rzo_sparse {
    SparseOut s;
};

int rzo_rasterize_sparse(const rzo_geoms* g, const rzo_raster_info* ri, int dtype, int fn, int all_touched,
                         const void* field, int field_is_scalar, uint64_t field_len, const uint8_t* field_valid,
                         const int32_t* band_of_geom, uint64_t by_len, int n_bands, const void* bg, int threads,
                         rzo_sparse** out, char* err, uint64_t errlen) {
    Job j;
    int rc = make_job(g, ri, fn, all_touched, field, field_is_scalar, field_len, field_valid, band_of_geom, by_len,
                      n_bands, bg, threads, j, err, errlen);
    if (rc) return rc;
    rzo_sparse* s = new rzo_sparse();
    RZO_DISPATCH(dtype, run_sparse<N>(j, s->s));
    *out = s;
    return 0;
}
uint64_t rzo_sparse_len(const rzo_sparse* s) { return s->s.rows.size(); }
uint64_t rzo_sparse_bands(const rzo_sparse* s) { return s->s.counts.size(); }
const uint64_t* rzo_sparse_rows(const rzo_sparse* s) { return s->s.rows.data(); }
const uint64_t* rzo_sparse_cols(const rzo_sparse* s) { return s->s.cols.data(); }
const uint64_t* rzo_sparse_counts(const rzo_sparse* s) { return s->s.counts.data(); }
const void* rzo_sparse_data(const rzo_sparse* s) { return s->s.data.data(); }
void rzo_sparse_free(rzo_sparse* s) { delete s; }

int rzo_sparse_replay(const rzo_raster_info* ri, int dtype, int fn, const void* bg, int n_bands,
                      const uint64_t* counts, const uint64_t* rows, const uint64_t* cols, const void* data, void* out) {
    RasterInfo r{ri->ncols, ri->nrows, ri->xmin, ri->xmax, ri->ymin, ri->ymax, ri->xres, ri->yres};
    RZO_DISPATCH(dtype, replay_sparse<N>(r, fn, bg, n_bands, counts, rows, cols, data, out));
    return 0;
}

}  // extern "C"
