// rz_host.cpp — geometry flattening (WKB / WKT / SoA -> pooled SoA + parts table), raster-grid
// math and band grouping.  Host only.
//
// Reference behaviour followed (paths relative to the reference repo):
//   rust/src/rasterization/burn_geometry.rs:24-210  which geometry types pool into one burn unit
//   python/src/geo/parse_geometry.rs:77-134          which inputs are dropped / rejected
//   rust/src/geo/raster.rs:50-156                    extent / shape / resolution rules
//   rust/src/rasterize.rs:199-205                    band order for `by`
#include "rz_host.hpp"

#include <emmintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>

namespace rz {

PinnedHooks g_pinned_hooks;
thread_local bool t_alloc_pinned = false;

// ------------------------------------------------------------------------------------------------
// Flattener
// ------------------------------------------------------------------------------------------------
void Flattener::begin_geometry() {
    for (int k = 0; k < 3; k++) {
        mark_pool_[k] = g_->pool[k].size();
        mark_seq_[k] = g_->pool[k].seq_end.size();
    }
    mark_parts_ = g_->part_kind.size();
    geom_has_bounds_ = false;
}

void Flattener::end_geometry(bool keep) {
    if (!keep) {  // roll back whatever the dropped geometry appended
        for (int k = 0; k < 3; k++) {
            g_->pool[k].x.resize(mark_pool_[k]);
            g_->pool[k].y.resize(mark_pool_[k]);
            g_->pool[k].tag.resize(mark_pool_[k]);
            g_->pool[k].seq_end.resize(mark_seq_[k]);
            g_->pool[k].seq_closed.resize(mark_seq_[k]);
        }
        g_->part_kind.resize(mark_parts_);
        g_->part_geom.resize(mark_parts_);
        g_->part_xlo.resize(mark_parts_);
        g_->part_xhi.resize(mark_parts_);
        g_->part_ylo.resize(mark_parts_);
        g_->part_yhi.resize(mark_parts_);
        g_->part_vbeg.resize(mark_parts_);
        g_->part_vend.resize(mark_parts_);
        return;
    }
    if (geom_has_bounds_) {
        if (!g_->has_bounds) {
            std::memcpy(g_->bounds, gb_, sizeof gb_);
            g_->has_bounds = true;
        } else {
            // rust/src/geo/raster.rs:81-84: f64::min / f64::max (NaN-ignoring)
            g_->bounds[0] = std::fmin(g_->bounds[0], gb_[0]);
            g_->bounds[1] = std::fmin(g_->bounds[1], gb_[1]);
            g_->bounds[2] = std::fmax(g_->bounds[2], gb_[2]);
            g_->bounds[3] = std::fmax(g_->bounds[3], gb_[3]);
        }
    }
    g_->n_geoms++;
}

void Flattener::begin_part(int kind) {
    kind_ = kind;
    size_t p = g_->part_kind.size();
    if (p >= TAG_PART_MASK) {
        ok_ = false;
        err_ = "Too many geometry parts (limit 2^30 - 1).";
        p = 0;
    }
    part_ = (uint32_t)p;
    g_->part_kind.push_back((uint8_t)kind);
    g_->part_geom.push_back(g_->n_geoms);
    g_->part_xlo.push_back(std::numeric_limits<double>::infinity());
    g_->part_xhi.push_back(-std::numeric_limits<double>::infinity());
    g_->part_ylo.push_back(std::numeric_limits<double>::infinity());
    g_->part_yhi.push_back(-std::numeric_limits<double>::infinity());
    g_->part_vbeg.push_back((uint32_t)g_->pool[kind].size());
    g_->part_vend.push_back((uint32_t)g_->pool[kind].size());
}

void Flattener::end_part() {
    g_->part_vend[part_] = (uint32_t)g_->pool[kind_].size();
    kind_ = -1;
}

void Flattener::begin_seq(bool counts_for_bounds) {
    seq_start_ = g_->pool[kind_].size();
    seq_bounds_ = counts_for_bounds;
}

void Flattener::bound(double x, double y) {
    // geo's bounding-rect fold: plain </> comparisons seeded by the first coordinate
    if (!geom_has_bounds_) {
        gb_[0] = gb_[2] = x;
        gb_[1] = gb_[3] = y;
        geom_has_bounds_ = true;
        return;
    }
    if (x < gb_[0]) gb_[0] = x;
    if (x > gb_[2]) gb_[2] = x;
    if (y < gb_[1]) gb_[1] = y;
    if (y > gb_[3]) gb_[3] = y;
}

void Flattener::coord(double x, double y) {
    Pool& p = g_->pool[kind_];
    p.x.push_back(x);
    p.y.push_back(y);
    p.tag.push_back(part_);
    if (seq_bounds_) bound(x, y);
    if (!(x - x == 0.0) || !(y - y == 0.0)) g_->nonfinite = true;
    if (kind_ == RZ_PART_POLYGON) {
        double& lo = g_->part_xlo[part_];
        double& hi = g_->part_xhi[part_];
        lo = std::fmin(lo, x);
        hi = std::fmax(hi, x);
        g_->part_ylo[part_] = std::fmin(g_->part_ylo[part_], y);
        g_->part_yhi[part_] = std::fmax(g_->part_yhi[part_], y);
    }
}

// Bounds / part extents of the n coordinates just appended at [first, first + n).  Equivalent to calling
// bound() and the fmin/fmax folds of coord() one coordinate at a time: every fold is "replace when strictly
// smaller / larger", which ignores NaN operands exactly like f64::min / f64::max do, so folding a run's own
// minimum (seeded with +-inf) into the running value gives the same result in any grouping.
void Flattener::extents(size_t first, size_t n) {
    if (n == 0) return;
    const Pool& p = g_->pool[kind_];
    const double* xs = p.x.data() + first;
    const double* ys = p.y.data() + first;
    if (seq_bounds_ && !geom_has_bounds_) bound(xs[0], ys[0]);  // geo's fold is seeded by the first coordinate (even a NaN one)
    const double inf = std::numeric_limits<double>::infinity();
    double xlo = inf, xhi = -inf, ylo = inf, yhi = -inf, zero = 0.0;
    for (size_t i = 0; i < n; i++) {
        const double x = xs[i], y = ys[i];
        xlo = x < xlo ? x : xlo;
        xhi = x > xhi ? x : xhi;
        ylo = y < ylo ? y : ylo;
        yhi = y > yhi ? y : yhi;
        zero += x * 0.0 + y * 0.0;  // NaN as soon as one coordinate is NaN or infinite
    }
    if (zero != zero) g_->nonfinite = true;
    if (seq_bounds_) {  // (folding the seed coordinate in again changes nothing)
        if (xlo < gb_[0]) gb_[0] = xlo;
        if (xhi > gb_[2]) gb_[2] = xhi;
        if (ylo < gb_[1]) gb_[1] = ylo;
        if (yhi > gb_[3]) gb_[3] = yhi;
    }
    if (kind_ == RZ_PART_POLYGON) {
        g_->part_xlo[part_] = std::fmin(g_->part_xlo[part_], xlo);
        g_->part_xhi[part_] = std::fmax(g_->part_xhi[part_], xhi);
        g_->part_ylo[part_] = std::fmin(g_->part_ylo[part_], ylo);
        g_->part_yhi[part_] = std::fmax(g_->part_yhi[part_], yhi);
    }
}

void Flattener::coords(const double* xs, const double* ys, size_t n) {
    if (n == 0) return;
    Pool& p = g_->pool[kind_];
    const size_t first = p.x.size();
    p.x.insert(p.x.end(), xs, xs + n);
    p.y.insert(p.y.end(), ys, ys + n);
    p.tag.insert(p.tag.end(), n, part_);
    extents(first, n);
}

void Flattener::coords_le(const uint8_t* rec, size_t stride, size_t n) {
    if (n == 0) return;
    Pool& p = g_->pool[kind_];
    const size_t first = p.x.size();
    p.x.resize(first + n);
    p.y.resize(first + n);
    p.tag.insert(p.tag.end(), n, part_);
    double* xd = p.x.data() + first;
    double* yd = p.y.data() + first;
    for (size_t i = 0; i < n; i++, rec += stride) {
        std::memcpy(xd + i, rec, 8);
        std::memcpy(yd + i, rec + 8, 8);
    }
    extents(first, n);
}

void Flattener::end_seq() {
    Pool& p = g_->pool[kind_];
    size_t n = p.size() - seq_start_;
    if (n == 0 || kind_ == RZ_PART_POINT) return;
    size_t a = seq_start_, b = p.size() - 1;
    bool closed = p.x[a] == p.x[b] && p.y[a] == p.y[b];
    if (kind_ == RZ_PART_POLYGON && !closed) {
        // geo_types::Polygon::new closes every ring
        double fx = p.x[a], fy = p.y[a];
        p.x.push_back(fx);
        p.y.push_back(fy);
        p.tag.push_back(part_);
        closed = true;
    }
    const bool flag_closed = kind_ == RZ_PART_LINE && closed;
    if (flag_closed)
        for (size_t i = seq_start_; i < p.size(); i++) p.tag[i] |= TAG_CLOSED;
    p.tag.back() |= TAG_SEQ_END;
    p.seq_end.push_back((uint32_t)(p.size() - 1));
    p.seq_closed.push_back(flag_closed ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// WKB
// ------------------------------------------------------------------------------------------------
namespace {

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    bool need(size_t n) {
        if ((size_t)(end - p) < n) ok = false;
        return ok;
    }
    uint8_t u8() {
        if (!need(1)) return 0;
        return *p++;
    }
    uint32_t u32(bool le) {
        if (!need(4)) return 0;
        uint32_t v = le ? ((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24)
                        : ((uint32_t)p[3] | (uint32_t)p[2] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[0] << 24);
        p += 4;
        return v;
    }
    double f64(bool le) {
        if (!need(8)) return 0;
        uint64_t v = 0;
        if (le) std::memcpy(&v, p, 8);
        else
            for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
        p += 8;
        double d;
        std::memcpy(&d, &v, 8);
        return d;
    }
};

struct WkbHeader {
    bool le;
    uint32_t type;
    int dims;
};

bool wkb_header(Cursor& c, WkbHeader& h) {
    uint8_t bo = c.u8();
    if (!c.ok || bo > 1) return c.ok = false;
    h.le = bo == 1;
    uint32_t t = c.u32(h.le);
    h.dims = 2;
    if (t & 0x80000000u) h.dims++;
    if (t & 0x40000000u) h.dims++;
    bool srid = (t & 0x20000000u) != 0;
    t &= 0x0fffffffu;
    uint32_t iso = t / 1000;
    if (iso == 1 || iso == 2) h.dims = 3;
    else if (iso == 3) h.dims = 4;
    h.type = t % 1000;
    if (srid) c.u32(h.le);
    return c.ok;
}

void wkb_coords(Cursor& c, const WkbHeader& h, Flattener& f) {
    uint32_t n = c.u32(h.le);
    if (!c.need((size_t)n * 8 * (size_t)h.dims)) return;
    if (h.le) {  // the common case: one pass over the run
        f.coords_le(c.p, 8 * (size_t)h.dims, n);
        c.p += (size_t)n * 8 * (size_t)h.dims;
        return;
    }
    for (uint32_t i = 0; i < n; i++) {
        double x = c.f64(h.le), y = c.f64(h.le);
        for (int k = 2; k < h.dims; k++) c.f64(h.le);
        f.coord(x, y);
    }
}

void wkb_poly_rings(Cursor& c, const WkbHeader& h, Flattener& f) {
    uint32_t nr = c.u32(h.le);
    for (uint32_t r = 0; r < nr && c.ok; r++) {
        f.begin_seq(r == 0);
        wkb_coords(c, h, f);
        f.end_seq();
    }
}

// `nested`: >0 while inside a GeometryCollection.  Returns whether the geometry exists in geo_types.
bool wkb_geom(Cursor& c, Flattener& f, int depth) {
    if (depth > 64) return c.ok = false;
    WkbHeader h;
    if (!wkb_header(c, h)) return false;
    switch (h.type) {
        case 1: {  // Point
            if (!c.need(8 * (size_t)h.dims)) return false;
            double x = c.f64(h.le), y = c.f64(h.le);
            for (int k = 2; k < h.dims; k++) c.f64(h.le);
            if (x != x && y != y) return false;  // POINT EMPTY -> try_to_geometry() == None
            f.begin_part(RZ_PART_POINT);
            f.begin_seq(true);
            f.coord(x, y);
            f.end_seq();
            f.end_part();
            return true;
        }
        case 2:  // LineString
            f.begin_part(RZ_PART_LINE);
            f.begin_seq(true);
            wkb_coords(c, h, f);
            f.end_seq();
            f.end_part();
            return c.ok;
        case 3:  // Polygon
            f.begin_part(RZ_PART_POLYGON);
            wkb_poly_rings(c, h, f);
            f.end_part();
            return c.ok;
        case 4: {  // MultiPoint
            uint32_t n = c.u32(h.le);
            f.begin_part(RZ_PART_POINT);
            f.begin_seq(true);
            for (uint32_t i = 0; i < n && c.ok; i++) {
                WkbHeader m;
                if (!wkb_header(c, m) || m.type != 1) return c.ok = false;
                if (!c.need(8 * (size_t)m.dims)) return false;
                double x = c.f64(m.le), y = c.f64(m.le);
                for (int k = 2; k < m.dims; k++) c.f64(m.le);
                if (!(x != x && y != y)) f.coord(x, y);
            }
            f.end_seq();
            f.end_part();
            return c.ok;
        }
        case 5: {  // MultiLineString
            uint32_t n = c.u32(h.le);
            f.begin_part(RZ_PART_LINE);
            for (uint32_t i = 0; i < n && c.ok; i++) {
                WkbHeader m;
                if (!wkb_header(c, m) || m.type != 2) return c.ok = false;
                f.begin_seq(true);
                wkb_coords(c, m, f);
                f.end_seq();
            }
            f.end_part();
            return c.ok;
        }
        case 6: {  // MultiPolygon
            uint32_t n = c.u32(h.le);
            f.begin_part(RZ_PART_POLYGON);
            for (uint32_t i = 0; i < n && c.ok; i++) {
                WkbHeader m;
                if (!wkb_header(c, m) || m.type != 3) return c.ok = false;
                wkb_poly_rings(c, m, f);
            }
            f.end_part();
            return c.ok;
        }
        case 7: {  // GeometryCollection: members burned one after another (burn_geometry.rs:64-74)
            uint32_t n = c.u32(h.le);
            for (uint32_t i = 0; i < n && c.ok; i++) wkb_geom(c, f, depth + 1);
            return c.ok;
        }
        default:
            return c.ok = false;
    }
}

}  // namespace

bool read_wkb(const uint8_t* buf, size_t len, Flattener& f, bool* keep) {
    Cursor c{buf, buf + len};
    *keep = wkb_geom(c, f, 0);
    return c.ok && f.ok();
}

// ------------------------------------------------------------------------------------------------
// WKT
// ------------------------------------------------------------------------------------------------
namespace {

struct Lex {
    const char* s;
    bool ok = true;
    void ws() {
        while (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r') s++;
    }
    bool eat(char ch) {
        ws();
        if (*s == ch) {
            s++;
            return true;
        }
        return false;
    }
    void expect(char ch) {
        if (!eat(ch)) ok = false;
    }
    // case-insensitive keyword
    bool word(std::string& out) {
        ws();
        out.clear();
        while ((*s >= 'A' && *s <= 'Z') || (*s >= 'a' && *s <= 'z')) out.push_back((char)(*s++ & ~0x20));
        return !out.empty();
    }
    bool peek_alpha() {
        ws();
        return (*s >= 'A' && *s <= 'Z') || (*s >= 'a' && *s <= 'z');
    }
    double num() {
        ws();
        char* e = nullptr;
        double v = std::strtod(s, &e);  // correctly rounded, like Rust's f64::from_str
        if (e == s) ok = false;
        s = e;
        return v;
    }
};

int wkt_dims(Lex& l, bool* empty) {
    int dims = 2;
    *empty = false;
    for (;;) {
        const char* save = l.s;
        std::string w;
        if (!l.peek_alpha()) break;
        l.word(w);
        if (w == "Z" || w == "M") dims = 3;
        else if (w == "ZM") dims = 4;
        else if (w == "EMPTY") {
            *empty = true;
            break;
        } else {
            l.s = save;
            l.ok = false;
            break;
        }
    }
    return dims;
}

void wkt_coord(Lex& l, int dims, double* x, double* y) {
    *x = l.num();
    *y = l.num();
    for (int k = 2; k < dims; k++) l.num();
}

void wkt_coord_list(Lex& l, int dims, Flattener& f) {  // "(x y, x y, ...)"
    l.expect('(');
    do {
        double x, y;
        wkt_coord(l, dims, &x, &y);
        if (!l.ok) return;
        f.coord(x, y);
    } while (l.eat(','));
    l.expect(')');
}

void wkt_poly_body(Lex& l, int dims, Flattener& f) {  // "((ring),(ring))"
    l.expect('(');
    int r = 0;
    do {
        f.begin_seq(r++ == 0);
        wkt_coord_list(l, dims, f);
        f.end_seq();
    } while (l.ok && l.eat(','));
    l.expect(')');
}

bool wkt_geom(Lex& l, Flattener& f, int depth) {
    if (depth > 64) return l.ok = false;
    std::string name;
    if (!l.word(name)) return l.ok = false;
    bool empty;
    int dims = wkt_dims(l, &empty);
    if (!l.ok) return false;
    if (name == "POINT") {
        if (empty) return false;
        l.expect('(');
        double x, y;
        wkt_coord(l, dims, &x, &y);
        l.expect(')');
        if (!l.ok) return false;
        f.begin_part(RZ_PART_POINT);
        f.begin_seq(true);
        f.coord(x, y);
        f.end_seq();
        f.end_part();
        return true;
    }
    if (name == "LINESTRING") {
        f.begin_part(RZ_PART_LINE);
        f.begin_seq(true);
        if (!empty) wkt_coord_list(l, dims, f);
        f.end_seq();
        f.end_part();
        return l.ok;
    }
    if (name == "POLYGON") {
        f.begin_part(RZ_PART_POLYGON);
        if (!empty) wkt_poly_body(l, dims, f);
        f.end_part();
        return l.ok;
    }
    if (name == "MULTIPOINT") {
        f.begin_part(RZ_PART_POINT);
        f.begin_seq(true);
        if (!empty) {
            l.expect('(');
            do {
                double x, y;
                if (l.eat('(')) {
                    wkt_coord(l, dims, &x, &y);
                    l.expect(')');
                } else {
                    std::string w;
                    const char* save = l.s;
                    if (l.peek_alpha() && l.word(w) && w == "EMPTY") continue;
                    l.s = save;
                    wkt_coord(l, dims, &x, &y);
                }
                if (!l.ok) break;
                f.coord(x, y);
            } while (l.eat(','));
            l.expect(')');
        }
        f.end_seq();
        f.end_part();
        return l.ok;
    }
    if (name == "MULTILINESTRING") {
        f.begin_part(RZ_PART_LINE);
        if (!empty) {
            l.expect('(');
            do {
                f.begin_seq(true);
                wkt_coord_list(l, dims, f);
                f.end_seq();
            } while (l.ok && l.eat(','));
            l.expect(')');
        }
        f.end_part();
        return l.ok;
    }
    if (name == "MULTIPOLYGON") {
        f.begin_part(RZ_PART_POLYGON);
        if (!empty) {
            l.expect('(');
            do {
                wkt_poly_body(l, dims, f);
            } while (l.ok && l.eat(','));
            l.expect(')');
        }
        f.end_part();
        return l.ok;
    }
    if (name == "GEOMETRYCOLLECTION") {
        if (!empty) {
            l.expect('(');
            do {
                wkt_geom(l, f, depth + 1);
            } while (l.ok && l.eat(','));
            l.expect(')');
        }
        return l.ok;
    }
    return l.ok = false;
}

}  // namespace

bool read_wkt(const char* s, Flattener& f, bool* keep) {
    Lex l{s};
    *keep = wkt_geom(l, f, 0);
    l.ws();
    if (*l.s != 0) l.ok = false;
    return l.ok && f.ok();
}

void finish_geoms(rz_geoms* g) {
    // give back over-allocation only when it is large: shrinking copies the whole vector, and reserved but
    // untouched memory is address space, not resident pages
    auto fit = [](auto& v) {
        if (v.capacity() > 2 * v.size() + 4096) v.shrink_to_fit();
    };
    for (int k = 0; k < 3; k++) {
        fit(g->pool[k].x);
        fit(g->pool[k].y);
        fit(g->pool[k].tag);
    }
}

// ------------------------------------------------------------------------------------------------
// Parallel SoA ingestion
// ------------------------------------------------------------------------------------------------
namespace {

// dst[i] = src[i] for one coordinate array, folding the run's minimum / maximum (strict compares seeded with
// +-inf: NaN operands are ignored, like Flattener::extents) and whether any value is NaN or infinite.  Four
// independent accumulators keep the min / max dependency chains off the critical path.
inline void copy_fold(const double* src, double* dst, size_t n, double& lo, double& hi, uint64_t& bad) {
    const double inf = std::numeric_limits<double>::infinity();
    double l0 = inf, l1 = inf, l2 = inf, l3 = inf, h0 = -inf, h1 = -inf, h2 = -inf, h3 = -inf;
    uint64_t b = 0;
    const uint64_t EXP = 0x7ff0000000000000ull;
    auto bits = [](double d) {
        uint64_t u;
        std::memcpy(&u, &d, 8);
        return u;
    };
    size_t i = 0;
    // The pools are written once and next read by the copy engine: streaming (non-temporal) stores keep them out of
    // the cache and save the read-for-ownership of every destination line - a third of the sweep's memory traffic.
    if (n < 32) {  // short runs: plain stores (streaming stores of partial lines from many open streams thrash the
                   // write-combining buffers: 10M five-vertex parcels flattened 4x slower with them)
        for (; i < n; i++) {
            const double a = src[i];
            dst[i] = a;
            l0 = a < l0 ? a : l0;
            h0 = a > h0 ? a : h0;
            b |= (uint64_t)((bits(a) & EXP) == EXP);
        }
        lo = l0;
        hi = h0;
        bad |= b;
        return;
    }
    if ((uintptr_t)dst & 15u) {  // align the destination to 16 bytes
        const double a = src[0];
        dst[0] = a;
        l0 = a < l0 ? a : l0;
        h0 = a > h0 ? a : h0;
        b |= (uint64_t)((bits(a) & EXP) == EXP);
        i = 1;
    }
    for (; i + 4 <= n; i += 4) {
        const double a0 = src[i], a1 = src[i + 1], a2 = src[i + 2], a3 = src[i + 3];
        _mm_stream_pd(dst + i, _mm_set_pd(a1, a0));
        _mm_stream_pd(dst + i + 2, _mm_set_pd(a3, a2));
        l0 = a0 < l0 ? a0 : l0;
        l1 = a1 < l1 ? a1 : l1;
        l2 = a2 < l2 ? a2 : l2;
        l3 = a3 < l3 ? a3 : l3;
        h0 = a0 > h0 ? a0 : h0;
        h1 = a1 > h1 ? a1 : h1;
        h2 = a2 > h2 ? a2 : h2;
        h3 = a3 > h3 ? a3 : h3;
        b |= (uint64_t)((bits(a0) & EXP) == EXP) | (uint64_t)((bits(a1) & EXP) == EXP) |
             (uint64_t)((bits(a2) & EXP) == EXP) | (uint64_t)((bits(a3) & EXP) == EXP);
    }
    for (; i < n; i++) {
        const double a = src[i];
        dst[i] = a;
        l0 = a < l0 ? a : l0;
        h0 = a > h0 ? a : h0;
        b |= (uint64_t)((bits(a) & EXP) == EXP);
    }
    l0 = l1 < l0 ? l1 : l0;
    l2 = l3 < l2 ? l3 : l2;
    l0 = l2 < l0 ? l2 : l0;
    h0 = h1 > h0 ? h1 : h0;
    h2 = h3 > h2 ? h3 : h2;
    h0 = h2 > h0 ? h2 : h0;
    lo = l0;
    hi = h0;
    bad |= b;
}

struct SoaChunk {
    uint64_t g0 = 0, g1 = 0;               // geometry range
    uint64_t pool_cnt[3] = {0, 0, 0};      // vertices written per pool (closing vertices included)
    uint64_t seq_cnt[2] = {0, 0};          // non-empty sequences of the polygon / line pools
    uint64_t part_cnt = 0, part_off = 0;   // parts written (all of the range's parts unless a keep mask is given)
    uint64_t pool_off[3] = {0, 0, 0}, seq_off[2] = {0, 0};
    bool has_bounds = false, nonfinite = false;
    double bounds[4] = {0, 0, 0, 0};
    const char* error = nullptr;
};

}  // namespace

int flatten_soa(const rz_geom_soa* soa, rz_geoms* g, unsigned threads, std::string& err, const FlattenHooks* hooks,
                const uint8_t* keep_part) {
    const uint64_t G = soa->n_geoms, NP = soa->n_parts, NS = soa->n_seqs, NC = soa->n_coords;
    if (G == 0) return RZ_OK;
    if (soa->geom_part_off[G] > NP || soa->part_seq_off[NP] > NS || soa->seq_coord_off[NS] > NC) {
        err = "Inconsistent SoA offsets";
        return RZ_VALUE_ERROR;
    }
    if (NP >= TAG_PART_MASK) {
        err = "Too many geometry parts (limit 2^30 - 1).";
        return RZ_RUNTIME_ERROR;
    }
    // first coordinate of geometry gi (non-decreasing in gi): where the chunk cuts are searched
    auto first_coord = [&](uint64_t gi) -> uint64_t {
        if (gi >= G) return NC;
        const uint64_t p = soa->geom_part_off[gi];
        if (p >= NP) return NC;
        const uint64_t s = soa->part_seq_off[p];
        return s >= NS ? NC : soa->seq_coord_off[s];
    };
    const unsigned T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)std::max(1u, threads), G, NC / 65536 + 1}));
    std::vector<SoaChunk> ch(T);
    for (unsigned t = 0; t < T; t++) {
        auto cut = [&](unsigned k) -> uint64_t {
            if (k == 0) return 0;
            if (k >= T) return G;
            const uint64_t target = NC / T * k;
            uint64_t lo = 0, hi = G;  // first geometry whose first coordinate is >= target
            while (lo < hi) {
                const uint64_t mid = (lo + hi) / 2;
                if (first_coord(mid) < target) lo = mid + 1;
                else hi = mid;
            }
            return lo;
        };
        ch[t].g0 = cut(t);
        ch[t].g1 = cut(t + 1);
    }
    auto run = [&](auto&& body) {
        if (T == 1) {
            body(0u);
            return;
        }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++) th.emplace_back([&, t]() { body(t); });
        for (auto& x : th) x.join();
    };
    // ---- pass A: exact output sizes ---------------------------------------------------------------
    run([&](unsigned t) {
        SoaChunk& c = ch[t];
        for (uint64_t p = soa->geom_part_off[c.g0]; p < soa->geom_part_off[c.g1]; p++) {
            const unsigned kind = soa->part_kind[p];
            if (kind > 2) {
                c.error = "Invalid part kind";
                return;
            }
            if (keep_part && !keep_part[p]) continue;
            c.part_cnt++;
            for (uint64_t s = soa->part_seq_off[p]; s < soa->part_seq_off[p + 1]; s++) {
                const uint64_t k0 = soa->seq_coord_off[s], k1 = soa->seq_coord_off[s + 1];
                if (k1 < k0 || k1 > NC) {
                    c.error = "Inconsistent SoA offsets";
                    return;
                }
                if (k1 == k0) continue;
                uint64_t n = k1 - k0;
                if (kind == RZ_PART_POLYGON) {  // geo_types::Polygon::new closes every ring
                    const bool closed = soa->x[k0] == soa->x[k1 - 1] && soa->y[k0] == soa->y[k1 - 1];
                    n += closed ? 0 : 1;
                }
                c.pool_cnt[kind] += n;
                if (kind != RZ_PART_POINT) c.seq_cnt[kind]++;
            }
        }
    });
    for (auto& c : ch)
        if (c.error) {
            err = c.error;
            return RZ_VALUE_ERROR;
        }
    uint64_t pool_tot[3] = {0, 0, 0}, seq_tot[2] = {0, 0}, part_tot = 0;
    for (auto& c : ch) {
        for (int k = 0; k < 3; k++) {
            c.pool_off[k] = pool_tot[k];
            pool_tot[k] += c.pool_cnt[k];
        }
        for (int k = 0; k < 2; k++) {
            c.seq_off[k] = seq_tot[k];
            seq_tot[k] += c.seq_cnt[k];
        }
        c.part_off = part_tot;
        part_tot += c.part_cnt;
    }
    for (int k = 0; k < 3; k++)
        if (pool_tot[k] >= 0xfffffff0ull) {
            err = "Too many vertices (limit 2^32 per pool).";
            return RZ_RUNTIME_ERROR;
        }
    {
        PinnedScope pinned;  // the pools and the parts table go straight into recycled page-locked blocks
        for (int k = 0; k < 3; k++) {
            g->pool[k].x.resize(pool_tot[k]);
            g->pool[k].y.resize(pool_tot[k]);
        }
        g->part_kind.resize(part_tot);
        g->part_geom.resize(part_tot);
        g->part_xlo.resize(part_tot);
        g->part_xhi.resize(part_tot);
        g->part_ylo.resize(part_tot);
        g->part_yhi.resize(part_tot);
        g->part_vbeg.resize(part_tot);
        g->part_vend.resize(part_tot);
    }
    for (int k = 0; k < 2; k++) {
        g->pool[k].seq_end.resize(seq_tot[k]);
        g->pool[k].seq_closed.resize(seq_tot[k]);
    }
    if (hooks && hooks->on_sized) hooks->on_sized(hooks->ctx, g);
    // ---- pass B: copy + extents ---------------------------------------------------------------------
    const double inf = std::numeric_limits<double>::infinity();
    const uint64_t NOTIFY = (uint64_t)1 << 19;  // vertices per on_range call (4 MB of x + 4 MB of y)
    run([&](unsigned t) {
        SoaChunk& c = ch[t];
        uint64_t at[3] = {c.pool_off[0], c.pool_off[1], c.pool_off[2]};
        uint64_t sent[3] = {c.pool_off[0], c.pool_off[1], c.pool_off[2]};
        uint64_t sq[2] = {c.seq_off[0], c.seq_off[1]};
        uint64_t bad = 0;
        uint64_t pi = c.part_off;  // index of the next part written (== p without a keep mask)
        auto notify = [&](int kind, bool flush) {
            if (!hooks || !hooks->on_range || at[kind] == sent[kind]) return;
            if (!flush && at[kind] - sent[kind] < NOTIFY) return;
            _mm_sfence();  // the streaming stores of this range are visible to the copy engine
            hooks->on_range(hooks->ctx, kind, sent[kind], at[kind]);
            sent[kind] = at[kind];
        };
        for (uint64_t gi = c.g0; gi < c.g1; gi++) {
            bool gb_has = false;
            double gb[4] = {0, 0, 0, 0};
            for (uint64_t p0 = soa->geom_part_off[gi]; p0 < soa->geom_part_off[gi + 1]; p0++) {
                if (keep_part && !keep_part[p0]) continue;
                const uint64_t p = pi++;
                const int kind = soa->part_kind[p0];
                Pool& pool = g->pool[kind];
                double pxlo = inf, pxhi = -inf, pylo = inf, pyhi = -inf;
                g->part_kind[p] = (uint8_t)kind;
                g->part_geom[p] = gi;
                g->part_vbeg[p] = (uint32_t)at[kind];
                for (uint64_t s = soa->part_seq_off[p0]; s < soa->part_seq_off[p0 + 1]; s++) {
                    const uint64_t k0 = soa->seq_coord_off[s], n = soa->seq_coord_off[s + 1] - k0;
                    if (n == 0) continue;
                    double* xd = pool.x.data() + at[kind];
                    double* yd = pool.y.data() + at[kind];
                    double xlo, xhi, ylo, yhi;
                    copy_fold(soa->x + k0, xd, n, xlo, xhi, bad);
                    copy_fold(soa->y + k0, yd, n, ylo, yhi, bad);
                    const double *xs = soa->x + k0, *ys = soa->y + k0;  // (read the source: xd / yd were streamed out)
                    if (!gb_has) {  // geo's fold is seeded by the first coordinate (Flattener::bound)
                        gb[0] = gb[2] = xs[0];
                        gb[1] = gb[3] = ys[0];
                        gb_has = true;
                    }
                    if (xlo < gb[0]) gb[0] = xlo;
                    if (xhi > gb[2]) gb[2] = xhi;
                    if (ylo < gb[1]) gb[1] = ylo;
                    if (yhi > gb[3]) gb[3] = yhi;
                    pxlo = xlo < pxlo ? xlo : pxlo;  // (the folds never return NaN: plain compares, no libm calls)
                    pxhi = xhi > pxhi ? xhi : pxhi;
                    pylo = ylo < pylo ? ylo : pylo;
                    pyhi = yhi > pyhi ? yhi : pyhi;
                    uint64_t m = n;
                    if (kind == RZ_PART_POINT) {
                        at[kind] += m;
                        notify(kind, false);
                        continue;
                    }
                    bool closed = xs[0] == xs[n - 1] && ys[0] == ys[n - 1];
                    if (kind == RZ_PART_POLYGON && !closed) {
                        xd[n] = xs[0];
                        yd[n] = ys[0];
                        m = n + 1;
                        closed = true;
                    }
                    at[kind] += m;
                    pool.seq_end[sq[kind]] = (uint32_t)(at[kind] - 1);
                    pool.seq_closed[sq[kind]] = (kind == RZ_PART_LINE && closed) ? 1 : 0;
                    sq[kind]++;
                    notify(kind, false);
                }
                g->part_vend[p] = (uint32_t)at[kind];
                const bool poly = kind == RZ_PART_POLYGON;
                g->part_xlo[p] = poly ? pxlo : inf;
                g->part_xhi[p] = poly ? pxhi : -inf;
                g->part_ylo[p] = poly ? pylo : inf;
                g->part_yhi[p] = poly ? pyhi : -inf;
            }
            if (gb_has) {  // Flattener::end_geometry (rust/src/geo/raster.rs:81-84)
                if (!c.has_bounds) {
                    std::memcpy(c.bounds, gb, sizeof gb);
                    c.has_bounds = true;
                } else if (gb[0] == gb[0] && gb[1] == gb[1] && gb[2] == gb[2] && gb[3] == gb[3] && c.bounds[0] == c.bounds[0] &&
                           c.bounds[1] == c.bounds[1] && c.bounds[2] == c.bounds[2] && c.bounds[3] == c.bounds[3]) {
                    c.bounds[0] = gb[0] < c.bounds[0] ? gb[0] : c.bounds[0];
                    c.bounds[1] = gb[1] < c.bounds[1] ? gb[1] : c.bounds[1];
                    c.bounds[2] = gb[2] > c.bounds[2] ? gb[2] : c.bounds[2];
                    c.bounds[3] = gb[3] > c.bounds[3] ? gb[3] : c.bounds[3];
                } else {  // NaN seeds (a geometry whose first coordinate is NaN): libm's fmin / fmax rules
                    c.bounds[0] = std::fmin(c.bounds[0], gb[0]);
                    c.bounds[1] = std::fmin(c.bounds[1], gb[1]);
                    c.bounds[2] = std::fmax(c.bounds[2], gb[2]);
                    c.bounds[3] = std::fmax(c.bounds[3], gb[3]);
                }
            }
        }
        c.nonfinite = bad != 0;
        _mm_sfence();  // streaming stores are visible before the thread is joined
        for (int k = 0; k < 3; k++) notify(k, true);
    });
    for (auto& c : ch) {
        if (c.nonfinite) g->nonfinite = true;
        if (!c.has_bounds) continue;
        if (!g->has_bounds) {
            std::memcpy(g->bounds, c.bounds, sizeof g->bounds);
            g->has_bounds = true;
        } else {
            g->bounds[0] = std::fmin(g->bounds[0], c.bounds[0]);
            g->bounds[1] = std::fmin(g->bounds[1], c.bounds[1]);
            g->bounds[2] = std::fmax(g->bounds[2], c.bounds[2]);
            g->bounds[3] = std::fmax(g->bounds[3], c.bounds[3]);
        }
    }
    g->n_geoms = G;
    return RZ_OK;
}

// world-y extent of every part of a caller's SoA (one parallel read of the y array): what a one-shot multi-device
// call needs to decide which device gets which part before anything is copied.  A part holding a NaN ordinate gets
// (-inf, +inf): kept everywhere.
int soa_part_y_extents(const rz_geom_soa* soa, unsigned threads, double* ylo, double* yhi, std::string& err) {
    const uint64_t NP = soa->n_parts, NS = soa->n_seqs, NC = soa->n_coords;
    if (soa->part_seq_off[NP] > NS || soa->seq_coord_off[NS] > NC) {
        err = "Inconsistent SoA offsets";
        return RZ_VALUE_ERROR;
    }
    const unsigned T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)std::max(1u, threads), NP, NC / 65536 + 1}));
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<std::thread> th;
    auto body = [&](unsigned t) {
        for (uint64_t p = NP * t / T; p < NP * (t + 1) / T; p++) {
            double lo = inf, hi = -inf;
            bool odd = false;
            const uint64_t s0 = soa->part_seq_off[p], s1 = soa->part_seq_off[p + 1];
            if (s1 > s0) {
                const double* y = soa->y;
                for (uint64_t k = soa->seq_coord_off[s0]; k < soa->seq_coord_off[s1]; k++) {
                    const double a = y[k];
                    lo = a < lo ? a : lo;
                    hi = a > hi ? a : hi;
                    odd |= !(a == a);
                }
            }
            ylo[p] = odd ? -inf : lo;
            yhi[p] = odd ? inf : hi;
        }
    };
    if (T == 1) body(0);
    else {
        for (unsigned t = 0; t < T; t++) th.emplace_back([&, t]() { body(t); });
        for (auto& x : th) x.join();
    }
    return RZ_OK;
}

void ensure_tags(rz_geoms* g, int kind) {
    std::lock_guard<std::mutex> lk(g->mu);
    Pool& pool = g->pool[kind];
    if (pool.tag.size() == pool.x.size()) return;
    pool.tag.resize(pool.x.size());
    for (size_t p = 0; p < g->part_kind.size(); p++)
        if (g->part_kind[p] == kind)
            for (uint32_t v = g->part_vbeg[p]; v < g->part_vend[p]; v++) pool.tag[v] = (uint32_t)p;
    size_t start = 0;
    for (size_t i = 0; i < pool.seq_end.size(); i++) {
        const size_t end = pool.seq_end[i];
        if (pool.seq_closed[i])
            for (size_t v = start; v <= end; v++) pool.tag[v] |= TAG_CLOSED;
        pool.tag[end] |= TAG_SEQ_END;
        start = end + 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Part subsets (multi-device sharding)
// ------------------------------------------------------------------------------------------------
// A new geometry set holding the parts keep[0] < keep[1] < ... of `src`, in the same order: vertex ranges,
// sequence lists and the parts table are copied (by `threads` threads, exact offsets from a counting pass);
// geometry indices are unchanged (n_geoms stays src's), so field / by arrays of the original call still apply.
rz_geoms* subset_parts(const rz_geoms* src, const uint32_t* keep, size_t n_keep, unsigned threads) {
    std::unique_ptr<rz_geoms> g(new rz_geoms());
    g->n_geoms = src->n_geoms;
    g->has_bounds = src->has_bounds;
    std::memcpy(g->bounds, src->bounds, sizeof g->bounds);
    g->nonfinite = src->nonfinite;
    if (n_keep == 0) return g.release();
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>({(size_t)std::max(1u, threads), n_keep / 4096 + 1}));
    struct Chunk {
        size_t j0, j1;
        uint64_t pool_cnt[3] = {0, 0, 0}, seq_cnt[2] = {0, 0}, pool_off[3], seq_off[2];
    };
    std::vector<Chunk> ch(T);
    for (unsigned t = 0; t < T; t++) {
        ch[t].j0 = n_keep * t / T;
        ch[t].j1 = n_keep * (t + 1) / T;
    }
    auto run = [&](auto&& body) {
        if (T == 1) {
            body(0u);
            return;
        }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < T; t++) th.emplace_back([&, t]() { body(t); });
        for (auto& x : th) x.join();
    };
    // sequences of part p inside its pool: those whose last vertex lies in [vbeg, vend).  Kept parts ascend, so do
    // their sequences: a cursor per kind (placed by one binary search, then advanced linearly) finds them.
    struct SeqCursor {
        size_t at[2] = {0, 0};
        bool placed[2] = {false, false};
    };
    auto seq_range = [&](SeqCursor& cur, uint32_t p, size_t& lo, size_t& hi) {
        const int k = src->part_kind[p];
        const auto& se = src->pool[k].seq_end;
        const uint32_t vb = src->part_vbeg[p], ve = src->part_vend[p];
        if (!cur.placed[k]) {
            cur.at[k] = std::lower_bound(se.begin(), se.end(), vb) - se.begin();
            cur.placed[k] = true;
        }
        size_t a = cur.at[k];
        if (a < se.size() && se[a] < vb && a + 64 < se.size() && se[a + 64] < vb)  // a long jump between kept parts
            a = std::lower_bound(se.begin() + a, se.end(), vb) - se.begin();
        while (a < se.size() && se[a] < vb) a++;
        size_t b = a;
        while (b < se.size() && se[b] < ve) b++;
        lo = a;
        hi = b;
        cur.at[k] = b;
    };
    run([&](unsigned t) {
        Chunk& c = ch[t];
        SeqCursor cur;
        for (size_t j = c.j0; j < c.j1; j++) {
            const uint32_t p = keep[j];
            const int k = src->part_kind[p];
            c.pool_cnt[k] += src->part_vend[p] - src->part_vbeg[p];
            if (k != RZ_PART_POINT) {
                size_t lo, hi;
                seq_range(cur, p, lo, hi);
                c.seq_cnt[k] += hi - lo;
            }
        }
    });
    uint64_t pool_tot[3] = {0, 0, 0}, seq_tot[2] = {0, 0};
    for (auto& c : ch) {
        for (int k = 0; k < 3; k++) {
            c.pool_off[k] = pool_tot[k];
            pool_tot[k] += c.pool_cnt[k];
        }
        for (int k = 0; k < 2; k++) {
            c.seq_off[k] = seq_tot[k];
            seq_tot[k] += c.seq_cnt[k];
        }
    }
    {
        PinnedScope pinned;
        for (int k = 0; k < 3; k++) {
            g->pool[k].x.resize(pool_tot[k]);
            g->pool[k].y.resize(pool_tot[k]);
        }
        g->part_kind.resize(n_keep);
        g->part_geom.resize(n_keep);
        g->part_xlo.resize(n_keep);
        g->part_xhi.resize(n_keep);
        g->part_ylo.resize(n_keep);
        g->part_yhi.resize(n_keep);
        g->part_vbeg.resize(n_keep);
        g->part_vend.resize(n_keep);
    }
    for (int k = 0; k < 2; k++) {
        g->pool[k].seq_end.resize(seq_tot[k]);
        g->pool[k].seq_closed.resize(seq_tot[k]);
    }
    run([&](unsigned t) {
        Chunk& c = ch[t];
        uint64_t at[3] = {c.pool_off[0], c.pool_off[1], c.pool_off[2]}, sq[2] = {c.seq_off[0], c.seq_off[1]};
        SeqCursor cur;
        for (size_t j = c.j0; j < c.j1; j++) {
            const uint32_t p = keep[j];
            const int k = src->part_kind[p];
            const uint32_t vb = src->part_vbeg[p], n = src->part_vend[p] - vb;
            std::memcpy(g->pool[k].x.data() + at[k], src->pool[k].x.data() + vb, (size_t)n * 8);
            std::memcpy(g->pool[k].y.data() + at[k], src->pool[k].y.data() + vb, (size_t)n * 8);
            g->part_kind[j] = (uint8_t)k;
            g->part_geom[j] = src->part_geom[p];
            g->part_xlo[j] = src->part_xlo[p];
            g->part_xhi[j] = src->part_xhi[p];
            g->part_ylo[j] = src->part_ylo[p];
            g->part_yhi[j] = src->part_yhi[p];
            g->part_vbeg[j] = (uint32_t)at[k];
            g->part_vend[j] = (uint32_t)(at[k] + n);
            if (k != RZ_PART_POINT) {
                size_t lo, hi;
                seq_range(cur, p, lo, hi);
                for (size_t q = lo; q < hi; q++) {
                    g->pool[k].seq_end[sq[k]] = (uint32_t)(src->pool[k].seq_end[q] - vb + at[k]);
                    g->pool[k].seq_closed[sq[k]] = src->pool[k].seq_closed[q];
                    sq[k]++;
                }
            }
            at[k] += n;
        }
    });
    return g.release();
}

// ------------------------------------------------------------------------------------------------
// Grid math — rust/src/geo/raster.rs:50-156
// ------------------------------------------------------------------------------------------------
namespace {
// Rust `f64 as usize`: truncate, saturate, NaN -> 0
uint64_t as_usize(double v) {
    if (!(v == v) || v <= 0.0) return 0;
    if (v >= 18446744073709551615.0) return UINT64_MAX;
    return (uint64_t)v;
}
bool is_pos_or_neg_zero_equal_zero(double v) {
    // total_cmp(&0.0) == Equal: only +0.0 qualifies (raster.rs:53)
    return v == 0.0 && !std::signbit(v);
}
}  // namespace

int build_raster_info(const rz_raw_raster_info* raw, const rz_geoms* g, rz_raster_info* out, std::string& err) {
    double xmin, ymin, xmax, ymax;
    bool inferred;
    if (raw->has_extent) {
        const double* e = raw->extent;
        if (is_pos_or_neg_zero_equal_zero(e[0]) && is_pos_or_neg_zero_equal_zero(e[1]) &&
            is_pos_or_neg_zero_equal_zero(e[2]) && is_pos_or_neg_zero_equal_zero(e[3])) {
            err = "Unspecified extent (all zeros).";
            return RZ_VALUE_ERROR;
        }
        xmin = e[0];
        ymin = e[1];
        xmax = e[2];
        ymax = e[3];
        inferred = false;
    } else {
        if (!g || !g->has_bounds) {
            err = "Cannot infer bounding box from geometry.";
            return RZ_RUNTIME_ERROR;
        }
        xmin = g->bounds[0];
        ymin = g->bounds[1];
        xmax = g->bounds[2];
        ymax = g->bounds[3];
        inferred = true;
    }
    bool has_shape = raw->has_shape != 0, has_res = raw->has_resolution != 0, tap = raw->tap != 0;
    if (!has_shape && !has_res) {
        err = "Must set at least one of `shape` or `resolution`";
        return RZ_VALUE_ERROR;
    }
    if (has_shape && has_res) {
        err = "Shape and resolution are mutually exclusive; provide only one";
        return RZ_VALUE_ERROR;
    }
    uint64_t nrows = has_shape ? raw->nrows : 0, ncols = has_shape ? raw->ncols : 0;
    double xres = has_res ? raw->xres : 0.0, yres = has_res ? raw->yres : 0.0;
    if (has_shape && (nrows == 0 || ncols == 0)) {
        err = "Shape values must be > 0.";
        return RZ_VALUE_ERROR;
    }
    if (has_res && (xres <= 0.0 || yres <= 0.0)) {
        err = "Resolution values must be > 0.";
        return RZ_VALUE_ERROR;
    }
    if (inferred && !tap && has_res) {  // half-pixel buffer
        xmin -= xres / 2.0;
        xmax += xres / 2.0;
        ymin -= yres / 2.0;
        ymax += yres / 2.0;
    }
    if (!has_res) {
        xres = (xmax - xmin) / (double)ncols;
        yres = (ymax - ymin) / (double)nrows;
    } else if (tap) {
        xmin = std::floor(xmin / xres) * xres;
        xmax = std::ceil(xmax / xres) * xres;
        ymin = std::floor(ymin / yres) * yres;
        ymax = std::ceil(ymax / yres) * yres;
    }
    if (!has_shape) {
        nrows = as_usize(0.5 + (ymax - ymin) / yres);
        ncols = as_usize(0.5 + (xmax - xmin) / xres);
    }
    out->nrows = nrows;
    out->ncols = ncols;
    out->xmin = xmin;
    out->ymin = ymin;
    out->xmax = xmax;
    out->ymax = ymax;
    out->xres = xres;
    out->yres = yres;
    out->epsg = raw->epsg;
    out->_pad = 0;
    return RZ_OK;
}

// rust/src/rasterize.rs:199-205 — BTreeMap<&String, Vec<usize>>: byte-lexicographic key order.
int64_t group_keys(const char* const* keys, uint64_t n, int32_t* band_of_geom, uint64_t* band_first) {
    std::map<std::string, int32_t> order;
    for (uint64_t i = 0; i < n; i++) order.emplace(std::string(keys[i]), 0);
    int32_t b = 0;
    for (auto& kv : order) kv.second = b++;
    std::vector<char> seen((size_t)b, 0);
    for (uint64_t i = 0; i < n; i++) {
        int32_t k = order[std::string(keys[i])];
        band_of_geom[i] = k;
        if (!seen[(size_t)k]) {
            seen[(size_t)k] = 1;
            band_first[k] = i;
        }
    }
    return b;
}

}  // namespace rz
