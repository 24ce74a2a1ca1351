// rz_host.hpp — host side of the burn path: flattened geometry pools, WKB/WKT readers, grid math.
// Pure C++17 (no CUDA) so it is testable on a CPU-only box.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <cstdlib>
#include <cstring>
#include <new>
#include <sys/mman.h>
#include <utility>
#include <vector>

#include "../../include/rz_b200.h"

namespace rz {

// Tag word stored next to every pooled vertex.
//   polygon / line pools: bits 0..29 part id, bit 30 = "owning line string is closed" (first coord ==
//   last coord, rust/src/geo/edges.rs:115), bit 31 = last vertex of its ring / line string (no
//   segment starts here).
//   point pool: part id.
constexpr uint32_t TAG_PART_MASK = 0x3fffffffu;
constexpr uint32_t TAG_CLOSED = 0x40000000u;
constexpr uint32_t TAG_SEQ_END = 0x80000000u;

// Allocator of the vertex pools: large blocks are 2 MiB-aligned and advised to use transparent huge pages.
// Filling a fresh 4 GB pool through 4 KiB pages spends most of its time in page faults (measured: 0.79 s
// against 0.23 s for 20M vertices); with huge pages the faults are 512 times fewer.
//
// Page-locked blocks: while a PinnedScope is alive on the calling thread, large allocations are served by the
// recycled, page-locked host blocks of rz_engine.cu (HostPool) through the hooks below - the final vertex pools
// of a geometry set are built straight into memory the copy engine can read at PCIe speed, and a freed
// geometry set hands its blocks to the next one (page-locking gigabytes per call costs more than the upload).
struct PinnedHooks {
    void* (*alloc)(size_t bytes) = nullptr;   // nullptr result: fall back to ordinary memory
    bool (*release)(void* p) = nullptr;       // false: `p` is not a pool block
};
extern PinnedHooks g_pinned_hooks;
extern thread_local bool t_alloc_pinned;
struct PinnedScope {
    bool prev;
    PinnedScope() : prev(t_alloc_pinned) { t_alloc_pinned = true; }
    ~PinnedScope() { t_alloc_pinned = prev; }
};

template <typename T> struct HugeAlloc {
    using value_type = T;
    HugeAlloc() = default;
    template <class U> HugeAlloc(const HugeAlloc<U>&) {}
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        void* p = nullptr;
        if (t_alloc_pinned && g_pinned_hooks.alloc && bytes >= ((size_t)4 << 20)) {
            p = g_pinned_hooks.alloc(bytes);
            if (p) return static_cast<T*>(p);
        }
        if (bytes >= ((size_t)4 << 20)) {
            const size_t huge = (size_t)2 << 20, rounded = (bytes + huge - 1) & ~(huge - 1);
            if (posix_memalign(&p, huge, rounded) != 0) throw std::bad_alloc();
            madvise(p, rounded, MADV_HUGEPAGE);  // advisory: failure only means ordinary pages
        } else {
            p = std::malloc(bytes ? bytes : 1);
            if (!p) throw std::bad_alloc();
        }
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t n) {
        if (n * sizeof(T) >= ((size_t)4 << 20) && g_pinned_hooks.release && g_pinned_hooks.release(p)) return;
        std::free(p);
    }
    // resize() default-initialises (no zero fill): the pools are always written right after they grow, and a
    // serial zero fill of gigabytes is exactly the first-touch cost the parallel ingestion avoids
    template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
    template <class U, class... Args> void construct(U* p, Args&&... args) {
        ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
    template <class U> bool operator==(const HugeAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const HugeAlloc<U>&) const { return false; }
};
template <typename T> using HVec = std::vector<T, HugeAlloc<T>>;

struct Pool {
    HVec<double> x, y;
    HVec<uint32_t> tag;
    // the sequences (rings / line strings) of the pool in order: index of the last vertex, "line string is
    // closed".  Together with the parts table they determine every tag, so the device rebuilds tag[] from them
    // (5 bytes per sequence on the wire instead of 4 bytes per vertex).
    std::vector<uint32_t> seq_end;
    std::vector<uint8_t> seq_closed;
    size_t size() const { return x.size(); }
};

struct DeviceGeoms;  // rz_engine.cu

// Upper bounds of the tile-binned engine's buffers for one (grid, row window, tile height): counted on the device
// with every polygon part active the first time a geometry set meets the grid, then reused so that later calls
// never wait for a count (rz_engine.cu).
struct TilePlanKey {
    uint64_t nrows, ncols;
    double xmin, ymax, xres, yres;
    uint32_t r0, r1, tile_r;
    bool operator<(const TilePlanKey& o) const {
        auto bits = [](double d) {
            uint64_t u;
            std::memcpy(&u, &d, 8);
            return u;
        };
        const uint64_t a[9] = {nrows, ncols, bits(xmin), bits(ymax), bits(xres), bits(yres), r0, r1, tile_r};
        const uint64_t b[9] = {o.nrows, o.ncols, bits(o.xmin), bits(o.ymax), bits(o.xres), bits(o.yres), o.r0, o.r1, o.tile_r};
        for (int i = 0; i < 9; i++)
            if (a[i] != b[i]) return a[i] < b[i];
        return false;
    }
};
struct TilePlan {
    uint64_t pairs = 0, units = 0, words = 0, edge_visits = 0, cross_lb = 0;
};

}  // namespace rz

// The opaque handle of include/rz_b200.h.
struct rz_geoms {
    uint64_t n_geoms = 0;
    rz::Pool pool[3];                  // indexed by RZ_PART_*
    rz::HVec<uint8_t> part_kind;    // [n_parts]
    rz::HVec<uint64_t> part_geom;   // [n_parts] owning geometry (index among kept geometries)
    rz::HVec<double> part_xlo, part_xhi;  // [n_parts] world-x extent of polygon parts (column-tile range)
    rz::HVec<double> part_ylo, part_yhi;  // [n_parts] world-y extent of polygon parts (row-tile range)
    rz::HVec<uint32_t> part_vbeg, part_vend;  // [n_parts] vertex range of the part inside its pool
    bool has_bounds = false;
    double bounds[4] = {0, 0, 0, 0};   // union of geo::BoundingRect, xmin ymin xmax ymax
    bool pinned = false;
    std::vector<std::pair<void*, size_t>> pinned_ranges;  // what cudaHostRegister accepted
    bool pool0_mapped = false;  // every byte of the polygon pool is page-locked and mapped: kernels may read it in place

    bool nonfinite = false;     // some coordinate is NaN or infinite (such sets never take the tile engine)

    std::mutex mu;
    std::map<int, rz::DeviceGeoms*> dev;  // cached device copies, by ordinal
    std::map<rz::TilePlanKey, rz::TilePlan> tile_plans;  // guarded by mu
    // multi-device calls: the part subsets of this set's shards (row bands of a grid / geometry ranges), built
    // once per (grid rows, band) and reused by later calls; guarded by mu
    std::map<std::vector<uint64_t>, std::shared_ptr<rz_geoms>> shards;

    ~rz_geoms();
};

namespace rz {

// Streaming builder shared by the WKB reader, the WKT reader and the SoA ingestion.  It applies
// the pooling rules of rust/src/rasterization/burn_geometry.rs:24-210:
//   * Polygon / MultiPolygon  -> ONE polygon part holding every ring of every member polygon
//   * LineString / MultiLineString -> ONE line part holding every member line string
//   * Point / MultiPoint -> ONE point part
//   * GeometryCollection -> its members' parts, in order (each burned independently)
class Flattener {
  public:
    explicit Flattener(rz_geoms* g) : g_(g) {}
    void begin_geometry();
    void end_geometry(bool keep);
    void begin_part(int kind);
    void end_part();
    // polygons: ring_is_exterior feeds geo::BoundingRect (exterior ring only); rings are closed
    // like geo_types::Polygon::new does.
    void begin_seq(bool counts_for_bounds);
    void coord(double x, double y);
    // n coordinates at once (same result as n coord() calls): separate x / y arrays, or interleaved
    // little-endian records of `stride` bytes starting with x, y (a WKB coordinate run)
    void coords(const double* xs, const double* ys, size_t n);
    void coords_le(const uint8_t* rec, size_t stride, size_t n);
    void end_seq();
    bool ok() const { return ok_; }
    const char* error() const { return err_; }

  private:
    void bound(double x, double y);
    void extents(size_t first, size_t n);
    rz_geoms* g_;
    int kind_ = -1;
    uint32_t part_ = 0;
    size_t seq_start_ = 0;
    bool seq_bounds_ = false;
    bool geom_has_bounds_ = false;
    double gb_[4];
    // rollback marks for a dropped top-level geometry
    size_t mark_pool_[3], mark_seq_[3], mark_parts_;
    bool ok_ = true;
    const char* err_ = "";
};

// ISO WKB / EWKB reader (little or big endian, Z/M dropped).  Returns false on malformed bytes.
// `*keep` = false when the top-level geometry has no geo_types equivalent (POINT EMPTY).
bool read_wkb(const uint8_t* buf, size_t len, Flattener& f, bool* keep);
// WKT reader with strtod-exact coordinates.
bool read_wkt(const char* s, Flattener& f, bool* keep);

void finish_geoms(rz_geoms* g);

// rz_geoms_from_soa: the caller's SoA form -> pools + parts table, written by `threads` threads straight into
// their final (page-locked) place: contiguous geometry ranges of equal coordinate counts, exact offsets from a
// counting pass, copy + extents in one sweep.  Same result as feeding the Flattener one geometry at a time
// (tests/test_host.py).  Returns RZ_OK or an error code with `err` set.
// Hooks let the caller start copying the pools to a device while they are still being written (rz_geoms_from_soa_to):
// on_sized runs once all sizes are known and the pools are allocated, on_range (from the worker threads) every time
// a few megabytes of a pool - vertices [v0, v1) of kind `kind`, x and y - have reached their final state.
struct FlattenHooks {
    void* ctx = nullptr;
    void (*on_sized)(void* ctx, rz_geoms* g) = nullptr;
    void (*on_range)(void* ctx, int kind, uint64_t v0, uint64_t v1) = nullptr;
};
// keep_part (nullable, [n_parts]): only the parts flagged non-zero are flattened - the geometry count and indices stay
// the caller's - which gives the part subset of a row band (subset_parts) without flattening the whole set first.
int flatten_soa(const rz_geom_soa* soa, rz_geoms* g, unsigned threads, std::string& err, const FlattenHooks* hooks = nullptr,
                const uint8_t* keep_part = nullptr);
int soa_part_y_extents(const rz_geom_soa* soa, unsigned threads, double* ylo, double* yhi, std::string& err);
// tag[] of a pool (part id | TAG_CLOSED | TAG_SEQ_END), rebuilt from the parts table and the sequence lists
// when a flattening path did not write it (the device never needs the host copy)
void ensure_tags(rz_geoms* g, int kind);
// the parts keep[0] < keep[1] < ... of `src` as a geometry set of their own (multi-device shards)
rz_geoms* subset_parts(const rz_geoms* src, const uint32_t* keep, size_t n_keep, unsigned threads);

int build_raster_info(const rz_raw_raster_info* raw, const rz_geoms* g, rz_raster_info* out, std::string& err);
int64_t group_keys(const char* const* keys, uint64_t n, int32_t* band_of_geom, uint64_t* band_first);

inline size_t dtype_size(int dt) {
    static const size_t s[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
    return (dt >= 0 && dt < 10) ? s[dt] : 0;
}

}  // namespace rz
