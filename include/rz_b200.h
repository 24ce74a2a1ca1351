/* rz_b200.h — C ABI of the B200-native rusterize burn path (librz_b200.so).
 *
 * The reference (ttrotto/rusterize) has no C ABI: its seam is the Rust generic call
 *     geoms.rasterize::<DenseArray<N> | SparseArray<N>>(RasterizeContext<N>)
 * (rust/src/rasterize.rs:54-62, implemented by ArrayBuilder::build at :71-116 and :118-157).
 * The entry points below are what an FFI shim for that seam binds (INTEGRATION.md shows the Rust
 * `extern "C"` block and the ctypes stub).  Plain pointers and sizes only; no torch types.
 *
 * Error convention (rust/src/error.rs:4-14): every fallible call returns RZ_OK, RZ_VALUE_ERROR or
 * RZ_RUNTIME_ERROR and writes the reference's message string into `err` (NUL-terminated, truncated
 * to errlen).  There is no CPU fallback: compute entry points return RZ_RUNTIME_ERROR when no CUDA
 * device is usable.
 *
 * Thread-safety: calls are re-entrant; the only shared state is a lazily created per-device
 * context (stream, scratch arena) guarded by a mutex, so concurrent calls on one device serialise.
 */
#ifndef RZ_B200_H
#define RZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RZ_OK 0
#define RZ_VALUE_ERROR 1   /* RusterizeError::ValueError   */
#define RZ_RUNTIME_ERROR 2 /* RusterizeError::RuntimeError */

/* Output dtypes, in the order of python/src/rusterize.rs:171-182. */
typedef enum {
    RZ_U8 = 0, RZ_U16, RZ_U32, RZ_U64, RZ_I8, RZ_I16, RZ_I32, RZ_I64, RZ_F32, RZ_F64
} rz_dtype;

/* rust/src/rasterization/pixel_functions.rs:8-16 */
typedef enum {
    RZ_SUM = 0, RZ_FIRST, RZ_LAST, RZ_MIN, RZ_MAX, RZ_COUNT, RZ_ANY
} rz_pixel_fn;

/* rust/src/geo/raster.rs:10-20 (RasterInfo).  epsg < 0 means None. */
typedef struct rz_raster_info {
    uint64_t nrows, ncols;
    double xmin, ymin, xmax, ymax;
    double xres, yres;
    int32_t epsg;
    int32_t _pad;
} rz_raster_info;

/* python/src/geo/raster.rs:6-14 (RawRasterInfo) / rust/src/geo/raster.rs:36-42 (RasterInfoBuilder). */
typedef struct rz_raw_raster_info {
    int32_t has_shape;
    int32_t has_extent;
    int32_t has_resolution;
    int32_t tap;
    uint64_t nrows, ncols;     /* shape = [nrows, ncols] */
    double extent[4];          /* xmin, ymin, xmax, ymax */
    double xres, yres;
    int32_t epsg;              /* < 0: None */
    int32_t _pad;
} rz_raw_raster_info;

/* Opaque geometry set: the host-side flattening of &[geo::Geometry<f64>] into SoA vertex pools
 * (polygon rings / line strings / points) plus a parts table, in pinned memory, with a cached
 * device copy.  Replaces the Vec<Geometry<f64>> argument of Rasterize::rasterize. */
typedef struct rz_geoms rz_geoms;

/* Part kinds of the flattened form (rust/src/rasterization/burn_geometry.rs:24-210). */
#define RZ_PART_POLYGON 0 /* Polygon / MultiPolygon / Rect / Triangle: all rings pooled (even-odd)   */
#define RZ_PART_LINE 1    /* LineString / MultiLineString / Line: all segments pooled                 */
#define RZ_PART_POINT 2   /* Point / MultiPoint                                                       */

/* Zero-parse ingestion for callers that already hold coordinates (the FFI shim walks
 * geo::Geometry and fills these; bench.py fills them from numpy).
 *   geometry g  = parts  [geom_part_off[g], geom_part_off[g+1])      (burn order; >1 only for collections)
 *   part p      = seqs   [part_seq_off[p],  part_seq_off[p+1])       (rings | line strings | one point run)
 *   sequence s  = coords [seq_coord_off[s], seq_coord_off[s+1])
 * Polygon rings must already be closed (geo_types::Polygon::new does that). */
typedef struct rz_geom_soa {
    uint64_t n_geoms, n_parts, n_seqs, n_coords;
    const uint64_t* geom_part_off; /* [n_geoms+1] */
    const uint8_t* part_kind;      /* [n_parts]   */
    const uint64_t* part_seq_off;  /* [n_parts+1] */
    const uint64_t* seq_coord_off; /* [n_seqs+1]  */
    const double* x;               /* [n_coords] world coordinates */
    const double* y;
} rz_geom_soa;

/* python/src/geo/parse_geometry.rs:109-121 (parse_sequence_wkb): ISO/EWKB, 2-D used, geometries with
 * no geo_types equivalent (POINT EMPTY) are dropped.  NULL + err on malformed input. */
rz_geoms* rz_geoms_from_wkb(const uint8_t* const* bufs, const uint64_t* lens, uint64_t n, char* err, size_t errlen);
/* python/src/geo/parse_geometry.rs:123-134 (parse_sequence_wkt) */
rz_geoms* rz_geoms_from_wkt(const char* const* strs, uint64_t n, char* err, size_t errlen);
rz_geoms* rz_geoms_from_soa(const rz_geom_soa* soa, char* err, size_t errlen);
/* The same, and the set is resident on `device` when the call returns: the pools are copied to the device WHILE they
 * are being flattened (the transfer hides behind the host sweep), so the first rasterize call on that device pays no
 * upload. */
rz_geoms* rz_geoms_from_soa_to(const rz_geom_soa* soa, int device, char* err, size_t errlen);
uint64_t rz_geoms_len(const rz_geoms* g);     /* geometries kept */
uint64_t rz_geoms_n_parts(const rz_geoms* g);
uint64_t rz_geoms_n_coords(const rz_geoms* g);
/* union of geo::BoundingRect (rust/src/geo/raster.rs:75-86); RZ_RUNTIME_ERROR when empty */
int rz_geoms_bounds(const rz_geoms* g, double out_xmin_ymin_xmax_ymax[4]);
/* Copy the flattened pools to `device` now (otherwise done lazily by the first rasterize call). */
int rz_geoms_upload(rz_geoms* g, int device, char* err, size_t errlen);
/* Drop cached device copies so the next call pays the host->device transfer again. */
void rz_geoms_evict(rz_geoms* g);
void rz_geoms_free(rz_geoms* g);
/* The parts of `g` that can write raster rows [row_begin, row_end) of the grid `ri`, as a geometry set of their
 * own: same order, same geometry indices (field / by arrays of the full set still apply), so burning the shard
 * with ctx->row_begin/row_end gives exactly those rows of the full result ("each GPU rasterises only
 * the edges intersecting its band"; the cull is the y-extent test of rust/src/geo/edges.rs:105 with band-local
 * bounds and a few rows of slack).  One process per GPU uses this to cut its share out of the full set;
 * rz_rasterize_dense_multi does the same internally.  Free with rz_geoms_free. */
rz_geoms* rz_geoms_row_shard(const rz_geoms* g, const rz_raster_info* ri, uint64_t row_begin, uint64_t row_end,
                             int all_touched, char* err, size_t errlen);
/* The same shard straight from the caller's arrays: only the band's parts are ever flattened (one parallel read of
 * the y ordinates decides which). */
rz_geoms* rz_geoms_from_soa_rows(const rz_geom_soa* soa, const rz_raster_info* ri, uint64_t row_begin, uint64_t row_end,
                                 int all_touched, char* err, size_t errlen);

/* Introspection of the flattened form (tests, FFI debugging). Returned pointers live as long as g. */
const uint8_t* rz_geoms_part_kind(const rz_geoms* g);
const uint64_t* rz_geoms_part_geom(const rz_geoms* g);
uint64_t rz_geoms_pool_len(const rz_geoms* g, int kind);
const double* rz_geoms_pool_x(const rz_geoms* g, int kind);
const double* rz_geoms_pool_y(const rz_geoms* g, int kind);
const uint32_t* rz_geoms_pool_tag(const rz_geoms* g, int kind);

/* rust/src/geo/raster.rs:50-156 (RasterInfoBuilder::build / build_with / finalize), same messages.
 * `g` may be NULL when raw->has_extent. */
int rz_raster_info_build(const rz_raw_raster_info* raw, const rz_geoms* g, rz_raster_info* out, char* err,
                         size_t errlen);

/* rust/src/rasterize.rs:199-205 (group_keys): bands are the distinct keys in byte-lexicographic
 * order.  Writes band_of_geom[n] and band_first[b] = first geometry index carrying band b's key.
 * Returns the number of bands. */
int64_t rz_group_keys(const char* const* keys, uint64_t n, int32_t* band_of_geom, uint64_t* band_first);

/* rust/src/prelude.rs:92-106 (RasterizeContext<N>) + execution controls that only exist here. */
typedef struct rz_context {
    rz_raster_info raster_info;
    int32_t dtype;    /* rz_dtype */
    int32_t pixel_fn; /* rz_pixel_fn */
    /* FieldSource (rust/src/rasterize.rs:24-31): one value of dtype, or field_len values */
    const void* field;
    int32_t field_is_scalar;
    int32_t all_touched;
    uint64_t field_len;
    const uint8_t* field_valid; /* nullable; 0 = null field => geometry skipped (rasterize.rs:187-192) */
    /* `by` after rz_group_keys; NULL => single band "band_1" */
    const int32_t* band_of_geom;
    uint64_t by_len;
    int32_t n_bands;
    int32_t device;         /* CUDA device ordinal */
    const void* background; /* one value of dtype */
    /* Row-band shard: only raster rows [row_begin, row_end) of every band are produced and `out`
     * is [n_bands][row_end-row_begin][ncols].  0,0 = all rows.  (north_star: row-band sharding.) */
    uint64_t row_begin, row_end;
    void* stream;           /* cudaStream_t to run on; NULL = the library's own per-device stream */
    uint32_t flags;
    uint32_t tile_bytes;    /* bytes of one shared-memory row tile per warp; 0 = default (4096) */
} rz_context;

#define RZ_FLAG_OUT_ON_DEVICE 1u   /* `out` is device memory (no D2H) */
#define RZ_FLAG_FORCE_H2D 2u       /* re-upload geometry even if a device copy is cached */
#define RZ_FLAG_SYNC_STAGES 4u     /* record per-stage CUDA-event timings into rz_stats */
#define RZ_FLAG_NO_TILE_ENGINE 8u      /* always use the crossing-record pipeline */
#define RZ_FLAG_FORCE_TILE_ENGINE 16u  /* use the tile-binned engine whenever the job is polygon-only */
#define RZ_FLAG_INPUTS_ON_DEVICE 128u   /* field, field_valid and band_of_geom are DEVICE pointers (same lengths and meaning):
                                          nothing is copied per call; band values are then not range-checked. Single-device
                                          calls only */
#define RZ_FLAG_OUT_ROW_COL_BAND 256u   /* dense `out` in R's array layout, (row, col, band) column-major = C-order
                                          [band][col][row] (R/rusterize/src/rust/src/encoding/rarrays.rs:9-17), instead of
                                          [band][row][col]: the R binding's permute done on the device */
#define RZ_FLAG_STREAMED_H2D 64u        /* accepted and ignored (round 1 experiment: pulling the polygon pool from mapped
                                          host memory under the raster's D2H did not pay) */

typedef struct rz_stats {
    uint64_t n_parts, n_poly_vertices, n_line_vertices, n_points;
    uint64_t n_records;        /* sort keys emitted (polygon crossings incl. tile replicas + line/point pixels) */
    uint64_t n_crossings;      /* polygon scanline crossings (X) */
    uint64_t n_tasks;          /* (band,row,column-tile) fill tasks */
    uint32_t key_bits, sort_passes, tile_width, n_windows;
    float h2d_ms, count_ms, emit_ms, sort_ms, index_ms, fill_ms, d2h_ms, total_ms;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t out_bytes;        /* B * rows * C * sizeof(dtype) */
    uint32_t kernel_launches;
    uint32_t engine;           /* 0 = crossing records + sort + row-tile fill, 1 = tile-binned polygon engine */
    uint64_t n_mask_words;     /* tile engine: 32-bit words of the compact inside-mask blocks (upper bound) */
    uint32_t host_syncs;       /* times the host waited for the device inside the call (0 for a steady-state
                                  tile-engine call with device output) */
    uint32_t plan_cached;      /* tile engine: buffer bounds came from the geometry handle's cache */
    float wall_ms;             /* host wall-clock time of the entry point (per device: its thread, shard_ms included) */
    float shard_ms;            /* multi-device calls: cutting this device's part subset out of the geometry set */
} rz_stats;

/* DenseArray::build (rust/src/rasterize.rs:71-116): out is [n_bands][rows][ncols] of ctx->dtype,
 * C-contiguous, host memory unless RZ_FLAG_OUT_ON_DEVICE.  Host memory may be page-locked (rz_host_alloc,
 * cudaHostAlloc, cudaHostRegister: row windows are copied straight into it at the link rate) or pageable (an array the
 * binding allocated itself): rasters of 64 MB and more then leave the device through two page-locked bounce blocks
 * and host copy threads instead of a driver-staged copy (env RZ_BOUNCE=0 restores the plain copy, RZ_BOUNCE_THREADS /
 * RZ_BOUNCE_BYTES / RZ_BOUNCE_MIN_BYTES tune it). */
int rz_rasterize_dense(rz_geoms* g, const rz_context* ctx, void* out, rz_stats* stats, char* err, size_t errlen);

/* SparseArray::build (rust/src/rasterize.rs:118-157): every (row, col, value) write, per band, in
 * burn order (rust/src/encoding/writers.rs:86-131). */
typedef struct rz_sparse rz_sparse;
int rz_rasterize_sparse(rz_geoms* g, const rz_context* ctx, rz_sparse** out, rz_stats* stats, char* err,
                        size_t errlen);
uint64_t rz_sparse_len(const rz_sparse* s);
uint64_t rz_sparse_n_bands(const rz_sparse* s);
const uint64_t* rz_sparse_rows(const rz_sparse* s);   /* [len] */
const uint64_t* rz_sparse_cols(const rz_sparse* s);   /* [len] */
const void* rz_sparse_data(const rz_sparse* s);       /* [len] of dtype */
const uint64_t* rz_sparse_counts(const rz_sparse* s); /* [n_bands] per-band triplet counts ("offsets") */
void rz_sparse_free(rz_sparse* s);
/* SparseArray::build_array (rust/src/encoding/arrays.rs:103-143): replay triplets through the pixel
 * function on the GPU.  rows/cols/data/counts are host arrays; out as in rz_rasterize_dense. */
int rz_sparse_build_array(const rz_context* ctx, uint64_t n_bands, const uint64_t* counts, const uint64_t* rows,
                          const uint64_t* cols, const void* data, void* out, rz_stats* stats, char* err,
                          size_t errlen);

/* Multi-device variants (the reference's only parallel section is rayon over `by` bands,
 * rust/src/rasterize.rs:89-101).  ONE call drives `n_devices` CUDA devices, one host thread per device, no
 * collective on the data path; ctx->device and ctx->stream are ignored.
 *   dense : device d owns the d-th of n_devices row bands of [row_begin,row_end); it uploads only the parts that
 *           can write those rows and copies its rows straight into the caller's host array `out`
 *           ([n_bands][rows][ncols], as rz_rasterize_dense; RZ_FLAG_OUT_ON_DEVICE is rejected).
 *   sparse: the triplet stream is ordered band -> geometry -> burn order (rust/src/encoding/writers.rs:101-131), so
 *           device d takes the d-th contiguous geometry range (ranges balanced by estimated work) and the streams
 *           are concatenated by offset into one rz_sparse.
 * Results are bit-identical to the single-device calls.  `stats` aggregates (sums of counts and bytes, maximum of
 * the stage times); `per_device` (nullable) receives n_devices entries. */
int rz_rasterize_dense_multi(rz_geoms* g, const rz_context* ctx, const int32_t* devices, int32_t n_devices, void* out,
                             rz_stats* stats, rz_stats* per_device, char* err, size_t errlen);
int rz_rasterize_sparse_multi(rz_geoms* g, const rz_context* ctx, const int32_t* devices, int32_t n_devices,
                              rz_sparse** out, rz_stats* stats, rz_stats* per_device, char* err, size_t errlen);

/* DenseArray::build (rust/src/rasterize.rs:77-115) in ONE call, for callers that hold the geometries and the context
 * at the same time (the FFI shim): flatten + upload + burn + copy back.  With several devices every device's host
 * thread flattens only the parts of its row band straight out of the caller's arrays (their extents come from one
 * parallel read of the y ordinates), so no full flattened copy is made first.  With env RZ_ONE_SHOT_RUNS=k a
 * device goes through its rows in k runs, flattening and uploading run i+1 while run i burns and copies back (off
 * by default: on the boxes measured the two compete for the host's memory system and nothing is gained).
 * `stats->shard_ms` = flattening time of the slowest device.  Same result as rz_geoms_from_soa +
 * rz_rasterize_dense_multi. */
int rz_rasterize_dense_soa(const rz_geom_soa* soa, const rz_context* ctx, const int32_t* devices, int32_t n_devices,
                           void* out, rz_stats* stats, rz_stats* per_device, char* err, size_t errlen);

/* Device plumbing */
/* Where the reference allocates its output with Array3::from_elem (RasterInfo::build_raster,
 * rust/src/geo/raster.rs:23-28), a caller of this library can take the array from here instead (the background fill
 * is fused into the burn, so the memory needs no initialisation).
 * Page-locked host memory for caller-owned outputs (and anything else handed to the library): huge-page backed,
 * faulted in by several threads, registered with CUDA as portable pinned memory, and placed on the memory nodes the
 * GPUs in use hang off (env RZ_DEVICES, else every visible device): interleaved over them when there are several,
 * so that devices attached to either socket copy into it at the same rate (a raster living on one socket is written
 * 40 % slower by the other socket's GPUs); env RZ_HOST_INTERLEAVE=0 turns the placement off, and it is skipped
 * silently where the kernel does not allow mbind.  Blocks are recycled by the library's host pool: rz_host_free
 * hands them back, it does not unmap them.  Returns NULL with the reason in `err`. */
void* rz_host_alloc(size_t bytes, char* err, size_t errlen);
void rz_host_free(void* p);
/* The library keeps freed page-locked blocks (vertex pools of freed geometry sets, freed sparse results, blocks given
 * back with rz_host_free) for the next call, up to RZ_HOST_POOL_BYTES (default: a quarter of the machine's memory,
 * 8 - 64 GiB; the oldest blocks go first when it is full).  rz_host_trim returns free blocks to the system until at
 * most `keep_bytes` stay pooled - between jobs of different shapes, or before handing memory to another consumer -
 * and returns the bytes still pooled. */
uint64_t rz_host_trim(uint64_t keep_bytes);

int rz_device_count(void);
const char* rz_version(void);
/* ABI self-description for bindings that mirror the structs by hand (ctypes, Rust #[repr(C)]): writes up to n of
 * 16 entries - sizeof rz_raster_info, rz_raw_raster_info, rz_geom_soa, rz_context, rz_stats; offsetof rz_context
 * .field .band_of_geom .background .row_begin .stream .flags; offsetof rz_stats .h2d_ms .h2d_bytes .kernel_launches
 * .n_mask_words .wall_ms - and returns 16. */
int rz_abi_layout(uint64_t* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* RZ_B200_H */
