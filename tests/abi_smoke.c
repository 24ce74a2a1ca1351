/* abi_smoke.c - a C caller of include/rz_b200.h, compiled with gcc against the header and linked to librz_b200.so
 * (tests/test_abi_c.py).  It burns the reference's own test geometries (python/test/test_many.py:19-28: GEOMS,
 * values 1..5, res (1,1), fun "sum", dtype uint8) and compares the value histogram of the raster with the one of the
 * reference's golden file python/test/data/standard_output_sum.tif (131 x 361).
 *   abi_smoke            -> layout + symbol checks only (no GPU needed), exit 0
 *   abi_smoke burn       -> the burn on device 0, exit 0 when the histogram matches */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rz_b200.h"

static const char* GEOMS[5] = {
    "POLYGON ((-180 -20, -140 55, 10 0, -140 -60, -180 -20), (-150 -20, -100 -10, -110 20, -150 -20))",
    "POLYGON ((-10 0, 140 60, 160 0, 140 -55, -10 0))",
    "POLYGON ((-125 0, 0 60, 40 5, 15 -45, -125 0))",
    "MULTILINESTRING ((-180 -70, -140 -50), (-140 -50, -100 -70), (-100 -70, -60 -50), (-60 -50, -20 -70), "
    "(-20 -70, 20 -50), (20 -50, 60 -70), (60 -70, 100 -50), (100 -50, 140 -70), (140 -70, 180 -50))",
    "GEOMETRYCOLLECTION (POINT (50 -40), POLYGON ((75 -40, 75 -30, 100 -30, 100 -40, 75 -40)), "
    "LINESTRING (60 -40, 80 0), GEOMETRYCOLLECTION (POLYGON ((100 20, 100 30, 110 30, 110 20, 100 20))))",
};
static const unsigned long long GOLDEN_HIST[8] = {22352, 6774, 8702, 4667, 3606, 826, 77, 287};

int main(int argc, char** argv) {
    char err[512] = {0};
    uint64_t layout[16];
    if (rz_abi_layout(layout, 16) != 16) return 10;
    const uint64_t mine[5] = {sizeof(rz_raster_info), sizeof(rz_raw_raster_info), sizeof(rz_geom_soa), sizeof(rz_context),
                              sizeof(rz_stats)};
    for (int i = 0; i < 5; i++)
        if (layout[i] != mine[i]) {
            fprintf(stderr, "struct %d: header %llu bytes, library %llu\n", i, (unsigned long long)mine[i],
                    (unsigned long long)layout[i]);
            return 11;
        }
    printf("%s, %d CUDA device(s)\n", rz_version(), rz_device_count());

    rz_geoms* g = rz_geoms_from_wkt(GEOMS, 5, err, sizeof err);
    if (!g) {
        fprintf(stderr, "rz_geoms_from_wkt: %s\n", err);
        return 12;
    }
    if (rz_geoms_len(g) != 5 || rz_geoms_n_parts(g) != 8) return 13; /* 3 polygons, 1 line part, 4 collection members */
    rz_raw_raster_info raw;
    memset(&raw, 0, sizeof raw);
    raw.has_resolution = 1;
    raw.xres = raw.yres = 1.0;
    raw.epsg = -1;
    rz_raster_info ri;
    if (rz_raster_info_build(&raw, g, &ri, err, sizeof err) != RZ_OK) {
        fprintf(stderr, "rz_raster_info_build: %s\n", err);
        return 14;
    }
    if (ri.nrows != 131 || ri.ncols != 361 || ri.xmin != -180.5 || ri.ymax != 60.5) return 15; /* half-pixel buffer */
    /* error strings are the reference's (rust/src/rasterize.rs:208-229) */
    const uint8_t values[5] = {1, 2, 3, 4, 5}, bg = 0;
    rz_context ctx;
    memset(&ctx, 0, sizeof ctx);
    ctx.raster_info = ri;
    ctx.dtype = RZ_U8;
    ctx.pixel_fn = RZ_SUM;
    ctx.field = values;
    ctx.field_len = 4; /* wrong on purpose */
    ctx.background = &bg;
    ctx.n_bands = 1;
    uint8_t* out = (uint8_t*)malloc(ri.nrows * ri.ncols);
    rz_stats st;
    int rc = rz_rasterize_dense(g, &ctx, out, &st, err, sizeof err);
    if (rc != RZ_VALUE_ERROR || strcmp(err, "Geometry and field lengths must match") != 0) {
        fprintf(stderr, "expected the length ValueError, got %d '%s'\n", rc, err);
        return 16;
    }
    if (argc < 2 || strcmp(argv[1], "burn") != 0) {
        rz_geoms_free(g);
        free(out);
        printf("abi ok (no burn requested)\n");
        return 0;
    }
    ctx.field_len = 5;
    rc = rz_rasterize_dense(g, &ctx, out, &st, err, sizeof err);
    if (rc != RZ_OK) {
        fprintf(stderr, "rz_rasterize_dense: %d %s\n", rc, err);
        return 17;
    }
    /* the same burn into the library's page-locked host memory (rz_host_alloc) gives the same bytes */
    uint8_t* pinned = (uint8_t*)rz_host_alloc(ri.nrows * ri.ncols, err, sizeof err);
    if (!pinned) {
        fprintf(stderr, "rz_host_alloc: %s\n", err);
        return 18;
    }
    rc = rz_rasterize_dense(g, &ctx, pinned, &st, err, sizeof err);
    if (rc != RZ_OK || memcmp(pinned, out, ri.nrows * ri.ncols) != 0) {
        fprintf(stderr, "burn into rz_host_alloc memory: %d %s\n", rc, err);
        return 19;
    }
    rz_host_free(pinned);
    unsigned long long hist[256] = {0}, sum = 0;
    for (uint64_t i = 0; i < ri.nrows * ri.ncols; i++) {
        hist[out[i]]++;
        sum += out[i];
    }
    int bad = sum != 59204ull;
    for (int v = 0; v < 256; v++) bad |= hist[v] != (v < 8 ? GOLDEN_HIST[v] : 0ull);
    /* the same job as a triplet stream: 29 363 writes (python/docs/python.md:106-136) */
    rz_sparse* sp = NULL;
    rc = rz_rasterize_sparse(g, &ctx, &sp, &st, err, sizeof err);
    if (rc != RZ_OK || rz_sparse_len(sp) != 29363ull || rz_sparse_rows(sp)[0] != 6 || rz_sparse_cols(sp)[0] != 40) bad |= 2;
    if (sp) rz_sparse_free(sp);
    rz_geoms_free(g);
    free(out);
    printf("histogram %s, sum %llu, %u kernel launches\n", bad ? "DIFFERS" : "matches the golden raster", sum,
           st.kernel_launches);
    return bad ? 18 : 0;
}
