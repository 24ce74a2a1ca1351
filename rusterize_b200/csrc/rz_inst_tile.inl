// rz_inst_tile.inl — tile_apply_kernel<N, FN, TILE_R, MODE, BGNAN> for one half of the dtypes (see rz_inst_fill.inl).
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "rz_dispatch.hpp"

namespace rz {

template <typename N> struct NanBackground {
    static constexpr bool possible = false;
    static bool is(N) { return false; }
};
template <> struct NanBackground<float> {
    static constexpr bool possible = true;
    static bool is(float v) { return v != v; }
};
template <> struct NanBackground<double> {
    static constexpr bool possible = true;
    static bool is(double v) { return v != v; }
};

// Which evaluation of the pixel-function rule a job gets (rz_tiles.cuh, MODE).  What depends on the burn VALUES
// (all finite / none equal to the background) is decided on the device: modes 1 and 3 fall back to the generic
// body inside the kernel.
// (the opt-in to more than 48 KB of dynamic shared memory is per device: set on every launch, it costs nothing)
template <typename K, typename... A>
static void launch_apply(K kfn, dim3 grid, int threads, size_t smem, cudaStream_t s, A... a) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;  // surfaces through cudaGetLastError
    kfn<<<grid, threads, smem, s>>>(a...);
}

template <typename N, int FN>
static void tile_launch(cudaStream_t s, KParams P, TileParams T, const uint32_t* tile_start, const BlockDesc* desc,
                        const uint32_t* masks, const TileCounters* tcnt, uint64_t bg, void* out) {
    constexpr int TR = sizeof(N) <= 4 ? 64 : 32;
    constexpr bool is_float = std::is_floating_point<N>::value;
    constexpr bool additive = FN == RZ_SUM || FN == RZ_COUNT;
    constexpr bool ordered = FN == RZ_FIRST || FN == RZ_MIN || FN == RZ_MAX;
    // one CTA per T.apply_tiles tiles of a tile row: (column groups, tile rows x bands) when that fits the grid
    // limits, else flattened
    const uint64_t gy = (uint64_t)T.n_tr * P.n_bands;
    const uint32_t groups = (T.n_tc + T.apply_tiles - 1) / T.apply_tiles;
    static const bool force_1d = std::getenv("RZ_APPLY_1D") != nullptr;  // tests: exercise the flattened grid
    const dim3 grid = (gy <= 65535 && !force_1d) ? dim3(groups, (unsigned)gy) : dim3((unsigned)(groups * gy));
    // flush staging (8 padded rows per consumer warp) + the staged mask blocks and their mbarriers
    const size_t smem = apply_smem_bytes<N, TR>();
    constexpr int THREADS = TR * 4 + 32;  // consumer warps + the producer warp
    N bgv;
    std::memcpy(&bgv, &bg, sizeof(N));
    const bool bg_nan = NanBackground<N>::is(bgv);
    if (additive && is_float && bg_nan)  // MODE 1 (when every value is finite): masked add + touched mask
        launch_apply(tile_apply_kernel<N, additive ? FN : RZ_SUM, TR, is_float ? 1 : 0, true>, grid, THREADS, smem, s, P, T, tile_start, desc, masks, tcnt, bg, (N*)out);
    else if (additive && !is_float && bg == 0)  // MODE 2: plain masked add
        launch_apply(tile_apply_kernel<N, additive ? FN : RZ_SUM, TR, is_float ? 0 : 2, false>, grid, THREADS, smem, s, P, T, tile_start, desc, masks, tcnt, bg, (N*)out);
    else if (ordered && (is_float ? bg_nan : true))  // MODE 3 (when no value can look like the background)
        launch_apply(tile_apply_kernel<N, ordered ? FN : RZ_FIRST, TR, 3, NanBackground<N>::possible>, grid, THREADS, smem, s, P, T, tile_start, desc, masks, tcnt, bg, (N*)out);
    else if (bg_nan)  // float dtypes with a NaN background: one comparison less per pixel
        launch_apply(tile_apply_kernel<N, FN, TR, 0, NanBackground<N>::possible>, grid, THREADS, smem, s, P, T, tile_start, desc, masks, tcnt, bg, (N*)out);
    else
        launch_apply(tile_apply_kernel<N, FN, TR, 0, false>, grid, THREADS, smem, s, P, T, tile_start, desc, masks, tcnt, bg, (N*)out);
}
template <typename N> static TileLaunch tile_for_fn(int fn) {
    switch (fn) {
        case RZ_SUM: return tile_launch<N, RZ_SUM>;
        case RZ_FIRST: return tile_launch<N, RZ_FIRST>;
        case RZ_LAST: return tile_launch<N, RZ_LAST>;
        case RZ_MIN: return tile_launch<N, RZ_MIN>;
        case RZ_MAX: return tile_launch<N, RZ_MAX>;
        case RZ_COUNT: return tile_launch<N, RZ_COUNT>;
        case RZ_ANY: return tile_launch<N, RZ_ANY>;
    }
    return nullptr;
}
#if RZ_INST_HALF == 0
TileLaunch tile_for_lo(int dtype, int fn) {
    switch (dtype) {
        case RZ_U8: return tile_for_fn<uint8_t>(fn);
        case RZ_U16: return tile_for_fn<uint16_t>(fn);
        case RZ_U32: return tile_for_fn<uint32_t>(fn);
        case RZ_U64: return tile_for_fn<uint64_t>(fn);
        case RZ_I8: return tile_for_fn<int8_t>(fn);
    }
    return nullptr;
}
#else
TileLaunch tile_for_hi(int dtype, int fn) {
    switch (dtype) {
        case RZ_I16: return tile_for_fn<int16_t>(fn);
        case RZ_I32: return tile_for_fn<int32_t>(fn);
        case RZ_I64: return tile_for_fn<int64_t>(fn);
        case RZ_F32: return tile_for_fn<float>(fn);
        case RZ_F64: return tile_for_fn<double>(fn);
    }
    return nullptr;
}
#endif

}  // namespace rz
