// rz_dispatch.hpp — launchers of the dtype x pixel-function kernel families.  The 10 x 7 instantiations of each
// family are compiled in their own translation units (rz_inst_*.cu) so that the library builds in parallel.
#pragma once

#include <cuda_runtime.h>

#include "rz_kernels.cuh"
#include "rz_tiles.cuh"

namespace rz {

typedef void (*FillLaunch)(dim3, size_t, cudaStream_t, FillParams, const uint64_t*, const uint32_t*, const PartInfo*,
                           const uint8_t*, uint64_t, void*, AliasCtx);
typedef void (*TileLaunch)(cudaStream_t, KParams, TileParams, const uint32_t*, const BlockDesc*, const uint32_t*,
                           const TileCounters*, uint64_t, void*);
typedef void (*ReplayLaunch)(dim3, size_t, cudaStream_t, FillParams, const uint64_t*, const uint32_t*,
                             const unsigned long long*, const void*, uint32_t, uint64_t, void*);

// rz_inst_fill_*.cu
FillLaunch fill_for_lo(int dtype, int fn);  // u8 u16 u32 u64 i8
FillLaunch fill_for_hi(int dtype, int fn);  // i16 i32 i64 f32 f64
inline FillLaunch fill_for(int dtype, int fn) { return dtype <= RZ_I8 ? fill_for_lo(dtype, fn) : fill_for_hi(dtype, fn); }
// rz_inst_tile_*.cu
TileLaunch tile_for_lo(int dtype, int fn);
TileLaunch tile_for_hi(int dtype, int fn);
inline TileLaunch tile_for(int dtype, int fn) { return dtype <= RZ_I8 ? tile_for_lo(dtype, fn) : tile_for_hi(dtype, fn); }
// rz_inst_replay.cu
ReplayLaunch replay_for(int dtype, int fn);

}  // namespace rz
