// rz_inst_fill.inl — fill_kernel<N, FN, ALL_POLY> for one half of the dtypes (RZ_INST_HALF = 0: u8 u16 u32 u64 i8,
// 1: i16 i32 i64 f32 f64).  Included by rz_inst_fill_lo.cu / rz_inst_fill_hi.cu.
#include "rz_dispatch.hpp"

namespace rz {

template <typename N, int FN>
static void fill_launch(dim3 grid, size_t smem, cudaStream_t s, FillParams F, const uint64_t* keys,
                        const uint32_t* task_start, const PartInfo* info, const uint8_t* kind, uint64_t bg, void* out,
                        AliasCtx A) {
    if (F.all_poly)
        fill_kernel<N, FN, true><<<grid, FILL_WARPS * 32, smem, s>>>(F, keys, task_start, info, kind, bg, (N*)out, A);
    else
        fill_kernel<N, FN, false><<<grid, FILL_WARPS * 32, smem, s>>>(F, keys, task_start, info, kind, bg, (N*)out, A);
}

template <typename N> static FillLaunch fill_for_fn(int fn) {
    switch (fn) {
        case RZ_SUM: return fill_launch<N, RZ_SUM>;
        case RZ_FIRST: return fill_launch<N, RZ_FIRST>;
        case RZ_LAST: return fill_launch<N, RZ_LAST>;
        case RZ_MIN: return fill_launch<N, RZ_MIN>;
        case RZ_MAX: return fill_launch<N, RZ_MAX>;
        case RZ_COUNT: return fill_launch<N, RZ_COUNT>;
        case RZ_ANY: return fill_launch<N, RZ_ANY>;
    }
    return nullptr;
}

#if RZ_INST_HALF == 0
FillLaunch fill_for_lo(int dtype, int fn) {
    switch (dtype) {
        case RZ_U8: return fill_for_fn<uint8_t>(fn);
        case RZ_U16: return fill_for_fn<uint16_t>(fn);
        case RZ_U32: return fill_for_fn<uint32_t>(fn);
        case RZ_U64: return fill_for_fn<uint64_t>(fn);
        case RZ_I8: return fill_for_fn<int8_t>(fn);
    }
    return nullptr;
}
#else
FillLaunch fill_for_hi(int dtype, int fn) {
    switch (dtype) {
        case RZ_I16: return fill_for_fn<int16_t>(fn);
        case RZ_I32: return fill_for_fn<int32_t>(fn);
        case RZ_I64: return fill_for_fn<int64_t>(fn);
        case RZ_F32: return fill_for_fn<float>(fn);
        case RZ_F64: return fill_for_fn<double>(fn);
    }
    return nullptr;
}
#endif

}  // namespace rz
