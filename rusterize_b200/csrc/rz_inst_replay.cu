// rz_inst_replay.cu — replay_fill_kernel<N, FN> (SparseArray::build_array, rust/src/encoding/arrays.rs:103-143).
#include "rz_dispatch.hpp"
#include "rz_sparse.cuh"

namespace rz {

template <typename N, int FN>
static void replay_launch(dim3 grid, size_t smem, cudaStream_t s, FillParams F, const uint64_t* keys,
                          const uint32_t* task_start, const unsigned long long* cols, const void* data, uint32_t idx_bits,
                          uint64_t bg, void* out) {
    replay_fill_kernel<N, FN><<<grid, FILL_WARPS * 32, smem, s>>>(F, keys, task_start, cols, (const N*)data, idx_bits, bg,
                                                                 (N*)out);
}
template <typename N> static ReplayLaunch replay_for_fn(int fn) {
    switch (fn) {
        case RZ_SUM: return replay_launch<N, RZ_SUM>;
        case RZ_FIRST: return replay_launch<N, RZ_FIRST>;
        case RZ_LAST: return replay_launch<N, RZ_LAST>;
        case RZ_MIN: return replay_launch<N, RZ_MIN>;
        case RZ_MAX: return replay_launch<N, RZ_MAX>;
        case RZ_COUNT: return replay_launch<N, RZ_COUNT>;
        case RZ_ANY: return replay_launch<N, RZ_ANY>;
    }
    return nullptr;
}
ReplayLaunch replay_for(int dtype, int fn) {
    switch (dtype) {
        case RZ_U8: return replay_for_fn<uint8_t>(fn);
        case RZ_U16: return replay_for_fn<uint16_t>(fn);
        case RZ_U32: return replay_for_fn<uint32_t>(fn);
        case RZ_U64: return replay_for_fn<uint64_t>(fn);
        case RZ_I8: return replay_for_fn<int8_t>(fn);
        case RZ_I16: return replay_for_fn<int16_t>(fn);
        case RZ_I32: return replay_for_fn<int32_t>(fn);
        case RZ_I64: return replay_for_fn<int64_t>(fn);
        case RZ_F32: return replay_for_fn<float>(fn);
        case RZ_F64: return replay_for_fn<double>(fn);
    }
    return nullptr;
}

}  // namespace rz
