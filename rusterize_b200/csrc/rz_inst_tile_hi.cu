#define RZ_INST_HALF 1
#include "rz_inst_tile.inl"
