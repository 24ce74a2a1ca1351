#define RZ_INST_HALF 0
#include "rz_inst_fill.inl"
