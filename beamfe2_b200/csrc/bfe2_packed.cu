// bfe2_packed.cu -- second translation unit of bfe2.cu: the packed-record entry points (bfe2_pattern_*,
// bfe2_solve_fused_packed) and the PACKED instantiations of k_solve, compiled side by side with the rest.
#define BFE2_TU_PACKED 1
#include "bfe2.cu"
