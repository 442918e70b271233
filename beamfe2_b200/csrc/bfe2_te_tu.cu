// bfe2_te_tu.cu -- third translation unit of bfe2.cu: the fused kernel with the two-ended modal engine
// (bfe2_te.cuh) and its launcher bfe2_te_launch, compiled side by side with the rest.
#define BFE2_TU_TE 1
#include "bfe2.cu"
