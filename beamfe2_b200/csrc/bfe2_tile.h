// bfe2_tile.h -- launcher of the static-only tile kernel (bfe2_tile.cu), called from bfe2_solve_fused.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bfe2_device.cuh"

struct bfe2_tile_args {
    long long B;
    int C;
    const double *coords, *props, *f_nodal, *r_local;   // dense layouts of bfe2_solve_fused; r_local may be NULL
    double *u, *endforces, *reactions;                   // endforces / reactions may be NULL
    int32_t* info;
};

// 0: launched; 1: this plan / batch is outside the tile kernel (the caller uses the group kernel); < 0: CUDA error
int bfe2_tile_static_launch(const bfe2::PlanDev& pd, const bfe2_tile_args& a, cudaStream_t stream);
