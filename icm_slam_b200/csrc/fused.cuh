// fused.cuh -- single-kernel (REDBLACK, NEWTON, PREV) sweep.  (stub: generic path only for now)
#pragma once
#include "common.cuh"
#include "assoc.cuh"
#include "pose.cuh"

struct FusedWorkspace {
    double* d_xin = nullptr;   // 3 x T copy of the sweep's input poses
    bool available = false;
};

static void fused_free(FusedWorkspace& fw)
{
    if (fw.d_xin) { cudaFree(fw.d_xin); fw.d_xin = nullptr; }
}

static int fused_sweep(FusedWorkspace&, cudaStream_t, const DevCfg&, const PoseArrays&, const double*, const double*, int64_t,
                       DevState*, const int*, const double*, const double*, const int*, int*, int*, double*, double*, int*,
                       double, int, unsigned long long*, char*, size_t)
{
    return ICMSLAM_ERR_UNSUPPORTED;
}
