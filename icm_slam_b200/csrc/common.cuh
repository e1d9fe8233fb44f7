// common.cuh -- shared device helpers for libicmslam (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/icmslam.h"

#define ICM_PI 3.141592653589793   // np.pi
#define ICM_TWOPI (2.0 * ICM_PI)
#define ICM_HALFPI (ICM_PI / 2.0)

#define WARP 32
#define FULLMASK 0xffffffffu

// Device-side copy of the configuration, passed by value to kernels.
struct DevCfg {
    double dt, q1, q2, r1, r2, r3, kod, cota, dist_thr, rmax, radio;
    int L;
};

// Status word written by kernels (checked by the host after the sweep).
enum { ST_OK = 0, ST_LABEL_CAP = 1, /* 4: empty map */ ST_P2P_TIMEOUT = 8 };

// ---- exact (non-contracted) arithmetic: nvcc fuses a*b+c into DFMA by default; where a result
// must be bit-identical to numpy/scipy we spell out the roundings.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dadd_rn(a, -b); }

// Euclidean distance exactly as scipy cdist/pdist: sqrt(dx*dx + dy*dy), each op rounded.
__device__ __forceinline__ double dist2_rn(double dx, double dy) { return add_rn(mul_rn(dx, dx), mul_rn(dy, dy)); }
__device__ __forceinline__ double dist_rn(double dx, double dy) { return __dsqrt_rn(dist2_rn(dx, dy)); }

// entrepi (ICM_SLAM.py:455-463): np.mod(a, 2*pi) (result in [0, 2*pi)), then -2*pi if > pi.
__device__ __forceinline__ double entrepi(double a)
{
    double m = fmod(a, ICM_TWOPI);
    if (m < 0.0) m += ICM_TWOPI;
    if (m > ICM_PI) m -= ICM_TWOPI;
    return m;
}

// tras_rot_z (ICM_SLAM.py:465-480).  numpy's matmul accumulates acc = a0*b0; acc = fma(a1,b1,acc)
// (verified bit-for-bit against the reference on the build host, tests/golden/units.npz).
struct Rot { double ct, st; };   // cos/sin of (theta - pi/2)
__device__ __forceinline__ Rot make_rot(double theta)
{
    Rot r;
    sincos(sub_rn(theta, ICM_HALFPI), &r.st, &r.ct);
    return r;
}
__device__ __forceinline__ void project(const Rot& r, double px, double py, double bx, double by, double& wx, double& wy)
{
    wx = add_rn(__fma_rn(by, -r.st, mul_rn(bx, r.ct)), px);
    wy = add_rn(__fma_rn(by, r.ct, mul_rn(bx, r.st)), py);
}

// Projection parameters of a pose: (x, y, sin(theta - pi/2), cos(theta - pi/2)) exactly as tras_rot_z forms them
// (ICM_SLAM.py:466-468; the subtraction rounded once, then libm-accurate sin/cos).  The sin/cos of the heading itself are
// (cos(theta - pi/2), -sin(theta - pi/2)).  One record per pose, rewritten by the solve for the poses it moves.
__device__ __forceinline__ double4 make_ppar(double x, double y, double th)
{
    double st, ct;
    sincos(sub_rn(th, ICM_HALFPI), &st, &ct);
    return make_double4(x, y, st, ct);
}

__device__ __forceinline__ double4 ldg_ppar(const double4* p)
{
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULLMASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLMASK, v, o));
    return v;
}

// Uniform landmark grid used by the association (ICM_SLAM.py:169-172) and by Mapa.filtrar's
// nearest-neighbour search (ICM_SLAM.py:241-245).  Cell edge h >= dist_thr*(1+2^-20), so every
// landmark within dist_thr of a point lies in the 3x3 block of cells around the point's cell.
struct Grid {
    double x0, y0, inv_h;
    int nx, ny;
    const int* cell_start;   // nx*ny + 1
    const double* lx;        // landmarks sorted by cell
    const double* ly;
    const int* lidx;         // original landmark index
    int n;
};

__device__ __forceinline__ int grid_coord(double v, double v0, double inv_h)
{
    double f = floor((v - v0) * inv_h);
    // clamp into int range before converting; callers clamp to the grid afterwards
    f = fmin(fmax(f, -2.0e9), 2.0e9);
    return (int)f;
}

// Nearest landmark of (wx, wy) among those within the 3x3 neighbourhood; ties -> lowest original
// index (np.argmin picks the first minimum).  Returns the rooted distance in `best` (INFINITY if
// the neighbourhood is empty) and the original index in `arg`.
__device__ __forceinline__ void grid_nearest(const Grid& g, double wx, double wy, double& best, int& arg, double& lxb,
                                             double& lyb)
{
    best = INFINITY;
    arg = 0;
    lxb = 0.0;
    lyb = 0.0;
    int cx = grid_coord(wx, g.x0, g.inv_h), cy = grid_coord(wy, g.y0, g.inv_h);
    if (cx < -1 || cy < -1 || cx > g.nx || cy > g.ny) return;
    int c0 = max(cx - 1, 0), c1 = min(cx + 1, g.nx - 1);
    int r0 = max(cy - 1, 0), r1 = min(cy + 1, g.ny - 1);
    if (c0 > c1) return;
    for (int r = r0; r <= r1; ++r) {
        int s = __ldg(g.cell_start + r * g.nx + c0), e = __ldg(g.cell_start + r * g.nx + c1 + 1);
        for (int k = s; k < e; ++k) {
            double lx = __ldg(g.lx + k), ly = __ldg(g.ly + k);
            double dd = dist_rn(lx - wx, ly - wy);
            int id = __ldg(g.lidx + k);
            if (dd < best || (dd == best && id < arg)) { best = dd; arg = id; lxb = lx; lyb = ly; }
        }
    }
}
