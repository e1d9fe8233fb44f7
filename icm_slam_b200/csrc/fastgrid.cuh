// fastgrid.cuh -- the landmark grid of the fused sweep path.
//
// Same contract as the grid in common.cuh (every landmark within dist_thr of a query point is
// visited, ties resolved exactly as cdist + np.argmin do, ICM_SLAM.py:169-172) with a layout made
// for the hot loop:
//   * cell edge h >= 2*thr1 (thr1 = dist_thr*(1+2^-20)), so the disc of radius dist_thr around a
//     query overlaps at most 2 x 2 cells: two cell_start loads per row instead of the 3 x 3 block;
//   * a FIXED cell budget (NC cells, host constant) so the counting sort's scan has a host-known
//     length and no host round trip: the geometry kernel grows h until nx*ny <= NC;
//   * landmarks stored cell-sorted as double2 (one 16-byte load per candidate); the original index
//     is only fetched for the winner (and for exact tie resolution);
//   * squared distances are compared; the rooted values the reference compares are only formed when
//     two candidates are within 2^-50 relative (sqrt_rn is monotone, so this is exact), and the
//     gate `amin > dist_thr` becomes `s > thr2_hi` with thr2_hi the largest double whose rooted
//     value is <= dist_thr (computed on the host).
// Points outside the bounding box clamp to the border cells on both the build and the query side,
// so a grid whose geometry was fixed for an earlier map stays exact for a later one.
#pragma once
#include "common.cuh"

struct FGeom {
    double x0, y0, inv_h, delta;   // delta = thr1 * inv_h  (< 0.5)
    int nx, ny;
};

struct FGrid {
    FGeom g;
    const int* cell_start;     // nx*ny + 1 (within the NC + 2 allocation)
    const double2* pts;        // cell-sorted coordinates
    const int* idx;            // cell-sorted original indices
};

__device__ __forceinline__ unsigned long long dkey(double v)   // order-preserving double -> u64
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dkey_inv(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// bb[0..3] = min x, min y, max x, max y as ordered keys; reset by k_fgrid_bbox_reset
__global__ void k_fgrid_bbox_reset(unsigned long long* bb)
{
    bb[0] = bb[1] = ~0ull;
    bb[2] = bb[3] = 0ull;
}

__global__ void __launch_bounds__(256)
k_fgrid_bbox(const double* __restrict__ px, const double* __restrict__ py, const int* __restrict__ n_ptr, unsigned long long* bb)
{
    const int n = *n_ptr;
    double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x = px[i], y = py[i];
        mnx = fmin(mnx, x); mxx = fmax(mxx, x);
        mny = fmin(mny, y); mxy = fmax(mxy, y);
    }
    mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
    if ((threadIdx.x % WARP) == 0 && mnx <= mxx) {
        atomicMin(bb + 0, dkey(mnx)); atomicMin(bb + 1, dkey(mny));
        atomicMax(bb + 2, dkey(mxx)); atomicMax(bb + 3, dkey(mxy));
    }
}

__global__ void k_fgrid_geom(const unsigned long long* bb, const int* n_ptr, double dist_thr, int max_cells, FGeom* out)
{
    double mnx = 0.0, mny = 0.0, mxx = 0.0, mxy = 0.0;
    if (*n_ptr > 0 && bb[0] != ~0ull) { mnx = dkey_inv(bb[0]); mny = dkey_inv(bb[1]); mxx = dkey_inv(bb[2]); mxy = dkey_inv(bb[3]); }
    const double thr1 = dist_thr * (1.0 + 9.5367431640625e-07);   // 1 + 2^-20
    double h = 2.0 * thr1 * (1.0 + 9.5367431640625e-07);
    if (!(h > 0.0)) h = 1.0;
    const double ex = mxx - mnx, ey = mxy - mny;
    for (;;) {
        double nxd = floor(ex / h) + 2.0, nyd = floor(ey / h) + 2.0;
        if (nxd * nyd <= (double)max_cells) break;
        h *= 1.25;
    }
    FGeom g;
    g.x0 = mnx; g.y0 = mny; g.inv_h = 1.0 / h; g.delta = thr1 * g.inv_h;
    g.nx = __double2int_rd((mxx - mnx) * g.inv_h) + 1;
    g.ny = __double2int_rd((mxy - mny) * g.inv_h) + 1;
    if (g.nx < 1) g.nx = 1;
    if (g.ny < 1) g.ny = 1;
    *out = g;
}

__device__ __forceinline__ int fgrid_cell(const FGeom& g, double x, double y)
{
    int cx = __double2int_rd((x - g.x0) * g.inv_h), cy = __double2int_rd((y - g.y0) * g.inv_h);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    return cy * g.nx + cx;
}

__global__ void __launch_bounds__(256)
k_fgrid_count(const double* __restrict__ px, const double* __restrict__ py, const int* __restrict__ n_ptr, const FGeom* __restrict__ geom,
              int* __restrict__ cell_cnt, int* __restrict__ cell_id)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ptr) return;
    const FGeom g = *geom;
    const int c = fgrid_cell(g, px[i], py[i]);
    cell_id[i] = c;
    atomicAdd(cell_cnt + c, 1);
}

__global__ void __launch_bounds__(256)
k_fgrid_fill(const double* __restrict__ px, const double* __restrict__ py, const int* __restrict__ n_ptr, const int* __restrict__ cell_id,
             const int* __restrict__ cell_start, int* __restrict__ cell_fill, double2* __restrict__ pts, int* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ptr) return;
    const int c = cell_id[i];
    const int p = cell_start[c] + atomicSub(cell_fill + c, 1) - 1;   // counts the cell back down to zero
    pts[p] = make_double2(px[i], py[i]);
    idx[p] = i;
}

// Nearest landmark of (wx, wy).  Returns the winner's position in the cell-sorted arrays (-1 if no
// candidate), its squared distance `best` (INFINITY if none) and its coordinates.
__device__ __forceinline__ int fgrid_nearest(const FGrid& G, double wx, double wy, double& best, double& lx, double& ly)
{
    const double fx = (wx - G.g.x0) * G.g.inv_h, fy = (wy - G.g.y0) * G.g.inv_h;
    int cx0 = __double2int_rd(fx - G.g.delta), cx1 = __double2int_rd(fx + G.g.delta);
    int cy0 = __double2int_rd(fy - G.g.delta), cy1 = __double2int_rd(fy + G.g.delta);
    // clamp BOTH ends into the grid: landmarks outside the box were binned into the border cells
    cx0 = min(max(cx0, 0), G.g.nx - 1); cx1 = min(max(cx1, 0), G.g.nx - 1);
    cy0 = min(max(cy0, 0), G.g.ny - 1); cy1 = min(max(cy1, 0), G.g.ny - 1);
    best = INFINITY;
    lx = 0.0; ly = 0.0;
    int bk = -1;
    for (int r = cy0; r <= cy1; ++r) {
        const int s = __ldg(G.cell_start + r * G.g.nx + cx0), e = __ldg(G.cell_start + r * G.g.nx + cx1 + 1);
        for (int k = s; k < e; ++k) {
            const double2 p = __ldg(G.pts + k);
            const double s2 = dist2_rn(p.x - wx, p.y - wy);
            bool take;
            if (bk < 0) take = true;
            else {
                const double lo = best * (1.0 - 8.8817841970012523e-16), hi = best * (1.0 + 8.8817841970012523e-16);   // 2^-50
                if (s2 < lo) take = true;
                else if (s2 > hi) take = false;
                else {   // (almost) equidistant: decide on the rooted values like np.argmin over cdist
                    const double dk = __dsqrt_rn(s2), db = __dsqrt_rn(best);
                    take = dk < db || (dk == db && __ldg(G.idx + k) < __ldg(G.idx + bk));
                }
            }
            if (take) { best = s2; bk = k; lx = p.x; ly = p.y; }
        }
    }
    return bk;
}
