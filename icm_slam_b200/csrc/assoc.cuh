// assoc.cuh -- landmark grid, data association, label assignment, landmark statistics
// (Mapa.actualizar Branch B, ICM_SLAM.py:167-194, for all scans of a sweep at once).
#pragma once
#include "common.cuh"

// Device-resident sweep state (no host round trip inside a sweep).
struct DevState {
    int lact;        // Mapa.landmarks_actuales
    int lsearch;     // min(lact, L_in): columns of map_in that are searched (ICM_SLAM.py:169)
    int lact0;       // lact at sweep start = first new label
    int raw_l;       // lact after association (before filtrar)
    int status;      // ST_*
    int n_far_scans;
    int kept;        // landmarks surviving cota
    int new_l;       // landmarks after filtrar
    int n_ind;       // landmarks with a neighbour closer than dist_thr
    int dirty_tiles; // record tiles that went through the association kernel this sweep (runs.cuh; filled on request)
    unsigned long long newton_iters;
    unsigned long long solved;
    // grid parameters of the association grid / filter grid
    double gx0, gy0, ginv_h;
    int gnx, gny;
    double fx0, fy0, finv_h;
    int fnx, fny;
    double f_extent;  // max bbox extent of the kept landmarks
    double cambio[3]; // calc_cambio of the sweep's map against the previous one, accumulated by the fast tail: min, max, sum
    int cambio_unres; // new landmarks whose nearest old landmark the tail could not certify (then cambio[] is not usable)
    int cambio_pad;
};

// ---- grid construction ---------------------------------------------------------------------
// bounding box of the first n points -> grid origin / dims (one block).  n is read from the
// device (pointer) so the whole sweep stays capturable.
__global__ void k_grid_setup(const double* __restrict__ px, const double* __restrict__ py, const int* __restrict__ n_ptr,
                             double dist_thr, int max_cells, double* gx0, double* gy0, double* ginv_h, int* gnx, int* gny,
                             double* extent)
{
    __shared__ double sh[4][32];
    const int n = *n_ptr;
    double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double x = px[i], y = py[i];
        mnx = fmin(mnx, x); mxx = fmax(mxx, x);
        mny = fmin(mny, y); mxy = fmax(mxy, y);
    }
    mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
    int w = threadIdx.x / WARP, l = threadIdx.x % WARP;
    if (l == 0) { sh[0][w] = mnx; sh[1][w] = mny; sh[2][w] = mxx; sh[3][w] = mxy; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x / WARP;
        mnx = l < nw ? sh[0][l] : INFINITY; mny = l < nw ? sh[1][l] : INFINITY;
        mxx = l < nw ? sh[2][l] : -INFINITY; mxy = l < nw ? sh[3][l] : -INFINITY;
        mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
        if (l == 0) {
            if (n <= 0) { mnx = mny = 0.0; mxx = mxy = 0.0; }
            double h = dist_thr * (1.0 + 9.5367431640625e-07);   // 1 + 2^-20
            if (!(h > 0.0)) h = 1.0;
            double ex = mxx - mnx, ey = mxy - mny;
            // grow the cell until the grid fits the fixed cell budget (still >= dist_thr: exact)
            for (;;) {   // (+1 slack per axis: the two floor() evaluations may differ by one at a boundary)
                double nxd = floor(ex / h) + 2.0, nyd = floor(ey / h) + 2.0;
                if (nxd * nyd <= (double)max_cells) break;
                h *= 1.5;
            }
            *gx0 = mnx; *gy0 = mny; *ginv_h = 1.0 / h;
            *gnx = grid_coord(mxx, mnx, 1.0 / h) + 1;
            *gny = grid_coord(mxy, mny, 1.0 / h) + 1;
            if (extent) *extent = fmax(ex, ey);
        }
    }
}

__device__ __forceinline__ int cell_of(double x, double y, double gx0, double gy0, double inv_h, int nx, int ny)
{
    int cx = min(max(grid_coord(x, gx0, inv_h), 0), nx - 1);
    int cy = min(max(grid_coord(y, gy0, inv_h), 0), ny - 1);
    return cy * nx + cx;
}

__global__ void k_cell_count(const double* __restrict__ px, const double* __restrict__ py, const int* __restrict__ n_ptr,
                             const double* gx0, const double* gy0, const double* ginv_h, const int* gnx, const int* gny,
                             int* __restrict__ cell_cnt, int* __restrict__ cell_id)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ptr) return;
    int c = cell_of(px[i], py[i], *gx0, *gy0, *ginv_h, *gnx, *gny);
    cell_id[i] = c;
    atomicAdd(cell_cnt + c, 1);
}

__global__ void k_cell_fill(const double* __restrict__ px, const double* __restrict__ py, const int* __restrict__ n_ptr,
                            const int* __restrict__ cell_id, const int* __restrict__ cell_start, int* __restrict__ cell_fill,
                            double* __restrict__ sx, double* __restrict__ sy, int* __restrict__ sidx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ptr) return;
    int c = cell_id[i];
    int p = cell_start[c] + atomicAdd(cell_fill + c, 1);
    sx[p] = px[i];
    sy[p] = py[i];
    sidx[p] = i;
}

__device__ __forceinline__ Grid load_grid(const double* gx0, const double* gy0, const double* ginv_h, const int* gnx,
                                          const int* gny, const int* cell_start, const double* lx, const double* ly,
                                          const int* lidx, int n)
{
    Grid g;
    g.x0 = *gx0; g.y0 = *gy0; g.inv_h = *ginv_h; g.nx = *gnx; g.ny = *gny;
    g.cell_start = cell_start; g.lx = lx; g.ly = ly; g.lidx = lidx; g.n = n;
    return g;
}

// ---- association ---------------------------------------------------------------------------
// One warp per scan, lanes over the scan's kept observations.  Projects with the sweep's INPUT
// pose (sensors.py:141,153), finds the nearest of the first `lsearch` previous-map landmarks
// (cdist + argmin, ICM_SLAM.py:169-171), gates at dist_thr (strict >, :172).  Matched
// observations add into the per-label sums; far ones are marked -1 and counted per scan.
__global__ void __launch_bounds__(256)
k_assoc(int T, const int* __restrict__ off, const double* __restrict__ bx, const double* __restrict__ by,
        const double* __restrict__ x, int64_t ldx, double x0x, double x0y, double x0t, DevState* st,
        const int* __restrict__ cell_start, const double* __restrict__ glx, const double* __restrict__ gly,
        const int* __restrict__ gidx, double dist_thr, int* __restrict__ c, int* __restrict__ nfar,
        double* __restrict__ sum_x, double* __restrict__ sum_y, int* __restrict__ cnt)
{
    const int lane = threadIdx.x % WARP;
    const int wpb = blockDim.x / WARP;
    Grid g = load_grid(&st->gx0, &st->gy0, &st->ginv_h, &st->gnx, &st->gny, cell_start, glx, gly, gidx, st->lsearch);
    for (int t = blockIdx.x * wpb + threadIdx.x / WARP; t < T; t += gridDim.x * wpb) {
        int o = off[t], e = off[t + 1];
        int far = 0;
        if (e > o) {
            double px, py, th;
            if (t == 0) { px = x0x; py = x0y; th = x0t; }
            else { px = x[t]; py = x[ldx + t]; th = x[2 * ldx + t]; }
            Rot r = make_rot(th);
            for (int i = o + lane; i < e; i += WARP) {
                double wx, wy, best, lxb, lyb;
                int arg;
                project(r, px, py, bx[i], by[i], wx, wy);
                if (g.n > 0) grid_nearest(g, wx, wy, best, arg, lxb, lyb);
                else { best = INFINITY; arg = 0; }
                if (best > dist_thr) { c[i] = -1; ++far; }
                else {
                    c[i] = arg;
                    atomicAdd(sum_x + arg, wx);
                    atomicAdd(sum_y + arg, wy);
                    atomicAdd(cnt + arg, 1);
                }
            }
            far = warp_sum_i(far);
        }
        if (lane == 0) nfar[t] = far;
    }
}

// flag[t] = scan t has >= 1 far observation (each such scan creates exactly ONE new label,
// ICM_SLAM.py:174-182).  Also used to turn counts into flags for cub scans.
__global__ void k_flag_positive(const int* __restrict__ v, int n, int* __restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = v[i] > 0 ? 1 : 0;
}

// New labels: label(t) = lact0 + (number of earlier scans with a far observation).  One warp per
// scan that has far observations: rewrite c, and own the new label's statistics (no atomics).
__global__ void __launch_bounds__(256)
k_new_labels(int T, const int* __restrict__ off, const double* __restrict__ bx, const double* __restrict__ by,
             const double* __restrict__ x, int64_t ldx, double x0x, double x0y, double x0t, DevState* st,
             const int* __restrict__ nfar, const int* __restrict__ far_prefix, int Lcap, int* __restrict__ c,
             double* __restrict__ sum_x, double* __restrict__ sum_y, int* __restrict__ cnt)
{
    const int lane = threadIdx.x % WARP;
    const int wpb = blockDim.x / WARP;
    const int lact0 = st->lact0;
    for (int t = blockIdx.x * wpb + threadIdx.x / WARP; t < T; t += gridDim.x * wpb) {
        if (t == T - 1 && lane == 0) {
            int total = far_prefix[t] + (nfar[t] > 0 ? 1 : 0);
            st->n_far_scans = total;
            st->raw_l = lact0 + total;
            if (lact0 + total > Lcap) st->status = ST_LABEL_CAP;   // IndexError at ICM_SLAM.py:191
        }
        if (nfar[t] == 0) continue;
        int label = lact0 + far_prefix[t];
        if (label >= Lcap) continue;
        int o = off[t], e = off[t + 1];
        double px, py, th;
        if (t == 0) { px = x0x; py = x0y; th = x0t; }
        else { px = x[t]; py = x[ldx + t]; th = x[2 * ldx + t]; }
        Rot r = make_rot(th);
        double sx = 0.0, sy = 0.0;
        for (int i0 = o; i0 < e; i0 += WARP) {       // sequential-in-lane-order sum of the far points
            int i = i0 + lane;
            double wx = 0.0, wy = 0.0;
            bool isfar = false;
            if (i < e && c[i] < 0) {
                project(r, px, py, bx[i], by[i], wx, wy);
                c[i] = label;
                isfar = true;
            }
            unsigned bal = __ballot_sync(FULLMASK, isfar);
            while (bal) {                              // np.sum(obs[c==i], axis=0): in row order
                int src = __ffs(bal) - 1;
                sx += __shfl_sync(FULLMASK, wx, src);
                sy += __shfl_sync(FULLMASK, wy, src);
                bal &= bal - 1;
            }
        }
        if (lane == 0) { sum_x[label] = sx; sum_y[label] = sy; cnt[label] = nfar[t]; }
    }
}

// raw map = per-label mean (the value the recursive running mean of ICM_SLAM.py:191-194 ends at);
// labels never observed keep the zero column of `y` (sensors.py:132).
__global__ void k_means(const DevState* st, const double* __restrict__ sum_x, const double* __restrict__ sum_y,
                        const int* __restrict__ cnt, double* __restrict__ raw_x, double* __restrict__ raw_y, int Lcap)
{
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= Lcap) return;
    int k = l < st->raw_l ? cnt[l] : 0;
    raw_x[l] = k > 0 ? sum_x[l] / (double)k : 0.0;
    raw_y[l] = k > 0 ? sum_y[l] / (double)k : 0.0;
}

// ---- running view --------------------------------------------------------------------------
// Reference semantics of what a pose sees (sensors.py:154-156): the recursive running mean of its
// label INCLUDING the current scan.  Observations are sorted by (label, time) (stable radix sort
// of labels; CSR order is time order); one thread per label walks its segment, grouping by scan
// and applying ICM_SLAM.py:191-194 literally.  Also produces the reference-arithmetic raw map.
__global__ void k_running_mean(const DevState* st, const int* __restrict__ seg_start, const int* __restrict__ cnt,
                               const int* __restrict__ sorted_obs, const int* __restrict__ scan_of,
                               const double* __restrict__ bx, const double* __restrict__ by, const double* __restrict__ x,
                               int64_t ldx, double x0x, double x0y, double x0t, double* __restrict__ seen_x,
                               double* __restrict__ seen_y, double* __restrict__ raw_x, double* __restrict__ raw_y, int Lcap)
{
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= Lcap) return;
    if (l >= st->raw_l || cnt[l] == 0) { raw_x[l] = 0.0; raw_y[l] = 0.0; return; }
    int s = seg_start[l], e = s + cnt[l];
    double mx = 0.0, my = 0.0, ni = 0.0;
    int j = s;
    while (j < e) {
        int t = scan_of[sorted_obs[j]];
        double px, py, th;
        if (t == 0) { px = x0x; py = x0y; th = x0t; }
        else { px = x[t]; py = x[ldx + t]; th = x[2 * ldx + t]; }
        Rot r = make_rot(th);
        double sx = 0.0, sy = 0.0;
        int k = 0, j2 = j;
        while (j2 < e && scan_of[sorted_obs[j2]] == t) {
            int i = sorted_obs[j2];
            double wx, wy;
            project(r, px, py, bx[i], by[i], wx, wy);
            sx = add_rn(sx, wx); sy = add_rn(sy, wy);
            ++k; ++j2;
        }
        double tot = ni + (double)k;
        mx = add_rn(__ddiv_rn(sx, tot), __ddiv_rn(mul_rn(mx, ni), tot));
        my = add_rn(__ddiv_rn(sy, tot), __ddiv_rn(mul_rn(my, ni), tot));
        ni = tot;
        for (int q = j; q < j2; ++q) { seen_x[sorted_obs[q]] = mx; seen_y[sorted_obs[q]] = my; }
        j = j2;
    }
    raw_x[l] = mx; raw_y[l] = my;
}

// labels -> sort keys (far observations already relabelled)
__global__ void k_iota(int n, int* __restrict__ v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
