// tail_coop.cuh -- the landmark update + Mapa.filtrar + grid of the new map (tail.cuh, part C) as ONE
// cooperative launch: the nine small kernels of the chain become phases separated by grid-wide
// barriers, which costs a few microseconds instead of a kernel boundary each (the chain moves ~1 MB,
// so it is pure launch latency).  Same arithmetic, same outputs as the kernel chain, which stays as the
// fallback when a cooperative launch of this size is not possible.
#pragma once
#include <cooperative_groups.h>

#include "tail.cuh"
#include "fused.cuh"

namespace cg = cooperative_groups;

#define TC_THREADS 512

struct TailCoopParams {
    DevState* st; TailState* ts;
    long long* fsum_x; long long* fsum_y; int* cnt;
    const double* map_x; const double* map_y;          // previous map rows
    double inv_scale, cota, dist_thr, thr2_lt, thr1sq, thr2_hi;
    double* newraw; double* raw_x; double* raw_y;
    int* kflag; int* kpos;
    double* kx; double* ky; double* kc; int* parent;
    unsigned long long* bb;
    int max_cells; FGeom* geom; int* cell_cnt; int* cell_start; double2* pts; int* gidx;
    int* nn; int* ind_flag; double* nnd2;
    double* map_out; int cap_out; int64_t ld_out; double* counts_state;
    LmRec* lmrec; int* remap;
    // merge path
    int* ind_pos; int* ind; int* lab; int* used; int* rank; double* ox; double* oy; double* oc;
    int* blk_scratch;                                  // >= 2 * gridDim.x ints
    int Lcap;
};

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = block total
__device__ __forceinline__ int tc_block_scan(int v, int* wsum /* >= 32 ints of shared memory */, int* total)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += o; }
    __syncthreads();
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int x = lane < nw ? wsum[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(FULLMASK, x, d); if (lane >= d) x += o; }
        wsum[lane] = x;
    }
    __syncthreads();
    *total = wsum[nw - 1];
    return inc - v + (warp ? wsum[warp - 1] : 0);
}

// sum over blocks b' < b of per_block[b'] (every thread of the block gets it)
__device__ __forceinline__ int tc_blocks_before(const int* per_block, int b, int* sh)
{
    int s = 0;
    for (int i = threadIdx.x; i < b; i += blockDim.x) s += per_block[i];
    s = warp_sum_i(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tail_coop(const TailCoopParams p)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ int sh[40];
    __shared__ FGeom sgeom;
    const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x, nb = gridDim.x;
    const int L = p.Lcap;
    DevState* st = p.st;
    TailState* ts = p.ts;
    int* blkcnt = p.blk_scratch;
    int* cellsum = p.blk_scratch + nb;
    // ---- phase 1: means + keep flags (k_fused_means); each block owns a contiguous chunk of labels -------------
    const int chunk = (L + nb - 1) / nb;
    const int l0 = min(b * chunk, L), l1 = min(l0 + chunk, L);
    const int raw_l = st->raw_l, ls = st->lsearch;
    int mycnt = 0;
    for (int l = l0 + tid; l < l1; l += nth) {
        const int k = l < raw_l ? p.cnt[l] : 0;
        double rx, ry;
        if (l < ls) {
            rx = k > 0 ? p.map_x[l] + ((double)p.fsum_x[l] * p.inv_scale) / (double)k : 0.0;
            ry = k > 0 ? p.map_y[l] + ((double)p.fsum_y[l] * p.inv_scale) / (double)k : 0.0;
        } else {
            const bool have = l < raw_l && k > 0;
            rx = have ? p.newraw[l] : 0.0;
            ry = have ? p.newraw[L + l] : 0.0;
        }
        p.raw_x[l] = rx; p.raw_y[l] = ry;
        p.newraw[l] = 0.0; p.newraw[L + l] = 0.0;
        p.fsum_x[l] = 0; p.fsum_y[l] = 0;
        const int f = (l < raw_l && !((double)k < p.cota)) ? 1 : 0;      // ICM_SLAM.py:232-236
        p.kflag[l] = f;
        mycnt += f;
    }
    {
        int tot;
        tc_block_scan(mycnt, sh, &tot);
        if (tid == 0) blkcnt[b] = tot;
    }
    if (b == 0 && tid == 0) { ts->n_ind = 0; ts->degenerate = 0; p.bb[0] = p.bb[1] = ~0ull; p.bb[2] = p.bb[3] = 0ull; }
    grid.sync();
    // ---- phase 2: ordered compaction of the survivors + bounding box (k_tail_compact) ----------------------------
    {
        int base = tc_blocks_before(blkcnt, b, sh);
        double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
        for (int l = l0; l < l1; l += nth) {
            const int i = l + tid;
            const int f = i < l1 ? p.kflag[i] : 0;
            int tot;
            const int pos = base + tc_block_scan(f, sh, &tot);
            if (i < l1) p.kpos[i] = pos;
            if (f) {
                const double x = p.raw_x[i], y = p.raw_y[i];
                p.kx[pos] = x; p.ky[pos] = y; p.kc[pos] = (double)p.cnt[i];
                p.parent[pos] = pos;
                mnx = fmin(mnx, x); mxx = fmax(mxx, x); mny = fmin(mny, y); mxy = fmax(mxy, y);
            }
            base += tot;
        }
        mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
        if ((tid & 31) == 0 && mnx <= mxx) {
            atomicMin(p.bb + 0, dkey(mnx)); atomicMin(p.bb + 1, dkey(mny));
            atomicMax(p.bb + 2, dkey(mxx)); atomicMax(p.bb + 3, dkey(mxy));
        }
        if (b == nb - 1 && tid == 0) {
            st->kept = base;
            st->n_ind = 0;
            if (base == 0) st->status |= 4;   // ValueError in the reference (ICM_SLAM.py:241-255)
        }
    }
    grid.sync();
    // ---- phase 3: grid geometry (every block derives the same one) + cell counts (k_tail_geom, k_fgrid_count) ----
    const int K = st->kept;
    if (tid == 0) {
        double mnx = 0.0, mny = 0.0, mxx = 0.0, mxy = 0.0;
        if (K > 0 && p.bb[0] != ~0ull) { mnx = dkey_inv(p.bb[0]); mny = dkey_inv(p.bb[1]); mxx = dkey_inv(p.bb[2]); mxy = dkey_inv(p.bb[3]); }
        sgeom = fgrid_make_geom(mnx, mny, mxx, mxy, p.dist_thr, p.max_cells);
        if (b == 0) { *p.geom = sgeom; ts->degenerate = (fmax(mxx - mnx, mxy - mny) >= p.dist_thr) ? 0 : 1; }
    }
    __syncthreads();
    const FGeom g = sgeom;
    for (int j = b * nth + tid; j < K; j += nb * nth) {
        int cx0, cx1, cy0, cy1;
        fgrid_cell_range(g, p.kx[j], p.ky[j], cx0, cx1, cy0, cy1);
        for (int cy = cy0; cy <= cy1; ++cy)
            for (int cx = cx0; cx <= cx1; ++cx) atomicAdd(p.cell_cnt + cy * g.nx + cx, 1);
    }
    grid.sync();
    // ---- phase 4/5: exclusive scan of the cell counts (two levels over the blocks) ---------------------------------
    const int ncell1 = g.nx * g.ny + 1;
    const int cchunk = (ncell1 + nb - 1) / nb;
    const int c0 = min(b * cchunk, ncell1), c1 = min(c0 + cchunk, ncell1);
    {
        int s = 0;
        for (int c = c0 + tid; c < c1; c += nth) s += p.cell_cnt[c];
        int tot;
        tc_block_scan(s, sh, &tot);
        if (tid == 0) cellsum[b] = tot;
    }
    grid.sync();
    {
        int base = tc_blocks_before(cellsum, b, sh);
        for (int c = c0; c < c1; c += nth) {
            const int i = c + tid;
            const int v = i < c1 ? p.cell_cnt[i] : 0;
            int tot;
            const int pos = base + tc_block_scan(v, sh, &tot);
            if (i < c1) p.cell_start[i] = pos;
            base += tot;
        }
    }
    grid.sync();
    // ---- phase 6: fill (k_fgrid_fill) --------------------------------------------------------------------------------
    for (int j = b * nth + tid; j < K; j += nb * nth) {
        const double x = p.kx[j], y = p.ky[j];
        int cx0, cx1, cy0, cy1;
        fgrid_cell_range(g, x, y, cx0, cx1, cy0, cy1);
        for (int cy = cy0; cy <= cy1; ++cy)
            for (int cx = cx0; cx <= cx1; ++cx) {
                const int c = cy * g.nx + cx;
                const int q = p.cell_start[c] + atomicSub(p.cell_cnt + c, 1) - 1;
                p.pts[q] = make_double2(x, y);
                p.gidx[q] = j;
            }
    }
    grid.sync();
    // ---- phase 7: nearest other survivor (k_tail_nn) -----------------------------------------------------------------
    const int degenerate = ts->degenerate;
    for (int j = b * nth + tid; j < L; j += nb * nth) {
        if (j >= K || degenerate) { p.ind_flag[j] = 0; p.nnd2[j] = 0.0; continue; }
        const double xj = p.kx[j], yj = p.ky[j];
        const int c = fgrid_cell(g, xj, yj);
        const int s = p.cell_start[c], e = p.cell_start[c + 1];
        double best = INFINITY, lo = INFINITY, hi = INFINITY;
        int arg = -1;
        for (int k = s; k < e; ++k) {
            const double2 q = p.pts[k];
            const int id = p.gidx[k];
            const double s2 = dist2_rn(q.x - xj, q.y - yj);
            if (id == j || s2 == 0.0) continue;
            bool take = s2 < lo;
            if (!take && s2 <= hi && arg >= 0) {
                const double dk = __dsqrt_rn(s2), db = __dsqrt_rn(best);
                take = dk < db || (dk == db && id < arg);
            }
            if (take) { best = s2; arg = id; lo = s2 * (1.0 - 8.8817841970012523e-16); hi = s2 * (1.0 + 8.8817841970012523e-16); }
        }
        const int f = (arg >= 0 && best <= p.thr2_lt) ? 1 : 0;      // amin < dist_thr (strict, :245)
        p.nn[j] = arg < 0 ? 0 : arg;
        p.ind_flag[j] = f;
        p.nnd2[j] = best;
        if (f) atomicAdd(&ts->n_ind, 1);
    }
    grid.sync();
    // ---- phase 8: the new map (k_tail_finalize), or the one-block merge path (k_tail_slow) ----------------------------
    if (ts->n_ind == 0 && !degenerate) {
        for (int r = b * nth + tid; r < L; r += nb * nth) {
            const double c = r < K ? p.kc[r] : 0.0;
            const double mx = r < K ? mul_rn(p.kx[r], c) / c : 0.0, my = r < K ? mul_rn(p.ky[r], c) / c : 0.0;
            if (r < p.cap_out) { p.map_out[r] = mx; p.map_out[p.ld_out + r] = my; }
            p.counts_state[r] = c;
            LmRec rec;
            rec.x = mx; rec.y = my; rec.r2 = r < K ? hint_radius2(p.nnd2[r], p.thr1sq, p.thr2_hi) : 0.0; rec.g = 0.0;
            p.lmrec[r] = rec;
            p.remap[r] = (r < raw_l && p.kflag[r]) ? p.kpos[r] : -1;
            if (r == 0) {
                st->new_l = K; st->lact = K; st->n_ind = 0;
                ts->remap_identity = (K == ls && raw_l == ls) ? 1 : 0;
            }
        }
    } else if (b == 0) {
        tail_slow_body(st, ts, p.dist_thr, p.kx, p.ky, p.kc, p.parent, p.nn, p.ind_flag, p.ind_pos, p.ind, p.lab, p.used, p.rank, p.ox, p.oy,
                       p.oc, p.map_out, p.cap_out, p.ld_out, p.counts_state, L, p.max_cells, p.geom, p.cell_cnt, p.cell_start, p.pts, p.gidx,
                       p.kflag, p.kpos, p.thr1sq, p.thr2_hi, p.lmrec, p.remap);
    }
}
