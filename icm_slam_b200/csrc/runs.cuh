// runs.cuh -- run records and the steady-state sweep kernel k_runs of the (REDBLACK, NEWTON, PREV) sweep.
//
// A RUN is a maximal group of consecutive kept beams of one scan that hit the same landmark (a trunk seen under ~3 beams).
// Everything the sweep needs from a run's observations is LINEAR in the sum of their body-frame points: with
// R = Rot(theta_t - pi/2), p = pose position, y = landmark (tras_rot_z ICM_SLAM.py:465-480, Mapa.actualizar :191-194, the
// energy's observation term sensors.py:194-204),
//     sum_i (w_i - y)            = R sum(b) + n (p - y)            -> the landmark update (running mean == sum / count)
//     sum_i (y - p), (y - p) (x) b_i = n (y - p), (y - p) (x) sum(b)   -> the six landmark moments of the pose's normal equations.
// So a run is kept as ONE 32-byte record (sum(b), n, label, a bound rho on |b_i - centroid|), built by the association
// kernel (assoc_tiles.cuh) whenever labels are (re)computed, and a sweep whose labels did not change reads ~0.31 records
// per observation instead of the observations themselves.
//
// "Did not change" is PROVEN every sweep, run by run, against the current poses and map: the landmark record carries the
// radius r inside which an observation is provably nearest to that landmark and inside the gate (tail.cuh hint_radius2);
// |w_i - y| <= |w_c - y| + |b_i - c| <= |S| / n + rho with S = sum_i (w_i - y), so |S| <= n (r - rho) certifies every label
// of the run exactly as the per-observation hint test would.  A run that fails (or a far run, whose scan creates a label)
// marks its scans dirty; the association kernel re-associates dirty tiles observation by observation in the same sweep.
//
// Records of a tile of RT_TILE scans are packed into CHUNKS of 32 (one per lane of a warp); a scan's runs are contiguous
// in one chunk (a scan with more than 32 runs starts a chunk and continues through the next ones: its pieces meet through
// a ticket).  Per chunk a warp
//   * loads 32 records (1 KB, coalesced), gathers each run's pose parameters and landmark record,
//   * validates (steady state), forms S and the six moments of every run,
//   * sums the moments over the runs of each scan with a segmented suffix scan (runs of a scan are adjacent lanes; the
//     tree depends only on the scan's run count, so the result does not depend on the packing or tiling) -- the scan's
//     first lane writes them for the solve (dyn) --,
//   * sums the fixed-point statistics over the lanes that name the same landmark (__match_any + pointer jumping) and adds
//     one int64 pair + count per (chunk, landmark) to the global statistics: integer addition is associative, so the
//     landmark update is bit-reproducible for any tiling or GPU count.
// No shared memory, no block barrier: the kernel is a stream of independent warps.
#pragma once
#include "common.cuh"
#include "assoc.cuh"
#include "tail.cuh"

#define RUN_PAD (-2)        // padding slot of a chunk
#define RUN_FAR (-1)        // run of far observations (its scan creates a new label, ICM_SLAM.py:172-182)
#define RF_LEADER 1         // first run of its scan (piece) in the chunk
#define RF_LONG 2           // the scan has more than 32 runs: this chunk holds one piece of it
#define RF_HALO 4           // the halo scan t_lo-1 of a time segment: moments only (its statistics belong to the previous segment)
#define RT_TILE 128         // scans per record tile
#define RT_RHO_UNIT 2048.0  // rho is staged as a 16-bit count of 1/2048 m (rounded up) before it becomes a float

struct __align__(32) RunRec {
    double sbx, sby;        // sum of the run's body-frame points
    int label;              // landmark index in the current map, RUN_FAR or RUN_PAD
    float rho;              // >= max_i |b_i - centroid| (+inf: not provable)
    unsigned short n;       // beams
    unsigned char lt;       // scan, relative to the tile's first scan
    unsigned char rem;      // runs of the same scan (piece) after this one in the chunk
    unsigned char piece;    // long scans: index of this chunk among the scan's chunks
    unsigned char flags;    // RF_*
    unsigned short npieces; // long scans: number of chunks the scan spans
};

__device__ __forceinline__ int64_t run_tile_base(const int* off, int t_start, int tile)
{
    // slot capacity of a tile = 2 * (its observations) + 33 >= 32 * (its chunks) for the greedy packing of assoc_tiles.cuh
    return (((int64_t)2 * off[t_start + tile * RT_TILE]) & ~(int64_t)31) + (int64_t)64 * tile;
}

struct RunParams {
    int t_start, t_hi;                    // record tiles cover scans [t_start, t_hi): t_start = t_lo, or t_lo - 1 with a halo scan
    const int* off;                       // CSR offsets of the kept observations (T + 1)
    const double4* ppar;                  // projection parameters of the input poses (solve.cuh make_ppar)
    const LmRec* lmrec;                   // landmarks by label: position, hint radius^2, hint radius
    RunRec* rec;                          // run records, tile regions at run_tile_base
    int* tile_nchunks; int* tile_epoch;   // valid chunks of each tile / label-numbering epoch its records belong to
    double* dyn;                          // 6 doubles per pose: (Yx, Mxx, Mxy, Yy, Myx, Myy)
    double* dynx; int* ticket;            // long scans: partial sums per chunk (12 doubles) / arrival counter per scan
    long long* fsum_x; long long* fsum_y; int* cnt;   // per previous-map landmark statistics (fixed point)
    double fix_scale;
    FarRec* far_list; TailState* ts; unsigned* farbits;   // scans with far observations (each creates one label)
    int* scan_dirty; int* tile_flag; int* dirty_list;     // steady state: what failed validation (tile_flag: 1 some scans, 2 all)
};

// The run's statistics S = sum_i (w_i - y) in fixed point, summed over the lanes of the warp that name the same landmark;
// the lowest such lane adds the total to the global sums.
__device__ __forceinline__ void chunk_statistics(const RunParams& p, bool act, int label, double Sx, double Sy, int n)
{
    const int lane = threadIdx.x & 31;
    long long vx = act ? __double2ll_rn(Sx * p.fix_scale) : 0ll, vy = act ? __double2ll_rn(Sy * p.fix_scale) : 0ll;
    int cn = act ? n : 0;
    const unsigned peers = __match_any_sync(FULLMASK, act ? label : (int)(0x80000000u | (unsigned)lane));
    const unsigned above = lane == 31 ? 0u : (peers >> (lane + 1));
    int nxt = above ? lane + 1 + (__ffs(above) - 1) : lane;          // next lane of my list (self: I am its tail)
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
        if (!__any_sync(FULLMASK, nxt != lane)) break;
        const long long ox = __shfl_sync(FULLMASK, vx, nxt), oy = __shfl_sync(FULLMASK, vy, nxt);
        const int oc = __shfl_sync(FULLMASK, cn, nxt), on = __shfl_sync(FULLMASK, nxt, nxt);
        if (nxt != lane) { vx += ox; vy += oy; cn += oc; nxt = (on == nxt) ? lane : on; }
    }
    if (act && (peers & ((1u << lane) - 1u)) == 0u) {
        atomicAdd((unsigned long long*)(p.fsum_x + label), (unsigned long long)vx);
        atomicAdd((unsigned long long*)(p.fsum_y + label), (unsigned long long)vy);
        atomicAdd(p.cnt + label, cn);
    }
}

// segmented suffix sum over the runs of a scan: lane i ends with the sum over lanes i .. i + rem
#define RUN_SEG6(rem, a0, a1, a2, a3, a4, a5)                                                                 \
    _Pragma("unroll") for (int d_ = 1; d_ < 32; d_ <<= 1) {                                                   \
        if (!__any_sync(FULLMASK, (rem) >= d_)) break;                                                        \
        const double b0 = __shfl_down_sync(FULLMASK, a0, d_), b1 = __shfl_down_sync(FULLMASK, a1, d_);        \
        const double b2 = __shfl_down_sync(FULLMASK, a2, d_), b3 = __shfl_down_sync(FULLMASK, a3, d_);        \
        const double b4 = __shfl_down_sync(FULLMASK, a4, d_), b5 = __shfl_down_sync(FULLMASK, a5, d_);        \
        if ((rem) >= d_) { a0 += b0; a1 += b1; a2 += b2; a3 += b3; a4 += b4; a5 += b5; }                      \
    }

// the scan's totals are complete in one lane: far observations see the mean of the scan's new label (PREV view), the scan
// is registered as label-creating, and the six moments go to the solve
__device__ __forceinline__ void finalize_scan(const RunParams& p, int t, int tile, int lt, bool halo, double px, double py, double m0, double m1,
                                              double m2, double m3, double m4, double m5, double nfar, double fsx, double fsy, double FBx, double FBy)
{
    if (nfar > 0.0) {
        const double yx = fsx / nfar - px, yy = fsy / nfar - py;
        m0 += nfar * yx; m3 += nfar * yy;
        m1 = fma(yx, FBx, m1); m2 = fma(yx, FBy, m2);
        m4 = fma(yy, FBx, m4); m5 = fma(yy, FBy, m5);
        if (!halo) {
            FarRec r;
            r.t = t; r.rank = 0; r.n = (int)nfar; r.pad = 0; r.sx = fsx; r.sy = fsy;
            p.far_list[atomicAdd(&p.ts->far_count, 1)] = r;
            atomicOr(p.farbits + (size_t)tile * 4 + (lt >> 5), 1u << (lt & 31));
        }
    }
    double2* d = reinterpret_cast<double2*>(p.dyn + (size_t)t * 6);
    d[0] = make_double2(m0, m1); d[1] = make_double2(m2, m3); d[2] = make_double2(m4, m5);
}

// One chunk by one warp.  STEADY: validate every run first; a chunk with a run that cannot be certified commits nothing,
// marks its scans dirty and returns false.  !STEADY (the association kernel has just written the records): no validation;
// only the scans with commit_all or a dirty flag are committed.
template <bool STEADY>
__device__ __forceinline__ bool process_chunk(const RunParams& p, const RunRec* chunk, int64_t chunk_gid, int t0, int tile, bool commit_all)
{
    const int lane = threadIdx.x & 31;
    const RunRec* rp = chunk + lane;
    const double2 sb = __ldcg(reinterpret_cast<const double2*>(rp));
    const int4 mw = __ldcg(reinterpret_cast<const int4*>(rp) + 1);
    const int label = mw.x;
    const float rho = __int_as_float(mw.y);
    const int n = mw.z & 0xffff, lt = (mw.z >> 16) & 0xff, rem = (mw.z >> 24) & 0xff;
    const int piece = mw.w & 0xff, flags = (mw.w >> 8) & 0xff, npieces = (mw.w >> 16) & 0xffff;
    const bool valid = label != RUN_PAD;
    const bool matched = valid && label >= 0;
    const int t = t0 + lt;
    double4 pp = make_double4(0.0, 0.0, 0.0, 1.0);
    if (valid) pp = ldg_ppar(p.ppar + t);
    double2 lm = make_double2(0.0, 0.0), lr = make_double2(0.0, 0.0);
    if (matched) {
        const double2* lp = reinterpret_cast<const double2*>(p.lmrec + label);
        lm = __ldg(lp); lr = __ldg(lp + 1);
    }
    const double px = pp.x, py = pp.y, st = pp.z, ct = pp.w;
    const double nd = (double)n;
    const double rwx = fma(ct, sb.x, -st * sb.y), rwy = fma(st, sb.x, ct * sb.y);   // sum of the rotated beams
    const double yx = lm.x - px, yy = lm.y - py;
    const double Sx = rwx - nd * yx, Sy = rwy - nd * yy;                          // sum of (observation - landmark)
    bool commit = true;
    if (STEADY) {
        const double a = (lr.y - (double)rho) * nd;
        const bool ok = !valid || (matched && !(flags & RF_LONG) && a > 0.0 && fma(Sx, Sx, Sy * Sy) <= a * a * (1.0 - 1e-9));
        if (!__all_sync(FULLMASK, ok)) {
            if (valid && (flags & RF_LEADER)) p.scan_dirty[t] = 1;
            if (lane == 0 && atomicCAS(p.tile_flag + tile, 0, 1) == 0) p.dirty_list[atomicAdd(&p.ts->n_dirty, 1)] = tile;
            return false;
        }
    } else {
        commit = valid && (commit_all || p.scan_dirty[t] != 0);
    }
    // ---- the six landmark moments of each scan ----------------------------------------------------------------------
    const bool mact = matched && commit;
    double m0 = mact ? nd * yx : 0.0, m1 = mact ? yx * sb.x : 0.0, m2 = mact ? yx * sb.y : 0.0;
    double m3 = mact ? nd * yy : 0.0, m4 = mact ? yy * sb.x : 0.0, m5 = mact ? yy * sb.y : 0.0;
    RUN_SEG6(rem, m0, m1, m2, m3, m4, m5)
    double nfar = 0.0, fsx = 0.0, fsy = 0.0, FBx = 0.0, FBy = 0.0, fpad = 0.0;
    if (!STEADY) {
        const bool fact = valid && label == RUN_FAR && commit;
        if (__any_sync(FULLMASK, fact)) {       // far runs: statistics of the scan's new label (ICM_SLAM.py:174-194)
            if (fact) { nfar = nd; fsx = fma(nd, px, rwx); fsy = fma(nd, py, rwy); FBx = sb.x; FBy = sb.y; }
            RUN_SEG6(rem, nfar, fsx, fsy, FBx, FBy, fpad)
        }
    }
    if (valid && (flags & RF_LEADER) && commit) {
        const bool halo = (flags & RF_HALO) != 0;
        if (!(flags & RF_LONG)) {
            finalize_scan(p, t, tile, lt, halo, px, py, m0, m1, m2, m3, m4, m5, nfar, fsx, fsy, FBx, FBy);
        } else {
            // a piece of a long scan: leave the partial sums, the last piece to arrive adds them up in piece order
            double* dx = p.dynx + (size_t)chunk_gid * 12;
            dx[0] = m0; dx[1] = m1; dx[2] = m2; dx[3] = m3; dx[4] = m4; dx[5] = m5;
            dx[6] = nfar; dx[7] = fsx; dx[8] = fsy; dx[9] = FBx; dx[10] = FBy;
            __threadfence();
            if (atomicAdd(p.ticket + t, 1) == npieces - 1) {
                __threadfence();
                p.ticket[t] = 0;
                double a[11];
                for (int k = 0; k < 11; ++k) a[k] = 0.0;
                for (int j = 0; j < npieces; ++j) {
                    const double* dj = p.dynx + (size_t)(chunk_gid - piece + j) * 12;
                    for (int k = 0; k < 11; ++k) a[k] += __ldcg(dj + k);
                }
                finalize_scan(p, t, tile, lt, halo, px, py, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]);
            }
        }
    }
    // ---- landmark statistics ------------------------------------------------------------------------------------------
    chunk_statistics(p, mact && !(flags & RF_HALO), label, Sx, Sy, n);
    return true;
}

#define RUNS_THREADS 256
#define RUNS_WARPS (RUNS_THREADS / 32)

// steady state: one block per record tile, a warp per chunk
__global__ void __launch_bounds__(RUNS_THREADS, 4)
k_runs(const RunParams p)
{
    const int tile = blockIdx.x;
    const int warp = threadIdx.x >> 5;
    const int nch = p.tile_nchunks[tile];
    if (p.tile_epoch[tile] != p.ts->epoch) {      // no records for this label numbering: the whole tile goes to the association kernel
        if (threadIdx.x == 0) { p.tile_flag[tile] = 2; p.dirty_list[atomicAdd(&p.ts->n_dirty, 1)] = tile; }
        return;
    }
    const int t0 = p.t_start + tile * RT_TILE;
    const int64_t base = run_tile_base(p.off, p.t_start, tile);
    for (int c = warp; c < nch; c += RUNS_WARPS)
        process_chunk<true>(p, p.rec + base + (int64_t)c * 32, (base >> 5) + c, t0, tile, true);
}

// every tile to the association kernel (first sweep of a handle that never built records, or ICMSLAM_RUNS=0)
__global__ void k_all_dirty(int* __restrict__ tile_flag, int* __restrict__ dirty_list, TailState* ts, int n_tiles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles) { tile_flag[i] = 2; dirty_list[i] = i; }
    if (i == 0) ts->n_dirty = n_tiles;
}
