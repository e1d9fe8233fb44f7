// fused.cuh -- the (REDBLACK, NEWTON, PREV) sweep: k_sweep_fused + k_solve_colour<0|1>.
//
// For every pose t of the trajectory (sensors.py:145-162 restated per SURVEY.md App. A): projection of the scan's kept
// beams with the sweep's INPUT pose (tras_rot_z, ICM_SLAM.py:465-480), nearest-landmark association against the previous
// map with the dist_thr gate (Mapa.actualizar Branch B, ICM_SLAM.py:169-182), per-landmark statistics for the landmark
// update (:191-194), and the pose's conditional minimiser of fun_xn / fun_x (sensors.py:224-282) in red-black order.
// Every observation is read from HBM exactly once.
//
// k_sweep_fused (block = 256 threads, 126 owned poses, one launch for the whole sweep):
//   * the block's contiguous slice of the observation records (bx, by) and last sweep's labels are staged in shared
//     memory with TMA bulk copies (cp.async.bulk + mbarrier); tiles whose observations exceed the budget are processed in
//     chunks of whole scans;
//   * phase A -- LANES OVER CONSECUTIVE OBSERVATIONS (four per lane in flight): validate last sweep's label inside the
//     landmark's proven-nearest radius, or project / associate through the grid (fastgrid.cuh); write the label.  Adjacent
//     lanes hold adjacent beams, i.e. mostly the same trunk, so the look-ups are coherent and every lane is busy;
//   * phase B -- THREAD PAIR PER POSE over the staged data.  Pass 1 collapses each run of beams that hit the same
//     landmark into an in-place record (run length, sum of bx, sum of by).  Pass 2 visits one RUN per step with the lanes
//     in lockstep: everything that depends on the landmark is linear in the run sums, so the landmark moments and the
//     landmark statistics cost one update per run, not per beam.  Statistics are added as int64 fixed point of
//     (observation - previous landmark) to a block-level shared-memory hash table (32-bit halves with carry: native shared
//     atomics) and flushed with one global integer atomic per (block, landmark): integer addition is associative, so the
//     landmark update is bit-reproducible for any block order or GPU count;
//   * split mode (default): the scan's six landmark moments and sin/cos of its input heading go to global memory and
//     k_solve_colour<0>, <1> (a lane pair per pose, no tile, no barrier) solve the odd poses from the OLD even neighbours
//     and then the even poses from the NEW odd ones.  ICMSLAM_SPLIT=0 keeps the solve inside this kernel (solve_phases:
//     odd warps, block barrier, even warps; the odd pose left of the tile is recomputed locally).
// The body-frame moments of a scan do not depend on poses or map: k_body_moments forms them once per dataset.
// Poses are double-buffered (xin -> xout).  Experimental, off by default: label certificates + run cache (see below).
#pragma once
#include "common.cuh"
#include "assoc.cuh"
#include "fastgrid.cuh"
#include "tail.cuh"

// Tile geometry, parametrised by HALF = pose slots per colour (64: 256 threads, 126 owned poses per block;
// 32: 128 threads, 62 owned poses -- more, smaller blocks per SM whose phases interleave better).
#define FS_THREADS (2 * TPP * HALF)  // TPP threads (a pair or a quad) per pose slot
#define FS_WARPS (FS_THREADS / 32)
#define FS_SLOTS (2 * HALF)    // slots 0..HALF-1: odd poses tb-1+2j (slot 0 = halo); HALF..: even poses tb+2j (last = spare)
#define FS_OWN (2 * HALF - 2)  // poses owned by a block: tb .. tb+FS_OWN-1 (tb even)
#define FS_XT (2 * HALF + 4)   // pose-tile entries: poses tb-2 .. tb+FS_OWN (2*HALF+1 used)
#define FS_HALF HALF
#define FS_HASH (HALF >= 64 ? 256 : HALF >= 32 ? 128 : 64)   // landmark slots of the block-level statistics table
#define FS_HASH_SHIFT (HALF >= 64 ? 24 : HALF >= 32 ? 25 : 26)
#define FS_PROBES 8
#ifndef FS_U
#define FS_U 4                 // observations per lane in flight in phase A (pend bitmask: 64 / FS_U iterations per warp)
#endif
#define FS_MINBLK (OCC / FS_THREADS)   // OCC = resident threads per SM the kernel is compiled for (register budget = 65536 / OCC)

struct FusedParams {
    int T;                                // columns of this handle's trajectory (a time segment incl. its halo columns)
    int t_lo, t_hi;                       // owned poses [t_lo, t_hi): whole trajectory = [0, T); a segment owns [2, T-1) etc.
    int first;                            // column 0 is the trajectory's first pose (pinned to self.x0, sensors.py:131)
    const int* off;                       // CSR offsets of the kept observations (T + 1)
    const double2* bxy;                   // body-frame observations (bx, by), CSR order
    const double* xin; int64_t ldin;      // 3 x T input poses
    double* xout; int64_t ldout;          // 3 x T output poses
    double x0[3];                         // self.x0 (sensors.py:131)
    const double* inc; int64_t ldinc;     // 3 x T odometry increments (see k_odo_increments)
    const double* u; int64_t ldu;         // 2 x T controls
    // Arrays that only the pose solve reads are stored COLOUR-MAJOR: pose t lives at cm(t) = (t & 1) * Th + (t >> 1), even poses
    // first, so that a launch solving one colour reads contiguous memory (in time order it would use half of every sector).
    int Th;                               // (T + 1) / 2
    const double* bm; int64_t ldbm;       // 6 x T static body-frame moments of each scan + its beam count (k_body_moments), colour-major
    const double* inc_cm; const double* u_cm;   // colour-major copies of inc (3 x T) and u (2 x T), leading dimension ldbm
    double* dyn; int64_t lddyn;           // 6 x T landmark moments of each scan (split mode: k_sweep_fused -> k_solve_colour), colour-major
    double* sc; double* scn; int64_t ldsc;   // 2 x T (sin, cos) of the input headings / of the new odd headings (split mode), colour-major
    // label certificates and the run cache (split mode; see "Label certificates" below)
    int cert;                             // certificates may be used this sweep (c[] and the stamps belong to this map chain)
    int stamp;                            // full passes stamp certificates and fill the run cache
    double rho, marg_scale;               // bound on |b| of a kept beam; 65535 / (2 sqrt(thr2_hi) dist_thr) rounded down
    double slack_unit;                    // one 8-bit slack step: dist_thr / 256, rounded down
    double2* rsum; int4* rmeta;           // run records: (sum bx, sum by) / (label, beams, landmark odometer when stamped, slack [float bits]);
                                          // a scan's first-half runs are stored from off[t] up, its second-half runs from off[t+1]-1 down
    int* rcnt;                            // runs of each scan: first half | second half << 16
    int* echk;                            // epoch in which the scan was stamped
    double* xchk; int64_t ldchk;          // 3 x T pose the scan was projected with when it was stamped
    DevCfg cfg;
    double thr2_hi;                       // largest s with sqrt_rn(s) <= dist_thr
    double fix_scale;                     // fixed-point scale of the statistics (power of two)
    double tol; int maxit;
    const DevState* st;
    const FGeom* geom;
    const int* cell_start; const double2* gpts; const int* gidx;
    const LmRec* lmrec;                   // landmarks by label: position + hint radius
    const int* remap;                     // last sweep's label -> this map's label (hints)
    int hints;                            // c[] holds last sweep's labels for this map chain
    int* c;                               // labels per observation (out)
    long long* fsum_x; long long* fsum_y; int* cnt;   // per previous-map landmark statistics
    FarRec* far_list; TailState* ts; int* blk_far;    // scans with far observations (each creates one new label)
    int obs_cap;                          // shared-memory capacity in observations
    int skip;                             // debug ablation mask (ICMSLAM_SKIP): 1 assoc, 2 pass 1, 4 pass 2, 8 statistics, 16 solve
    unsigned long long* iters;
    long long* prof;                      // debug (ICMSLAM_PROF): 16 clock stamps of warp 0 per block
};

// odometry increments, sweep-invariant: D_t = Rota(o_t.theta) (o_{t+1}.xy - o_t.xy), dtheta_t
// (sensors.py:236-238, :250-253).  inc[:, t] for t < T-1; the last column is zero.
__global__ void k_odo_increments(const double* __restrict__ odo, int64_t ldo, int T, double* __restrict__ inc, int64_t ldi)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double dx = 0.0, dy = 0.0, dth = 0.0;
    if (t + 1 < T) {
        double s, c;
        sincos(odo[2 * ldo + t], &s, &c);
        const double vx = odo[t + 1] - odo[t], vy = odo[ldo + t + 1] - odo[ldo + t];
        dx = c * vx + s * vy;
        dy = -s * vx + c * vy;
        dth = odo[2 * ldo + t + 1] - odo[2 * ldo + t];
    }
    inc[t] = dx; inc[ldi + t] = dy; inc[2 * ldi + t] = dth;
}

// body-frame moment sums of every scan's kept beams: sum bx, sum by, sum bx^2, sum by^2, sum bx*by.  They do not depend on
// the poses or the map, so they are formed once per dataset (one thread per scan, fixed order: the result does not depend
// on any tiling) and the sweep kernel only reads them (40 B per pose).
__device__ __forceinline__ int64_t cm_index(int t, int Th) { return (int64_t)(t & 1) * Th + (t >> 1); }

// rows x T array in time order -> colour-major (see FusedParams::Th)
__global__ void k_to_colour_major(const double* __restrict__ src, int64_t lds, int rows, int T, double* __restrict__ dst, int64_t ldd)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t c = cm_index(t, (T + 1) >> 1);
    for (int r = 0; r < rows; ++r) dst[r * ldd + c] = src[r * lds + t];
}

__global__ void __launch_bounds__(128)
k_body_moments(const int* __restrict__ off, const double2* __restrict__ bxy, int T, double* __restrict__ bm, int64_t ld)
{
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (t0 >= T) return;
    const int t = t0;
    double Bx = 0.0, By = 0.0, Bxx = 0.0, Byy = 0.0, Bxy = 0.0;
    for (int i = off[t]; i < off[t + 1]; ++i) {
        const double2 b = bxy[i];
        Bx += b.x; By += b.y;
        Bxx = fma(b.x, b.x, Bxx); Byy = fma(b.y, b.y, Byy); Bxy = fma(b.x, b.y, Bxy);
    }
    const int64_t c = cm_index(t, (T + 1) >> 1);      // colour-major, + the beam count as a sixth row
    bm[c] = Bx; bm[ld + c] = By; bm[2 * ld + c] = Bxx; bm[3 * ld + c] = Byy; bm[4 * ld + c] = Bxy;
    bm[5 * ld + c] = (double)(off[t + 1] - off[t]);
}

__device__ __forceinline__ double entrepi_fast(double a)
{
    return (fabs(a) <= ICM_PI) ? a : entrepi(a);
}

// sin/cos of a small increment (|d| <= 0.125): Taylor to d^11 / d^10, error < 1e-18
__device__ __forceinline__ void sincos_small(double d, double& s, double& c)
{
    const double z = d * d;
    s = d * fma(z, fma(z, fma(z, fma(z, fma(z, -2.5052108385441719e-08, 2.7557319223985893e-06), -1.9841269841269841e-04),
                              8.3333333333333332e-03), -1.6666666666666666e-01), 1.0);
    c = fma(z, fma(z, fma(z, fma(z, fma(z, -2.7557319223985888e-07, 2.4801587301587302e-05), -1.3888888888888889e-03),
                          4.1666666666666664e-02), -0.5), 1.0);
}

struct Mom {   // moment sums of one pose's observations (landmark coordinates relative to the origin)
    double n, Bx, By, Bxx, Byy, Bxy, Yx, Yy, Mxx, Mxy, Myx, Myy;
};

struct PoseIn {
    double ax, ay, ath, sa, ca;     // x_{t-1} and sin/cos of its heading
    double bx, by, bth;             // x_{t+1}
    double uav, uaw, ucv, ucw;      // u_{t-1}, u_t
    double D0x, D0y, dth0;          // odometry increment t-1 -> t
    double D1x, D1y, dth1;          // odometry increment t -> t+1
    int has_next;
};

// Exact conditional minimiser (same reduced 1-D Newton as newton_moments in pose.cuh, with the
// odometry rotations and the neighbour's sin/cos supplied, the start heading's sin/cos supplied,
// and the heading's sin/cos carried through the iterations by small-angle rotation).
__device__ __forceinline__ int newton_lean(const DevCfg& cfg, const PoseIn& P, const Mom& M, double ox, double oy, double th,
                                           double s, double c, double tol, int maxit, double out[3], double& s_out, double& c_out)
{
    const double dt = cfg.dt, k = cfg.kod;
    const double gax = (P.ax - ox) + dt * (P.ca * P.uav), gay = (P.ay - oy) + dt * (P.sa * P.uav), th_ga = P.ath + dt * P.uaw;
    const double e0x = (P.ax - ox) + (P.ca * P.D0x - P.sa * P.D0y);
    const double e0y = (P.ay - oy) + (P.sa * P.D0x + P.ca * P.D0y);
    const double hn = P.has_next ? 1.0 : 0.0;
    const double bxp = hn * (P.bx - ox), byp = hn * (P.by - oy);
    const double D1x = hn * P.D1x, D1y = hn * P.D1y;
    const double dv = hn * dt * P.ucv;
    const double Sx = cfg.r1 + k + M.n * cfg.q1 + hn * (cfg.r1 + k), Sy = cfg.r2 + k + M.n * cfg.q2 + hn * (cfg.r2 + k);
    const double iSx = 1.0 / Sx, iSy = 1.0 / Sy;
    const double KAx = cfg.r1 * gax + k * e0x + cfg.q1 * M.Yx + (cfg.r1 + k) * bxp;
    const double KAy = cfg.r2 * gay + k * e0y + cfg.q2 * M.Yy + (cfg.r2 + k) * byp;
    const double ang2 = (2.0 * cfg.r3 + 2.0 * k) * (1.0 + hn);
    const double Pxc = (cfg.r1 * dv + k * D1x) + cfg.q1 * M.By;
    const double Pxs = cfg.q1 * M.Bx - k * D1y;
    const double Pys = (cfg.r2 * dv + k * D1x) + cfg.q2 * M.By;
    const double Pyc = k * D1y - cfg.q2 * M.Bx;
    const double c3 = P.dth0 + P.ath;
    const double c4 = P.dth1 - P.bth;
    const double wb = dt * P.ucw - P.bth;
    const double dByx = M.Byy - M.Bxx;
    int it = 0;
    for (;;) {
        const double ss = s * s, cc = c * c, sc = s * c;
        const double Ax = KAx - c * Pxc - s * Pxs, Ax1 = s * Pxc - c * Pxs, Ax2 = KAx - Ax;
        const double Ay = KAy - s * Pys - c * Pyc, Ay1 = -c * Pys + s * Pyc, Ay2 = KAy - Ay;
        // motion / odometry terms towards t+1 (all zero when !has_next: dv = D1 = bxp = 0 and hn scales them)
        const double r1x = c * D1x - s * D1y, r1y = s * D1x + c * D1y;     // Rota(th)^T D1
        const double f2 = bxp - dv * c, f21 = dv * s;
        const double f4 = bxp - r1x;
        const double h2 = byp - dv * s, h21 = -dv * c;
        const double h4 = byp - r1y;
        double Cx1 = hn * (cfg.r1 * f2 * f21 + k * f4 * r1y);
        double Cx2 = hn * (cfg.r1 * (f21 * f21 + f2 * (dv * c)) + k * (r1y * r1y + f4 * r1x));
        double Cy1 = hn * (cfg.r2 * h2 * h21 - k * h4 * r1x);
        double Cy2 = hn * (cfg.r2 * (h21 * h21 + h2 * (dv * s)) + k * (r1x * r1x + h4 * r1y));
        // observation sums with w = Rot(th - pi/2) b = (bx s + by c, -bx c + by s)
        const double Swxwy = sc * dByx + (ss - cc) * M.Bxy;
        const double Swxx = ss * M.Bxx + 2.0 * sc * M.Bxy + cc * M.Byy;
        const double Swyy = cc * M.Bxx - 2.0 * sc * M.Bxy + ss * M.Byy;
        const double Yxwy = -c * M.Mxx + s * M.Mxy, Yxwx = s * M.Mxx + c * M.Mxy;
        const double Yywx = s * M.Myx + c * M.Myy, Yywy = -c * M.Myx + s * M.Myy;
        Cx1 += cfg.q1 * (Yxwy - Swxwy);
        Cx2 += cfg.q1 * (Swyy + Yxwx - Swxx);
        Cy1 += cfg.q2 * (-Yywx + Swxwy);
        Cy2 += cfg.q2 * (Swxx + Yywy - Swyy);
        double ang1 = 2.0 * cfg.r3 * entrepi_fast(th - th_ga) - 2.0 * k * entrepi_fast(c3 - th);
        if (P.has_next) ang1 += 2.0 * cfg.r3 * entrepi_fast(th + wb) + 2.0 * k * entrepi_fast(c4 + th);
        const double p1 = 2.0 * (Cx1 - Ax * Ax1 * iSx + Cy1 - Ay * Ay1 * iSy) + ang1;
        double p2 = 2.0 * (Cx2 - (Ax1 * Ax1 + Ax * Ax2) * iSx + Cy2 - (Ay1 * Ay1 + Ay * Ay2) * iSy) + ang2;
        if (!(p2 > 0.0)) p2 = ang2;
        const double dth = -p1 / p2;
        th += dth;
        ++it;
        if (fabs(dth) <= 0.125) {
            double sd, cd;
            sincos_small(dth, sd, cd);
            const double s2 = s * cd + c * sd;
            c = c * cd - s * sd;
            s = s2;
        } else {
            sincos(th, &s, &c);
        }
        const bool done = fabs(dth) <= tol || it >= maxit;
        if (__all_sync(__activemask(), done)) break;
    }
    out[0] = (KAx - c * Pxc - s * Pxs) * iSx + ox;
    out[1] = (KAy - s * Pys - c * Pyc) * iSy + oy;
    out[2] = th;
    s_out = s; c_out = c;
    return it;
}

// The same solve split over a lane pair: lane role 0 carries the x equations, role 1 the y equations.
// With (a, b) = (cos, sin) of theta for role 0 and of (theta - pi/2) for role 1, both roles evaluate
// IDENTICAL formulas on their own constants (derived from newton_lean: the y rows are the x rows
// rotated by -pi/2), so the pair halves the dependent chain of every iteration and the two partial
// derivatives meet through one shuffle each.  Returns the role's coordinate in `coord`, theta in `th`.
__device__ __forceinline__ int newton_pair(const DevCfg& cfg, const PoseIn& P, const Mom& M, int role, double ox, double oy, double& th,
                                           double s, double c, double tol, int maxit, double& coord, double& s_out, double& c_out)
{
    const double dt = cfg.dt, k = cfg.kod;
    const double hn = P.has_next ? 1.0 : 0.0;
    const double D1x = hn * P.D1x, D1y = hn * P.D1y;
    const double dv = hn * dt * P.ucv;
    // role constants
    const double r = role ? cfg.r2 : cfg.r1, q = role ? cfg.q2 : cfg.q1;
    const double o_ = role ? oy : ox;
    const double a_ = (role ? P.ay : P.ax) - o_;
    const double ga = a_ + dt * ((role ? P.sa : P.ca) * P.uav);
    const double e0 = a_ + (role ? (P.sa * P.D0x + P.ca * P.D0y) : (P.ca * P.D0x - P.sa * P.D0y));
    const double bp = hn * ((role ? P.by : P.bx) - o_);
    const double iS = 1.0 / (r + k + M.n * q + hn * (r + k));
    const double KA = r * ga + k * e0 + q * (role ? M.Yy : M.Yx) + (r + k) * bp;
    const double P1 = (r * dv + k * D1x) + q * M.By;
    const double P2 = q * M.Bx - k * D1y;
    const double M1 = role ? M.Myx : M.Mxx, M2 = role ? M.Myy : M.Mxy;
    const double dB = M.Bxx - M.Byy;
    const double ang2 = (2.0 * cfg.r3 + 2.0 * k) * (1.0 + hn);
    // angular residual constants: role 0 owns the terms towards t-1, role 1 those towards t+1
    const double th_ga = P.ath + dt * P.uaw, c3 = P.dth0 + P.ath, c4 = P.dth1 - P.bth, wb = dt * P.ucw - P.bth;
    // (a, b): role 0 = (cos th, sin th); role 1 = (sin th, -cos th)
    double a = role ? s : c, b = role ? -c : s;
    int it = 0;
    bool done = false;
    for (;;) {
        const double A = KA - a * P1 - b * P2, A1 = b * P1 - a * P2, A2 = KA - A;
        const double ra = a * D1x - b * D1y, rb = b * D1x + a * D1y;
        const double f2 = bp - dv * a, f21 = dv * b, f4 = bp - ra;
        double C1 = r * f2 * f21 + k * f4 * rb;
        double C2 = r * (f21 * f21 + f2 * (dv * a)) + k * (rb * rb + f4 * ra);
        const double ab = a * b, aa_bb = a * a - b * b;
        const double Dq = -ab * dB - aa_bb * M.Bxy;                 // sum_i w_x w_y in the role's frame
        const double Eq = aa_bb * dB - 4.0 * ab * M.Bxy;            // sum_i (w_y^2 - w_x^2)
        C1 += q * ((b * M2 - a * M1) - Dq);
        C2 += q * (Eq + (b * M1 + a * M2));
        double g1 = 2.0 * (C1 - A * A1 * iS);
        double g2 = 2.0 * (C2 - (A1 * A1 + A * A2) * iS);
        if (role == 0) g1 += 2.0 * cfg.r3 * entrepi_fast(th - th_ga) - 2.0 * k * entrepi_fast(c3 - th);
        else if (P.has_next) g1 += 2.0 * cfg.r3 * entrepi_fast(th + wb) + 2.0 * k * entrepi_fast(c4 + th);
        const double p1 = g1 + __shfl_xor_sync(FULLMASK, g1, 1);
        double p2 = g2 + __shfl_xor_sync(FULLMASK, g2, 1) + ang2;
        if (!(p2 > 0.0)) p2 = ang2;
        const double dth = -p1 * (double)__frcp_rn((float)p2);   // quasi-Newton: 24-bit reciprocal of the curvature, same fixed point
        if (!done) {       // a converged pair is frozen: its result does not depend on how long its warp-mates iterate
            th += dth;
            ++it;
            if (fabs(dth) <= 0.125) {
                double sd, cd;
                sincos_small(dth, sd, cd);
                const double b2 = b * cd + a * sd;
                a = a * cd - b * sd;
                b = b2;
            } else {
                sincos(role ? th - ICM_HALFPI : th, &b, &a);
            }
            done = fabs(dth) <= tol || it >= maxit;
        }
        if (__all_sync(FULLMASK, done)) break;
    }
    coord = (KA - a * P1 - b * P2) * iS + o_;
    s_out = role ? a : b;          // sin th
    c_out = role ? -b : a;         // cos th
    return it;
}

// The reduced 1-D problem in closed form.  Expanding newton_lean's phi'(theta) with s = sin(theta), c = cos(theta) gives a
// trigonometric polynomial plus the (piecewise linear) angular residuals,
//     phi'(theta)  = 2 (as s + ac c + au s c + av (c^2 - s^2)) + ang1(theta)
//     phi''(theta) = 2 (as c - ac s + au (c^2 - s^2) - 4 av s c) + ang2,
// whose four coefficients depend on the pose's moments and neighbours but not on theta: they are formed ONCE (lane role 0
// contributes the x rows, role 1 the y rows, which are the x rows rotated by -pi/2; one shuffle each), and an iteration
// costs a dozen FMAs instead of re-evaluating every term.  Same root as newton_lean / the oracle's newton_pose.
// Both lanes of the pair iterate on identical values.  Returns the role's coordinate in `coord`, theta in `th`.
__device__ __forceinline__ int newton_trig(const DevCfg& cfg, const PoseIn& P, const Mom& M, int role, double ox, double oy, double& th,
                                           double s, double c, double tol, int maxit, double& coord, double& s_out, double& c_out)
{
    const double dt = cfg.dt, k = cfg.kod;
    const double hn = P.has_next ? 1.0 : 0.0;
    const double D1x = hn * P.D1x, D1y = hn * P.D1y;
    const double dv = hn * dt * P.ucv;
    // role constants (x rows for role 0, y rows for role 1), frame (a, b) = (c, s) resp. (s, -c)
    const double r = role ? cfg.r2 : cfg.r1, q = role ? cfg.q2 : cfg.q1;
    const double o_ = role ? oy : ox;
    const double a_ = (role ? P.ay : P.ax) - o_;
    const double ga = a_ + dt * ((role ? P.sa : P.ca) * P.uav);
    const double e0 = a_ + (role ? (P.sa * P.D0x + P.ca * P.D0y) : (P.ca * P.D0x - P.sa * P.D0y));
    const double bp = hn * ((role ? P.by : P.bx) - o_);
    const double iS = 1.0 / (r + k + M.n * q + hn * (r + k));
    const double KA = r * ga + k * e0 + q * (role ? M.Yy : M.Yx) + (r + k) * bp;
    const double P1 = (r * dv + k * D1x) + q * M.By;
    const double P2 = q * M.Bx - k * D1y;
    const double M1 = role ? M.Myx : M.Mxx, M2 = role ? M.Myy : M.Mxy;
    const double iKA = iS * KA;
    const double ba = fma(iKA, P2, fma(k * bp, D1y, -q * M1));
    const double bb = fma(-iKA, P1, fma(bp, fma(r, dv, k * D1x), q * M2));
    const double bab = fma(-iS, fma(P2, P2, -P1 * P1), fma(q, M.Bxx - M.Byy, -fma(r * dv, dv, k * fma(D1x, D1x, -D1y * D1y))));
    const double bd = fma(-iS * P1, P2, fma(q, M.Bxy, -k * D1x * D1y));
    double as2 = role ? ba : bb, ac2 = role ? -bb : ba, au2 = role ? -bab : bab, av2 = role ? -bd : bd;
    as2 += __shfl_xor_sync(FULLMASK, as2, 1); ac2 += __shfl_xor_sync(FULLMASK, ac2, 1);
    au2 += __shfl_xor_sync(FULLMASK, au2, 1); av2 += __shfl_xor_sync(FULLMASK, av2, 1);
    as2 *= 2.0; ac2 *= 2.0; au2 *= 2.0; av2 *= 2.0;
    const double av8 = 4.0 * av2;
    const double ang2 = (2.0 * cfg.r3 + 2.0 * k) * (1.0 + hn);
    const double r3_2 = 2.0 * cfg.r3, k_2 = 2.0 * k;
    const double th_ga = P.ath + dt * P.uaw, c3 = P.dth0 + P.ath, c4 = P.dth1 - P.bth, wb = dt * P.ucw - P.bth;
    int it = 0;
    bool done = false;
    for (;;) {
        const double u = s * c, v = fma(c, c, -s * s);
        double ang1 = fma(r3_2, entrepi_fast(th - th_ga), -k_2 * entrepi_fast(c3 - th));
        if (P.has_next) ang1 += fma(r3_2, entrepi_fast(th + wb), k_2 * entrepi_fast(c4 + th));
        const double p1 = fma(as2, s, fma(ac2, c, fma(au2, u, fma(av2, v, ang1))));
        double p2 = fma(as2, c, fma(-ac2, s, fma(au2, v, fma(-av8, u, ang2))));
        if (!(p2 > 0.0)) p2 = ang2;
        const double dth = -p1 * (double)__frcp_rn((float)p2);   // quasi-Newton: 24-bit reciprocal of the curvature, same fixed point
        if (!done) {       // a converged pair is frozen: its result does not depend on how long its warp-mates iterate
            th += dth;
            ++it;
            if (fabs(dth) <= 0.125) {
                double sd, cd;
                sincos_small(dth, sd, cd);
                const double s2 = fma(s, cd, c * sd);
                c = fma(c, cd, -s * sd);
                s = s2;
            } else {
                sincos(th, &s, &c);
            }
            done = fabs(dth) <= tol || it >= maxit;
        }
        if (__all_sync(FULLMASK, done)) break;
    }
    const double a = role ? s : c, b = role ? -c : s;
    coord = (KA - a * P1 - b * P2) * iS + o_;
    s_out = s; c_out = c;
    return it;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int HALF, bool SPLIT = false>
struct __align__(16) FusedSmemFixed {
    double xs[3][FS_XT];           // input poses tb-2 .. tb+126, index lt = t - (tb-2)
    double2 pp[FS_XT];             // projection origin of each scan (self.x0 for scan 0)
    double2 rsc[FS_XT];            // (sin, cos) of (projection heading - pi/2)
    // (split mode: the solve runs in k_pose_solve, so its inputs are not staged here)
    double sn[SPLIT ? 1 : FS_XT], cs[SPLIT ? 1 : FS_XT];   // sin/cos of the input headings (after phase B1: of the new odd headings)
    double xn[3][SPLIT ? 1 : FS_XT];           // new poses (same indexing)
    double inc[3][SPLIT ? 1 : FS_XT];          // odometry increments
    double u[2][SPLIT ? 1 : FS_XT];
    int off[FS_XT];                // off[t] at lt (clamped)
    int hkey[FS_HASH];
    int hcnt[FS_HASH];
    unsigned hs[FS_HASH][4];       // fixed-point sums: x (lo, hi), y (lo, hi)
    unsigned farbits[4];           // owned scans (bit t - tb) that have far observations (FS_OWN <= 128)
    unsigned long long mbar;
};

// adds one run (label, sum of (obs - landmark), count) to the block-level table.  Shared memory has no
// native 64-bit atomic add, so a 64-bit fixed-point sum is kept as (lo, hi) 32-bit halves: the low add
// returns the old value, from which the carry into the high half follows exactly.
// (Tried and dropped: grouping the warp's lanes by label with __match_any + __reduce_add before touching the table --
//  the partial-mask reductions cost ten times what the conflicting atomics do.)
template <int HALF, bool SPLIT>
__device__ __forceinline__ void stat_add(FusedSmemFixed<HALF, SPLIT>& S, const FusedParams& p, bool active, int arg, double rdx, double rdy, int rn)
{
    if (!active) return;
    const long long vx = __double2ll_rn(rdx * p.fix_scale), vy = __double2ll_rn(rdy * p.fix_scale);
    unsigned h = ((unsigned)arg * 2654435761u) >> FS_HASH_SHIFT;
#pragma unroll 1
    for (int q = 0; q < FS_PROBES; ++q) {
        int key = ((volatile int*)S.hkey)[h];
        if (key == -1) {
            const int prev = atomicCAS(&S.hkey[h], -1, arg);
            key = (prev == -1) ? arg : prev;
        }
        if (key == arg) {
            const unsigned xl = (unsigned)vx, xh = (unsigned)((unsigned long long)vx >> 32);
            const unsigned yl = (unsigned)vy, yh = (unsigned)((unsigned long long)vy >> 32);
            const unsigned ox = atomicAdd(&S.hs[h][0], xl);
            const unsigned oy = atomicAdd(&S.hs[h][2], yl);
            atomicAdd(&S.hcnt[h], rn);
            atomicAdd(&S.hs[h][1], xh + ((unsigned)(ox + xl) < ox ? 1u : 0u));      // carry out of the low half
            atomicAdd(&S.hs[h][3], yh + ((unsigned)(oy + yl) < oy ? 1u : 0u));
            return;
        }
        h = (h + 1) & (FS_HASH - 1);
    }
    atomicAdd((unsigned long long*)(p.fsum_x + arg), (unsigned long long)vx);   // table crowded: go to global directly
    atomicAdd((unsigned long long*)(p.fsum_y + arg), (unsigned long long)vy);
    atomicAdd(p.cnt + arg, rn);
}

__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
}

// The red-black pose solve of one tile: the slots of the odd poses (warps of the first half) solve from the OLD even
// neighbours, then, after a block barrier, the slots of the even poses solve from the NEW odd poses held in shared memory.
// Shared by the single-launch kernel and by k_pose_solve (split mode).  S needs xs, xn, inc, u, sn, cs (FS_XT entries each).
template <int HALF, int TPP, class SM>
__device__ __forceinline__ void solve_phases(SM& S, const FusedParams& p, const Mom& M, int q, int sub, int t, int li, bool qvalid)
{
    const int half = sub & 1;
    const int grp = q / FS_HALF;
    const int T = p.T;
    unsigned long long my_iters = 0;
    for (int phase = 0; phase < 2; ++phase) {
        if (phase == grp && !(p.skip & 64)) {       // warp-uniform: a warp holds 16 slots of one colour
            double res = 0.0, th = 0.0, s_new = 0.0, c_new = 1.0;
            if (qvalid) {
                // neighbours: old poses for the odd phase, new (odd) poses for the even phase
                double (*X)[FS_XT] = phase == 0 ? S.xs : S.xn;
                const bool has_next = t + 1 < T;
                const bool solve = !(t == 0 && p.first) && M.n > 0.0;
                PoseIn P;
                P.ax = X[0][li - 1]; P.ay = X[1][li - 1]; P.ath = X[2][li - 1];
                P.sa = S.sn[li - 1]; P.ca = S.cs[li - 1];
                P.bx = has_next ? X[0][li + 1] : 0.0; P.by = has_next ? X[1][li + 1] : 0.0; P.bth = has_next ? X[2][li + 1] : 0.0;
                P.uav = S.u[0][li - 1]; P.uaw = S.u[1][li - 1];
                P.ucv = S.u[0][li]; P.ucw = S.u[1][li];
                P.D0x = S.inc[0][li - 1]; P.D0y = S.inc[1][li - 1]; P.dth0 = S.inc[2][li - 1];
                P.D1x = S.inc[0][li]; P.D1y = S.inc[1][li]; P.dth1 = S.inc[2][li];
                P.has_next = has_next ? 1 : 0;
                th = S.xs[2][li];
                if (t == 0 && p.first) {
                    res = S.xs[half][li];
                    s_new = S.sn[li]; c_new = S.cs[li];
                } else if (!solve) {       // sensors.py:147-151: no observation, average of the neighbours
                    const bool t1 = t == 1 && p.first;
                    const double pv = t1 ? p.x0[half] : X[half][li - 1];
                    res = (pv + X[half][li + 1]) / 2.0;
                    th = ((t1 ? p.x0[2] : X[2][li - 1]) + X[2][li + 1]) / 2.0;
                    sincos(th, &s_new, &c_new);
                }
                // (lanes that do not solve still run the loop below with harmless values: the pair shuffles inside
                //  newton_trig need every lane of the warp)
                double r2 = 0.0, th2 = S.xs[2][li], s2 = 0.0, c2 = 1.0;
                const int it = newton_trig(p.cfg, P, M, half, S.xs[0][li], S.xs[1][li], th2, S.sn[li], S.cs[li], p.tol,
                                           (solve && !(p.skip & 16)) ? p.maxit : 1, r2, s2, c2);
                if (solve) { res = r2; th = th2; s_new = s2; c_new = c2; my_iters += (unsigned long long)(sub == 0 ? it : 0); }
            } else {
                PoseIn P;
                P.ax = P.ay = P.ath = P.sa = 0.0; P.ca = 1.0; P.bx = P.by = P.bth = 0.0; P.uav = P.uaw = P.ucv = P.ucw = 0.0;
                P.D0x = P.D0y = P.dth0 = P.D1x = P.D1y = P.dth1 = 0.0; P.has_next = 0;
                double r2, th2 = 0.0, s2, c2;
                newton_trig(p.cfg, P, M, half, 0.0, 0.0, th2, 0.0, 1.0, p.tol, 1, r2, s2, c2);
            }
            if (qvalid && sub < 2) {     // (with a quad per slot, lanes 2 and 3 repeat lanes 0 and 1)
                S.xn[half][li] = res;
                if (half == 0) {
                    S.xn[2][li] = th;
                    if (phase == 0) { S.sn[li] = s_new; S.cs[li] = c_new; }   // (old odd headings are no longer needed)
                }
            }
        }
        __syncthreads();
    }
    if (p.iters) {
        my_iters = (unsigned long long)warp_sum_i((int)my_iters);
        if ((threadIdx.x & 31) == 0 && my_iters) atomicAdd(p.iters, my_iters);
    }
}

// Label certificates (split mode).  When phase A accepts last sweep's label l for an observation at distance d from the
// landmark, inside its proven-nearest radius r (d^2 <= r^2), the label stays the strict argmin inside the gate for as long as
// the observation's world position and the landmark record together move by less than r - d >= (r^2 - d^2) / (2 sqrt(thr2_hi)).
// A run of beams keeps the least such slack of its observations (0 if one of them went through the grid search or is far)
// and the landmark's odometer g (LmRec::g, tail.cuh: how far the record has moved in total) at that time; the scan keeps the
// pose it was projected with.  On a later sweep the labels of a run -- hence its cached record (label, beams, sum of the
// body-frame points) -- are PROVABLY what phase A would produce if
//     (|dx| + |dy| + |dtheta| rho + (g_now - g_then)) (1 + 1e-6) + 1e-9 < slack,
// rho bounding the beam length.  A tile whose runs all pass skips staging, association and the run pass: it forms the
// moments from ~5 cached run records per scan instead of ~25 observations.  The attempt is speculative (it only touches the
// block's shared-memory table); a tile that fails it runs the full path, which re-stamps its scans.  Anything that
// renumbers landmarks bumps the epoch and voids every certificate.
#define FS_STAMP(k) if (p.prof && threadIdx.x == 0) p.prof[(size_t)blockIdx.x * 24 + (k)] = clock64()
template <int HALF, int TPP, int OCC, bool SPLIT>
__global__ void __launch_bounds__(FS_THREADS, FS_MINBLK)
k_sweep_fused(const FusedParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmemFixed<HALF, SPLIT>& S = *reinterpret_cast<FusedSmemFixed<HALF, SPLIT>*>(smem_raw);
    double2* sb = reinterpret_cast<double2*>(smem_raw + sizeof(FusedSmemFixed<HALF, SPLIT>));   // staged observations; later run sums
    int* sbk_raw = reinterpret_cast<int*>(sb + p.obs_cap);                          // hints in (TMA), labels out; +4 ints of alignment slack
    unsigned short* srn = reinterpret_cast<unsigned short*>(sbk_raw + p.obs_cap + 4);   // run length at run heads (phase A: cell entry count)
    unsigned char* slt = reinterpret_cast<unsigned char*>(srn + p.obs_cap);             // scan (local pose index) of each observation

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    FS_STAMP(0);
    const int tb = p.t_lo + blockIdx.x * FS_OWN;      // t_lo is even: colours stay aligned with the global time index
    const int T = p.T;
    // ---- tile loads ---------------------------------------------------------------------------
    for (int li = tid; li < FS_XT; li += FS_THREADS) {
        const int t = tb - 2 + li;
        const bool ok = t >= 0 && t < T;
        for (int r = 0; r < 3; ++r) {
            S.xs[r][li] = ok ? p.xin[r * p.ldin + t] : 0.0;
            if (!SPLIT) S.inc[r][li] = ok ? p.inc[r * p.ldinc + t] : 0.0;
        }
        if (!SPLIT) {
            S.u[0][li] = ok ? p.u[t] : 0.0;
            S.u[1][li] = ok ? p.u[p.ldu + t] : 0.0;
        }
        S.off[li] = p.off[min(max(t, 0), T)];
    }
    for (int h = tid; h < FS_HASH; h += FS_THREADS) { S.hkey[h] = -1; S.hcnt[h] = 0; for (int k = 0; k < 4; ++k) S.hs[h][k] = 0u; }
    const uint32_t mb = smem_u32(&S.mbar);
    if (tid < 4) S.farbits[tid] = 0u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    FS_STAMP(1);
    // ---- per-pose projection parameters: thread tid <-> pose tb-2+tid --------------------------------
    // (split mode without certificates: deferred until the observations' bulk copy is in flight, see below)
    auto proj_params = [&]() {
        if (tid < FS_SLOTS) {
            const int t = tb - 2 + tid;
            if (t >= 0 && t < T) {
                double px = S.xs[0][tid], py = S.xs[1][tid], th = S.xs[2][tid];
                double st, ct, sh, ch;
                if (t == 0 && p.first) {                             // scan 0 is projected with self.x0; x[:,0] is only a neighbour
                    sincos(th, &sh, &ch);
                    px = p.x0[0]; py = p.x0[1]; th = p.x0[2];
                    sincos(sub_rn(th, ICM_HALFPI), &st, &ct);
                } else {
                    sincos(sub_rn(th, ICM_HALFPI), &st, &ct);        // make_rot: cos/sin of (theta - pi/2)
                    sh = ct; ch = -st;
                }
                S.pp[tid] = make_double2(px, py); S.rsc[tid] = make_double2(st, ct);
                if (!SPLIT) { S.sn[tid] = sh; S.cs[tid] = ch; }
                else if (tid >= 2 ? t < p.t_hi : blockIdx.x == 0) { const int64_t c = cm_index(t, p.Th); p.sc[c] = sh; p.sc[p.ldsc + c] = ch; }
            }
        }
    };
    const bool proj_early = !SPLIT || p.cert != 0;
    if (proj_early) proj_params();
    // ---- pose slots: slot q is worked by the thread pair (2q, 2q+1) in phase B and solved by thread q -----
    //      slots 0..63: odd poses tb-1+2j (slot 0 = halo), slots 64..127: even poses tb+2j (slot 127 spare)
    const int q = tid / TPP, sub = tid % TPP, half = sub & 1;   // half = the lane's role in the solve (x / y rows)
    const int qt = q < FS_HALF ? tb - 1 + 2 * q : tb + 2 * (q - FS_HALF);
    const int qli = qt - (tb - 2);
    // (split mode: the moments of pose tb-1 come from the tile that owns it; only a segment's first tile computes them for the
    //  halo column t_lo-1, which no tile of this handle owns)
    const bool qvalid = q != FS_SLOTS - 1 && qt >= 0 && qt < p.t_hi && (q == 0 ? (!SPLIT || blockIdx.x == 0) : qt >= p.t_lo);
    const bool qowned = qvalid && q != 0;                       // the halo pose tb-1 is recomputed, not owned
    Mom M;
    M.n = M.Bx = M.By = M.Bxx = M.Byy = M.Bxy = M.Yx = M.Yy = M.Mxx = M.Mxy = M.Myx = M.Myy = 0.0;
    if (qvalid && !SPLIT) {     // static body-frame moments of the slot's scan (in flight until the solve)
        M.n = (double)(S.off[qli + 1] - S.off[qli]);
        const double* bmq = p.bm + cm_index(qt, p.Th);
        M.Bx = __ldg(bmq); M.By = __ldg(bmq + p.ldbm); M.Bxx = __ldg(bmq + 2 * p.ldbm);
        M.Byy = __ldg(bmq + 3 * p.ldbm); M.Bxy = __ldg(bmq + 4 * p.ldbm);
    }
    int nfar = 0;
    double fsx = 0.0, fsy = 0.0, FBx = 0.0, FBy = 0.0;

    FGrid G;
    G.g = *p.geom;
    G.cell_start = p.cell_start; G.pts = p.gpts; G.idx = p.gidx;
    const bool have_map = p.st->lsearch > 0;
    const int t_first = (SPLIT && blockIdx.x != 0) ? tb : max(tb - 1, 0), t_last = min(tb + FS_OWN - 1, p.t_hi - 1);     // scans processed by this block
    const int lt_first = t_first - (tb - 2), lt_last = t_last - (tb - 2);
    // ---- label certificates: try to form the tile's moments from the cached run records ------------------------------
    constexpr bool CERT = SPLIT && TPP == 2;
    bool fast = false;
    int epoch = 0;
    double pth_ = 0.0;                      // heading the slot's scan is projected with
    if constexpr (CERT) {
        epoch = p.ts->epoch;
        bool ok_all = true;
        double dpose = 0.0;
        if (qvalid) {
            const bool pin = qt == 0 && p.first;
            const double ppx = pin ? p.x0[0] : S.xs[0][qli], ppy = pin ? p.x0[1] : S.xs[1][qli];
            pth_ = pin ? p.x0[2] : S.xs[2][qli];
            if (p.cert) {
                dpose = fabs(ppx - p.xchk[qt]) + fabs(ppy - p.xchk[p.ldchk + qt]) + fabs(pth_ - p.xchk[2 * p.ldchk + qt]) * p.rho;
                ok_all = p.echk[qt] == epoch && dpose < 1e30;
            }
        }
        if (p.cert) __syncthreads();        // (the projection parameters are in place)
        FS_STAMP(18);
        // (without certificates no barrier is needed here: nothing was written to shared memory since the last one)
        if (p.cert && __syncthreads_and(ok_all)) {
            FS_STAMP(19);
            // The slot's cached run records, then the landmark records they name, are brought into shared memory with
            // 16-byte asynchronous copies (every copy of a stage is in flight at once: one memory round trip per stage,
            // whatever the number of runs); the moments are then formed from shared memory.
            struct __align__(16) RunSlot { double2 sum; int4 meta; double2 lm; double2 lg; };
            RunSlot* rs = reinterpret_cast<RunSlot*>(sb);
            const int cap_runs = (int)(((size_t)p.obs_cap * 23) / sizeof(RunSlot));
            int nr = 0, gi = 0;
            const int step = sub == 0 ? 1 : -1;
            if (qvalid) {
                const int rcw = __ldg(p.rcnt + qt);
                nr = sub == 0 ? (rcw & 0xffff) : (rcw >> 16);
                gi = sub == 0 ? S.off[qli] : S.off[qli + 1] - 1;
            }
            // block-wide exclusive scan of the run counts: compact slots
            int inc = nr;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULLMASK, inc, d); if (lane >= d) inc += v; }
            int* wsum = reinterpret_cast<int*>(S.hcnt);          // (the statistics table is empty at this point; restored below)
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            int base = inc - nr, total = 0;
            for (int w = 0; w < FS_WARPS; ++w) { const int v = wsum[w]; if (w < warp) base += v; total += v; }
            __syncthreads();
            if (tid < FS_WARPS) wsum[tid] = 0;
            __syncthreads();
            FS_STAMP(20);
            if (total <= cap_runs) {
                for (int k = 0; k < nr; ++k) {
                    const uint32_t d0 = smem_u32(&rs[base + k]);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(p.rsum + gi + k * step) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u), "l"(p.rmeta + gi + k * step) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                FS_STAMP(21);
                for (int k = 0; k < nr; ++k) {
                    const int lab = rs[base + k].meta.x;
                    if (lab >= 0) {
                        const uint32_t d0 = smem_u32(&rs[base + k].lm);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(p.lmrec + lab) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u), "l"(reinterpret_cast<const double2*>(p.lmrec + lab) + 1) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                FS_STAMP(22);
                const double2 pq = qvalid ? S.pp[qli] : make_double2(0.0, 0.0), rc = qvalid ? S.rsc[qli] : make_double2(0.0, 1.0);
                const double px = pq.x, py = pq.y, st = rc.x, ct = rc.y;
                int k = 0;
                while (__any_sync(FULLMASK, k < nr)) {
                    const bool act = k < nr;
                    double2 sxy = make_double2(0.0, 0.0), lm = make_double2(0.0, 0.0), lg = make_double2(0.0, 0.0);
                    int4 meta = make_int4(-1, 0, 0, 0);
                    if (act) {
                        const RunSlot& r = rs[base + k];
                        sxy = r.sum; meta = r.meta;
                        if (meta.x >= 0) { lm = r.lm; lg = r.lg; }
                        // the run's certificate
                        const double drift = (dpose + (lg.y - (double)__int_as_float(meta.z))) * (1.0 + 1e-6) + 1e-9;
                        ok_all = ok_all && meta.x >= 0 && drift < (double)__int_as_float(meta.w);
                    }
                    const bool matched = act && meta.x >= 0;
                    const double dn = (double)meta.y;
                    const double rwx = fma(ct, sxy.x, -st * sxy.y), rwy = fma(st, sxy.x, ct * sxy.y);
                    const double yx = lm.x - px, yy = lm.y - py;
                    if (matched) {
                        M.Yx = fma(dn, yx, M.Yx); M.Yy = fma(dn, yy, M.Yy);
                        M.Mxx = fma(yx, sxy.x, M.Mxx); M.Mxy = fma(yx, sxy.y, M.Mxy);
                        M.Myx = fma(yy, sxy.x, M.Myx); M.Myy = fma(yy, sxy.y, M.Myy);
                    }
                    stat_add(S, p, matched && qowned && !(p.skip & 8), meta.x, rwx - dn * yx, rwy - dn * yy, meta.y);
                    if (act) ++k;
                }
            } else ok_all = false;
            FS_STAMP(23);
            fast = __syncthreads_and(ok_all) != 0;
            if (!fast) {       // some run lost its certificate: undo the attempt, the full path follows
                M.Yx = M.Yy = M.Mxx = M.Mxy = M.Myx = M.Myy = 0.0;
                for (int h = tid; h < FS_HASH; h += FS_THREADS) { S.hkey[h] = -1; S.hcnt[h] = 0; for (int kk = 0; kk < 4; ++kk) S.hs[h][kk] = 0u; }
                __syncthreads();
            } else if (tid == 0) atomicAdd(&const_cast<DevState*>(p.st)->cert_tiles, 1);
        }
    } else {
        __syncthreads();
    }
    FS_STAMP(2);
    // ---- chunks of whole scans whose observations fit the shared-memory budget (normally one) -----------
    uint32_t parity = 0;
    int myruns = 0;
    for (int c_lo = fast ? lt_last + 1 : lt_first; c_lo <= lt_last;) {
        int c_hi = lt_last;
        if (S.off[lt_last + 1] - S.off[c_lo] > p.obs_cap) {
            int lo = c_lo, hi = lt_last;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (S.off[mid + 1] - S.off[c_lo] <= p.obs_cap) lo = mid; else hi = mid - 1;
            }
            c_hi = lo;
        }
        const int co = S.off[c_lo], ce = S.off[c_hi + 1];
        const bool use_hints = p.hints != 0 && have_map && !(p.skip & 32);
        const int co4 = co & ~3;                         // 16-byte aligned start of the labels' slice
        int* sbk = sbk_raw + (co - co4);                 // label of each observation (-1 far), same slot as its hint
        if (tid == 0) {
            const uint32_t bytes = (p.skip & 128) ? 0u : (uint32_t)(ce - co) * 16u;
            const uint32_t hbytes = (use_hints && ce > co) ? (uint32_t)((ce - co4 + 3) & ~3) * 4u : 0u;
            if (bytes + hbytes > 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes + hbytes) : "memory");
                if (bytes) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                        ::"r"(smem_u32(sb)), "l"(p.bxy + co), "r"(bytes), "r"(mb) : "memory");
                // last sweep's labels of the same observations (the hints) land in the label array itself
                if (hbytes) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(smem_u32(sbk_raw)), "l"(p.c + co4), "r"(hbytes), "r"(mb) : "memory");
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
            }
        }
        // the slot's scan and this thread's half of it (indices relative to the chunk)
        const bool mine = qvalid && qli >= c_lo && qli <= c_hi;
        int o = 0, e = 0, so = 0, se = 0;
        if (mine) {
            so = S.off[qli] - co; se = S.off[qli + 1] - co;
            const int h1 = (se - so + TPP - 1) / TPP;
            o = min(so + sub * h1, se);
            e = min(o + h1, se);
        }
        for (int i = o; i < e; ++i) slt[i] = (unsigned char)qli;     // (while the bulk copies are in flight)
        if (!proj_early && c_lo == lt_first) proj_params();          // (likewise; the barrier below publishes them)
        FS_STAMP(3);
        mbar_wait(mb, parity);
        parity ^= 1u;
        __syncthreads();
        FS_STAMP(4);
        // ---- phase A: lanes over consecutive observations (adjacent lanes = adjacent beams), four per lane in flight --
        //   A0 (hints): last sweep's label of the same observation, carried through the filter's renumbering, is
        //       accepted when the observation lies inside the landmark's proven-nearest radius (tail.cuh hint_radius2):
        //       one staged value and one gather instead of the grid search.  Lanes it cannot settle stay pending for
        //   A1: project, locate the grid cell, fetch its entry range     (first level of gathers)
        //   A2: fetch the candidates, pick the nearest, gate, label      (second level of gathers)
        {
            const int m = ce - co;
            const int per = ((((m + FS_WARPS - 1) / FS_WARPS) + FS_U * 32 - 1) / (FS_U * 32)) * (FS_U * 32);
            const int wa = warp * per, wb = min(wa + per, m);
            unsigned long long pend = 0ull;
            const bool ident = p.ts->remap_identity != 0;
            const int lsearch = p.st->lsearch;
            int slot = 0;
            for (int base = wa; base < wb; base += FS_U * 32, slot += FS_U) {
                int h_[FS_U];
                double wx_[FS_U], wy_[FS_U];
#pragma unroll
                for (int u = 0; u < FS_U; ++u) {
                    const int i = base + u * 32 + lane;
                    h_[u] = -1; wx_[u] = 0.0; wy_[u] = 0.0;
                    if (i < wb && use_hints) {
                        h_[u] = sbk[i];
                        if (h_[u] >= p.cfg.L) h_[u] = -1;          // (labels of scans this handle does not own are not maintained)
                        const double2 bq = sb[i];
                        const int lt = slt[i];
                        const double2 pq = S.pp[lt], rc = S.rsc[lt];
                        wx_[u] = add_rn(__fma_rn(bq.y, -rc.x, mul_rn(bq.x, rc.y)), pq.x);
                        wy_[u] = add_rn(__fma_rn(bq.y, rc.y, mul_rn(bq.x, rc.x)), pq.y);
                    }
                }
                if (!ident) {
#pragma unroll
                    for (int u = 0; u < FS_U; ++u) if (h_[u] >= 0) h_[u] = __ldg(p.remap + h_[u]);
                }
                double2 q_[FS_U];
                double r2_[FS_U];
#pragma unroll
                for (int u = 0; u < FS_U; ++u) {
                    q_[u] = make_double2(0.0, 0.0); r2_[u] = -1.0;
                    if (h_[u] >= 0 && h_[u] < lsearch) {
                        const double2* rp = reinterpret_cast<const double2*>(p.lmrec + h_[u]);
                        q_[u] = __ldg(rp); r2_[u] = __ldg(rp + 1).x;
                    }
                }
#pragma unroll
                for (int u = 0; u < FS_U; ++u) {
                    const int i = base + u * 32 + lane;
                    if (i < wb) {
                        const double d2 = dist2_rn(q_[u].x - wx_[u], q_[u].y - wy_[u]);
                        const bool ok = d2 <= r2_[u];
                        if (ok) {
                            sbk[i] = h_[u];
                            // certified slack of this label, in 1/65535 of dist_thr, rounded down (the grid search below leaves 0)
                            if constexpr (CERT) if (p.stamp) srn[i] = (unsigned short)min(__double2int_rd((r2_[u] - d2) * p.marg_scale), 65535);
                            if (!ident && slt[i] >= 2) p.c[co + i] = h_[u];    // (with the identity renumbering c[] already holds it)
                        } else pend |= 1ull << (slot + u);
                    }
                }
            }
            const bool any_pend = __any_sync(FULLMASK, pend != 0ull) && !(p.skip & 32);
            if (p.skip & 32) for (int i = wa + lane; i < wb; i += 32) sbk[i] = -1;
            slot = 0;
            for (int base = wa; base < wb && any_pend; base += FS_U * 32, slot += FS_U) {
                int s_[FS_U], e_[FS_U];
#pragma unroll
                for (int u = 0; u < FS_U; ++u) {
                    const int i = base + u * 32 + lane;
                    s_[u] = 0; e_[u] = 0;
                    if (i < wb && have_map && ((pend >> (slot + u)) & 1ull)) {
                        const double2 bq = sb[i];
                        const int lt = slt[i];
                        const double2 pq = S.pp[lt], rc = S.rsc[lt];
                        const double wx = add_rn(__fma_rn(bq.y, -rc.x, mul_rn(bq.x, rc.y)), pq.x);
                        const double wy = add_rn(__fma_rn(bq.y, rc.y, mul_rn(bq.x, rc.x)), pq.y);
                        const int cc = fgrid_cell(G.g, wx, wy);
                        s_[u] = __ldg(G.cell_start + cc); e_[u] = __ldg(G.cell_start + cc + 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < FS_U; ++u) {
                    const int i = base + u * 32 + lane;
                    if (i < wb && ((pend >> (slot + u)) & 1ull)) { sbk[i] = s_[u]; srn[i] = (unsigned short)min(e_[u] - s_[u], 65535); }
                }
            }
            slot = 0;
            for (int base = wa; base < wb && any_pend; base += FS_U * 32, slot += FS_U) {
                double wx_[FS_U], wy_[FS_U];
                double2 p_[FS_U];
                int id_[FS_U], s_[FS_U], n_[FS_U];
#pragma unroll
                for (int u = 0; u < FS_U; ++u) {
                    const int i = base + u * 32 + lane;
                    s_[u] = 0; n_[u] = 0; id_[u] = -1; p_[u] = make_double2(0.0, 0.0); wx_[u] = 0.0; wy_[u] = 0.0;
                    if (i < wb && ((pend >> (slot + u)) & 1ull)) {
                        s_[u] = sbk[i]; n_[u] = (p.skip & 1) ? 0 : srn[i];
                        const double2 bq = sb[i];
                        const int lt = slt[i];
                        const double2 pq = S.pp[lt], rc = S.rsc[lt];
                        // tras_rot_z: numpy's matmul order, acc = a0*b0; acc = fma(a1, b1, acc); + translation
                        wx_[u] = add_rn(__fma_rn(bq.y, -rc.x, mul_rn(bq.x, rc.y)), pq.x);
                        wy_[u] = add_rn(__fma_rn(bq.y, rc.y, mul_rn(bq.x, rc.x)), pq.y);
                        if (n_[u] > 0) { p_[u] = __ldg(G.pts + s_[u]); id_[u] = __ldg(G.idx + s_[u]); }
                    }
                }
#pragma unroll
                for (int u = 0; u < FS_U; ++u) {
                    const int i = base + u * 32 + lane;
                    if (i < wb && ((pend >> (slot + u)) & 1ull)) {
                        double best;
                        int bid;
                        const int bk = fgrid_scan_pre(G, wx_[u], wy_[u], s_[u], n_[u], p_[u], id_[u], best, bid);
                        const bool far = bk < 0 || best > p.thr2_hi;      // amin > dist_thr (ICM_SLAM.py:172)
                        sbk[i] = far ? -1 : bid;                          // the label (index in the previous map)
                        if constexpr (CERT) srn[i] = 0;                   // (no certificate for a searched label)
                        if (slt[i] >= 2) p.c[co + i] = far ? -1 : bid;    // the halo scan (lt == 1) is not owned
                    }
                }
            }
        }
        FS_STAMP(5);
        __syncthreads();     // the labels of the whole chunk are visible to the pose threads
        FS_STAMP(6);
        // ---- phase B, pass 1 (thread pair per pose): runs of equal winners -> in-place run records --------
        const bool stamping = CERT && p.stamp;
        if (!(p.skip & 2) && o < e) {
            // (the next observation is fetched before the current one is consumed; the body-frame moments of the scan are
            //  static and come from k_body_moments).  When certificates are stamped, srn[] holds each observation's slack
            //  (16 bits) on entry; a run head leaves with beams | (least slack of the run >> 8) << 8.
            int run_start = o, bk = sbk[o];
            double2 b = sb[o];
            int sl = stamping ? srn[o] : 0, rsl = 65535;
            double Sbx = 0.0, Sby = 0.0;
            for (int i = o; i < e; ++i) {
                const int inx = min(i + 1, e - 1);
                const double2 bn = sb[inx];
                const int bkn = sbk[inx];
                const int sln = stamping ? srn[inx] : 0;
                Sbx += b.x; Sby += b.y;
                rsl = min(rsl, sl);
                if (i + 1 == e || bkn != bk) {
                    sb[run_start] = make_double2(Sbx, Sby);
                    srn[run_start] = stamping ? (unsigned short)((i + 1 - run_start) | (rsl & 0xff00)) : (unsigned short)(i + 1 - run_start);
                    run_start = i + 1; Sbx = 0.0; Sby = 0.0; rsl = 65535;
                }
                b = bn; bk = bkn; sl = sln;
            }
        }
        FS_STAMP(7);
        // ---- phase B, pass 2: one run per step, lanes in lockstep: moments + landmark statistics ---------
        {
            const double2 pq = mine ? S.pp[qli] : make_double2(0.0, 0.0), rc = mine ? S.rsc[qli] : make_double2(0.0, 1.0);
            const double px = pq.x, py = pq.y, st = rc.x, ct = rc.y;
            const int nmask = stamping ? 0xff : 0xffff;       // (beams <= 255 whenever certificates are stamped: checked on the host)
            int i = (p.skip & 6) ? e : o;
            int nw = 0, bk = -1;
            double2 sxy = make_double2(0.0, 0.0), lm = make_double2(0.0, 0.0), lg = make_double2(0.0, 0.0);
            if (i < e) {
                nw = srn[i]; bk = sbk[i]; sxy = sb[i];
                if (bk >= 0) {
                    const double2* rp = reinterpret_cast<const double2*>(p.lmrec + bk);
                    lm = __ldg(rp);
                    if (stamping) lg = __ldg(rp + 1);
                }
            }
            while (__any_sync(FULLMASK, i < e)) {
                const bool act = i < e;
                const int n = nw & nmask;
                const int i2 = i + n;                     // the next run's record and landmark are fetched first
                int nw2 = 0, bk2 = -1;
                double2 sxy2 = make_double2(0.0, 0.0), lm2 = make_double2(0.0, 0.0), lg2 = make_double2(0.0, 0.0);
                if (act && i2 < e) {
                    nw2 = srn[i2]; bk2 = sbk[i2]; sxy2 = sb[i2];
                    if (bk2 >= 0) {
                        const double2* rp = reinterpret_cast<const double2*>(p.lmrec + bk2);
                        lm2 = __ldg(rp);
                        if (stamping) lg2 = __ldg(rp + 1);
                    }
                }
                const bool matched = act && bk >= 0;
                const double dn = (double)n;
                const double rwx = fma(ct, sxy.x, -st * sxy.y), rwy = fma(st, sxy.x, ct * sxy.y);   // sum of rotated beams
                const double yx = lm.x - px, yy = lm.y - py;
                if (act && bk < 0) {          // far run: statistics of the scan's new label
                    nfar += n;
                    fsx += fma(dn, px, rwx); fsy += fma(dn, py, rwy);
                    FBx += sxy.x; FBy += sxy.y;
                }
                if (matched) {
                    M.Yx = fma(dn, yx, M.Yx); M.Yy = fma(dn, yy, M.Yy);
                    M.Mxx = fma(yx, sxy.x, M.Mxx); M.Mxy = fma(yx, sxy.y, M.Mxy);
                    M.Myx = fma(yy, sxy.x, M.Myx); M.Myy = fma(yy, sxy.y, M.Myy);
                }
                stat_add(S, p, matched && qowned && !(p.skip & 8), bk, rwx - dn * yx, rwy - dn * yy, n);   // sum of (obs - landmark)
                if constexpr (CERT) if (stamping && act) {
                    // the run cache: this half's records in order, from its end of the scan's slice.  The landmark's odometer is
                    // rounded DOWN and the slack was rounded down, so the certificate errs on the side of the full path.
                    const int gi = sub == 0 ? co + so + myruns : co + se - 1 - myruns;
                    const float slack = bk >= 0 ? __double2float_rd((double)(nw >> 8) * p.slack_unit) : 0.0f;
                    p.rsum[gi] = sxy;
                    p.rmeta[gi] = make_int4(bk, n, __float_as_int(__double2float_rd(lg.y)), __float_as_int(slack));
                    ++myruns;
                }
                i = i2; nw = nw2; bk = bk2; sxy = sxy2; lm = lm2; lg = lg2;
            }
        }
        if constexpr (CERT) {     // stamp the scans of this chunk
            const int rpart = __shfl_xor_sync(FULLMASK, myruns, 1);
            if (mine && sub == 0 && p.stamp) {
                const bool pin = qt == 0 && p.first;
                p.rcnt[qt] = myruns | (rpart << 16);
                p.echk[qt] = epoch;
                p.xchk[qt] = pin ? p.x0[0] : S.xs[0][qli]; p.xchk[p.ldchk + qt] = pin ? p.x0[1] : S.xs[1][qli]; p.xchk[2 * p.ldchk + qt] = pth_;
            }
            myruns = 0;
        }
        FS_STAMP(8);
        c_lo = c_hi + 1;
        if (c_lo <= lt_last) __syncthreads();   // the next chunk overwrites the staging buffers
    }
    // ---- combine the pair's partial sums (both lanes end up with the scan's totals) ----------------------------
    int far_idx = -1;
    {
#define FS_PAIR(v) { v += __shfl_xor_sync(FULLMASK, v, 1); if (TPP == 4) v += __shfl_xor_sync(FULLMASK, v, 2); }
        FS_PAIR(M.Yx); FS_PAIR(M.Yy); FS_PAIR(M.Mxx); FS_PAIR(M.Mxy); FS_PAIR(M.Myx); FS_PAIR(M.Myy);
        FS_PAIR(fsx); FS_PAIR(fsy); FS_PAIR(FBx); FS_PAIR(FBy);
        nfar += __shfl_xor_sync(FULLMASK, nfar, 1);
        if (TPP == 4) nfar += __shfl_xor_sync(FULLMASK, nfar, 2);
#undef FS_PAIR
        if (sub == 0 && qowned && nfar > 0) {
            // a scan with far observations creates one label (ICM_SLAM.py:174-182): its record now, its rank in time
            // order within the tile once every slot has reported (after the solve; only the record's index is kept)
            atomicOr(&S.farbits[(qt - tb) >> 5], 1u << ((qt - tb) & 31));
            FarRec r;
            r.t = qt; r.rank = 0; r.n = nfar; r.pad = 0; r.sx = fsx; r.sy = fsy;
            far_idx = atomicAdd(&p.ts->far_count, 1);
            p.far_list[far_idx] = r;
        }
        if (nfar > 0) {   // far observations see the mean of the scan's new label (PREV view, raw = sum / k)
            const double px = S.pp[qli].x, py = S.pp[qli].y;
            const double yx = fsx / (double)nfar - px, yy = fsy / (double)nfar - py;
            M.Yx += (double)nfar * yx; M.Yy += (double)nfar * yy;
            M.Mxx = fma(yx, FBx, M.Mxx); M.Mxy = fma(yx, FBy, M.Mxy);
            M.Myx = fma(yy, FBx, M.Myx); M.Myy = fma(yy, FBy, M.Myy);
        }
    }
    FS_STAMP(9);
    if (SPLIT) {
        // ---- split mode: hand the scan's landmark moments to k_pose_solve ---------------------------------------------
        if (qvalid && sub == 0) {
            double* d = p.dyn + cm_index(qt, p.Th);
            d[0] = M.Yx; d[p.lddyn] = M.Yy; d[2 * p.lddyn] = M.Mxx; d[3 * p.lddyn] = M.Mxy; d[4 * p.lddyn] = M.Myx; d[5 * p.lddyn] = M.Myy;
        }
        __syncthreads();     // every slot has reported its far observations
    }
    // ---- pose solve by the lane pairs: warps 0-3 red (odd poses), then warps 4-7 black (even poses) --------------
    if constexpr (!SPLIT) solve_phases<HALF, TPP>(S, p, M, q, sub, qt, qli, qvalid);
    FS_STAMP(13);
    if (far_idx >= 0) {
        const int bit = qt - tb, w = bit >> 5;
        int rank = __popc(S.farbits[w] & ((1u << (bit & 31)) - 1u));
        for (int k = 0; k < w; ++k) rank += __popc(S.farbits[k]);
        p.far_list[far_idx].rank = rank;
    }
    FS_STAMP(14);
    if (tid == 0) p.blk_far[blockIdx.x] = __popc(S.farbits[0]) + __popc(S.farbits[1]) + __popc(S.farbits[2]) + __popc(S.farbits[3]);
    FS_STAMP(15);
    // ---- outputs ----------------------------------------------------------------------------------
    if (!SPLIT) {
        const int n_own = min(FS_OWN, p.t_hi - tb);
        for (int r = 0; r < 3; ++r)
            for (int k = tid; k < n_own; k += FS_THREADS) p.xout[r * p.ldout + tb + k] = S.xn[r][k + 2];
    }
    FS_STAMP(16);
    for (int h = tid; h < FS_HASH; h += FS_THREADS) {
        const int key = S.hkey[h];
        if (key >= 0 && S.hcnt[h] > 0) {
            atomicAdd((unsigned long long*)(p.fsum_x + key), ((unsigned long long)S.hs[h][1] << 32) | S.hs[h][0]);
            atomicAdd((unsigned long long*)(p.fsum_y + key), ((unsigned long long)S.hs[h][3] << 32) | S.hs[h][2]);
            atomicAdd(p.cnt + key, S.hcnt[h]);
        }
    }
    FS_STAMP(17);
}

// ---- split mode: the pose solve as two launches, one per colour -----------------------------------------------------
// k_sweep_fused<.., SPLIT = true> leaves, per scan, the landmark moments (p.dyn) and the sin/cos of the input heading
// (p.sc).  The solve then needs no tile, no shared memory and no barrier: k_solve_colour<0> solves every odd pose from its
// OLD even neighbours (a lane pair per pose, everything read straight from global memory), k_solve_colour<1> every even
// pose from the NEW odd poses.  Same arithmetic as solve_phases, so the results are identical to the single-launch kernel.
// A segment re-solves the odd halo pose t_lo-1 itself (its moments come from the segment's first tile).
template <int COLOUR, int OCC>
__global__ void __launch_bounds__(128, OCC / 128)
k_solve_colour(const FusedParams p)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 1, half = threadIdx.x & 1;
    const int T = p.T;
    const int t = COLOUR == 0 ? (p.t_lo >= 2 ? p.t_lo - 1 : 1) + 2 * g : p.t_lo + 2 * g;
    const bool qvalid = t < p.t_hi;
    const double* X = COLOUR == 0 ? p.xin : p.xout;          // the neighbours: old poses for the odd colour, new ones for the even
    const int64_t ldX = COLOUR == 0 ? p.ldin : p.ldout;
    const double* SC = COLOUR == 0 ? p.sc : p.scn;           // ... and the sin/cos of the previous pose's heading
    Mom M;
    M.n = M.Bx = M.By = M.Bxx = M.Byy = M.Bxy = M.Yx = M.Yy = M.Mxx = M.Mxy = M.Myx = M.Myy = 0.0;
    PoseIn P;
    P.ax = P.ay = P.ath = P.sa = 0.0; P.ca = 1.0; P.bx = P.by = P.bth = 0.0; P.uav = P.uaw = P.ucv = P.ucw = 0.0;
    P.D0x = P.D0y = P.dth0 = P.D1x = P.D1y = P.dth1 = 0.0; P.has_next = 0;
    double ox = 0.0, oy = 0.0, th0 = 0.0, s0 = 0.0, c0 = 1.0;
    bool solve = false;
    const bool pinned = qvalid && t == 0;                    // x[:,0] is never updated (sensors.py:131)
    const int64_t ct = cm_index(t, p.Th), cp = cm_index(t - 1, p.Th);     // colour-major positions of pose t and of pose t-1
    if (qvalid) {
        ox = p.xin[t]; oy = p.xin[p.ldin + t]; th0 = p.xin[2 * p.ldin + t];
        s0 = p.sc[ct]; c0 = p.sc[p.ldsc + ct];
        if (!pinned) {
            const bool has_next = t + 1 < T;
            P.has_next = has_next ? 1 : 0;
            const double* bmq = p.bm + ct;
            M.n = __ldg(bmq + 5 * p.ldbm);
            P.ax = X[t - 1]; P.ay = X[ldX + t - 1]; P.ath = X[2 * ldX + t - 1];
            P.sa = SC[cp]; P.ca = SC[p.ldsc + cp];
            if (has_next) { P.bx = X[t + 1]; P.by = X[ldX + t + 1]; P.bth = X[2 * ldX + t + 1]; }
            P.uav = __ldg(p.u_cm + cp); P.uaw = __ldg(p.u_cm + p.ldbm + cp); P.ucv = __ldg(p.u_cm + ct); P.ucw = __ldg(p.u_cm + p.ldbm + ct);
            P.D0x = __ldg(p.inc_cm + cp); P.D0y = __ldg(p.inc_cm + p.ldbm + cp); P.dth0 = __ldg(p.inc_cm + 2 * p.ldbm + cp);
            P.D1x = __ldg(p.inc_cm + ct); P.D1y = __ldg(p.inc_cm + p.ldbm + ct); P.dth1 = __ldg(p.inc_cm + 2 * p.ldbm + ct);
            M.Bx = __ldg(bmq); M.By = __ldg(bmq + p.ldbm); M.Bxx = __ldg(bmq + 2 * p.ldbm);
            M.Byy = __ldg(bmq + 3 * p.ldbm); M.Bxy = __ldg(bmq + 4 * p.ldbm);
            const double* d = p.dyn + ct;
            M.Yx = d[0]; M.Yy = d[p.lddyn]; M.Mxx = d[2 * p.lddyn]; M.Mxy = d[3 * p.lddyn]; M.Myx = d[4 * p.lddyn]; M.Myy = d[5 * p.lddyn];
            solve = M.n > 0.0;
        }
    }
    double res = half ? oy : ox, th = th0, s_new = s0, c_new = c0;
    if (qvalid && !pinned && !solve) {       // sensors.py:147-151: no observation, average of the neighbours
        const bool t1 = t == 1 && p.first;
        const double pv = t1 ? p.x0[half] : (half ? P.ay : P.ax);
        res = (pv + (half ? P.by : P.bx)) / 2.0;
        th = ((t1 ? p.x0[2] : P.ath) + P.bth) / 2.0;
        sincos(th, &s_new, &c_new);
    }
    {
        // (lanes that do not solve still run the iteration with harmless values: newton_trig's shuffles need the whole warp)
        double r2 = 0.0, th2 = th0, s2 = 0.0, c2 = 1.0;
        const int it = newton_trig(p.cfg, P, M, half, ox, oy, th2, s0, c0, p.tol, (solve && !(p.skip & 16)) ? p.maxit : 1, r2, s2, c2);
        if (solve) { res = r2; th = th2; s_new = s2; c_new = c2; }
        if (p.iters) {
            const int tot = warp_sum_i((solve && half == 0) ? it : 0);
            if ((threadIdx.x & 31) == 0 && tot) atomicAdd(p.iters, (unsigned long long)tot);
        }
    }
    if (qvalid) {
        p.xout[(half ? p.ldout : 0) + t] = res;
        if (half == 0) p.xout[2 * p.ldout + t] = th;
        if (COLOUR == 0) p.scn[(half ? p.ldsc : 0) + ct] = half ? c_new : s_new;
    }
}

// interleaves the extraction's (bx, by) arrays into the double2 records the fused kernel stages
__global__ void k_interleave(const double* __restrict__ bx, const double* __restrict__ by, int64_t n, double2* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(bx[i], by[i]);
}

static size_t fused_smem_fixed(int half, bool split)
{
    if (split) return half == 16 ? sizeof(FusedSmemFixed<16, true>) : half == 32 ? sizeof(FusedSmemFixed<32, true>) : sizeof(FusedSmemFixed<64, true>);
    return half == 16 ? sizeof(FusedSmemFixed<16>) : half == 32 ? sizeof(FusedSmemFixed<32>) : sizeof(FusedSmemFixed<64>);
}
static size_t fused_smem_bytes(int half, bool split, int obs_cap)   // obs_cap is even
{
    return fused_smem_fixed(half, split) + (size_t)obs_cap * 23 + 64;
}

// the compiled variants: (pose slots per colour, threads per slot, resident threads per SM)
#define FUSED_DEFAULT_HALF 64
#define FUSED_DEFAULT_TPP 2
#define FUSED_DEFAULT_OCC 512
typedef void (*fused_kernel_t)(const FusedParams);
static fused_kernel_t fused_variant(int half, int tpp, int occ, bool split)
{
#define FV(H, P, O) if (half == H && tpp == P && occ == O) return split ? k_sweep_fused<H, P, O, true> : k_sweep_fused<H, P, O, false>;
    FV(64, 2, 512) FV(32, 2, 512) FV(16, 2, 512) FV(32, 2, 640) FV(32, 2, 768) FV(64, 2, 768)
#undef FV
    return nullptr;
}
static fused_kernel_t solve_variant(int colour, int occ)     // k_solve_colour: 64 poses of one colour per 128-thread block
{
    if (occ == 1024) return colour ? k_solve_colour<1, 1024> : k_solve_colour<0, 1024>;
    if (occ == 512) return colour ? k_solve_colour<1, 512> : k_solve_colour<0, 512>;
    return colour ? k_solve_colour<1, 768> : k_solve_colour<0, 768>;
}
#define SOLVE_THREADS 128
#undef FS_THREADS
#undef FS_WARPS
#undef FS_SLOTS
#undef FS_OWN
#undef FS_XT
#undef FS_HALF
#undef FS_HASH
#undef FS_HASH_SHIFT
#undef FS_MINBLK

// raw map of the previous-map landmarks from the fixed-point statistics + keep flags for all labels;
// clears the statistics for the next sweep.
__global__ void __launch_bounds__(256)
k_fused_means(const DevState* st, long long* fsum_x, long long* fsum_y, const int* __restrict__ cnt,
              const double* __restrict__ map_x, const double* __restrict__ map_y, double inv_scale, double cota,
              double* newraw /* 2 x Lcap: means of this sweep's new labels, zero elsewhere; cleared here.  ALIASES fsum_x / fsum_y
                                (old labels use a word as int64 sum, new labels as double mean: disjoint index ranges) */,
              double* __restrict__ raw_x, double* __restrict__ raw_y, int* __restrict__ flag, int Lcap)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= Lcap) return;
    const int raw_l = st->raw_l, ls = st->lsearch;
    const int k = l < raw_l ? cnt[l] : 0;
    if (l < ls) {
        raw_x[l] = k > 0 ? map_x[l] + ((double)fsum_x[l] * inv_scale) / (double)k : 0.0;
        raw_y[l] = k > 0 ? map_y[l] + ((double)fsum_y[l] * inv_scale) / (double)k : 0.0;
    } else {
        const bool have = l < raw_l && k > 0;
        raw_x[l] = have ? newraw[l] : 0.0;
        raw_y[l] = have ? newraw[Lcap + l] : 0.0;
    }
    newraw[l] = 0.0; newraw[Lcap + l] = 0.0;
    fsum_x[l] = 0; fsum_y[l] = 0;
    flag[l] = (l < raw_l && !((double)k < cota)) ? 1 : 0;      // ICM_SLAM.py:232-236
}
