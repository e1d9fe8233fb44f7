/*
 * icm_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference ICM-SLAM hot path.
 *
 * This is the parity oracle: a plain-C restatement of the numpy/scipy algorithm of
 * Seba-san/icm-slam for the offline sweep path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product package
 * (icm_slam_b200/) never does and has no CPU fallback.
 *
 * PINNING: checked against fixtures minted by running the UNMODIFIED reference modules
 * (oracle/ref_runner.py + oracle/make_golden.py -> tests/golden/*.npz): extraction sets,
 * association label streams, raw/filtered maps, counts, energies, Nelder-Mead poses.  The
 * reference itself has no tests, so those fixtures are the pin.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/scripts/).  Arithmetic is IEEE fp64, compiled with -ffp-contract=off so no
 * FMA contraction changes the numpy operation order.
 *
 * Layout conventions: numpy C-order.  A "3xT" array is three rows of length ld (>= T).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct orc_config {
    double deltat;          /* ConfigICM.deltat  ICM_SLAM.py:73 */
    double q1, q2;          /* diag(Q)           ICM_SLAM.py:76-78 (used as weights, sensors.py:202) */
    double r1, r2, r3;      /* diag(R)           ICM_SLAM.py:80-83 (weights, sensors.py:240) */
    double cte_odom;        /* ICM_SLAM.py:85 */
    double cota;            /* ICM_SLAM.py:86 */
    double dist_thr;        /* ICM_SLAM.py:87 */
    double rango_laser_max; /* ICM_SLAM.py:89 */
    double radio;           /* ICM_SLAM.py:90 */
    int32_t L;              /* label capacity, ICM_SLAM.py:74 */
    int32_t pad_;
} orc_config;

enum { ORC_SCHED_SEQUENTIAL = 0, ORC_SCHED_REDBLACK = 1 };
enum { ORC_SOLVER_NM = 0, ORC_SOLVER_NEWTON = 1 };
enum { ORC_VIEW_RUNNING = 0, ORC_VIEW_FULL = 1, ORC_VIEW_PREV = 2 };

enum {
    ORC_OK = 0,
    ORC_ERR_LABEL_CAP = -2,    /* reference: IndexError at ICM_SLAM.py:191 when label >= config.L */
    ORC_ERR_EMPTY_LAST = -3,   /* reference: IndexError at sensors.py:148 (x[:,t+1] with t==T-1) */
    ORC_ERR_EMPTY_MAP = -4,    /* reference: ValueError at ICM_SLAM.py:241-255 (nothing survives) */
    ORC_ERR_ALLOC = -5
};

static const double ORC_PI = 3.141592653589793; /* np.pi */

/* ------------------------------------------------------------------------------------------
 * entrepi  (ICM_SLAM.py:455-463): np.mod(angle, 2*pi), then -2*pi if > pi.
 * np.mod on doubles = fmod with the result moved to the divisor's sign (npy_divmod). */
static double entrepi(double a)
{
    const double twopi = 2 * ORC_PI;
    double m = fmod(a, twopi);
    if (m != 0.0) {
        if (m < 0.0) m += twopi;
    } else {
        m = 0.0;
    }
    if (m > ORC_PI) m = m - twopi;
    return m;
}

ORC_API double orc_entrepi(double a) { return entrepi(a); }

/* g  (sensors.py:206-211): unicycle Euler step; S@u keeps the "+0*w" / "0*v+" terms of matmul. */
static void g_step(const double p[3], double v, double w, double dt, double out[3])
{
    double c = cos(p[2]), s = sin(p[2]);
    out[0] = p[0] + dt * (c * v + 0.0 * w);
    out[1] = p[1] + dt * (s * v + 0.0 * w);
    out[2] = p[2] + dt * (0.0 * v + 1.0 * w);
}

ORC_API void orc_g(const orc_config* cfg, const double p[3], const double u[2], double out[3])
{
    g_step(p, u[0], u[1], cfg->deltat, out);
}

/* h  (sensors.py:175-204): sum_i (p_xy + d_i*(cos,sin)(alpha_i+theta-pi/2) - seen_i)^T Q (...) */
static double h_obs(const orc_config* cfg, const double x[3], int n, const double* d,
                    const double* alpha, const double* sx, const double* sy)
{
    double pot = 0.0;
    for (int i = 0; i < n; ++i) {
        double alfa = alpha[i] + x[2] - ORC_PI / 2.0;
        double zc = d[i] * cos(alfa), zs = d[i] * sin(alfa);
        double dx = (x[0] + zc) - sx[i];
        double dy = (x[1] + zs) - sy[i];
        double ax = dx * cfg->q1 + dy * 0.0; /* distancias @ Q, Q diagonal */
        double ay = dx * 0.0 + dy * cfg->q2;
        pot += ax * dx;
        pot += ay * dy;
    }
    return pot;
}

/* Rota(phi) @ v   (ICM_SLAM.py:482-488): [[c,s],[-s,c]] */
static void rota_mul(double phi, double vx, double vy, double out[2])
{
    double c = cos(phi), s = sin(phi);
    out[0] = c * vx + s * vy;
    out[1] = -s * vx + c * vy;
}

/* gg^T R gg for diagonal R, in matmul order: (gg^T R) then @ gg. */
static double quad3(const double g[3], double r1, double r2, double r3)
{
    double a0 = g[0] * r1 + g[1] * 0.0 + g[2] * 0.0;
    double a1 = g[0] * 0.0 + g[1] * r2 + g[2] * 0.0;
    double a2 = g[0] * 0.0 + g[1] * 0.0 + g[2] * r3;
    return a0 * g[0] + a1 * g[1] + a2 * g[2];
}

/* fun_x  (sensors.py:266-282): causal energy (pass 0 and the last pose of a sweep).
 * odo0 = odometry[:,t-1], odo1 = odometry[:,t]. */
ORC_API double orc_fun_x(const orc_config* cfg, const double x[3], const double x_ant[3],
                         const double u_ant[2], const double odo0[3], const double odo1[3], int n,
                         const double* d, const double* alpha, const double* sx, const double* sy)
{
    double ga[3], gg[3], o1[2], o2[2], ooo[3];
    g_step(x_ant, u_ant[0], u_ant[1], cfg->deltat, ga);
    gg[0] = x[0] - ga[0];
    gg[1] = x[1] - ga[1];
    gg[2] = entrepi(x[2] - ga[2]);
    double hh = h_obs(cfg, x, n, d, alpha, sx, sy);
    rota_mul(odo0[2], odo1[0] - odo0[0], odo1[1] - odo0[1], o1);
    rota_mul(x_ant[2], x[0] - x_ant[0], x[1] - x_ant[1], o2);
    ooo[0] = o1[0] - o2[0];
    ooo[1] = o1[1] - o2[1];
    ooo[2] = entrepi(odo1[2] - odo0[2] - x[2] + x_ant[2]);
    double oo = ooo[0] * ooo[0] + ooo[1] * ooo[1] + ooo[2] * ooo[2];
    return quad3(gg, cfg->r1, cfg->r2, cfg->r3) + hh + cfg->cte_odom * oo;
}

/* fun_xn  (sensors.py:224-256): smoothing energy of an interior pose.
 * odo0/1/2 = odometry[:,t-1], [:,t], [:,t+1]; u_ant = u[:,t-1], u_act = u[:,t]. */
ORC_API double orc_fun_xn(const orc_config* cfg, const double x[3], const double x_ant[3],
                          const double x_pos[3], const double u_ant[2], const double u_act[2],
                          const double odo0[3], const double odo1[3], const double odo2[3], int n,
                          const double* d, const double* alpha, const double* sx, const double* sy)
{
    double gx[3], gg[3], o1[2], o2[2], ooo[3];
    g_step(x, u_act[0], u_act[1], cfg->deltat, gx);
    gg[0] = gx[0] - x_pos[0];
    gg[1] = gx[1] - x_pos[1];
    gg[2] = entrepi(gx[2] - x_pos[2]);
    rota_mul(odo1[2], odo2[0] - odo1[0], odo2[1] - odo1[1], o1);
    rota_mul(x[2], x_pos[0] - x[0], x_pos[1] - x[1], o2);
    ooo[0] = o1[0] - o2[0];
    ooo[1] = o1[1] - o2[1];
    ooo[2] = entrepi(odo2[2] - odo1[2] - x_pos[2] + x[2]);
    double oo = ooo[0] * ooo[0] + ooo[1] * ooo[1] + ooo[2] * ooo[2];
    double f = quad3(gg, cfg->r1, cfg->r2, cfg->r3) + cfg->cte_odom * oo;

    double ga[3];
    g_step(x_ant, u_ant[0], u_ant[1], cfg->deltat, ga);
    gg[0] = x[0] - ga[0];
    gg[1] = x[1] - ga[1];
    gg[2] = entrepi(x[2] - ga[2]);
    double hh = h_obs(cfg, x, n, d, alpha, sx, sy);
    rota_mul(odo0[2], odo1[0] - odo0[0], odo1[1] - odo0[1], o1);
    rota_mul(x_ant[2], x[0] - x_ant[0], x[1] - x_ant[1], o2);
    ooo[0] = o1[0] - o2[0];
    ooo[1] = o1[1] - o2[1];
    ooo[2] = entrepi(odo1[2] - odo0[2] - x[2] + x_ant[2]);
    oo = ooo[0] * ooo[0] + ooo[1] * ooo[1] + ooo[2] * ooo[2];
    f = f + quad3(gg, cfg->r1, cfg->r2, cfg->r3) + hh + cfg->cte_odom * oo;
    return f;
}

/* ------------------------------------------------------------------------------------------
 * Pose problem context shared by the two inner solvers. */
typedef struct pose_problem {
    const orc_config* cfg;
    int has_next; /* 1: fun_xn, 0: fun_x */
    double x_ant[3], x_pos[3], u_ant[2], u_act[2], odo0[3], odo1[3], odo2[3];
    int n;
    const double *d, *alpha, *sx, *sy;
    long nev;
} pose_problem;

static double pose_energy(const double x[3], pose_problem* P)
{
    P->nev++;
    if (P->has_next)
        return orc_fun_xn(P->cfg, x, P->x_ant, P->x_pos, P->u_ant, P->u_act, P->odo0, P->odo1,
                          P->odo2, P->n, P->d, P->alpha, P->sx, P->sy);
    return orc_fun_x(P->cfg, x, P->x_ant, P->u_ant, P->odo0, P->odo1, P->n, P->d, P->alpha, P->sx,
                     P->sy);
}

/* Nelder-Mead, restating scipy.optimize.fmin -> _minimize_neldermead as the reference calls it
 * (sensors.py:221,263: xtol=1e-3, default ftol=1e-4, maxiter=maxfun=3*200).  scipy is a pinned
 * third-party dependency (requisitos.txt:21, scipy==1.5.4; this container: 1.18.1) absent from
 * /root/reference; constants rho=1, chi=2, psi=sigma=0.5, nonzdelt=0.05, zdelt=0.00025. */
static void nm_sort4(double sim[4][3], double fs[4])
{
    for (int i = 1; i < 4; ++i) { /* stable insertion sort == np.argsort on 4 elements */
        double f = fs[i], v[3] = {sim[i][0], sim[i][1], sim[i][2]};
        int j = i - 1;
        while (j >= 0 && fs[j] > f) {
            fs[j + 1] = fs[j];
            memcpy(sim[j + 1], sim[j], sizeof v);
            --j;
        }
        fs[j + 1] = f;
        memcpy(sim[j + 1], v, sizeof v);
    }
}

static void nelder_mead(pose_problem* P, const double start[3], double xtol, double ftol,
                        double out[3])
{
    const int N = 3, maxiter = 600;
    const long maxfun = 600;
    double sim[4][3], fs[4];
    long nev0 = P->nev;
    for (int k = 0; k < 3; ++k) sim[0][k] = start[k];
    for (int k = 0; k < N; ++k) {
        for (int j = 0; j < 3; ++j) sim[k + 1][j] = start[j];
        if (sim[k + 1][k] != 0.0)
            sim[k + 1][k] = (1 + 0.05) * sim[k + 1][k];
        else
            sim[k + 1][k] = 0.00025;
    }
    for (int k = 0; k < 4; ++k) fs[k] = pose_energy(sim[k], P);
    nm_sort4(sim, fs);
    int iterations = 1;
    while ((P->nev - nev0) < maxfun && iterations < maxiter) {
        double dxm = 0.0, dfm = 0.0;
        for (int k = 1; k < 4; ++k) {
            for (int j = 0; j < 3; ++j) {
                double a = fabs(sim[k][j] - sim[0][j]);
                if (a > dxm) dxm = a;
            }
            double b = fabs(fs[0] - fs[k]);
            if (b > dfm) dfm = b;
        }
        if (dxm <= xtol && dfm <= ftol) break;
        double xbar[3], xr[3], xe[3], xc[3];
        for (int j = 0; j < 3; ++j) xbar[j] = ((sim[0][j] + sim[1][j]) + sim[2][j]) / N;
        for (int j = 0; j < 3; ++j) xr[j] = (1 + 1) * xbar[j] - 1 * sim[3][j];
        double fxr = pose_energy(xr, P);
        int doshrink = 0;
        if (fxr < fs[0]) {
            for (int j = 0; j < 3; ++j) xe[j] = (1 + 1 * 2) * xbar[j] - 1 * 2 * sim[3][j];
            double fxe = pose_energy(xe, P);
            if (fxe < fxr) {
                memcpy(sim[3], xe, sizeof xe);
                fs[3] = fxe;
            } else {
                memcpy(sim[3], xr, sizeof xr);
                fs[3] = fxr;
            }
        } else if (fxr < fs[2]) {
            memcpy(sim[3], xr, sizeof xr);
            fs[3] = fxr;
        } else {
            if (fxr < fs[3]) {
                for (int j = 0; j < 3; ++j) xc[j] = (1 + 0.5 * 1) * xbar[j] - 0.5 * 1 * sim[3][j];
                double fxc = pose_energy(xc, P);
                if (fxc <= fxr) {
                    memcpy(sim[3], xc, sizeof xc);
                    fs[3] = fxc;
                } else
                    doshrink = 1;
            } else {
                for (int j = 0; j < 3; ++j) xc[j] = (1 - 0.5) * xbar[j] + 0.5 * sim[3][j];
                double fxcc = pose_energy(xc, P);
                if (fxcc < fs[3]) {
                    memcpy(sim[3], xc, sizeof xc);
                    fs[3] = fxcc;
                } else
                    doshrink = 1;
            }
            if (doshrink) {
                for (int k = 1; k < 4; ++k) {
                    for (int j = 0; j < 3; ++j)
                        sim[k][j] = sim[0][j] + 0.5 * (sim[k][j] - sim[0][j]);
                    fs[k] = pose_energy(sim[k], P);
                }
            }
        }
        iterations++;
        nm_sort4(sim, fs);
    }
    for (int j = 0; j < 3; ++j) out[j] = sim[0][j];
}

/* "Exact" inner solver (the restated mode the CUDA path is compared against; NOT in the
 * reference, which uses the NM above).  For fixed theta the energy of sensors.py:224-282 is
 * quadratic in (x,y) with a diagonal Hessian, so it is minimised in closed
 * form and theta is found by Newton on the reduced 1-D function phi(theta).  Terms are
 * evaluated one by one (centres c_k(theta), weights W_k) -- deliberately not via the moment
 * sums the CUDA kernel uses.  Stops when |dtheta| <= tol. */
typedef struct { double w, c, c1, c2; } vp_term; /* weight, centre and its 1st/2nd theta-derivative */

static void newton_pose(pose_problem* P, const double start[3], double tol, int maxit,
                        double out[3], int* iters)
{
    const orc_config* cfg = P->cfg;
    const double dt = cfg->deltat, k = cfg->cte_odom;
    double ga[3], D0[2], D1[2], e0c[2];
    g_step(P->x_ant, P->u_ant[0], P->u_ant[1], dt, ga);
    rota_mul(P->odo0[2], P->odo1[0] - P->odo0[0], P->odo1[1] - P->odo0[1], D0);
    { /* centre of the e0 term: a_xy + Rota(a_th)^T D0 */
        double c = cos(P->x_ant[2]), s = sin(P->x_ant[2]);
        e0c[0] = P->x_ant[0] + (c * D0[0] - s * D0[1]);
        e0c[1] = P->x_ant[1] + (s * D0[0] + c * D0[1]);
    }
    if (P->has_next) rota_mul(P->odo1[2], P->odo2[0] - P->odo1[0], P->odo2[1] - P->odo1[1], D1);
    double th = start[2], xs = start[0], ys = start[1];
    const double ox = start[0], oy = start[1]; /* local origin: keeps the variance-like sums small */
    int it = 0;
    for (; it < maxit; ++it) {
        double s = sin(th), c = cos(th);
        /* accumulators: S=sum W, A=sum W c, A1=sum W c', and the pieces of phi', phi'' */
        double Sx = 0, Ax = 0, Ax1 = 0, Ax2 = 0, Ccx1 = 0, Cx2 = 0;
        double Sy = 0, Ay = 0, Ay1 = 0, Ay2 = 0, Ccy1 = 0, Cy2 = 0;
#define ADDX(W, C, C1, C2)                                                              \
    do {                                                                                \
        double w_ = (W), c_ = (C) - ox, c1_ = (C1), c2_ = (C2);                         \
        Sx += w_; Ax += w_ * c_; Ax1 += w_ * c1_; Ax2 += w_ * c2_;                      \
        Ccx1 += w_ * c_ * c1_; Cx2 += w_ * (c1_ * c1_ + c_ * c2_);                      \
    } while (0)
#define ADDY(W, C, C1, C2)                                                              \
    do {                                                                                \
        double w_ = (W), c_ = (C) - oy, c1_ = (C1), c2_ = (C2);                         \
        Sy += w_; Ay += w_ * c_; Ay1 += w_ * c1_; Ay2 += w_ * c2_;                      \
        Ccy1 += w_ * c_ * c1_; Cy2 += w_ * (c1_ * c1_ + c_ * c2_);                      \
    } while (0)
        ADDX(cfg->r1, ga[0], 0, 0);
        ADDY(cfg->r2, ga[1], 0, 0);
        ADDX(k, e0c[0], 0, 0);
        ADDY(k, e0c[1], 0, 0);
        if (P->has_next) {
            double v = P->u_act[0];
            ADDX(cfg->r1, P->x_pos[0] - dt * v * c, dt * v * s, dt * v * c);
            ADDY(cfg->r2, P->x_pos[1] - dt * v * s, -dt * v * c, dt * v * s);
            ADDX(k, P->x_pos[0] - (c * D1[0] - s * D1[1]), s * D1[0] + c * D1[1],
                 c * D1[0] - s * D1[1]);
            ADDY(k, P->x_pos[1] - (s * D1[0] + c * D1[1]), -(c * D1[0] - s * D1[1]),
                 s * D1[0] + c * D1[1]);
        }
        for (int i = 0; i < P->n; ++i) {
            double bx = P->d[i] * cos(P->alpha[i]), by = P->d[i] * sin(P->alpha[i]);
            double wx = bx * s + by * c, wy = -bx * c + by * s; /* Rot(th - pi/2) b */
            /* centre = seen - w ; w' = (bx c - by s, bx s + by c) = (-wy, wx) ; w'' = -w */
            ADDX(cfg->q1, P->sx[i] - wx, wy, wx);
            ADDY(cfg->q2, P->sy[i] - wy, -wx, wy);
        }
#undef ADDX
#undef ADDY
        xs = Ax / Sx + ox;
        ys = Ay / Sy + oy;
        double w1 = entrepi(th - ga[2]);
        double w3 = entrepi(P->odo1[2] - P->odo0[2] - th + P->x_ant[2]);
        double ang1 = 2 * cfg->r3 * w1 - 2 * k * w3;
        double ang2 = 2 * cfg->r3 + 2 * k;
        if (P->has_next) {
            double w2 = entrepi(th + dt * P->u_act[1] - P->x_pos[2]);
            double w4 = entrepi(P->odo2[2] - P->odo1[2] - P->x_pos[2] + th);
            ang1 += 2 * cfg->r3 * w2 + 2 * k * w4;
            ang2 += 2 * cfg->r3 + 2 * k;
        }
        /* phi = sum W c^2 - A^2/S ;  phi' = 2 sum W c c' - 2 A A'/S ;
         * phi'' = 2 sum W (c'^2 + c c'') - 2 (A'^2 + A A'')/S */
        double p1 = 2 * Ccx1 - 2 * Ax * Ax1 / Sx + 2 * Ccy1 - 2 * Ay * Ay1 / Sy + ang1;
        double p2 = 2 * Cx2 - 2 * (Ax1 * Ax1 + Ax * Ax2) / Sx + 2 * Cy2 -
                    2 * (Ay1 * Ay1 + Ay * Ay2) / Sy + ang2;
        if (!(p2 > 0)) p2 = ang2; /* safeguard far from the minimum */
        double dth = -p1 / p2;
        th += dth;
        if (fabs(dth) <= tol) { ++it; break; }
    }
    { /* closed-form xy at the final theta */
        double s = sin(th), c = cos(th);
        double Sx = cfg->r1 + k, Ax = cfg->r1 * (ga[0] - ox) + k * (e0c[0] - ox);
        double Sy = cfg->r2 + k, Ay = cfg->r2 * (ga[1] - oy) + k * (e0c[1] - oy);
        if (P->has_next) {
            double v = P->u_act[0];
            Sx += cfg->r1 + k;
            Sy += cfg->r2 + k;
            Ax += cfg->r1 * (P->x_pos[0] - dt * v * c - ox) +
                  k * (P->x_pos[0] - (c * D1[0] - s * D1[1]) - ox);
            Ay += cfg->r2 * (P->x_pos[1] - dt * v * s - oy) +
                  k * (P->x_pos[1] - (s * D1[0] + c * D1[1]) - oy);
        }
        for (int i = 0; i < P->n; ++i) {
            double bx = P->d[i] * cos(P->alpha[i]), by = P->d[i] * sin(P->alpha[i]);
            Sx += cfg->q1;
            Sy += cfg->q2;
            Ax += cfg->q1 * (P->sx[i] - (bx * s + by * c) - ox);
            Ay += cfg->q2 * (P->sy[i] - (-bx * c + by * s) - oy);
        }
        xs = Ax / Sx + ox;
        ys = Ay / Sy + oy;
    }
    out[0] = xs;
    out[1] = ys;
    out[2] = th;
    if (iters) *iters = it;
}

/* Solve one pose problem from Python (used to pin the solvers against reference fmin output). */
ORC_API long orc_solve_pose(const orc_config* cfg, int solver, int has_next, const double x_ant[3],
                            const double x_pos[3], const double u_ant[2], const double u_act[2],
                            const double odo0[3], const double odo1[3], const double odo2[3], int n,
                            const double* d, const double* alpha, const double* sx,
                            const double* sy, double out[3])
{
    pose_problem P;
    memset(&P, 0, sizeof P);
    P.cfg = cfg;
    P.has_next = has_next;
    memcpy(P.x_ant, x_ant, 24);
    memcpy(P.u_ant, u_ant, 16);
    memcpy(P.odo0, odo0, 24);
    memcpy(P.odo1, odo1, 24);
    if (has_next) {
        memcpy(P.x_pos, x_pos, 24);
        memcpy(P.u_act, u_act, 16);
        memcpy(P.odo2, odo2, 24);
    }
    P.n = n; P.d = d; P.alpha = alpha; P.sx = sx; P.sy = sy;
    double start[3];
    if (has_next) {
        for (int j = 0; j < 3; ++j) start[j] = (x_ant[j] + x_pos[j]) / 2.0; /* sensors.py:221 */
    } else {
        g_step(x_ant, u_ant[0], u_ant[1], cfg->deltat, start);              /* sensors.py:262 */
    }
    if (solver == ORC_SOLVER_NM)
        nelder_mead(&P, start, 1e-3, 1e-4, out);
    else
        newton_pose(&P, start, 1e-14, 60, out, 0);
    return P.nev;
}

/* ------------------------------------------------------------------------------------------
 * filtrar_z  (ICM_SLAM.py:22-58) on one scan.  z has B beams with element stride `stride`
 * (a column of the BxT `mediciones`).  ang/cosb/sinb are the beam tables
 * ang[i]=(i*pi)/180, cos(ang[i]), sin(ang[i]) as numpy computes them (ICM_SLAM.py:44,51-53);
 * they are passed in so the host's numpy values are used verbatim.
 * Outputs (capacity B): beam index, d (median-filtered range), bx=d*cos, by=d*sin.
 * Returns n_t (0 covers both the `np.array([])` and the `(0,4)` outcome). */
ORC_API int orc_filtrar_z(const double* z, int B, long stride, const double* cosb,
                          const double* sinb, double rango_laser_max, double dist_thr,
                          int32_t* beam, double* d, double* bx, double* by)
{
    double* zf = (double*)malloc(sizeof(double) * (size_t)B * 3);
    int32_t* nind = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
    if (!zf || !nind) { free(zf); free(nind); return ORC_ERR_ALLOC; }
    double *px = zf + B, *py = zf + 2 * B;
    /* scipy.signal.medfilt kernel 3, zero padded ends (ICM_SLAM.py:37) */
    for (int i = 0; i < B; ++i) {
        double a = i > 0 ? z[(long)(i - 1) * stride] : 0.0;
        double b = z[(long)i * stride];
        double c = i + 1 < B ? z[(long)(i + 1) * stride] : 0.0;
        double lo = a < b ? a : b, hi = a < b ? b : a;
        zf[i] = c < lo ? lo : (c > hi ? hi : c);
    }
    int k = 0;
    for (int i = 0; i < B; ++i)
        if (zf[i] < rango_laser_max) nind[k++] = i; /* :41 */
    int n = 0;
    if (k > 1) {
        for (int j = 0; j < k; ++j) { /* :44-45 */
            px[j] = cosb[nind[j]] * zf[nind[j]];
            py[j] = sinb[nind[j]] * zf[nind[j]];
        }
        for (int j = 0; j < k; ++j) { /* :46-50: min over the other points; zeros become 100 */
            double m = INFINITY;
            for (int i = 0; i < k; ++i) {
                double dx = px[i] - px[j], dy = py[i] - py[j];
                double dd = sqrt(dx * dx + dy * dy);
                if (dd == 0.0) dd = 100.0;
                if (dd < m) m = dd;
            }
            if (m <= dist_thr) {
                int b_ = nind[j];
                beam[n] = b_;
                d[n] = zf[b_];
                bx[n] = zf[b_] * cosb[b_]; /* :52-53 */
                by[n] = zf[b_] * sinb[b_];
                ++n;
            }
        }
    }
    free(zf);
    free(nind);
    return n;
}

/* filtrar_z over every column of the BxT array -> CSR.  off has T+1 entries. */
ORC_API long orc_extract_all(const double* med, int B, int T, long ld, const double* cosb,
                             const double* sinb, double rango_laser_max, double dist_thr,
                             int32_t* off, int32_t* beam, double* d, double* bx, double* by)
{
    long n = 0;
    off[0] = 0;
    for (int t = 0; t < T; ++t) {
        int nt = orc_filtrar_z(med + t, B, ld, cosb, sinb, rango_laser_max, dist_thr, beam + n,
                               d + n, bx + n, by + n);
        if (nt < 0) return nt;
        n += nt;
        off[t + 1] = (int32_t)n;
    }
    return n;
}

/* tras_rot_z  (ICM_SLAM.py:465-480): world = b @ [[ct,st],[-st,ct]] + xy, ct=cos(th-pi/2). */
ORC_API void orc_tras_rot(const double pose[3], int n, const double* bx, const double* by,
                          double* wx, double* wy)
{
    double ct = cos(pose[2] - ORC_PI / 2.0), st = sin(pose[2] - ORC_PI / 2.0);
    for (int i = 0; i < n; ++i) {
        /* numpy's matmul accumulates acc = a0*b0; acc = fma(a1, b1, acc) on this host (verified
         * bit-for-bit against tests/golden/units.npz tr_out); the CUDA path uses __fma_rn. */
        wx[i] = fma(by[i], -st, bx[i] * ct) + pose[0];
        wy[i] = fma(by[i], ct, bx[i] * st) + pose[1];
    }
}

/* ------------------------------------------------------------------------------------------
 * Mapa state (ICM_SLAM.py:104-126). */
typedef struct orc_mapa {
    int32_t Lact;  /* landmarks_actuales */
    int32_t L;     /* capacity */
    double* cant;  /* cant_obs_i, length L */
} orc_mapa;


ORC_API orc_mapa* orc_mapa_new(int L)
{
    orc_mapa* M = (orc_mapa*)calloc(1, sizeof *M);
    if (!M) return 0;
    M->L = L;
    M->cant = (double*)calloc((size_t)(L > 0 ? L : 1), sizeof(double));
    return M;
}
ORC_API void orc_mapa_free(orc_mapa* M) { if (M) { free(M->cant); free(M); } }
ORC_API void orc_mapa_clear_obs(orc_mapa* M) { memset(M->cant, 0, sizeof(double) * (size_t)M->L); } /* :119-126 */
ORC_API int orc_mapa_get_lact(const orc_mapa* M) { return M->Lact; }
ORC_API void orc_mapa_set_lact(orc_mapa* M, int v) { M->Lact = v; }
ORC_API double* orc_mapa_counts(orc_mapa* M) { return M->cant; }

/* Mapa.actualizar, Branch B (ICM_SLAM.py:167-194).
 * mapa: 2 x L running-mean buffer (ld = L).  ref: 2 x Lref, leading dimension ldref; only the
 * first min(Lact, Lref) columns are searched (the `[:,:Lact]` slice clips to the array width).
 * Far observations (min_dist > dist_thr, strict) of one scan ALL receive the single new label
 * Lact: `ztt[:,2:4]` of a (k,2) array is (k,0), so pdist is all zeros and fcluster returns one
 * cluster (ICM_SLAM.py:174-180).  Then the recursive running mean :184-194.
 * Returns 0 or ORC_ERR_LABEL_CAP (reference: IndexError when the new label >= config.L). */
ORC_API int orc_actualizar(orc_mapa* M, double dist_thr, double* mapa, const double* ref, int Lref,
                           long ldref, int n, const double* wx, const double* wy, int32_t* c)
{
    int Lact = M->Lact, L = M->L;
    int Ls = Lact < Lref ? Lact : Lref;
    int nfar = 0;
    for (int i = 0; i < n; ++i) {
        double best = INFINITY;
        int arg = 0;
        for (int l = 0; l < Ls; ++l) { /* cdist euclidean + argmin: first minimum wins */
            double dx = ref[l] - wx[i], dy = ref[ldref + l] - wy[i];
            double dd = sqrt(dx * dx + dy * dy);
            if (dd < best) { best = dd; arg = l; }
        }
        if (best > dist_thr) { c[i] = -1; ++nfar; } else c[i] = arg;
    }
    if (nfar > 0) {
        if (Lact >= L) return ORC_ERR_LABEL_CAP;
        for (int i = 0; i < n; ++i) if (c[i] < 0) c[i] = Lact;
        Lact = Lact + 1; /* :182 */
    }
    for (int lbl = 0; lbl < Lact; ++lbl) { /* :184-194 */
        int k = 0;
        double sx = 0.0, sy = 0.0;
        for (int i = 0; i < n; ++i)
            if (c[i] == lbl) { sx += wx[i]; sy += wy[i]; ++k; }
        if (k > 0) {
            double ni = M->cant[lbl], tot = ni + (double)k;
            mapa[lbl] = sx / tot + mapa[lbl] * ni / tot;
            mapa[L + lbl] = sy / tot + mapa[L + lbl] * ni / tot;
            M->cant[lbl] = tot;
        }
    }
    M->Lact = Lact;
    return ORC_OK;
}

/* Mapa.filtrar (ICM_SLAM.py:204-265).  mapa: 2 x L (ld = L).  out: 2 x L, zero filled beyond the
 * new Lact.  Updates M->Lact and M->cant.  O(K^2) like the reference.
 * Deviation (documented): when no label is pruned the reference indexes a 2 x L buffer with a
 * length-Lact mask and raises IndexError; here the first Lact columns are used. */
ORC_API int orc_filtrar(orc_mapa* M, double cota, double dist_thr, const double* mapa, double* out)
{
    int Lact = M->Lact, L = M->L;
    int K = 0;
    double* kx = (double*)malloc(sizeof(double) * (size_t)(Lact + 1) * 3);
    int32_t* c = (int32_t*)malloc(sizeof(int32_t) * (size_t)(Lact + 1) * 2);
    if (!kx || !c) { free(kx); free(c); return ORC_ERR_ALLOC; }
    double *ky = kx + (Lact + 1), *kc = kx + 2 * (Lact + 1);
    int32_t* b = c + (Lact + 1);
    for (int i = 0; i < Lact; ++i) /* :231-239 */
        if (!(M->cant[i] < cota)) { kx[K] = mapa[i]; ky[K] = mapa[L + i]; kc[K] = M->cant[i]; ++K; }
    if (K == 0) { free(kx); free(c); return ORC_ERR_EMPTY_MAP; }
    /* :241-245 */
    double amax = 0.0;
    for (int i = 0; i < K; ++i)
        for (int j = i + 1; j < K; ++j) {
            double dx = kx[i] - kx[j], dy = ky[i] - ky[j];
            double dd = sqrt(dx * dx + dy * dy);
            if (dd > amax) amax = dd;
        }
    double* amin = (double*)malloc(sizeof(double) * (size_t)K);
    if (!amin) { free(kx); free(c); return ORC_ERR_ALLOC; }
    for (int j = 0; j < K; ++j) {
        double best = INFINITY;
        int arg = 0;
        for (int i = 0; i < K; ++i) {
            double dx = kx[i] - kx[j], dy = ky[i] - ky[j];
            double dd = (i == j) ? 0.0 : sqrt(dx * dx + dy * dy);
            if (dd == 0.0) dd = amax;
            if (dd < best) { best = dd; arg = i; }
        }
        amin[j] = best;
        b[j] = arg;
        c[j] = j;
    }
    for (int i = 0; i < K; ++i) /* :247-249, ascending i over ind */
        if (amin[i] < dist_thr) {
            int from = c[b[i]], to = c[i];
            for (int j = 0; j < K; ++j) if (c[j] == from) c[j] = to;
        }
    for (int i = K - 1; i >= 0; --i) { /* :251-253 */
        int present = 0;
        for (int j = 0; j < K; ++j) if (c[j] == i) { present = 1; break; }
        if (!present)
            for (int j = 0; j < K; ++j) if (c[j] >= i) c[j] -= 1;
    }
    int newL = 0;
    for (int j = 0; j < K; ++j) if (c[j] + 1 > newL) newL = c[j] + 1;
    memset(out, 0, sizeof(double) * 2 * (size_t)L);
    memset(M->cant, 0, sizeof(double) * (size_t)L);
    for (int i = 0; i < newL; ++i) { /* :258-260 */
        double cs = 0.0, sx = 0.0, sy = 0.0;
        for (int j = 0; j < K; ++j)
            if (c[j] == i) { cs += kc[j]; sx += kx[j] * kc[j]; sy += ky[j] * kc[j]; }
        M->cant[i] = cs;
        out[i] = sx / cs;
        out[L + i] = sy / cs;
    }
    M->Lact = newL;
    free(amin); free(kx); free(c);
    return ORC_OK;
}

/* calc_cambio (ICM_SLAM.py:490-495): for each NEW landmark the distance to the nearest OLD one;
 * returns min, max, mean.  y: 2 x Ln (ldn), old: 2 x Lo (ldo). */
ORC_API void orc_calc_cambio(const double* y, int Ln, long ldn, const double* old, int Lo, long ldo,
                             double out[3])
{
    double mn = INFINITY, mx = -INFINITY, sum = 0.0;
    for (int j = 0; j < Ln; ++j) {
        double best = INFINITY;
        for (int i = 0; i < Lo; ++i) {
            double dx = old[i] - y[j], dy = old[ldo + i] - y[ldn + j];
            double dd = sqrt(dx * dx + dy * dy);
            if (dd < best) best = dd;
        }
        if (best < mn) mn = best;
        if (best > mx) mx = best;
        sum += best;
    }
    out[0] = mn; out[1] = mx; out[2] = sum / (double)Ln;
}

/* ------------------------------------------------------------------------------------------
 * The sweep: ICM_ROS.iterations_process_offline (sensors.py:125-168), with the extraction
 * (filtrar_z, sweep-invariant) hoisted into a CSR computed once by orc_extract_all.
 *
 * mode (ORC_SCHED_SEQUENTIAL, ORC_SOLVER_NM, ORC_VIEW_RUNNING) is the reference's own
 * semantics.  The other switch values are the restated variants of SURVEY.md 7.1:
 *   schedule redblack : odd poses first (from the sweep's input neighbours), then even poses
 *                       (from the updated odd ones); t=0 is never updated (sensors.py:145).
 *   solver   newton   : exact conditional minimiser (newton_pose above).
 *   view     full     : a pose sees the final per-label mean of the whole sweep.
 *   view     prev     : a pose sees the previous (input) map's landmark it was matched to;
 *                       labels created in this sweep see their own single-scan mean.
 * Association / labels / running means / raw map never depend on the pose updates (they use the
 * input poses, sensors.py:153), so they are computed first for all t -- same results.
 *
 * x: 3 x T (ld = ldx), updated in place.  map_in: 2 x Lin (ld = ldm).  M->Lact on entry selects
 * the search width and the first new label.  Outputs: c[n] labels; seen_x/seen_y[n] the
 * landmark each observation was fitted against; raw_map 2 x L / raw counts in M before
 * filtrar are copied to raw_map/raw_counts/raw_L; map_out 2 x L and M hold the filtered map.
 * Returns ORC_OK, 1 if scan 0 is empty (reference returns its inputs unchanged, :137-139),
 * or a negative error. */
ORC_API int orc_sweep(const orc_config* cfg, orc_mapa* M, int T, const int32_t* off,
                      const int32_t* beam, const double* d, const double* bx, const double* by,
                      const double* ang, const double* odo, long ldo, const double* u, long ldu,
                      const double x0[3], const double* map_in, int Lin, long ldm, double* x,
                      long ldx, int schedule, int solver, int view, double newton_tol,
                      int32_t* c, double* seen_x, double* seen_y, double* raw_map,
                      double* raw_counts, int32_t* raw_L, double* map_out, long* nev_out)
{
    const int L = cfg->L;
    long n_all = off[T];
    (void)beam;
    orc_mapa_clear_obs(M); /* :133 */
    if (off[1] - off[0] == 0) return 1; /* :137-139 */
    if (off[T] - off[T - 1] == 0 && T > 1) return ORC_ERR_EMPTY_LAST;
    const int Lact0 = M->Lact;
    double* y = (double*)calloc((size_t)2 * L, sizeof(double)); /* :132 */
    double* wx = (double*)malloc(sizeof(double) * (size_t)(n_all + 1) * 3);
    if (!y || !wx) { free(y); free(wx); return ORC_ERR_ALLOC; }
    double* wy = wx + (n_all + 1);
    double* alpha = wy + (n_all + 1);
    for (long i = 0; i < n_all; ++i) alpha[i] = ang[beam[i]];
    int rc = ORC_OK;
    /* stage 1: projection with the INPUT poses + association + running mean, in time order */
    for (int t = 0; t < T && rc == ORC_OK; ++t) {
        int o = off[t], nt = off[t + 1] - off[t];
        if (nt == 0) continue;
        double pose[3];
        if (t == 0) { pose[0] = x0[0]; pose[1] = x0[1]; pose[2] = x0[2]; } /* :141 */
        else { pose[0] = x[t]; pose[1] = x[ldx + t]; pose[2] = x[2 * ldx + t]; } /* :153 */
        orc_tras_rot(pose, nt, bx + o, by + o, wx + o, wy + o);
        rc = orc_actualizar(M, cfg->dist_thr, y, map_in, Lin, ldm, nt, wx + o, wy + o, c + o);
        for (int i = 0; i < nt; ++i) { /* y[:,c].T right after actualizar (:156) */
            seen_x[o + i] = y[c[o + i]];
            seen_y[o + i] = y[L + c[o + i]];
        }
    }
    if (rc != ORC_OK) { free(y); free(wx); return rc; }
    if (view == ORC_VIEW_FULL) {
        for (long i = 0; i < n_all; ++i) { seen_x[i] = y[c[i]]; seen_y[i] = y[L + c[i]]; }
    } else if (view == ORC_VIEW_PREV) {
        for (long i = 0; i < n_all; ++i) {
            if (c[i] < Lact0 && c[i] < Lin) { seen_x[i] = map_in[c[i]]; seen_y[i] = map_in[ldm + c[i]]; }
            else { seen_x[i] = y[c[i]]; seen_y[i] = y[L + c[i]]; }
        }
    }
    /* stage 2: pose updates */
    long nev = 0;
    int npass = schedule == ORC_SCHED_REDBLACK ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
        double xt_prev[3] = {x0[0], x0[1], x0[2]}; /* `xt` of sensors.py:131 */
        for (int t = 1; t < T; ++t) {
            if (schedule == ORC_SCHED_REDBLACK) {
                if ((t & 1) != (pass == 0 ? 1 : 0)) continue;
                if (t > 1) { xt_prev[0] = x[t - 1]; xt_prev[1] = x[ldx + t - 1]; xt_prev[2] = x[2 * ldx + t - 1]; }
            }
            int o = off[t], nt = off[t + 1] - off[t];
            double res[3];
            if (nt == 0) { /* :147-151 */
                res[0] = (xt_prev[0] + x[t + 1]) / 2.0;
                res[1] = (xt_prev[1] + x[ldx + t + 1]) / 2.0;
                res[2] = (xt_prev[2] + x[2 * ldx + t + 1]) / 2.0;
            } else {
                pose_problem P;
                memset(&P, 0, sizeof P);
                P.cfg = cfg;
                P.has_next = (t + 1 < T);
                for (int j = 0; j < 3; ++j) {
                    P.x_ant[j] = x[j * ldx + t - 1];
                    P.odo0[j] = odo[j * ldo + t - 1];
                    P.odo1[j] = odo[j * ldo + t];
                }
                P.u_ant[0] = u[t - 1]; P.u_ant[1] = u[ldu + t - 1];
                double start[3];
                if (P.has_next) {
                    for (int j = 0; j < 3; ++j) {
                        P.x_pos[j] = x[j * ldx + t + 1];
                        P.odo2[j] = odo[j * ldo + t + 1];
                        start[j] = (P.x_ant[j] + P.x_pos[j]) / 2.0; /* :221 */
                    }
                    P.u_act[0] = u[t]; P.u_act[1] = u[ldu + t];
                } else {
                    g_step(P.x_ant, P.u_ant[0], P.u_ant[1], cfg->deltat, start); /* :262 */
                }
                P.n = nt; P.d = d + o; P.alpha = alpha + o; P.sx = seen_x + o; P.sy = seen_y + o;
                if (solver == ORC_SOLVER_NM) nelder_mead(&P, start, 1e-3, 1e-4, res);
                else newton_pose(&P, start, newton_tol, 60, res, 0);
                nev += P.nev;
            }
            x[t] = res[0]; x[ldx + t] = res[1]; x[2 * ldx + t] = res[2]; /* :162 */
            xt_prev[0] = res[0]; xt_prev[1] = res[1]; xt_prev[2] = res[2];
        }
    }
    if (nev_out) *nev_out = nev;
    /* stage 3: raw map snapshot, then filtrar (:165-166) */
    if (raw_L) *raw_L = M->Lact;
    if (raw_map) memcpy(raw_map, y, sizeof(double) * 2 * (size_t)L);
    if (raw_counts) memcpy(raw_counts, M->cant, sizeof(double) * (size_t)L);
    rc = orc_filtrar(M, cfg->cota, cfg->dist_thr, y, map_out);
    free(y);
    free(wx);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * filtrar_obs.m (scripts/filtrar_obs.m:6-50), the offline scan gate that produced
 * data_IJAC2018.mat from datos_palomar1.mat.  obs: B x T (ld), column = scan.
 *  :8-9   ranges > max_dist are invalid (NaN)
 *  :12-17 a(t) = number of valid beams, sentinel cant_max appended
 *  :23-27 entries with a > cant_max dropped, a = fix(interp1(tt, a, t, 'linear')) over t=1..T+1
 *  :33-48 keep the a(t) smallest ranges of each scan (stable ascending sort, NaN last)
 *  :50    invalid -> max_dist */
ORC_API int orc_filtrar_obs(const double* obs, int B, int T, long ld, double max_dist, int cant_max,
                            double* out, long ldout, int32_t* a_out)
{
    double* a = (double*)malloc(sizeof(double) * (size_t)(T + 1));
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
    if (!a || !idx) { free(a); free(idx); return ORC_ERR_ALLOC; }
    for (int t = 0; t < T; ++t) {
        int k = 0;
        for (int i = 0; i < B; ++i) {
            double v = obs[(long)i * ld + t];
            if (!(v > max_dist) && !isnan(v)) ++k;
        }
        a[t] = k;
    }
    a[T] = cant_max;
    /* linear interpolation over the kept knots (value <= cant_max), query every t */
    int prev = -1;
    for (int t = 0; t <= T;) {
        if (a[t] <= cant_max) { prev = t; ++t; continue; }
        int nxt = t;
        while (a[nxt] > cant_max) ++nxt; /* sentinel at T guarantees termination */
        for (int q = t; q < nxt; ++q) {
            if (prev < 0) { free(a); free(idx); return -6; } /* interp1 gives NaN before the first knot */
            double slope = (a[nxt] - a[prev]) / (double)(nxt - prev);
            double v = slope * (double)(q - prev) + a[prev];
            a[q] = -1.0 - trunc(v); /* mark, decoded below (fix = trunc) */
        }
        t = nxt;
    }
    for (int t = 0; t < T; ++t) {
        int keep = (int)(a[t] < 0 ? -(a[t] + 1.0) : a[t]);
        if (a_out) a_out[t] = keep;
        int k = 0;
        for (int i = 0; i < B; ++i) {
            double v = obs[(long)i * ld + t];
            if (!(v > max_dist) && !isnan(v)) idx[k++] = i;
        }
        for (int i = 1; i < k; ++i) { /* stable insertion sort by range */
            int id = idx[i];
            double v = obs[(long)id * ld + t];
            int j = i - 1;
            while (j >= 0 && obs[(long)idx[j] * ld + t] > v) { idx[j + 1] = idx[j]; --j; }
            idx[j + 1] = id;
        }
        for (int i = 0; i < B; ++i) out[(long)i * ldout + t] = max_dist;
        for (int i = 0; i < k && i < keep; ++i) out[(long)idx[i] * ldout + t] = obs[(long)idx[i] * ld + t];
    }
    free(a);
    free(idx);
    return ORC_OK;
}

/* range pre-conditioning (sensors_definitions.py:21-22 == IJAC2018_python.txt:43) */
ORC_API void orc_precondition(const double* z, long n, double radio, double rango_laser_max, double* out)
{
    for (long i = 0; i < n; ++i) {
        double v = isnan(z[i]) ? rango_laser_max : z[i];
        v = v + radio;
        out[i] = v < rango_laser_max ? v : rango_laser_max; /* np.minimum */
    }
}
