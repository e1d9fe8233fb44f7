"""CPU pin of the math of `newton_trig` / `k_solve_tile` (icm_slam_b200/csrc/solve.cuh): the reduced 1-D problem in theta as a
trigonometric polynomial whose four coefficients are formed once from the moment sums.

`_coefficients` below is a line-by-line numpy mirror of the device code (role 0 = x rows, role 1 = y rows).  It is held to the
oracle's own restatement of the reference energy (`orc.fun_xn` / `orc.fun_x`, sensors.py:224-282), which is pinned by the
golden fixtures: (i) phi'(theta) from the coefficients equals the derivative of the energy along the closed-form (x*, y*)(theta)
(envelope theorem), (ii) the root of phi' is the pose the oracle's exact solver returns."""
import numpy as np
import pytest

from helpers import CONFIG_ROS

PI = 3.141592653589793


def _entrepi(a):
    m = np.mod(a, 2 * PI)
    return m - 2 * PI if m > PI else m


def _coefficients(cfg, pose, a, b, u_ant, u_act, odo3, d, alpha, sx, sy, has_next):
    """Mirror of newton_trig's set-up.  Returns (coef, role constants, angular constants)."""
    dt, k = cfg.deltat, cfg.cte_odom
    ox, oy = pose[0], pose[1]
    bx, by = d * np.cos(alpha), d * np.sin(alpha)
    yx, yy = sx - ox, sy - oy
    M = dict(n=float(d.size), Bx=bx.sum(), By=by.sum(), Bxx=(bx * bx).sum(), Byy=(by * by).sum(), Bxy=(bx * by).sum(),
             Yx=yx.sum(), Yy=yy.sum(), Mxx=(yx * bx).sum(), Mxy=(yx * by).sum(), Myx=(yy * bx).sum(), Myy=(yy * by).sum())

    def inc(o0, o1):       # k_odo_increments: Rota(o0.theta) (o1.xy - o0.xy), dtheta
        s, c = np.sin(o0[2]), np.cos(o0[2])
        vx, vy = o1[0] - o0[0], o1[1] - o0[1]
        return c * vx + s * vy, -s * vx + c * vy, o1[2] - o0[2]
    D0x, D0y, dth0 = inc(odo3[:, 0], odo3[:, 1])
    D1x, D1y, dth1 = inc(odo3[:, 1], odo3[:, 2]) if has_next else (0.0, 0.0, 0.0)
    hn = 1.0 if has_next else 0.0
    sa, ca = np.sin(a[2]), np.cos(a[2])
    D1x, D1y = hn * D1x, hn * D1y
    dv = hn * dt * (u_act[0] if has_next else 0.0)
    co = np.zeros(4)       # as2, ac2, au2, av2 (already doubled at the end)
    roles = []
    for role in (0, 1):
        r, q = (cfg.r2, cfg.q2) if role else (cfg.r1, cfg.q1)
        o_ = oy if role else ox
        a_ = (a[1] if role else a[0]) - o_
        ga = a_ + dt * ((sa if role else ca) * u_ant[0])
        e0 = a_ + ((sa * D0x + ca * D0y) if role else (ca * D0x - sa * D0y))
        bp = hn * (((b[1] if role else b[0]) - o_) if has_next else 0.0)
        iS = 1.0 / (r + k + M["n"] * q + hn * (r + k))
        KA = r * ga + k * e0 + q * (M["Yy"] if role else M["Yx"]) + (r + k) * bp
        P1 = (r * dv + k * D1x) + q * M["By"]
        P2 = q * M["Bx"] - k * D1y
        M1, M2 = (M["Myx"], M["Myy"]) if role else (M["Mxx"], M["Mxy"])
        ba = iS * KA * P2 + (k * bp * D1y - q * M1)
        bb = -iS * KA * P1 + (bp * (r * dv + k * D1x) + q * M2)
        bab = -iS * (P2 * P2 - P1 * P1) + (q * (M["Bxx"] - M["Byy"]) - (r * dv * dv + k * (D1x * D1x - D1y * D1y)))
        bd = -iS * P1 * P2 + (q * M["Bxy"] - k * D1x * D1y)
        co += np.array([ba, -bb, -bab, -bd]) if role else np.array([bb, ba, bab, bd])
        roles.append(dict(KA=KA, P1=P1, P2=P2, iS=iS, o=o_))
    ang = dict(r3_2=2.0 * cfg.r3, k_2=2.0 * k, th_ga=a[2] + dt * u_ant[1], c3=dth0 + a[2],
               c4=(dth1 - b[2]) if has_next else 0.0, wb=(dt * u_act[1] - b[2]) if has_next else 0.0,
               ang2=(2.0 * cfg.r3 + 2.0 * k) * (1.0 + hn), has_next=has_next)
    return 2.0 * co, roles, ang


def _phi1_phi2(co, ang, th):
    s, c = np.sin(th), np.cos(th)
    u, v = s * c, c * c - s * s
    a1 = ang["r3_2"] * _entrepi(th - ang["th_ga"]) - ang["k_2"] * _entrepi(ang["c3"] - th)
    if ang["has_next"]:
        a1 += ang["r3_2"] * _entrepi(th + ang["wb"]) + ang["k_2"] * _entrepi(ang["c4"] + th)
    p1 = co[0] * s + co[1] * c + co[2] * u + co[3] * v + a1
    p2 = co[0] * c - co[1] * s + co[2] * v - 4.0 * co[3] * u + ang["ang2"]
    return p1, p2


def _xy(roles, th):
    s, c = np.sin(th), np.cos(th)
    x = (roles[0]["KA"] - c * roles[0]["P1"] - s * roles[0]["P2"]) * roles[0]["iS"] + roles[0]["o"]
    y = (roles[1]["KA"] - s * roles[1]["P1"] + c * roles[1]["P2"]) * roles[1]["iS"] + roles[1]["o"]
    return x, y


def _problem(rng, n):
    a = np.array([1.0, 2.0, 0.3]) + rng.normal(size=3) * [0.05, 0.05, 0.02]
    pose = a + np.array([0.19 * np.cos(a[2]), 0.19 * np.sin(a[2]), 0.01]) + rng.normal(size=3) * [0.02, 0.02, 0.01]
    b = pose + np.array([0.2 * np.cos(pose[2]), 0.2 * np.sin(pose[2]), -0.015]) + rng.normal(size=3) * [0.02, 0.02, 0.01]
    odo3 = np.stack([a, pose, b], axis=1) + rng.normal(size=(3, 3)) * 0.01 + np.array([[3.0], [-1.0], [0.2]])
    u_ant, u_act = np.array([2.0, 0.1]) + rng.normal(size=2) * 0.05, np.array([2.0, -0.15]) + rng.normal(size=2) * 0.05
    alpha = np.sort(rng.choice(181, size=n, replace=False)) * PI / 180.0
    d = rng.uniform(1.0, 9.0, size=n)
    wx = pose[0] + d * np.cos(alpha + pose[2] - PI / 2)
    wy = pose[1] + d * np.sin(alpha + pose[2] - PI / 2)
    sx, sy = wx + rng.normal(size=n) * 0.05, wy + rng.normal(size=n) * 0.05      # seen landmarks near the projected beams
    return pose, a, b, u_ant, u_act, odo3, d, alpha, sx, sy


@pytest.mark.parametrize("has_next", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_trigonometric_newton_matches_the_oracle_energy(seed, has_next):
    from oracle import oracle as orc
    cfg = orc.make_cfg(**dict(CONFIG_ROS, Q=[1.3, 0.8], R=[0.9, 1.2, 0.7], cte_odom=1.4))
    rng = np.random.default_rng(20181 + seed)
    pose, a, b, u_ant, u_act, odo3, d, alpha, sx, sy = _problem(rng, 14)
    co, roles, ang = _coefficients(cfg, pose, a, b, u_ant, u_act, odo3, d, alpha, sx, sy, has_next)

    def energy(th):
        x, y = _xy(roles, th)
        p = np.array([x, y, th])
        if has_next:
            return orc.fun_xn(cfg, p, a, b, u_ant, u_act, odo3, d, alpha, sx, sy)
        return orc.fun_x(cfg, p, a, u_ant, odo3[:, :2], d, alpha, sx, sy)

    # (i) phi' and phi'' against central differences of the oracle's energy along the closed-form (x*, y*)(theta)
    for th in pose[2] + np.array([-0.05, 0.0, 0.03]):
        p1, p2 = _phi1_phi2(co, ang, th)
        h = 1e-5
        num1 = (energy(th + h) - energy(th - h)) / (2 * h)
        num2 = (energy(th + h) - 2 * energy(th) + energy(th - h)) / (h * h)
        assert abs(p1 - num1) <= 1e-6 * max(1.0, abs(num1)), (th, p1, num1)
        assert abs(p2 - num2) <= 2e-3 * max(1.0, abs(num2)), (th, p2, num2)
    # (ii) Newton on phi' lands on the oracle's exact conditional minimiser
    th = pose[2]
    for _ in range(30):
        p1, p2 = _phi1_phi2(co, ang, th)
        step = -p1 / (p2 if p2 > 0 else ang["ang2"])
        th += step
        if abs(step) <= 1e-14:
            break
    x, y = _xy(roles, th)
    # (the oracle's solver starts at the midpoint of the neighbours / the motion prediction; same minimum)
    ref, _ = orc.solve_pose(cfg, "newton", a, b if has_next else None, u_ant, u_act if has_next else None,
                            odo3 if has_next else odo3[:, :2], d, alpha, sx, sy)
    assert abs(x - ref[0]) <= 1e-9 and abs(y - ref[1]) <= 1e-9 and abs(th - ref[2]) <= 1e-10, (x, y, th, ref)
