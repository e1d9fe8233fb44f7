"""Synthetic 2-D range-bearing datasets of the shapes BASELINE.json names.

The reference ships only two real logs (T=1833); the larger configurations
("100k poses / 10k landmarks", "1M poses / 100k landmarks", batches of trajectories) need a
generator.  This one is deterministic (numpy Generator seeded with 20181 + config_id) and
produces arrays in exactly the layout of the reference's `.mat` files
(`observations` B x T, `odometry` 3 x T, `velocities` 2 x T; createbag.py:124-127):

* landmarks ("trunks", radius 0.137 m) on a G x G grid of pitch 4 m with uniform jitter
  +-0.75 m (so any two are >= 2.5 m apart, > 2*dist_thr);
* a boustrophedon trajectory driven by unicycle controls (v = 2 m/s, dt = 0.1 s -> 0.2 m per
  pose), rows 8 m apart running midway between landmark rows, U-turns of radius 4 m; the true
  poses are the Euler integration `g` of the true controls (sensors.py:206-211), so
  T = 10 * L_true poses cover the field once;
* a 181-beam lidar at 1 deg over the forward half-plane (beam i looks along
  theta - pi/2 + i deg, the convention hard-wired in ICM_SLAM.py:44 and sensors.py:195),
  ray-cast against the trunk circles, no-return = 10.0, range noise N(0, 0.02^2), stored
  rounded to float32 like the real logs;
* controls = true (v, w) + N(0, [0.02, 0.01]^2); odometry = Euler integration of the noisy
  controls (it drifts; the energies only use its increments, sensors.py:237,252).

`x_init` / `map_init` stand in for the reference's causal pass 0 (sensors.py:51-123, which is
inherently sequential and not part of the sweeps/s metric): truth plus N(0, 0.05 m / 0.01 rad)
on poses and N(0, 0.1^2 m) on landmarks.
"""
from __future__ import annotations

import numpy as np

TRUNK_RADIUS = 0.137
RANGE_MAX = 10.0
PITCH = 4.0
JITTER = 0.75
ROW_SPACING = 8.0
SPEED = 2.0
DT = 0.1
TURN_RADIUS = ROW_SPACING / 2.0


def _controls(T: int, G: int):
    """True (v, w) per step for a boustrophedon over a field of side PITCH*G."""
    row_len = PITCH * G
    n_straight = int(round(row_len / (SPEED * DT)))
    w_turn = SPEED / TURN_RADIUS
    n_turn = int(round(np.pi / (w_turn * DT)))
    w_turn = np.pi / (n_turn * DT)  # exact half turn in n_turn Euler steps
    v = np.full(T, SPEED)
    w = np.zeros(T)
    t = 0
    sign = 1.0
    while t < T:
        t += n_straight
        if t >= T:
            break
        e = min(T, t + n_turn)
        w[t:e] = sign * w_turn
        t = e
        sign = -sign
    return v, w


def _integrate(v, w, start):
    """Euler unicycle, p[t+1] = g(p[t], u[t])."""
    T = v.shape[0]
    th = np.empty(T)
    th[0] = start[2]
    th[1:] = start[2] + np.cumsum(w[:-1] * DT)
    x = np.empty(T)
    y = np.empty(T)
    x[0], y[0] = start[0], start[1]
    x[1:] = start[0] + np.cumsum(v[:-1] * np.cos(th[:-1]) * DT)
    y[1:] = start[1] + np.cumsum(v[:-1] * np.sin(th[:-1]) * DT)
    return np.stack([x, y, th])


def _raycast(poses, lm, G, B):
    """Exact ray/circle ranges for every (pose, beam) of one chunk of poses: landmark-centric,
    using the grid layout (landmark gy*G + gx lives within JITTER of (PITCH*gx, PITCH*gy))."""
    T = poses.shape[1]
    scans = np.full((T, B), RANGE_MAX, dtype=np.float64)
    reach = RANGE_MAX + TRUNK_RADIUS + JITTER
    R = int(np.ceil(reach / PITCH))
    offs = np.arange(-R, R + 1)
    ox, oy = np.meshgrid(offs, offs, indexing="ij")
    ox = ox.ravel()
    oy = oy.ravel()
    kmax = 2 * int(np.ceil(np.degrees(np.arcsin(TRUNK_RADIUS / 1.0)))) + 2  # trunks are >= 1 m away
    ks = np.arange(kmax)
    px, py, th = poses[0], poses[1], poses[2]
    cgx = np.rint(px / PITCH).astype(np.int64)
    cgy = np.rint(py / PITCH).astype(np.int64)
    gx = cgx[:, None] + ox[None, :]
    gy = cgy[:, None] + oy[None, :]
    ok = (gx >= 0) & (gx < G) & (gy >= 0) & (gy < G)
    ti, ci = np.nonzero(ok)
    li = gy[ti, ci] * G + gx[ti, ci]
    dx = lm[0, li] - px[ti]
    dy = lm[1, li] - py[ti]
    D = np.hypot(dx, dy)
    keep = (D < RANGE_MAX + TRUNK_RADIUS) & (D > 1.0)
    ti, dx, dy, D = ti[keep], dx[keep], dy[keep], D[keep]
    beta = np.arctan2(dy, dx) - th[ti] + np.pi / 2.0  # bearing in beam coordinates
    beta = np.mod(beta + np.pi, 2 * np.pi) - np.pi
    gam = np.arcsin(TRUNK_RADIUS / D)
    lo = np.ceil(np.degrees(beta - gam)).astype(np.int64)
    hi = np.floor(np.degrees(beta + gam)).astype(np.int64)
    bi = lo[:, None] + ks[None, :]
    valid = (bi <= hi[:, None]) & (bi >= 0) & (bi < B)
    pi_, ki = np.nonzero(valid)
    b = bi[pi_, ki]
    delta = np.radians(b.astype(np.float64)) - beta[pi_]
    Dp = D[pi_]
    disc = TRUNK_RADIUS ** 2 - (Dp * np.sin(delta)) ** 2
    good = disc >= 0
    rho = Dp[good] * np.cos(delta[good]) - np.sqrt(disc[good])
    tt = ti[pi_][good]
    bb = b[good]
    inr = (rho > 0) & (rho < RANGE_MAX)
    np.minimum.at(scans, (tt[inr], bb[inr]), rho[inr])
    return scans


def _raycast_chunked(poses, lm, G, B, chunk=20000):
    T = poses.shape[1]
    out = np.empty((T, B), dtype=np.float64)
    for s in range(0, T, chunk):
        e = min(T, s + chunk)
        out[s:e] = _raycast(poses[:, s:e], lm, G, B)
    return out


def make_synthetic(L_true: int, T: int | None = None, seed: int = 20181, beams: int = 181,
                   range_noise: float = 0.02, ctrl_noise=(0.02, 0.01), init_pose_noise=(0.05, 0.01),
                   init_map_noise: float = 0.1, dtype_obs=np.float64):
    """Returns a dict with observations (B x T, raw ranges as a lidar reports them: distance to
    the trunk SURFACE, no-return 10.0), odometry (3 x T), velocities (2 x T), x_true, x_init,
    landmarks_true (2 x L_true), map_init (2 x L_true) and the generator parameters."""
    G = int(round(np.sqrt(L_true)))
    if G * G != L_true:
        raise ValueError("L_true must be a perfect square (G x G landmark grid)")
    if T is None:
        T = 10 * L_true
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(G), np.arange(G), indexing="xy")
    lm = np.stack([PITCH * gx.ravel() + rng.uniform(-JITTER, JITTER, L_true),
                   PITCH * gy.ravel() + rng.uniform(-JITTER, JITTER, L_true)])
    v, w = _controls(T, G)
    start = np.array([0.0, PITCH / 2.0, 0.0])  # first row midway between landmark rows 0 and 1
    x_true = _integrate(v, w, start)
    scans = _raycast_chunked(x_true, lm, G, beams)
    hit = scans < RANGE_MAX
    scans[hit] += rng.normal(0.0, range_noise, int(hit.sum()))
    scans = np.minimum(scans, RANGE_MAX)
    scans = np.maximum(scans, 0.05)
    scans = scans.astype(np.float32).astype(dtype_obs)  # float32-exact like the real logs
    vn = v + rng.normal(0.0, ctrl_noise[0], T)
    wn = w + rng.normal(0.0, ctrl_noise[1], T)
    odo = _integrate(vn, wn, start)
    x_init = x_true.copy()
    x_init[0:2] += rng.normal(0.0, init_pose_noise[0], (2, T))
    x_init[2] += rng.normal(0.0, init_pose_noise[1], T)
    x_init[:, 0] = odo[:, 0]  # the reference pins pose 0 to odometry[:,0] (sensors.py:61)
    map_init = lm + rng.normal(0.0, init_map_noise, lm.shape)
    return dict(observations=np.ascontiguousarray(scans.T), odometry=odo, velocities=np.stack([vn, wn]),
                x_true=x_true, x_init=x_init, landmarks_true=lm, map_init=map_init,
                params=dict(L_true=L_true, T=T, seed=seed, beams=beams, pitch=PITCH, jitter=JITTER,
                            row_spacing=ROW_SPACING, speed=SPEED, dt=DT, trunk_radius=TRUNK_RADIUS,
                            range_noise=range_noise, ctrl_noise=list(ctrl_noise),
                            init_pose_noise=list(init_pose_noise), init_map_noise=init_map_noise))


def make_synthetic_loop(L_true: int = 16, T: int = 2048, seed: int = 20181, beams: int = 181, range_noise: float = 0.02,
                        ctrl_noise=(0.02, 0.01), init_pose_noise=(0.05, 0.01), init_map_noise: float = 0.1):
    """One trajectory of the BATCH workload (BASELINE configs[4], SURVEY.md 8d "C5": T = 2048 poses, 16 landmarks, per-trajectory
    seed): the same sensor and noise model as make_synthetic, but the robot CIRCLES through a small G x G field (pitch 4 m,
    G = sqrt(L_true)) -- radius 0.3 of the field's width around its centre -- so that every one of the T scans sees landmarks."""
    G = int(round(np.sqrt(L_true)))
    if G * G != L_true:
        raise ValueError("L_true must be a perfect square (G x G landmark grid)")
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(G), np.arange(G), indexing="xy")
    lm = np.stack([PITCH * gx.ravel() + rng.uniform(-JITTER, JITTER, L_true),
                   PITCH * gy.ravel() + rng.uniform(-JITTER, JITTER, L_true)])
    width = PITCH * (G - 1)
    radius = max(0.3 * width, 2.0)
    centre = np.array([width / 2.0, width / 2.0])
    n_loop = int(round(2 * np.pi * radius / (SPEED * DT)))
    w_true = 2 * np.pi / (n_loop * DT)                      # an exact loop in n_loop Euler steps
    v = np.full(T, radius * w_true)
    w = np.full(T, w_true)
    start = np.array([centre[0] + radius, centre[1], np.pi / 2.0])
    x_true = _integrate(v, w, start)
    scans = _raycast_chunked(x_true, lm, G, beams)
    hit = scans < RANGE_MAX
    scans[hit] += rng.normal(0.0, range_noise, int(hit.sum()))
    scans = np.maximum(np.minimum(scans, RANGE_MAX), 0.05).astype(np.float32).astype(np.float64)
    vn = v + rng.normal(0.0, ctrl_noise[0], T)
    wn = w + rng.normal(0.0, ctrl_noise[1], T)
    odo = _integrate(vn, wn, start)
    x_init = x_true.copy()
    x_init[0:2] += rng.normal(0.0, init_pose_noise[0], (2, T))
    x_init[2] += rng.normal(0.0, init_pose_noise[1], T)
    x_init[:, 0] = odo[:, 0]
    map_init = lm + rng.normal(0.0, init_map_noise, lm.shape)
    return dict(observations=np.ascontiguousarray(scans.T), odometry=odo, velocities=np.stack([vn, wn]), x_true=x_true, x_init=x_init,
                landmarks_true=lm, map_init=map_init, params=dict(L_true=L_true, T=T, seed=seed, beams=beams, radius=radius))
