"""CPU check of the inequality behind the run records (icm_slam_b200/csrc/runs.cuh): a run whose record passes
|S|^2 <= (n (r - rho))^2 (1 - 1e-9), with S = R sum(b) + n (p - y) and rho the (rounded-up, float-coded) largest distance
of its body-frame points from their centroid, has EVERY observation inside the landmark's proven-nearest radius, i.e. the per-observation hint
test of assoc_tiles.cuh (d2 <= r2) accepts each of them -- randomised, with the kernel's own quantisation of rho and r."""
import numpy as np


def _rho_code(b):
    c = b.sum(axis=0) * (1.0 / b.shape[0])
    rho = np.sqrt(((b - c) ** 2).sum(axis=1).max()) * (1.0 + 1e-6) + 1e-9
    code = min(int(np.ceil(rho * 2048.0)), 65535)
    return np.inf if code >= 65535 else float(np.float32(code) * np.float32(1.0 / 2048.0))


def test_run_record_bound_implies_hint_acceptance():
    rng = np.random.default_rng(20181)
    trials = certified = 0
    for _ in range(20000):
        n = int(rng.integers(1, 9))
        centre = rng.uniform(-9.0, 9.0, 2)
        b = centre + rng.normal(0.0, rng.choice([0.02, 0.1, 0.4]), (n, 2))
        th = rng.uniform(-20.0, 20.0)
        p = rng.uniform(-500.0, 500.0, 2)
        st, ct = np.sin(th - np.pi / 2), np.cos(th - np.pi / 2)
        w = np.stack([b[:, 0] * ct - b[:, 1] * st + p[0], b[:, 0] * st + b[:, 1] * ct + p[1]], axis=1)
        r2 = float(rng.choice([1.0, 0.6, 0.3, 0.05])) * (1.0 - 2.0 ** -30)
        y = w.mean(axis=0) + rng.normal(0.0, rng.choice([0.05, 0.3, 0.8]), 2)
        r = np.nextafter(np.sqrt(r2), 0.0)             # sqrt rounded down (LmRec.r)
        rho = _rho_code(b)
        sb = b.sum(axis=0)
        yx, yy = y[0] - p[0], y[1] - p[1]
        Sx = (ct * sb[0] - st * sb[1]) - n * yx
        Sy = (st * sb[0] + ct * sb[1]) - n * yy
        a = (r - rho) * n
        ok = a > 0.0 and Sx * Sx + Sy * Sy <= a * a * (1.0 - 1e-9)
        trials += 1
        if not ok:
            continue
        certified += 1
        d2 = (w[:, 0] - y[0]) ** 2 + (w[:, 1] - y[1]) ** 2
        assert np.all(d2 <= r2), (d2.max(), r2)
    assert trials >= 20000 and certified > 2000, (trials, certified)
