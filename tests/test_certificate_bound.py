"""CPU check of the inequality behind the label certificates (icm_slam_b200/csrc/fused.cuh, "Label certificates"; experimental,
ICMSLAM_CERT=1): if an observation was accepted at distance d inside the hint radius r of its landmark, and since then the
pose moved by (dx, dy, dtheta) and the landmark record by at most g (position in the 1-norm plus loss of radius), then

    (|dx| + |dy| + |dtheta| rho + g) (1 + 1e-6) + 1e-9 < slack,   slack = floor8((r^2 - d^2) / (2 sqrt(thr2_hi)))

implies that the hint test d'^2 <= r'^2 still passes -- so the label is provably unchanged.  Randomised: the implication must
hold in every trial, with the slack quantised exactly as the kernel quantises it (16 then 8 bits, rounded down)."""
import numpy as np


def test_certificate_implies_hint_acceptance():
    rng = np.random.default_rng(20181)
    dist_thr, rmax = 1.0, 10.0
    thr2_hi = dist_thr * dist_thr
    rho = rmax * (1.0 + 1e-9)
    marg_scale = (1.0 - 1e-9) * 65535.0 / (2.0 * np.sqrt(thr2_hi) * dist_thr)
    slack_unit = (1.0 - 1e-9) * dist_thr / 256.0
    trials = certified = 0
    for _ in range(20000):
        r = rng.uniform(0.3, 1.0) * np.sqrt(thr2_hi)                 # proven-nearest radius of the landmark (<= sqrt(thr2_hi))
        lm = rng.normal(size=2) * 30.0
        # a beam (body frame, |b| <= rmax) whose projection lands at distance d < r from the landmark
        d = r * np.sqrt(rng.uniform(0.0, 1.0))
        ang = rng.uniform(0, 2 * np.pi)
        w = lm + d * np.array([np.cos(ang), np.sin(ang)])
        rb, ab = rng.uniform(0.5, rmax), rng.uniform(0, np.pi)
        b = rb * np.array([np.cos(ab), np.sin(ab)])
        heading = rng.uniform(-7, 7)
        th = heading - np.pi / 2
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        pose = np.array([*(w - R @ b), heading])
        d2 = float(np.sum((w - lm) ** 2))
        q16 = min(int(np.floor((r * r - d2) * marg_scale)), 65535)
        if q16 < 0:
            continue
        slack = np.float32((q16 >> 8) * slack_unit)                   # what the run record stores
        # drift since the stamp
        scale = rng.choice([1e-4, 1e-2, 0.3])
        dpose = rng.normal(size=3) * scale * [1.0, 1.0, 0.05]
        dlm = rng.normal(size=2) * scale
        dr = abs(rng.normal()) * scale * 0.5                           # loss of radius
        g = abs(dlm[0]) + abs(dlm[1]) + dr
        drift = (abs(dpose[0]) + abs(dpose[1]) + abs(dpose[2]) * rho + g) * (1.0 + 1e-6) + 1e-9
        trials += 1
        if not drift < float(slack):
            continue
        certified += 1
        p2 = pose + dpose
        th2 = p2[2] - np.pi / 2
        R2 = np.array([[np.cos(th2), -np.sin(th2)], [np.sin(th2), np.cos(th2)]])
        w2 = R2 @ b + p2[:2]
        lm2, r2 = lm + dlm, r - dr
        assert r2 > 0 and float(np.sum((w2 - lm2) ** 2)) <= r2 * r2, (d, r, drift, slack)
    assert trials > 10000 and certified > 2000, (trials, certified)
