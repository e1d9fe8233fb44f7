"""CPU pin of the two facts the steady tail (csrc/tail.cuh k_tail_steady) rests on, with the kernel's constants:

1. a landmark registered in every cell its (thr1 + margin)-disc overlaps at BUILD time is still found in the cell of any query
   point within thr1 of it after it has moved by at most `margin` (fastgrid.cuh fgrid_make_geom / fgrid_cell_range);
2. nnd0_i - D_i - margin, with nnd0_i = min(distance to the nearest other landmark at build time, 2 thr1) and D the displacements
   since (all <= margin), is a lower bound of the distance from landmark i to every other landmark now -- so the proven radius
   derived from it is valid, and a value above dist_thr rules out a merge.
"""
import numpy as np

FG_MARGIN = 0.05          # fastgrid.cuh


def _geom(pts, dist_thr):
    thr1 = dist_thr * (1.0 + 2.0 ** -20) + FG_MARGIN * dist_thr
    h = 2.0 * thr1 * (1.0 + 2.0 ** -20)
    x0, y0 = pts[0].min(), pts[1].min()
    inv_h = 1.0 / h
    nx = int(np.floor((pts[0].max() - x0) * inv_h)) + 1
    ny = int(np.floor((pts[1].max() - y0) * inv_h)) + 1
    return dict(x0=x0, y0=y0, inv_h=inv_h, delta=thr1 * inv_h, nx=max(nx, 1), ny=max(ny, 1))


def _cell_range(g, x, y):
    fx, fy = (x - g["x0"]) * g["inv_h"], (y - g["y0"]) * g["inv_h"]
    c = lambda v, n: min(max(int(np.floor(v)), 0), n - 1)
    return c(fx - g["delta"], g["nx"]), c(fx + g["delta"], g["nx"]), c(fy - g["delta"], g["ny"]), c(fy + g["delta"], g["ny"])


def _cell(g, x, y):
    c = lambda v, n: min(max(int(np.floor(v)), 0), n - 1)
    return c((y - g["y0"]) * g["inv_h"], g["ny"]) * g["nx"] + c((x - g["x0"]) * g["inv_h"], g["nx"])


def test_margin_registered_cells_stay_supersets_after_a_move_within_the_margin():
    rng = np.random.default_rng(11)
    for dist_thr in (1.0, 0.37):
        margin = FG_MARGIN * dist_thr
        thr1 = dist_thr * (1.0 + 2.0 ** -20)
        pts = rng.uniform(-20.0, 35.0, (2, 400))
        g = _geom(pts, dist_thr)
        assert g["delta"] < 0.5                                     # at most 2 x 2 cells per landmark
        member = {}
        for i in range(pts.shape[1]):
            cx0, cx1, cy0, cy1 = _cell_range(g, pts[0, i], pts[1, i])
            assert cx1 - cx0 <= 1 and cy1 - cy0 <= 1
            for cy in range(cy0, cy1 + 1):
                for cx in range(cx0, cx1 + 1):
                    member.setdefault(cy * g["nx"] + cx, set()).add(i)
        for _ in range(4000):
            i = int(rng.integers(0, pts.shape[1]))
            a = rng.uniform(0, 2 * np.pi)
            moved = pts[:, i] + rng.uniform(0.0, margin) * np.array([np.cos(a), np.sin(a)])       # where the landmark is now
            b = rng.uniform(0, 2 * np.pi)
            q = moved + rng.uniform(0.0, thr1) * np.array([np.cos(b), np.sin(b)])                 # a query point that must find it
            assert i in member.get(_cell(g, q[0], q[1]), set()), (dist_thr, i)                    # (points outside the box clamp to border cells)


def test_displacement_bound_on_the_nearest_other_landmark():
    rng = np.random.default_rng(12)
    dist_thr = 1.0
    margin = FG_MARGIN * dist_thr
    thr1 = dist_thr * (1.0 + 2.0 ** -20)
    for _ in range(60):
        n = int(rng.integers(2, 60))
        p0 = rng.uniform(0.0, 12.0, (2, n))
        d0 = np.hypot(p0[0][:, None] - p0[0][None, :], p0[1][:, None] - p0[1][None, :])
        np.fill_diagonal(d0, np.inf)
        nnd0 = np.minimum(d0.min(axis=1), 2.0 * thr1)               # what k_tail_nn stores (sqrt of min(wide, 4 thr1^2))
        ang = rng.uniform(0, 2 * np.pi, n)
        D = rng.uniform(0.0, margin, n)
        p1 = p0 + D * np.stack([np.cos(ang), np.sin(ang)])
        d1 = np.hypot(p1[0][:, None] - p1[0][None, :], p1[1][:, None] - p1[1][None, :])
        np.fill_diagonal(d1, np.inf)
        lower = nnd0 - D * (1.0 + 1e-12) - margin                   # k_tail_steady
        assert np.all(d1.min(axis=1) >= lower - 1e-12)             # every other landmark is at least `lower` away now
        # no merge can happen where the bound exceeds dist_thr
        safe = lower > dist_thr * (1.0 + 1e-9)
        assert np.all(d1.min(axis=1)[safe] > dist_thr)
