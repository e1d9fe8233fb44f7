"""The host restatement of scipy's fcluster(linkage(pdist(.)), t) (csrc/fcluster.h, through the C ABI) against golden
vectors minted from scipy (oracle/make_golden_fcluster.py).  No GPU needed."""
import ctypes as C

import numpy as np

from helpers import golden


def _fcluster(P, t):
    from icm_slam_b200 import _lib
    L = _lib.lib()
    px = np.ascontiguousarray(P[:, 0]); py = np.ascontiguousarray(P[:, 1])
    lab = np.empty(P.shape[0], np.int32)
    k = C.c_int32()
    assert L.icmslam_fcluster(C.c_void_p(px.ctypes.data), C.c_void_p(py.ctypes.data), P.shape[0], float(t),
                              C.c_void_p(lab.ctypes.data), C.byref(k)) == 0
    return lab, k.value


def test_fcluster_matches_scipy_goldens():
    g = golden("fcluster.npz")
    n = int(g["n_cases"])
    assert n > 200
    bad = []
    for i in range(n):
        lab, k = _fcluster(g["P%d" % i], float(g["t%d" % i]))
        if not np.array_equal(lab, g["c%d" % i]) or k != int(g["c%d" % i].max()) + 1:
            bad.append(i)
    assert not bad, bad[:20]


def test_fcluster_matches_scipy_live():
    """(and against the scipy of this environment, on fresh random inputs)"""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import pdist
    rng = np.random.default_rng(7)
    for _ in range(200):
        n = int(rng.integers(2, 60))
        P = rng.uniform(-6, 6, (n, 2)) if rng.random() < 0.6 else np.round(rng.uniform(-3, 3, (n, 2)))
        want = fcluster(linkage(pdist(P)), 1.0) - 1
        got, _ = _fcluster(P, 1.0)
        assert np.array_equal(got, want)
