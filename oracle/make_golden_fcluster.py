"""Golden vectors for the single scipy call of pass 0 (ICM_SLAM.py:161):
    fcluster(linkage(pdist(obs)), dist_thr) - 1      (single linkage, criterion='inconsistent', depth=2)
minted from scipy itself in this container (scipy 1.18.1 at the time of minting); run from the repo root:
    python oracle/make_golden_fcluster.py
TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy
from scipy.cluster.hierarchy import fcluster, linkage
from scipy.spatial.distance import pdist

rng = np.random.default_rng(20181)
cases = []
def add(P, t=1.0):
    P = np.ascontiguousarray(P, dtype=np.float64)
    c = fcluster(linkage(pdist(P)), t) - 1
    cases.append((P, float(t), c.astype(np.int32)))

for n in (2, 3, 4, 5, 7, 10, 15, 17, 31, 64, 119, 181):
    for rep in range(6):
        add(rng.uniform(-10, 10, (n, 2)))                                     # scattered
        k = max(1, n // 4)
        ctr = rng.uniform(-8, 8, (k, 2))
        add(ctr[rng.integers(0, k, n)] + rng.normal(0, 0.08, (n, 2)))         # trunks: tight groups
        add(np.round(rng.uniform(-3, 3, (n, 2)) * 2) / 2)                     # lattice: many exact ties and duplicates
for t in (0.5, 0.7, 1.0, 1.15, 2.0):
    for rep in range(5):
        add(rng.uniform(-5, 5, (12, 2)), t)
# a real first scan shape: points along an arc at trunk positions
ang = np.deg2rad(np.array([10, 11, 12, 40, 41, 80, 81, 82, 83, 120, 150, 151]))
add(np.stack([6 * np.cos(ang), 6 * np.sin(ang)], 1) + rng.normal(0, 0.02, (12, 2)))
out = {"n_cases": np.int32(len(cases)), "scipy_version": np.array(scipy.__version__)}
for i, (P, t, c) in enumerate(cases):
    out["P%d" % i] = P; out["t%d" % i] = np.float64(t); out["c%d" % i] = c
np.savez_compressed("tests/golden/fcluster.npz", **out)
print("wrote %d cases (scipy %s)" % (len(cases), scipy.__version__))
