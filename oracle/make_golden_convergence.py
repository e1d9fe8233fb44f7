"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/c1_converged.npz by running the UNMODIFIED reference for config_ros.yaml's
N = 30 sweeps on data_IJAC2018.mat (pass 0 first), the run whose figures the reference's notebook shows.

    python oracle/make_golden_convergence.py        (build container only: needs /root/reference; ~12 minutes, 1 core)

Stored: the pass-0 poses / map the 30 sweeps start from, the reference's poses and map after 30 sweeps, its calc_cambio history
and the joint energy of its final state evaluated with the reference's own fun_x / fun_xn (sum over poses, each against the
landmarks its observations are associated with in the final map).  tests/test_convergence.py holds the fast (red-black /
Newton / previous-map) mode to it: same 11 landmarks within a stated tolerance, joint energy not above the reference's.
"""
import os
import sys
import time

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_runner as rr  # noqa: E402


def main(nsweeps=30):
    _, ref_icm = rr.load_reference()
    z, odo, u = rr.load_ijac()
    cfg = rr.make_config()
    s = rr.make_solver(cfg, rr.precondition(z, cfg), odo, u)
    t0 = time.time()
    p0 = rr.pass0(s)
    print("pass 0: %.1f s, %d landmarks" % (time.time() - t0, p0["mapa"].shape[1]))
    mapa, x = p0["mapa"].copy(), np.ascontiguousarray(p0["x"].copy())
    cambios = []
    for k in range(nsweeps):
        t0 = time.time()
        mapa_new, x, _ = rr.sweep(s, mapa, x)
        cambios.append(ref_icm.calc_cambio(mapa_new, mapa))
        mapa = mapa_new
        print("sweep %d: %.1f s, L = %d, cambio = %s" % (k + 1, time.time() - t0, mapa.shape[1], cambios[-1]), flush=True)
    out = os.path.join(ROOT, "tests", "golden", "c1_converged.npz")
    np.savez_compressed(out, p0_x=p0["x"], p0_map=p0["mapa"], x_ref=x, map_ref=mapa, cambios=np.array(cambios), nsweeps=np.int32(nsweeps),
                        versions=np.array(["numpy=%s" % np.__version__, "scipy=%s" % scipy.__version__]))
    print("wrote", out)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 30)
