"""where a host-memory sweep's time goes: raw pinned copies, Engine.sweep, the ICM_SLAM facade"""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icm_slam_b200.config import ConfigICM
from icm_slam_b200.engine import Engine
from icm_slam_b200.icm import ICM_SLAM
from icm_slam_b200.synthetic import make_synthetic
L = 316 * 316; T = 1_000_000
d = make_synthetic(L, T=T, seed=20181 + 4)
cfg = ConfigICM.from_values(N=1, L=2 * L, cota=20.0)
xh = torch.from_numpy(d["x_init"].copy()).pin_memory()
xd = torch.empty_like(xh, device="cuda")
def tm(f, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("pinned H2D 24 MB: %.3f ms   D2H: %.3f ms" % (tm(lambda: xd.copy_(xh, non_blocking=True)), tm(lambda: xh.copy_(xd, non_blocking=True))))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2): xh2.copy_(xd2, non_blocking=True)
xh2 = torch.empty_like(xh).pin_memory(); xd2 = torch.empty_like(xd)
print("H2D + D2H concurrently: %.3f ms" % tm(both))
from icm_slam_b200.icm import Mapa
eng = Engine(cfg)
eng.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
eng.extract()
icm = ICM_SLAM(cfg, x0=d["odometry"][:, 0].copy())
icm.mediciones, icm.odometria, icm.u = d["observations"], d["odometry"], d["velocities"]
icm.mapa_obj = Mapa(cfg)
icm.adopt_engine(eng)
icm.mapa_obj.landmarks_actuales = d["map_init"].shape[1]
x = xh.numpy()
mapa = d["map_init"].copy()
for _ in range(4): mapa, x = icm.iterations_process_offline(mapa, x)
t0 = time.perf_counter()
for _ in range(10): mapa, x = icm.iterations_process_offline(mapa, x)
print("ICM_SLAM.iterations_process_offline: %.3f ms/call" % ((time.perf_counter() - t0) / 10 * 1e3))
e = icm._engine
x0 = np.asarray(icm.x0).reshape(3)
t0 = time.perf_counter()
for _ in range(10): st, Lout, mapa = e.sweep(mapa, x, x0)
print("Engine.sweep: %.3f ms/call" % ((time.perf_counter() - t0) / 10 * 1e3))
for ch in (1, 4, 16):
    pass
t0 = time.perf_counter()
for _ in range(10): icm._sync_data()
print("_sync_data: %.4f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
