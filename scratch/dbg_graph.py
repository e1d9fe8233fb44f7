import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_parity import _synthetic_case, _cfg, _engine
from icm_slam_b200.multigpu import SegmentedSolver
d, cfgd = _synthetic_case(625, 4000, 20181 + 15)
z, odo, u = d["observations"], d["odometry"], d["velocities"]
cfg = _cfg(**cfgd)
refs = {}
single = _engine(cfg, z, odo, u)
single.set_map(d["map_init"]); single.set_poses(d["x_init"])
for k in range(1, 10):
    single.iterate(None, odo[:, 0], 1)
    refs[k] = single.get_poses()
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    for use_graph in (False, True):
        sol = SegmentedSolver(cfg, 0, 1, device=0)
        sol.engine.set_stream(stream.cuda_stream)
        sol.load(z, odo, u, precondition=True)
        sol.set_map(d["map_init"]); sol.set_poses(d["x_init"])
        tot = 0
        for n in (4, 2, 1, 2):
            sol.sweep(n, use_graph=use_graph); tot += n
            stream.synchronize()
            x = sol.owned_poses()
            best = min(range(1, 10), key=lambda k: np.abs(x - refs[k]).max())
            print("graph" if use_graph else "eager", "after", tot, "sweeps: closest to ref sweep", best, "maxdiff", np.abs(x - refs[best]).max(), "graph?", sol._graph is not None)
        sol.close()
