"""CPU checks of bench.py's reference arm (the unmodified reference from baseline/_ref or /root/reference, driven through
oracle/ref_runner.py -- the one place besides tests/ and smoke() that may execute oracle/) and of the result writers: the JSON
contract the driver parses, no GPU needed."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "0", "--ref-poses", "200"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["extrapolated"] is True and d["cpu_baseline"]["extrapolated"] is True and "EXTRAPOLATED" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_result_writers_round_trip(tmp_path):
    from icm_slam_b200.offline import save_result, load_result
    rng = np.random.default_rng(7)
    res = dict(x=rng.normal(size=(3, 50)), mapa=rng.normal(size=(2, 7)), cambios=rng.random((4, 3)), mapa_inicial=rng.normal(size=(2, 9)),
               x_inicial=rng.normal(size=(3, 50)), labels=np.arange(20, dtype=np.int32))
    for ext in ("mat", "npz"):
        back = load_result(save_result(str(tmp_path / ("r." + ext)), res))
        for k in ("x", "mapa", "cambios", "mapa_inicial", "x_inicial"):
            assert np.array_equal(np.asarray(back[k]), res[k]), (ext, k)
        assert np.array_equal(np.asarray(back["labels"]).ravel(), res["labels"])
