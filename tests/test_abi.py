"""The C-ABI library loads without a GPU and exports every symbol include/icmslam.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "icmslam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icmslam_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from icm_slam_b200 import _lib
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names, set(names) ^ set(_lib.EXPORTS)
    assert lib.icmslam_abi_version() == 1


def test_no_cpu_fallback_create_fails_without_device():
    import torch
    if torch.cuda.is_available():
        return
    from icm_slam_b200 import _lib
    from icm_slam_b200.config import ConfigICM
    from icm_slam_b200.engine import Engine
    import pytest
    with pytest.raises(_lib.IcmSlamError):
        Engine(ConfigICM.from_values())


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "icm_slam_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            txt = open(os.path.join(pkg, f)).read()
            assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", ""), f
