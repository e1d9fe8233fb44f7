"""Copies the UNMODIFIED reference (its Python modules, YAML configs and the two .mat logs) from /root/reference/scripts
into the git-ignored baseline/_ref/scripts, so that `gpurun` carries it to the GPU box where /root/reference does not exist.

The reference is not a pip-installable package (no setup.py / pyproject), so "installing" it is this copy.
Nothing under baseline/_ref is committed; bench.py --impl reference and the cpu_baseline leg import it through
oracle/ref_runner.py behind roslibpy / matplotlib stubs, with no file modified.
"""
import os
import shutil
import sys

SRC = "/root/reference/scripts"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "scripts")
FILES = ["ICM_SLAM.py", "ICM_SLAM_tools.py", "funciones_varias.py", "sensors.py", "sensors_definitions.py", "config_ros.yaml", "config_default.yaml", "data_IJAC2018.mat",
         "datos_palomar1.mat", "requisitos.txt"]


def install(verbose=False):
    if not os.path.isdir(SRC):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if os.path.isfile(s) and (not os.path.isfile(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s)):
            shutil.copyfile(s, d)
            if verbose:
                print("copied", f)
    return True


if __name__ == "__main__":
    sys.exit(0 if install(verbose=True) else 1)
