#!/bin/bash
# one-GPU evidence run: full GPU test suite, then ncu launch list and --set full capture of the default path on C4
timeout 1500 python -m pytest tests -m gpu -q --timeout=1200 2>&1 | tail -6
python dev/prof_c4.py 6 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_c4.csv python dev/prof_c4.py 6 > gpurun_out/ncu_launch2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_runs|k_solve_tile|k_assoc_tiles|k_tail_steady|k_tail_labels' --launch-skip 10 --launch-count 12 -o gpurun_out/r02f_prof -f python dev/prof_c4.py 6 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/r02f_prof.ncu-rep; tail -3 gpurun_out/ncu_full2.log
