#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 40 --warmup 12 --no-cpu-baseline --kernel-times 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.4f fused-kernel %.4f e2e %.1f'%(d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['value']))"; }
run ICMSLAM_CERT=0
run ICMSLAM_CERT=1
ICMSLAM_CERT=0 WARM=14 timeout 300 python scratch/prof_run.py 2>&1 | grep -E "prof|kernel ms" | grep -v "e+\|barrier\|solve"
ICMSLAM_CERT=1 WARM=14 timeout 300 python scratch/prof_run.py 2>&1 | grep -E "prof|kernel ms" | grep -v "e+\|barrier\|solve"
