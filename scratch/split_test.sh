#!/bin/bash
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --kernel-times 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.4f fused-kernel %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms']))"; }
ICMSLAM_SPLIT=1 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run ICMSLAM_SPLIT=0
run ICMSLAM_SPLIT=1 ICMSLAM_SOLVE_OCC=1024
run ICMSLAM_SPLIT=1 ICMSLAM_SOLVE_OCC=768
run ICMSLAM_SPLIT=1 ICMSLAM_SOLVE_OCC=512
run ICMSLAM_SPLIT=1 ICMSLAM_TILE=32
