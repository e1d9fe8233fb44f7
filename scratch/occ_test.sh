#!/bin/bash
echo "== 1 block/SM (8 warps)"; ICMSLAM_SMEM_PAD=60000 timeout 200 python scratch/prof_run.py 2>&1 | grep -E "mean cycles|kernel ms"
echo "== tile 32: 4 blocks"; ICMSLAM_TILE=32 timeout 200 python scratch/prof_run.py 2>&1 | grep -E "mean cycles|kernel ms"
echo "== tile 32: 3 blocks"; ICMSLAM_TILE=32 ICMSLAM_SMEM_PAD=16000 timeout 200 python scratch/prof_run.py 2>&1 | grep -E "mean cycles|kernel ms"
echo "== tile 32: 2 blocks"; ICMSLAM_TILE=32 ICMSLAM_SMEM_PAD=50000 timeout 200 python scratch/prof_run.py 2>&1 | grep -E "mean cycles|kernel ms"
echo "== tile 32: 1 block"; ICMSLAM_TILE=32 ICMSLAM_SMEM_PAD=100000 timeout 200 python scratch/prof_run.py 2>&1 | grep -E "mean cycles|kernel ms"
