#!/bin/bash
# retry wrapper around gpurun: exit code 3 = no slot right now (nothing charged) -> wait and retry
# usage: dev/gpu.sh [--gpus N] [--timeout S] -- 'command'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
