#!/bin/bash
# usage: [GPUS=N] gpuretry.sh <timeout> <logfile> <command...>   retries while the pod answers "transient / busy" (exit 3)
T=$1; LOG=$2; shift 2
G=""; if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
