#!/bin/bash
# usage: scripts/gpurun_retry.sh LOG TIMEOUT_S 'command' [--gpus N]: retries while the pod answers "busy" (exit 3)
log=$1; to=$2; cmd=$3; shift 3
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" --timeout "$to" -- "$cmd" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
