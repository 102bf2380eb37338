#!/bin/bash
# gpurun with retries while the pod is busy (exit 3 / "transient"): tools/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8; do
    /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
    if ! grep -q "status=transient\|already running" "$log"; then break; fi
    sleep 90
done
