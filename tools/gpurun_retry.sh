#!/bin/bash
# usage: tools/gpurun_retry.sh <log file> <gpurun args...>   -- retries while the pod answers "transient" (busy, nothing charged)
log="$1"; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
    /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
    if grep -q "status=transient" "$log" || grep -q "retry in a few minutes" "$log"; then
        sleep 45
        continue
    fi
    break
done
