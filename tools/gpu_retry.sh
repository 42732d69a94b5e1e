#!/bin/bash
# usage: tools/gpu_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" (rc 3)
log=$1; shift
for i in $(seq 1 40); do
	/usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
	if [ $rc -ne 3 ]; then break; fi
	sleep 45
done
echo "final rc=$rc" >> "$log"
