#!/bin/sh
# retry a gpurun call while the pool answers "busy" (exit 3); everything else ends the loop
# usage: sh tools/gpurun_retry.sh <logfile> <gpurun args...>
LOG="$1"; shift
i=0
while [ $i -lt 30 ]; do
	/usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
	rc=$?
	if [ $rc -ne 3 ]; then echo "gpurun exit $rc after $i retries"; exit $rc; fi
	i=$((i + 1)); sleep 90
done
echo "gpurun: still busy after $i tries"; exit 3
