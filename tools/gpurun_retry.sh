#!/bin/bash
# gpurun with retries while the pod answers "busy" (nothing is charged for those)
for attempt in 1 2 3 4 5 6 7 8; do
	out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
	echo "$out" > gpurun_out/.last_retry.log
	echo "$out" | tail -${GPURUN_TAIL:-40}
	if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
	sleep 120
done
