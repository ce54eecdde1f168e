#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   -- retries while the pod answers "transient"/busy
log=$1; shift
for try in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=ok\|status=fail\|status=timeout" "$log"; then exit 0; fi
  if ! grep -q "status=transient\|busy\|retry" "$log"; then exit 0; fi
  sleep 150
done
