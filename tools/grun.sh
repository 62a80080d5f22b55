#!/bin/bash
# gpurun with retries on "transient" / busy answers: tools/grun.sh <logname> [gpurun options] -- '<command>'
# writes gpurun_out/call_<logname>.log; the last line is "done rc=<exit code of gpurun>"
name=$1; shift
log=gpurun_out/call_$name.log
mkdir -p gpurun_out
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" $log || [ $rc -eq 3 ]; then sleep 75; continue; fi
  break
done
echo "done rc=$rc attempts=$attempt" >> $log
