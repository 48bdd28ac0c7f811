#!/bin/bash
# usage: gpurun_retry.sh <timeout> <script>
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $1 -- "bash $2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
