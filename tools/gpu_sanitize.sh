#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over every kernel at small sizes.  usage: gpu_sanitize.sh TAG
TAG=${1:-x}; O=gpurun_out; mkdir -p $O
for TOOL in memcheck racecheck synccheck initcheck; do
  EXTRA=""; [ $TOOL = racecheck ] && EXTRA="--racecheck-report all"
  [ $TOOL = initcheck ] && EXTRA="--track-unused-memory no"
  NPK_SANITIZE_QUICK=1 timeout 900 compute-sanitizer --tool $TOOL $EXTRA --print-limit 40 \
      python tools/sanitizer_workload.py > $O/${TAG}_sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|finished|^ok" $O/${TAG}_sanitize_$TOOL.log | tail -4
done
