#!/bin/bash
# Builds the library here, then submits a session script to a GPU box, retrying while the pod has
# no free slot (nothing is charged for those attempts).  usage: scripts/gpu_submit.sh <script> <tag> [gpurun args]
SCRIPT=$1; TAG=$2; shift 2
python -m xtrack_b200.build > /dev/null || exit 1
rm -rf /tmp/build_$TAG; cp -r xtrack_b200/csrc/_build /tmp/build_$TAG      # objects matching the profiles
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout 2400 "$@" -- "bash $SCRIPT $TAG" > /tmp/gpu_$TAG.log 2>&1
  if grep -q "status=transient" /tmp/gpu_$TAG.log; then sleep 150; continue; fi
  break
done
tail -40 /tmp/gpu_$TAG.log
