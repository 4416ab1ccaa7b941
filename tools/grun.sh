#!/bin/bash
# usage: tools/grun.sh <command-file> <log-file> [timeout-seconds]
# Runs the command file's content on the GPU box through gpurun, retrying while the pod answers "transient" (nothing charged).
cmd="$(cat "$1")"
log="$2"
to="${3:-900}"
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$cmd" > "$log" 2>&1
  if grep -q "status=transient" "$log"; then sleep 120; else break; fi
done
tail -40 "$log"
