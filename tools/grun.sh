#!/bin/bash
# retry wrapper around gpurun for transient "busy" answers: tools/grun.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
