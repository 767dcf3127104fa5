#!/bin/bash
# retry wrapper around gpurun for multi-GPU calls: tools/grun2.sh <gpus> <timeout> '<command>'
N=$1; T=$2; shift; shift
for i in $(seq 1 40); do
  out=$(gpurun --gpus "$N" --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
