#!/bin/bash
# run ON THE GPU BOX: each configuration in its own process
cd tools/probe
for cfg in "2 96 38 -16 -3 1" "3 96 38 368 -3 1" "3 96 38 -16 445 1" "3 96 38 624 477 1" "3 96 38 48 29 0" "3 80 38 -16 -3 1"; do
  timeout 60 ./tma_probe $cfg; echo "   exit $?"
done
