#!/bin/bash
# Sweep of the workspace layout pads (doubles) of large_ws_carve: where w and p sit relative to X and to each other.
out=gpurun_out/$1_pad_sweep.txt; : > $out
run() { echo "== PAD_W=$1 PAD_P=$2" >> $out; CHUB_WS_PAD_W=$1 CHUB_WS_PAD_P=$2 timeout 100 python tools/time_update.py 16384 20 2>&1 | grep time >> $out; }
run 0 0
for p in 32 64 128 256 512 1024 2048 4096 8192; do run 0 $p; done
for w in 32 256 1024 4096; do run $w 0; done
run 0 0
cat $out
