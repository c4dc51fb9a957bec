#!/bin/bash
# Persistent-kernel hand-shake stress (round 2): work-item mixes that break item-level barrier protocols - hundreds of
# one-tile items per CTA, query blocks with no visible key (causal, Nq > Nk), half-empty query blocks - on watchdog
# builds (-DFCSA_WATCHDOG: a thread stuck on an mbarrier for > 2e9 cycles prints the barrier and traps).
# Build first:  cd tests/cuda && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DFCSA_WATCHDOG -o wd_fwd trace_fwd.cu -lcuda
#               (same for wd_bwd from trace_bwd.cu).   Usage: tests/gpu_stress_persistent.sh [reps]
reps=${1:-2}
for rep in $(seq 1 $reps); do
for shape in "1 4000 16 16 1" "1 4000 300 300 1" "1 3000 130 130 0" "2 600 520 260 1" "1 2000 384 384 1" "3 500 16 1024 1" "1 1500 700 300 1" "4 8 4096 4096 1"; do
for v in wd_fwd wd_bwd; do
  out=$(timeout -k 5 20 ./tests/cuda/$v $shape 2>&1 | sed "s/thread [0-9]* /thread X /; s/block [0-9]* /block Y /" | sort | uniq -c | sort -rn | grep -v "barrier 0 at" | head -4 | tr '\n' ';')
  echo "$v [$shape] $out"
done; done; done
