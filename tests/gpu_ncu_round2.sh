#!/bin/bash
# ncu captures of every kernel of a training step (round 2 evidence).  Usage: tests/gpu_ncu_round2.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:l2norm_fwd_pair -s 2 -c 1 -f -o $out/${tag}_ncu_l2norm ./tests/cuda/bench_aux > $out/${tag}_ncu_l2norm.log 2>&1
timeout -k 10 300 $NCU -k regex:bwd_prep -s 2 -c 1 -f -o $out/${tag}_ncu_prep ./tests/cuda/bench_aux > $out/${tag}_ncu_prep.log 2>&1
timeout -k 10 300 $NCU -k regex:dq_finish64 -s 14 -c 1 -f -o $out/${tag}_ncu_finish ./tests/cuda/bench_aux > $out/${tag}_ncu_finish.log 2>&1
timeout -k 10 300 $NCU -k regex:fcsa_bwd_kernel -s 2 -c 1 -f -o $out/${tag}_ncu_bwd ./tests/cuda/time_bwd > $out/${tag}_ncu_bwd.log 2>&1
timeout -k 10 300 $NCU -k regex:fcsa_fwd_kernel -s 2 -c 1 -f -o $out/${tag}_ncu_fwd ./tests/cuda/time_fwd > $out/${tag}_ncu_fwd.log 2>&1
# launch list of one bench step sequence (shares, not absolutes)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 > $out/${tag}_launches_bench.log 2>&1
ls -la $out/${tag}_ncu_*.ncu-rep
tail -2 $out/${tag}_ncu_*.log
