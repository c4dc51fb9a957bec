#!/bin/bash
# One gpurun call of round 2: parity suite, smoke, bench, host overhead, A/B harnesses.  Usage: tests/gpu_round2.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
timeout -k 10 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout -k 10 120 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
timeout -k 10 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout -k 10 120 python tests/gpu_cpu_overhead.py > $out/${tag}_overhead.log 2>&1
timeout -k 10 120 ./tests/cuda/time_bwd > $out/${tag}_time_bwd.log 2>&1
for v in tests/cuda/time_bwd_*; do [ -x "$v" ] && { echo "variant $v" >> $out/${tag}_time_bwd.log; timeout -k 10 120 $v >> $out/${tag}_time_bwd.log 2>&1; }; done
timeout -k 10 120 ./tests/cuda/time_fwd > $out/${tag}_time_fwd.log 2>&1
timeout -k 10 120 ./tests/cuda/bench_aux > $out/${tag}_aux.log 2>&1
if [ -n "$2" ]; then timeout -k 10 1500 python tests/gpu_reference_scripts.py --out $out --only $2 > $out/${tag}_refscripts.log 2>&1; cat $out/${tag}_refscripts.log; fi
tail -25 $out/${tag}_pytest.log
cat $out/${tag}_smoke.log | tail -2
cat $out/${tag}_bench.json | cut -c1-1500
tail -3 $out/${tag}_bench.err
cat $out/${tag}_overhead.log $out/${tag}_time_bwd.log $out/${tag}_time_fwd.log $out/${tag}_aux.log
