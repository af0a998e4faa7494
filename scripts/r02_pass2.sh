#!/bin/bash
# Round-2 GPU pass 2: whole GPU test-suite after the prune / new kernels, bench line, stage-1 HBM kernels under ncu.
tag=${1:-r02b}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/gpu_tests_$tag.log 2>&1; echo "gpu tests rc=$?"; tail -5 $O/gpu_tests_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -c 400 $O/bench_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 --optimizer --no-gpu-baseline --sustained 0 > $O/bench_opt_$tag.log 2>&1; echo "bench --optimizer rc=$?"; tail -c 300 $O/bench_opt_$tag.log
DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage1.csv python bench.py --stage 1 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s1.log 2>&1
ls -la $O/*$tag*
