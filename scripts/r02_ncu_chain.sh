#!/bin/bash
tag=${1:-r02e}
O=gpurun_out
mkdir -p $O
for st in 1 2; do
  DSF_NCU_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"chain_" -c 4 \
     -o $O/prof_chain_s${st}_$tag python bench.py --stage $st --no-graph --quick --steps 1 --warmup 3 > $O/ncu_chain_s${st}_$tag.log 2>&1
done
DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage1.csv python bench.py --stage 1 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s1.log 2>&1
DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage2.csv python bench.py --stage 2 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_optim.py -q -m gpu -k "upsample or optim" > $O/gpu_tests_$tag.log 2>&1; tail -3 $O/gpu_tests_$tag.log
ls -la $O/*$tag*
