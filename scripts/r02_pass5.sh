#!/bin/bash
tag=${1:-r02g}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_optim.py tests/test_gpu_parity.py tests/test_gpu_stage.py tests/test_gpu_kernels.py -q -m gpu -k "optim or chain or parity or stage" > $O/gpu_tests_$tag.log 2>&1; echo "tests rc=$?"; tail -4 $O/gpu_tests_$tag.log
for st in 1 2; do for v in 0 1 2; do DSF_CHAIN_BWD=$v timeout 300 python bench.py --stage $st --quick --steps 30 --warmup 5 > $O/chainbwd_${tag}_s${st}_$v.log 2>&1; echo "stage $st DSF_CHAIN_BWD=$v: $(tail -n 1 $O/chainbwd_${tag}_s${st}_$v.log)"; done; done
DSF_CHAIN_BWD=2 DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage1_chainbwd.csv python bench.py --stage 1 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s1.log 2>&1
DSF_CHAIN_BWD=2 DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage2_chainbwd.csv python bench.py --stage 2 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s2.log 2>&1
