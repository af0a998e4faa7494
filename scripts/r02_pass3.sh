#!/bin/bash
tag=${1:-r02c}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_optim.py tests/test_gpu_kernels.py -q -m gpu -k "optim or chain or tokens or upsample" > $O/gpu_tests_a_$tag.log 2>&1; echo "kernel tests rc=$?"; tail -4 $O/gpu_tests_a_$tag.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stage.py -q -m gpu > $O/gpu_tests_b_$tag.log 2>&1; echo "stage tests rc=$?"; tail -4 $O/gpu_tests_b_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-gpu-baseline > $O/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -c 300 $O/bench_$tag.log
for st in 1 2; do for v in "DSF_CHAIN=0" "DSF_CHAIN_BWD=0" "DSF_CHAIN=1"; do env $v timeout 300 python bench.py --stage $st --quick --steps 30 --warmup 5 > $O/chain_${tag}_s${st}_$v.log 2>&1; echo "stage $st $v: $(tail -n 1 $O/chain_${tag}_s${st}_$v.log)"; done; done
timeout 600 python bench.py --steps 20 --warmup 5 --optimizer --no-gpu-baseline --sustained 0 > $O/bench_opt_$tag.log 2>&1; echo "bench --optimizer rc=$?"; tail -c 200 $O/bench_opt_$tag.log
