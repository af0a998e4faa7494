#!/bin/bash
tag=${1:-r02f}
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -q -m gpu > $O/gpu_tests_$tag.log 2>&1; echo "gpu tests rc=$?"; tail -6 $O/gpu_tests_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -c 300 $O/bench_$tag.log
for st in 1 2; do for v in 0 1 2; do DSF_CHAIN=$v timeout 300 python bench.py --stage $st --quick --steps 30 --warmup 5 > $O/chain_${tag}_s${st}_$v.log 2>&1; echo "stage $st DSF_CHAIN=$v: $(tail -n 1 $O/chain_${tag}_s${st}_$v.log)"; done; done
DSF_CHAIN=2 DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage2_chain.csv python bench.py --stage 2 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s2.log 2>&1
DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage1.csv python bench.py --stage 1 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s1.log 2>&1
timeout 900 python bench.py --workload model --steps 10 --warmup 3 > $O/bench_model_$tag.log 2>&1; echo "model rc=$?"; tail -c 400 $O/bench_model_$tag.log
timeout 900 python bench.py --workload model --torch-optimizer --steps 10 --warmup 3 > $O/bench_model_torchopt_$tag.log 2>&1; echo "model torch-opt rc=$?"; tail -c 200 $O/bench_model_torchopt_$tag.log
timeout 600 python bench.py --workload stage4 --anchors 16 --steps 10 --warmup 3 --no-gpu-baseline --sustained 0 > $O/bench_a16_$tag.log 2>&1; echo "a16 rc=$?"; tail -c 200 $O/bench_a16_$tag.log
timeout 900 python scripts/bench_missing_modality.py > $O/missing_modality_$tag.log 2>&1; cat $O/missing_modality_$tag.log
