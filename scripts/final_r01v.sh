#!/bin/bash
# Final evidence pass of the round on one B200: whole GPU suite, ncu launch list + full captures, bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests_r01v.log 2>&1
echo "gpu tests exit=$?"; tail -n 4 gpurun_out/gpu_tests_r01v.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01v.log 2>&1; echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke_r01v.log
bash scripts/profile_round.sh r01v > gpurun_out/profile_r01v.log 2>&1; tail -n 3 gpurun_out/profile_r01v.log
python bench.py > gpurun_out/bench_r01v.log 2>&1; tail -n 1 gpurun_out/bench_r01v.log | cut -c1-300
python bench.py --impl reference > gpurun_out/bench_ref_r01v.log 2>&1; tail -n 1 gpurun_out/bench_ref_r01v.log | cut -c1-200
python bench.py --workload model --steps 10 --warmup 3 > gpurun_out/bench_model_r01v.log 2>&1; tail -n 1 gpurun_out/bench_model_r01v.log | cut -c1-300
python bench.py --quick --anchors 16 --steps 10 --warmup 3 > gpurun_out/bench_a16_r01v.log 2>&1; tail -n 1 gpurun_out/bench_a16_r01v.log | cut -c1-200
python bench.py --quick --dropout 0.1 --steps 20 --warmup 5 > gpurun_out/bench_drop_r01v.log 2>&1; tail -n 1 gpurun_out/bench_drop_r01v.log | cut -c1-200
