#!/bin/bash
# Final single-GPU pass of the round: smoke, whole GPU test-suite, driver-style bench (both arms), ncu evidence.
tag=${1:-r02z}
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"; cat $O/smoke_$tag.log | tail -2
timeout 1800 python -m pytest tests -q -m gpu > $O/gpu_tests_$tag.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/gpu_tests_$tag.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -c 200 $O/bench_$tag.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref_$tag.log 2>&1; echo "reference arm rc=$?"; tail -c 300 $O/bench_ref_$tag.log
timeout 600 python bench.py --workload stage4 --steps 20 --warmup 5 > $O/bench_stage4_$tag.log 2>&1; echo "stage4 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --optimizer --no-gpu-baseline --sustained 0 > $O/bench_opt_$tag.log 2>&1; echo "bench --optimizer rc=$?"
timeout 900 python bench.py --workload model --steps 10 --warmup 3 > $O/bench_model_$tag.log 2>&1; echo "model rc=$?"
timeout 600 python bench.py --workload stage4 --anchors 16 --steps 10 --warmup 3 --no-gpu-baseline --sustained 0 > $O/bench_a16_$tag.log 2>&1; echo "a16 rc=$?"
timeout 200 python scripts/bench_attn_parts.py final > $O/attn_parts_$tag.log 2>&1; tail -5 $O/attn_parts_$tag.log
timeout 200 python scripts/bench_hbm_kernels.py > $O/hbm_kernels_$tag.log 2>&1; tail -4 $O/hbm_kernels_$tag.log
bash scripts/r02_profile_final.sh $tag
DSF_NCU_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tokens_|upsample_|chain_" -c 6 \
   -o $O/prof_stage1_$tag python bench.py --stage 1 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_stage1_$tag.log 2>&1
ls $O/*$tag* | head -30
