#!/bin/bash
# Round-2 GPU pass 1 (one gpurun call): parity tables, new parity tests, default bench line, per-stage launch lists, ncu metric names.
tag=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$tag.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py::test_top1_beam_agreement_over_64_samples -x -q -s -m gpu > $O/parity_tests_$tag.log 2>&1; echo "parity tests rc=$?"; tail -3 $O/parity_tests_$tag.log
for sh in "12 64 64 8" "12 128 32 8" "12 256 16 8" "12 512 8 8"; do
  timeout 600 python tests/tools/diag_bf16_error.py $sh --summary > "$O/errtab_${tag}_$(echo $sh | tr ' ' '_').txt" 2>&1
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -c 600 $O/bench_$tag.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref_$tag.log 2>&1; echo "ref rc=$?"
for st in 1 2 3; do
  for mb in 0 1; do
    DSF_FWD_MICROBATCH=$mb timeout 300 python bench.py --stage $st --quick --steps 30 --warmup 5 > $O/mb_${tag}_s${st}_$mb.log 2>&1
    echo "stage $st microbatch=$mb: $(tail -n 1 $O/mb_${tag}_s${st}_$mb.log)"
  done
done
ncu --query-metrics 2>/dev/null | grep -i "tensor" > $O/ncu_tensor_metrics_$tag.txt
for st in 1 2 3 4; do
  DSF_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_${tag}_stage$st.csv python bench.py --stage $st --no-graph --quick --steps 1 --warmup 3 > $O/ncu_launches_${tag}_s$st.log 2>&1
done
for pat in "tokens_|upsample_:4" "layernorm_fwd:1" "layernorm_bwd:2"; do
  k=${pat%%:*}; c=${pat##*:}; n=$(echo $k | tr -d '_|')
  DSF_NCU_RANGE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"$k" -c $c -o $O/prof_hbm_stage1_${n}_$tag python bench.py --stage 1 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_hbm_${n}_$tag.log 2>&1
done
ls -la $O/*$tag* | head -40
