#!/bin/bash
# ncu evidence for one round (run on the GPU box through gpurun, 1 GPU):  scripts/profile_round.sh <tag>
#   1. launch list of one eager step (every kernel with its device time; cold-cache, serialised -> compare SHARES)
#   2. --set full captures of the hot kernels, forward and backward, taken from the same step
# Outputs land in gpurun_out/; summarise them into profiles/ with scripts/ncu_summary.py.
tag=${1:-rXX}
mkdir -p gpurun_out
CMD="python bench.py --no-graph --quick --steps 1 --warmup 3"
$CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
DSF_NCU_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
DSF_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_nt3|attn_fwd3|layernorm_fwd" -c 7 -o gpurun_out/prof_fwd_$tag $CMD > gpurun_out/ncu_fwd_$tag.log 2>&1
DSF_NCU_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"attn_bwd|layernorm_bwd|gemm_tn2|colsum" -c 9 -o gpurun_out/prof_bwd_$tag $CMD > gpurun_out/ncu_bwd_$tag.log 2>&1
ls -la gpurun_out/*$tag*
