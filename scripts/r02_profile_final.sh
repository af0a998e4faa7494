#!/bin/bash
# Final ncu evidence of the round (one GPU): per-launch time + DRAM bytes + tensor-pipe counters of one eager step of the default
# workload (all four stages), and --set full captures of the top kernels of stage 4.  Summarise with scripts/ncu_summary.py.
tag=${1:-r02z}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --no-graph --quick --steps 1 --warmup 3"
$CMD > $O/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$tag.log; exit 1; }
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,sm__inst_executed_pipe_tensor.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
DSF_NCU_RANGE=1 timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $O/launches_$tag.csv $CMD > $O/ncu_launches_$tag.log 2>&1
echo "launch list rc=$?"
DSF_NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_nt3|attn_fwd3|gemm_tn2" -c 6 -o $O/prof_fwd_$tag python bench.py --stage 4 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_fwd_$tag.log 2>&1
DSF_NCU_RANGE=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"attn_bwd" -c 2 -o $O/prof_bwd_$tag python bench.py --stage 4 --no-graph --quick --steps 1 --warmup 3 > $O/ncu_bwd_$tag.log 2>&1
ls -la $O/*$tag*
