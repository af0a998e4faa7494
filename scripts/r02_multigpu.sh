#!/bin/bash
# multi-GPU lines of the round: bash scripts/r02_multigpu.sh <N> <tag>
N=${1:-2}; tag=${2:-r02h}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-gpu-baseline > $O/bench_n${N}_$tag.log 2>&1; echo "fusion4 N=$N rc=$?"; grep "^{" $O/bench_n${N}_$tag.log | cut -c1-250
DSF_DEFER_REDUCE=0 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --quick > $O/bench_n${N}_nodefer_$tag.log 2>&1; echo "fusion4 N=$N no-defer: $(grep '^{' $O/bench_n${N}_nodefer_$tag.log)"
timeout 900 $TR bench.py --gpus $N --workload stage4 --steps 20 --warmup 5 --quick > $O/bench_n${N}_stage4_$tag.log 2>&1; echo "stage4 N=$N: $(grep '^{' $O/bench_n${N}_stage4_$tag.log)"
timeout 900 $TR bench.py --gpus $N --workload model --steps 10 --warmup 3 > $O/bench_n${N}_model_$tag.log 2>&1; echo "model N=$N rc=$?"; grep "^{" $O/bench_n${N}_model_$tag.log | cut -c1-250
timeout 900 $TR bench.py --gpus $N --workload stage4 --anchors 16 --steps 10 --warmup 3 --no-gpu-baseline --sustained 0 > $O/bench_n${N}_a16_$tag.log 2>&1; echo "a16 N=$N rc=$?"; grep "^{" $O/bench_n${N}_a16_$tag.log | cut -c1-250
