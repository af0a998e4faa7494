#!/bin/bash
N=2; O=gpurun_out; tag=${1:-r02i}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for d in 1 0; do
  DSF_DEFER_REDUCE=$d timeout 600 $TR bench.py --gpus $N --workload stage4 --steps 10 --warmup 3 --quick > $O/sync_${tag}_stage4_defer$d.log 2>&1; echo "stage4 defer=$d: $(grep '^{' $O/sync_${tag}_stage4_defer$d.log)"
  DSF_DEFER_REDUCE=$d timeout 600 $TR bench.py --gpus $N --workload stage4 --steps 10 --warmup 3 --quick --no-graph > $O/sync_${tag}_stage4_eager_defer$d.log 2>&1; echo "stage4 eager defer=$d: $(grep '^{' $O/sync_${tag}_stage4_eager_defer$d.log)"
done
DSF_DEFER_REDUCE=0 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --quick > $O/sync_${tag}_fusion4_defer0.log 2>&1; echo "fusion4 defer=0: $(grep '^{' $O/sync_${tag}_fusion4_defer0.log)"
