#!/bin/bash
N=2; O=gpurun_out; tag=${1:-r02k}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --workload stage4 --steps 10 --warmup 3 --quick > $O/sync_${tag}_stage4.log 2>&1; echo "stage4: $(grep '^{' $O/sync_${tag}_stage4.log)"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --quick > $O/sync_${tag}_fusion4.log 2>&1; echo "fusion4: $(grep '^{' $O/sync_${tag}_fusion4.log)"
timeout 900 $TR bench.py --gpus $N --workload model --steps 10 --warmup 3 > $O/sync_${tag}_model.log 2>&1; echo "model: $(grep '^{' $O/sync_${tag}_model.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["config"]["params_identical_across_ranks"], d["config"]["params_not_identical"])')"
