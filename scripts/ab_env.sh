#!/bin/bash
# Generic A/B on the GPU box:  scripts/ab_env.sh name1:VAR=VAL,VAR2=VAL name2: ...   (empty list after ':' = defaults)
# Each configuration runs `bench.py --quick` twice, interleaved; logs land in gpurun_out/ab_<name>_<k>.log.
mkdir -p gpurun_out
for k in 1 2; do
  for spec in "$@"; do
    name=${spec%%:*}; envs=${spec#*:}
    env $(echo "$envs" | tr ',' ' ') DSF_AB=1 timeout 300 python bench.py --quick --steps 30 --warmup 5 > gpurun_out/ab_${name}_$k.log 2>&1
    echo "$name[$k]: $(tail -n 1 gpurun_out/ab_${name}_$k.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>&1 | tail -n 1)"
  done
done
