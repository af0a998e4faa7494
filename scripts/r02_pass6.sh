#!/bin/bash
tag=${1:-r02y}
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -q -m gpu > $O/gpu_tests_$tag.log 2>&1; echo "optim tests rc=$?"; tail -2 $O/gpu_tests_$tag.log
timeout 600 python bench.py --steps 20 --warmup 5 --optimizer --no-gpu-baseline --sustained 0 > $O/bench_opt_$tag.log 2>&1; echo "bench --optimizer rc=$?"; grep "^{" $O/bench_opt_$tag.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])'
timeout 900 python bench.py --workload model --steps 10 --warmup 3 > $O/bench_model_$tag.log 2>&1; echo "model rc=$?"; grep "^{" $O/bench_model_$tag.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])'
