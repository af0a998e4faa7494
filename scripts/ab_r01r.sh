#!/bin/bash
# A/B of the round's late changes on the GPU box: LayerNorm v2 (cp.async row rings), N-split tail of the pair GEMM,
# bf16 dL/d(LN out).  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests_r01r.log 2>&1
echo "gpu tests (defaults) exit=$?"; tail -n 5 gpurun_out/gpu_tests_r01r.log
DSF_LN_DY_BF16=1 timeout 600 python -m pytest tests/test_gpu_stage.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests_dybf16.log 2>&1
echo "stage/model tests with DSF_LN_DY_BF16=1 exit=$?"; tail -n 15 gpurun_out/gpu_tests_dybf16.log
b() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 30 --warmup 5 > gpurun_out/ab_$name.log 2>&1
  echo "$name: $(tail -n 1 gpurun_out/ab_$name.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>&1 | tail -n 1)"
}
b base DSF_LN_IMPL=1 DSF_GEMM_TAIL_NSPLIT=0
b ln2 DSF_GEMM_TAIL_NSPLIT=0
b nsplit DSF_LN_IMPL=1
b both DSF_X=0
b both_dybf16 DSF_LN_DY_BF16=1
b base2 DSF_LN_IMPL=1 DSF_GEMM_TAIL_NSPLIT=0
(DSF_LN_IMPL=1 python scripts/bench_kernels.py ln; python scripts/bench_kernels.py ln; DSF_GEMM_TAIL_NSPLIT=0 python scripts/bench_kernels.py gemm; python scripts/bench_kernels.py gemm) > gpurun_out/ab_kernels.log 2>&1
cat gpurun_out/ab_kernels.log | grep -v Warning | tail -n 40
