#!/bin/bash
# Runs the GPU test-suite in separate bounded processes so that one faulting kernel cannot take the
# others' results with it.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  timeout $to python -m pytest "$@" -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
run simple 600 tests/test_gpu_kernels.py -m gpu -k "tokens or layernorm or upsample or gemm_f32 or softmax or error"
run gemm_nt 300 tests/test_gpu_kernels.py -m gpu -k "gemm_bf16_nt"
run gemm_tn 300 tests/test_gpu_kernels.py -m gpu -k "gemm_bf16_tn"
run attn 600 tests/test_gpu_kernels.py -m gpu -k "attention"
run stage_f32 900 tests/test_gpu_stage.py -m gpu -k "fp32 or dropin or float32"
run stage_bf16 900 tests/test_gpu_stage.py -m gpu -k "bf16 or bfloat16 or scaled"
cat gpurun_out/summary.txt
