#!/bin/bash
# GPU box: TMA-prefetched residual in the pair GEMM epilogue (fp32 output + residual).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:cacheprovider -k "gemm_bf16_nt" > gpurun_out/gpu_tests_restma.log 2>&1
rc=$?; echo "gemm nt kernel tests exit=$rc"; tail -n 12 gpurun_out/gpu_tests_restma.log
if [ $rc -ne 0 ]; then export DSF_GEMM_RES_TMA=0; echo "RES_TMA FAILED: the rest runs with DSF_GEMM_RES_TMA=0"; fi
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_baseline.py -k "not gemm_bf16_nt" > gpurun_out/gpu_tests_r01u.log 2>&1
echo "gpu tests exit=$?"; tail -n 5 gpurun_out/gpu_tests_r01u.log
b() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 30 --warmup 5 > gpurun_out/ab_$name.log 2>&1
  echo "$name: $(tail -n 1 gpurun_out/ab_$name.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>&1 | tail -n 1)"
}
b u_off DSF_GEMM_RES_TMA=0
if [ $rc -eq 0 ]; then
b u_on DSF_GEMM_RES_TMA=1
b u_off2 DSF_GEMM_RES_TMA=0
b u_on2 DSF_GEMM_RES_TMA=1
b u_on_lnfuse DSF_GEMM_RES_TMA=1 DSF_GEMM_LN_FUSE=1
(DSF_GEMM_RES_TMA=0 python scripts/bench_kernels.py gemm; python scripts/bench_kernels.py gemm) 2>&1 | grep -v Warning | grep "N=512" > gpurun_out/ab_kernels_u.log; cat gpurun_out/ab_kernels_u.log
fi
