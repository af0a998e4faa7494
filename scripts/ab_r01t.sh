#!/bin/bash
# GPU box: fused GEMM + LayerNorm epilogue (dsf_gemm_bf16_nt_ln): kernel test first (bounded), then the suite, then A/B.
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -p no:cacheprovider -k "nt_ln_fused" > gpurun_out/gpu_tests_lnfuse.log 2>&1
rc=$?; echo "nt_ln kernel tests exit=$rc"; tail -n 12 gpurun_out/gpu_tests_lnfuse.log
if [ $rc -ne 0 ]; then export DSF_GEMM_LN_FUSE=0; echo "FUSED KERNEL FAILED: the rest runs with DSF_GEMM_LN_FUSE=0"; fi
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_baseline.py -k "not nt_ln_fused" > gpurun_out/gpu_tests_r01t.log 2>&1
echo "gpu tests exit=$?"; tail -n 5 gpurun_out/gpu_tests_r01t.log
b() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 30 --warmup 5 > gpurun_out/ab_$name.log 2>&1
  echo "$name: $(tail -n 1 gpurun_out/ab_$name.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>&1 | tail -n 1)"
}
b t_nofuse DSF_GEMM_LN_FUSE=0
if [ $rc -eq 0 ]; then
b t_fuse DSF_GEMM_LN_FUSE=1
b t_nofuse2 DSF_GEMM_LN_FUSE=0
b t_fuse2 DSF_GEMM_LN_FUSE=1
python scripts/bench_kernels.py lnfuse > gpurun_out/ab_kernels_t.log 2>&1; grep -v Warning gpurun_out/ab_kernels_t.log | tail -n 8
fi
