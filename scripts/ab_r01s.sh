#!/bin/bash
# A/B on the GPU box: attention forward with 128-row CTAs (two per SM) vs 256-row CTAs; LayerNorm forward v1 vs v2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests_r01s.log 2>&1
echo "gpu tests (defaults) exit=$?"; tail -n 5 gpurun_out/gpu_tests_r01s.log
b() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 30 --warmup 5 > gpurun_out/ab_$name.log 2>&1
  echo "$name: $(tail -n 1 gpurun_out/ab_$name.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>&1 | tail -n 1)"
}
b s_default DSF_X=0
b s_nwg2 DSF_ATTN_FWD_NWG=2
b s_nwg1 DSF_ATTN_FWD_NWG=1
b s_lnfwd1 DSF_LN_FWD_IMPL=1
b s_default2 DSF_X=0
b s_a16 DSF_X=0 
(python scripts/bench_kernels.py attnq) > gpurun_out/ab_kernels_s.log 2>&1
grep -v Warning gpurun_out/ab_kernels_s.log | tail -n 20
python bench.py --quick --steps 10 --warmup 3 --anchors 16 > gpurun_out/ab_s_anchors16.log 2>&1; tail -n 1 gpurun_out/ab_s_anchors16.log | cut -c1-200
DSF_ATTN_FWD_NWG=1 python bench.py --quick --steps 10 --warmup 3 --anchors 16 > gpurun_out/ab_s_anchors16_nwg1.log 2>&1; tail -n 1 gpurun_out/ab_s_anchors16_nwg1.log | cut -c1-200
