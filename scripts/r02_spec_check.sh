#!/bin/bash
# One short GPU pass for the speculative-tile switch of the forward attention kernel (DSF_ATTN_SPEC): the new moving-maximum test
# under both settings, the forward kernel timed alone under both, the whole GPU suite with the switch on, two quick bench lines.
O=gpurun_out; mkdir -p $O; tag=${1:-r02sp}
for s in 0 1; do
  DSF_ATTN_SPEC=$s timeout 60 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "running_maximum" -p no:cacheprovider > $O/newtest_${tag}_spec$s.log 2>&1
  echo "moving-maximum test, spec=$s: rc=$? $(tail -n 1 $O/newtest_${tag}_spec$s.log)"
done
for s in 0 1 0 1; do
  DSF_ATTN_SPEC=$s timeout 40 python scripts/bench_attn_parts.py spec$s fwd >> $O/attn_fwd_${tag}.log 2>&1
done
cat $O/attn_fwd_${tag}.log
DSF_ATTN_SPEC=1 timeout 120 python -m pytest tests -q -m gpu --maxfail=10 -p no:cacheprovider > $O/gpu_tests_${tag}_spec1.log 2>&1
echo "gpu suite, spec=1: rc=$? $(tail -n 1 $O/gpu_tests_${tag}_spec1.log)"
for s in 1 0; do
  DSF_ATTN_SPEC=$s timeout 60 python bench.py --quick --steps 20 --warmup 5 > $O/bench_quick_${tag}_spec$s.log 2>&1
  echo "bench --quick, spec=$s: $(tail -n 1 $O/bench_quick_${tag}_spec$s.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>&1 | tail -n 1)"
done
