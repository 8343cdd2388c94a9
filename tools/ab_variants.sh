#!/bin/bash
# A/B harness: times alternative builds of the same ABI (exp/liblb200_<name>.so, selected through LB200_LIB) next to the
# shipped library in ONE gpurun session: VARIANTS="shipped x y" WORKLOADS="cfg4 cfg3" bash tools/ab_variants.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-ab}
B="python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 10 --warmup 3"
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 300 $B --workload $wl > $O/${T}_bench_${name}_$wl.json 2> $O/${T}_bench_${name}_$wl.err; echo "$name $wl rc=$? $(python -c "import json,sys; d=json.load(open('$O/${T}_bench_${name}_$wl.json')); print('kernel_ms', round(d['roofline']['kernel_ms'],4), 'pass_ms', round(d['ms_per_step']/d['detail']['passes_per_step'],4), 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"; }
for v in $VARIANTS; do
  if [ $v = shipped ]; then L="X=1"; else L="LB200_LIB=$PWD/exp/liblb200_$v.so"; fi
  for wl in ${WORKLOADS:-cfg4}; do run $v $wl $L; done
done
timeout 600 python -m pytest tests/test_reference_cufft_gpu.py -x -q > $O/${T}_cufft_test.log 2>&1; tail -15 $O/${T}_cufft_test.log
LB200_REF_BACKTRACE=1 timeout 300 python -X faulthandler -c "
import bench, json
print(json.dumps(bench.cufft_reference_path('cfg5', seconds=4)))
print(json.dumps(bench.cufft_reference_path('cfg1', seconds=4)))
print(json.dumps(bench.cpu_baseline_quick('cfg5', seconds=4)))
" > $O/${T}_cufft_path.txt 2>&1; tail -5 $O/${T}_cufft_path.txt
echo done
