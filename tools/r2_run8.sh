#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2p}
B="python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 10 --warmup 3"
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 300 $B --workload $wl > $O/${T}_bench_${name}_$wl.json 2> $O/${T}_bench_${name}_$wl.err; echo "$name $wl rc=$? $(python -c "import json,sys; d=json.load(open('$O/${T}_bench_${name}_$wl.json')); print('kernel_ms', round(d['roofline']['kernel_ms'],4), 'pass_ms', round(d['ms_per_step']/d['detail']['passes_per_step'],4), 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"; }
for v in $VARIANTS; do
  if [ $v = shipped ]; then L="X=1"; else L="LB200_LIB=$PWD/exp/liblb200_$v.so"; fi
  env $L timeout 900 python -m pytest tests/test_pipe_gpu.py -x -q > $O/${T}_pipe_tests_$v.log 2>&1; echo "$v tests: $(tail -1 $O/${T}_pipe_tests_$v.log)"
  for wl in ${WORKLOADS:-cfg4 cfg3}; do run $v $wl $L; done
done
run shipped_again cfg4 X=1
echo done
