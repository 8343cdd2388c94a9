#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2m}
timeout 900 python -m pytest tests/test_pipe_gpu.py -x -q > $O/${T}_pipe_tests.log 2>&1; tail -3 $O/${T}_pipe_tests.log
B="python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --no-per-config --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/${T}_bench_$name.json 2> $O/${T}_bench_$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/${T}_bench_$name.json')); print('kernel_ms', round(d['roofline']['kernel_ms'],4), 'pass_ms', round(d['ms_per_step']/d['detail']['passes_per_step'],4), 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"; }
run shipped X=1
run e4 LB200_LIB=$PWD/exp/liblb200_e4.so
run e7 LB200_LIB=$PWD/exp/liblb200_e7.so
run e7_lag12 LB200_LIB=$PWD/exp/liblb200_e7.so LB200_PIPE_LAG=12 LB200_PIPE_SLOTS=24
run shipped_lag12 LB200_PIPE_LAG=12 LB200_PIPE_SLOTS=24
run shipped_lag6 LB200_PIPE_LAG=6 LB200_PIPE_SLOTS=12
run shipped_again X=1
echo done
