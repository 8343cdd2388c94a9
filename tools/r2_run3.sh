#!/bin/bash
# round 2: pipeline iteration -- pipe tests, A/B bench lines at configs[3], ncu of the pipe and mix1 kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2e}
timeout 900 python -m pytest tests/test_pipe_gpu.py tests/test_wide_graph_gpu.py -x -q > $O/${T}_pipe_tests.log 2>&1
echo "pipe tests rc=$?" >> $O/${T}_pipe_tests.log
tail -3 $O/${T}_pipe_tests.log
B="python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --no-per-config --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/${T}_bench_$name.json 2> $O/${T}_bench_$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/${T}_bench_$name.json')); print('kernel_ms', round(d['roofline']['kernel_ms'],4), 'pass_ms', round(d['ms_per_step']/d['detail']['passes_per_step'],4), 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"; }
run default X=1
run legacy LB200_LARGE_LEGACY=1
run tmaout LB200_PIPE_TMA_OUT=1
run lag8 LB200_PIPE_LAG=8 LB200_PIPE_SLOTS=16
run lag10 LB200_PIPE_LAG=10 LB200_PIPE_SLOTS=18
run lag16 LB200_PIPE_LAG=16 LB200_PIPE_SLOTS=26
run lag20 LB200_PIPE_LAG=20 LB200_PIPE_SLOTS=30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_large_pipe -s 2 -c 1 -o $O/${T}_pipe_cfg4 $B --steps 1 --warmup 1 --step-ms 1 > $O/${T}_ncu_cfg4.log 2>&1
echo done
