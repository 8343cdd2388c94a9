#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2k}
timeout 900 python -m pytest tests/test_pipe_gpu.py -x -q > $O/${T}_pipe_tests.log 2>&1; tail -3 $O/${T}_pipe_tests.log
B="python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --no-per-config --steps 3 --warmup 2 --step-ms 1"
for v in "X=1" "LB200_PIPE_LAG=10 LB200_PIPE_SLOTS=20" "LB200_PIPE_LAG=16 LB200_PIPE_SLOTS=32"; do
  echo "== $v"; env $v LB200_PIPE_STATS=1 timeout 300 $B 2>&1 >/dev/null | tail -2
done
B="python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --no-per-config --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/${T}_bench_$name.json 2> $O/${T}_bench_$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/${T}_bench_$name.json')); print('kernel_ms', round(d['roofline']['kernel_ms'],4), 'pass_ms', round(d['ms_per_step']/d['detail']['passes_per_step'],4), 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"; }
run v5b_default X=1
run v5b_lag10 LB200_PIPE_LAG=10 LB200_PIPE_SLOTS=20
run v5b_lag16 LB200_PIPE_LAG=16 LB200_PIPE_SLOTS=32
run v1_stg_13_26 LB200_LIB=$PWD/exp/liblb200_v1.so LB200_PIPE_TMA_OUT=0 LB200_PIPE_LAG=13 LB200_PIPE_SLOTS=26
run v1_stg_10_20 LB200_LIB=$PWD/exp/liblb200_v1.so LB200_PIPE_TMA_OUT=0 LB200_PIPE_LAG=10 LB200_PIPE_SLOTS=20
run v1_stg_8_16 LB200_LIB=$PWD/exp/liblb200_v1.so LB200_PIPE_TMA_OUT=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_large_pipe -s 2 -c 1 -o $O/${T}_pipe_cfg4 $B --steps 1 --warmup 1 --step-ms 1 > $O/${T}_ncu_cfg4.log 2>&1
LB200_LIB=$PWD/exp/liblb200_v1.so LB200_PIPE_TMA_OUT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_large_pipe -s 2 -c 1 -o $O/${T}_pipe_cfg4_v1 $B --steps 1 --warmup 1 --step-ms 1 > $O/${T}_ncu_cfg4_v1.log 2>&1
echo done
