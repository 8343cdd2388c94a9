#!/bin/bash
# round 2: pipeline v3 -- tests, A/B bench lines at configs[3], the new default bench line, ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2c}
timeout 900 python -m pytest tests/test_pipe_gpu.py -x -q > $O/${T}_pipe_tests.log 2>&1
echo "pipe tests rc=$?" >> $O/${T}_pipe_tests.log
tail -3 $O/${T}_pipe_tests.log
timeout 1200 python -m pytest tests -x -q -m gpu > $O/${T}_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> $O/${T}_gpu_tests.log
tail -3 $O/${T}_gpu_tests.log
B="python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --no-per-config --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/${T}_bench_$name.json 2> $O/${T}_bench_$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/${T}_bench_$name.json')); print('kernel_ms', round(d['roofline']['kernel_ms'],4), 'pass_ms', round(d['ms_per_step']/d['detail']['passes_per_step'],4), 'frac', round(d['roofline']['frac'],3))" 2>&1 | tail -1)"; }
run default X=1
run legacy LB200_LARGE_LEGACY=1
run tmaout LB200_PIPE_TMA_OUT=1
run notmain LB200_PIPE_TMA_IN=0
run lag6 LB200_PIPE_LAG=6 LB200_PIPE_SLOTS=12
run lag8 LB200_PIPE_LAG=8 LB200_PIPE_SLOTS=16
run lag16 LB200_PIPE_LAG=16 LB200_PIPE_SLOTS=32
run nopf LB200_PIPE_PREFETCH=0
timeout 900 python bench.py > $O/${T}_bench_full.json 2> $O/${T}_bench_full.err; echo "full bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; echo "ref bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${T}_launches_cfg4.csv $B --steps 1 --warmup 1 --step-ms 1 > $O/${T}_launches_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_large_pipe -s 2 -c 1 -o $O/${T}_pipe_cfg4 $B --steps 1 --warmup 1 --step-ms 1 > $O/${T}_ncu_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix1_kernel -s 2 -c 1 -o $O/${T}_mix1_cfg4 $B --steps 1 --warmup 1 --step-ms 1 > $O/${T}_ncu_mix1_cfg4.log 2>&1
echo done
