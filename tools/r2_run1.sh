#!/bin/bash
# round 2, first GPU run of the persistent four-step kernel: tests, then A/B bench lines at configs[3]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2b_smi.txt 2>&1
timeout 900 python -m pytest tests/test_pipe_gpu.py -x -q > $O/r2b_pipe_tests.log 2>&1
echo "pipe tests rc=$?" >> $O/r2b_pipe_tests.log
tail -5 $O/r2b_pipe_tests.log
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r2b_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> $O/r2b_gpu_tests.log
tail -4 $O/r2b_gpu_tests.log
B="python bench.py --workload cfg4 --no-e2e --no-cpu-baseline --steps 20 --warmup 5"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/r2b_bench_$name.json 2> $O/r2b_bench_$name.err; echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$O/r2b_bench_$name.json')); print(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'])" 2>&1 | tail -1)"; }
run default X=1
run tmaout LB200_PIPE_TMA_OUT=1
run notmain LB200_PIPE_TMA_IN=0
run lag4 LB200_PIPE_LAG=4 LB200_PIPE_SLOTS=8
run lag6s10 LB200_PIPE_LAG=6 LB200_PIPE_SLOTS=10
run lag12 LB200_PIPE_LAG=12 LB200_PIPE_SLOTS=24
run lag16 LB200_PIPE_LAG=16 LB200_PIPE_SLOTS=32
run nopf LB200_PIPE_PREFETCH=0
timeout 300 python bench.py --workload cfg3 --no-e2e --no-cpu-baseline > $O/r2b_bench_cfg3.json 2> $O/r2b_bench_cfg3.err; echo "cfg3 rc=$?"
LB200_LARGE_LEGACY=1 timeout 300 python bench.py --workload cfg3 --no-e2e --no-cpu-baseline > $O/r2b_bench_cfg3_legacy.json 2>&1
timeout 300 python bench.py --workload cfg5 --no-e2e --no-cpu-baseline > $O/r2b_bench_cfg5.json 2> $O/r2b_bench_cfg5.err; echo "cfg5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2b_launches_cfg4.csv $B --steps 2 --warmup 1 > $O/r2b_launches_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_large_pipe -s 1 -c 1 -o $O/r2b_pipe_cfg4 $B --steps 2 --warmup 1 > $O/r2b_ncu_cfg4.log 2>&1
echo done
