#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU): launch lists of the default bench command and
# one `ncu --set full` capture of the dominant kernels per workload.  Outputs land in gpurun_out/;
# tools/ncu_summary.py turns the .ncu-rep files into the text summaries committed under profiles/.
# usage: tools/profile_round.sh TAG
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
# launch list of the default bench command (B200_PROFILING.md: time-only pass, unmodified clocks)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_launches_cfg2.log 2>&1
for w in cfg1 cfg3 cfg4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_launches_$w.log 2>&1
done
# full captures (one launch of each dominant kernel, after warm-up)
ncu --set full --import-source on --clock-control none -k regex:fft1_fused --launch-skip 3 -c 1 -f -o $O/${TAG}_fused_cfg2 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_a.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:mix1_kernel --launch-skip 3 -c 1 -f -o $O/${TAG}_mix1_cfg2 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fft1_fused --launch-skip 3 -c 1 -f -o $O/${TAG}_fused_cfg1 \
    python bench.py --workload cfg1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fft1_large --launch-skip 12 -c 2 -f -o $O/${TAG}_large_cfg4 \
    python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_d.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"fft1_large|fft1_real" --launch-skip 27 -c 3 -f -o $O/${TAG}_real_cfg3 \
    python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_e.log 2>&1
ls -la $O
