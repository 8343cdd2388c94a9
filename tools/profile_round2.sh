#!/bin/bash
# round 2: the evidence run -- GPU tests, the default bench line and the reference arm as the driver runs them,
# ncu launch list of the default workload, ncu --set full of the dominant kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2_final}
rm -f $O/parity_report.jsonl
timeout 1800 python -m pytest tests -x -q -m gpu > $O/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> $O/${T}_gpu_tests.log; tail -4 $O/${T}_gpu_tests.log
cp $O/parity_report.jsonl $O/${T}_parity_report.jsonl 2>/dev/null
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference arm rc=$?"
python -c "
import json
d=json.load(open('$O/${T}_bench.json')); r=json.load(open('$O/${T}_bench_reference.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'ref', round(r['value'],1), 'same_config', d['config']==r['config'])
for k,v in d['per_config'].items(): print(k, round(v['value']), round(v['frac'],3), round(v['kernel_ms'],4))
print(d['e2e'].get('pcie_probe'))
"
Q="--no-e2e --no-cpu-baseline --no-per-config --steps 1 --warmup 1 --step-ms 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${T}_launches_cfg5.csv python bench.py $Q > $O/${T}_launches_cfg5.log 2>&1
for w in cfg1 cfg2 cfg3 cfg4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/${T}_launches_$w.csv python bench.py --workload $w $Q > $O/${T}_launches_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_large_pipe -s 2 -c 1 -o $O/${T}_pipe_cfg4 python bench.py --workload cfg4 $Q > $O/${T}_ncu_pipe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_large_pipe -s 2 -c 1 -o $O/${T}_pipe_cfg3 python bench.py --workload cfg3 $Q > $O/${T}_ncu_pipe3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_fused -s 2 -c 1 -o $O/${T}_fused_cfg2 python bench.py --workload cfg2 $Q > $O/${T}_ncu_fused2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft1_fused -s 2 -c 1 -o $O/${T}_fused_cfg1 python bench.py --workload cfg1 $Q > $O/${T}_ncu_fused1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix1_kernel -s 2 -c 1 -o $O/${T}_mix1_cfg4 python bench.py --workload cfg4 $Q > $O/${T}_ncu_mix1.log 2>&1
echo done
