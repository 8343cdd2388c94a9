#!/bin/bash
# round 2: the default bench line at N GPUs (weak scaling, lb200_reduce_* spectrum sum) + the multi-GPU test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
N=${NGPU:-4}
nvidia-smi --query-gpu=index,name --format=csv > $O/r2s_smi_$N.txt 2>&1
if [ "$N" = "4" ]; then timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q > $O/r2s_mgpu_tests_$N.log 2>&1; tail -3 $O/r2s_mgpu_tests_$N.log; cp $O/mgpu_worker_4.log $O/r2s_mgpu_worker_4.log 2>/dev/null; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-per-config > $O/r2s_bench_${N}gpu.json 2> $O/r2s_bench_${N}gpu.err
echo "bench $N gpus rc=$? $(python -c "import json; d=json.load(open('$O/r2s_bench_${N}gpu.json')); print(round(d['value']), 'Msamples/s, ms/step', round(d['ms_per_step'],2), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('pcie_probe'))" 2>&1 | tail -1)"
echo done
