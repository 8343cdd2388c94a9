#!/bin/bash
# round 2, multi-GPU: the reducer / time-block sharding test on real devices, then the bench line at N GPUs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2o}
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv > $O/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q > $O/${T}_mgpu_tests.log 2>&1; echo "mgpu tests rc=$?" >> $O/${T}_mgpu_tests.log; tail -15 $O/${T}_mgpu_tests.log
for red in p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-per-config --reduce $red > $O/${T}_bench_${N}gpu_$red.json 2> $O/${T}_bench_${N}gpu_$red.err
echo "bench $N gpus $red rc=$? $(python -c "import json; d=json.load(open('$O/${T}_bench_${N}gpu_$red.json')); print(round(d['value']), 'Msamples/s, ms/step', round(d['ms_per_step'],2), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['detail']['spectrum_reduction'][:40])" 2>&1 | tail -1)"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-per-config > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err
echo "bench 1 gpu rc=$? $(python -c "import json; d=json.load(open('$O/${T}_bench_1gpu.json')); print(round(d['value']), 'Msamples/s, ms/step', round(d['ms_per_step'],2), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']))" 2>&1 | tail -1)"
echo done
