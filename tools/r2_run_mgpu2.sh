#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q > $O/r2p_mgpu_tests.log 2>&1; echo "mgpu tests rc=$?" >> $O/r2p_mgpu_tests.log; tail -5 $O/r2p_mgpu_tests.log
grep -n "reduce:\|time-block\|MGPU_OK\|Error" $O/mgpu_worker_2.log | head
