#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T=${TAG:-r2q}
timeout 900 python -m pytest tests/test_timf2_gpu.py -x -q > $O/${T}_timf2_tests.log 2>&1; echo "rc=$?" >> $O/${T}_timf2_tests.log; tail -25 $O/${T}_timf2_tests.log
echo done
