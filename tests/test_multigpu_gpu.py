"""Multi-GPU paths on real devices: needs >= 2 GPUs, otherwise skipped (the
world-size-2 gloo tests of tests/test_oracle_cpu.py cover the host logic on CPU).  One process per
GPU under torch.distributed.run; see tests/mgpu_worker.py for what is checked."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu]


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_reduce_and_time_block_sharding_on_real_devices(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 400) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):                  # keep the whole transcript of the ranks (pytest abbreviates it)
        with open(os.path.join(out_dir, f"mgpu_worker_{world}.log"), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "MGPU_OK" in r.stdout, r.stdout[-3000:]
