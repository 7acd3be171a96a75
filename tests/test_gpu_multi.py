"""GPU, >= 2 devices: gene-sharded fit == unsharded fit (scripts/check_multi_gpu.py under torchrun).
Skipped on a single-GPU box; the CPU (gloo) analogue is tests/test_sharded_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_unsharded_on_two_gpus():
    from bayesvg_b200 import _lib
    n = _lib.lib().bvg_device_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs (have {n})")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(ROOT, "scripts", "check_multi_gpu.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU check ok" in r.stdout
