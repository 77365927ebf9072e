"""GPU, >= 2 devices: sharded frame == single-GPU frame, bit for bit (tests/mgpu_worker.py under torch.distributed.run,
one process per GPU, NCCL). Skipped on a single-GPU box; the world_size-2 gloo test covers the host-side plan on CPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_frame_equals_single_gpu_frame(gpu):
    import tg_b200
    n = tg_b200.lib().tgb200_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29641", os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.count("MGPU_OK") == 2, out.stdout[-3000:] + out.stderr[-4000:]
