"""Real multi-process, multi-GPU parity of the exchange step (camera.hpp:55-65 is the sum being split): when the box has
>= 2 devices, 2 (and 4, 8 if present) ranks render one job sharded by sample index; rank 0's reduced int64 accumulator must
equal the single-GPU one BIT FOR BIT, through our push kernel over peer memory and through the ncclInt64 reduce."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_on_real_gpus_reproduce_the_single_gpu_accumulator(rtb, gpu_ctx, world):
    if _device_count() < world:
        pytest.skip(f"needs {world} CUDA devices, this box has {_device_count()}")
    scene, width, spp, seed = "book2_final", 200, 37, 21  # 37 spp: ragged shards
    sc = rtb.Scene(scene, rand_seed=1)
    cam = sc.camera_copy(image_width=width, samples_per_pixel=spp)
    gpu_ctx.upload_scene(sc.desc)
    gpu_ctx.render(cam, seed=seed)
    want = hashlib.sha256(gpu_ctx.download_accum().tobytes()).hexdigest()
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multirank_worker.py"), scene, str(width), str(spp), str(seed)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert r.returncode == 0 and line, (r.stdout[-1500:], r.stderr[-3000:])
    out = json.loads(line[-1][7:])
    assert out["world"] == world
    assert out["peer"] == want, "push kernel over peer memory: reduced accumulator differs from the single-GPU render"
    assert out["nccl"] == want, "ncclInt64 reduce: reduced accumulator differs from the single-GPU render"
    assert out["foreign_peer_accum_refused"]
