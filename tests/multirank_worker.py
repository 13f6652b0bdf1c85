"""Worker of tests/test_gpu_multirank.py (launched by torch.distributed.run, one process per GPU): renders one job
sharded over the ranks with both forms of the exchange step and prints rank 0's accumulator hash."""
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

rtb = importlib.import_module("raytracing-practice_b200")
dist = importlib.import_module("raytracing-practice_b200.dist")


def main():
    scene, width, spp, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, local_rank, world = dist.init_process_group()
    torch.cuda.set_device(local_rank)
    ctx = rtb.Context(local_rank)
    sc = rtb.Scene(scene, rand_seed=1)
    cam = sc.camera_copy(image_width=width, samples_per_pixel=spp)
    ctx.upload_scene(sc.desc)
    out = {"world": world}
    # 1. our push kernel over peer memory (CUDA IPC), no collective
    peer = dist.PeerReduce(ctx, cam)
    peer.render(seed)
    if rank == 0:
        out["peer"] = hashlib.sha256(ctx.download_accum().tobytes()).hexdigest()
    # 2. the collective form: exact ncclInt64 reduce to rank 0
    dist.render_sharded(ctx, cam, seed=seed)
    if rank == 0:
        out["nccl"] = hashlib.sha256(ctx.download_accum().tobytes()).hexdigest()
    # 3. another GPU's accumulator as peer_accum is refused (device-scope adds): rank 1 tries rank 0's IPC-mapped buffer
    if world > 1:
        torch.distributed.barrier()
        if rank == 1:
            try:
                ctx.render(cam, seed=seed, peer_accum=peer.ptr)
                out_r1 = "accepted"
            except rtb.RtError as e:
                out_r1 = "refused"
            t = torch.tensor([1 if out_r1 == "refused" else 0], device=f"cuda:{local_rank}")
        else:
            t = torch.tensor([1], device=f"cuda:{local_rank}")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN)
        out["foreign_peer_accum_refused"] = bool(t.item())
    peer.close()
    if rank == 0:
        print("RESULT " + json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
