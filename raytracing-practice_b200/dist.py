"""Multi-GPU plumbing: one process per GPU (torchrun), samples sharded by index, exact int64
reduce to rank 0.  torch.distributed is plumbing only — the data path is the CUDA library.

The path shards naturally: every (pixel, sample) path is independent; the only
exchange step is the per-pixel sum (camera.hpp:55-65).  Each rank renders sample indices
[begin, begin+count) of EVERY pixel into its own fixed-point int64 accumulator; because the
Philox stream is keyed on (pixel, sample, bounce) and integer addition is associative, the reduced
image is bit-identical for any world size.
"""
import os


def shard_samples(spp, rank, world):
    """Contiguous, exhaustive, non-overlapping split of sample indices 0..spp-1 (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    begin = (spp * rank) // world
    end = (spp * (rank + 1)) // world
    return begin, end - begin


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    """Rendezvous from the torchrun environment (127.0.0.1 by default)."""
    import torch
    import torch.distributed as dist

    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def reduce_accum_to_rank0(accum_tensor):
    """Exact sum of the per-rank int64 accumulators onto rank 0 (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return accum_tensor
    assert accum_tensor.dtype == torch.int64
    dist.reduce(accum_tensor, dst=0, op=dist.ReduceOp.SUM)
    return accum_tensor


def clear_accum(ctx, cam):
    """An all-zero accumulator of `cam`'s size — what a rank with an EMPTY sample shard contributes.  (rt_render cannot do
    it: a sample_count of 0 means "all remaining samples" in the C-ABI.)"""
    import numpy as np

    from . import image_height

    ctx.upload_accum(cam, np.zeros((image_height(cam), cam.image_width, 3), np.int64))


def render_sharded(ctx, cam, seed=0):
    """rt_render of this rank's sample shard, then the reduce.  Rank 0's accumulator holds the image."""
    import torch

    rank, _, world = env_rank()
    begin, count = shard_samples(cam.samples_per_pixel, rank, world)
    if count > 0:
        ctx.render(cam, seed=seed, sample_begin=begin, sample_count=count, clear=True)
    else:
        clear_accum(ctx, cam)  # spp < world: this rank has no samples, but the reduce still needs its (zero) accumulator
    ctx.synchronize()
    if world > 1:
        reduce_accum_to_rank0(ctx.accum_tensor())
        torch.cuda.synchronize()


class PeerReduce:
    """The exchange step without a collective call: rank 0 owns a reduce buffer; stream-ordered behind its render
    kernel, every rank's push kernel adds the rank's accumulator into it (rt_render_opts.push_accum: system-scope
    red.add.u64, over NVLink peer memory for the other ranks).  torch.distributed only ships the 64-byte IPC handle
    once and provides the barriers."""

    def __init__(self, ctx, cam):
        import torch
        import torch.distributed as dist

        self.ctx, self.cam = ctx, cam
        self.rank, _, self.world = env_rank()
        self.multi = self.world > 1 and dist.is_available() and dist.is_initialized()
        if self.rank == 0:
            self.ptr, handle = ctx.reduce_buffer(cam)
        else:
            handle = bytes(64)
        if self.multi:
            dev = f"cuda:{ctx.device}" if torch.cuda.is_available() else "cpu"
            t = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            dist.broadcast(t, src=0)
            if self.rank != 0:
                self.ptr = ctx.peer_open(bytes(t.cpu().tolist()))

    def render(self, seed=0, cam=None):
        """One frame: zero the buffer (rank 0), barrier, every rank renders its shard and pushes it, barrier, rank 0 adopts.
        `cam`: another camera of the SAME image size (e.g. another sample count); a rank whose shard is empty pushes nothing."""
        import torch
        import torch.distributed as dist

        cam = self.cam if cam is None else cam
        begin, count = shard_samples(cam.samples_per_pixel, self.rank, self.world)
        if self.rank == 0:
            self.ctx.reduce_buffer(cam)  # re-zero, same pointer
        if self.multi:
            dist.barrier()
        if count > 0:
            self.ctx.render(cam, seed=seed, sample_begin=begin, sample_count=count, clear=True, push_accum=self.ptr)
        elif self.rank == 0:
            clear_accum(self.ctx, cam)  # rank 0 adopts below: its own accumulator must exist
        self.ctx.synchronize()
        if self.multi:
            dist.barrier()
        self.adopt_ms = 0.0
        if self.rank == 0:
            import time

            t0 = time.time()
            self.ctx.adopt_reduce_buffer()
            self.adopt_ms = (time.time() - t0) * 1e3

    def close(self):
        if self.rank != 0 and self.multi:
            self.ctx.peer_close(self.ptr)
