"""Multi-GPU plumbing, one process per GPU.  Inference shards by image with NO data-path collective: torch.distributed
is used only for the rendezvous, the barrier around the timed region and the max-over-ranks reduction of the timing.  Training has one real
exchange step per iteration -- the reference's DistributedDataParallel gradient all-reduce (engine/trainer.py:273) -- done here as ONE
all-reduce of the flat fp32 gradient arena, plus DDP's rank-0 broadcast of the BatchNorm buffers."""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return world, rank, local


def shard(total, rank, world):
    """Contiguous split of `total` images over `world` ranks (the reference's per-rank batch = batch // world, engine/trainer.py:290,
    generalised to ragged totals): returns (start, count)."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (device-timed milliseconds of the slowest rank)"""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def exchange_gradients(tp, group=None):
    """DDP semantics of the reference trainer on the flat arenas of train_params.TrainParams: gradients are averaged by DDP and the loss is
    multiplied by world_size (engine/trainer.py:387), i.e. the applied gradient is the SUM over ranks; module buffers (BatchNorm running
    statistics) follow rank 0 (DistributedDataParallel(broadcast_buffers=True), its default)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(tp.grad, op=dist.ReduceOp.SUM, group=group)
    if tp.btotal:
        dist.broadcast(tp.bufs, src=0, group=group)
