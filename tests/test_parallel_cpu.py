"""CPU, world_size 2 over gloo: the multi-rank plumbing bench.py uses at N > 1 (image sharding with no collective on the data path, barrier,
max-over-ranks timing)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo_ad_refine_b200 import parallel


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    w, r, _ = parallel.init("gloo")
    start, count = parallel.shard(131, r, w)
    parallel.barrier()
    ms = parallel.max_over_ranks(10.0 + 5.0 * r)
    total = parallel.sum_over_ranks(count)
    q.put((r, start, count, ms, total))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, c0, ms0, t0), (r1, s1, c1, ms1, t1) = res
    assert (s0, c0, s1, c1) == (0, 66, 66, 65)          # contiguous, ragged split covers every image exactly once
    assert ms0 == ms1 == 15.0                           # every rank reports the slowest rank's time
    assert t0 == t1 == 131.0


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    parallel.init("gloo")
    from yolo_ad_refine_b200.train_params import TrainParams
    sd = {"a.conv.weight": torch.ones(4, 3, 3, 3), "a.bn.weight": torch.ones(4), "a.bn.running_mean": torch.full((4,), float(rank)),
          "a.bn.running_var": torch.ones(4)}
    tp = TrainParams(sd, torch.float32, "cpu")
    tp.grad.fill_(float(rank + 1))
    parallel.exchange_gradients(tp)
    q.put((rank, float(tp.grad.min()), float(tp.grad.max()), float(tp.buf("a.bn.running_mean").max())))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_exchange():
    """training's one collective: SUM all-reduce of the flat gradient arena, rank-0 BatchNorm buffers (DDP semantics, trainer.py:273,387)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, lo, hi, rm in res:
        assert lo == hi == 3.0 and rm == 0.0


def test_shard_properties():
    for total in (0, 1, 7, 64, 129):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert parallel.max_over_ranks(3.5) == 3.5 and not dist.is_initialized()
