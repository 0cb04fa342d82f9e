"""Row f4, GPU: DevicePrefetcher (the device end of the reference's InfiniteDataLoader, data/build.py:28-141) delivers every batch of the host
iterable, in order, bit for bit, with the collate format preserved, over several epochs, and feeds a graphed training step."""
import numpy as np
import pytest
import torch

from yolo_ad_refine_b200.data import DevicePrefetcher

pytestmark = pytest.mark.gpu


def _batches(n, b=3, hw=64):
    rs = np.random.RandomState(1)
    out = []
    for i in range(n):
        m = int(rs.randint(0, 7))
        out.append(dict(img=torch.from_numpy(rs.randint(0, 256, (b, 3, hw, hw), dtype=np.uint8)), batch_idx=torch.from_numpy(np.sort(rs.randint(0, b, m)).astype(np.float32)),
                        cls=torch.from_numpy(rs.randint(0, 80, (m, 1)).astype(np.float32)), bboxes=torch.from_numpy(rs.rand(m, 4).astype(np.float32)),
                        im_file=[f"{i}_{k}.jpg" for k in range(b)]))
    return out


def test_prefetcher_delivers_every_batch_in_order_over_epochs():
    host = _batches(7)
    pf = DevicePrefetcher(host)
    assert len(pf) == 7
    for epoch in range(2):
        seen = 0
        for got, want in zip(pf, host):
            for k in ("img", "batch_idx", "cls", "bboxes"):
                assert got[k].is_cuda
                # consume on the current stream (a kernel), then compare on the host
                assert torch.equal((got[k] + 0).cpu(), want[k]), (epoch, seen, k)
            assert got["im_file"] == want["im_file"]
            seen += 1
        assert seen == 7
    pf.reset()
    assert sum(1 for _ in pf) == 7


def test_prefetcher_feeds_graphed_training(state_dict):
    from yolo_ad_refine_b200.trainer import TrainEngine
    rs = np.random.RandomState(2)
    host = []
    for i in range(3):
        bi = np.sort(rs.randint(0, 2, 5)).astype(np.float32)
        host.append(dict(img=torch.from_numpy(rs.randint(0, 256, (2, 3, 160, 160), dtype=np.uint8)), batch_idx=torch.from_numpy(bi),
                         cls=torch.from_numpy(rs.randint(0, 80, (5, 1)).astype(np.float32)),
                         bboxes=torch.from_numpy(np.concatenate([rs.uniform(0.3, 0.7, (5, 2)), rs.uniform(0.05, 0.3, (5, 2))], 1).astype(np.float32))))
    eng = TrainEngine(state_dict, dtype=torch.bfloat16)
    eng.capture(2, 160, n_max=8)
    losses = [eng.step_graphed(b["img"], b["batch_idx"], b["cls"], b["bboxes"]).cpu().numpy().copy() for b in DevicePrefetcher(host)]
    assert len(losses) == 3 and all(np.isfinite(l).all() and l[3] > 0 for l in losses) and eng.tp.steps == 3
