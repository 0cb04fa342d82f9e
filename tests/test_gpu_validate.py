"""GPU parity (through the C ABI: yad_val_labels, yad_val_match, yad_val_ap) of the validator statistics against the
fixtures written by the live reference (tests/golden/match_cases.npz, val_cases.npz) and against the oracle (oracle/metrics.py).
Integer / boolean outputs (the correct matrix, label counts, the max-F1 index) and the fp32 label boxes must be IDENTICAL; the fp64 curves and APs
follow numpy's operation order and are compared at rtol 1e-12."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import metrics as om
from util_gpu import DEV
from yolo_ad_refine_b200 import ops
from yolo_ad_refine_b200.validate import DeviceDetectionStats, ap_per_class, process_batch

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-12, 1e-15


@pytest.mark.parametrize("name", list(cases.MATCH_CASES))
def test_process_batch_bit_exact_vs_reference_golden(gold, name):
    g = gold("match_cases.npz")
    det, gt, gt_cls = cases.match_inputs(*cases.MATCH_CASES[name])
    got = process_batch(torch.from_numpy(det).to(DEV), torch.from_numpy(gt).to(DEV), torch.from_numpy(gt_cls).to(DEV))
    assert got.dtype == torch.bool and tuple(got.shape) == g[name].shape
    np.testing.assert_array_equal(got.cpu().numpy(), g[name])
    np.testing.assert_array_equal(got.cpu().numpy(), om.process_batch(det, gt, gt_cls))


def _feed(stats, v, batch_size, as_list=False):
    """drive update_metrics the way the validator does: NMS output of `batch_size` images + the collated label dict of those images"""
    n_img = len(v["dets"])
    labels, rows = [], []
    for b0 in range(0, n_img, batch_size):
        ids = list(range(b0, min(b0 + batch_size, n_img)))
        sel = np.isin(v["batch_idx"], ids)
        batch = dict(batch_idx=torch.from_numpy(v["batch_idx"][sel] - b0), cls=torch.from_numpy(v["cls"][sel])[:, None],
                     bboxes=torch.from_numpy(v["bboxes"][sel]), ori_shape=[v["ori_shape"][i] for i in ids],
                     ratio_pad=[v["ratio_pad"][i] for i in ids], imgsz=(v["imgsz"], v["imgsz"]))
        if as_list:
            preds = [torch.from_numpy(v["dets"][i]).to(DEV) for i in ids]
        else:
            det = torch.full((len(ids), stats.max_det, 6), 123.0, device=DEV)  # rows >= count hold garbage, as after yad_nms
            for j, i in enumerate(ids):
                det[j, :len(v["dets"][i])] = torch.from_numpy(v["dets"][i]).to(DEV)
            preds = (det, torch.tensor([len(v["dets"][i]) for i in ids], dtype=torch.int32, device=DEV))
            keep = det.clone()
        _, _, gt = stats.update_metrics(preds, batch)
        if not as_list:
            assert torch.equal(preds[0], keep)  # the caller's NMS output is not modified (the reference clones in _prepare_pred)
        labels.append(gt.cpu().numpy())
        rows += [(stats.seen - len(ids) + j, len(v["dets"][i])) for j, i in enumerate(ids)]
    return np.concatenate(labels), rows


@pytest.mark.parametrize("name,batch_size,as_list", [("small", 4, False), ("coco_like", 16, False), ("few_classes_many_dets", 32, True),
                                                     ("sparse", 5, False), ("coco_like", 7, True)])
def test_validation_run_vs_reference_golden(gold, name, batch_size, as_list):
    g = gold("val_cases.npz")
    args = cases.VAL_CASES[name]
    v = cases.val_inputs(*args)
    stats = DeviceDetectionStats(nc=args[1], max_det=300, device=DEV, capacity_images=8)  # small capacity: the arena has to grow
    labels, rows = _feed(stats, v, batch_size, as_list)
    np.testing.assert_array_equal(labels, g[name + "_labels"])  # _prepare_batch: fp32, bit for bit
    n = int(g[name + "_n"])
    tp = np.concatenate([stats._tp[i, :k].cpu().numpy() for i, k in rows if k])
    np.testing.assert_array_equal(tp.astype(bool), np.unpackbits(g[name + "_tp"], axis=0)[:n].astype(bool))  # _process_batch: bit for bit
    for i, k in rows:  # padding rows carry "no detection"
        assert (stats._cls[i, k:] == -1).all() and not stats._tp[i, k:].any()
    res = stats.get_stats()
    np.testing.assert_array_equal(stats.ap_class_index, g[name + "_classes"])
    assert int(stats.result.f1_index.item()) == int(g[name + "_f1_index"])
    u = stats.ap_class_index
    np.testing.assert_allclose(stats.all_ap, g[name + "_ap"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(stats.p, g[name + "_p"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(stats.r, g[name + "_r"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(stats.f1, g[name + "_f1"], rtol=RTOL, atol=ATOL)
    summ = stats.result.summary.cpu().numpy()[u]
    np.testing.assert_array_equal(summ[:, 3], g[name + "_tpn"])
    np.testing.assert_array_equal(summ[:, 4], g[name + "_fpn"])
    for k, t in (("p_curve", stats.result.p_curve), ("r_curve", stats.result.r_curve), ("f1_curve", stats.result.f1_curve)):
        np.testing.assert_allclose(t.cpu().numpy()[u][:, ::8], g[f"{name}_{k}"], rtol=RTOL, atol=ATOL, err_msg=k)
    ap = g[name + "_ap"]
    want = [g[name + "_p"].mean(), g[name + "_r"].mean(), ap[:, 0].mean(), ap.mean()]
    np.testing.assert_allclose([res[k] for k in stats.keys], want, rtol=1e-12)
    np.testing.assert_allclose(res["fitness"], 0.9 * want[2] + 0.1 * want[3], rtol=1e-12)  # this fork's weights (utils/metrics.py:1357)
    np.testing.assert_array_equal(stats.nt_per_class, np.bincount(v["cls"].astype(int), minlength=args[1]))
    per_img = np.zeros(args[1], np.int64)
    for i in range(len(v["dets"])):
        per_img[np.unique(v["cls"][v["batch_idx"] == i]).astype(int)] += 1
    np.testing.assert_array_equal(stats.nt_per_image, per_img)


def _random_stats(n, m, nc, seed):
    rs = np.random.RandomState(seed)
    conf = rs.permutation(n).astype(np.float32) / np.float32(n)  # distinct confidences: no ties
    pred_cls = rs.randint(0, nc, n).astype(np.float32)
    target_cls = rs.randint(0, max(nc - 2, 1), m).astype(np.float32)  # the last classes have predictions but no labels
    hit = rs.rand(n) < (0.6 * conf + 0.1) * min(1.0, 0.9 * m / n)
    for c in range(nc):  # a label is matched at most once: true positives of a class never exceed its labels
        idx = np.nonzero(hit & (pred_cls == c))[0]
        hit[idx[int((target_cls == c).sum()):]] = False
    tp = hit[:, None] & (rs.rand(n, 10).cumsum(1) < 2.5)  # correct at the loose thresholds first, like real matches
    return tp, conf, pred_cls, target_cls


@pytest.mark.parametrize("n,m,nc,seed", [(5000, 900, 12, 0), (257, 40, 3, 1), (100000, 20000, 80, 2), (1, 1, 1, 3), (300, 50, 1, 4)])
def test_ap_per_class_mirror_vs_oracle(n, m, nc, seed):
    tp, conf, pred_cls, target_cls = _random_stats(n, m, nc, seed)
    r = om.ap_per_class(tp, conf, pred_cls, target_cls)
    got = ap_per_class(tp, conf, pred_cls, target_cls)
    assert len(got) == 12
    np.testing.assert_array_equal(got[6], r["unique_classes"])
    for i, k in ((0, "tp"), (1, "fp")):
        np.testing.assert_array_equal(got[i], r[k])
    for i, k in ((2, "p"), (3, "r"), (4, "f1"), (5, "ap"), (7, "p_curve"), (8, "r_curve"), (9, "f1_curve")):
        np.testing.assert_allclose(got[i], r[k], rtol=RTOL, atol=ATOL, err_msg=k)
    np.testing.assert_array_equal(got[10], np.linspace(0, 1, 1000))


def test_ap_per_class_ties_keep_input_order():
    """equal confidences: stable order (the reference's np.argsort leaves it unspecified; the oracle sorts stably)"""
    tp, conf, pred_cls, target_cls = _random_stats(4000, 500, 5, 7)
    conf = np.round(conf * 50) / np.float32(50)
    r = om.ap_per_class(tp, conf, pred_cls, target_cls)
    got = ap_per_class(tp, conf, pred_cls, target_cls)
    np.testing.assert_allclose(got[5], r["ap"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got[8], r["r_curve"], rtol=RTOL, atol=ATOL)


def test_edge_cases():
    # no labels at all, some detections: every row is a false positive, no class has data
    stats = DeviceDetectionStats(nc=4, max_det=50, device=DEV)
    det = torch.rand(2, 50, 6, device=DEV)
    det[:, :, 5] = 1
    cnt = torch.tensor([10, 0], dtype=torch.int32, device=DEV)
    empty = dict(batch_idx=torch.zeros(0), cls=torch.zeros(0, 1), bboxes=torch.zeros(0, 4), ori_shape=[(480, 640)] * 2, imgsz=(640, 640))
    stats.update_metrics((det, cnt), empty)
    res = stats.get_stats()
    assert res == {**{k: 0.0 for k in stats.keys}, "fitness": 0.0} and len(stats.ap_class_index) == 0
    assert not stats._tp[:2].any() and (stats._cls[0, :10] == 1).all() and (stats._cls[0, 10:] == -1).all() and (stats._cls[1] == -1).all()
    # labels but no detections: recall 0 for the labelled classes; nothing to process (no true positive)
    stats.init_metrics()
    batch = dict(batch_idx=torch.tensor([0, 0, 1]), cls=torch.tensor([[0.], [2.], [2.]]), bboxes=torch.tensor([[.5, .5, .2, .2]] * 3),
                 ori_shape=[(480, 640)] * 2, imgsz=(640, 640))
    stats.update_metrics((det, torch.zeros(2, dtype=torch.int32, device=DEV)), batch)
    res = stats.get_stats()
    assert res["metrics/mAP50(B)"] == 0.0 and list(stats.nt_per_class) == [1, 0, 2, 0] and list(stats.nt_per_image) == [1, 0, 2, 0]
    # a run with nothing in it
    stats.init_metrics()
    assert stats.get_stats()["fitness"] == 0.0
    # a perfect detector: AP 1 at every threshold
    stats.init_metrics()
    gt = torch.tensor([[100., 100., 200., 220.], [300., 50., 400., 150.]])
    det = torch.zeros(1, 50, 6, device=DEV)
    det[0, :2, :4] = gt.to(DEV)
    det[0, :2, 4] = torch.tensor([0.9, 0.8], device=DEV)
    det[0, :2, 5] = torch.tensor([1., 3.], device=DEV)
    xywhn = torch.cat([(gt[:, :2] + gt[:, 2:]) / 2, gt[:, 2:] - gt[:, :2]], 1) / 640
    batch = dict(batch_idx=torch.zeros(2), cls=torch.tensor([1., 3.]), bboxes=xywhn, ori_shape=[(640, 640)], imgsz=(640, 640))
    stats.update_metrics((det, torch.tensor([2], dtype=torch.int32, device=DEV)), batch)
    res = stats.get_stats()
    assert list(stats.ap_class_index) == [1, 3]
    np.testing.assert_allclose(stats.all_ap, om.ap_per_class(np.ones((2, 10), bool), np.array([.9, .8], np.float32), np.array([1., 3.]),
                                                             np.array([1., 3.]))["ap"], rtol=RTOL)
    assert res["metrics/mAP50-95(B)"] > 0.99


def test_argument_errors():
    det = torch.zeros(1, 10, 5, device=DEV)
    with pytest.raises(RuntimeError, match="row_ld"):
        ops.val_match(det, None, torch.zeros(1, 4, device=DEV), torch.zeros(1, device=DEV), torch.tensor([0, 1], dtype=torch.int32, device=DEV), 1,
                      torch.linspace(0.5, 0.95, 10, device=DEV), torch.zeros(1, 10, 10, dtype=torch.uint8, device=DEV))
    det = torch.zeros(1, 10, 6, device=DEV)
    with pytest.raises(RuntimeError, match="niou"):
        ops.val_match(det, None, torch.zeros(1, 4, device=DEV), torch.zeros(1, device=DEV), torch.tensor([0, 1], dtype=torch.int32, device=DEV), 1,
                      torch.linspace(0.1, 0.95, 13, device=DEV), torch.zeros(1, 10, 13, dtype=torch.uint8, device=DEV))
    with pytest.raises(RuntimeError, match="shared memory"):
        ops.val_match(det, None, torch.zeros(1, 4, device=DEV), torch.zeros(1, device=DEV), torch.tensor([0, 1], dtype=torch.int32, device=DEV), 100000,
                      torch.linspace(0.5, 0.95, 10, device=DEV), torch.zeros(1, 10, 10, dtype=torch.uint8, device=DEV))


def test_full_size_validation_run_properties():
    """5000 images x 300 detections (a COCO val2017-sized run): the statistics are invariant under the order in which the batches arrive, equal to
    the oracle's on the same rows, and many labels per image (crowded scenes) go through the large-shared-memory path."""
    tp, conf, pred_cls, target_cls = _random_stats(5000 * 300, 36000, 80, 11)
    a = ap_per_class(tp, conf, pred_cls, target_cls)
    perm = np.random.RandomState(5).permutation(len(conf))
    b = ap_per_class(tp[perm], conf[perm], pred_cls[perm], target_cls[::-1].copy())
    for x, y in zip(a[:10], b[:10]):
        np.testing.assert_array_equal(x, y)
    r = om.ap_per_class(tp, conf, pred_cls, target_cls)
    np.testing.assert_allclose(a[5], r["ap"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(a[9], r["f1_curve"], rtol=RTOL, atol=ATOL)
    # crowded image: 3000 labels, 300 detections
    det, gt, gt_cls = cases.match_inputs(300, 3000, 4, 3.0, 123)
    got = process_batch(torch.from_numpy(det).to(DEV), torch.from_numpy(gt).to(DEV), torch.from_numpy(gt_cls).to(DEV))
    np.testing.assert_array_equal(got.cpu().numpy(), om.process_batch(det, gt, gt_cls))
