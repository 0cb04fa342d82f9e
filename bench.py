#!/usr/bin/env python
"""bench.py -- YOLO-AD-Refine inference hot path (forward + fused DFL decode + batched NMS) at 640x640, batch 64 per GPU, bf16.

Contract (see DESIGN.md "Measurement"):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  value        img/s, whole job, inputs already resident in HBM (CUDA-graph replay of the full hot path), device-timed with CUDA events,
               max over ranks.
  e2e          same metric through the public API with HOST buffers: pinned uint8 image batch -> H2D -> hot path -> D2H of the detections,
               every step inside the timed region.
  roofline     for the dominant kernel (the implicit-GEMM convolution entry point yad_conv2d): algorithmic FLOPs / CUDA-event time of its
               launches, measured live in one instrumented eager step, against the measured bf16 peak in MEASURED_PEAKS.json.
  cpu_baseline the oracle port of the reference's PyTorch path, fp32, all host threads, on a bounded sample (rank 0, N=1 only).
--impl reference times that same CPU path as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))  # set explicitly, before torch is imported

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "img/s @640^2 b64 fwd+decode+NMS"
NMS_ARGS = dict(conf_thres=0.25, iou_thres=0.7, max_det=300)
FLOPS_PER_IMG_640 = 12.44e9  # SURVEY.md section 8d (conv + linear + bmm + DCN + MHA, 2*MAC)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.samples, self._stop, self._t = gpu, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def make_config(args, world):
    """the `config` object of the JSON line -- identical for both arms (the reference arm times a bounded sample of this workload)"""
    return {"workload": f"YOLO-AD-Refine (yolo11-701 yaml, scale n) inference batch {args.batch}/GPU at {args.imgsz}x{args.imgsz}, "
                        "forward + DFL decode + NMS(conf .25, iou .7, max_det 300), random-init synthetic weights",
            "global_batch": world * args.batch, "parallelism": f"batch-sharded replicas x{world}, no collective",
            "pipeline": ("one batch in flight: " if getattr(args, "no_overlap", False) else
                         f"{getattr(args, 'in_flight', 2)} batches in flight (each buffer set has its own input, forward stream and CUDA graph; every batch "
                         "is computed exactly as alone, the GPU interleaves the kernels of consecutive batches): ") +
                        "forward + decode of batch i + 1 overlaps the NMS of batch i (two CUDA graphs, NMS stream, double-buffered predictions)",
            "l2": "inputs + activations of one step (>1 GB) exceed the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's OWN implementation (oracle/_ref, the copy of its Python package made by oracle/build_ref.py; /root/reference in the
# build container) on the host cores; the oracle port only where neither exists.  Baseline / reference arm only -- never the product path.
# ---------------------------------------------------------------------------------------------------------------------------
def cpu_steps(batch, imgsz, steps, warmup):
    """(seconds per step, threads, kind): `steps` timed passes of forward + decode + NMS over `batch` synthetic images after `warmup` untimed ones"""
    from oracle import refrun
    if refrun.available():
        sec, cores = refrun.time_steps(batch, imgsz, steps, warmup)
        return sec, cores, "reference"
    from oracle import model as om
    from oracle import postprocess as op
    from yolo_ad_refine_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_state_dict(seed=1)
    img = torch.from_numpy(synth.make_images(batch, imgsz, imgsz, seed=2))
    with torch.inference_mode():
        def one():
            y, _ = om.forward(sd, img)
            op.non_max_suppression(y.numpy(), **NMS_ARGS)
        for _ in range(warmup):
            one()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
    return (time.perf_counter() - t0) / max(1, steps), torch.get_num_threads(), "port"


def cpu_sample_text(kind, b, imgsz):
    who = ("the reference's own DetectionModel + ops.non_max_suppression (oracle/_ref, unmodified, fused eval model)" if kind == "reference"
           else "oracle port (fp32 torch-CPU restatement of the reference path)")
    return f"{who}, fp32, batch {b} at {imgsz}^2 per step"


def run_reference(args):
    """Reference arm: exactly --steps timed steps after --warmup untimed ones; one step = forward + decode + NMS of a bounded sample (batch 4) of the
    workload through the reference's own CPU code on all host threads.  value = images of the sample / measured seconds per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = args.ref_batch
    t0 = time.perf_counter()
    sec, cores, kind = cpu_steps(b, args.imgsz, args.steps, args.warmup)
    v = b / sec
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": make_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": kind, "sample": cpu_sample_text(kind, b, args.imgsz)},
            "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "images_per_step": b, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def measure_training(args, world, rank, sd, dtype, W, barrier, B=None, scaling="weak"):
    """BASELINE.json configs[3]: one training step (train()-mode forward + TaskAlignedAssigner + v8DetectionLoss + full backward + gradient
    all-reduce over NCCL when N > 1 + clip / SGD / EMA), B images per GPU.  The step is a captured CUDA graph (TrainEngine.capture): one graph at
    N = 1; forward + backward | NCCL all-reduce | optimizer at N > 1.  Device-timed, max over ranks.  The e2e figure adds, per step, the H2D copy
    of the pinned uint8 image batch + targets (prefetched on a copy stream under the previous step) and the D2H read of the loss items."""
    from yolo_ad_refine_b200 import ops, parallel, synth
    from yolo_ad_refine_b200.trainer import TrainEngine
    B = B or args.train_batch
    eng = TrainEngine(sd, dtype=dtype, world_size=world)
    # uint8 images, as the reference's dataloader delivers them (the /255 happens on the device: models/yolo/detect/train.py:57-59)
    rs = np.random.RandomState(200 + rank)
    img_host = torch.from_numpy(rs.randint(0, 256, (B, 3, args.imgsz, args.imgsz), dtype=np.uint8)).pin_memory()
    tg_host = [torch.from_numpy(a).float().pin_memory() for a in synth.make_targets(B, seed=300 + rank, max_per_img=8, empty_images=())]
    loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
    l0 = ops.LAUNCHES
    eng.capture(B, args.imgsz, m_cap=8 * B, n_max=8)
    launches = (ops.LAUNCHES - l0) // 3  # two warm-up passes + the captured one
    eng.load_static(img_host, *tg_host)
    torch.cuda.synchronize()

    def timed(fn, k):
        for _ in range(max(W, 3)):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        e.record()
        barrier()
        return parallel.max_over_ranks(s.elapsed_time(e), device="cuda") / k

    ms = timed(lambda: eng.step_graphed(), args.train_steps)

    # end to end: pinned host batch -> (copy stream) staging -> static inputs -> step -> loss items on the host, every step
    copy_stream = torch.cuda.Stream()
    stage = [(torch.empty_like(eng._s_img), [torch.empty(t.shape, dtype=torch.float32, device="cuda") for t in tg_host]) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()
    for ev in free:
        ev.record(main)
    state = {"i": 0, "primed": False}

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i % 2])
            stage[i % 2][0].copy_(img_host, non_blocking=True)
            for d, h in zip(stage[i % 2][1], tg_host):
                d.copy_(h, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_step():
        i = state["i"]
        if not state["primed"]:
            prefetch(i)
            state["primed"] = True
        prefetch(i + 1)                        # the next batch's H2D runs under this step
        main.wait_event(ready[i % 2])
        out4 = eng.step_graphed(stage[i % 2][0], *stage[i % 2][1])
        free[i % 2].record(main)
        loss_host.copy_(out4, non_blocking=True)
        main.synchronize()                     # the trainer formats the loss items every iteration
        state["i"] = i + 1

    e2e_ms = timed(e2e_step, args.train_steps)
    # where the step goes at N > 1 (CUDA events on this rank, a few steps): compute graph | all-reduce (wire + waiting for the slowest rank) | optimizer
    decomp = None
    if world > 1:
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(4)]
        barrier()
        for k in range(4):
            eng.tp.set_hyper(lr=eng.lr, momentum=eng.momentum, weight_decay=eng.weight_decay, optimizer=eng.optimizer)
            evs[k][0].record()
            eng._gA.replay()
            evs[k][1].record()
            eng.exchange()
            evs[k][2].record()
            eng._gB.replay()
            evs[k][3].record()
        torch.cuda.synchronize()
        decomp = {"compute_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in evs[1:]])),
                  "allreduce_ms": float(np.mean([e[1].elapsed_time(e[2]) for e in evs[1:]])),
                  "optimizer_ms": float(np.mean([e[2].elapsed_time(e[3]) for e in evs[1:]])),
                  "note": "rank 0; allreduce_ms = NCCL SUM all-reduce of the 16.4 MB fp32 gradient arena + rank-0 broadcast of the BatchNorm buffers "
                          "(DDP semantics) + waiting for the slowest rank's backward"}
    # roofline of the training step's dominant kernels, measured live: one instrumented EAGER step, CUDA events around every libyad entry point
    torch.cuda.synchronize()
    img, tg = eng._s_img, (eng._s_bi, eng._s_cls, eng._s_box)
    ops.PROFILE = {}
    torch.cuda._sleep(int(200e6))  # keep the host ahead of the device (see the inference profile): ~100 ms of queued spin
    eng.forward_backward(img, *tg, static_n_max=8)
    eng.tp.optimizer_step_dev(eng.optimizer)
    torch.cuda.synchronize()
    prof = {k: (sum(a.elapsed_time(b) for a, b, _ in v), sum(m["flops"] for _, _, m in v if m), len(v)) for k, v in ops.PROFILE.items()}
    ops.PROFILE = None
    tot_ms = sum(v[0] for v in prof.values())
    pk = peaks()
    roof = {}
    for key, label in (("yad_conv_wgrad", "weight gradients: wgrad_tc_kernel (tcgen05 / TMEM, TMA-fed MN-major operands) + small-channel / split-K mma.sync kernels"),
                       ("yad_conv2d", "forward convolutions + input gradients: conv2_kernel / conv_tma_kernel / conv_tc_kernel / conv_small_kernel")):
        ms_k, fl, n = prof[key]
        roof[key] = {"kernel": label, "bound": "tensor", "achieved": fl / (ms_k / 1000.0) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": fl / (ms_k / 1000.0) / 1e12 / pk["tf_sustained"], "launches": n, "ms": ms_k, "share_of_step": ms_k / tot_ms,
                     "traffic": None, "note": "most launches are HBM-bound (1x1 wgrad reads 5.6 TB/s = 86 % of the measured HBM peak, profiles/r1_ncu_wgrad_tc.json)"}
    out = {"metric": "train img/s (forward + loss + backward + optimizer)", "value": world * B / (ms / 1000.0), "unit": "img/s",
           "ms_per_step": ms, "batch_per_gpu": B, "global_batch": world * B, "steps": args.train_steps, "launches_per_step": launches,
           "scaling": scaling, "graph": "one CUDA graph per step" if world == 1 else "graph(forward+backward) | NCCL all-reduce | graph(optimizer)",
           "exchange": "none (1 GPU)" if world == 1 else f"NCCL all-reduce (SUM) of the {eng.tp.total * 4 / 1e6:.1f} MB fp32 gradient arena",
           "e2e": {"value": world * B / (e2e_ms / 1000.0), "unit": "img/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": img_host.numel() + sum(t.numel() * 4 for t in tg_host), "d2h_bytes_per_step": 16,
                   "h2d": "prefetched on a copy stream under the previous step (double-buffered staging)"},
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "loss": [float(v) for v in loss_host], "roofline": roof,
           "profile_ms": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:12]}}
    if decomp:
        out["decomposition"] = decomp
    del eng, stage
    torch.cuda.empty_cache()
    return out


def run_nms_micro(args):
    """BASELINE.json configs[2]: fused DFL decode (yad_decode) + batched NMS (yad_nms) on synthetic head logits (batch 256,
    8400 anchors, 80 classes, bf16 level maps, class logits -6 + 1.5 randn => ~350 candidates per image at conf 0.25).  HBM-bound: algorithmic bytes =
    the level maps read once + the kept rows written."""
    from yolo_ad_refine_b200 import ops, parallel, synth
    from yolo_ad_refine_b200.postprocess import nms_raw
    world, rank, local = parallel.env_world()
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    parallel.init("nccl")
    B, N = 256, 8400
    raw = torch.from_numpy(synth.make_head_logits(B, N, seed=rank))
    host = [raw[:, :, :6400].reshape(B, 144, 80, 80).permute(0, 2, 3, 1).contiguous().bfloat16().pin_memory(),
            raw[:, :, 6400:8000].reshape(B, 144, 40, 40).permute(0, 2, 3, 1).contiguous().bfloat16().pin_memory(),
            raw[:, :, 8000:].reshape(B, 144, 20, 20).permute(0, 2, 3, 1).contiguous().bfloat16().pin_memory()]
    lv = [t.cuda() for t in host]
    levels = [ops.Act(t) for t in lv]
    y = torch.empty(B, 84, N, device="cuda")
    proj = torch.arange(16, dtype=torch.float32, device="cuda")
    l0 = ops.LAUNCHES

    def step():
        ops.decode(levels, (8, 16, 32), 80, 16, proj, y)
        return nms_raw(y, **NMS_ARGS)

    det, _, cnt = step()
    launches = ops.LAUNCHES - l0
    W = max(args.warmup, 3)

    def timed(fn, k):
        for _ in range(W):
            fn()
        parallel.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        e.record()
        parallel.barrier()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(s.elapsed_time(e), device="cuda") / k

    with ClockSampler(local) as clk:
        ms = timed(step, args.steps)
    ms_dec = timed(lambda: ops.decode(levels, (8, 16, 32), 80, 16, proj, y), args.steps)
    det_host = torch.empty((B, NMS_ARGS["max_det"], 6), dtype=torch.float32).pin_memory()
    cnt_host = torch.empty((B,), dtype=torch.int32).pin_memory()

    def e2e_step():
        for d, h in zip(lv, host):
            d.copy_(h, non_blocking=True)
        dd, _, cc = step()
        det_host.copy_(dd, non_blocking=True)
        cnt_host.copy_(cc, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_ms = timed(e2e_step, max(3, args.steps // 4))
    in_bytes = sum(t.numel() * 2 for t in host)
    alg = in_bytes + int(cnt.sum()) * 6 * 4
    pk = peaks()
    line = {"metric": "img/s decode+NMS microbench (8400 anchors x 80 classes, batch 256)", "value": world * B / (ms / 1000.0), "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[2]: fused DFL decode + batched NMS(conf .25, iou .7, max_det 300), batch 256/GPU, 8400 anchors, 80 classes, "
                                   "synthetic bf16 head logits", "global_batch": world * B, "l2": "620 MB of level maps per step exceed the 126 MB L2; no flush"},
            "e2e": {"value": world * B / (e2e_ms / 1000.0), "unit": "img/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": det_host.numel() * 4 + cnt_host.numel() * 4},
            "gpu_launches": launches * args.steps, "launches_per_step": launches, "clocks": clk.summary(),
            "detections_per_image": float(cnt.float().mean()), "candidates_per_image": float((y[:, 4:].amax(1) > NMS_ARGS["conf_thres"]).sum(1).float().mean()),
            "roofline": {"kernel": "decode_kernel + nms_* (yad_decode + yad_nms, whole step)", "bound": "hbm", "achieved": alg / (ms / 1000.0) / 1e9, "peak": pk["hbm"],
                         "unit": "GB/s", "frac": alg / (ms / 1000.0) / 1e9 / pk["hbm"], "traffic": None, "algorithmic_bytes": alg,
                         "decode_only_ms": ms_dec, "decode_only_GBps": (in_bytes + 84 * N * B * 4) / (ms_dec / 1000.0) / 1e9,
                         "peak_source": f"hbm_gbs of MEASURED_PEAKS.json ({pk['src']})"}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--in-flight", type=int, default=2, help="batches in flight (buffer sets with their own input, forward stream and graph)")
    ap.add_argument("--no-overlap", action="store_true", help="one batch in flight (forward graphs of consecutive batches on one stream)")
    ap.add_argument("--ref-batch", type=int, default=4, help="images per step of the CPU reference arm (a bounded sample of the workload)")
    ap.add_argument("--train-batch", type=int, default=128, help="images per GPU of the training-step measurement (0 = skip)")
    ap.add_argument("--train-steps", type=int, default=5)
    ap.add_argument("--profile-json", default=None, help="write the per-entry-point CUDA-event profile of one eager step here")
    ap.add_argument("--workload", default="inference", choices=["inference", "nms_micro"],
                    help="inference = BASELINE.json configs[1] (default; --imgsz 1280 --batch 4 = configs[4] per GPU); nms_micro = configs[2]: fused DFL "
                         "decode + batched NMS on synthetic head logits, batch 256 x 8400 anchors x 80 classes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "nms_micro":
        return run_nms_micro(args)

    from yolo_ad_refine_b200 import parallel
    world, rank, local = parallel.env_world()
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    parallel.init("nccl")
    from yolo_ad_refine_b200 import ops, synth
    from yolo_ad_refine_b200.engine import RefineEngine

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    W = max(args.warmup, 3)
    sd = synth.make_state_dict(seed=1)
    eng = RefineEngine(sd, batch=args.batch, imgsz=args.imgsz, dtype=dtype, nms_args=NMS_ARGS, input_u8=True, overlap_batches=False if args.no_overlap else args.in_flight)
    rs = np.random.RandomState(100 + rank)
    host_u8 = torch.from_numpy(rs.randint(0, 256, (args.batch, 3, args.imgsz, args.imgsz), dtype=np.uint8)).pin_memory()
    eng.fill_inputs(host_u8)
    eng.step()  # builds weights, warms the allocator, captures the graph
    torch.cuda.synchronize()

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        for _ in range(W):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(k):
            fn()
        eng.join()  # the NMS of the last batches runs on the engine's second stream: the closing event waits for it
        e.record()
        barrier()
        return parallel.max_over_ranks(s.elapsed_time(e), device="cuda")

    # ---- device-resident throughput (graph replay of forward + decode + NMS)
    with ClockSampler(local) as clk:
        ms = timed(eng.step, args.steps)
    ms_per_step = ms / args.steps
    value = world * args.batch / (ms_per_step / 1000.0)

    # ---- end to end through the public API: pinned host uint8 -> H2D -> hot path -> D2H of detections and counts
    det_host = torch.empty((args.batch, NMS_ARGS["max_det"], 6), dtype=torch.float32).pin_memory()
    cnt_host = torch.empty((args.batch,), dtype=torch.int32).pin_memory()

    def e2e_run(k):
        # public streaming API: every step copies its own pinned uint8 batch to the device and reads its detections back on the host;
        # the copy of step i+1 overlaps the compute of step i (engine.detect_many)
        n_det = 0
        for dh, ch in eng.detect_many((host_u8 for _ in range(k)), det_host, cnt_host):
            n_det += int(ch[0])
        return n_det

    e2e_run(W)
    barrier()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_ev.record()
    e2e_run(args.steps)
    e_ev.record()
    barrier()
    e2e_ms = parallel.max_over_ranks(s_ev.elapsed_time(e_ev), device="cuda") / args.steps
    e2e = {"value": world * args.batch / (e2e_ms / 1000.0), "unit": "img/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": host_u8.numel(), "d2h_bytes_per_step": det_host.numel() * 4 + cnt_host.numel() * 4}

    # ---- roofline of the dominant kernel: one instrumented eager step, CUDA events around every libyad entry point
    torch.cuda.synchronize()
    ops.PROFILE = {}
    # the host must stay AHEAD of the device here: an event pair brackets the C call, so host work inside the call (tensor-map encoding, ctypes)
    # would be billed to the kernel whenever the device has caught up.  A ~150 ms spin kernel is queued first and the whole step is enqueued behind it
    # (30 ms was not always enough: the second half of the 253 calls then carried 10 - 15 us of host time each and the fraction read 0.108 instead of 0.13).
    torch.cuda._sleep(int(300e6))
    eng._run()
    torch.cuda.synchronize()
    prof = {}
    for name, evs in ops.PROFILE.items():
        t = sum(s.elapsed_time(e) for s, e, _ in evs)
        fl = sum(m["flops"] for _, _, m in evs if m)
        prof[name] = {"calls": len(evs), "ms": t, "flops": fl}
        if name == "yad_conv2d":
            prof[name]["per_call"] = [dict(m, ms=s.elapsed_time(e)) for s, e, m in evs]
    conv_evs = ops.PROFILE.get("yad_conv2d", [])
    ops.PROFILE = None
    eager_ms = sum(p["ms"] for p in prof.values())
    conv = prof["yad_conv2d"]
    pk = peaks()
    ach = conv["flops"] / (conv["ms"] / 1000.0) / 1e12
    # DRAM bytes per launch of the same kernels: from the committed ncu capture of THIS build (profiles/r2_conv_traffic.json, written by
    # tools/conv_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); used only when it describes the same launches
    traffic, traffic_note = None, "no ncu capture for this configuration"
    tpath = os.path.join(ROOT, "profiles", "r2_conv_traffic.json")
    if os.path.exists(tpath) and args.batch == 64 and args.imgsz == 640:
        tj = json.load(open(tpath))
        if tj.get("launches") == conv["calls"]:
            traffic, traffic_note = tj["dram_bytes_per_launch"], f"profiles/r2_conv_traffic.json ({tj['launches']} launches, ncu)"
        else:
            traffic_note = f"profiles/r2_conv_traffic.json covers {tj.get('launches')} launches, this build issues {conv['calls']}: stale, not reported"
    roofline = {"kernel": "yad_conv2d = conv2_kernel (resident weights, haloed 3x3 patch) / conv_tma_kernel / conv_tc_kernel / conv_small_kernel "
                          "(tcgen05 implicit-GEMM convolution, all launches of one step)",
                "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                "traffic": traffic, "traffic_source": traffic_note, "algorithmic_bytes_per_launch": sum(m["bytes"] for _, _, m in conv_evs) / max(1, len(conv_evs)),
                "achieved_GBps_algorithmic": sum(m["bytes"] for _, _, m in conv_evs) / (conv["ms"] / 1000.0) / 1e9, "hbm_peak_GBps": pk["hbm"],
                "peak_source": f"bf16_tflops_sustained of MEASURED_PEAKS.json ({pk['src']})", "launches": conv["calls"],
                "avg_launch_us": 1000.0 * conv["ms"] / conv["calls"], "share_of_step": conv["ms"] / eager_ms,
                "whole_step_frac_of_roofline": (args.batch * FLOPS_PER_IMG_640 * (args.imgsz / 640.0) ** 2 / (ms_per_step / 1000.0) / 1e12) / pk["tf_sustained"]}
    if args.profile_json and rank == 0:
        json.dump({"eager_ms": eager_ms, "graph_ms_per_step": ms_per_step, "entries": prof}, open(args.profile_json, "w"), indent=1)

    launches_per_step = eng.launches_per_step
    eng.step()
    eng.join()
    _y, _, _det, _, _cnt = eng._out
    det_per_img = float(_cnt.float().mean())
    cand_per_img = float((_y[:, 4:].amax(1) > NMS_ARGS["conf_thres"]).sum(1).float().mean())
    train = train_strong = None
    if args.train_batch > 0:
        del eng  # the inference engine's graphs and buffers are not needed any more
        torch.cuda.empty_cache()
        train = measure_training(args, world, rank, sd, dtype, W, barrier)
        if world > 1 and args.train_batch % world == 0:
            # BASELINE.json configs[3] as the reference runs it: GLOBAL batch 128 split over the ranks (engine/trainer.py:290) -- strong scaling
            train_strong = measure_training(args, world, rank, sd, dtype, W, barrier, B=args.train_batch // world, scaling="strong")

    line = {"metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": make_config(args, world),
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
            "clocks": clk.summary(), "detections_per_image": det_per_img, "candidates_per_image": cand_per_img, "roofline": roofline}
    if train is not None:
        line["train"] = train
    if train_strong is not None:
        line["train_strong"] = train_strong
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            sec, cores, kind = cpu_steps(args.ref_batch, args.imgsz, 4, 1)
            line["cpu_baseline"] = {"value": args.ref_batch / sec, "unit": "img/s", "cores": cores, "kind": kind,
                                    "sample": cpu_sample_text(kind, args.ref_batch, args.imgsz) + ", 4 timed steps after 1 warm-up"}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
