#!/usr/bin/env python
"""bench.py — headline benchmark of the detection post-processing + RoIAlign hot path.

    python bench.py --gpus N --steps K --warmup W            (our arm, one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   (reference CPU arm, rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic inputs of BASELINE.json's
config 2 (Faster R-CNN R50-FPN, batch 4, 1024^2 images resized to 800^2, 1000 RPN proposals per
image, 256-channel pyramid, 7x7 RoIAlign, 300 detections per image): RPN post-head stage ->
MultiScaleRoIAlign -> detection post-processing -> score filter + crop extraction. The CNN parts
(backbone, RPN head, box head) are not on this path; their outputs are seeded random tensors.

  value  detections/s (detections that pass miso's score filter and are cropped), inputs resident
         in HBM, CUDA-event timed, max over ranks.
  e2e    the same metric with HOST (pinned) inputs: per step H2D of every input, the hot path,
         D2H of the detections and crop bytes.
  roofline  RoIAlign kernel: algorithmic bytes / mean CUDA-event duration of its
         launches inside the timed region, against the measured HBM peak.
  cpu_baseline  the CPU oracle port of the same path on one image of the batch (rank 0, N=1).
With N > 1 every rank runs the same per-GPU batch as a shard of mosaic tiles (weak scaling); the
step then also contains the path's one exchange: an NCCL all-gather of the per-rank detection
blocks followed by the cross-tile seam NMS.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "detections_per_sec_posthead_path"
UNIT = "detections/s"


# ----------------------------------------------------------------------------------------------
def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--fast-roi-align", action="store_true", help="FMA RoIAlign (<=1e-5) instead of the bit-exact order")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--serial", action="store_true", help="skip the additional three-batches-in-flight measurement")
    ap.add_argument("--features-layout", default="channels_last", choices=["channels_last", "nchw"],
                    help="memory format of the synthetic FPN maps: channels_last = what a torch.channels_last "
                         "cuDNN backbone produces (gathered in place); nchw = the reference's default layout")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
print("max", nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), flush=True)
while True:
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksEventReasons(h), flush=True)
    time.sleep(0.001)
"""


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: a helper process polls NVML every ~1 ms
    (its own interpreter, so it neither holds this process's GIL nor delays the launch loop); only the
    samples stamped inside [t0, t1] are kept."""

    def __init__(self, index: int):
        import tempfile
        self.out = tempfile.TemporaryFile(mode="w+")
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdout=self.out,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def summary(self, t0: float, t1: float):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.seek(0)
        mx, rows, before = None, [], None
        for line in self.out.read().splitlines():
            f = line.split()
            if len(f) == 2 and f[0] == "max":
                mx = int(f[1])
            elif len(f) == 3:
                ts, mhz, mask = float(f[0]), int(f[1]), int(f[2])
                if t0 <= ts <= t1:
                    rows.append((mhz, mask))
                elif ts < t0:
                    before = (mhz, mask)
        if not rows and before is not None:
            rows = [before]
        sm = sorted(r[0] for r in rows)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        reasons = sorted({n for _, m in rows for n, b in bits.items() if m & b})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(rows)}


def roi_align_algorithmic_bytes(proposals, counts, shapes, thresholds, scales, channels, pooled, sr=2):
    """20*K + 4*C*P^2*K + 4*C*sum_levels |union of touched pixels|, the union
    rasterised exactly from the sample coordinates of the reference formula (Appendix B.2)."""
    import numpy as np
    F = np.float32
    n = proposals.shape[0]
    maps = [np.zeros((n, gh, gw), dtype=bool) for gh, gw in shapes.feature_grids]
    k_live = 0
    for i in range(n):
        b = proposals[i, :counts[i]]
        k_live += len(b)
        area = ((b[:, 2] - b[:, 0]).astype(F) * (b[:, 3] - b[:, 1]).astype(F)).astype(F)
        lvl = np.zeros(len(b), dtype=np.int64)
        for t in thresholds:
            lvl += area >= F(t)
        for j in range(len(b)):
            l = lvl[j]
            gh, gw = shapes.feature_grids[l]
            sc = F(scales[l])
            sets = []
            for lo_c, hi_c, size in ((b[j, 1], b[j, 3], gh), (b[j, 0], b[j, 2], gw)):
                s0, e0 = F(lo_c * sc), F(hi_c * sc)
                r = max(F(e0 - s0), F(1.0))
                binsz = F(r / F(pooled))
                pp = np.repeat(np.arange(pooled, dtype=F), sr)
                ii = np.tile(np.arange(sr, dtype=F), pooled)
                v = ((s0 + (pp * binsz).astype(F)).astype(F) + (((ii + F(0.5)).astype(F) * binsz).astype(F) / F(sr)).astype(F)).astype(F)
                ok = ~((v < -1.0) | (v > size))
                v = np.maximum(v[ok], 0)
                lo = np.minimum(v.astype(np.int64), size - 1)
                hi = np.minimum(lo + 1, size - 1)
                sets.append(np.unique(np.concatenate([lo, hi])))
            if len(sets[0]) and len(sets[1]):
                maps[l][i][np.ix_(sets[0], sets[1])] = True
    touched = sum(int(m.sum()) for m in maps)
    return 20 * k_live + 4 * channels * pooled * pooled * k_live + 4 * channels * touched, k_live, touched


# ----------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path: torchvision 0.26 CPU ops and modules
    (filter_proposals, MultiScaleRoIAlign, postprocess_detections, resize_boxes) plus miso's score
    filter / coords_int / crop slice (restated: miso.* needs lxml/skimage, absent in this image).
    Bounded sample: one image of the batch per step. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    import torch
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    from torchvision.models.detection.image_list import ImageList
    from torchvision.models.detection.roi_heads import RoIHeads
    from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead, concat_box_prediction_layers
    from torchvision.models.detection.transform import resize_boxes
    from torchvision.ops.poolers import MultiScaleRoIAlign
    from miso_b200 import workload
    from oracle import miso_path as M

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = workload.faster_rcnn_batch(num_images=1, seed=0, pin=False)
    h = w.host
    ag = AnchorGenerator(w.rpn.sizes, w.rpn.aspect_ratios)
    rpn = RegionProposalNetwork(ag, RPNHead(8, 3), 0.7, 0.3, 256, 0.5, dict(training=2000, testing=w.rpn.pre_nms_top_n),
                                dict(training=2000, testing=w.rpn.post_nms_top_n), w.rpn.nms_thresh).eval()
    heads = RoIHeads(None, None, None, 0.5, 0.5, 512, 0.25, (10.0, 10.0, 5.0, 5.0), w.det.score_thresh,
                     w.det.nms_thresh, w.det.detections_per_img)
    pool = MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    il = ImageList(torch.zeros(1, 3, *w.shapes.padded_image_size), list(w.shapes.image_sizes))
    img = h["images"][0].numpy()

    def step():
        with torch.inference_mode():
            anchors = ag(il, h["objectness"])
            napl = [o.shape[1] * o.shape[2] * o.shape[3] for o in h["objectness"]]
            objectness, deltas = concat_box_prediction_layers(list(h["objectness"]), list(h["deltas"]))
            proposals = rpn.box_coder.decode(deltas, anchors).view(1, -1, 4)
            boxes, _ = rpn.filter_proposals(proposals, objectness, il.image_sizes, napl)
            feats = pool({str(i): f for i, f in enumerate(h["features"])}, boxes, il.image_sizes)
            k = boxes[0].shape[0]
            db, ds, dl = heads.postprocess_detections(h["class_logits"][0][:k], h["box_regression"][0][:k], boxes, il.image_sizes)
            rb = resize_boxes(db[0], list(w.shapes.image_sizes[0]), list(w.shapes.original_image_sizes[0]))
            _, _, _, _, _, crops = M.filter_and_crop(img, rb.numpy(), ds[0].numpy(), dl[0].numpy(), w.threshold)
        return len(crops), feats.shape[0]

    for _ in range(min(args.warmup, 2)):
        step()
    t0 = time.perf_counter()
    dets = 0
    for _ in range(args.steps):
        d, _ = step()
        dets += d
    dt = time.perf_counter() - t0
    value = dets / dt
    sample = "1 image of the batch per step (torchvision CPU ops + miso filter/crop restated in numpy)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name.replace(f"batch {w.shapes.num_images}", "batch 4 (sampled: 1 image/step)")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
def cpu_baseline_port(w_full, steps=2):
    """The CPU oracle port of the same path on ONE image of the batch (bounded sample)."""
    import numpy as np
    from oracle import native
    from oracle import pipeline as ref_pipeline
    from miso_b200 import workload
    native.lib()
    w = workload.faster_rcnn_batch(num_images=1, seed=0, pin=False)
    h = {k: [t.numpy() for t in v] for k, v in w.host.items()}
    kw = dict(padded_image_size=w.shapes.padded_image_size, image_sizes=w.shapes.image_sizes,
              original_image_sizes=w.shapes.original_image_sizes, sizes=w.rpn.sizes, aspect_ratios=w.rpn.aspect_ratios,
              pre_nms_top_n=w.rpn.pre_nms_top_n, post_nms_top_n=w.rpn.post_nms_top_n,
              detections_per_img=w.det.detections_per_img, threshold=w.threshold)
    args_ = (h["objectness"], h["deltas"], h["features"], h["class_logits"][0], h["box_regression"][0], h["images"])
    ref_pipeline.run(*args_, **kw)
    t0 = time.perf_counter()
    dets = 0
    tm = {}
    for _ in range(steps):
        out = ref_pipeline.run(*args_, timings=tm, **kw)
        dets += sum(len(o["crops"]) for o in out)
    dt = time.perf_counter() - t0
    return {"value": dets / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"1 image of the batch, {steps} passes of the oracle (C NMS/RoIAlign with OpenMP + numpy); "
                      f"last pass: rpn {tm.get('rpn_s', 0):.3f}s roi_align {tm.get('roi_align_s', 0):.3f}s "
                      f"det+crop {tm.get('det_crop_s', 0):.3f}s"}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from miso_b200 import mosaic, pipeline, workload

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None     # started early: it must be polling before the timed region
    w = workload.faster_rcnn_batch(num_images=args.batch, seed=rank, features_layout=args.features_layout)
    hp = pipeline.HotPath(w.shapes, w.rpn, w.det, threshold=w.threshold, crop_capacity_bytes=512 << 20,
                          exact_roi_align=not args.fast_roi_align, device=dev)
    d = workload.to_device(w, dev)
    hp.bind(d["objectness"], d["deltas"], d["features"], d["class_logits"][0], d["box_regression"][0], d["images"])
    n, dpi = args.batch, hp.dpi
    # mosaic framing for N > 1: this rank's images are tiles at (row=rank, col=i) of a tile grid with 128 px overlap
    origins = torch.tensor([[rank * 896.0, i * 896.0] for i in range(n)], dtype=torch.float32, device=dev)

    # The timed region runs one batch at a time (rpn -> roi_align -> detections -> crops, event-chained), so that
    # the RoIAlign launch durations and their share of the step are those of the kernel running alone. A second
    # measurement ("pipelined") keeps three batches in flight on three streams (pipeline.OverlappedHotPath): the
    # few-CTA kernel chains of the RPN and detection stages then run beside another batch's RoIAlign.
    hps = [hp]
    if not args.serial:
        for _ in range(2):
            h2 = pipeline.HotPath(w.shapes, w.rpn, w.det, threshold=w.threshold, crop_capacity_bytes=64 << 20,
                                  exact_roi_align=not args.fast_roi_align, device=dev)
            h2.bind(d["objectness"], d["deltas"], d["features"], d["class_logits"][0], d["box_regression"][0], d["images"])
            hps.append(h2)
    plan1 = pipeline.OverlappedHotPath(hps[:1])
    plan3 = pipeline.OverlappedHotPath(hps) if len(hps) > 1 else None

    # Exchange plumbing (N > 1): the block is packed on the detection stream right behind mb_det_postprocess (one
    # 3 us kernel, three rotating buffers), so the next batch never waits for the exchange; all-gather + seam NMS
    # alternate between two communication streams with their own buffers — the seam NMS is a latency chain on a
    # couple of CTAs, two of them in flight cost nothing and double its throughput.
    NBLK, NCOMM = 3, 2
    comms = [torch.cuda.Stream(device=dev) for _ in range(NCOMM)] if world > 1 else []
    seams = [mosaic.SeamNms(world * n * dpi, w.shapes.num_classes, dev) for _ in range(NCOMM)] if world > 1 else []
    x_block = [torch.empty((n * dpi, 6), dtype=torch.float32, device=dev) for _ in range(NBLK)]
    x_gathered = [torch.empty((world * n * dpi, 6), dtype=torch.float32, device=dev) for _ in range(NCOMM)]
    ev_packed = [torch.cuda.Event() for _ in range(NBLK)]
    ev_gathered = [None] * NBLK           # all-gather that last read x_block[b] has completed
    xstep = [0]
    roi_ev = []           # (start, end) CUDA events around every RoIAlign launch of the timed region, on its stream
    recording = [False]

    def before_roi(i, hp_i, st):
        if recording[0]:
            e = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            roi_ev.append(e)
            e[0].record(st)

    def after_roi(i, hp_i, st):
        if recording[0]:
            roi_ev[-1][1].record(st)

    def after_det(i, hp_i, st):
        """The path's one exchange: fixed-size block -> one all_gather_into_tensor -> seam NMS right behind it,
        no host sync anywhere."""
        if world == 1:
            return
        k_, b_, c_ = xstep[0], xstep[0] % NBLK, xstep[0] % NCOMM
        xstep[0] += 1
        if ev_gathered[b_] is not None:
            st.wait_event(ev_gathered[b_])           # (three steps old: long done)
        with torch.cuda.stream(st):
            mosaic.pack_block(hp_i.det_boxes, hp_i.det_scores, hp_i.det_labels, hp_i.det_counts, origins,
                              w.threshold, n * dpi, out=x_block[b_])
        ev_packed[b_].record(st)
        with torch.cuda.stream(comms[c_]):
            comms[c_].wait_event(ev_packed[b_])
            g_ = mosaic.exchange(x_block[b_], world, out=x_gathered[c_])
            if ev_gathered[b_] is None:
                ev_gathered[b_] = torch.cuda.Event()
            ev_gathered[b_].record(comms[c_])
            if k_ % world == rank:                   # the gathered rows are identical on every rank: the seam NMS of
                seams[c_].launch(g_, w.det.nms_thresh)   # step k runs once, on rank k mod N (not replicated N times)

    for pl in (plan1, plan3):
        if pl is not None:
            pl.hooks.update(before_roi=before_roi, after_roi=after_roi, after_det=after_det)
    host_enqueue = [0.0]

    def run_steps(plan, k):
        t_h = time.perf_counter()
        for _ in range(k):
            plan.submit()
        host_enqueue[0] = (time.perf_counter() - t_h) / k
        plan.drain()
        if world > 1:
            for c_ in comms:
                torch.cuda.current_stream(dev).wait_stream(c_)   # the region ends with the last seam NMS

    def timed(plan, k):
        """K batches submitted and completed between two events on the current stream; max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        run_steps(plan, k)
        e1.record()
        barrier()
        t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms[0])

    run_steps(plan1, max(args.warmup, 3))
    if plan3 is not None:
        run_steps(plan3, max(args.warmup, 3))
    barrier()
    tot = hp.crop_totals.tolist()
    if tot[2]:
        raise SystemExit("crop buffer overflow")
    dets_per_step = tot[0]
    crop_bytes = tot[1]
    for h2 in hps[1:]:
        assert h2.crop_totals.tolist() == tot

    # ---- stage breakdown from a short pass with events between the stages (not the timed region) ----
    sev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(5)]
    for es in sev:
        es[0].record(); hp.rpn(); es[1].record(); hp.roi_align(); es[2].record(); hp.detections(); es[3].record()
        hp.crops(); es[4].record()
    barrier()
    names = ("rpn", "roi_align", "det_postprocess", "filter_crop")
    serial_stage_ms = {nm: sorted(es[j].elapsed_time(es[j + 1]) for es in sev)[len(sev) // 2] for j, nm in enumerate(names)}

    # ---- timed region: K steps, device timed; every RoIAlign launch bracketed by events on its stream ----
    wall0 = time.time()
    recording[0] = True
    ms = timed(plan1, args.steps)
    recording[0] = False
    clocks = sampler.summary(wall0, time.time()) if rank == 0 else None
    host_ms = 1e3 * host_enqueue[0]
    t = torch.tensor([float(dets_per_step)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total_dets = float(t[0])
    ms_per_step = ms / args.steps
    value = total_dets / (ms_per_step * 1e-3)
    roi_ms = sorted(e[0].elapsed_time(e[1]) for e in roi_ev)
    roi_mean_ms = sum(roi_ms) / len(roi_ms)

    # ---- the RoIAlign launch in the other arithmetic mode (same proposals), for reference ----
    other_exact = int(bool(args.fast_roi_align))
    hp.roi_params.exact = other_exact
    oev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a_, b_ in oev:
        hp.rpn(); a_.record(); hp.roi_align(); b_.record(); hp.detections()
    hp.roi_params.exact = 1 - other_exact
    hp.step()
    barrier()
    other_ms = sorted(a_.elapsed_time(b_) for a_, b_ in oev)[len(oev) // 2]

    # ---- the same step replayed from a CUDA graph (one host call per step, no per-launch gaps) ----
    graph_ms = None
    if world == 1:
        hp.capture()
        for _ in range(3):
            hp.replay()
        ge0, ge1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ge0.record()
        for _ in range(args.steps):
            hp.replay()
        ge1.record()
        barrier()
        graph_ms = ge0.elapsed_time(ge1) / args.steps

    # ---- the same K batches, three in flight ----
    pipelined = None
    if plan3 is not None:
        k3 = 3 * max(args.steps // 3, 1)
        ms3 = timed(plan3, k3) / k3
        pipelined = {"value": total_dets / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3, "steps": k3, "batches_in_flight": 3,
                     "note": "rpn(k+2) | roi_align(k+1) | detections+crops(k) on three streams, per-batch stage order kept by events"}

    # ---- e2e: host (pinned) inputs, H2D + path + D2H of results every step, through the public host-facing
    # API (pipeline.HostPipeline: copy-in / compute / copy-out streams, two slots) ----
    e2e_seams = {}

    def make_hp():
        return pipeline.HotPath(w.shapes, w.rpn, w.det, threshold=w.threshold, crop_capacity_bytes=128 << 20,
                                exact_roi_align=not args.fast_roi_align, device=dev)

    def on_computed(hp_s, slot):
        if world == 1:
            return
        sm_ = e2e_seams.setdefault(id(slot), mosaic.SeamNms(world * n * dpi, w.shapes.num_classes, dev))
        block = mosaic.pack_block(hp_s.det_boxes, hp_s.det_scores, hp_s.det_labels, hp_s.det_counts, origins, w.threshold, n * dpi)
        sm_.launch(mosaic.exchange(block, world), w.det.nms_thresh)
        slot["seam_done"] = torch.cuda.Event(); slot["seam_done"].record()

    def on_collect(hp_s, slot):
        if world == 1:
            return None
        slot["seam_done"].synchronize()
        b, s_, l = e2e_seams[id(slot)].finish()
        return b.cpu(), s_.cpu(), l.cpu()

    pipe = pipeline.HostPipeline(make_hp, w.host, depth=2, on_computed=on_computed, on_collect=on_collect)
    d2h_small = sum(t_.numel() * t_.element_size() for t_ in pipe.slots[0]["out"].values())

    def e2e_run(k):
        got = 0
        for _ in range(k):
            r = pipe.submit(w.host)
            got += r is not None
        got += len(pipe.flush())
        assert got == k
        return r

    e2e_steps = max(3, min(args.steps, 20))
    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)          # every batch: H2D of all inputs, the path, D2H of detections + crops; drained at the end
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_dets / float(te[0])
    h2d = w.input_bytes()
    d2h = d2h_small + crop_bytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (RoIAlign) ----
    peak, peak_src = peaks()
    props = hp.proposals.cpu().numpy()
    cnts = hp.prop_counts.cpu().numpy()
    q = hp.roi_params
    thr = [q.level_thresholds[i] for i in range(q.num_levels - 1)]
    scl = [q.spatial_scale[i] for i in range(q.num_levels)]
    alg_bytes, k_live, touched = roi_align_algorithmic_bytes(props, cnts, w.shapes, thr, scl, w.shapes.channels, w.shapes.pooled)
    achieved = alg_bytes / (roi_mean_ms * 1e-3) / 1e9
    nchw_route = "k_nchw_to_nhwc+k_roi_align_nhwc4d" if hp.roi_ws is not None else "k_roi_align_sr2"
    roi_kernel = "k_roi_align_nhwc4d" if hp.features_layout == "channels_last" else nchw_route
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roi_align_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(roi_kernel, {}).get("dram_bytes_per_launch")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_port(w)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": w.name, "per_gpu_batch": args.batch, "features_layout": hp.features_layout, "l2": "inputs larger than L2 (218 MB pyramid + 201 MB RoIAlign output per step)",
                   "roi_align_mode": "fast(fma)" if args.fast_roi_align else "exact(reference op order)",
                   "multi_gpu": "per-rank batch = shard of mosaic tiles; NCCL all_gather every step on every rank, the seam NMS of step k on rank k mod N; two communication streams overlap the following batches" if world > 1 else "single GPU",
                   "detections_per_step": total_dets, "crop_bytes_per_step": crop_bytes,
                   "images_per_s": world * args.batch / (ms_per_step * 1e-3)},
        "roofline": {"bound": "hbm", "kernel": roi_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                     "kernel_ms_mean": roi_mean_ms, "kernel_ms_min": roi_ms[0], "rois": k_live, "touched_pixels": touched,
                     "kernel_share_of_step": roi_mean_ms / ms_per_step,
                     "limiter": "L2->SM bandwidth (1.4 GB per launch of L1 misses at ~10 TB/s), see DESIGN.md section 7",
                     "other_mode": {"mode": "exact" if other_exact else "fast(fma)", "kernel_ms": other_ms,
                                    "frac": alg_bytes / (other_ms * 1e-3) / 1e9 / peak}},
        "stage_ms": serial_stage_ms, "host_enqueue_ms_per_step": host_ms, "pipelined": pipelined,
        "graph_replay": None if graph_ms is None else {"ms_per_step": graph_ms, "value": total_dets / (graph_ms * 1e-3), "unit": UNIT},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * float(te[0]), "steps": e2e_steps,
                "api": "miso_b200.pipeline.HostPipeline (pinned host in/out, copy-in | compute | copy-out streams, 2 batches in flight)"},
        "gpu_launches": hp.kernel_launches_per_step * args.steps,
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
