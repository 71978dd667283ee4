"""Decompose the TMA RoIAlign kernel's time on the bench input: MB_TMA_PROBE bit 0 = no copies, bit 1 = no
arithmetic, bit 2 = no output stores (each run is a fresh process: the probe value is read once)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ROOT)
    from miso_b200 import pipeline, workload
    dev = torch.device("cuda:0")
    w = workload.faster_rcnn_batch(num_images=4, seed=0, features_layout="channels_last")
    hp = pipeline.HotPath(w.shapes, w.rpn, w.det, threshold=w.threshold, crop_capacity_bytes=64 << 20, device=dev)
    d = workload.to_device(w, dev)
    hp.bind(d["objectness"], d["deltas"], d["features"], d["class_logits"][0], d["box_regression"][0], d["images"])
    hp.rpn()
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    for _ in range(3):
        hp.roi_align()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(); hp.roi_align(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    print(json.dumps({"probe": os.environ.get("MB_TMA_PROBE", "0"), "median_ms": t[len(t) // 2], "min_ms": t[0]}))
else:
    for probe in sys.argv[1:] or ["0", "1", "2", "4", "3", "6", "7"]:
        env = dict(os.environ, MB_TMA_PROBE=probe)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
