"""BASELINE config 5: a synthetic S x S mosaic (default 16384), 1024 px tiles with 128 px overlap, sharded over the
ranks of one box; patched Faster R-CNN R50-FPN (random init) per tile, NCCL all-gather, seam NMS, crops.
    python tools/mosaic_prof.py [S]                                  (1 GPU)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mosaic_prof.py [S]"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso.object_detection.models import get_object_detection_model  # noqa: E402
from miso_b200 import mosaic  # noqa: E402
from miso_b200.patch import patch_model  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
model = get_object_detection_model(3).eval().to(dev)
with torch.no_grad():
    model.roi_heads.box_predictor.cls_score.weight.mul_(8.0)
patch_model(model)
g = torch.Generator(device=dev).manual_seed(0)
mos = torch.randint(0, 256, (S, S, 3), dtype=torch.uint8, device=dev, generator=g)
tiles = len(mosaic.tile_grid(S, S, 1024, 128))


def run():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    b, s, l, crops = mosaic.infer_mosaic(model, mos, tile=1024, overlap=128, threshold=0.5, batch_size=4, rank=rank, world=world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.perf_counter() - t0, b, crops


run()                                # warm-up (cuDNN autotune, lazy init)
dt, b, crops = run()
if rank == 0:
    nb = 0 if crops is None else int(crops.totals[1])
    print(f"mosaic {S}x{S}: {tiles} tiles over {world} GPU(s): {dt * 1e3:.1f} ms ({tiles / dt:.0f} tiles/s), "
          f"{b.shape[0]} detections after the seam NMS, rank-0 crop bytes {nb}")
if world > 1:
    dist.destroy_process_group()
