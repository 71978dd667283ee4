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
grid = mosaic.tile_grid(S, S, 1024, 128)
tiles = len(grid)
y0, y1 = mosaic.rank_band(grid, 1024, S, world, rank)            # only this rank's pixel rows live on its GPU
g = torch.Generator(device=dev).manual_seed(1 + rank)
band = torch.randint(0, 256, (y1 - y0, S, 3), dtype=torch.uint8, device=dev, generator=g)


def run():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = mosaic.infer_mosaic(model, band, (S, S), band_y0=y0, tile=1024, overlap=128, threshold=0.5, batch_size=4,
                              rank=rank, world=world, crop_capacity_bytes=16 << 30)      # random-init weights on noise: large boxes
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.perf_counter() - t0, out


run()                                # warm-up (cuDNN autotune, lazy init)
dt, out = run()
if rank == 0:
    kept = int((out["state"] == 1).sum())
    print(f"mosaic {S}x{S}: {tiles} tiles over {world} GPU(s): {dt * 1e3:.1f} ms ({tiles / dt:.0f} tiles/s) including the "
          f"backbone, {kept} detections after the seam NMS, rank-0 crops {out['crops']['count']} ({out['crops']['bytes']} bytes)")
if world > 1:
    dist.destroy_process_group()
