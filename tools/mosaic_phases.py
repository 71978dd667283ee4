"""Phase breakdown of one rank's share of the 16384^2 mosaic (config 5, world 8 emulated on one GPU): model, pack, seam NMS, crops."""
import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from miso.object_detection.models import get_object_detection_model
from miso_b200 import mosaic, detection
from miso_b200.patch import patch_model, forward_uint8
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = get_object_detection_model(3).eval().to(dev)
with torch.no_grad():
    model.roi_heads.box_predictor.cls_score.weight.mul_(8.0)
patch_model(model)
S = 16384
g = torch.Generator(device=dev).manual_seed(0)
mos = torch.randint(0, 256, (S, S, 3), dtype=torch.uint8, device=dev, generator=g)
grid = mosaic.tile_grid(S, S, 1024, 128)
world = 8; rank = 0
mine = list(mosaic.rank_tiles(len(grid), world, rank))
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = T()
    dpi = 300
    boxes = torch.zeros((len(mine), dpi, 4), device=dev); scores = torch.zeros((len(mine), dpi), device=dev)
    labels = torch.zeros((len(mine), dpi), dtype=torch.int64, device=dev); counts = torch.zeros((len(mine),), dtype=torch.int32, device=dev)
    with torch.inference_mode():
        for i0 in range(0, len(mine), 4):
            idx = mine[i0:i0 + 4]
            tiles = [mos[grid[t][0]:grid[t][0] + 1024, grid[t][1]:grid[t][1] + 1024] for t in idx]
            res = forward_uint8(model, tiles)
            for j, r in enumerate(res):
                k = int(r["boxes"].shape[0]); boxes[i0 + j, :k], scores[i0 + j, :k], labels[i0 + j, :k] = r["boxes"], r["scores"], r["labels"]; counts[i0 + j] = k
    t1 = T()
    # emulate the gathered block of 8 ranks by tiling this rank's block
    origins = torch.tensor([[float(grid[t][0]), float(grid[t][1])] for t in mine], dtype=torch.float32, device=dev)
    tmax = mosaic.tiles_per_rank_max(len(grid), world)
    block = mosaic.pack_block(boxes, scores, labels, counts, origins, 0.5, tmax * dpi)
    gathered = block.repeat(world, 1)
    gathered[:, 0] += torch.arange(gathered.shape[0], device=dev) // block.shape[0] * 20000.0   # keep ranks apart
    gathered[:, 2] += torch.arange(gathered.shape[0], device=dev) // block.shape[0] * 20000.0
    t2 = T()
    seam = mosaic.SeamNms(gathered.shape[0], 3, dev)
    t3 = T()
    seam.launch(gathered, 0.5); fb, fs, fl = seam.finish()
    t4 = T()
    share = torch.arange(rank, fb.shape[0], world, device=dev)
    sb = fb[share]; sb[:, 0::2] = sb[:, 0::2] % 16000
    oc = detection.filter_and_crop([mos], sb[None].contiguous(), torch.ones((1, sb.shape[0]), device=dev), torch.tensor([sb.shape[0]], dtype=torch.int32, device=dev), 0.5)
    t5 = T()
    print(f"tiles {len(mine)}: model {1e3*(t1-t0):.1f} ms | pack+emulated gather {1e3*(t2-t1):.1f} | SeamNms alloc {1e3*(t3-t2):.1f} | seam nms {1e3*(t4-t3):.1f} | crops {1e3*(t5-t4):.1f} ms ({int(oc.totals[1])/1e6:.0f} MB)")
