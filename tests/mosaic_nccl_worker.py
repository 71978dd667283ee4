"""2-rank NCCL run of the small mosaic plan; every rank checks its share against its own world-size-1 run.
Launched by tests/test_gpu_mosaic.py through torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
import tests.test_gpu_mosaic as T  # noqa: E402
from miso_b200 import mosaic  # noqa: E402

T.DEV = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = T.small_mosaic()
p1, b1, grid, _ = T.make_plan(w, 0, 1)
p1.run(b1)
r1 = p1.results()
dpi = p1.dpi
plan, batches, _, mine = T.make_plan(w, rank, world)
plan.run(batches)
res = plan.results()
to_w = np.array([mosaic.gathered_row(r // dpi, r % dpi, len(grid), world, dpi) for r in range(len(grid) * dpi)])
assert np.array_equal(plan.gathered.cpu().numpy()[to_w], p1.gathered.cpu().numpy())
keep1 = np.nonzero(p1.seam.state.cpu().numpy() == 1)[0]
assert np.array_equal(np.nonzero(plan.seam.state.cpu().numpy() == 1)[0], np.sort(to_w[keep1]))
src1 = r1["src"].cpu().numpy()
own = (src1 // dpi >= mine[0]) & (src1 // dpi <= mine[-1])
assert res["count"] == int(own.sum())
assert np.array_equal(res["rects"].cpu().numpy(), r1["rects"].cpu().numpy()[own])
offs1 = r1["offsets"].cpu().numpy()
pix1 = r1["pixels"].cpu().numpy()
want = np.concatenate([pix1[offs1[j]:offs1[j + 1]] for j in np.nonzero(own)[0]]) if own.any() else np.zeros(0, np.uint8)
assert np.array_equal(res["pixels"].cpu().numpy(), want)
dist.barrier()
if rank == 0:
    print("world 2 == world 1: ok")
dist.destroy_process_group()
