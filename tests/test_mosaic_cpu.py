"""Host logic of the tiled-mosaic path without a GPU: tile grid, contiguous rank partition, row mapping, pixel
bands, the one collective (gloo, world_size 2). Packing / seam NMS run here through the CPU restatement of
tests/mosaic_ref.py (the product has no CPU branch); the GPU tests run the same checks through libmisob200."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from miso_b200 import mosaic
from oracle import detection as D
from tests import mosaic_ref as R


def test_tile_grid_matches_config5():
    starts = mosaic.tile_starts(16384, 1024, 128)
    assert len(starts) == 19 and starts[0] == 0 and starts[1] == 896 and starts[-1] == 16384 - 1024
    assert all(b - a <= 896 for a, b in zip(starts, starts[1:]))
    grid = mosaic.tile_grid(16384, 16384, 1024, 128)
    assert len(grid) == 361 and grid[1] == (0, 896) and grid[19] == (896, 0)
    assert mosaic.tile_starts(1000, 1024, 128) == [0]
    assert mosaic.tile_starts(1920, 1024, 128) == [0, 896]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_rank_partition_is_contiguous_and_balanced(world):
    parts = [mosaic.rank_tiles(361, world, r) for r in range(world)]
    assert [i for p in parts for i in p] == list(range(361))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1 and max(sizes) == mosaic.tiles_per_rank_max(361, world)


def synth_tiles(num_tiles=5, dpi=40, seed=0):
    """Detections of every tile (tile-local coordinates) on a 2-column grid with 128 px overlap."""
    rng = np.random.default_rng(seed)
    boxes = np.zeros((num_tiles, dpi, 4), np.float32); scores = np.zeros((num_tiles, dpi), np.float32)
    labels = np.zeros((num_tiles, dpi), np.int64); counts = np.zeros(num_tiles, np.int32)
    origins = np.array([[(t // 2) * 896.0, (t % 2) * 896.0] for t in range(num_tiles)], np.float32)
    for t in range(num_tiles):
        k = int(rng.integers(dpi // 2, dpi + 1))
        c = rng.uniform(0, 1024, (k, 2)); s = rng.uniform(20, 200, (k, 2))
        boxes[t, :k] = np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 1024)
        scores[t, :k] = np.sort(rng.uniform(0.2, 1.0, k))[::-1]
        labels[t, :k] = rng.integers(1, 3, k); counts[t] = k
    # duplicate a few objects across the seam of tiles 0|1 so that the seam NMS has work to do
    boxes[1, :5] = boxes[0, :5] + np.array([-896, 0, -896, 0], np.float32)
    labels[1, :5] = labels[0, :5]
    return boxes, scores, labels, counts, origins


def single_process(thr=0.5, iou=0.5):
    b, s, l, c, o = synth_tiles()
    T, dpi = s.shape
    block = R.pack_block(b, s, l, c, o, thr, T * dpi)
    return block, R.seam_keep_rows(block, iou)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, s, l, c, o = synth_tiles()
    T, dpi = s.shape
    mine = list(mosaic.rank_tiles(T, world, rank))
    rows = mosaic.tiles_per_rank_max(T, world) * dpi
    block = torch.from_numpy(R.pack_block(b[mine], s[mine], l[mine], c[mine], o[mine], 0.5, rows))
    gathered = mosaic.exchange(block, world).numpy()          # the product's collective call, gloo backend
    keep = R.seam_keep_rows(gathered, 0.5)
    q.put((rank, gathered, keep))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_equals_world_size_1():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    block1, keep1 = single_process()
    T, dpi = synth_tiles()[1].shape
    assert len(keep1) < int((block1[:, 5] >= 0).sum())        # the seam duplicates were suppressed
    # world-1 row (tile*dpi + slot) -> row in the 2-rank gathered buffer
    to2 = np.array([mosaic.gathered_row(r // dpi, r % dpi, T, 2, dpi) for r in range(T * dpi)])
    for _, gathered, keep in res:                             # replicated, deterministic result on every rank
        assert np.array_equal(gathered[to2], block1)
        assert np.array_equal(keep, np.sort(to2[keep1]))


def test_seam_nms_matches_reference_composition():
    """single-process definition: per tile detections -> +origin -> concat ->
    _batched_nms_vanilla -> score filter; the filter commutes with NMS."""
    from torchvision.ops import boxes as tvb
    b, s, l, c, o = synth_tiles()
    T, dpi = s.shape
    allb, alls, alll, rows = [], [], [], []
    for t in range(len(c)):
        off = np.array([o[t, 1], o[t, 0], o[t, 1], o[t, 0]], np.float32)
        allb.append(b[t, :c[t]] + off); alls.append(s[t, :c[t]]); alll.append(l[t, :c[t]])
        rows.append(t * dpi + np.arange(c[t]))
    B, S, L = (torch.from_numpy(np.concatenate(x)) for x in (allb, alls, alll))
    rows = np.concatenate(rows)
    keep = tvb._batched_nms_vanilla(B, S, L, 0.5)
    keep = keep[S[keep] > 0.5].numpy()
    block, kept_rows = single_process(0.5, 0.5)
    assert np.array_equal(np.sort(rows[keep]), kept_rows)
    assert np.array_equal(block[kept_rows, :4], B.numpy()[np.sort(keep)])


def test_rank_bands_cover_own_tiles_only():
    grid = mosaic.tile_grid(16384, 16384, 1024, 128)
    for world in (1, 2, 4, 8):
        for r in range(world):
            y0, y1 = mosaic.rank_band(grid, 1024, 16384, world, r)
            for t in mosaic.rank_tiles(len(grid), world, r):
                assert y0 <= grid[t][0] and grid[t][0] + 1024 <= y1
            assert (y1 - y0) <= 1024 + 896 * (len(mosaic.rank_tiles(len(grid), world, r)) // 19 + 1)
    assert mosaic.rank_band(grid, 1024, 16384, 1, 0) == (0, 16384)
