"""GPU parity tests of the sparse seam NMS (mb_seam_nms): the kept set must equal torchvision's per-label
strategy over all live rows (tests/mosaic_ref.seam_keep_rows -> oracle C NMS) and the dense mb_nms mode 1,
bit for bit, for any input — tile-shaped or not."""
import numpy as np
import pytest
import torch

from tests import mosaic_ref as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F = np.float32


@pytest.fixture(scope="module")
def mosaic():
    from miso_b200 import mosaic as m
    m._lib.load()
    return m


def seam_block(rng, nty=5, ntx=5, dpi=120, objects=1400, thr=0.5, classes=2):
    """Detections of a tile grid whose objects repeat (jittered) in every tile they fall into."""
    tile, stride = 1024, 896
    H, W = stride * (nty - 1) + tile, stride * (ntx - 1) + tile
    c = rng.uniform(0, [W, H], (objects, 2))
    s = np.exp(rng.uniform(np.log(24), np.log(220), (objects, 2)))
    obj = np.concatenate([c - s / 2, c + s / 2], 1)
    lab = rng.integers(1, classes + 1, objects)
    T = nty * ntx
    boxes = np.zeros((T, dpi, 4), F); scores = np.zeros((T, dpi), F); labels = np.zeros((T, dpi), np.int64)
    counts = np.zeros(T, np.int32); origins = np.zeros((T, 2), F)
    for ty in range(nty):
        for tx in range(ntx):
            t = ty * ntx + tx
            oy, ox = ty * stride, tx * stride
            origins[t] = (oy, ox)
            inside = (obj[:, 2] > ox + 4) & (obj[:, 0] < ox + tile - 4) & (obj[:, 3] > oy + 4) & (obj[:, 1] < oy + tile - 4)
            idx = np.nonzero(inside)[0][:dpi]
            b = obj[idx] + rng.normal(0, 1.5, (len(idx), 4)) - np.array([ox, oy, ox, oy])
            b = np.clip(b, 0, tile)
            sc = rng.uniform(0.3, 1.0, len(idx))
            order = np.argsort(-sc, kind="stable")
            k = len(idx)
            boxes[t, :k], scores[t, :k], labels[t, :k], counts[t] = b[order], sc[order], lab[idx][order], k
    return R.pack_block(boxes, scores, labels, counts, origins, thr, T * dpi + 37), dpi


def run_sparse(mosaic, block, dpi, thr, **kw):
    g = torch.from_numpy(block).to(DEV)
    s = mosaic.SparseSeamNms(block.shape[0], dpi, DEV, want_keep=True, **kw)
    s.launch(g, thr)
    n, edges = s.check()
    state = s.state.cpu().numpy()
    keep = s.keep[:n].cpu().numpy()
    assert np.array_equal(keep, np.nonzero(state == 1)[0])
    assert np.array_equal(state == 3, block[:, 5] < 0)
    return keep, edges


@pytest.mark.parametrize("thr", [0.5, 0.3, 0.0])
def test_sparse_seam_equals_dense_and_oracle(mosaic, thr):
    rng = np.random.default_rng(17)
    block, dpi = seam_block(rng)
    keep, edges = run_sparse(mosaic, block, dpi, thr)
    ref = R.seam_keep_rows(block, thr)
    assert 0 < len(ref) < int((block[:, 5] >= 0).sum())          # the seam duplicates were suppressed
    assert np.array_equal(keep, ref)
    assert edges > 0
    g = torch.from_numpy(block).to(DEV)
    dense = mosaic.SeamNms(block.shape[0], 3, DEV)
    dense.launch(g, thr)
    _, _, _, rows = dense.finish()
    assert np.array_equal(np.sort(rows.cpu().numpy()), keep)
    assert np.array_equal(mosaic.by_score(g, torch.from_numpy((np.isin(np.arange(len(block)), keep)).astype(np.int32)).to(DEV)).cpu().numpy(),
                          rows.cpu().numpy())


def test_sparse_seam_arbitrary_rows_ties_and_degenerate_boxes(mosaic):
    """No geometric assumption: boxes anywhere, groups of 64 rows, equal scores (lower row wins), zero-area
    duplicates (NaN IoU -> kept), ignored rows sprinkled in."""
    rng = np.random.default_rng(23)
    n = 64 * 40 + 11
    c = rng.uniform(0, 900, (n, 2)); s = np.exp(rng.uniform(np.log(8), np.log(200), (n, 2)))
    block = np.zeros((n, 6), F)
    block[:, :4] = np.concatenate([c - s / 2, c + s / 2], 1)
    block[:, 4] = np.floor(rng.uniform(0, 1, n) * 16) / 16                      # heavy ties
    block[:, 5] = rng.integers(0, 3, n)
    block[rng.choice(n, 300, replace=False), 5] = -1
    block[100:104, :4] = np.array([50, 50, 50, 50], F)
    block[100:104, 5] = 1
    for thr in (0.5, 0.7):
        keep, _ = run_sparse(mosaic, block, 64, thr)
        assert np.array_equal(keep, R.seam_keep_rows(block, thr))


def test_sparse_seam_long_dependency_chain(mosaic):
    """A chain of 600 boxes, each suppressing the next when kept: the fixed rounds cannot settle it, the finishing
    kernel does (kept = every other box)."""
    n = 600
    block = np.zeros((n, 6), F)
    x = np.arange(n) * 10.0
    block[:, 0], block[:, 1], block[:, 2], block[:, 3] = x, 0, x + 40, 40              # IoU(i, i+1) = 30/50 = 0.6
    block[:, 4] = np.linspace(0.99, 0.51, n)
    block[:, 5] = 1
    keep, _ = run_sparse(mosaic, block, 50, 0.5)
    ref = R.seam_keep_rows(block, 0.5)
    assert np.array_equal(keep, ref) and len(ref) < n


def test_sparse_seam_reports_edge_overflow_and_rejects_negative_threshold(mosaic):
    from miso_b200 import MisoB200Error
    rng = np.random.default_rng(5)
    block = np.zeros((4096, 6), F)
    block[:, :4] = np.array([10, 10, 60, 60], F) + rng.normal(0, 0.5, (4096, 4))      # everything overlaps everything
    block[:, 4] = rng.permutation(4096) / 4096.0
    block[:, 5] = 1
    g = torch.from_numpy(block).to(DEV)
    s = mosaic.SparseSeamNms(4096, 256, DEV, edges_per_row=1)
    s.launch(g, 0.5)
    with pytest.raises(MisoB200Error):
        s.check()
    big = mosaic.SparseSeamNms(4096, 256, DEV, edges_per_row=2100, want_keep=True)
    big.launch(g, 0.5)
    n, edges = big.check()
    assert n == 1 and edges == 4096 * 4095 // 2
    with pytest.raises(MisoB200Error):
        mosaic.SparseSeamNms(4096, 256, DEV).launch(g, -0.1)


def test_sparse_seam_more_neighbour_tiles_than_the_list_holds(mosaic):
    """300 groups of 16 rows whose bounding boxes all touch: the per-tile neighbour list (96 entries) overflows and
    the pairs kernel walks the tiles in order instead; rows with more than four suppressors take the second pass."""
    rng = np.random.default_rng(41)
    n = 16 * 300 + 5
    c = rng.uniform(0, 700, (n, 2)); s = np.exp(rng.uniform(np.log(30), np.log(260), (n, 2)))
    block = np.zeros((n, 6), F)
    block[:, :4] = np.concatenate([c - s / 2, c + s / 2], 1)
    block[:, 4] = rng.uniform(0, 1, n)
    block[:, 5] = rng.integers(0, 2, n)
    block[rng.choice(n, 200, replace=False), 5] = -1
    keep, edges = run_sparse(mosaic, block, 16, 0.3, edges_per_row=64)
    assert np.array_equal(keep, R.seam_keep_rows(block, 0.3))
    assert edges > 4 * n
