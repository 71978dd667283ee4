"""GPU parity tests of the TMA-staged RoIAlign kernel (k_roi_align_tma, miso_b200/csrc/roi_align_tma.cu):
bit-exact against the CPU oracle (tv-csrc:ops/cpu/roi_align_kernel.cpp:393 restated in oracle/c) and
bit-identical to the register-gather kernel it replaces, over the cases that exercise its moving parts —
ring wrap-around, more than 64 staged rows in flight, rows wider than the ring allows (direct route),
samples outside the map (skipped, not multiplied by zero), dead rows, channel counts that do not fill a
128-channel pass, many RoIs per CTA."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import detection as D
from oracle import native
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F = np.float32


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def ops():
    from miso_b200 import ops as o
    o._lib.load()
    return o


def mixed_rois(rng, n_img, k, hw):
    """[K,5] RoIs: stress distribution + tiny + huge + far outside + extreme aspect + on the borders."""
    h, w = hw
    parts = [cases.stress_rois(rng, k, hw, side=(8.0, 1000.0))]
    c = rng.uniform(0, [w, h], (k // 8, 2))
    s = rng.uniform(0.0, 5.0, (k // 8, 2))
    parts.append(np.concatenate([c, c + s], 1))                                      # sub-bin-sized
    parts.append(np.array([[0, 0, w, h], [-40, -40, w + 40, h + 40], [-300, -300, -200, -250], [w - 1, h - 1, w, h],
                           [w + 10, 5, w + 90, 60], [3, h + 30, 80, h + 90], [0, 10, w, 14], [10, 0, 13, h]], np.float64))
    c = rng.uniform(-30, [w + 30, h + 30], (k // 4, 2))
    s = np.exp(rng.uniform(np.log(4), np.log(600), (k // 4, 2)))
    parts.append(np.concatenate([c - s / 2, c + s / 2], 1))                          # partly outside, any aspect
    b = np.concatenate(parts, 0).astype(F)
    idx = rng.integers(0, n_img, (b.shape[0], 1)).astype(F)
    r = np.concatenate([idx, b], 1).astype(F)
    return r[rng.permutation(r.shape[0])]


@pytest.mark.parametrize("channels", [256, 136, 32])
def test_tma_equals_gather_and_oracle_single_level(ops, channels):
    rng = np.random.default_rng(channels)
    n, h, w = 2, 60, 76
    x = rng.standard_normal((n, channels, h, w)).astype(F)
    rois = mixed_rois(rng, n, 600, (h * 4, w * 4))
    xc = cu(x).contiguous(memory_format=torch.channels_last)
    lib = ops._lib.load()
    for exact in (True, False):
        n0 = lib.mb_roi_align_tma_launches()
        a = ops.roi_align(xc, cu(rois), 7, 0.25, 2, False, exact=exact, force_gather="tma")
        assert lib.mb_roi_align_tma_launches() == n0 + 1          # the TMA-staged kernel produced `a`
        b = ops.roi_align(xc, cu(rois), 7, 0.25, 2, False, exact=exact, force_gather="gather")
        assert lib.mb_roi_align_tma_launches() == n0 + 1          # ... and the gather kernel `b`
        if exact:
            assert torch.equal(a, b)
        else:
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(np.abs(x).max()))
    pick = np.sort(rng.choice(rois.shape[0], 96, replace=False))
    ref = native.roi_align(x, rois[pick], 0.25, 7, 7, 2, False)
    for route in ("tma", "gather"):
        got = ops.roi_align(xc, cu(rois), 7, 0.25, 2, False, force_gather=route)[torch.from_numpy(pick).to(DEV)].cpu().numpy()
        assert np.array_equal(got, ref), route


def test_tma_multiscale_many_rois_per_cta(ops):
    """4000 RoIs over a 4-level 256-channel pyramid (27 RoIs per CTA: geometry double buffer, >64 staged rows,
    ring wrap-around, output double buffer all cycle many times)."""
    rng = np.random.default_rng(11)
    n, c = 2, 256
    hw = (416, 544)
    feats = [rng.standard_normal((n, c, hw[0] // s, hw[1] // s)).astype(F) for s in (4, 8, 16, 32)]
    boxes = [cases.stress_rois(rng, 2000, hw, side=(8.0, 700.0)) for _ in range(n)]
    x = {str(i): cu(f).contiguous(memory_format=torch.channels_last) for i, f in enumerate(feats)}
    tb = [cu(b) for b in boxes]
    from miso_b200 import _lib
    n0 = _lib.load().mb_roi_align_tma_launches()
    a, la = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, force_gather="tma")(x, tb, [hw] * n, return_levels=True)
    assert _lib.load().mb_roi_align_tma_launches() == n0 + 1                      # the module passes the route through
    b, lb = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, force_gather="gather")(x, tb, [hw] * n, return_levels=True)
    assert _lib.load().mb_roi_align_tma_launches() == n0 + 1
    assert torch.equal(la, lb) and set(la.cpu().tolist()) == {0, 1, 2, 3}
    assert torch.equal(a, b)
    pick = np.sort(rng.choice(4000, 80, replace=False))
    sub = [boxes[i][pick[pick // 2000 == i] % 2000] for i in range(n)]
    ref = D.multiscale_roi_align(feats, sub, [hw] * n, 7, 2)
    assert np.array_equal(a[torch.from_numpy(pick).to(DEV)].cpu().numpy(), ref)
    fast = ops.MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, exact=False, force_gather="tma")(x, tb, [hw] * n)
    assert torch.allclose(fast, a, rtol=1e-5, atol=5e-5)


def test_tma_skips_invalid_samples_like_the_reference(ops):
    """A sample outside [-1, size] contributes nothing in the reference — it is skipped, not multiplied by a zero
    weight — so an Inf/NaN feature elsewhere in the map must not leak into the bin."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 32, 40, 40)).astype(F)
    x[0, :, 0, 0] = np.inf
    x[0, 5, 0, 1] = np.nan
    rois = np.array([[0, 100, 100, 260, 260], [0, -60, 20, 30, 90], [0, 20, -60, 90, 30], [0, 120, 120, 150, 150]], F)
    xc = cu(x).contiguous(memory_format=torch.channels_last)
    ref = native.roi_align(x, rois, 0.25, 7, 7, 2, False)
    for route in ("tma", "gather"):
        got = ops.roi_align(xc, cu(rois), 7, 0.25, 2, False, force_gather=route).cpu().numpy()
        assert np.array_equal(got, ref, equal_nan=True), route
        assert np.isfinite(got[0]).all() and np.isfinite(got[3]).all()


def test_tma_per_image_layout_with_dead_rows(ops):
    """The fused path's RoI layout: [N, R, 4] + live counts; rows beyond the count produce zeros."""
    from miso_b200 import _lib
    from miso_b200._lib import RoiAlignParams
    from miso_b200.ops import _ptr, _stream, level_thresholds
    rng = np.random.default_rng(5)
    n, c, R = 3, 64, 300
    hw = (320, 384)
    feats = [rng.standard_normal((n, c, hw[0] // s, hw[1] // s)).astype(F) for s in (4, 8, 16, 32)]
    boxes = np.stack([cases.stress_rois(rng, R, hw, side=(8.0, 500.0)) for _ in range(n)])
    counts = np.array([R, 117, 0], np.int32)
    tf = [cu(f).contiguous(memory_format=torch.channels_last) for f in feats]
    tb, tc = cu(boxes), cu(counts)
    outs = []
    for force in (2, 1):
        p = RoiAlignParams()
        p.num_levels, p.num_images, p.channels, p.pooled_h, p.pooled_w = 4, n, c, 7, 7
        p.sampling_ratio, p.aligned, p.exact, p.channels_last, p.force_gather = 2, 0, 1, 1, force
        for l, f in enumerate(tf):
            p.height[l], p.width[l], p.spatial_scale[l], p.features[l] = f.shape[2], f.shape[3], 1.0 / (4 << l), f.data_ptr()
        for i, t in enumerate(level_thresholds(2, 5)):
            p.level_thresholds[i] = t
        p.boxes_per_image, p.box_counts = R, tc.data_ptr()
        out = torch.full((n * R, c, 7, 7), 7.0, device=DEV)
        nb = _lib.load().mb_roi_align_workspace_bytes(C.byref(p), n * R)
        assert nb > 0
        ws = torch.empty((nb,), dtype=torch.uint8, device=DEV)
        _lib.check(_lib.load().mb_multiscale_roi_align(C.byref(p), _ptr(tb), n * R, _ptr(out), None, _ptr(ws), nb, _stream(tb)), "roi")
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    o = outs[0].cpu().numpy().reshape(n, R, c, 7, 7)
    assert not o[1, 117:].any() and not o[2].any()
    ref = D.multiscale_roi_align(feats, [boxes[0][:40], boxes[1][:40], boxes[2][:0]], [hw] * n, 7, 2)
    assert np.array_equal(np.concatenate([o[0, :40], o[1, :40]]), ref)
