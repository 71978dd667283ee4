"""GPU parity tests of the training-side siblings against torchvision's own CPU code:
box_iou bit-exact, Matcher indices bit-exact (both RPN and RoI-head settings, ties, gts without any overlap),
BoxCoder.encode within 1e-5 (log), RoIAlign backward against torchvision's CPU autograd (atomic accumulation order:
1e-5 relative to the gradient scale), and the dispatcher route of torchvision's own autograd node."""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
F = np.float32


@pytest.fixture(scope="module")
def ops():
    from miso_b200 import ops as o
    o._lib.load()
    return o


def boxes_case(rng, n, extent=800.0):
    return torch.from_numpy(cases.random_boxes(rng, n, extent=extent, wh=(8.0, 300.0)))


def test_box_iou_bit_exact(ops):
    from torchvision.ops import box_iou
    rng = np.random.default_rng(0)
    a, b = boxes_case(rng, 777), boxes_case(rng, 1234)
    a[5] = b[7]                                   # IoU exactly 1
    b[9] = torch.tensor([10.0, 10.0, 10.0, 10.0])  # zero area
    got = ops.box_iou(a.to(DEV), b.to(DEV)).cpu()
    assert torch.equal(got, box_iou(a, b))
    assert ops.box_iou(a[:0].to(DEV), b.to(DEV)).shape == (0, 1234)


@pytest.mark.parametrize("high,low,allow,weights,n_gt,n_anchor", [
    (0.7, 0.3, True, (1.0, 1.0, 1.0, 1.0), 37, 50000),       # RPN: assign_targets_to_anchors
    (0.5, 0.5, False, (10.0, 10.0, 5.0, 5.0), 12, 2037),      # RoI heads: assign_targets_to_proposals
    (0.7, 0.3, True, (1.0, 1.0, 1.0, 1.0), 1500, 4096),       # more gts than one shared-memory chunk
])
def test_match_and_encode_equals_torchvision(ops, high, low, allow, weights, n_gt, n_anchor):
    from torchvision.models.detection._utils import BoxCoder, Matcher
    from torchvision.ops import box_iou
    rng = np.random.default_rng(n_gt)
    gt, an = boxes_case(rng, n_gt), boxes_case(rng, n_anchor)
    an[:n_gt // 2] = gt[:n_gt // 2]                                # exact hits (IoU 1)
    an[100:110] = an[100]                                          # identical anchors: ties for a gt's best anchor
    gt[-1] = torch.tensor([5000.0, 5000.0, 5100.0, 5100.0])        # a gt nothing overlaps: highest quality 0 -> torchvision restores
    gt[-2] = gt[0]                                                 # duplicate gt: first maximum's index wins
    m = Matcher(high, low, allow_low_quality_matches=allow)
    ref = m(box_iou(gt, an))
    ref_vals = box_iou(gt, an).max(dim=0).values
    ref_t = BoxCoder(weights).encode_single(gt[ref.clamp(min=0)], an)
    got, vals, tg = ops.match_and_encode(gt.to(DEV), an.to(DEV), high, low, allow, weights)
    assert torch.equal(got.cpu(), ref)
    assert torch.equal(vals.cpu(), ref_vals)
    err = (tg.cpu() - ref_t).abs()
    assert float((err / (ref_t.abs() + 1.0)).max()) < 1e-5


@pytest.mark.parametrize("P,sr,aligned,scale", [(7, 2, False, 0.25), (14, 2, False, 0.125), (7, 0, True, 0.25), (5, 3, False, 1.0)])
def test_roi_align_backward_equals_torchvision_cpu_autograd(ops, P, sr, aligned, scale):
    import torchvision
    rng = np.random.default_rng(P + sr)
    n, c, h, w = 2, 40, 50, 38
    x = torch.from_numpy(rng.standard_normal((n, c, h, w)).astype(F)).requires_grad_(True)
    extent = (h / scale, w / scale)
    rois = np.concatenate([rng.integers(0, n, (60, 1)).astype(F), cases.stress_rois(rng, 60, (int(extent[0]), int(extent[1])), side=(4.0, extent[0]))], 1)
    rois = np.concatenate([rois, np.array([[0, -30, -30, 20, 25], [1, 10, 10, 10.2, 10.3], [0, extent[1] - 3, extent[0] - 3, extent[1] + 40, extent[0] + 40]], F)]).astype(F)
    r = torch.from_numpy(rois)
    out = torchvision.ops.roi_align(x, r, P, scale, sr, aligned)
    g = torch.from_numpy(rng.standard_normal(tuple(out.shape)).astype(F))
    out.backward(g)
    got = ops.roi_align_backward(g.to(DEV), r.to(DEV), scale, P, P, n, c, h, w, sr, aligned).cpu()
    tol = 1e-5 * float(x.grad.abs().max()) + 1e-6
    assert float((got - x.grad).abs().max()) <= tol


def test_torchvision_autograd_routes_through_the_override(ops):
    """patch.override_torchvision_ops registers torchvision::roi_align AND torchvision::_roi_align_backward on the CUDA
    key: torchvision's own autograd node then runs forward and backward in libmisob200."""
    import torchvision
    from miso_b200.patch import override_torchvision_ops
    override_torchvision_ops()
    rng = np.random.default_rng(1)
    x_cpu = torch.from_numpy(rng.standard_normal((1, 16, 32, 32)).astype(F)).requires_grad_(True)
    x_gpu = x_cpu.detach().to(DEV).requires_grad_(True)
    rois = torch.tensor([[0, 2.5, 3.5, 60.0, 75.0], [0, 10, 10, 100, 50]], dtype=torch.float32)
    torchvision.ops.roi_align(x_cpu, rois, 7, 0.25, 2).square().sum().backward()
    torchvision.ops.roi_align(x_gpu, rois.to(DEV), 7, 0.25, 2).square().sum().backward()
    assert float((x_gpu.grad.cpu() - x_cpu.grad).abs().max()) <= 1e-5 * float(x_cpu.grad.abs().max()) + 1e-6
