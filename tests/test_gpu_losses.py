"""Fused image-side losses against the reference's torch expressions.

`_ref_*` restate utils/loss_utils.py:17-63, utils/graphics_utils.py:97-126 and train.py:143-163 in eager torch.  They are
evaluated in float64 so the comparison measures OUR fp32 error, not the sum of two fp32 errors; tolerances: losses 2e-6
absolute (values are O(1)), gradients 1e-3 of the gradient's largest magnitude per map (the fp32 eager reference itself
sits at ~1e-4 of that scale).
"""
import math
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _window(channel, dtype, device):
    g = torch.tensor([math.exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, 11, 11).contiguous().to(device=device, dtype=dtype)


def _ref_ssim(img1, img2):
    if img1.dim() == 3:
        img1 = img1.unsqueeze(0)
    img2 = img2.reshape(img1.shape)
    ch = img1.size(-3)
    w = _window(ch, img1.dtype, img1.device)
    mu1 = F.conv2d(img1, w, padding=5, groups=ch)
    mu2 = F.conv2d(img2, w, padding=5, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=5, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=5, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=5, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def _ref_l1(a, b):
    return torch.abs(a - b.reshape(a.shape)).mean()


def _ref_points_from_depth(view, H, W, d1, d2):
    fx = W / (2 * math.tan(view.FoVx / 2.))
    fy = H / (2 * math.tan(view.FoVy / 2.))
    intrins_inv = torch.tensor([[1 / fx, 0., -W / (2 * fx)], [0., 1 / fy, -H / (2 * fy)], [0., 0., 1.0]]).float().cuda().to(d1.dtype)
    gx, gy = torch.meshgrid(torch.arange(W) + 0.5, torch.arange(H) + 0.5, indexing='xy')
    pts = torch.stack([gx, gy, torch.ones_like(gx)], dim=0).reshape(3, -1).float().cuda().to(d1.dtype)
    rays = intrins_inv @ pts
    return (d1.reshape(1, -1) * rays).reshape(3, H, W), (d2.reshape(1, -1) * rays).reshape(3, H, W)


def _ref_normal_loss(normal, p1, p2, ratio=0.6):
    points = torch.stack([p1, p2], dim=0)
    out = torch.zeros_like(points)
    dx = points[..., 2:, 1:-1] - points[..., :-2, 1:-1]
    dy = points[..., 1:-1, 2:] - points[..., 1:-1, :-2]
    out[..., 1:-1, 1:-1] = F.normalize(torch.cross(dx, dy, dim=1), dim=1)
    err = 1 - (normal.unsqueeze(0) * out).sum(dim=1)
    return (1 - ratio) * err[0].mean() + ratio * err[1].mean()


def _images(C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(C, H, W, generator=g)
    img = (gt + 0.15 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    img[:, : H // 4] = gt[:, : H // 4]  # a region with img == gt exactly: sign(0) = 0 in the L1 gradient
    return img.cuda(), gt.cuda()


def _grad_close(ours, ref, rel=1e-3):
    scale = ref.abs().max().item()
    err = (ours.double() - ref).abs().max().item()
    assert err <= rel * scale, (err, scale)


@pytest.mark.parametrize("C,H,W", [(3, 67, 101), (3, 16, 32), (1, 5, 7), (3, 200, 333)])
def test_ssim_and_l1_match_reference(C, H, W):
    from rade_gs_b200 import losses

    img, gt = _images(C, H, W, 3 * H + W)
    for name, fn, ref in (("ssim", losses.ssim, _ref_ssim), ("l1", losses.l1_loss, _ref_l1)):
        a = img.clone().requires_grad_(True)
        v = fn(a, gt.unsqueeze(0))  # the reference passes gt_image.unsqueeze(0) (train.py:163)
        v.backward()
        b = img.double().requires_grad_(True)
        rv = ref(b, gt.double())
        rv.backward()
        assert v.dtype == torch.float32 and v.dim() == 0
        assert abs(v.item() - rv.item()) < 2e-6, (name, v.item(), rv.item())
        _grad_close(a.grad, b.grad)
    # evaluation call: no graph, no derivative maps
    with torch.no_grad():
        assert abs(losses.ssim(img, gt).item() - _ref_ssim(img.double(), gt.double()).item()) < 2e-6


def test_l1_ssim_loss_is_train_py_line_163():
    from rade_gs_b200 import losses

    lam = 0.2
    img, gt = _images(3, 120, 160, 5)
    a = img.clone().requires_grad_(True)
    (losses.l1_ssim_loss(a, gt, lam) * 1.7).backward()  # upstream gradient other than 1
    b = img.double().requires_grad_(True)
    ref = (1.0 - lam) * _ref_l1(b, gt.double()) + lam * (1.0 - _ref_ssim(b, gt.double()))
    (ref * 1.7).backward()
    assert abs(losses.l1_ssim_loss(img, gt, lam).item() - ref.item()) < 2e-6
    _grad_close(a.grad, b.grad)
    # against the eager fp32 expression too (what train.py runs): same loss to fp32 accuracy
    ref32 = (1.0 - lam) * _ref_l1(img, gt) + lam * (1.0 - _ref_ssim(img, gt))
    assert abs(losses.l1_ssim_loss(img, gt, lam).item() - ref32.item()) < 1e-5


@pytest.mark.parametrize("H,W", [(37, 53), (16, 16), (3, 3), (2, 9), (130, 70)])
def test_depth_normal_consistency_matches_reference(H, W):
    from rade_gs_b200 import losses

    g = torch.Generator().manual_seed(H * 7 + W)
    view = SimpleNamespace(FoVx=0.9, FoVy=0.7, image_width=W, image_height=H)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    base = 3.0 + 0.02 * xx + 0.03 * yy + 0.3 * torch.sin(xx * 0.4) * torch.cos(yy * 0.3)
    d1 = (base + 0.01 * torch.randn(H, W, generator=g))[None].cuda()
    d2 = (base + 0.05 * torch.randn(H, W, generator=g))[None].cuda()
    nrm = F.normalize(torch.randn(3, H, W, generator=g), dim=0).cuda() * 0.9

    leaves = [t.clone().requires_grad_(True) for t in (nrm, d1, d2)]
    v = losses.depth_normal_consistency_loss(view, *leaves)
    (v * 0.05).backward()  # lambda_depth_normal
    ref_leaves = [t.double().requires_grad_(True) for t in (nrm, d1, d2)]
    p1, p2 = _ref_points_from_depth(view, H, W, ref_leaves[1], ref_leaves[2])
    rv = _ref_normal_loss(ref_leaves[0], p1, p2)
    (rv * 0.05).backward()
    assert abs(v.item() - rv.item()) < 2e-6, (v.item(), rv.item())
    for a, b in zip(leaves, ref_leaves):
        assert a.grad.shape == a.shape
        if b.grad.abs().max() > 0:
            _grad_close(a.grad, b.grad, rel=2e-3)
        else:
            assert a.grad.abs().max().item() == 0.0


def test_point_normal_consistency_matches_reference():
    from rade_gs_b200 import losses

    H, W = 45, 61
    g = torch.Generator().manual_seed(8)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    z = 4.0 + 0.2 * torch.sin(xx * 0.3) + 0.1 * torch.cos(yy * 0.5)
    pts = torch.stack([(xx - W / 2) * z / 50, (yy - H / 2) * z / 50, z])
    c1 = (pts + 0.01 * torch.randn(3, H, W, generator=g)).cuda()
    c2 = (pts + 0.03 * torch.randn(3, H, W, generator=g)).cuda()
    nrm = F.normalize(torch.randn(3, H, W, generator=g), dim=0).cuda()
    leaves = [t.clone().requires_grad_(True) for t in (nrm, c1, c2)]
    v = losses.point_normal_consistency_loss(*leaves)
    v.backward()
    ref_leaves = [t.double().requires_grad_(True) for t in (nrm, c1, c2)]
    rv = _ref_normal_loss(*ref_leaves)
    rv.backward()
    assert abs(v.item() - rv.item()) < 2e-6
    for a, b in zip(leaves, ref_leaves):
        _grad_close(a.grad, b.grad, rel=2e-3)


def test_losses_refuse_cpu_and_gt_gradients():
    from rade_gs_b200 import losses

    a, b = torch.rand(3, 8, 8), torch.rand(3, 8, 8)
    with pytest.raises(RuntimeError):
        losses.ssim(a, b)
    with pytest.raises(ValueError):
        losses.l1_ssim_loss(a.cuda(), b.cuda().requires_grad_(True), 0.2)


def test_training_step_slice_with_fused_losses():
    """render -> fused photometric + normal-consistency loss -> backward: gradients arrive at the Gaussians."""
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import losses, scenes
    from test_gpu_api import _settings

    sc, _, _ = scenes.make_config("C1")
    sc = sc.to("cuda")
    view = SimpleNamespace(FoVx=2 * math.atan(sc.tanfovx), FoVy=2 * math.atan(sc.tanfovy))
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    means2D = torch.zeros_like(sc.means3D, requires_grad=True)
    color, radii, _, _, depth, mdepth, alpha, normal = dgr.GaussianRasterizer(_settings(dgr, sc, False, True, ks=0.1))(
        means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"], shs=leaves["shs"], scales=leaves["scales"],
        rotations=leaves["rotations"])
    gt = torch.rand_like(color)
    loss = losses.l1_ssim_loss(color, gt, 0.2) + 0.05 * losses.depth_normal_consistency_loss(view, normal, depth, mdepth)
    ref = (0.8 * _ref_l1(color, gt) + 0.2 * (1 - _ref_ssim(color, gt))) + 0.05 * _ref_normal_loss(
        normal, *_ref_points_from_depth(view, sc.height, sc.width, depth, mdepth))
    assert abs(loss.item() - ref.item()) < 1e-5
    loss.backward()
    for k, t in leaves.items():
        assert t.grad is not None and torch.isfinite(t.grad).all(), k
    assert leaves["means3D"].grad.abs().max() > 0


def test_row_range_ssim_kernels_add_up_to_the_whole_image():
    """The row-sharded form of the L1 + SSIM kernels (multi-GPU slab-local loss): sums over disjoint row ranges add up to the
    whole-image sums, and the per-range gradients (which reach 5 rows into the neighbouring ranges) add up to the whole gradient."""
    import diff_gaussian_rasterization as dgr
    C = dgr._C
    DEV = "cuda:0"
    g = torch.Generator().manual_seed(17)
    H, W = 16 * 7 + 9, 150
    gt = torch.rand(3, H, W, generator=g).to(DEV)
    img = (gt + 0.2 * torch.randn(3, H, W, generator=g).to(DEV)).clamp(0, 1)
    sums, dmaps = C.ssim_l1_forward(img, gt, True)
    d_whole = C.ssim_l1_backward(img, gt, dmaps, -0.2, 0.8, torch.empty(0))
    for cuts in ([0, 48, 80, H], [0, 16, 32, 64, 96, H]):
        tot = torch.zeros(2, dtype=torch.float64, device=DEV)
        d_sum = torch.zeros_like(img)
        for r0, r1 in zip(cuts[:-1], cuts[1:]):
            s, dm = C.ssim_l1_forward(img, gt, True, r0, r1)
            tot += s
            d = C.ssim_l1_backward(img, gt, dm, -0.2, 0.8, torch.empty(0), r0, r1)
            far = d.clone()
            far[:, max(r0 - 5, 0):min(r1 + 5, H)] = 0
            assert not far.any()                              # nothing further than the 5 halo rows
            d_sum += d
        assert torch.allclose(tot, sums, rtol=1e-9, atol=1e-9)
        assert torch.allclose(d_sum, d_whole, rtol=1e-5, atol=1e-7)
