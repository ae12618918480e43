"""Fused activations / densification statistics against the reference's torch expressions.

The torch functions below restate reference scene/gaussian_model.py:156-166, :125-126, :743-747 and train.py:187-188 in
plain fp32 eager torch (this is a floating-point kernel, so the checker is the torch fp32 reference of the same op).
Tolerances: forward 2 ulp-ish (rel 1e-6), backward rel 1e-4 of the row-wise gradient scale (autograd sums the same
terms in another order); the statistics are exact.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_activate(raw_scaling, raw_opacity, raw_rotation, filter_3D):
    opacity = torch.sigmoid(raw_opacity)
    scales = torch.exp(raw_scaling)
    scales_square = torch.square(scales)
    det1 = scales_square.prod(dim=1)
    scales_after_square = scales_square + torch.square(filter_3D)
    det2 = scales_after_square.prod(dim=1)
    coef = torch.sqrt(det1 / det2)
    return torch.sqrt(scales_after_square), opacity * coef[..., None], torch.nn.functional.normalize(raw_rotation)


def _raw(P, seed, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    raw_scaling = (torch.randn(P, 3, generator=g) * 1.5 - 3.0).to(device)
    raw_opacity = (torch.randn(P, 1, generator=g) * 2.0).to(device)
    raw_rotation = torch.randn(P, 4, generator=g).to(device)
    filter_3D = (torch.rand(P, 1, generator=g) * 0.05 + 1e-4).to(device)
    return raw_scaling, raw_opacity, raw_rotation, filter_3D


@pytest.mark.parametrize("P", [1, 257, 100_003])
def test_activate_forward_matches_torch(P):
    from rade_gs_b200.fused import activate_gaussians

    raw = _raw(P, 11 + P)
    s, o, r = activate_gaussians(*raw)
    rs, ro, rr = _ref_activate(*raw)
    assert s.shape == rs.shape and o.shape == ro.shape and r.shape == rr.shape
    torch.testing.assert_close(s, rs, rtol=1e-6, atol=0)
    torch.testing.assert_close(o, ro, rtol=2e-6, atol=1e-12)
    torch.testing.assert_close(r, rr, rtol=1e-6, atol=1e-7)


def test_activate_backward_matches_autograd():
    from rade_gs_b200.fused import activate_gaussians

    P = 50_021
    raw = _raw(P, 5)
    g = torch.Generator(device="cpu").manual_seed(99)
    gs, go, gr = (torch.randn(P, k, generator=g).cuda() for k in (3, 1, 4))

    def run(fn):
        leaves = [t.clone().requires_grad_(True) for t in raw[:3]]
        s, o, r = fn(*leaves, raw[3])
        torch.autograd.backward([s, o, r], [gs, go, gr])
        return [t.grad for t in leaves]

    ours, ref = run(activate_gaussians), run(_ref_activate)
    for a, b, name in zip(ours, ref, ("scaling", "opacity", "rotation")):
        scale = b.abs().amax(dim=1, keepdim=True).clamp_min(1e-20)
        err = ((a - b).abs() / scale).max().item()
        assert err < 1e-4, f"d_raw_{name}: {err}"


def test_activate_backward_partial_outputs():
    """Only the opacity output is used downstream: autograd passes zeros for the rest."""
    from rade_gs_b200.fused import activate_gaussians

    raw = _raw(1000, 3)
    leaves = [t.clone().requires_grad_(True) for t in raw[:3]]
    _, o, _ = activate_gaussians(*leaves, raw[3])
    o.sum().backward()
    ref_leaves = [t.clone().requires_grad_(True) for t in raw[:3]]
    _, ro, _ = _ref_activate(*ref_leaves, raw[3])
    ro.sum().backward()
    torch.testing.assert_close(leaves[1].grad, ref_leaves[1].grad, rtol=1e-4, atol=1e-9)
    # d coef / d raw_scaling is the difference of two nearly equal chain-rule terms (through det1 and det2) of size
    # ~|g o coef| <= 1: both sides carry ~1e-7 of cancellation noise, so the tolerance is absolute
    torch.testing.assert_close(leaves[0].grad, ref_leaves[0].grad, rtol=1e-3, atol=2e-6)
    assert leaves[2].grad.abs().max().item() == 0.0


def test_activate_rejects_cpu_tensors():
    from rade_gs_b200.fused import activate_gaussians

    raw = _raw(8, 1, device="cpu")
    with pytest.raises(RuntimeError):
        activate_gaussians(*raw)


def test_densification_stats_match_reference_expressions():
    from rade_gs_b200.fused import add_densification_stats_

    P = 70_001
    g = torch.Generator(device="cpu").manual_seed(2)
    grad = torch.randn(P, 3, generator=g).cuda() * 1e-3
    radii = torch.randint(-1, 40, (P,), generator=g, dtype=torch.int32).cuda()
    radii[radii < 0] = 0
    state = [torch.rand(P, 1, generator=g).cuda() for _ in range(4)]
    max_radii = (torch.rand(P, generator=g) * 30).cuda()

    ref = [t.clone() for t in state]
    ref_max = max_radii.clone()
    vis = radii > 0
    ref_max[vis] = torch.max(ref_max[vis], radii[vis])
    ref[0][vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
    ref[1][vis] += torch.norm(grad[vis, 2:], dim=-1, keepdim=True)
    ref[2][vis] = torch.max(ref[2][vis], torch.norm(grad[vis, 2:], dim=-1, keepdim=True))
    ref[3][vis] += 1

    add_densification_stats_(grad, radii, state[0], state[1], state[2], state[3], max_radii)
    torch.testing.assert_close(state[0], ref[0], rtol=2e-7, atol=0)
    for a, b in zip(state[1:], ref[1:]):
        assert torch.equal(a, b)
    assert torch.equal(max_radii, ref_max)

    # without the radius maximum
    before = max_radii.clone()
    add_densification_stats_(grad, radii, state[0], state[1], state[2], state[3])
    assert torch.equal(max_radii, before)
    assert torch.equal(state[3], ref[3] + vis[:, None].float())


def test_fused_step_through_rasterizer():
    """activate -> rasterize -> backward -> statistics as one training-step slice; gradients reach the raw parameters."""
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import scenes
    from rade_gs_b200.fused import activate_gaussians, add_densification_stats_
    from test_gpu_api import _settings

    sc, coord, depth = scenes.make_config("C1")
    sc = sc.to("cuda")
    P = sc.means3D.shape[0]
    filter_3D = _raw(P, 21)[3] * 0.1
    # raw parameters whose activations reproduce the synthetic scene (up to the 3D filter)
    raw_scaling = torch.log(sc.scales.clamp_min(1e-6)).requires_grad_(True)
    raw_opacity = torch.logit(sc.opacities.clamp(1e-4, 1 - 1e-4)).requires_grad_(True)
    raw_rotation = (sc.rotations * 1.7).requires_grad_(True)
    s, o, r = activate_gaussians(raw_scaling, raw_opacity, raw_rotation, filter_3D)
    means2D = torch.zeros(P, 3, device="cuda", requires_grad=True)
    out = dgr.GaussianRasterizer(_settings(dgr, sc, coord, depth, ks=0.1))(
        means3D=sc.means3D, means2D=means2D, opacities=o, shs=sc.shs, scales=s, rotations=r)
    color, radii = out[0], out[1]
    color.mean().backward()
    for t in (raw_scaling, raw_opacity, raw_rotation):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().max() > 0
    stats = [torch.zeros(P, 1, device="cuda") for _ in range(4)]
    max_radii = torch.zeros(P, device="cuda")
    add_densification_stats_(means2D.grad, radii, *stats, max_radii)
    vis = radii > 0
    assert vis.any()
    assert torch.equal(stats[3][:, 0] > 0, vis)
    assert torch.equal(max_radii, radii.float())


@pytest.mark.parametrize("deg,M", [(3, 16), (1, 4), (2, 16)])
def test_split_sh_layout_is_bit_identical_to_concatenated(deg, M):
    """shs=(features_dc, features_rest) must give the same bits as shs=torch.cat(...) (gaussian_model.py:133-136)."""
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import scenes
    from test_gpu_api import _settings

    sc, coord, depth = scenes.make_config("C1")
    sc = sc.to("cuda")
    P = sc.means3D.shape[0] - 13  # ragged last warp in the SH kernel
    dc = sc.shs[:P, :1].contiguous()
    rest = sc.shs[:P, 1:M].contiguous()
    up = scenes.make_upstream_grads(sc.height, sc.width, device="cuda")

    def run(split):
        leaves = {k: getattr(sc, k)[:P].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities")}
        f_dc, f_rest = dc.clone().requires_grad_(True), rest.clone().requires_grad_(True)
        shs = (f_dc, f_rest) if split else torch.cat((f_dc, f_rest), dim=1)
        means2D = torch.zeros(P, 3, device="cuda", requires_grad=True)
        out = dgr.GaussianRasterizer(_settings(dgr, sc, coord, depth, ks=0.1, deg=deg))(
            means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"], shs=shs, scales=leaves["scales"],
            rotations=leaves["rotations"])
        color, radii, _, _, d, md, alpha, normal = out
        loss = (color * up["color"]).sum() + (d * up["depth"]).sum() + (alpha * up["alpha"]).sum() + (normal * up["normal"]).sum()
        loss.backward()
        return [o.detach() for o in out], f_dc.grad, f_rest.grad, leaves["means3D"].grad

    o_cat, dc_cat, rest_cat, m_cat = run(False)
    o_spl, dc_spl, rest_spl, m_spl = run(True)
    for a, b in zip(o_cat, o_spl):
        assert torch.equal(a, b)
    assert dc_spl.shape == dc.shape and rest_spl.shape == rest.shape
    # per-Gaussian SH arithmetic is identical; only the screen-space accumulation order (atomics) differs run to run, which
    # single elements with cancellation amplify: compare like every other gradient check (relative L2 + max-norm)
    from tolerances import grad_close_gpu
    for a, b, name in ((dc_spl, dc_cat, "dc"), (rest_spl, rest_cat, "rest"), (m_spl, m_cat, "means3D")):
        grad_close_gpu(a.cpu().numpy(), b.cpu().numpy(), name)
    assert rest_spl.abs().max() > 0 and dc_spl.abs().max() > 0


def _ref_compute_3D_filter(xyz, cameras):
    """scene/gaussian_model.py:179-232 restated (same torch expressions)."""
    import math

    distance = torch.ones((xyz.shape[0]), device=xyz.device) * 100000.0
    valid_points = torch.zeros((xyz.shape[0]), device=xyz.device, dtype=torch.bool)
    focal_length = 0.
    for camera in cameras:
        W, H = camera.image_width, camera.image_height
        focal_x = W / (2 * math.tan(camera.FoVx / 2.))
        focal_y = H / (2 * math.tan(camera.FoVy / 2.))
        R = torch.tensor(camera.R, device=xyz.device, dtype=torch.float32)
        T = torch.tensor(camera.T, device=xyz.device, dtype=torch.float32)
        xyz_cam = xyz @ R + T[None, :]
        valid_depth = xyz_cam[:, 2] > 0.2
        x, y, z = xyz_cam[:, 0], xyz_cam[:, 1], xyz_cam[:, 2]
        z = torch.clamp(z, min=0.001)
        x = x / z * focal_x + camera.image_width / 2.0
        y = y / z * focal_y + camera.image_height / 2.0
        in_screen = torch.logical_and(torch.logical_and(x >= -0.15 * camera.image_width, x <= camera.image_width * 1.15),
                                      torch.logical_and(y >= -0.15 * camera.image_height, y <= 1.15 * camera.image_height))
        valid = torch.logical_and(valid_depth, in_screen)
        distance[valid] = torch.min(distance[valid], z[valid])
        valid_points = torch.logical_or(valid_points, valid)
        if focal_length < focal_x:
            focal_length = focal_x
    distance[~valid_points] = distance[valid_points].max()
    filter_3D = distance / focal_length * (0.2 ** 0.5)
    return filter_3D[..., None], valid_points


@pytest.mark.parametrize("n_cams", [1, 7, 300])
def test_compute_3D_filter_matches_reference_loop(n_cams):
    import math
    from types import SimpleNamespace

    import numpy as np
    from rade_gs_b200.fused import compute_3D_filter

    rng = np.random.default_rng(n_cams)
    P = 40_013
    xyz = torch.from_numpy(rng.normal(size=(P, 3)).astype(np.float32) * 3.0).cuda()
    cams = []
    for k in range(n_cams):
        # random orthonormal R, camera a few units away looking roughly at the cloud
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        W, H = int(rng.integers(300, 1700)), int(rng.integers(200, 1300))
        cams.append(SimpleNamespace(R=q, T=rng.normal(size=3) * 0.5 + np.array([0, 0, 6.0]), FoVx=float(rng.uniform(0.5, 1.2)),
                                    FoVy=float(rng.uniform(0.4, 1.0)), image_width=W, image_height=H))
    ours = compute_3D_filter(xyz, cams)
    ref, valid = _ref_compute_3D_filter(xyz, cams)
    assert ours.shape == ref.shape == (P, 1)
    assert 0 < valid.sum().item() < P or n_cams > 1
    # a point within one ulp of a frustum / depth threshold may flip its validity against cuBLAS's summation order:
    # allow a handful of such points, everything else agrees to fp32 rounding
    close = torch.isclose(ours, ref, rtol=2e-6, atol=0)
    assert (~close).sum().item() <= max(2, P // 20000), (~close).sum().item()


def test_compute_3D_filter_raises_when_nothing_is_seen():
    from types import SimpleNamespace

    import numpy as np
    from rade_gs_b200.fused import compute_3D_filter

    xyz = torch.randn(100, 3, device="cuda")
    behind = SimpleNamespace(R=np.eye(3), T=np.array([0, 0, -50.0]), FoVx=0.8, FoVy=0.6, image_width=640, image_height=480)
    with pytest.raises(RuntimeError):
        compute_3D_filter(xyz, [behind])
