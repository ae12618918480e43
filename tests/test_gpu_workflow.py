"""The pieces in concert, the way train.py / render.py use them: a short optimisation must reduce its
loss, and a model saved to PLY and loaded back must render the same image."""
import math
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(sc, filter_3D):
    return {
        "xyz": sc.means3D.clone().requires_grad_(True),
        "scaling": torch.log(sc.scales).requires_grad_(True),
        "opacity": torch.logit(sc.opacities.clamp(1e-4, 1 - 1e-4)).requires_grad_(True),
        "rotation": sc.rotations.clone().requires_grad_(True),
        "f_dc": sc.shs[:, :1].contiguous().requires_grad_(True),
        "f_rest": sc.shs[:, 1:].contiguous().requires_grad_(True),
    }


def _render(dgr, fused, settings, m, filter_3D, means2D=None):
    s, o, r = fused.activate_gaussians(m["scaling"], m["opacity"], m["rotation"], filter_3D)
    if means2D is None:
        means2D = torch.zeros_like(m["xyz"])
    return dgr.GaussianRasterizer(settings)(means3D=m["xyz"], means2D=means2D, opacities=o, shs=(m["f_dc"], m["f_rest"]), scales=s, rotations=r)


def test_short_optimisation_reduces_the_training_loss():
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import fused, losses, scenes
    from test_gpu_api import _settings

    sc = scenes.make_scene(20_000, 256, 192, 220.0, -2.6, seed=3).to(DEV)
    view = SimpleNamespace(FoVx=2 * math.atan(sc.tanfovx), FoVy=2 * math.atan(sc.tanfovy))
    settings = _settings(dgr, sc, False, True, ks=0.1)
    P = sc.means3D.shape[0]
    filter_3D = torch.full((P, 1), 1e-3, device=DEV)
    with torch.no_grad():  # target: the same scene with other colours and slightly different geometry
        tgt = _model(scenes.make_scene(20_000, 256, 192, 220.0, -2.6, seed=3, sh_rest_std=0.3).to(DEV), filter_3D)
        tgt["f_dc"] = tgt["f_dc"] + 0.8
        gt = _render(dgr, fused, settings, tgt, filter_3D)[0].clamp(0, 1)
    m = _model(sc, filter_3D)
    opt = torch.optim.Adam([{"params": [m["f_dc"]], "lr": 0.05}, {"params": [m["f_rest"]], "lr": 0.0025}, {"params": [m["opacity"]], "lr": 0.05},
                            {"params": [m["scaling"]], "lr": 0.005}, {"params": [m["rotation"]], "lr": 0.001}, {"params": [m["xyz"]], "lr": 1e-4}],
                           eps=1e-15)
    stats = [torch.zeros(P, 1, device=DEV) for _ in range(4)]
    max_radii = torch.zeros(P, device=DEV)
    history = []
    for it in range(40):
        means2D = torch.zeros(P, 3, device=DEV, requires_grad=True)
        color, radii, _, _, d, md, alpha, normal = _render(dgr, fused, settings, m, filter_3D, means2D)
        loss = losses.l1_ssim_loss(color, gt, 0.2) + 0.05 * losses.depth_normal_consistency_loss(view, normal, d, md)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        fused.add_densification_stats_(means2D.grad, radii, *stats, max_radii)
        opt.step()
        history.append(loss.item())
        assert math.isfinite(history[-1])
    assert history[-1] < 0.7 * history[0], history[::8]
    assert stats[3].max().item() == 40 and (max_radii > 0).sum() == (stats[3][:, 0] > 0).sum()
    for k, t in m.items():
        assert torch.isfinite(t).all(), k


def test_model_saved_to_ply_renders_identically_after_loading(tmp_path):
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import fused, ply_io, scenes
    from test_gpu_api import _settings

    sc = scenes.make_scene(5000, 160, 120, 140.0, -2.4, seed=8).to(DEV)
    settings = _settings(dgr, sc, False, True, ks=0.1)
    cams = [SimpleNamespace(R=sc.viewmatrix[:3, :3].cpu().numpy(), T=sc.viewmatrix[3, :3].cpu().numpy(), FoVx=2 * math.atan(sc.tanfovx),
                            FoVy=2 * math.atan(sc.tanfovy), image_width=sc.width, image_height=sc.height)]
    filter_3D = fused.compute_3D_filter(sc.means3D, cams)
    m = {k: v.detach() for k, v in _model(sc, filter_3D).items()}
    with torch.no_grad():
        before = _render(dgr, fused, settings, m, filter_3D)
    path = os.path.join(tmp_path, "point_cloud", "iteration_7000", "point_cloud.ply")
    ply_io.save_gaussian_ply(path, m["xyz"], m["f_dc"], m["f_rest"], m["opacity"], m["scaling"], m["rotation"], filter_3D)
    back = ply_io.load_gaussian_ply(path, max_sh_degree=3)
    t = lambda k: torch.from_numpy(back[k]).to(DEV)  # noqa: E731
    m2 = {"xyz": t("xyz"), "scaling": t("scaling"), "opacity": t("opacity"), "rotation": t("rotation"), "f_dc": t("features_dc"),
          "f_rest": t("features_rest")}
    with torch.no_grad():
        after = _render(dgr, fused, settings, m2, t("filter_3D"))
    for a, b in zip(before, after):
        assert torch.equal(a, b)
    assert (before[1] > 0).sum() > 1000


def test_fused_render_entry_point_matches_the_reference_style_call():
    """`rade_gs_b200.renderer.render` (fused activations, split SH) against the call sequence of the reference's render()
    (eager activations, concatenated SH): same maps bit for bit up to the activation's last-bit differences, same gradients."""
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import renderer, scenes
    from test_gpu_api import _settings
    from test_gpu_fused import _ref_activate
    from tolerances import grad_close_gpu

    sc = scenes.make_scene(20_000, 256, 192, 220.0, -2.6, seed=13).to(DEV)
    P = sc.means3D.shape[0]
    filter_3D = torch.full((P, 1), 2e-3, device=DEV)

    def make_pc():
        m = _model(sc, filter_3D)
        return SimpleNamespace(_xyz=m["xyz"], _scaling=m["scaling"], _opacity=m["opacity"], _rotation=m["rotation"], _features_dc=m["f_dc"],
                               _features_rest=m["f_rest"], filter_3D=filter_3D, active_sh_degree=3)

    cam = SimpleNamespace(image_height=sc.height, image_width=sc.width, FoVx=2 * math.atan(sc.tanfovx), FoVy=2 * math.atan(sc.tanfovy),
                          world_view_transform=sc.viewmatrix, full_proj_transform=sc.projmatrix, camera_center=sc.campos)
    pipe = SimpleNamespace(debug=False)
    g = scenes.make_upstream_grads(sc.height, sc.width, seed=3, device=DEV)

    def loss_of(pkg):
        return (pkg["render"] * g["color"]).sum() + (pkg["expected_depth"] * g["depth"]).sum() + (pkg["normal"] * g["normal"]).sum() + \
            (pkg["mask"] * g["alpha"]).sum()

    pc = make_pc()
    pkg = renderer.render(cam, pc, pipe, sc.bg, 0.1, require_coord=False, require_depth=True)
    assert set(pkg) == {"render", "mask", "expected_coord", "median_coord", "expected_depth", "median_depth", "viewspace_points",
                        "visibility_filter", "radii", "normal"}
    loss_of(pkg).backward()

    ref = make_pc()
    s, o, r = _ref_activate(ref._scaling, ref._opacity, ref._rotation, filter_3D)
    means2D = torch.zeros(P, 3, device=DEV, requires_grad=True)
    out = dgr.GaussianRasterizer(_settings(dgr, sc, False, True, ks=0.1))(
        means3D=ref._xyz, means2D=means2D, opacities=o, shs=torch.cat((ref._features_dc, ref._features_rest), dim=1), scales=s, rotations=r)
    ref_pkg = {"render": out[0], "expected_depth": out[4], "normal": out[7], "mask": out[6]}
    loss_of(ref_pkg).backward()

    assert torch.equal(pkg["radii"], out[1])
    for k in ("render", "expected_depth", "normal", "mask"):
        d = (pkg[k] - ref_pkg[k]).abs()
        assert (d > 1e-4 + 1e-4 * ref_pkg[k].abs()).float().mean().item() < 1e-4 and d.max().item() < 8e-3, (k, d.max().item())
    for name in ("_xyz", "_scaling", "_opacity", "_rotation", "_features_dc", "_features_rest"):
        grad_close_gpu(getattr(pc, name).grad.cpu().numpy(), getattr(ref, name).grad.cpu().numpy(), name, rel=2e-3, elem=1e-2)
    grad_close_gpu(pkg["viewspace_points"].grad.cpu().numpy(), means2D.grad.cpu().numpy(), "viewspace_points", rel=2e-3, elem=1e-2)
