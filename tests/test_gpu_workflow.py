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
