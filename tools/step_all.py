"""One training-step slice at a config, every fused piece in the order train.py runs them:
activate -> rasterize (split SH) -> L1+SSIM + normal consistency -> backward -> densification statistics.
Used under ncu for the per-kernel table (tools/gpu_kernel_table.sh) and timed end to end against the eager
equivalents.  python tools/step_all.py [cfg] [iters] [--eager]"""
import math
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import diff_gaussian_rasterization as dgr  # noqa: E402
from rade_gs_b200 import fused, losses, scenes  # noqa: E402
from test_gpu_api import _settings  # noqa: E402
from test_gpu_fused import _ref_activate  # noqa: E402
from test_gpu_losses import _ref_l1, _ref_normal_loss, _ref_points_from_depth, _ref_ssim  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
cfg = args[0] if args else "C2"
iters = int(args[1]) if len(args) > 1 else 3
eager = "--eager" in sys.argv

sc, coord, depth = scenes.make_config(cfg)
sc = sc.to("cuda:0")
P = sc.means3D.shape[0]
view = SimpleNamespace(FoVx=2 * math.atan(sc.tanfovx), FoVy=2 * math.atan(sc.tanfovy))
filter_3D = torch.full((P, 1), 1e-3, device="cuda")
raw = {
    "xyz": sc.means3D.clone().requires_grad_(True),
    "scaling": torch.log(sc.scales).requires_grad_(True),
    "opacity": torch.logit(sc.opacities.clamp(1e-4, 1 - 1e-4)).requires_grad_(True),
    "rotation": sc.rotations.clone().requires_grad_(True),
    "f_dc": sc.shs[:, :1].contiguous().requires_grad_(True),
    "f_rest": sc.shs[:, 1:].contiguous().requires_grad_(True),
}
gt = torch.rand(3, sc.height, sc.width, device="cuda")
stats = [torch.zeros(P, 1, device="cuda") for _ in range(4)]
max_radii = torch.zeros(P, device="cuda")
settings = _settings(dgr, sc, coord, True, ks=0.1)


def step():
    for t in raw.values():
        t.grad = None
    means2D = torch.zeros(P, 3, device="cuda", requires_grad=True)
    if eager:
        s, o, r = _ref_activate(raw["scaling"], raw["opacity"], raw["rotation"], filter_3D)
        shs = torch.cat((raw["f_dc"], raw["f_rest"]), dim=1)
    else:
        s, o, r = fused.activate_gaussians(raw["scaling"], raw["opacity"], raw["rotation"], filter_3D)
        shs = (raw["f_dc"], raw["f_rest"])
    color, radii, _, _, d, md, alpha, normal = dgr.GaussianRasterizer(settings)(
        means3D=raw["xyz"], means2D=means2D, opacities=o, shs=shs, scales=s, rotations=r)
    if eager:
        loss = 0.8 * _ref_l1(color, gt) + 0.2 * (1.0 - _ref_ssim(color, gt.unsqueeze(0)))
        loss = loss + 0.05 * _ref_normal_loss(normal, *_ref_points_from_depth(view, sc.height, sc.width, d, md))
    else:
        loss = losses.l1_ssim_loss(color, gt, 0.2) + 0.05 * losses.depth_normal_consistency_loss(view, normal, d, md)
    loss.backward()
    with torch.no_grad():
        if eager:
            vis = radii > 0
            max_radii[vis] = torch.max(max_radii[vis], radii[vis])
            g = means2D.grad
            stats[0][vis] += torch.norm(g[vis, :2], dim=-1, keepdim=True)
            stats[1][vis] += torch.norm(g[vis, 2:], dim=-1, keepdim=True)
            stats[2][vis] = torch.max(stats[2][vis], torch.norm(g[vis, 2:], dim=-1, keepdim=True))
            stats[3][vis] += 1
        else:
            fused.add_densification_stats_(means2D.grad, radii, *stats, max_radii)
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    loss = step()
b.record()
torch.cuda.synchronize()
print({"cfg": cfg, "mode": "eager-around-raster" if eager else "fused", "ms_per_step": round(a.elapsed_time(b) / iters, 4), "loss": float(loss)})
