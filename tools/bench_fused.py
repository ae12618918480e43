"""Time the fused activation / statistics kernels against the reference's eager-torch expressions (1M Gaussians)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rade_gs_b200.fused import activate_gaussians, add_densification_stats_  # noqa: E402
from test_gpu_fused import _raw, _ref_activate  # noqa: E402
from test_gpu_losses import _ref_l1, _ref_normal_loss, _ref_points_from_depth, _ref_ssim  # noqa: E402
from rade_gs_b200 import losses  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    raw = _raw(P, 7)
    g = [torch.randn(P, k, device="cuda") for k in (3, 1, 4)]

    def step(fn):
        leaves = [t.detach().requires_grad_(True) for t in raw[:3]]
        out = fn(*leaves, raw[3])
        torch.autograd.backward(list(out), g)

    grad = torch.randn(P, 3, device="cuda")
    radii = torch.randint(0, 30, (P,), dtype=torch.int32, device="cuda")
    st = [torch.zeros(P, 1, device="cuda") for _ in range(4)]
    mr = torch.zeros(P, device="cuda")

    def ref_stats():
        vis = radii > 0
        mr[vis] = torch.max(mr[vis], radii[vis])
        st[0][vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
        st[1][vis] += torch.norm(grad[vis, 2:], dim=-1, keepdim=True)
        st[2][vis] = torch.max(st[2][vis], torch.norm(grad[vis, 2:], dim=-1, keepdim=True))
        st[3][vis] += 1

    # image-side losses at the C2 image size (1600 x 1200)
    import math
    from types import SimpleNamespace
    H, W = 1200, 1600
    gt = torch.rand(3, H, W, device="cuda")
    img = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)
    view = SimpleNamespace(FoVx=0.9, FoVy=0.7)
    depth = 3 + torch.rand(1, H, W, device="cuda")
    mdepth = depth + 0.01 * torch.randn_like(depth)
    normal = torch.nn.functional.normalize(torch.randn(3, H, W, device="cuda"), dim=0)

    def photo(fused):
        a = img.detach().requires_grad_(True)
        if fused:
            l = losses.l1_ssim_loss(a, gt, 0.2)
        else:
            l = 0.8 * _ref_l1(a, gt) + 0.2 * (1.0 - _ref_ssim(a, gt.unsqueeze(0)))
        l.backward()

    def nrm(fused):
        leaves = [t.detach().requires_grad_(True) for t in (normal, depth, mdepth)]
        if fused:
            l = losses.depth_normal_consistency_loss(view, *leaves)
        else:
            l = _ref_normal_loss(leaves[0], *_ref_points_from_depth(view, H, W, leaves[1], leaves[2]))
        l.backward()

    res = {
        "P": P,
        "image": [H, W],
        "l1_ssim_fwd_bwd_ms": {"fused": timed(lambda: photo(True)), "torch_eager": timed(lambda: photo(False))},
        "normal_consistency_fwd_bwd_ms": {"fused": timed(lambda: nrm(True)), "torch_eager": timed(lambda: nrm(False))},
        "activate_fwd_bwd_ms": {"fused": timed(lambda: step(activate_gaussians)), "torch_eager": timed(lambda: step(_ref_activate))},
        "densification_stats_ms": {"fused": timed(lambda: add_densification_stats_(grad, radii, *st, mr)), "torch_eager": timed(ref_stats)},
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
