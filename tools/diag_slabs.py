"""Whole image vs K tile-row slabs on ONE GPU (the multi-GPU forward without NCCL): where do they differ?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
import torch  # noqa: E402

import diff_gaussian_rasterization as dgr  # noqa: E402
from rade_gs_b200 import multigpu, scenes  # noqa: E402

C = dgr._C
sc = scenes.make_scene(60000, 640, 400, 500.0, -3.8, seed=21, view=scenes.look_at_view((0.3, 0.2, -0.4), (0.0, 0.1, 6.0)), bg=(0.2, 0.1, 0.3)).to("cuda")
E = torch.Tensor([])
args = (sc.bg, sc.means3D, E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx, sc.tanfovy, 0.1, sc.height,
        sc.width, sc.shs, 3, sc.campos, False, True, True, False)
whole = C.rasterize_gaussians(*args)
R = whole[0]
wl = whole[10][: 4 * R].view(torch.int32).clone()
gy = (sc.height + 15) // 16
for K in (2, 3, 8):
    slabs = multigpu.partition_tile_rows(gy, K)
    tot, lists, bad = 0, [], []
    for (b, e) in slabs:
        s = C.rasterize_gaussians_slab(*args, b, e)
        tot += s[0]
        lists.append(s[10][: 4 * s[0]].view(torch.int32).clone())
        for idx, name in ((1, "color"), (4, "alpha"), (6, "depth"), (5, "normal"), (2, "coord")):
            d = (s[idx][:, b * 16: e * 16] != whole[idx][:, b * 16: e * 16]).any(0)
            if d.any():
                ys, xs = torch.nonzero(d, as_tuple=True)
                bad.append((name, (b, e), int(d.sum()), [(int(y) + b * 16, int(x)) for y, x in zip(ys[:4], xs[:4])],
                            float((s[idx][:, b * 16: e * 16] - whole[idx][:, b * 16: e * 16]).abs().max())))
            out = s[idx].clone()
            out[:, b * 16: e * 16] = 0
            if out.any():
                bad.append((name, (b, e), "nonzero outside slab", int((out != 0).sum())))
    cat = torch.cat(lists)
    print(f"K={K} slabs={slabs} sumR={tot} R={R} lists_equal={bool(cat.numel() == wl.numel() and torch.equal(cat, wl))} issues={bad[:6]}")
# run-to-run determinism of the whole image
w2 = C.rasterize_gaussians(*args)
print("whole deterministic:", all(torch.equal(a, b) for a, b in zip(whole[1:9], w2[1:9])))
