"""Time `integrate` at mesh-extraction scale (C1: 300k splats, 800x800; 9 query points per splat) next to the reference build.
python tools/bench_integrate.py [P] [W] [H]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("rade-gs_b200", "oracle", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402

import diff_gaussian_rasterization as dgr  # noqa: E402
from gen_golden_integrate import call_reference, make_points  # noqa: E402
from rade_gs_b200 import scenes  # noqa: E402
from test_gpu_integrate import _settings  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 800
H = int(sys.argv[3]) if len(sys.argv) > 3 else 800
sc = scenes.make_scene(P, W, H, 1100.0 * W / 800, -4.3, seed=1234)
pts = make_points(sc, 9 * P, 99).cuda()
sc = sc.to("cuda")
rast = dgr.GaussianRasterizer(_settings(dgr, sc=sc))


def ours():
    return rast.integrate(points3D=pts, means3D=sc.means3D, means2D=None, opacities=sc.opacities, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


res = {"P": P, "PN": int(pts.shape[0]), "image": [H, W], "ours_ms": timed(ours, 5)}
o = ours()
res["points_touched"] = int((o[4] != -1000.0).sum())
res["max_points_per_pixel"] = int(o[0][8].max())
try:
    import build_ref
    ref = build_ref.load()
    res["reference_ms"] = timed(lambda: call_reference(ref, sc, pts, 3), 2)
    r = call_reference(ref, sc, pts, 3)
    res["max_abs_diff_alpha"] = float((r[2] - o[1]).abs().max())
    res["frac_alpha_diff_gt_1e-4"] = float(((r[2] - o[1]).abs() > 1e-4).float().mean())
except Exception as e:  # reference build absent
    res["reference_ms"] = None
    res["reference_error"] = str(e)[:200]
print(json.dumps(res))
