"""Live drop-in run: the reference's UNMODIFIED `train.py`, `render.py` and `gaussian_renderer.render()` executed against this
repo's `diff_gaussian_rasterization` (north_star: "so train.py / render.py drop in unchanged").

The reference's Python lives in /root/reference, which does not exist on the GPU box; `oracle/stage_ref.py` stages
byte-identical copies under oracle/_ref/refsrc/ (git-ignored, shipped like the reference build).  The third-party imports
this image lacks (plyfile, simple_knn, trimesh, open3d, matplotlib) are satisfied by tests/ref_stubs/.  Skipped only when the
staged copy is absent.
  1. a tiny Blender-format dataset is rendered with this rasterizer (known Gaussians, 10 cameras on a ring);
  2. `python train.py -s <data> -m <out> --iterations 60 ...` (densification at 20 / 40, depth-normal regularisation from 20, opacity
     reset at 45, the 3D filter recomputed after every densification) must run to completion, save a PLY and report a PSNR;
  3. `python render.py -m <out>` must write the renders;
  4. `gaussian_renderer.render()` on the trained model must return exactly the bits of a raw `_C.rasterize_gaussians` call on
     the same activated tensors, for the depth and the coordinate-map variant.
"""
import json
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

REFSRC = os.path.join(ROOT, "oracle", "_ref", "refsrc")
STUBS = os.path.join(ROOT, "tests", "ref_stubs")
PKG = os.path.join(ROOT, "rade-gs_b200")


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([STUBS, REFSRC, PKG, env.get("PYTHONPATH", "")])
    env["PYTHONUNBUFFERED"] = "1"
    return env


def _make_dataset(root, W=208, H=160, n_train=8, n_test=2):
    """Self-rendered Blender-format scene (scene/dataset_readers.py:245-321): PNGs + transforms_{train,test}.json."""
    from PIL import Image
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import rawapi, scenes
    g = torch.Generator().manual_seed(11)
    P = 4000
    means = (torch.rand(P, 3, generator=g) * 2 - 1) * 0.9
    scales = torch.exp(torch.randn(P, 3, generator=g) * 0.4 - 3.2)
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opac = torch.sigmoid(torch.randn(P, 1, generator=g) * 1.5 + 1.0)
    sh = torch.cat([torch.randn(P, 1, 3, generator=g) * 1.2, 0.1 * torch.randn(P, 15, 3, generator=g)], dim=1)
    fovx = 0.9
    focal = W / (2 * math.tan(fovx / 2))
    fovy = 2 * math.atan(H / (2 * focal))
    os.makedirs(os.path.join(root, "train"), exist_ok=True)
    os.makedirs(os.path.join(root, "test"), exist_ok=True)
    frames = {"train": [], "test": []}
    n = n_train + n_test
    for i in range(n):
        split = "train" if i < n_train else "test"
        ang = 2 * math.pi * i / n
        eye = (4.0 * math.cos(ang), 0.8 * math.sin(3 * ang), 4.0 * math.sin(ang))
        V = scenes.look_at_view(eye, (0.0, 0.0, 0.0))                    # world -> camera, +z forward, y down (COLMAP convention)
        view = V.t().contiguous()
        proj = (view @ scenes.projection_matrix(0.01, 100.0, fovx, fovy).t()).contiguous()
        sc = scenes.Scene(means, scales, rot, opac, sh.contiguous(), view, proj, view.inverse()[3, :3].contiguous(), torch.zeros(3), W, H,
                          math.tan(fovx / 2), math.tan(fovy / 2)).to("cuda")
        img = rawapi.forward(dgr._C, sc, False, False)["color"].clamp(0, 1)
        Image.fromarray((img.permute(1, 2, 0).cpu().numpy() * 255 + 0.5).astype(np.uint8), "RGB").save(os.path.join(root, split, f"r_{i}.png"))
        c2w = torch.linalg.inv(V.double())
        c2w[:3, 1:3] *= -1                                                # the reader flips these two axes back (OpenGL -> COLMAP)
        frames[split].append({"file_path": f"./{split}/r_{i}", "transform_matrix": c2w.tolist()})
    for split in ("train", "test"):
        with open(os.path.join(root, f"transforms_{split}.json"), "w") as f:
            json.dump({"camera_angle_x": fovx, "frames": frames[split]}, f)


_RENDER_CHECK = r'''
import json, sys, torch
from argparse import Namespace
from gaussian_renderer import render                      # the reference's own file, unmodified
from scene.gaussian_model import GaussianModel
from scene.cameras import MiniCam
from utils.graphics_utils import getProjectionMatrix
import diff_gaussian_rasterization as dgr
import math
ply = sys.argv[1]
pc = GaussianModel(3)
pc.load_ply(ply)
W, H, fovx = 208, 160, 0.9
fovy = 2 * math.atan(H / (2 * (W / (2 * math.tan(fovx / 2)))))
V = torch.tensor([[0.8, 0.0, -0.6, 0.1], [0.0, 1.0, 0.0, -0.05], [0.6, 0.0, 0.8, 4.0], [0.0, 0.0, 0.0, 1.0]], device="cuda")
view = V.t().contiguous()
proj = view @ getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1).cuda()
cam = MiniCam(W, H, fovy, fovx, 0.01, 100.0, view, proj)
pipe = Namespace(convert_SHs_python=False, compute_cov3D_python=False, debug=False)
bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
out = {}
E = torch.Tensor([])
for name, coord, depth, ks in (("depth", False, True, 0.0), ("coord", True, False, 0.1), ("none", False, False, 0.0)):
    with torch.no_grad():
        pkg = render(cam, pc, pipe, bg, ks, require_coord=coord, require_depth=depth)
        scales, opacity = pc.get_scaling_n_opacity_with_3D_filter
        raw = dgr._C.rasterize_gaussians(bg, pc.get_xyz, E, opacity, scales, pc.get_rotation, 1.0, E, cam.world_view_transform, cam.full_proj_transform,
                                         math.tan(fovx * 0.5), math.tan(fovy * 0.5), ks, H, W, pc.get_features, pc.active_sh_degree, cam.camera_center,
                                         False, coord, depth, False)
    keys = {"render": 1, "expected_coord": 2, "median_coord": 3, "mask": 4, "normal": 5, "expected_depth": 6, "median_depth": 7, "radii": 8}
    assert set(pkg) == set(keys) | {"viewspace_points", "visibility_filter"}, sorted(pkg)
    out[name] = {k: int((pkg[k] != raw[i]).sum()) for k, i in keys.items()}
    out[name]["nonzero"] = int((pkg["render"] != 0).sum())
    out[name]["visible"] = int(pkg["visibility_filter"].sum())
# and the autograd path through the reference's render(): gradients reach every parameter group
pkg = render(cam, pc, pipe, bg, 0.0, require_coord=False, require_depth=True)
(pkg["render"].mean() + 0.1 * pkg["expected_depth"].mean() + 0.1 * pkg["normal"].abs().mean()).backward()
out["grads_finite"] = all(bool(torch.isfinite(p.grad).all()) and bool((p.grad != 0).any())
                          for p in (pc._xyz, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation, pc._opacity))
out["viewspace_grad_nonzero"] = bool((pkg["viewspace_points"].grad != 0).any())
print("RESULT " + json.dumps(out))
'''


@pytest.fixture(scope="module")
def trained(tmp_path_factory):
    if not os.path.isfile(os.path.join(REFSRC, "train.py")):
        pytest.skip("staged reference sources (oracle/_ref/refsrc) not present: run `python oracle/stage_ref.py` where /root/reference exists")
    root = tmp_path_factory.mktemp("dropin")
    data, out = str(root / "data"), str(root / "out")
    _make_dataset(data)
    cmd = [sys.executable, os.path.join(REFSRC, "train.py"), "-s", data, "-m", out, "--iterations", "60", "--test_iterations", "60",
           "--save_iterations", "60", "--checkpoint_iterations", "60", "--densify_from_iter", "10", "--densification_interval", "20",
           "--densify_until_iter", "50", "--opacity_reset_interval", "45", "--regularization_from_iter", "20", "--eval"]
    # schedule: densify + prune at 20 and 40, depth-normal regularisation from 20, opacity reset at 45, no pruning after the reset (the
    # reference's compute_3D_filter fails on an EMPTY model -- gaussian_model.py:227 -- which a prune right after a reset produces)
    r = subprocess.run(cmd, cwd=REFSRC, env=_env(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "train.py failed:\n" + r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return data, out, r.stdout


def test_reference_train_py_runs_unchanged(trained):
    data, out, log = trained
    assert "Training complete." in log
    ply = os.path.join(out, "point_cloud", "iteration_60", "point_cloud.ply")
    assert os.path.isfile(ply) and os.path.getsize(ply) > 100_000
    assert os.path.isfile(os.path.join(out, "chkpnt60.pth"))
    m = re.search(r"\[ITER 60\] Evaluating train: L1 ([0-9.eE+-]+) PSNR ([0-9.eE+-]+)", log)
    assert m, log[-2000:]
    l1, psnr = float(m.group(1)), float(m.group(2))
    assert math.isfinite(l1) and math.isfinite(psnr) and psnr > 5.0, (l1, psnr)   # 60 iterations from a random cloud, opacities reset at 45: a sanity bar, not a quality claim


def test_reference_render_py_runs_unchanged(trained):
    data, out, _ = trained
    r = subprocess.run([sys.executable, os.path.join(REFSRC, "render.py"), "-m", out], cwd=REFSRC, env=_env(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "render.py failed:\n" + r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    from PIL import Image
    for split, n in (("train", 8), ("test", 2)):
        d = os.path.join(out, split, "ours_60", "renders")
        files = sorted(os.listdir(d))
        assert len(files) == n, (split, files)
        img = np.asarray(Image.open(os.path.join(d, files[0])))
        assert img.shape == (160, 208, 3) and img.max() > 0


def test_reference_render_function_matches_the_raw_extension_call(trained, tmp_path):
    data, out, _ = trained
    ply = os.path.join(out, "point_cloud", "iteration_60", "point_cloud.ply")
    script = tmp_path / "render_check.py"
    script.write_text(_RENDER_CHECK)
    r = subprocess.run([sys.executable, str(script), ply], cwd=REFSRC, env=_env(), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    for name in ("depth", "coord", "none"):
        v = res[name]
        assert v["nonzero"] > 1000 and v["visible"] > 100, (name, v)
        for k in ("render", "expected_coord", "median_coord", "mask", "normal", "expected_depth", "median_depth", "radii"):
            assert v[k] == 0, (name, k, v)       # render() is a thin wrapper: identical bits
    assert res["grads_finite"] and res["viewspace_grad_nonzero"]
