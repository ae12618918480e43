"""`rade_gs_b200.renderer.render / integrate` keep the reference's signatures and result keys (checked against the reference's
source when it is present) and wire the fused pieces correctly (checked on CPU with the CUDA calls replaced by torch stand-ins)."""
import ast
import inspect
import math
import os
from types import SimpleNamespace

import pytest
import torch

REF = "/root/reference/gaussian_renderer/__init__.py"


def _ref_function(name):
    tree = ast.parse(open(REF).read())
    return next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")
@pytest.mark.parametrize("name", ["render", "integrate"])
def test_signature_and_result_keys_equal_the_reference(name):
    from rade_gs_b200 import renderer
    fn = _ref_function(name)
    ref_args = [a.arg for a in fn.args.args]
    ours = inspect.signature(getattr(renderer, name))
    assert list(ours.parameters) == ref_args
    n_def = len(fn.args.defaults)
    ref_defaults = [ast.literal_eval(d) for d in fn.args.defaults]
    our_defaults = [p.default for p in list(ours.parameters.values())[-n_def:]]
    assert our_defaults == ref_defaults
    ret = next(n for n in ast.walk(fn) if isinstance(n, ast.Return) and isinstance(n.value, ast.Dict))
    ref_keys = [k.value for k in ret.value.keys]
    src = ast.parse(inspect.getsource(getattr(renderer, name)))
    our_ret = next(n for n in ast.walk(src) if isinstance(n, ast.Return) and isinstance(n.value, ast.Dict))
    assert [k.value for k in our_ret.value.keys] == ref_keys


def test_render_wires_activations_split_sh_and_settings(monkeypatch):
    from rade_gs_b200 import renderer
    P, H, W = 7, 32, 48
    seen = {}

    class FakeRasterizer:
        def __init__(self, raster_settings):
            seen["settings"] = raster_settings

        def __call__(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
            seen.update(means2D=means2D, shs=shs, opacities=opacities, scales=scales, rotations=rotations)
            img = lambda c: torch.zeros(c, H, W) + opacities.sum() + scales.sum() + rotations.sum() + means3D.sum() + shs[0].sum() + shs[1].sum()  # noqa: E731
            return img(3), torch.arange(P, dtype=torch.int32) % 2, img(3), img(3), img(1), img(1), img(1), img(3)

    def fake_activate(raw_s, raw_o, raw_r, f3d):
        return torch.sqrt(torch.exp(raw_s) ** 2 + f3d ** 2), torch.sigmoid(raw_o), torch.nn.functional.normalize(raw_r)

    monkeypatch.setattr(renderer, "GaussianRasterizer", FakeRasterizer)
    monkeypatch.setattr(renderer.fused, "activate_gaussians", fake_activate)
    g = torch.Generator().manual_seed(0)
    pc = SimpleNamespace(_xyz=torch.randn(P, 3, generator=g).requires_grad_(True), _scaling=torch.randn(P, 3, generator=g).requires_grad_(True),
                         _opacity=torch.randn(P, 1, generator=g).requires_grad_(True), _rotation=torch.randn(P, 4, generator=g).requires_grad_(True),
                         _features_dc=torch.randn(P, 1, 3, generator=g).requires_grad_(True),
                         _features_rest=torch.randn(P, 15, 3, generator=g).requires_grad_(True), filter_3D=torch.full((P, 1), 0.01),
                         active_sh_degree=2)
    cam = SimpleNamespace(image_height=H, image_width=W, FoVx=0.9, FoVy=0.7, world_view_transform=torch.eye(4), full_proj_transform=torch.eye(4),
                          camera_center=torch.zeros(3))
    out = renderer.render(cam, pc, SimpleNamespace(debug=False), torch.zeros(3), 0.1, require_coord=False)
    s = seen["settings"]
    assert (s.image_height, s.image_width, s.sh_degree, s.kernel_size, s.require_coord, s.require_depth) == (H, W, 2, 0.1, False, True)
    assert abs(s.tanfovx - math.tan(0.45)) < 1e-12 and s.prefiltered is False and s.debug is False
    assert isinstance(seen["shs"], tuple) and seen["shs"][0] is pc._features_dc and seen["shs"][1] is pc._features_rest
    assert out["viewspace_points"] is seen["means2D"] and out["viewspace_points"].requires_grad
    assert torch.equal(out["visibility_filter"], out["radii"] > 0)
    out["render"].sum().backward()  # gradients reach every raw parameter through the (stand-in) activation
    for k in ("_xyz", "_scaling", "_opacity", "_rotation", "_features_dc", "_features_rest"):
        assert getattr(pc, k).grad is not None, k
