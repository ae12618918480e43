"""`render()` / `integrate()` with the reference's signatures and result dictionaries (gaussian_renderer/__init__.py:19-95,
98-195), on the fused fast paths (opt-in -- the reference's own `gaussian_renderer` keeps working with
the drop-in extension):

  * scales / opacity / rotations come from ONE kernel over the raw parameters (`fused.activate_gaussians`) instead of
    `pc.get_scaling_n_opacity_with_3D_filter` + `pc.get_rotation` (~12 element-wise kernels, ~20 more in backward);
  * the SH coefficients are handed over as the model's two tensors (`shs=(pc._features_dc, pc._features_rest)`) instead of
    `pc.get_features` (a 192 B/Gaussian `torch.cat` per iteration and the split of its gradient).

`pc` is duck-typed: it needs `_xyz` (or `get_xyz`), `_scaling`, `_opacity`, `_rotation`, `_features_dc`, `_features_rest`,
`filter_3D`, `active_sh_degree` -- the attributes `GaussianModel` has (scene/gaussian_model.py:58-75).
"""
from __future__ import annotations

import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer

from . import fused


def _settings(viewpoint_camera, pc, pipe, bg_color, kernel_size, scaling_modifier, require_coord, require_depth):
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        kernel_size=kernel_size,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        require_depth=require_depth,
        require_coord=require_coord,
        debug=bool(getattr(pipe, "debug", False)),
    )


def _xyz(pc):
    return pc.get_xyz if hasattr(pc, "get_xyz") else pc._xyz


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, kernel_size, scaling_modifier=1.0, require_coord: bool = True,
           require_depth: bool = True):
    """Same arguments, same dictionary as the reference's `render` (gaussian_renderer/__init__.py:19-95)."""
    means3D = _xyz(pc)
    screenspace_points = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device=means3D.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:  # no graph (torch.no_grad()): nothing to retain
        pass
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pc, pipe, bg_color, kernel_size, scaling_modifier,
                                                              require_coord, require_depth))
    scales, opacity, rotations = fused.activate_gaussians(pc._scaling, pc._opacity, pc._rotation, pc.filter_3D)
    (rendered_image, radii, rendered_expected_coord, rendered_median_coord, rendered_expected_depth, rendered_median_depth, rendered_alpha,
     rendered_normal) = rasterizer(means3D=means3D, means2D=screenspace_points, shs=(pc._features_dc, pc._features_rest), colors_precomp=None,
                                   opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None)
    return {"render": rendered_image,
            "mask": rendered_alpha,
            "expected_coord": rendered_expected_coord,
            "median_coord": rendered_median_coord,
            "expected_depth": rendered_expected_depth,
            "median_depth": rendered_median_depth,
            "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0,
            "radii": radii,
            "normal": rendered_normal,
            }


@torch.no_grad()
def integrate(points3D, viewpoint_camera, pc, pipe, bg_color: torch.Tensor, kernel_size, scaling_modifier=1.0, override_color=None):
    """Same arguments, same dictionary as the reference's `integrate` (gaussian_renderer/__init__.py:98-195).  The reference
    takes the two activations from separate properties here (`get_opacity_with_3D_filter`, `get_scaling_with_3D_filter`); they
    are the same two tensors `get_scaling_n_opacity_with_3D_filter` returns (scene/gaussian_model.py:114-166)."""
    means3D = _xyz(pc)
    rasterizer = GaussianRasterizer(raster_settings=_settings(viewpoint_camera, pc, pipe, bg_color, kernel_size, scaling_modifier, True, True))
    scales, opacity, rotations = fused.activate_gaussians(pc._scaling, pc._opacity, pc._rotation, pc.filter_3D)
    if override_color is None:
        shs, colors_precomp = torch.cat((pc._features_dc, pc._features_rest), dim=1), None  # integrate takes the concatenated layout
    else:
        shs, colors_precomp = None, override_color
    rendered_image, alpha_integrated, color_integrated, point_coordinate, point_sdf, radii = rasterizer.integrate(
        points3D=points3D, means3D=means3D, means2D=None, shs=shs, colors_precomp=colors_precomp, opacities=opacity, scales=scales,
        rotations=rotations, cov3D_precomp=None, view2gaussian_precomp=None)
    return {"render": rendered_image,
            "alpha_integrated": alpha_integrated,
            "color_integrated": color_integrated,
            "point_coordinate": point_coordinate,
            "point_sdf": point_sdf,
            "visibility_filter": radii > 0,
            "radii": radii}
