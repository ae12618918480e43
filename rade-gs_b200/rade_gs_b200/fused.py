"""Fused per-Gaussian arithmetic either side of the rasterizer call (opt-in).

The reference computes the activated scales / opacity / rotations with a dozen element-wise torch kernels per
iteration (plus about twenty in autograd's backward), and the densification statistics with five masked-index
passes.  At 1M Gaussians each of those is a full HBM round trip of a [P, k] tensor; the three kernels behind this
module (csrc/rgs_activation.cu) do one round trip each.

    scales, opacity = activate_scaling_n_opacity(pc._scaling, pc._opacity, pc.filter_3D)   # + rotations below
    scales, opacity, rotations = activate_gaussians(pc._scaling, pc._opacity, pc._rotation, pc.filter_3D)
    add_densification_stats_(viewspace_points.grad, radii, pc.xyz_gradient_accum, pc.xyz_gradient_accum_abs,
                             pc.xyz_gradient_accum_abs_max, pc.denom, pc.max_radii2D)

Reference formulas: scene/gaussian_model.py:156-166 (get_scaling_n_opacity_with_3D_filter), :125-126 (get_rotation ->
F.normalize), :743-747 (add_densification_stats), train.py:187-188 (max_radii2D update, update_filter = radii > 0).
There is no CPU path: the calls raise if the CUDA extension is missing or the tensors are not CUDA tensors.
"""
from __future__ import annotations

import torch

from diff_gaussian_rasterization import _C


class _ActivateGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw_scaling, raw_opacity, raw_rotation, filter_3D):
        scales, opacity, rotations = _C.activate_forward(raw_scaling, raw_opacity, raw_rotation, filter_3D)
        ctx.save_for_backward(raw_scaling, raw_opacity, raw_rotation, filter_3D)
        return scales, opacity, rotations

    @staticmethod
    def backward(ctx, g_scales, g_opacity, g_rotations):
        raw_scaling, raw_opacity, raw_rotation, filter_3D = ctx.saved_tensors
        # autograd hands None-free zero tensors for unused outputs because the outputs are not marked non-differentiable
        d_s, d_o, d_r = _C.activate_backward(raw_scaling, raw_opacity, raw_rotation, filter_3D, g_scales, g_opacity, g_rotations)
        return d_s, d_o, d_r, None  # filter_3D is a buffer computed under no_grad in the reference (compute_3D_filter)


def activate_gaussians(raw_scaling: torch.Tensor, raw_opacity: torch.Tensor, raw_rotation: torch.Tensor, filter_3D: torch.Tensor):
    """(scales [P,3], opacity [P,1], rotations [P,4]) = GaussianModel.get_scaling_n_opacity_with_3D_filter + get_rotation."""
    return _ActivateGaussians.apply(raw_scaling, raw_opacity, raw_rotation, filter_3D)


@torch.no_grad()
def add_densification_stats_(means2D_grad: torch.Tensor, radii: torch.Tensor, xyz_gradient_accum: torch.Tensor,
                             xyz_gradient_accum_abs: torch.Tensor, xyz_gradient_accum_abs_max: torch.Tensor, denom: torch.Tensor,
                             max_radii2D: torch.Tensor | None = None) -> None:
    """In-place train.py:187-188 for the Gaussians with ``radii > 0``; ``max_radii2D=None`` skips the radius maximum."""
    empty = torch.empty(0, device=means2D_grad.device)
    _C.densification_stats(means2D_grad, radii, xyz_gradient_accum, xyz_gradient_accum_abs, xyz_gradient_accum_abs_max, denom,
                           empty if max_radii2D is None else max_radii2D)


@torch.no_grad()
def compute_3D_filter(xyz: torch.Tensor, cameras) -> torch.Tensor:
    """``GaussianModel.compute_3D_filter`` (scene/gaussian_model.py:179-232): returns ``filter_3D`` [P,1].

    ``cameras`` is the iterable the reference passes (objects with ``R``, ``T``, ``FoVx``, ``FoVy``, ``image_width``,
    ``image_height``).  One kernel walks all cameras per point instead of ~15 torch kernels per camera (the camera
    table streams through shared memory 256 cameras at a time).
    """
    import math

    import numpy as np

    rows, focal_length = [], 0.0
    for cam in cameras:
        W, H = cam.image_width, cam.image_height
        focal_x = W / (2 * math.tan(cam.FoVx / 2.))
        focal_y = H / (2 * math.tan(cam.FoVy / 2.))
        R = np.asarray(cam.R, dtype=np.float32).reshape(9)
        T = np.asarray(cam.T, dtype=np.float32).reshape(3)
        rows.append(np.concatenate([R, T, np.asarray([focal_x, focal_y, W, H], dtype=np.float32)]))
        focal_length = max(focal_length, focal_x)
    if not rows:
        raise ValueError("compute_3D_filter needs at least one camera")
    table = torch.from_numpy(np.stack(rows)).to(xyz.device)
    out, mx = _C.compute_3d_filter(xyz, table, focal_length)
    if float(mx) == 0.0:  # the reference fails in distance[valid_points].max() on an empty selection
        raise RuntimeError("compute_3D_filter: no point is seen by any camera")
    return out
