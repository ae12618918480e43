"""Fused image-side losses of the training step (opt-in).

Drop-in for the reference's loss helpers, same names and argument meaning, computed by the kernels in
csrc/rgs_image_loss.cu instead of ~100 eager torch kernels per iteration:

    ssim(img1, img2)                              utils/loss_utils.py:35-63   (window 11, sigma 1.5, zero padding)
    l1_loss(network_output, gt)                   utils/loss_utils.py:17-18
    l1_ssim_loss(image, gt, lambda_dssim)         train.py:163  (1-l) * L1 + l * (1 - SSIM), one forward + one backward kernel
    depth_normal_consistency_loss(view, normal, expected_depth, median_depth)      train.py:143-156 with require_depth
    point_normal_consistency_loss(normal, expected_coord, median_coord)            train.py:143-156 with require_coord

All are differentiable w.r.t. the rendered maps (not w.r.t. the ground-truth image).  No CPU path: CUDA tensors only.
"""
from __future__ import annotations

import math

import torch

from diff_gaussian_rasterization import _C

__all__ = ["ssim", "l1_loss", "l1_ssim_loss", "depth_normal_consistency_loss", "point_normal_consistency_loss"]


class _SsimL1(torch.autograd.Function):
    """const + a_ssim * mean(SSIM map) + a_l1 * mean |img - gt|."""

    @staticmethod
    def forward(ctx, img, gt, a_ssim: float, a_l1: float, const: float):
        need_grad = ctx.needs_input_grad[0]
        sums, dmaps = _C.ssim_l1_forward(img, gt, need_grad)
        n = img.numel()
        ctx.weights = (a_ssim / n, a_l1 / n)
        if need_grad:
            ctx.save_for_backward(img, gt, dmaps)
        return (const + (a_ssim / n) * sums[0] + (a_l1 / n) * sums[1]).float()

    @staticmethod
    def backward(ctx, grad_out):
        img, gt, dmaps = ctx.saved_tensors
        w_ssim, w_l1 = ctx.weights
        return _C.ssim_l1_backward(img, gt, dmaps, w_ssim, w_l1, grad_out.float()), None, None, None, None


def _check_images(img1, img2):
    if img2.requires_grad:
        raise ValueError("the fused losses differentiate w.r.t. the first (rendered) image only")
    if not img1.is_cuda:
        raise RuntimeError("fused losses need CUDA tensors: there is no CPU path")


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """Mean SSIM of ``img1`` ([C,H,W] or [B,C,H,W]) against ``img2`` (same number of elements)."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("fused ssim covers the configuration the reference trains and evaluates with: window 11, size_average=True")
    _check_images(img1, img2)
    return _SsimL1.apply(img1, img2, 1.0, 0.0, 0.0)


def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    _check_images(network_output, gt)
    return _SsimL1.apply(network_output, gt, 0.0, 1.0, 0.0)


def l1_ssim_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float) -> torch.Tensor:
    """``(1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt))`` (train.py:163)."""
    _check_images(image, gt)
    return _SsimL1.apply(image, gt, -float(lambda_dssim), 1.0 - float(lambda_dssim), float(lambda_dssim))


class _NormalConsistency(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rendered_normal, map_expected, map_median, from_depth: bool, inv_fx: float, inv_fy: float, cx: float, cy: float,
                depth_ratio: float):
        H, W = rendered_normal.shape[-2:]
        n = H * W
        loss, d_normal, d_e, d_m = _C.normal_consistency(rendered_normal, map_expected, map_median, from_depth, inv_fx, inv_fy, cx, cy,
                                                         (1.0 - depth_ratio) / n, depth_ratio / n)
        ctx.save_for_backward(d_normal, d_e, d_m)
        return loss[0].float()

    @staticmethod
    def backward(ctx, grad_out):
        d_normal, d_e, d_m = ctx.saved_tensors
        return d_normal * grad_out, d_e * grad_out, d_m * grad_out, None, None, None, None, None, None


def depth_normal_consistency_loss(view, rendered_normal: torch.Tensor, expected_depth: torch.Tensor, median_depth: torch.Tensor,
                                  depth_ratio: float = 0.6) -> torch.Tensor:
    """``view`` needs ``FoVx``, ``FoVy`` (radians); image size is taken from the maps ([3,H,W] normal, [1,H,W] depths)."""
    H, W = rendered_normal.shape[-2:]
    fx = W / (2 * math.tan(view.FoVx / 2.))
    fy = H / (2 * math.tan(view.FoVy / 2.))
    # entries of intrins_inv (utils/graphics_utils.py:101-105)
    return _NormalConsistency.apply(rendered_normal, expected_depth, median_depth, True, 1 / fx, 1 / fy, -W / (2 * fx), -H / (2 * fy), depth_ratio)


def point_normal_consistency_loss(rendered_normal: torch.Tensor, expected_coord: torch.Tensor, median_coord: torch.Tensor,
                                  depth_ratio: float = 0.6) -> torch.Tensor:
    return _NormalConsistency.apply(rendered_normal, expected_coord, median_coord, False, 0.0, 0.0, 0.0, 0.0, depth_ratio)
