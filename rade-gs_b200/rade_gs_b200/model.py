"""A parameter container with `GaussianModel`'s attribute names (scene/gaussian_model.py:58-75, 515-559), enough to drive
`renderer.render` / `renderer.integrate` from a saved `point_cloud.ply` the way `render.py` does
without the reference's scene / dataset code.  It holds the raw (pre-activation) parameters; training policy
(densification, optimizer groups, appearance network) stays with the reference's `GaussianModel`."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ply_io


@dataclass
class GaussianParameters:
    _xyz: torch.Tensor            # [P,3]
    _features_dc: torch.Tensor    # [P,1,3]
    _features_rest: torch.Tensor  # [P,M-1,3]
    _opacity: torch.Tensor        # [P,1]  logit
    _scaling: torch.Tensor        # [P,3]  log
    _rotation: torch.Tensor       # [P,4]  unnormalised quaternion (r,x,y,z)
    filter_3D: torch.Tensor       # [P,1]
    max_sh_degree: int
    active_sh_degree: int

    @classmethod
    def from_ply(cls, path: str, device="cuda", max_sh_degree: int | None = None, requires_grad: bool = False) -> "GaussianParameters":
        """`GaussianModel.load_ply` (scene/gaussian_model.py:515-559): the active degree is the stored degree."""
        d = ply_io.load_gaussian_ply(path, max_sh_degree=max_sh_degree)

        def t(k):
            return torch.from_numpy(d[k]).to(device=device, dtype=torch.float32).contiguous().requires_grad_(requires_grad)

        deg = int(d["sh_degree"])
        return cls(t("xyz"), t("features_dc"), t("features_rest"), t("opacity"), t("scaling"), t("rotation"),
                   torch.from_numpy(d["filter_3D"]).to(device=device, dtype=torch.float32).contiguous(), deg, deg)

    def save_ply(self, path: str) -> None:
        ply_io.save_gaussian_ply(path, self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation,
                                 self.filter_3D)

    def to(self, device) -> "GaussianParameters":
        kw = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return GaussianParameters(**kw)

    @property
    def get_xyz(self) -> torch.Tensor:
        return self._xyz

    @property
    def num_points(self) -> int:
        return int(self._xyz.shape[0])

    def compute_3D_filter(self, cameras) -> None:
        """In place, like the reference's method (scene/gaussian_model.py:179-232)."""
        from . import fused
        self.filter_3D = fused.compute_3D_filter(self._xyz.detach(), cameras)
