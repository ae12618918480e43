"""Stand-in for simple_knn._C.distCUDA2 (test infrastructure): mean squared distance to the 3 nearest neighbours, which is
what the reference's missing submodule computes for the initial scales (scene/gaussian_model.py:315)."""
import torch


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    P = points.shape[0]
    out = torch.empty(P, device=points.device, dtype=torch.float32)
    chunk = max(1, min(P, (256 << 20) // max(4 * P, 1)))
    for s in range(0, P, chunk):
        d2 = torch.cdist(points[s:s + chunk].double(), points.double()).pow(2)
        k = min(4, P)
        nearest = torch.topk(d2, k, dim=1, largest=False).values[:, 1:]   # drop the point itself
        out[s:s + chunk] = nearest.mean(dim=1).float() if k > 1 else 1.0
    return out
