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

    res = {
        "P": P,
        "activate_fwd_bwd_ms": {"fused": timed(lambda: step(activate_gaussians)), "torch_eager": timed(lambda: step(_ref_activate))},
        "densification_stats_ms": {"fused": timed(lambda: add_densification_stats_(grad, radii, *st, mr)), "torch_eager": timed(ref_stats)},
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
