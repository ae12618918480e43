"""One rank's share of a sharded step, in one process on one GPU (for `ncu`, which must not wrap a multi-rank command):

    python tools/slab_step.py --config C2 --world 8 --rank 3 [--steps 2]

runs forward (slab), backward-render (slab), backward-preprocess -- everything a rank executes except the cross-rank exchange --
so that per-kernel durations and DRAM bytes of "what one of N ranks does" can be captured with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... python tools/slab_step.py ...
and turned into achieved HBM GB/s per N (tools/per_n_hbm.py)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
import torch  # noqa: E402

import diff_gaussian_rasterization as dgr  # noqa: E402
from rade_gs_b200 import multigpu, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
C = dgr._C
sc, coord, depth = scenes.make_config(a.config)
sc = sc.to("cuda")
g = scenes.make_upstream_grads(sc.height, sc.width, device="cuda")
gy = (sc.height + 15) // 16
b, e = multigpu.partition_tile_rows(gy, a.world)[a.rank]
E = torch.Tensor([])
for _ in range(a.steps):
    out = C.rasterize_gaussians_slab(sc.bg, sc.means3D, E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx, sc.tanfovy,
                                     0.0, sc.height, sc.width, sc.shs, 3, sc.campos, False, coord, depth, False, b, e)
    acc = C.rasterize_gaussians_backward_render(sc.bg, sc.means3D, out[8], E, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx,
                                                sc.tanfovy, 0.0, g["color"], g["coord"], g["mcoord"], g["depth"], g["mdepth"], g["alpha"], g["normal"], out[5],
                                                sc.shs, 3, sc.campos, out[9], out[0], out[10], out[11], out[4], coord, depth, False, b, e)
    gr = C.rasterize_gaussians_backward_preprocess(acc, sc.bg, sc.means3D, out[8], E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix,
                                                   sc.projmatrix, sc.tanfovx, sc.tanfovy, 0.0, sc.height, sc.width, sc.shs, 3, sc.campos, out[9], coord, depth,
                                                   False)
torch.cuda.synchronize()
print(f"{a.config} world {a.world} rank {a.rank}: tile rows {b}-{e}, num_rendered {out[0]}")
