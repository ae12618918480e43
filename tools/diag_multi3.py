"""Cross-rank consistency probe (torchrun, N ranks): are the INPUTS identical on all ranks?  the preprocess records?  the whole renders?
the slab renders?  Prints one line per check; rank 0 reports where things differ (which tensors / record columns / how many rows).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/diag_multi3.py
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import diff_gaussian_rasterization as dgr  # noqa: E402
from rade_gs_b200 import multigpu, rawapi, scenes  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
C = dgr._C
sc_cpu = scenes.make_scene(60000, 640, 400, 500.0, -3.8, seed=21, view=scenes.look_at_view((0.3, 0.2, -0.4), (0.0, 0.1, 6.0)), bg=(0.2, 0.1, 0.3))
sc = sc_cpu.to(dev)
print(f"[rank {rank}] torch threads {torch.get_num_threads()} OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}", flush=True)


def same_everywhere(t, name):
    """number of elements of `t` that differ from rank 0's copy (bitwise)"""
    t = t.contiguous()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    n = int((t.view(torch.uint8) != ref.view(torch.uint8)).view(-1, t.element_size()).any(1).sum()) if t.numel() else 0
    tot = torch.tensor([n], device=dev)
    dist.all_reduce(tot)
    if rank == 0:
        print(f"  {name:28s}: {int(tot)} elements differ from rank 0's across the other ranks", flush=True)
    return n


for k in ("means3D", "scales", "rotations", "opacities", "shs", "viewmatrix", "projmatrix", "campos", "bg"):
    same_everywhere(getattr(sc, k), "input " + k)
coord, depth, ks = True, True, 0.1
f = rawapi.forward(C, sc, coord, depth, kernel_size=ks)
v = rawapi.ours_views(f, sc)
vis = f["radii"] > 0
same_everywhere(f["radii"], "radii")
rec = v["records"].clone()
rec[~vis] = 0
nrec = same_everywhere(rec, "records (visible rows)")
if nrec:
    ref = rec.clone()
    dist.broadcast(ref, src=0)
    cols = (rec.view(torch.int32) != ref.view(torch.int32)).sum(0).tolist()
    rows = int((rec.view(torch.int32) != ref.view(torch.int32)).any(1).sum())
    print(f"[rank {rank}] record columns differing from rank 0 (count per column): {cols}; rows {rows}; max abs {float((rec - ref).abs().max()):.3e}", flush=True)
same_everywhere(v["depths"] * vis, "depths")
for k in ("color", "alpha", "depth", "normal", "coord"):
    same_everywhere(f[k], "whole " + k)
gy = (sc.height + 15) // 16
b, e = multigpu.partition_tile_rows(gy, world)[rank]
E = torch.Tensor([])
s = C.rasterize_gaussians_slab(sc.bg, sc.means3D, E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx, sc.tanfovy, ks,
                               sc.height, sc.width, sc.shs, 3, sc.campos, False, coord, depth, False, b, e)
names = {1: "color", 4: "alpha", 6: "depth", 5: "normal", 2: "coord"}
for i, n in names.items():
    own = int((s[i][:, b * 16:e * 16] != f[n][:, b * 16:e * 16]).sum())
    print(f"[rank {rank}] slab {n} vs own whole: {own} mismatches", flush=True)
dist.barrier()
dist.destroy_process_group()
