"""Does any result depend on memory the kernels did not write themselves?

Every scratch buffer and every empty()-allocated output is pre-filled with a byte pattern (`_C.set_poison`) before the
kernels run; the whole-image render, K tile-row slab renders and both backward stages must give the same bits whatever
the pattern.  (Round-1 open issue: 8 ranks in 8 processes disagreed with one GPU in a handful of pixels while 8 slabs in
ONE process composed bit-exactly -- the signature of a stale-memory read hidden by the caching allocator.)

    python tools/diag_poison.py [K ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rade-gs_b200"))
import torch  # noqa: E402

import diff_gaussian_rasterization as dgr  # noqa: E402
from rade_gs_b200 import multigpu, scenes  # noqa: E402

C = dgr._C
dev = torch.device("cuda")
E = torch.Tensor([])
IMG = {1: "color", 2: "coord", 3: "mcoord", 4: "alpha", 5: "normal", 6: "depth", 7: "mdepth"}
bad_total = 0


def run(sc, g, coord, depth, ks, b, e):
    args = (sc.bg, sc.means3D, E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx, sc.tanfovy, ks, sc.height,
            sc.width, sc.shs, 3, sc.campos, False, coord, depth, False)
    out = C.rasterize_gaussians_slab(*args, b, e)
    acc = C.rasterize_gaussians_backward_render(sc.bg, sc.means3D, out[8], E, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix,
                                                sc.tanfovx, sc.tanfovy, ks, g["color"], g["coord"], g["mcoord"], g["depth"], g["mdepth"], g["alpha"],
                                                g["normal"], out[5], sc.shs, 3, sc.campos, out[9], out[0], out[10], out[11], out[4], coord, depth, False,
                                                b, e)
    return out, acc


def stage2(sc, acc, out, coord, depth, ks):
    return C.rasterize_gaussians_backward_preprocess(acc, sc.bg, sc.means3D, out[8], E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix,
                                                     sc.projmatrix, sc.tanfovx, sc.tanfovy, ks, sc.height, sc.width, sc.shs, 3, sc.campos, out[9], coord,
                                                     depth, False)


def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    global bad_total
    Ks = [int(k) for k in sys.argv[1:]] or [2, 8]
    sc = scenes.make_scene(60000, 640, 400, 500.0, -3.8, seed=21, view=scenes.look_at_view((0.3, 0.2, -0.4), (0.0, 0.1, 6.0)), bg=(0.2, 0.1, 0.3)).to(dev)
    g = scenes.make_upstream_grads(sc.height, sc.width, seed=5, device=dev)
    gy = (sc.height + 15) // 16
    for coord, depth in ((True, True), (False, True), (False, False)):
        ks = 0.1
        C.set_poison(-1)
        ref_out, ref_acc = run(sc, g, coord, depth, ks, 0, gy)
        ref_g = stage2(sc, ref_acc, ref_out, coord, depth, ks)
        for pat in (0xFF, 0x00, 0x7F, 0x3B):
            C.set_poison(pat)
            out, acc = run(sc, g, coord, depth, ks, 0, gy)
            bad = {n: int((out[i] != ref_out[i]).sum()) for i, n in IMG.items()}
            bad["radii"] = int((out[8] != ref_out[8]).sum())
            bad["R"] = int(out[0] != ref_out[0])
            grads = stage2(sc, ref_acc, out, coord, depth, ks)     # same accumulator -> stage 2 is deterministic
            gbad = [int((a != b).sum()) + int(torch.isnan(a).sum()) for a, b in zip(grads, ref_g)]
            nb = sum(bad.values()) + sum(gbad)
            bad_total += nb
            print(f"coord={coord} depth={depth} poison=0x{pat:02X} whole: image mismatches {bad}  acc rel {rel(acc, ref_acc):.2e} nan {int(torch.isnan(acc).sum())}"
                  f"  stage-2 mismatches {gbad}", flush=True)
            for K in Ks:
                slabs = multigpu.partition_tile_rows(gy, K)
                acc_sum = torch.zeros_like(ref_acc)
                sb = {n: 0 for n in IMG.values()}
                outside = 0
                Rs = 0
                for (b, e) in slabs:
                    o, a = run(sc, g, coord, depth, ks, b, e)
                    Rs += o[0]
                    acc_sum += a
                    for i, n in IMG.items():
                        sb[n] += int((o[i][:, b * 16: e * 16] != ref_out[i][:, b * 16: e * 16]).sum())
                        z = o[i].clone()
                        z[:, b * 16: e * 16] = 0
                        outside += int((z != 0).sum())
                    sb["radii"] = sb.get("radii", 0) + int((o[8] != ref_out[8]).sum())
                nb = sum(sb.values()) + outside + int(Rs != ref_out[0]) + int(torch.isnan(acc_sum).sum())
                bad_total += nb
                print(f"    K={K}: slab mismatches {sb} outside {outside} sumR {Rs} vs {ref_out[0]}  acc-sum rel {rel(acc_sum, ref_acc):.2e} "
                      f"nan {int(torch.isnan(acc_sum).sum())}", flush=True)
    C.set_poison(-1)
    print("TOTAL MISMATCHES", bad_total)
    sys.exit(1 if bad_total else 0)


if __name__ == "__main__":
    main()
