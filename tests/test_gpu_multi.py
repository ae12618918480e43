"""Multi-GPU tests of the tile-row sharded rasterizer: world sizes 2, 4 and 8 (as many as the box has GPUs; skipped on
single-GPU boxes).  One process per GPU; the images must be bit-identical to the single-GPU render, the gradients within the
float tolerance, and -- with the device-side peer exchange -- bit-identical on every rank.  The same control flow runs on
CPU with gloo in tests/test_multigpu_host.py."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, outdir, exchange_mode, variant, compact=False):
    for p in (os.path.join(ROOT, "rade-gs_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import scenes
    from rade_gs_b200.multigpu import GradExchange, ShardedGaussianRasterizer, broadcast_scene_
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ex = None
    try:
        coord, depth, ks = variant
        sc = scenes.make_scene(60000, 640, 400, 500.0, -3.8, seed=21, view=scenes.look_at_view((0.3, 0.2, -0.4), (0.0, 0.1, 6.0)), bg=(0.2, 0.1, 0.3)).to(dev)
        sc = broadcast_scene_(sc)   # replicated state = rank 0's bits on every rank, as in a trainer
        st = dgr.GaussianRasterizationSettings(sc.height, sc.width, sc.tanfovx, sc.tanfovy, ks, sc.bg, 1.0, sc.viewmatrix, sc.projmatrix, 3, sc.campos,
                                               False, depth, coord, False)
        P = sc.means3D.shape[0]
        ex = GradExchange(P, dgr._C.grad_stride(coord, depth), dev, mode=exchange_mode)
        assert ex.mode == exchange_mode

        def run(rast, seed, n=None):
            g = scenes.make_upstream_grads(sc.height, sc.width, seed=seed, device=dev)
            lv = {k: getattr(sc, k)[:n].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
            m2 = torch.zeros_like(lv["means3D"], requires_grad=True)
            color, radii, co, mco, dep, mdep, alpha, normal = rast(lv["means3D"], m2, lv["opacities"], shs=lv["shs"], scales=lv["scales"],
                                                                    rotations=lv["rotations"])
            if color.shape[1] != sc.height:     # compact slab maps: the upstream gradients of this rank are the rows of its slab
                r0, r1 = rast.pixel_rows()
                g = {k: v[:, r0:r1] for k, v in g.items()}
            loss = (color * g["color"]).sum() + (dep * g["depth"]).sum() + (normal * g["normal"]).sum() + (alpha * g["alpha"]).sum() + \
                (co * g["coord"]).sum() + (mdep * g["mdepth"]).sum() + (mco * g["mcoord"]).sum()
            loss.backward()
            return dict(color=color.detach(), depth=dep.detach(), normal=normal.detach(), coord=co.detach(), alpha=alpha.detach(), radii=radii,
                        **{"g_" + k: v.grad for k, v in lv.items()}, g_means2D=m2.grad)

        sharded = ShardedGaussianRasterizer(st, exchange=ex, compact=compact)
        single = dgr.GaussianRasterizer(st)
        res = {}
        # three steps on the same exchange object: different upstream gradients, then fewer Gaussians (rows re-associated), so the
        # self-cleaning accumulators and the changed-rows bookkeeping of the exchange are exercised, not just its first call
        for step, (seed, n) in enumerate(((5, None), (6, None), (7, P - 1234))):
            a = run(sharded, seed, n)
            for k in ("color", "depth", "normal", "coord", "alpha"):
                a[k] = sharded.gather_image(a[k])
            b = run(single, seed, n)                     # the single-GPU answer, computed on every rank's own GPU
            for k in a:
                x, y = a[k].float(), b[k].float()
                res[f"{step}/{k}"] = (float((x - y).abs().max()), float(y.abs().max()), float((x - y).norm() / (y.norm() + 1e-30)))
                if k in ("color", "depth", "normal", "coord", "alpha") and not torch.equal(a[k], b[k]):
                    # evidence for an intermittent few-pixel difference: where, which rank's slab, and which of the two renders repeats
                    bad = torch.nonzero((a[k] != b[k]).any(0))[:4].tolist()
                    rows = [(min(s0 * 16, sc.height), min(s1 * 16, sc.height)) for s0, s1 in sharded.slabs]
                    with torch.no_grad():
                        again_single = single(sc.means3D[:n], torch.zeros_like(sc.means3D[:n]), sc.opacities[:n], shs=sc.shs[:n], scales=sc.scales[:n],
                                              rotations=sc.rotations[:n])
                    idx = {"color": 0, "coord": 2, "depth": 4, "alpha": 6, "normal": 7}[k]
                    res[f"{step}/{k}/evidence"] = {
                        "pixels(y,x)": bad, "slab_owner": [next(r for r, (r0, r1) in enumerate(rows) if r0 <= y < r1) for y, x in bad],
                        "gathered": [a[k][:, y, x].tolist() for y, x in bad], "single": [b[k][:, y, x].tolist() for y, x in bad],
                        "single_again": [again_single[idx][:, y, x].tolist() for y, x in bad]}
            # are the replicated gradients the same bits on every rank?
            flat = torch.cat([a[k].reshape(-1).float() for k in sorted(a) if k.startswith("g_")])
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            res[f"{step}/bits_differ_from_rank0"] = int((flat.view(torch.int32) != ref.view(torch.int32)).sum())
        np.save(os.path.join(outdir, f"res{rank}.npy"), res, allow_pickle=True)
        ev = {k: v for k, v in res.items() if k.endswith("/evidence")}
        if ev:   # keep the full evidence where the driver of the GPU box can collect it
            try:
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                with open(os.path.join(ROOT, "gpurun_out", "multi_mismatch_evidence.txt"), "a") as f:
                    f.write(f"world {world} rank {rank} exchange {exchange_mode} variant {variant} compact {compact}: {ev}\n")
            except OSError:
                pass
    finally:
        if ex is not None:
            ex.close()
        dist.destroy_process_group()


_WORLDS = [w for w in (2, 4, 8)]
_VARIANTS = {"both_ks01": (True, True, 0.1), "depth_ks0": (False, True, 0.0), "coord_ks0": (True, False, 0.0)}
_CASES = [(w, "peer", v, False) for w in _WORLDS for v in ("both_ks01",)] + [(2, "dense", "both_ks01", False), (2, "peer", "depth_ks0", True),
                                                                             (4, "peer", "coord_ks0", True), (8, "peer", "coord_ks0", True)]


@pytest.mark.parametrize("world,exchange,variant,compact", _CASES)
def test_sharded_equals_single(world, exchange, variant, compact, tmp_path):
    """`world` ranks against the single-GPU answer (computed on EVERY rank's GPU): images bit-exact, gradients within tolerance,
    and identical bits on all ranks."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, 29600 + os.getpid() % 1000 + world, str(tmp_path), exchange, _VARIANTS[variant], compact), nprocs=world, join=True)
    for rank in range(world):
        res = np.load(tmp_path / f"res{rank}.npy", allow_pickle=True).item()
        for key, val in res.items():
            step, k = key.split("/")[:2]
            if k == "bits_differ_from_rank0":
                assert val == 0, (rank, key, val)                 # replicated gradients: same bits everywhere
            elif key.endswith("/evidence"):
                continue
            elif k in ("color", "depth", "normal", "coord", "alpha", "radii"):
                assert val[0] == 0.0, (rank, key, val, res.get(key + "/evidence"))   # slabs reproduce the single-GPU image bit for bit
            else:
                mx, ref, rel = val
                assert rel < 1e-3 and mx <= 1e-2 * ref + 1e-6, (rank, key, val)


def _slab_loss_worker(rank, world, port, outdir):
    for p in (os.path.join(ROOT, "rade-gs_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from rade_gs_b200 import losses, multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        H, W = 400, 640
        g = torch.Generator().manual_seed(5)
        gt = torch.rand(3, H, W, generator=g).to(dev)
        whole = (gt + 0.2 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1)
        slabs = multigpu.partition_tile_rows(multigpu.tile_rows(H), world)
        r0, r1 = min(slabs[rank][0] * 16, H), min(slabs[rank][1] * 16, H)
        mine = torch.zeros_like(whole)
        mine[:, r0:r1] = whole[:, r0:r1]
        mine.requires_grad_(True)
        loss = multigpu.slab_l1_ssim_loss(mine, gt, 0.2, (r0, r1))
        loss.backward()
        x = whole.clone().requires_grad_(True)
        ref = losses.l1_ssim_loss(x, gt, 0.2)              # the single-GPU fused loss on the whole image, on this rank's GPU
        ref.backward()
        out = mine.grad.clone()
        out[:, r0:r1] = 0
        res = {"loss": float(loss), "ref": float(ref), "outside": int((out != 0).sum()),
               "grad_err": float((mine.grad[:, r0:r1] - x.grad[:, r0:r1]).abs().max()), "grad_scale": float(x.grad.abs().max())}
        np.save(os.path.join(outdir, f"slabloss{rank}.npy"), res, allow_pickle=True)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_local_ssim_loss_equals_the_whole_image_loss(world, tmp_path):
    """Row-sharded L1 + SSIM with halo rows exchanged between neighbouring ranks (no image gather): same loss, same gradient rows."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_slab_loss_worker, args=(world, 29700 + os.getpid() % 1000 + world, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        r = np.load(tmp_path / f"slabloss{rank}.npy", allow_pickle=True).item()
        assert abs(r["loss"] - r["ref"]) < 2e-6, r
        assert r["outside"] == 0 and r["grad_err"] <= 1e-5 * r["grad_scale"] + 1e-12, r
