"""GPU tests of the drop-in surface: the autograd-facing `GaussianRasterizer` (what gaussian_renderer.render() calls,
reference gaussian_renderer/__init__.py:53-78), `markVisible`, empty / ragged inputs, and the raw C ABI driven
through ctypes with plain device pointers."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, ROOT, load_golden
from tolerances import IMG_OUTLIER_FRAC_GPU, grad_close_gpu, image_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _settings(dgr, sc, coord, depth, ks=0.0, deg=3, debug=False):
    return dgr.GaussianRasterizationSettings(
        image_height=sc.height, image_width=sc.width, tanfovx=sc.tanfovx, tanfovy=sc.tanfovy, kernel_size=ks, bg=sc.bg, scale_modifier=1.0,
        viewmatrix=sc.viewmatrix, projmatrix=sc.projmatrix, sh_degree=deg, campos=sc.campos, prefiltered=False, require_depth=depth,
        require_coord=coord, debug=debug)


@pytest.mark.parametrize("case", [c for c in ("depth_ks0", "both_ks01") if c in GOLDEN_CASES])
def test_autograd_module_matches_golden(case):
    """The call sequence of render(): means2D dummy with retain_grad, rasterizer(...), loss.backward()."""
    import diff_gaussian_rasterization as dgr
    from test_gpu_parity import _scene_from_golden
    d = load_golden(case)
    sc, _ = _scene_from_golden(d)
    leaves = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True) + 0
    means2D.retain_grad()
    rast = dgr.GaussianRasterizer(_settings(dgr, sc, bool(d["meta_coord"]), bool(d["meta_depth"]), float(d["meta_ks"]), int(d["meta_deg"])))
    color, radii, coord, mcoord, depth, mdepth, alpha, normal = rast(
        means3D=leaves["means3D"], means2D=means2D, shs=leaves["shs"], colors_precomp=None, opacities=leaves["opacities"],
        scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None)
    assert color.shape == (3, sc.height, sc.width) and radii.dtype == torch.int32 and alpha.shape == (1, sc.height, sc.width)
    up = {k: torch.from_numpy(d["gin_" + k]).to(DEV) for k in ("color", "coord", "mcoord", "depth", "mdepth", "alpha", "normal")}
    loss = (color * up["color"]).sum() + (coord * up["coord"]).sum() + (mcoord * up["mcoord"]).sum() + (depth * up["depth"]).sum() + \
           (mdepth * up["mdepth"]).sum() + (alpha * up["alpha"]).sum() + (normal * up["normal"]).sum()
    loss.backward()
    image_close(color.detach().cpu().numpy(), d["out_color"], IMG_OUTLIER_FRAC_GPU, "color")
    assert np.array_equal(radii.cpu().numpy(), d["out_radii"])
    for name, t in (("means3D", leaves["means3D"]), ("means2D", means2D), ("sh", leaves["shs"]), ("opacity", leaves["opacities"]),
                    ("scales", leaves["scales"]), ("rotations", leaves["rotations"])):
        noise = float(d["grad_noise_" + name]) / (np.abs(d["grad_" + name]).max() + 1e-30)
        grad_close_gpu(t.grad.cpu().numpy(), d["grad_" + name], name, rel=1e-3 + 4 * noise, elem=1e-3 + 4 * noise)
    # densification statistics consumer (scene/gaussian_model.py:743-747) reads these two slices
    assert means2D.grad.shape == (sc.means3D.shape[0], 3) and bool((means2D.grad[:, 2] >= 0).all())


def test_mark_visible_and_empty_inputs():
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import scenes
    sc = scenes.make_scene(1000, 64, 48, 60.0, -3.0, seed=3).to(DEV)
    sc.means3D[:10, 2] = -1.0
    sc.means3D[10, 2] = 0.2
    rast = dgr.GaussianRasterizer(_settings(dgr, sc, False, True))
    vis = rast.markVisible(sc.means3D)
    assert vis.dtype == torch.bool and not bool(vis[:11].any()) and bool(vis[11:].all())
    # P == 0: all-zero maps, zero rendered, empty gradients (reference rasterize_points.cu:90,195)
    E = torch.Tensor([])
    z3, z1 = torch.zeros(0, 3, device=DEV), torch.zeros(0, 1, device=DEV)
    out = dgr._C.rasterize_gaussians(sc.bg, z3, E, z1, z3, torch.zeros(0, 4, device=DEV), 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx,
                                     sc.tanfovy, 0.0, sc.height, sc.width, torch.zeros(0, 16, 3, device=DEV), 3, sc.campos, False, True, True, False)
    assert out[0] == 0 and not bool(out[1].any()) and out[8].numel() == 0
    # everything behind the camera: background only, nothing rendered
    sc2 = scenes.make_scene(500, 70, 50, 60.0, -3.0, seed=4, bg=(0.2, 0.4, 0.6)).to(DEV)  # 70x50: ragged last tile column / row
    sc2.means3D[:, 2] = -sc2.means3D[:, 2]
    from rade_gs_b200 import rawapi
    f = rawapi.forward(dgr._C, sc2, True, True)
    assert f["num_rendered"] == 0 and not bool(f["radii"].any())
    assert torch.allclose(f["color"], sc2.bg.view(3, 1, 1).expand_as(f["color"]))
    assert not bool(f["alpha"].any()) and not bool(f["depth"].any()) and not bool(f["normal"].any())
    g = scenes.make_upstream_grads(sc2.height, sc2.width, device=DEV)
    b = rawapi.backward(dgr._C, sc2, f, g)
    assert all(not bool(v.any()) for v in b.values())


def test_prefiltered_is_rejected_and_debug_mode_runs():
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import rawapi, scenes
    sc = scenes.make_scene(2000, 64, 48, 60.0, -3.0, seed=5).to(DEV)
    E = torch.Tensor([])
    with pytest.raises(RuntimeError, match="prefiltered"):
        dgr._C.rasterize_gaussians(sc.bg, sc.means3D, E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx,
                                   sc.tanfovy, 0.0, sc.height, sc.width, sc.shs, 3, sc.campos, True, False, True, False)
    a = rawapi.forward(dgr._C, sc, False, True, debug=True)
    b = rawapi.forward(dgr._C, sc, False, True, debug=False)
    assert a["num_rendered"] == b["num_rendered"] and torch.equal(a["color"], b["color"])
    # lower active SH degree than stored coefficients, and scale_modifier
    c = rawapi.forward(dgr._C, sc, False, True, sh_degree=1, scale_modifier=0.5)
    assert c["num_rendered"] < b["num_rendered"] and torch.isfinite(c["color"]).all()


def test_c_abi_direct_call_with_raw_pointers():
    """include/rgs_b200.h driven from ctypes: device pointers, a resize callback, a stream -- no torch types cross."""
    from rade_gs_b200 import rawapi, scenes
    import diff_gaussian_rasterization as dgr

    lib = ctypes.CDLL(os.path.join(ROOT, "rade-gs_b200", "rade_gs_b200", "librgs_b200.so"))
    sc = scenes.make_scene(4000, 96, 80, 90.0, -2.8, seed=6).to(DEV)
    P, H, W = sc.means3D.shape[0], sc.height, sc.width

    class Cam(ctypes.Structure):
        _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("tan_fovx", ctypes.c_float), ("tan_fovy", ctypes.c_float),
                    ("kernel_size", ctypes.c_float), ("scale_modifier", ctypes.c_float), ("viewmatrix", ctypes.c_void_p),
                    ("projmatrix", ctypes.c_void_p), ("cam_pos", ctypes.c_void_p), ("background", ctypes.c_void_p), ("sh_degree", ctypes.c_int32),
                    ("sh_coeffs", ctypes.c_int32), ("require_coord", ctypes.c_int32), ("require_depth", ctypes.c_int32),
                    ("prefiltered", ctypes.c_int32), ("debug", ctypes.c_int32), ("tile_row_begin", ctypes.c_int32), ("tile_row_end", ctypes.c_int32),
                    ("compact_slab", ctypes.c_int32)]

    class Gs(ctypes.Structure):
        _fields_ = [("P", ctypes.c_int32), ("means3D", ctypes.c_void_p), ("opacities", ctypes.c_void_p), ("shs", ctypes.c_void_p),
                    ("colors_precomp", ctypes.c_void_p), ("scales", ctypes.c_void_p), ("rotations", ctypes.c_void_p), ("cov3D_precomp", ctypes.c_void_p),
                    ("shs_rest", ctypes.c_void_p)]

    class Out(ctypes.Structure):
        _fields_ = [(n, ctypes.c_void_p) for n in ("out_color", "out_coord", "out_mcoord", "out_alpha", "out_normal", "out_depth", "out_mdepth", "radii")]

    RESIZE = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

    class Bufs(ctypes.Structure):
        _fields_ = [("geom", RESIZE), ("geom_user", ctypes.c_void_p), ("binning", RESIZE), ("binning_user", ctypes.c_void_p),
                    ("image", RESIZE), ("image_user", ctypes.c_void_p)]

    held = {}

    def make_cb(name):
        def cb(user, nbytes):
            held[name] = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=DEV)
            return held[name].data_ptr()
        return RESIZE(cb)

    cbs = [make_cb(n) for n in ("geom", "binning", "image")]
    cam = Cam(W, H, sc.tanfovx, sc.tanfovy, 0.0, 1.0, sc.viewmatrix.data_ptr(), sc.projmatrix.data_ptr(), sc.campos.data_ptr(), sc.bg.data_ptr(),
              3, 16, 0, 1, 0, 0, 0, -1)
    gs = Gs(P, sc.means3D.data_ptr(), sc.opacities.data_ptr(), sc.shs.data_ptr(), None, sc.scales.data_ptr(), sc.rotations.data_ptr(), None)
    maps = {n: torch.empty(c, H, W, device=DEV) for n, c in (("color", 3), ("coord", 3), ("mcoord", 3), ("alpha", 1), ("normal", 3), ("depth", 1), ("mdepth", 1))}
    radii = torch.empty(P, dtype=torch.int32, device=DEV)
    out = Out(*[maps[n].data_ptr() for n in ("color", "coord", "mcoord", "alpha", "normal", "depth", "mdepth")], radii.data_ptr())
    bufs = Bufs(cbs[0], None, cbs[1], None, cbs[2], None)
    lib.rgs_forward.restype = ctypes.c_int64
    lib.rgs_last_error.restype = ctypes.c_char_p
    stream = torch.cuda.current_stream().cuda_stream
    R = lib.rgs_forward(ctypes.byref(cam), ctypes.byref(gs), ctypes.byref(out), ctypes.byref(bufs), ctypes.c_void_p(stream))
    assert R >= 0, lib.rgs_last_error()
    torch.cuda.synchronize()
    ref = rawapi.forward(dgr._C, sc, False, True)
    assert R == ref["num_rendered"] and torch.equal(radii, ref["radii"])
    for n in ("color", "alpha", "depth", "normal"):
        assert torch.equal(maps[n], ref[n]), n
    # mark_visible through the C ABI
    present = torch.zeros(P, dtype=torch.uint8, device=DEV)
    lib.rgs_mark_visible.restype = ctypes.c_int32
    rc = lib.rgs_mark_visible(ctypes.c_int32(P), ctypes.c_void_p(sc.means3D.data_ptr()), ctypes.c_void_p(sc.viewmatrix.data_ptr()),
                              ctypes.c_void_p(sc.projmatrix.data_ptr()), ctypes.c_void_p(present.data_ptr()), ctypes.c_void_p(stream))
    torch.cuda.synchronize()
    assert rc == 0 and int(present.sum()) == P
    # an invalid slab is refused with a message, not a crash
    cam.tile_row_begin, cam.tile_row_end = 3, 99
    assert lib.rgs_forward(ctypes.byref(cam), ctypes.byref(gs), ctypes.byref(out), ctypes.byref(bufs), ctypes.c_void_p(stream)) == -1
    assert b"slab" in lib.rgs_last_error()
    lib.rgs_launch_count.restype = ctypes.c_int64
    assert lib.rgs_launch_count() > 0


def test_sharded_rasterizer_single_rank_equals_plain():
    """world_size 1 path of the multi-GPU wrapper: same outputs and gradients as GaussianRasterizer."""
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import scenes
    from rade_gs_b200.multigpu import ShardedGaussianRasterizer
    sc = scenes.make_scene(5000, 128, 96, 110.0, -2.8, seed=8).to(DEV)
    st = _settings(dgr, sc, False, True, ks=0.1)
    res = []
    for cls in (dgr.GaussianRasterizer, lambda s: ShardedGaussianRasterizer(s, rank=0, world_size=1)):
        lv = {k: getattr(sc, k).clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
        m2 = torch.zeros_like(lv["means3D"], requires_grad=True)
        o = cls(st)(lv["means3D"], m2, lv["opacities"], shs=lv["shs"], scales=lv["scales"], rotations=lv["rotations"])
        (o[0].sum() + 0.1 * o[4].sum() + 0.1 * o[7].sum()).backward()
        res.append((o, lv, m2))
    (o0, l0, m0), (o1, l1, m1) = res
    for a, b in zip(o0, o1):
        assert torch.equal(a, b)
    for k in l0:
        grad_close_gpu(l1[k].grad.cpu().numpy(), l0[k].grad.cpu().numpy(), k)   # two runs of the same kernels: atomic-order noise only (measured up to 2e-4 here)


def test_compact_slab_maps_equal_the_rows_of_the_whole_image():
    """`compact` slab calls (multi-GPU: maps hold only the slab's pixel rows, [C, Hs, W]) against the whole-image call: forward
    maps bit-identical to the corresponding rows, the slab accumulators add up to the whole-image accumulator (image height not a
    multiple of 16, uneven slabs, all four variants)."""
    import diff_gaussian_rasterization as dgr
    from rade_gs_b200 import multigpu, scenes
    C = dgr._C
    sc = scenes.make_scene(30000, 200, 150, 260.0, -3.2, seed=8, view=scenes.look_at_view((0.2, 0.1, -0.3), (0.0, 0.0, 6.0)), bg=(0.3, 0.1, 0.2)).to(DEV)
    g = scenes.make_upstream_grads(sc.height, sc.width, seed=9, device=DEV)
    E = torch.Tensor([])
    gy = (sc.height + 15) // 16
    order = ("color", "coord", "mcoord", "depth", "mdepth", "alpha", "normal")

    def fwd(b, e, compact):
        return C.rasterize_gaussians_slab(sc.bg, sc.means3D, E, sc.opacities, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx,
                                          sc.tanfovy, 0.1, sc.height, sc.width, sc.shs, 3, sc.campos, False, coord, depth, False, b, e, compact)

    def bwd(out, grads, b, e, compact):
        return C.rasterize_gaussians_backward_render(sc.bg, sc.means3D, out[8], E, sc.scales, sc.rotations, 1.0, E, sc.viewmatrix, sc.projmatrix, sc.tanfovx,
                                                     sc.tanfovy, 0.1, *[grads[k] for k in order], out[5], sc.shs, 3, sc.campos, out[9], out[0], out[10],
                                                     out[11], out[4], coord, depth, False, b, e, compact, sc.height)

    for coord, depth in ((True, True), (False, True), (True, False), (False, False)):
        whole = fwd(0, gy, False)
        acc_whole = bwd(whole, g, 0, gy, False)
        acc_sum = torch.zeros_like(acc_whole)
        for b, e in multigpu.partition_tile_rows(gy, 3):
            r0, r1 = b * 16, min(e * 16, sc.height)
            out = fwd(b, e, True)
            for i in range(1, 8):
                assert out[i].shape[1] == r1 - r0 and torch.equal(out[i], whole[i][:, r0:r1]), (coord, depth, i, b, e)
            assert torch.equal(out[8], whole[8])
            acc_sum += bwd(out, {k: v[:, r0:r1].contiguous() for k, v in g.items()}, b, e, True)
        rel = float((acc_sum - acc_whole).norm() / (acc_whole.norm() + 1e-30))
        assert rel < 1e-5, (coord, depth, rel)
