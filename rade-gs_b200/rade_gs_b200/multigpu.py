"""Tile-row sharding of one camera across the GPUs of a box (the reference is single-GPU only).

Every rank holds the full Gaussian set and runs preprocess on all of it (identical on every rank); rank r bins,
sorts and blends only the tile rows of its slab, so the union of the per-rank sorted lists is the reference's
list.  Backward has exactly one exchange step: the packed screen-space gradient rows written by
backward-render are plain sums over pixels, hence additive across slabs -- they are summed with ONE all-reduce
(NCCL over NVLink on GPUs, gloo in the CPU tests) and only then turned into parameter gradients.  The reduce
has to sit before backward-preprocess, not after it: with kernel_size > 0 that stage multiplies two
accumulated quantities (the reference's mip-gradient aliasing), so per-slab results would not
add up to the single-GPU answer.

`backward_two_stage` is written against three callables so that the same control flow is exercised on CPU
(gloo + the oracle) and on GPUs (NCCL + the CUDA stages).
"""
from __future__ import annotations

import os
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 16


def tile_rows(height: int) -> int:
    return (height + TILE - 1) // TILE


def partition_tile_rows(grid_y: int, world_size: int, weights: Sequence[float] | None = None) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) tile-row slabs, one per rank.

    Without weights rows are split as evenly as possible (earlier ranks take the remainder, e.g. 68 rows over 8
    ranks -> 9,9,9,9,8,8,8,8).  With per-row weights (e.g. instances per tile row from a previous frame) the cut
    points equalise the weight prefix sums; every rank still gets a (possibly empty) contiguous range and the
    ranges tile [0, grid_y) exactly.
    """
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if weights is None:
        base, rem = divmod(grid_y, world_size)
        out, b = [], 0
        for r in range(world_size):
            e = b + base + (1 if r < rem else 0)
            out.append((b, e))
            b = e
        return out
    if len(weights) != grid_y:
        raise ValueError("need one weight per tile row")
    total = float(sum(weights))
    if total <= 0:
        return partition_tile_rows(grid_y, world_size)
    cuts, acc, r = [0], 0.0, 1
    for y, w in enumerate(weights):
        acc += float(w)
        while r < world_size and acc >= total * r / world_size:
            cuts.append(y + 1)
            r += 1
    while len(cuts) < world_size:
        cuts.append(grid_y)
    cuts.append(grid_y)
    return [(cuts[i], max(cuts[i], cuts[i + 1])) for i in range(world_size)]


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks; a no-op outside an initialised process group (single-GPU path)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def exchange_sum_(acc: torch.Tensor, group=None, mode: str | None = None) -> torch.Tensor:
    """Sum of the slab-local accumulator rows [P, stride] over the ranks, in place, bit-identical on every rank.

    ``dense`` (default, or ``RGS_GRAD_EXCHANGE``): one all-reduce of the whole tensor.
    ``sparse`` (opt-in): a rank's rows are zero for every Gaussian that does not reach its slab (with 8 slabs a splat touches
    1.25 of them on average), so each rank ships only its non-zero rows -- index + row, one all-gather -- and every rank
    rebuilds the sum by adding the ranks' rows in rank order (fixed order: same bits everywhere, like an all-reduce).
    Measured on 8 B200 at C2 it LOSES to NCCL's all-reduce (1.67 vs 1.43 ms per step, profiles/r01_bench_ours_n8*.json): the
    two size exchanges with their host synchronisations and eight index_add_ passes cost more than the 0.5 ms all-reduce
    they replace.  Kept because the control flow is the starting point for a fused device-side exchange (no host syncs).
    """
    if not (dist.is_available() and dist.is_initialized()):
        return acc
    W = dist.get_world_size(group)
    if W == 1:
        return acc
    mode = mode or os.environ.get("RGS_GRAD_EXCHANGE", "dense")
    if mode == "owner":
        # Host-side mirror of the device exchange (csrc/rgs_exchange.cu), collective by collective: rows are owned in contiguous
        # blocks of ceil(P / W); every block is reduced INTO its owner and the owner's sum is then copied to everybody -- one sum per
        # row, the same bits on every rank.  (The device version moves only the rows that carry something; this one is dense and
        # exists so that the ownership / reduce / spread logic runs under gloo in the CPU tests.)
        P = acc.shape[0]
        rpr = (P + W - 1) // W
        for owner in range(W):
            blk = acc[owner * rpr: min(P, (owner + 1) * rpr)]
            if blk.numel():
                dist.reduce(blk, dst=dist.get_global_rank(group, owner) if group is not None else owner, op=dist.ReduceOp.SUM, group=group)
        for owner in range(W):
            blk = acc[owner * rpr: min(P, (owner + 1) * rpr)]
            if blk.numel():
                dist.broadcast(blk, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        return acc
    if mode != "sparse":
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        return acc
    idx = torch.nonzero((acc != 0).any(dim=1)).squeeze(1)
    n_mine = torch.tensor([idx.numel()], dtype=torch.int64, device=acc.device)
    n_all = [torch.zeros_like(n_mine) for _ in range(W)]
    dist.all_gather(n_all, n_mine, group=group)
    counts = [int(c) for c in torch.cat(n_all).tolist()]  # same list on every rank: the branch below is taken by all or none
    maxn, stride = max(counts), acc.shape[1]
    rows = torch.zeros(maxn, stride, dtype=acc.dtype, device=acc.device)
    ids = torch.zeros(maxn, dtype=torch.int32, device=acc.device)
    rows[: idx.numel()] = acc[idx]
    ids[: idx.numel()] = idx.to(torch.int32)
    rows_all = torch.empty(W * maxn, stride, dtype=acc.dtype, device=acc.device)
    ids_all = torch.empty(W * maxn, dtype=torch.int32, device=acc.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(rows_all, rows, group=group)
        dist.all_gather_into_tensor(ids_all, ids, group=group)
    else:  # gloo (CPU tests) has no flat all-gather
        dist.all_gather(list(rows_all.view(W, maxn, stride).unbind(0)), rows, group=group)
        dist.all_gather(list(ids_all.view(W, maxn).unbind(0)), ids, group=group)
    acc.zero_()
    for r in range(W):  # rank order, own rows included: indices are unique within a rank, so each add is deterministic
        k = counts[r]
        if k:
            acc.index_add_(0, ids_all[r * maxn: r * maxn + k].long(), rows_all[r * maxn: r * maxn + k])
    return acc


def broadcast_scene_(scene, src: int = 0, group=None):
    """Make a replicated scene (any object whose tensor attributes are the model / camera state) bit-identical on every rank by
    broadcasting rank `src`'s copy -- what a trainer does with its replicated parameters.  Data-parallel slabs only compose to the
    single-GPU image if every rank rasterizes THE SAME numbers: two processes that each synthesise "the same" scene on their CPUs can
    differ in the last bits (threaded vectorised exp / sigmoid / matmul tails), and a few-ulp difference in one splat is a visible
    few-pixel difference in the gathered image."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return scene
    for k, v in list(vars(scene).items()):
        if isinstance(v, torch.Tensor):
            t = v.contiguous()
            dist.broadcast(t, src=src, group=group)
            setattr(scene, k, t)
    return scene


class GradExchange:
    """The one exchange step of the sharded backward: slab-local accumulator rows -> summed rows on every rank.

    ``peer`` (default on CUDA with NCCL): the library's device-side exchange over peer memory (csrc/rgs_exchange.cu: touched
    rows are reduced into their owner rank with vector reductions through NVLink, the sums are spread to every rank, two
    flag barriers, no host synchronisation; bit-identical rows on all ranks).  torch.distributed is used ONCE, to all-gather
    the 64-byte IPC handles.  ``dense`` (``RGS_GRAD_EXCHANGE=dense``, and the only mode off-GPU): stage 1 into a fresh tensor
    + one NCCL / gloo all-reduce of the whole tensor.

    All ranks must create the object collectively and call ``backward_render`` the same number of times with the same P.
    """

    def __init__(self, capacity_rows: int, row_floats: int, device, group=None, mode: str | None = None):
        self.group, self.device = group, torch.device(device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        mode = mode or os.environ.get("RGS_GRAD_EXCHANGE", "peer")
        if self.world == 1 or self.device.type != "cuda":
            mode = "dense"
        self.mode, self.capacity, self.row_floats, self._ex = mode, int(capacity_rows), int(row_floats), None
        self.window = None
        if mode == "peer":
            from diff_gaussian_rasterization import _C
            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            if os.environ.get("RGS_EXCHANGE_WINDOW", "ipc") == "symm":
                # opt-in: windows from torch's symmetric memory (cuMem VMM) with an NVLS multicast mapping -> the spread phase
                # issues one store per 16 bytes and the NVSwitch replicates it; any failure falls back to the CUDA-IPC windows
                try:
                    self._attach_symmetric(_C, idx)
                except Exception as e:  # noqa: BLE001
                    import warnings
                    warnings.warn(f"symmetric-memory windows unavailable ({type(e).__name__}: {e}); using CUDA IPC windows")
                    self._ex = None
            if self._ex is None:
                self._ex, handle = _C.exchange_create(self.rank, self.world, self.capacity, self.row_floats, idx)
                mine = handle.to(self.device)
                gathered = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(gathered, mine, group=group)
                _C.exchange_connect(self._ex, torch.stack(gathered).cpu(), idx)
                self.window = "cuda-ipc"
            dist.barrier(group=group)   # nobody pushes before every rank has mapped every window
            why = self._self_check(_C, idx)
            if why is not None:         # the same verdict on every rank (all-reduced): fall back together
                import warnings
                warnings.warn(f"peer gradient exchange failed its self-check ({why}); using the dense all-reduce")
                self.fallback_reason = why
                self.mode = "dense"

    fallback_reason = None

    def _attach_symmetric(self, C, idx):
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = int(C.exchange_window_bytes(self.world, self.capacity, self.row_floats))
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", idx))
        g = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(buf, g.group_name)
        buf.zero_()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        mc = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
        self._ex = C.exchange_attach(self.rank, self.world, self.capacity, self.row_floats, [int(p) for p in hdl.buffer_ptrs], mc, idx)
        self._symm = (buf, hdl)   # keep the allocation alive
        self.window = "symmetric-memory" + (" + NVLS multicast" if mc else " (no multicast)")

    def _self_check(self, C, idx):
        """Two synthetic exchange steps with a known answer before the object is trusted: rank r touches every 3rd row starting at
        (r + step) % 3 (+ row 0, touched by all) and fills it with f(step, rank, row); the result must be the sum over the ranks that
        touched the row, bit for bit (small integers: exact in float), on every rank.  Returns None or the reason of the failure."""
        P = min(self.capacity, 4096 + 17)
        dev = torch.device("cuda", idx)
        rows = torch.arange(P, device=dev)
        radii = torch.ones(P, dtype=torch.int32, device=dev)
        radii[5::11] = 0                                         # rows nobody renders are neither pushed nor spread
        ok = True
        try:
            for step in (1, 2):
                expect = torch.zeros(P, self.row_floats, device=dev)
                mine = None
                for r in range(self.world):
                    touched = ((rows % 3) == ((r + step) % 3)) | (rows == 0)   # step 2 leaves some of step 1's rows untouched: they must read zero again
                    touched &= radii > 0
                    val = (touched.float() * (step * 64 + r + 1))[:, None] * (1 + (rows % 7).float())[:, None] * torch.ones(1, self.row_floats, device=dev)
                    expect += val
                    if r == self.rank:
                        mine = (touched, val)
                acc = C.exchange_accumulator(self._ex, P, self.row_floats, idx)
                acc.copy_(mine[1])
                C.exchange_rows(self._ex, mine[0].int(), radii)
                if C.exchange_status(self._ex) != 0:
                    ok = False
                    break
                got = C.exchange_result(self._ex, P, self.row_floats, idx)
                vis = radii > 0
                if not torch.equal(got[vis], expect[vis]) or bool(acc.any()):
                    ok = False
                    break
        except Exception as e:   # noqa: BLE001  (a failing CUDA call: report, fall back)
            ok = False
            why_local = f"{type(e).__name__}: {e}"
        flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            return None
        return "a rank saw a wrong result or a barrier timeout" if ok else locals().get("why_local", "wrong result or barrier timeout on this rank")

    def backward_render(self, C, *stage1_args) -> torch.Tensor:
        """stage 1 (`rasterize_gaussians_backward_render` arguments) + exchange -> summed accumulator rows [P, row_floats]."""
        if self.mode == "peer":
            return C.rasterize_gaussians_backward_render_exchange(self._ex, *stage1_args)
        acc = C.rasterize_gaussians_backward_render(*stage1_args)
        return exchange_sum_(acc, self.group, mode="dense")

    def close(self):
        if self._ex is not None:
            from diff_gaussian_rasterization import _C
            if dist.is_initialized():
                dist.barrier(group=self.group)   # peers may still be reading / writing this rank's window
            _C.exchange_destroy(self._ex)
            self._ex = None


def backward_two_stage(stage1_render: Callable[[], torch.Tensor], stage2_preprocess: Callable[[torch.Tensor], tuple], group=None):
    """slab-local scatter -> one exchange of the accumulator rows -> replicated parameter gradients."""
    acc = stage1_render()
    exchange_sum_(acc, group)
    return stage2_preprocess(acc)


class _ShardedRasterize(torch.autograd.Function):
    """Autograd node of the row-sharded rasterizer; same argument / gradient order as the single-GPU
    `_RasterizeGaussians` (reference: diff_gaussian_rasterization/__init__.py:44-169)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, slab, group, exchange=None,
                compact=False):
        from diff_gaussian_rasterization import _C
        s = raster_settings
        out = _C.rasterize_gaussians_slab(
            s.bg, means3D, colors_precomp, opacities, scales, rotations, s.scale_modifier, cov3Ds_precomp, s.viewmatrix, s.projmatrix,
            s.tanfovx, s.tanfovy, s.kernel_size, s.image_height, s.image_width, sh, s.sh_degree, s.campos, s.prefiltered,
            s.require_coord, s.require_depth, s.debug, slab[0], slab[1], compact)
        ctx.compact = compact
        num_rendered, color, coord, mcoord, alpha, normal, depth, mdepth, radii, geom, binning, img = out
        ctx.s, ctx.slab, ctx.group, ctx.num_rendered, ctx.exchange = s, slab, group, num_rendered, exchange
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, normal, radii, sh, geom, binning, img, alpha, opacities)
        return color, radii, coord, mcoord, depth, mdepth, alpha, normal

    @staticmethod
    def backward(ctx, g_color, g_radii, g_coord, g_mcoord, g_depth, g_mdepth, g_alpha, g_normal):
        from diff_gaussian_rasterization import _C
        s = ctx.s
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, normal, radii, sh, geom, binning, img, alpha, opacities = ctx.saved_tensors

        stage1_args = (s.bg, means3D, radii, colors_precomp, scales, rotations, s.scale_modifier, cov3Ds_precomp, s.viewmatrix, s.projmatrix,
                       s.tanfovx, s.tanfovy, s.kernel_size, g_color, g_coord, g_mcoord, g_depth, g_mdepth, g_alpha, g_normal, normal, sh,
                       s.sh_degree, s.campos, geom, ctx.num_rendered, binning, img, alpha, s.require_coord, s.require_depth, s.debug,
                       ctx.slab[0], ctx.slab[1], ctx.compact, s.image_height)

        def stage1():
            return _C.rasterize_gaussians_backward_render(*stage1_args)

        def stage2(acc):
            return _C.rasterize_gaussians_backward_preprocess(
                acc, s.bg, means3D, radii, colors_precomp, opacities, scales, rotations, s.scale_modifier, cov3Ds_precomp, s.viewmatrix,
                s.projmatrix, s.tanfovx, s.tanfovy, s.kernel_size, s.image_height, s.image_width, sh, s.sh_degree, s.campos, geom,
                s.require_coord, s.require_depth, s.debug)

        if ctx.exchange is not None:   # device-side exchange (or its dense fallback) owned by the caller
            g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot = stage2(ctx.exchange.backward_render(_C, *stage1_args))
        else:
            g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot = backward_two_stage(stage1, stage2, ctx.group)
        return g_means3D, g_means2D, g_sh, g_colors, g_opac, g_scales, g_rot, g_cov3D, None, None, None, None, None


class _GatherSlabs(torch.autograd.Function):
    """All-gather of the ranks' slab rows into the whole map (pure copies: the gathered map holds exactly the bits each
    rank rendered); identity backward (every rank holds the same full-image loss, and a rank's backward reads only the rows
    of its own slab)."""

    @staticmethod
    def forward(ctx, img, group, pixel_rows, rank):
        world = len(pixel_rows)
        if not (dist.is_available() and dist.is_initialized()) or world == 1:
            return img.detach().clone()
        C, H, W = img.shape
        hmax = max(e - b for b, e in pixel_rows)
        b, e = pixel_rows[rank]
        mine = img.new_zeros(C, hmax, W)
        mine[:, : e - b] = img.detach()[:, b:e]
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        out = torch.empty_like(img)
        for r, (rb, re) in enumerate(pixel_rows):
            out[:, rb:re] = parts[r][:, : re - rb]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return grad_out, None, None, None


class ShardedGaussianRasterizer(torch.nn.Module):
    """`GaussianRasterizer` whose work is split by tile rows over the ranks of a process group.

    Each rank returns full-size maps of which only its slab rows are filled (the rest is zero) or, with ``compact=True``, maps
    that hold only the slab's pixel rows ([C, rows, W]); callers that need the whole image gather the slabs (`gather_image`).  Gradients returned on every rank are the full,
    already-reduced parameter gradients.
    """

    def __init__(self, raster_settings, rank: int | None = None, world_size: int | None = None, group=None, row_weights=None,
                 exchange: GradExchange | None = None, compact: bool = False):
        super().__init__()
        self.compact = compact     # True: the returned maps are [C, slab rows, W] (per-rank image memory and fill scale with 1/ranks)
        self.raster_settings = raster_settings
        self.group = group
        self.exchange = exchange   # a GradExchange created once by the caller (collective); None: one dense all-reduce per backward
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world_size = rank, world_size
        self.slabs = partition_tile_rows(tile_rows(raster_settings.image_height), world_size, row_weights)
        self.slab = self.slabs[rank]

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        from diff_gaussian_rasterization import _absent, _check_exclusive
        _check_exclusive(shs, colors_precomp, scales, rotations, cov3D_precomp)
        shs = _absent() if shs is None else shs
        colors_precomp = _absent() if colors_precomp is None else colors_precomp
        scales = _absent() if scales is None else scales
        rotations = _absent() if rotations is None else rotations
        cov3D_precomp = _absent() if cov3D_precomp is None else cov3D_precomp
        return _ShardedRasterize.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                       self.raster_settings, self.slab, self.group, self.exchange, self.compact)

    def pixel_rows(self) -> Tuple[int, int]:
        H = self.raster_settings.image_height
        return min(self.slab[0] * TILE, H), min(self.slab[1] * TILE, H)

    def gather_image(self, img: torch.Tensor) -> torch.Tensor:
        """The whole [C,H,W] map from the ranks' slabs (all-gather of the slab rows; rows outside a rank's slab are zero locally).

        Differentiable: with the whole image on every rank an unchanged full-image loss (train.py:130-165: L1, SSIM,
        normal consistency) can be evaluated redundantly per rank; its gradient flows back unchanged and the sharded
        rasterizer's backward reads only the rows of its own slab, so nothing is counted twice.
        """
        H = self.raster_settings.image_height
        rows = [(min(b * TILE, H), min(e * TILE, H)) for b, e in self.slabs]
        if self.compact:   # [C, slab rows, W] -> place into a full-size map first (the gather copies slab rows only)
            full = img.new_zeros(img.shape[0], H, img.shape[2])
            r0, r1 = rows[self.rank]
            full[:, r0:r1] = img
            img = full
        return _GatherSlabs.apply(img, self.group, rows, self.rank)


# ---- slab-local image loss with halo rows (multi-GPU half) ----------------------------------------------------
HALO = 5   # 11x11 SSIM window (utils/loss_utils.py:35: window_size=11, padding 5)


def _exchange_rows(t: torch.Tensor, send_up, send_down, recv_up, recv_down, rank: int, world: int, group, add: bool):
    """Neighbour exchange of row blocks of a [C,H,W] tensor: rows `send_up` go to rank-1, `send_down` to rank+1; what the
    neighbours send lands in rows `recv_up` (from rank-1) / `recv_down` (from rank+1), overwriting or (add=True) accumulating."""
    ops, bufs = [], []
    for peer, srows, rrows in ((rank - 1, send_up, recv_up), (rank + 1, send_down, recv_down)):
        if peer < 0 or peer >= world:
            continue
        if srows[1] > srows[0]:
            ops.append(dist.P2POp(dist.isend, t[:, srows[0]:srows[1]].contiguous(), peer, group=group))
        if rrows[1] > rrows[0]:
            buf = torch.empty_like(t[:, rrows[0]:rrows[1]]).contiguous()
            bufs.append((rrows, buf))
            ops.append(dist.P2POp(dist.irecv, buf, peer, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for (r0, r1), buf in bufs:
        if add:
            t[:, r0:r1] += buf
        else:
            t[:, r0:r1] = buf


class _SlabSsimL1(torch.autograd.Function):
    """(1 - l) * L1 + l * (1 - SSIM) of the WHOLE image (train.py:163), evaluated slab by slab: a rank counts the SSIM-map and L1
    rows of its own slab; the 11x11 window reaches 5 rows into the neighbours' slabs, so 5 halo rows travel to each neighbour in
    forward and the gradient a rank holds for its neighbours' halo pixels travels back (and is added) in backward.  Two scalars are
    all-reduced; the image itself is never gathered.  `fwd` / `bwd` are the per-slab kernels (CUDA kernels in the product,
    torch expressions in the CPU test): fwd(img, gt, r0, r1, need_grad) -> (sums[2] float64, state), bwd(img, gt, state, w_ssim,
    w_l1, r0, r1) -> d_img [C,H,W] (non-zero in rows [r0-5, r1+5) only)."""

    @staticmethod
    def forward(ctx, image, gt, lambda_dssim, rows, rank, world, group, fwd, bwd):
        r0, r1 = rows
        H = image.shape[-2]
        img = image.detach().clone()
        # my first / last 5 rows are the neighbours' halos; theirs are mine
        _exchange_rows(img, (r0, min(r0 + HALO, r1)), (max(r1 - HALO, r0), r1), (max(r0 - HALO, 0), r0), (r1, min(r1 + HALO, H)), rank, world, group, add=False)
        sums, state = fwd(img, gt, r0, r1, ctx.needs_input_grad[0])
        sums = sums.clone()
        if world > 1:
            dist.all_reduce(sums, group=group)
        n = image.numel()
        ctx.save_for_backward(img, gt, *state)
        ctx.meta = (float(lambda_dssim), n, r0, r1, rank, world, group, bwd, H)
        return ((1.0 - lambda_dssim) * sums[1] / n + lambda_dssim * (1.0 - sums[0] / n)).float()

    @staticmethod
    def backward(ctx, grad_out):
        img, gt, *state = ctx.saved_tensors
        lam, n, r0, r1, rank, world, group, bwd, H = ctx.meta
        d = bwd(img, gt, state, -lam / n, (1.0 - lam) / n, r0, r1)
        # rows [r0-5, r0) and [r1, r1+5) hold what MY slab's SSIM values contribute to the neighbours' pixels: send them home
        _exchange_rows(d, (max(r0 - HALO, 0), r0), (r1, min(r1 + HALO, H)), (r0, min(r0 + HALO, r1)), (max(r1 - HALO, r0), r1), rank, world, group, add=True)
        out = torch.zeros_like(d)
        out[:, r0:r1] = d[:, r0:r1]
        return out * grad_out, None, None, None, None, None, None, None, None


def _cuda_slab_fwd(img, gt, r0, r1, need_grad):
    from diff_gaussian_rasterization import _C
    sums, dmaps = _C.ssim_l1_forward(img, gt, need_grad, r0, r1)
    return sums, (dmaps,)


def _cuda_slab_bwd(img, gt, state, w_ssim, w_l1, r0, r1):
    from diff_gaussian_rasterization import _C
    return _C.ssim_l1_backward(img, gt, state[0], w_ssim, w_l1, torch.empty(0), r0, r1)


def slab_l1_ssim_loss(image: torch.Tensor, gt: torch.Tensor, lambda_dssim: float, pixel_rows: Tuple[int, int], rank: int | None = None,
                      world_size: int | None = None, group=None, kernels=None) -> torch.Tensor:
    """train.py:163's `(1 - l) * l1_loss(image, gt) + l * (1 - ssim(image, gt))` for a row-sharded render: `image` [3,H,W] holds this
    rank's slab rows `pixel_rows` (`ShardedGaussianRasterizer.pixel_rows()`), `gt` the ground truth (at least those rows +- 5).
    Returns the full-image loss (the same value on every rank); its gradient w.r.t. `image` is non-zero in this rank's rows only and
    already contains the neighbours' contributions.  Slabs must be at least 5 rows tall (they are multiples of 16)."""
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size > 1 and pixel_rows[1] - pixel_rows[0] < HALO:
        raise ValueError("slab-local SSIM needs slabs of at least 5 pixel rows on every rank (a shorter slab breaks the neighbour halo exchange)")
    fwd, bwd = kernels if kernels is not None else (_cuda_slab_fwd, _cuda_slab_bwd)
    return _SlabSsimL1.apply(image, gt, float(lambda_dssim), (int(pixel_rows[0]), int(pixel_rows[1])), rank, world_size, group, fwd, bwd)
