"""Row-block sharding of the SSGI chain across N GPUs (one process per GPU, torch.distributed / NCCL over NVLink).

Every kernel of the path writes disjoint output rows, so the produced planes are partitioned by row block
(SURVEY.md §8e).  Two kinds of inputs cross row-block borders:

  * bounded stencils — K2's 5x5 neighbourhood of ssgiOut (2 rows), K3's Poisson taps (ceil(radius)+1 rows per pass,
    bilinear footprint included), K4's pixel-centre bilinear fetch of dnB (1 row).  Instead of one halo exchange per
    pass, each rank RECOMPUTES the halo rows itself: pass k is launched on a row range widened by the halos of all the
    passes after it (`ShardPlan`).  Kernels are bit-identical under row sharding, so the recomputed rows equal the
    owner's rows bit for bit, and no per-pass NCCL latency is paid (cost: ~1 % extra rows per block at 4K).
  * arbitrary-uv gathers — K1 samples `composed` at ray hit points and K2 samples the history `dnB[0..1]` at
    reprojected uvs anywhere on screen, so these three produced planes are all-gathered once per frame (32 B/px of each
    rank's rows).  The static inputs (depth, gBuffer, velocity, directLight) are given to every rank in full.

Load balance: sky rows are nearly free (background pixels are discarded) while floor rows are the most expensive, so
contiguous bands would leave the max-over-ranks time ~1/3 above the mean.  Rows are therefore assigned BLOCK-CYCLICALLY:
with B blocks per rank the frame is cut into N*B blocks and rank r owns blocks r, r+N, r+2N, ...; super-block j (blocks
j*N .. j*N+N-1) is contiguous and in rank order, so each plane is all-gathered in place with B collectives.

Overlap: a frame is issued in two phases.  K1 only needs last frame's `composed`; K2..K4 need `dnB`.  The all-gathers
are launched asynchronously after K4 (composed first); the next frame waits for the composed gather before K1 and for the
dnB gathers only before K2, so the dnB transfer hides behind K1.

The result on N GPUs is bit-identical to the single-GPU result (tests/test_sharding_cpu.py with gloo + the oracle as
compute; tests/test_gpu_multi.py on >= 2 GPUs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass
class ShardPlan:
    """Row ranges [a, b) per launch of one frame, in chain order: K1, K2, K3 pass 0..n-1, K4 (SSGI mode only),
    for every row block this rank owns."""

    height: int
    world: int
    rank: int
    n_poisson_passes: int  # 2 * denoiseIterations
    radius: float
    ssgi_mode: bool = True
    blocks_per_rank: int = 1
    mirror: bool = False  # True: odd super-blocks are assigned in REVERSE rank order (boustrophedon), see block_of()

    K2_NEIGHBOURHOOD_ROWS = 2  # 5x5 clamp window (reproject.frag:57-59)
    K4_INPUT_ROWS = 1          # literal bilinear fetch of the LINEAR Poisson targets at the pixel centre

    def block_of(self, rank: int, j: int):
        """rows of the block `rank` owns inside super-block j.  With `mirror`, odd super-blocks run in reverse rank order, so a
        rank that gets the cheapest end of one super-block (sky) gets the most expensive end of the next (floor): 2 blocks per
        rank already balance a vertical cost gradient, at half the halo recompute of a 4-block cyclic assignment."""
        pos = (self.world - 1 - rank) if (self.mirror and j % 2 == 1) else rank
        b = j * self.world + pos
        return (b * self.block_rows, (b + 1) * self.block_rows)

    def reversed_order(self, j: int) -> bool:
        return self.mirror and j % 2 == 1

    def __post_init__(self):
        nb = self.world * self.blocks_per_rank
        if self.height % nb:
            raise ValueError(f"height {self.height} is not divisible by world size x blocks per rank = {nb}")
        self.block_rows = self.height // nb
        self.rows_per_rank = self.block_rows * self.blocks_per_rank
        self.blocks = [self.block_of(self.rank, j) for j in range(self.blocks_per_rank)]
        self.r0, self.r1 = self.blocks[0]  # (single-block plans: the contiguous band)
        self.poisson_halo = int(math.ceil(self.radius)) + 1  # taps reach ceil(radius) rows, +1 for the bilinear footprint

    def _expand(self, rng, rows):
        return (max(0, rng[0] - rows), min(self.height, rng[1] + rows))

    def ranges_for(self, own) -> list:
        n = self.n_poisson_passes
        k3 = [None] * n
        nxt = own
        if n:
            k3[n - 1] = self._expand(own, self.K4_INPUT_ROWS if self.ssgi_mode else 0)
            for j in range(n - 2, -1, -1):
                k3[j] = self._expand(k3[j + 1], self.poisson_halo)
            nxt = self._expand(k3[0], self.poisson_halo)
        k2 = nxt
        k1 = self._expand(k2, self.K2_NEIGHBOURHOOD_ROWS)
        out = [k1, k2, *k3]
        if self.ssgi_mode:
            out.append(own)
        return out

    @property
    def ranges(self) -> list:
        """single-block plans: the per-launch ranges of the band"""
        return self.ranges_for(self.blocks[0])

    @property
    def block_ranges(self) -> list:
        """[block][launch] -> (row0, row1)"""
        return [self.ranges_for(b) for b in self.blocks]

    @property
    def n_launches(self) -> int:
        return 2 + self.n_poisson_passes + (1 if self.ssgi_mode else 0)

    def super_block(self, j: int):
        """rows of super-block j: N consecutive blocks, one per rank, in rank order (an in-place all-gather unit)"""
        return (j * self.world * self.block_rows, (j + 1) * self.world * self.block_rows)

    @property
    def recompute_overhead(self) -> float:
        """extra rows launched / rows owned (the price of exchanging nothing per pass)"""
        tot = sum((b - a) for rs in self.block_ranges for a, b in rs)
        return tot / (self.n_launches * self.rows_per_rank) - 1.0

    @property
    def local_input_rows(self) -> list:
        """Rows of the PER-PIXEL input planes (G-buffer, direct light) this rank reads: the K1 range of each block, which contains
        every later launch's range and the rows its Poisson taps reach.  The planes sampled at arbitrary screen positions (depth:
        ray-march taps; velocity: reprojected uv) are needed whole and are all-gathered from the ranks' own rows instead."""
        return [rs[0] for rs in self.block_ranges]

    @property
    def gathered_planes(self):
        """chain outputs (`rfx_ssgi_chain_output` index) that are all-gathered after the frame; `composed` first"""
        return (0, 4, 5) if self.ssgi_mode else (4,)


class _CudaBytes:
    """__cuda_array_interface__ view of a raw device allocation (an rfx_plane) so torch/NCCL can address it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class _PlaneRef:
    """adapter: SsgiChain.render takes objects with a `.p` rfx_plane"""

    def __init__(self, p):
        self.p = p


class ShardedSsgiChain:
    """The native SSGI chain on this rank's row blocks of a W x H frame + the per-frame all-gathers of the produced planes."""

    def __init__(self, ctx, chain_options, group=None, blocks_per_rank: int = 4, overlap: bool = True, mirror: bool = False):
        import torch
        import torch.distributed as dist

        from . import abi, engine

        self.dist, self.torch, self.group = dist, torch, group
        self._abi = abi
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.chain = engine.SsgiChain(ctx, chain_options)
        self.ctx = ctx
        self.overlap = overlap
        self.coalesce = True
        self.plan = ShardPlan(chain_options.height, self.world, self.rank, 2 * chain_options.denoise_iterations, chain_options.radius,
                              chain_options.mode == abi.MODE_SSGI, blocks_per_rank, mirror)
        # A mirrored (boustrophedon) assignment needs an all-gather whose output order is the reverse rank order; torch.distributed
        # sorts the ranks of every new_group(), so that order is not expressible with NCCL collectives here.  The plan supports it
        # (ShardPlan.block_of) for the planned peer-store exchange; the NCCL path uses the forward cyclic order.
        self.group_rev = None
        if self.world > 1 and mirror and blocks_per_rank > 1:
            raise ValueError("mirror=True needs a reverse-rank-order gather, which torch.distributed process groups cannot express")
        # a dedicated torch stream: the kernels and the NCCL collectives are ordered on it.  (A NULL stream handle means "the
        # context's own stream" to the C ABI, so torch's default stream cannot be used here.)
        self.stream = torch.cuda.Stream(device=torch.device("cuda", ctx.device))
        self._tensors = {}
        for which in self.plan.gathered_planes:
            p = self.chain.output(which)
            nbytes = int(p.pitch) * int(p.height)
            t = torch.as_tensor(_CudaBytes(p.ptr, nbytes), device=torch.device("cuda", ctx.device))
            self._tensors[which] = (t, int(p.pitch))
        self._pending = {}  # plane index -> [Work]: all-gathers of the previous frame not yet waited for

    def _wait(self, planes):
        for which in planes:
            for w in self._pending.pop(which, []):
                w.wait()  # makes self.stream wait for the collective

    def _gather(self, planes):
        """Launches the in-place all-gathers of `planes` as ONE coalesced NCCL group (ncclGroupStart/End: a single launch
        instead of len(planes) x blocks_per_rank); falls back to one async collective per super-block."""
        def calls(rev):
            for which in planes:
                t, pitch = self._tensors[which]
                for j, (b0, b1) in enumerate(self.plan.blocks):
                    if self.plan.reversed_order(j) != rev:
                        continue
                    s0, s1 = self.plan.super_block(j)
                    yield t[s0 * pitch:s1 * pitch], t[b0 * pitch:b1 * pitch]  # in place: every rank's block lands at its own rows

        works = []
        cm_fn = getattr(self.dist, "_coalescing_manager", None)
        for rev, group in ((False, self.group), (True, self.group_rev)):  # reversed super-blocks gather over the reversed group
            todo = list(calls(rev))
            if not todo:
                continue
            done = False
            if self.coalesce and cm_fn is not None:
                try:
                    with cm_fn(group=group, device=self.torch.device("cuda", self.ctx.device), async_ops=True) as cm:
                        for out, own in todo:
                            self.dist.all_gather_into_tensor(out, own, group=group)
                    works.append(cm)
                    done = True
                except Exception:  # private API: keep working if its signature changes
                    self.coalesce = False
            if not done:
                works += [self.dist.all_gather_into_tensor(out, own, group=group, async_op=True) for out, own in todo]
        self._pending[planes[0]] = works

    def render(self, cam, depth, gbuffer, velocity, direct_light, camera_pos, camera_moved: bool):
        """Enqueues one frame on self.stream (two phases) and the asynchronous all-gathers of its outputs."""
        torch, plan = self.torch, self.plan
        args = (cam, depth, gbuffer, velocity, direct_light, camera_pos, camera_moved)
        with torch.cuda.stream(self.stream):
            if self.world == 1:
                self.chain.render(*args, stream=self.stream.cuda_stream)
                return
            br, nl = plan.block_ranges, plan.n_launches
            first = plan.gathered_planes[0]
            if self.overlap and plan.ssgi_mode:
                self._wait([first])                                   # K1 samples last frame's `composed`
                self.chain.render(*args, stream=self.stream.cuda_stream, ranges=br, launches=(0, 1))
                self._wait(plan.gathered_planes[1:])                  # K2 samples last frame's dnB history
                self.chain.render(*args, stream=self.stream.cuda_stream, ranges=br, launches=(1, nl))
            else:
                self._wait(plan.gathered_planes)
                self.chain.render(*args, stream=self.stream.cuda_stream, ranges=br, launches=(0, nl))
            self._gather(plan.gathered_planes[:1])                    # `composed` first: the next frame needs it first
            if len(plan.gathered_planes) > 1:
                self._gather(plan.gathered_planes[1:])                # dnB[0..1]: pending under key gathered_planes[1]
            if not self.overlap:
                self._wait(plan.gathered_planes)

    def finish(self):
        """wait for the outstanding all-gathers (before reading the planes or tearing down)"""
        with self.torch.cuda.stream(self.stream):
            self._wait(list(self._pending))
        self.stream.synchronize()

    # ---- host-buffer path ------------------------------------------------------------------------------------------------
    # Every rank holds (or maps) the frame's host planes but moves only its share over PCIe: its own rows of depth and velocity -
    # the two planes sampled anywhere on screen - which are then all-gathered over NVLink (own communicator, so the gather of
    # frame i+1 is not queued behind frame i's output gathers), and the K1-range rows of the G-buffer and direct light.  H2D,
    # kernels and D2H run on three streams with two staging sets, like rfx_ssgi_chain_submit_host on one GPU.
    INPUTS = (("depth", 4, True), ("gbuffer", 16, False), ("velocity", 16, True), ("direct", 8, False))  # name, bytes/px, gathered

    def _host_init(self):
        torch, abi = self.torch, self._abi
        dev = torch.device("cuda", self.ctx.device)
        W, H = self.chain.opt.width, self.chain.opt.height
        fmts = dict(depth=abi.FMT_R32F, gbuffer=abi.FMT_RGBA32F, velocity=abi.FMT_RGBA32F, direct=abi.FMT_RGBA16F)
        self._in = []
        for _ in range(2):
            st = {}
            for name, bpp, _g in self.INPUTS:
                t = torch.zeros(H * W * bpp, dtype=torch.uint8, device=dev)
                pl = abi.Plane()
                pl.ptr, pl.width, pl.height, pl.pitch, pl.format = t.data_ptr(), W, H, W * bpp, fmts[name]
                st[name] = (t, pl, W * bpp)
            self._in.append(st)
        self._out_dev = [torch.empty(self.plan.rows_per_rank * W * 16, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.up_stream, self.dn_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._ev_up = [torch.cuda.Event() for _ in range(2)]
        self._ev_rendered = [torch.cuda.Event() for _ in range(2)]
        self._ev_dn = [torch.cuda.Event() for _ in range(2)]
        self.in_group = self.dist.new_group(ranks=list(range(self.world))) if self.world > 1 else None
        self._host_frames = 0

    def submit_host(self, cam, host: dict, camera_pos, camera_moved: bool, out_host):
        """One frame from host planes (dict name -> CPU tensor of the FULL frame, pinned for asynchronous copies) to this rank's
        rows of `composed` in out_host (CPU float32 tensor (rows_per_rank, W, 4), blocks in plan order).  Returns after enqueueing."""
        torch, plan = self.torch, self.plan
        if not hasattr(self, "_in"):
            self._host_init()
        W, H = self.chain.opt.width, self.chain.opt.height
        k = self._host_frames & 1
        st = self._in[k]
        works = []
        with torch.cuda.stream(self.up_stream):
            if self._host_frames >= 2:
                self.up_stream.wait_event(self._ev_rendered[k])   # frame i-2 no longer reads this staging set
            for name, bpp, gathered in self.INPUTS:
                if name not in host or host[name] is None:
                    continue
                t, _pl, pitch = st[name]
                dev2d, host2d = t.view(H, pitch), host[name].view(torch.uint8).view(H, pitch)
                for a, b in (plan.blocks if (gathered and self.world > 1) else plan.local_input_rows if self.world > 1 else [(0, H)]):
                    dev2d[a:b].copy_(host2d[a:b], non_blocking=True)
            self._ev_up[k].record(self.up_stream)
            if self.world > 1:
                todo = []
                for name, bpp, gathered in self.INPUTS:
                    if gathered and host.get(name) is not None:
                        t, _pl, pitch = st[name]
                        for j, (b0, b1) in enumerate(plan.blocks):
                            s0, s1 = plan.super_block(j)
                            todo.append((t[s0 * pitch:s1 * pitch], t[b0 * pitch:b1 * pitch]))
                works = self._all_gather(todo, self.in_group)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._ev_up[k])
            for w in works:
                w.wait()
        pw = lambda n: _PlaneRef(st[n][1]) if host.get(n) is not None else None  # noqa: E731
        self.render(cam, pw("depth"), pw("gbuffer"), pw("velocity"), pw("direct"), camera_pos, camera_moved)
        comp, cpitch = self._composed()
        with torch.cuda.stream(self.stream):
            # snapshot this rank's rows of `composed`, so the next frame's K4 may overwrite them while the D2H copy still runs
            out2d, off = self._out_dev[k].view(plan.rows_per_rank, W * 16), 0
            for b0, b1 in plan.blocks:
                out2d[off:off + (b1 - b0)].copy_(comp.view(H, cpitch)[b0:b1, :W * 16], non_blocking=True)
                off += b1 - b0
            self._ev_rendered[k].record(self.stream)
        with torch.cuda.stream(self.dn_stream):
            self.dn_stream.wait_event(self._ev_rendered[k])
            out_host.view(torch.uint8).view(-1).copy_(self._out_dev[k], non_blocking=True)
            self._ev_dn[k].record(self.dn_stream)
        self._host_frames += 1

    def wait_host(self, max_in_flight: int = 0):
        """Blocks until at most max_in_flight (0 or 1) submitted frames are incomplete (their out_host rows not yet written)."""
        n = getattr(self, "_host_frames", 0)
        if n == 0:
            return
        if max_in_flight <= 0:
            self._ev_dn[(n - 1) & 1].synchronize()
        elif n >= 2:
            self._ev_dn[(n - 2) & 1].synchronize()

    def _composed(self):
        if 0 in self._tensors:
            return self._tensors[0]
        p = self.chain.output(0)
        t = self.torch.as_tensor(_CudaBytes(p.ptr, int(p.pitch) * int(p.height)), device=self.torch.device("cuda", self.ctx.device))
        self._tensors[0] = (t, int(p.pitch))
        return self._tensors[0]

    def _all_gather(self, todo, group):
        """in-place all-gathers [(out, own)] as one coalesced NCCL group when the private coalescing API is usable"""
        if not todo:
            return []
        cm_fn = getattr(self.dist, "_coalescing_manager", None)
        if self.coalesce and cm_fn is not None:
            try:
                with cm_fn(group=group, device=self.torch.device("cuda", self.ctx.device), async_ops=True) as cm:
                    for out, own in todo:
                        self.dist.all_gather_into_tensor(out, own, group=group)
                return [cm]
            except Exception:
                self.coalesce = False
        return [self.dist.all_gather_into_tensor(out, own, group=group, async_op=True) for out, own in todo]

    @property
    def host_bytes_per_frame(self):
        """(H2D, D2H) bytes this rank moves per frame on the host path"""
        W = self.chain.opt.width
        own = self.plan.rows_per_rank if self.world > 1 else self.chain.opt.height
        loc = sum(b - a for a, b in self.plan.local_input_rows) if self.world > 1 else self.chain.opt.height
        h2d = sum((own if g else loc) * W * bpp for _n, bpp, g in self.INPUTS)
        return h2d, own * W * 16

    @property
    def exchange_bytes_per_frame(self) -> int:
        """bytes this rank RECEIVES per frame"""
        return sum((len(t) // self.world) * (self.world - 1) for t, _ in self._tensors.values())

    def close(self):
        if self.world > 1:
            self.finish()
        self.chain.close()
