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

Overlap: a frame is issued in three parts (rfx_ssgi_chain_render_part).  K1's ray march reads the depth plane only; K1's
shading samples last frame's `composed`; K2..K4 need the `dnB` history.  The exchanges are launched asynchronously after K4
(composed first); the next frame marches its rays immediately, waits for `composed` only before the K1 shading part and for
`dnB` only before K2 - so the `composed` transfer hides behind the march and the `dnB` transfer behind the whole of K1.

Adaptive bands (`balance="adaptive"`, the default of bench.py): one CONTIGUOUS band per rank whose height follows the measured
kernel time.  Every rank times its own kernels with CUDA events; every few frames the ranks exchange those times (a few floats
over a gloo control group), model the cost per row as piecewise constant over the bands and move the band borders towards equal
cost (`rebalance`, damped, 16-row aligned).  All cross-frame state lives in the exchanged planes, so the split may change from
frame to frame without copying anything, and the result stays bit-identical.  Against block-cyclic assignment this keeps the
K1 tap working set of a rank local (one band + ray reach instead of the whole frame), halves the halo rows and removes the
partial-tile waste of many small blocks; the bands have different heights, so the exchange is a grouped send/recv.

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
    bounds: tuple | None = None  # explicit band borders (world + 1 ascending rows, 0 .. height): one contiguous band per rank

    K2_NEIGHBOURHOOD_ROWS = 2  # 5x5 clamp window (reproject.frag:57-59)
    K4_INPUT_ROWS = 1          # literal bilinear fetch of the LINEAR Poisson targets at the pixel centre

    def block_of(self, rank: int, j: int):
        """rows of the block `rank` owns inside super-block j.  With `mirror`, odd super-blocks run in reverse rank order, so a
        rank that gets the cheapest end of one super-block (sky) gets the most expensive end of the next (floor): 2 blocks per
        rank already balance a vertical cost gradient, at half the halo recompute of a 4-block cyclic assignment."""
        if self.bounds is not None:
            return (self.bounds[rank], self.bounds[rank + 1])
        pos = (self.world - 1 - rank) if (self.mirror and j % 2 == 1) else rank
        b = j * self.world + pos
        return (b * self.block_rows, (b + 1) * self.block_rows)

    def reversed_order(self, j: int) -> bool:
        return self.mirror and j % 2 == 1

    @property
    def p2p(self) -> bool:
        """the exchange cannot be a rank-ordered in-place all-gather (mirrored order or unequal bands)"""
        return self.mirror or self.bounds is not None

    def __post_init__(self):
        if self.bounds is not None:
            b = tuple(int(x) for x in self.bounds)
            if len(b) != self.world + 1 or b[0] != 0 or b[-1] != self.height or any(b[i] >= b[i + 1] for i in range(self.world)):
                raise ValueError(f"bounds {b} are not {self.world + 1} ascending rows from 0 to {self.height}")
            self.bounds, self.blocks_per_rank, self.block_rows = b, 1, None
            self.blocks = [self.block_of(self.rank, 0)]
            self.rows_per_rank = self.blocks[0][1] - self.blocks[0][0]
            self.r0, self.r1 = self.blocks[0]
            self.poisson_halo = int(math.ceil(self.radius)) + 1
            return
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


def rebalance(bounds, costs, measured_bounds=None, align: int = 16, min_rows: int = 64, max_share: float = 4.0, damping: float = 0.6):
    """New band borders from per-rank costs (any unit) measured with `measured_bounds` (default: `bounds`).

    The cost per row is modelled as constant inside each measured band; the ideal border k is the row where the cumulative
    cost reaches k/N of the total.  The borders move `damping` of the way from `bounds` to the ideal ones (the model is coarse:
    damping avoids overshoot), are rounded to `align` rows (the kernels' tile height, so no band ends in a partial tile) and
    kept at least `min_rows` and at most max_share x the mean height apart.  Pure and deterministic: every rank computes the
    same borders from the same gathered costs."""
    mb = list(measured_bounds if measured_bounds is not None else bounds)
    n, H = len(costs), bounds[-1]
    costs = [max(float(c), 1e-9) for c in costs]
    total = sum(costs)
    ideal, k, acc = [0], 0, 0.0
    for i in range(1, n):  # invert the piecewise-linear cumulative cost at i/n of the total
        want = total * i / n
        while k < n - 1 and acc + costs[k] < want:
            acc += costs[k]
            k += 1
        ideal.append(mb[k] + (want - acc) / costs[k] * (mb[k + 1] - mb[k]))
    ideal.append(H)
    lo_h, hi_h = min_rows, max(min_rows, int(max_share * H / n))
    out = [0]
    for i in range(1, n):
        b = bounds[i] + damping * (ideal[i] - bounds[i])
        b = int(round(b / align)) * align
        b = max(b, out[-1] + lo_h)                    # not thinner than min_rows ...
        b = min(b, out[-1] + hi_h)                    # ... nor taller than max_share x the mean
        b = min(b, H - (n - i) * lo_h)                # leave room for the bands below
        out.append(b)
    out.append(H)
    return tuple(out)


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

    def __init__(self, ctx, chain_options, group=None, blocks_per_rank: int = 4, overlap: bool = True, mirror: bool = False,
                 balance: str = "static", rebalance_every: int = 4, rebalance_lag: int = 2, split_k1: bool = True, dual_comm: bool = False):
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
        self.split_k1 = split_k1  # overlap mode: K1 as ray march + shading, so only the shading waits for the `composed` exchange
        self.coalesce = True
        self._plan_args = (chain_options.height, self.world, self.rank, 2 * chain_options.denoise_iterations, chain_options.radius,
                           chain_options.mode == abi.MODE_SSGI)
        if balance not in ("static", "adaptive"):
            raise ValueError("balance must be 'static' (block-cyclic / mirrored blocks) or 'adaptive' (one cost-balanced band per rank)")
        self.balance = balance if self.world > 1 else "static"
        self.rebalance_every, self.rebalance_lag = max(1, rebalance_every), max(1, rebalance_lag)
        self._frame, self._timing, self._ev_pool, self.last_costs, self._host_span = 0, [], [], None, None
        if self.balance == "adaptive":
            H, n = chain_options.height, self.world
            if H < n * 64:
                raise ValueError("adaptive bands need at least 64 rows per rank")
            eq = tuple(int(round(H * i / n / 16.0)) * 16 for i in range(n)) + (H,)
            self.plan = ShardPlan(*self._plan_args, bounds=eq)
            # host-side control plane: a few floats per rebalance, over gloo (no device sync, no NCCL stream involved)
            self.ctl_group = dist.new_group(ranks=[self._global_rank(r) for r in range(self.world)], backend="gloo")
        else:
            self.plan = ShardPlan(*self._plan_args, blocks_per_rank, mirror)
        # A mirrored (boustrophedon) assignment puts the ranks of odd super-blocks in reverse order, which a rank-ordered
        # all_gather_into_tensor cannot write in place (torch.distributed sorts the ranks of every new_group()).  Mirrored plans
        # therefore exchange with grouped point-to-point transfers (one ncclGroup of send/recv pairs per exchange, every block
        # landing directly at its rows); plain cyclic plans keep the in-place all-gather.
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
        # dual_comm (experiment, unmeasured in round 1): the dnB exchange runs on a second communicator so it moves concurrently with
        # the `composed` exchange instead of queueing behind it on one NCCL stream.  Every rank issues both in the same order.
        self.group2 = None
        if dual_comm and self.world > 1:
            self.group2 = dist.new_group(ranks=[self._global_rank(r) for r in range(self.world)])

    def _wait(self, planes):
        for which in planes:
            for w in self._pending.pop(which, []):
                w.wait()  # makes self.stream wait for the collective

    def _exchange(self, items, group):
        """Every rank's blocks of the planes `items` = [(flat byte tensor, pitch)] to every other rank, in place, as ONE NCCL group:
        in-place all-gathers per super-block for cyclic plans, send/recv pairs for mirrored plans and unequal bands.  Returns the Work handles."""
        plan = self.plan
        if plan.p2p:
            ops = []
            for t, pitch in items:
                for j, (b0, b1) in enumerate(plan.blocks):
                    for peer in range(self.world):  # both sides walk (plane, block) in the same order, so the k-th send to a peer meets its k-th recv
                        if peer == self.rank:
                            continue
                        p0, p1 = plan.block_of(peer, j)
                        ops.append(self.dist.P2POp(self.dist.isend, t[b0 * pitch:b1 * pitch], self._global_rank(peer), group))
                        ops.append(self.dist.P2POp(self.dist.irecv, t[p0 * pitch:p1 * pitch], self._global_rank(peer), group))
            return list(self.dist.batch_isend_irecv(ops)) if ops else []
        todo = []
        for t, pitch in items:
            for j, (b0, b1) in enumerate(plan.blocks):
                s0, s1 = plan.super_block(j)
                todo.append((t[s0 * pitch:s1 * pitch], t[b0 * pitch:b1 * pitch]))  # in place: every rank's block lands at its own rows
        return self._all_gather(todo, group)

    def _gather(self, planes, group=None):
        """launches the exchange of the chain outputs `planes`; the handles wait under the first plane's key"""
        self._pending[planes[0]] = self._exchange([self._tensors[w] for w in planes], group if group is not None else self.group)

    def _global_rank(self, r: int) -> int:
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def _event(self):
        return self._ev_pool.pop() if self._ev_pool else self.torch.cuda.Event(enable_timing=True)

    def _maybe_rebalance(self):
        """adaptive bands: every `rebalance_every` frames move the band borders towards equal measured kernel time.  Uses the
        newest measurement that is at least `rebalance_lag` frames old (long finished, so the event wait is immediate); the ranks
        hold identical frame counters and gather identical costs, so they all derive the same borders."""
        if self.balance != "adaptive":
            return
        f = self._frame
        if f % self.rebalance_every == 0:
            cand = [t for t in self._timing if t[0] <= f - self.rebalance_lag]
            if cand:
                _fr, mb, spans, hspan = cand[-1]
                spans[-1][1].synchronize()
                cost = sum(a.elapsed_time(b) for a, b in spans)
                if hspan is not None:  # host path: a rank is as slow as the slower of its kernels and its PCIe upload (both grow with the band)
                    hspan[1].synchronize()
                    cost = max(cost, hspan[0].elapsed_time(hspan[1]))
                mine = self.torch.tensor([cost], dtype=self.torch.float64)
                allc = [self.torch.zeros(1, dtype=self.torch.float64) for _ in range(self.world)]
                self.dist.all_gather(allc, mine, group=self.ctl_group)
                self.last_costs = [float(c) for c in allc]
                nb = rebalance(self.plan.bounds, self.last_costs, mb, max_share=self.MAX_SHARE)
                if nb != self.plan.bounds:
                    self.plan = ShardPlan(*self._plan_args, bounds=nb)
        keep = []
        for t in self._timing:  # recycle the events of measurements that can no longer be chosen
            if t[0] > f - self.rebalance_lag - self.rebalance_every - 1:
                keep.append(t)
            else:
                for a, b in t[2] + ([t[3]] if t[3] is not None else []):
                    self._ev_pool += [a, b]
        self._timing = keep

    def render(self, cam, depth, gbuffer, velocity, direct_light, camera_pos, camera_moved: bool, _rebalanced: bool = False):
        """Enqueues one frame on self.stream (two phases) and the asynchronous exchange of its outputs."""
        torch = self.torch
        args = (cam, depth, gbuffer, velocity, direct_light, camera_pos, camera_moved)
        with torch.cuda.stream(self.stream):
            if self.world == 1:
                self.chain.render(*args, stream=self.stream.cuda_stream)
                return
            if not _rebalanced:
                self._maybe_rebalance()
            plan = self.plan
            br, nl = plan.block_ranges, plan.n_launches
            first = plan.gathered_planes[0]
            spans = []

            def timed(launches=None, part=None):  # kernels only: the waits for the exchanges stay outside the measured spans
                a, b = self._event(), self._event()
                a.record(self.stream)
                if part is None:
                    self.chain.render(*args, stream=self.stream.cuda_stream, ranges=br, launches=launches)
                else:
                    self.chain.render_part(part, *args, stream=self.stream.cuda_stream, ranges=br)
                b.record(self.stream)
                spans.append((a, b))

            if self.overlap and plan.ssgi_mode and self.split_k1:
                timed(part=0)                                         # K1 ray march: depth only - runs under the exchange of `composed`
                self._wait([first])                                   # K1 shading samples last frame's `composed`
                timed(part=1)
                self._wait(plan.gathered_planes[1:])                  # K2 samples last frame's dnB history
                timed(part=2)
            elif self.overlap and plan.ssgi_mode:
                self._wait([first])                                   # K1 samples last frame's `composed`
                timed((0, 1))
                self._wait(plan.gathered_planes[1:])                  # K2 samples last frame's dnB history
                timed((1, nl))
            else:
                self._wait(plan.gathered_planes)
                timed((0, nl))
            self._gather(plan.gathered_planes[:1])                    # `composed` first: the next frame needs it first
            if len(plan.gathered_planes) > 1:                         # dnB[0..1]: pending under key gathered_planes[1]
                self._gather(plan.gathered_planes[1:], self.group2)   # (dual_comm: on its own communicator, concurrent with `composed`)
            if not self.overlap:
                self._wait(plan.gathered_planes)
            self._timing.append((self._frame, plan.bounds, spans, self._host_span))
            self._host_span = None
            self._frame += 1

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
    MAX_SHARE = 4.0  # tallest adaptive band, in units of the mean band height (sizes the read-back staging)
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
        cap = H if self.world == 1 else min(H, int(self.MAX_SHARE * H / self.world) + 16)   # adaptive bands stay below MAX_SHARE x the mean height
        self._out_dev = [torch.empty(cap * W * 16, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.up_stream, self.dn_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._ev_up = [torch.cuda.Event() for _ in range(2)]
        self._ev_rendered = [torch.cuda.Event() for _ in range(2)]
        self._ev_dn = [torch.cuda.Event() for _ in range(2)]
        self.in_group = self.dist.new_group(ranks=[self._global_rank(r) for r in range(self.world)]) if self.world > 1 else None
        self._host_frames = 0

    def submit_host(self, cam, host: dict, camera_pos, camera_moved: bool, out_host):
        """One frame from host planes (dict name -> CPU tensor of the FULL frame, pinned for asynchronous copies) to this rank's
        rows of `composed`, written to the start of out_host (CPU float32 tensor of at least rows_per_rank x W x 4; with adaptive
        bands up to MAX_SHARE x the mean band height).  Returns the row blocks [(r0, r1), ...] the frame's rows belong to, after enqueueing."""
        torch = self.torch
        if not hasattr(self, "_in"):
            self._host_init()
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                self._maybe_rebalance()   # before the uploads: they follow this frame's bands
        plan = self.plan
        W, H = self.chain.opt.width, self.chain.opt.height
        k = self._host_frames & 1
        st = self._in[k]
        works = []
        with torch.cuda.stream(self.up_stream):
            if self._host_frames >= 2:
                self.up_stream.wait_event(self._ev_rendered[k])   # frame i-2 no longer reads this staging set
            eu0, eu1 = self._event(), self._event()
            eu0.record(self.up_stream)
            for name, bpp, gathered in self.INPUTS:
                if name not in host or host[name] is None:
                    continue
                t, _pl, pitch = st[name]
                dev2d, host2d = t.view(H, pitch), host[name].view(torch.uint8).view(H, pitch)
                for a, b in (plan.blocks if (gathered and self.world > 1) else plan.local_input_rows if self.world > 1 else [(0, H)]):
                    dev2d[a:b].copy_(host2d[a:b], non_blocking=True)
            self._ev_up[k].record(self.up_stream)
            eu1.record(self.up_stream)
            self._host_span = (eu0, eu1)
            if self.world > 1:
                works = self._exchange([(st[n][0], st[n][2]) for n, _b, g in self.INPUTS if g and host.get(n) is not None], self.in_group)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self._ev_up[k])
            for w in works:
                w.wait()
        pw = lambda n: _PlaneRef(st[n][1]) if host.get(n) is not None else None  # noqa: E731
        self.render(cam, pw("depth"), pw("gbuffer"), pw("velocity"), pw("direct"), camera_pos, camera_moved, _rebalanced=True)
        comp, cpitch = self._composed()
        with torch.cuda.stream(self.stream):
            # snapshot this rank's rows of `composed`, so the next frame's K4 may overwrite them while the D2H copy still runs
            nbytes = plan.rows_per_rank * W * 16
            out2d, off = self._out_dev[k][:nbytes].view(plan.rows_per_rank, W * 16), 0
            for b0, b1 in plan.blocks:
                out2d[off:off + (b1 - b0)].copy_(comp.view(H, cpitch)[b0:b1, :W * 16], non_blocking=True)
                off += b1 - b0
            self._ev_rendered[k].record(self.stream)
        with torch.cuda.stream(self.dn_stream):
            self.dn_stream.wait_event(self._ev_rendered[k])
            out_host.view(torch.uint8).view(-1)[:nbytes].copy_(self._out_dev[k][:nbytes], non_blocking=True)
            self._ev_dn[k].record(self.dn_stream)
        self._host_frames += 1
        return list(plan.blocks)

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
        H = self.chain.opt.height
        return sum((H - self.plan.rows_per_rank) * pitch for _t, pitch in self._tensors.values())

    def close(self):
        if self.world > 1:
            self.finish()
        self.chain.close()
