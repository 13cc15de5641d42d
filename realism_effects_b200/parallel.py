"""Row-block sharding of the SSGI chain across N GPUs (one process per GPU, torch.distributed / NCCL over NVLink).

Every kernel of the path writes disjoint output rows, so GPU g of N owns rows [g*H/N, (g+1)*H/N) of every produced
plane.  Two kinds of inputs cross row-block borders:

  * bounded stencils — K2's 5x5 neighbourhood of ssgiOut (2 rows), K3's Poisson taps (ceil(radius)+1 rows per pass,
    bilinear footprint included), K4's pixel-centre bilinear fetch of dnB (1 row).  Instead of one halo exchange per
    pass, each rank RECOMPUTES the halo rows itself: pass k is launched on a row range widened by the halos of all the
    passes after it (`ShardPlan`).  Kernels are bit-identical under row sharding, so the recomputed rows equal the
    owner's rows bit for bit, and no per-pass NCCL latency is paid.
  * arbitrary-uv gathers — K1 samples `composed` at ray hit points and K2 samples the history `dnB[0..1]` at
    reprojected uvs anywhere on screen, so these three produced planes are all-gathered once per frame (32 B/px of each
    rank's rows).  The static inputs (depth, gBuffer, velocity, directLight) are given to every rank in full.

The result on N GPUs is bit-identical to the single-GPU result (tests/test_sharding_cpu.py with gloo + the oracle as
compute; tests/test_gpu_multi.py on >= 2 GPUs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass
class ShardPlan:
    """Row ranges [a, b) per launch of one frame, in chain order: K1, K2, K3 pass 0..n-1, K4 (SSGI mode only)."""

    height: int
    world: int
    rank: int
    n_poisson_passes: int  # 2 * denoiseIterations
    radius: float
    ssgi_mode: bool = True

    K2_NEIGHBOURHOOD_ROWS = 2  # 5x5 clamp window (reproject.frag:57-59)
    K4_INPUT_ROWS = 1          # literal bilinear fetch of the LINEAR Poisson targets at the pixel centre

    def __post_init__(self):
        if self.height % self.world:
            raise ValueError(f"height {self.height} is not divisible by world size {self.world}")
        self.rows_per_rank = self.height // self.world
        self.r0, self.r1 = self.rank * self.rows_per_rank, (self.rank + 1) * self.rows_per_rank
        self.poisson_halo = int(math.ceil(self.radius)) + 1  # taps reach ceil(radius) rows, +1 for the bilinear footprint

    def _expand(self, rng, rows):
        return (max(0, rng[0] - rows), min(self.height, rng[1] + rows))

    @property
    def ranges(self) -> list:
        own = (self.r0, self.r1)
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
    def recompute_overhead(self) -> float:
        """extra rows launched / rows owned (the price of exchanging nothing per pass)"""
        rs = self.ranges
        return sum((b - a) for a, b in rs) / (len(rs) * self.rows_per_rank) - 1.0

    @property
    def gathered_planes(self):
        """chain outputs (`rfx_ssgi_chain_output` index) that are all-gathered after the frame"""
        return (0, 4, 5) if self.ssgi_mode else (4,)


class _CudaBytes:
    """__cuda_array_interface__ view of a raw device allocation (an rfx_plane) so torch/NCCL can address it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class ShardedSsgiChain:
    """The native SSGI chain on this rank's row block of a W x H frame + the per-frame all-gather of the produced planes."""

    def __init__(self, ctx, chain_options, group=None):
        import torch
        import torch.distributed as dist

        from . import abi, engine

        self.dist, self.torch, self.group = dist, torch, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.chain = engine.SsgiChain(ctx, chain_options)
        self.ctx = ctx
        self.plan = ShardPlan(chain_options.height, self.world, self.rank, 2 * chain_options.denoise_iterations, chain_options.radius,
                              chain_options.mode == abi.MODE_SSGI)
        # a dedicated torch stream: the kernels and the NCCL collectives are ordered on it.  (A NULL stream handle means "the
        # context's own stream" to the C ABI, so torch's default stream cannot be used here.)
        self.stream = torch.cuda.Stream(device=torch.device("cuda", ctx.device))
        self._tensors = {}
        for which in self.plan.gathered_planes:
            p = self.chain.output(which)
            nbytes = int(p.pitch) * int(p.height)
            t = torch.as_tensor(_CudaBytes(p.ptr, nbytes), device=torch.device("cuda", ctx.device))
            self._tensors[which] = (t, int(p.pitch))

    def render(self, cam, depth, gbuffer, velocity, direct_light, camera_pos, camera_moved: bool):
        """Enqueues the frame on self.stream, then the all-gathers (torch's NCCL wrapper orders them after the kernels on that
        stream and makes the stream wait for their completion)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.chain.render(cam, depth, gbuffer, velocity, direct_light, camera_pos, camera_moved, stream=self.stream.cuda_stream,
                              ranges=self.plan.ranges if self.world > 1 else None)
            if self.world > 1:
                for which, (t, pitch) in self._tensors.items():
                    own = t[self.plan.r0 * pitch:self.plan.r1 * pitch]
                    self.dist.all_gather_into_tensor(t, own, group=self.group)  # in place: block g lands at rows [g*H/N, (g+1)*H/N)

    @property
    def exchange_bytes_per_frame(self) -> int:
        """bytes this rank RECEIVES per frame"""
        return sum((len(t) // self.world) * (self.world - 1) for t, _ in self._tensors.values())

    def close(self):
        self.chain.close()
