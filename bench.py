#!/usr/bin/env python
"""bench.py — SSGI+denoise Mpixels/s at 3840x2160 on N B200s (BASELINE.json metric, config C3).

One "step" = one frame of the SSGI chain over one batch of synthetic G-buffer planes:
  K1 SSGI trace (steps 20 / refine 5) -> K2 temporal reprojection (2 planes) ->
  K3 Poisson denoise x4 (denoiseIterations = 2) -> K4 GI compose          (432 B/px algorithmic)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 : one 3840x2160 frame per step on one GPU (the configuration the metric is quoted on).
N > 1 : weak scaling — a 3840 x (2160*N) frame, row-block sharded: each rank renders its own 2160 rows (halo rows
        recomputed locally, realism_effects_b200/parallel.py) and the ranks all-gather the produced planes the next frame
        samples at arbitrary uv (composed + dnB[0..1], 32 B/px) over NCCL once per frame.
value   : whole-job Mpixels/s with the input planes resident in HBM (CUDA events on the launching stream, max over ranks).
e2e     : the same metric through host buffers (pinned host planes -> H2D -> chain -> D2H of `composed`), copies inside
          the timed region; at N = 1 this is the single C-ABI call rfx_ssgi_chain_render_host.
roofline: the dominant kernel's algorithmic bytes / its mean CUDA-event duration over the timed frames, against the
          measured HBM copy bandwidth in MEASURED_PEAKS.json (+ the chain-level figure).
cpu_baseline / --impl reference: the CPU restatement in oracle/ (the reference itself is WebGL-only and cannot run
          here: no GL, no JS engine) on the box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WIDTH, HEIGHT = 3840, 2160
DENOISE_ITERATIONS = 2
ALGO_BYTES = {  # algorithmic bytes per output pixel
    "K1_ssgi_trace": 76, "K2_temporal_reproject": 80, "K3_poisson_pass0": 68, "K3_poisson_pass1plus": 52, "K4_gi_compose": 52,
}
CHAIN_BYTES_PER_PX = 76 + 80 + (68 + 3 * 52) + 52  # = 432 (C3)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def opts_for_bench():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import chain_harness as ch  # parameter builders only (the oracle is imported lazily by the CPU legs)

    return ch, ch.Opts(denoise_iterations=DENOISE_ITERATIONS)


def make_gpu_frames(width, height, n, device, aspect):
    """Synthetic planes generated on the device with torch (plumbing)."""
    import torch

    from realism_effects_b200 import synth

    frames = []
    for t in range(n):
        fr = synth.render_frame(width, height, t + 1, device=device, aspect=aspect)
        frames.append(dict(depth=fr.depth, gbuffer=fr.gbuffer, velocity=fr.velocity, direct=fr.direct_light, cam=fr.cam.uniforms(), moved=True))
    torch.cuda.synchronize()
    return frames


def tensor_plane(t, fmt):
    from realism_effects_b200 import abi

    p = abi.Plane()
    p.ptr = t.data_ptr()
    p.height, p.width = t.shape[0], t.shape[1]
    p.pitch = t.shape[1] * abi.FMT_BYTES[fmt]
    p.format = fmt
    return p


class _PW:  # adapter so SsgiChain.render can take raw planes
    def __init__(self, p):
        self.p = p


def run_ours(args):
    import math

    import torch
    import torch.distributed as dist

    from realism_effects_b200 import abi, engine, parallel, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU oracle timing)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ch, o = opts_for_bench()
    W, Hr = args.width, args.height          # per-rank block
    emu = max(1, args.emulate_world) if world == 1 else 1   # experiment: one process plays rank 0 of `emu` (no exchange)
    H = Hr * world * emu                     # global frame height (weak scaling)
    K, Wm = args.steps, args.warmup
    # weak scaling keeps the 4K VIEW (same camera, same content mix) and samples it with N x more rows (non-square pixels), so
    # the per-pixel work statistics are those of the N = 1 frame
    aspect = W / (args.view_height or Hr)

    ctx = engine.Context(local)
    env = synth.synthetic_env(1024, 512)
    marg, cond, total = synth.build_env_cdf(env.astype(np.float32))
    ctx.set_env(env, marg, cond, total)

    class _I:  # minimal Inputs for chain_options
        width, height = W, H

    copt = ch.chain_options(_I, o)
    frames = make_gpu_frames(W, H, 2, dev, aspect)
    planes = [dict(depth=tensor_plane(f["depth"], abi.FMT_R32F), gbuffer=tensor_plane(f["gbuffer"], abi.FMT_RGBA32F),
                   velocity=tensor_plane(f["velocity"], abi.FMT_RGBA32F), direct=tensor_plane(f["direct"], abi.FMT_RGBA16F)) for f in frames]
    cams = [abi.make_camera(f["cam"]) for f in frames]

    if world == 1:
        chain = engine.SsgiChain(ctx, copt)
        native = chain
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

        force_ranges = None
        if args.force_blocks:  # experiment: issue the frame as B row blocks (with their recomputed halos) on one GPU
            force_ranges = parallel.ShardPlan(H, emu, 0, 2 * o.denoise_iterations, o.radius, True, args.force_blocks).block_ranges

        def render(i):
            j = i % len(frames)
            pl = planes[j]
            a = (cams[j], _PW(pl["depth"]), _PW(pl["gbuffer"]), _PW(pl["velocity"]), _PW(pl["direct"]), frames[j]["cam"]["position"], True)
            if args.split_parts:  # experiment: the three-part frame of the sharded path on one GPU (cost of splitting K1)
                for part in (0, 1, 2):
                    chain.render_part(part, *a, ranges=force_ranges)
            else:
                chain.render(*a, ranges=force_ranges)
    else:
        # auto: measured up to 4 GPUs, adaptive bands win (N = 2: 4.70 vs 5.14 ms, N = 4: 6.02 vs 7.04 ms against cyclic blocks).  At 8
        # GPUs the frame is exchange-bound and the grouped send/recv of unequal bands sustains only ~211 GB/s per rank with 7 peers
        # (the overlap model reproduces the measured 11.8 ms), so from 8 GPUs on the block-cyclic assignment is used:
        # equal blocks can be exchanged with NCCL's in-place all-gather (a real collective, not 7 point-to-point pairs per rank).
        # This choice is model-based - round 1 had no GPU time left to measure it; tools/next_round_sweeps.sh n8 does.
        balance = args.balance if args.balance != "auto" else ("adaptive" if world <= 4 else "static")
        chain = parallel.ShardedSsgiChain(ctx, copt, blocks_per_rank=args.blocks_per_rank, overlap=not args.no_overlap, mirror=args.mirror, balance=balance,
                                            split_k1=bool(args.split_k1), dual_comm=bool(args.dual_comm))
        native = chain.chain
        stream = chain.stream  # kernels + NCCL all-gathers are ordered on this stream

        def render(i):
            j = i % len(frames)
            pl = planes[j]
            chain.render(cams[j], _PW(pl["depth"]), _PW(pl["gbuffer"]), _PW(pl["velocity"]), _PW(pl["direct"]), frames[j]["cam"]["position"], True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------------
    calib = 24 if (world > 1 and chain.balance == "adaptive") else 0  # untimed frames in which the band borders settle (on top of --warmup)
    for i in range(calib + Wm):
        render(i)
    barrier()
    launches0 = ctx.launch_count
    native.set_profiling(True)
    native.get_profile()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        render(Wm + i)
    if world > 1:
        chain.finish()  # the last frame's all-gathers belong to the timed region
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    prof = native.get_profile()
    native.set_profiling(False)
    launches = ctx.launch_count - launches0
    ms_per_step = ms_total / K
    bounds_dev = list(chain.plan.bounds) if (world > 1 and chain.plan.bounds is not None) else None  # band borders of the timed region
    rank_kernel_ms = [{k: round(ms / K, 3) for k, (ms, n) in prof.items() if n}]  # per-frame kernel time of every rank (load balance)
    if world > 1:
        rank_kernel_ms = [None] * world
        dist.all_gather_object(rank_kernel_ms, {k: round(ms / K, 3) for k, (ms, n) in prof.items() if n})
    mpx = W * H / emu / 1e6
    value = mpx / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (this rank's owned pixels / its event-timed duration) -------
    peak, peak_src = measured_peak()
    per_kernel = {}
    own_rows = chain.plan.rows_per_rank if world > 1 else Hr  # (adaptive bands: this rank's band at the end of the run)
    for k, (ms, n) in prof.items():
        if n:
            per_kernel[k] = {"ms_per_launch": ms / n, "launches": n, "share_of_step": ms / max(ms_total, 1e-9),
                             "algo_GBps": ALGO_BYTES[k] * W * own_rows / (ms / n * 1e-3) / 1e9}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"] * per_kernel[k]["launches"]) if per_kernel else None
    roof = None
    if dom:
        ach = per_kernel[dom]["algo_GBps"]
        chain_ach = CHAIN_BYTES_PER_PX * W * Hr / (ms_per_step * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
                "peak_source": peak_src, "chain_achieved": round(chain_ach, 1), "chain_frac": round(chain_ach / peak, 4),
                "note": "per GPU; this path is instruction/SFU-bound, not HBM-bound (DESIGN.md §4): ncu DRAM traffic is at or below the algorithmic bytes",
                "per_kernel": {k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in per_kernel.items()}}
        tf = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
        if world == 1 and (W, Hr) == (3840, 2160) and os.path.exists(tf):  # measured once per round under ncu (tools/ncu_traffic.py), quoted here
            t = json.load(open(tf))
            if dom in t.get("bytes_per_launch", {}):
                roof["traffic"] = t["bytes_per_launch"][dom]
                roof["algorithmic_bytes_per_launch"] = ALGO_BYTES[dom] * W * Hr
                roof["traffic_source"] = t.get("source")

    # ---- e2e through host buffers --------------------------------------------------------------------
    # pinned host copies of the input frames (N > 1: one frame serves both parities - at N = 8 a frame is 2.9 GB per rank)
    host = [{k: f[k].cpu().pin_memory() for k in ("depth", "gbuffer", "velocity", "direct")} for f in frames[:(2 if world == 1 else 1)]]
    if world > 1:
        host.append(host[0])
    h2d = sum(host[0][k].numel() * host[0][k].element_size() for k in host[0])
    ke = max(3, min(K, 10))
    if world == 1:
        outs = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        out_host = outs[0]
        hfs = []
        for j, hb in enumerate(host):
            hf = abi.SsgiHostFrame()
            hf.cam = cams[j]
            hf.depth, hf.gbuffer, hf.velocity, hf.direct_light = hb["depth"].data_ptr(), hb["gbuffer"].data_ptr(), hb["velocity"].data_ptr(), hb["direct"].data_ptr()
            hf.camera_pos[:] = [float(x) for x in frames[j]["cam"]["position"]]
            hf.camera_moved = 1
            hf.out_composed = outs[j].data_ptr()
            hfs.append(hf)

        def e2e_step(i):
            # pipelined host path (include/rfx.h): frame i's H2D / kernels / D2H are enqueued on three streams; the call then waits
            # for frame i-1, whose host buffer set is reused by frame i+1.  Every step still moves its own 365 MB in and 133 MB out.
            chain.submit_host(hfs[i % 2])
            chain.wait_host(1)

        def e2e_drain():
            chain.wait_host(0)
    else:
        # sharded host path (parallel.ShardedSsgiChain.submit_host): each rank uploads its own rows of depth / velocity and the
        # K1-range rows of gbuffer / direct light, the two sampled-anywhere planes are all-gathered over NVLink, and the rank
        # reads back its own rows of `composed`; two frames in flight
        outs = [torch.empty((min(H, int(chain.MAX_SHARE * Hr) + 16), W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]  # tallest adaptive band
        out_host = outs[0]

        def e2e_step(i):
            j = i % 2
            chain.submit_host(cams[j], host[j], frames[j]["cam"]["position"], True, outs[j])
            chain.wait_host(1)

        def e2e_drain():
            chain.wait_host(0)
            chain.finish()
    d2h = out_host.numel() * 4
    input_planes_bytes = h2d  # size of one frame's input planes (the working set the L2 note in `config` refers to)
    for i in range(3 + calib):  # (adaptive bands re-settle for the host path: upload time grows with the band as well)
        e2e_step(i)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for i in range(ke):
        e2e_step(i)
    e2e_drain()  # the last frame's result has landed in host memory before the clock stops
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / ke)
    checksum = float(out_host[::97, ::89, :3].double().sum())
    if world > 1:
        h2d, d2h = chain.host_bytes_per_frame  # this rank's share (own rows + K1-range rows), after the bands settled
    e2e = {"value": round(mpx / e2e_s, 2), "unit": "Mpixels/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": round(e2e_s * 1e3, 3), "steps": ke, "result_checksum": checksum, "bytes_are": "per rank"}
    if world == 1:  # latency of one frame through the synchronous call (no overlap between frames)
        t0 = time.perf_counter()
        for i in range(3):
            chain.render_host(hfs[i % 2])
        e2e["sync_call_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
        e2e["mode"] = "pipelined submit_host/wait_host, 2 frames in flight; sync_call_ms = rfx_ssgi_chain_render_host latency"

    # ---- CPU baseline (rank 0, N=1 only): one full-resolution frame through the oracle ------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if (args.cpu_width, args.cpu_height) == (W, H):  # reuse the planes already generated on the device
            f0 = frames[0]
            fr = dict(depth=f0["depth"].cpu().numpy(), gbuffer=f0["gbuffer"].cpu().numpy(), velocity=f0["velocity"].cpu().numpy(),
                      direct=f0["direct"].cpu().numpy(), cam=f0["cam"], moved=True)
            cpu_inp = ch.Inputs(W, H, [fr], env, marg, cond, total, synth.load_blue_noise())
        else:
            cpu_inp = ch.make_inputs(args.cpu_width, args.cpu_height, 1, env_size=(1024, 512))
        cpu = cpu_baseline_sample(ch, o, cpu_inp)

    if rank == 0:
        cfg = {"workload": f"C3 SSGI+PoissonDenoise(denoiseIterations={DENOISE_ITERATIONS} => {2 * DENOISE_ITERATIONS} passes)+compose, steps=20 refineSteps=5, "
                           f"{W}x{Hr} per GPU" + (f" (frame {W}x{H} = the 4K view sampled with {world}x the rows, row-sharded over {world} GPUs)" if world > 1 else ""),
               "inputs": f"2 alternating synthetic G-buffer frames ({input_planes_bytes / 1e6:.0f} MB of input planes per frame > 126 MB L2), moving camera",
               "l2": "inputs larger than L2; no explicit flush", "fast_math": True}
        if world > 1:
            if chain.balance == "adaptive":
                shard = (f"one contiguous band per rank, borders rebalanced every {chain.rebalance_every} frames from the ranks' event-timed kernel time; "
                         f"borders during the timed frames {bounds_dev}, halo rows recomputed locally")
            else:
                shard = ("mirrored (boustrophedon) " if chain.plan.mirror else "block-cyclic ") + \
                    f"row blocks ({chain.plan.blocks_per_rank} x {chain.plan.block_rows} rows per rank), halo rows recomputed locally"
            cfg["multi_gpu"] = {"sharding": shard,
                                "recompute_overhead": round(chain.plan.recompute_overhead, 4), "calibration_frames": calib,
                                "exchange": ("NCCL grouped send/recv" if chain.plan.p2p else "NCCL all-gather") + " of composed + dnB[0..1] once per frame" + ("" if args.no_overlap else ("; the composed exchange overlaps the next frame's K1 ray march, the dnB exchange its K1 shading" if args.split_k1 else "; dnB exchange overlaps the next frame's K1")),
                                "exchange_recv_bytes_per_rank_per_frame": chain.exchange_bytes_per_frame,
                                "per_rank_kernel_ms_per_frame": [round(sum(d.values()), 3) for d in rank_kernel_ms],
                                "per_rank_K1_ms": [d.get("K1_ssgi_trace") for d in rank_kernel_ms]}
        line = {
            "metric": "SSGI+denoise Mpixels/s at 4K", "value": round(value, 2), "unit": "Mpixels/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 accumulate planes)",
            "data": "synthetic", "impl": "ours", "config": cfg, "gpu_launches": int(launches), "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
        }
        print(json.dumps(line))
    chain.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_baseline_sample(ch, o, inp):
    """Times the CPU oracle chain on the frames of `inp` (all host threads)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc

    width, height, frames_n = inp.width, inp.height, len(inp.frames)
    orc.lib()
    t0 = time.perf_counter()
    ch.run_oracle_chain(inp, o, capture=("composed",))
    dt = time.perf_counter() - t0
    cores = int(orc.lib().orc_num_threads())
    return {"value": round(width * height * frames_n / 1e6 / dt, 4), "unit": "Mpixels/s", "cores": cores, "kind": "port",
            "sample": f"{frames_n} frame(s) of the same chain at {width}x{height} ({dt:.1f} s of CPU work; first frame => empty history)"}


def run_reference(args):
    """--impl reference: the reference's own implementation is WebGL-only (no GL / JS engine here), so the
    arm times the CPU restatement in oracle/ with all host threads, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ch, o = opts_for_bench()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc

    sw, sh = args.cpu_width, args.cpu_height
    if args.cpu_width == WIDTH:  # default: 1/16 of the 4K frame per step
        sw, sh = 960, 540
    K, Wm = args.steps, args.warmup
    inp = ch.make_inputs(sw, sh, 2, env_size=(1024, 512))
    frames = inp.frames
    cores = int(orc.lib().orc_num_threads())

    def step_block(n):
        inp.frames = [frames[i % 2] for i in range(n)]
        ch.run_oracle_chain(inp, o, capture=("composed",))

    step_block(Wm)
    t0 = time.perf_counter()
    step_block(K)
    dt = (time.perf_counter() - t0) / K
    v = round(sw * sh / 1e6 / dt, 4)
    line = {"metric": "SSGI+denoise Mpixels/s at 4K", "value": v, "unit": "Mpixels/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K,
            "warmup": Wm, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": f"C3 SSGI+PoissonDenoise(denoiseIterations={DENOISE_ITERATIONS})+compose, CPU restatement (oracle/), bounded sample {sw}x{sh} per step"},
            "cpu_baseline": {"value": v, "unit": "Mpixels/s", "cores": cores, "kind": "port", "sample": f"{K} steps x one {sw}x{sh} frame (1/16 of the 4K frame)"},
            "e2e": {"value": v, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT, help="rows per GPU")
    ap.add_argument("--cpu-width", type=int, default=WIDTH)
    ap.add_argument("--cpu-height", type=int, default=HEIGHT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emulate-world", type=int, default=1, help="experiment (N = 1, with --force-blocks): play rank 0 of this many ranks")
    ap.add_argument("--force-blocks", type=int, default=0, help="experiment (N = 1): issue each pass as this many row-block launches")
    ap.add_argument("--view-height", type=int, default=0, help="experiment: rows of the VIEW (aspect = width / view_height) when --height differs")
    ap.add_argument("--blocks-per-rank", type=int, default=4, help="N > 1: block-cyclic row blocks per rank (content balance)")
    ap.add_argument("--balance", default="auto", choices=("auto", "adaptive", "static"),
                    help="N > 1: adaptive = one band per rank, borders follow the measured kernel time (grouped send/recv exchange); static = "
                         "block-cyclic / mirrored blocks (in-place all-gather); auto = adaptive up to 4 GPUs, static from 8 (see run_ours)")
    ap.add_argument("--split-parts", action="store_true", help="experiment (N = 1): issue every frame as K1 march / K1 shading / K2..K4")
    ap.add_argument("--dual-comm", type=int, default=0, help="experiment (N > 1): 1 = dnB exchange on a second NCCL communicator, concurrent with composed")
    ap.add_argument("--split-k1", type=int, default=1, help="N > 1: 1 = K1 as ray march + shading so the `composed` exchange hides behind the march")
    ap.add_argument("--mirror", type=int, default=0, help="N > 1: 1 = boustrophedon block assignment (odd super-blocks in reverse rank order), P2P exchange")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: wait for all all-gathers at the end of every frame")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
