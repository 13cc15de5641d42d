"""Multi-GPU test (needs >= 2 GPUs on the box; skipped otherwise): the row-sharded chain over NCCL must equal the
single-GPU chain bit for bit on every rank (gathered planes: whole frame; other planes: the rank's own rows) — for
contiguous bands and for block-cyclic blocks with the overlapped two-phase frame."""
import os
import socket

import numpy as np
import pytest
import torch

import chain_harness as ch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, bpr, overlap, mirror=False, balance="static"):
    import torch.distributed as dist

    from realism_effects_b200 import abi, engine, parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        o = ch.Opts(denoise_iterations=2)
        inp = ch.make_inputs(256, 256 if balance == "adaptive" else 128, 3)
        ctx = engine.Context(rank, inp.blue)
        ctx.set_env(inp.env_map, inp.env_marginal, inp.env_conditional, inp.env_total)
        chain = parallel.ShardedSsgiChain(ctx, ch.chain_options(inp, o), blocks_per_rank=bpr, overlap=overlap, mirror=mirror, balance=balance, rebalance_every=1, rebalance_lag=1)
        keep = []
        for fr in inp.frames:
            pl = [ctx.upload(fr[k]) for k in ("depth", "gbuffer", "velocity", "direct")]
            keep.append(pl)  # inputs must outlive the asynchronously enqueued frame
            chain.render(abi.make_camera(fr["cam"]), *pl, fr["cam"]["position"], fr["moved"])
        chain.finish()
        out = {k: chain.chain.download(w).tobytes() for k, w in (("composed", 0), ("ssgi", 1), ("tr0", 2), ("dn0", 4), ("dn1", 5))}
        q.put((rank, out, chain.plan.blocks))
        chain.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("bpr,overlap,mirror,balance", [(1, False, False, "static"), (2, True, False, "static"), (2, True, True, "static"),
                                                        (1, True, False, "adaptive")])
def test_sharded_chain_equals_single_gpu_bit_exact(built, bpr, overlap, mirror, balance):
    import torch.multiprocessing as mp

    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q, bpr, overlap, mirror, balance)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict()
    for _ in procs:
        rank, out, blocks = q.get(timeout=600)
        res[rank] = (out, blocks)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = ch.Opts(denoise_iterations=2)
    inp = ch.make_inputs(256, 256 if balance == "adaptive" else 128, 3)
    single, _ = ch.run_cuda_chain(inp, o)
    ref = single[-1]
    for rank, (out, blocks) in res.items():
        for k in ("composed", "dn0", "dn1"):
            assert out[k] == ref[k].tobytes(), (rank, k)
        for k in ("ssgi", "tr0"):
            got = np.frombuffer(out[k], ref[k].dtype).reshape(ref[k].shape)
            for r0, r1 in blocks:
                assert got[r0:r1].tobytes() == ref[k][r0:r1].tobytes(), (rank, k)


def _host_worker(rank, world, port, q, mirror):
    import torch.distributed as dist

    from realism_effects_b200 import abi, engine, parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        o = ch.Opts(denoise_iterations=1)
        inp = ch.make_inputs(256, 128, 4)
        ctx = engine.Context(rank, inp.blue)
        ctx.set_env(inp.env_map, inp.env_marginal, inp.env_conditional, inp.env_total)
        chain = parallel.ShardedSsgiChain(ctx, ch.chain_options(inp, o), blocks_per_rank=2, overlap=True, mirror=mirror)
        rows = chain.plan.rows_per_rank
        outs = [torch.zeros((rows, inp.width, 4), dtype=torch.float32).pin_memory() for _ in inp.frames]
        hosts = [{k: torch.from_numpy(np.ascontiguousarray(fr[k])).pin_memory() for k in ("depth", "gbuffer", "velocity", "direct")} for fr in inp.frames]
        for i, fr in enumerate(inp.frames):
            chain.submit_host(abi.make_camera(fr["cam"]), hosts[i], fr["cam"]["position"], fr["moved"], outs[i])
            chain.wait_host(1)
        chain.wait_host(0)
        chain.finish()
        q.put((rank, [o_.numpy().tobytes() for o_ in outs], chain.plan.blocks, chain.host_bytes_per_frame))
        chain.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
@pytest.mark.parametrize("mirror", [False, True])
def test_sharded_host_path_equals_single_gpu_bit_exact(built, mirror):
    """submit_host / wait_host on 2 ranks: each rank uploads its share, depth + velocity are all-gathered, and the rows of
    `composed` it reads back are, for every frame, the single-GPU chain's rows."""
    import torch.multiprocessing as mp

    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_host_worker, args=(r, world, port, q, mirror)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict()
    for _ in procs:
        rank, outs, blocks, nbytes = q.get(timeout=600)
        res[rank] = (outs, blocks, nbytes)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = ch.Opts(denoise_iterations=1)
    inp = ch.make_inputs(256, 128, 4)
    single, _ = ch.run_cuda_chain(inp, o, capture=("composed",))
    full_h2d = inp.width * inp.height * 44
    for rank, (outs, blocks, nbytes) in res.items():
        assert nbytes[0] < full_h2d and nbytes[1] == inp.width * inp.height * 16 // world
        for i, got in enumerate(outs):
            want = np.concatenate([single[i]["composed"][r0:r1] for r0, r1 in blocks], axis=0)
            assert got == want.tobytes(), (rank, i)
