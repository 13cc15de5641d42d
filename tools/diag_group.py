"""torchrun diagnostic: row-sharded group vs single-GPU render on every rank, per frame and per output plane.
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/diag_group.py [W H frames]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import chain_harness as ch  # noqa: E402
from realism_effects_b200 import abi, engine, parallel, synth  # noqa: E402

W, H, F = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160, 4)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = engine.Context(local, synth.load_blue_noise())
env, marg, cond, total = bench.env_for_bench() if os.environ.get("DIAG_ENV", "ref") == "ref" else (synth.synthetic_env(1024, 512),) + synth.build_env_cdf(synth.synthetic_env(1024, 512).astype(np.float32))
ctx.set_env(env, marg, cond, total)
o = ch.Opts(denoise_iterations=2)
copt = bench.chain_options(ch, o, W, H)
frames = bench.make_gpu_frames(W, H, 2, dev)
planes = [bench.frame_planes(f) for f in frames]
cams = [abi.make_camera(f["cam"]) for f in frames]
single = engine.SsgiChain(ctx, copt)
sh = parallel.ShardedSsgiChain(ctx, copt, rebalance_every=int(os.environ.get("DIAG_REBALANCE", "0")), rebalance_lag=2)
names = {0: "composed", 1: "ssgi", 2: "tr0", 3: "tr1", 4: "dn0", 5: "dn1"}
for t in range(F):
    j = t % 2
    single.render(cams[j], *planes[j], frames[j]["cam"]["position"], True)
    sh.render(cams[j], *planes[j], frames[j]["cam"]["position"], True)
    b0, b1 = sh.band_of_last_frame
    msg = [f"rank {rank} frame {t} band [{b0},{b1}) peer={sh.uses_peer_reads}"]
    for wch, nm in names.items():
        a, b = single.download(wch)[b0:b1], sh.chain.download(wch)[b0:b1]
        ne = (a.view(np.uint8).reshape(a.shape[0], -1) != b.view(np.uint8).reshape(b.shape[0], -1))
        rows = np.nonzero(ne.any(1))[0]
        if len(rows):
            px = (a.reshape(a.shape[0], a.shape[1], -1).view(np.uint8) != b.reshape(b.shape[0], b.shape[1], -1).view(np.uint8)).any(-1)
            cols = np.nonzero(px.any(0))[0]
            msg.append(f"{nm}: {int(px.sum())} px differ, rows {b0 + rows[0]}..{b0 + rows[-1]} ({len(rows)} rows), cols {cols[0]}..{cols[-1]}")
        else:
            msg.append(f"{nm}: equal")
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(" | ".join(msg), flush=True)
sh.close()
single.close()
ctx.close()
dist.destroy_process_group()
