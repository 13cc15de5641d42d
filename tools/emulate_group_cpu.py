"""CPU emulation of one rank of a row-sharded group at the bench workload (3840x2160, 4 bands, the bench's two frames, HDR environment),
with the ORACLE as compute (tests/test_sharding_cpu.py: every pass commits only its planned rows; the other ranks' rows of the exchanged
planes are taken from the single-process chain).  It reproduces, pixel for pixel, what the N = 4 GPU run showed before the A-target fix —
frame 1, rank 2: composed 2 px, dn0 1 px, dn1 1 px, all on row 1615 (profiles/r02_group_diag_n4_before_fix.log) — and shows that taking
the A Poisson target's rows from the rank that owns them (what csrc/rfx_api.cu now does by carrying discarded texels from the owner) makes the
band bit-identical.  ~4 min on 8 cores.   python tools/emulate_group_cpu.py [--world N] [rank ...]      TEST INFRASTRUCTURE (uses oracle/)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import chain_harness as ch  # noqa: E402
from realism_effects_b200 import synth  # noqa: E402
from test_sharding_cpu import sharded_oracle_chain  # noqa: E402

W, H = 3840, 2160
args = sys.argv[1:]
world = int(args.pop(args.index("--world") + 1)) if "--world" in args else 4
if "--world" in args:
    args.remove("--world")
bounds = tuple(int(round(H * i / world / 16.0)) * 16 for i in range(world)) + (H,)  # rfx_group_attach_chain's initial borders (4 ranks: 0, 544, 1088, 1616, 2160)
ranks = [int(a) for a in args] or [2]
o = ch.Opts(denoise_iterations=2)
frames = []
for t in (1, 2):  # bench.make_gpu_frames
    fr = synth.render_frame(W, H, t)
    frames.append(dict(depth=fr.depth.numpy(), gbuffer=fr.gbuffer.numpy(), velocity=fr.velocity.numpy(), direct=fr.direct_light.numpy(), cam=fr.cam.uniforms(), moved=True))
img, gl = synth.load_reference_env()
marg, cond, total = synth.build_env_cdf(img.astype(np.float32), flip_y=True)
inp = ch.Inputs(W, H, frames, gl, marg, cond, total, synth.load_blue_noise())
t0 = time.time()
recs = ch.run_oracle_chain(inp, o)
print(f"single-process chain, 2 frames: {time.time() - t0:.0f} s", flush=True)
ref = [{k: r[k] for k in ("composed", "dn0", "dn1")} for r in recs]
a_ref = [(r["_k3"][2]["out0"], r["_k3"][2]["out1"]) for r in recs]  # the A target after its last pass of each frame
del recs
for rank, exchange_a in [(r, e) for r in ranks for e in (False, True)]:
    st = {"k": 0, "f": 0}

    def all_gather_rows(plane, plan, st=st, exchange_a=exchange_a, rank=rank):
        name = ("composed", "dn0", "dn1", "a0", "a1")[st["k"]]
        src = a_ref[st["f"]][int(name[1])] if name[0] == "a" else ref[st["f"]][name]
        for g in range(world):
            if g != rank:
                plane[bounds[g]:bounds[g + 1]] = src[bounds[g]:bounds[g + 1]]
        st["k"] += 1
        if st["k"] == (5 if exchange_a else 3):
            st["k"], st["f"] = 0, st["f"] + 1

    t0 = time.time()
    out = sharded_oracle_chain(inp, o, rank, world, all_gather_rows, bounds_per_frame=[bounds, bounds], exchange_a=exchange_a)
    b0, b1 = bounds[rank], bounds[rank + 1]
    msg = []
    for k in ("composed", "dn0", "dn1"):
        a, b = out[k][b0:b1], ref[-1][k][b0:b1]
        px = (a.reshape(b1 - b0, W, -1).view(np.uint8) != b.reshape(b1 - b0, W, -1).view(np.uint8)).any(-1)
        rows = np.nonzero(px.any(1))[0]
        msg.append(f"{k}: " + ("equal" if len(rows) == 0 else f"{int(px.sum())} px differ, rows {b0 + rows[0]}..{b0 + rows[-1]}, cols {np.nonzero(px.any(0))[0].tolist()}"))
    print(("A rows from their owner (the fix)" if exchange_a else "A single-buffered and local (before the fix)") + f" | rank {rank} band [{b0},{b1}) frame 1 | " + " | ".join(msg)
          + f" ({time.time() - t0:.0f} s)", flush=True)
