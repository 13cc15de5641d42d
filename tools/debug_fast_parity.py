"""GPU debug: per-pass isolated parity of the fast variants on a dark scene (no env, no direct light)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import chain_harness as ch
from realism_effects_b200 import abi, engine

kw = dict(use_envmap=False, use_direct_light=False, steps=8, refine_steps=0)
o = ch.Opts(**kw)
inp = ch.make_inputs(128, 72, 2)
ref = ch.run_oracle_chain(inp, o)
for fast in (True, False):
    ctx = engine.Context(0, inp.blue); ctx.set_fast_math(fast)
    fr, rec = inp.frames[1], ref[1]
    d, g = ctx.upload(fr["depth"]), ctx.upload(fr["gbuffer"])
    o0, o1 = ctx.upload(rec["_k2_prev_out"][0]), ctx.upload(rec["_k2_prev_out"][1])
    ctx.temporal_reproject(rec["_k2_params"], ctx.upload(rec["ssgi"]), ctx.upload(fr["velocity"]), ctx.upload(rec["_k2_hist"][0]), ctx.upload(rec["_k2_hist"][1]), o0, o1)
    for nm, w, gt in (("K2 d", rec["tr0"], o0.download()), ("K2 s", rec["tr1"], o1.download())):
        print(fast, nm, ch.compare(w, gt))
    for i, k3 in enumerate(rec["_k3"]):
        p0, p1 = ctx.upload(k3["prev0"]), ctx.upload(k3["prev1"])
        ctx.poisson_denoise(k3["params"], d, g, ctx.upload(k3["in0"]), ctx.upload(k3["in1"]), p0, p1)
        for nm, w, gt in ((f"K3.{i} d", k3["out0"], p0.download()), (f"K3.{i} s", k3["out1"], p1.download())):
            c = ch.compare(w, gt)
            print(fast, nm, c)
            if c["n_bad"]:
                a, b = w.astype(np.float64), gt.astype(np.float64)
                bad = np.abs(a - b) > ch.RTOL * np.maximum(np.abs(a), np.abs(b)) + ch.ATOL
                ys, xs, cs = np.where(bad)
                for y, x, c_ in list(zip(ys, xs, cs))[:8]:
                    print("    px", y, x, c_, "want", a[y, x], "got", b[y, x], "in0", k3["in0"][y, x], "depth", fr["depth"][y, x])
    ctx.close()
