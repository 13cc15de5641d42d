"""GPU debug: isolated K1 parity of the fast variant against the oracle, with the differing pixels classified.

Unpacks the (diffuse rgb, roughness | specular rgb, rayLength) halves and reports, for pixels outside the 1e-3 band, whether the
ray length changed (a different hit / miss decision or hit texel) or only the colour did.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import chain_harness as ch
from realism_effects_b200 import engine, abi

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 144)
o = ch.Opts()
inp = ch.make_inputs(W, H, 3)
ref = ch.run_oracle_chain(inp, o)
ctx = engine.Context(0, inp.blue)
ctx.set_env(inp.env_map, inp.env_marginal, inp.env_conditional, inp.env_total)
for fast in (False, True):
    ctx.set_fast_math(fast)
    for t in (0, 2):
        fr, rec = inp.frames[t], ref[t]
        out = ctx.alloc(abi.FMT_RGBA32F, W, H)
        ctx.ssgi_trace(rec["_k1_params"], ctx.upload(fr["depth"]), ctx.upload(fr["gbuffer"]), None, ctx.upload(fr["direct"]),
                       ctx.upload(rec["_k1_accumulated"]), out)
        got = ch.unpack_halves(out.download()).astype(np.float64)
        want = ch.unpack_halves(rec["ssgi"]).astype(np.float64)
        bad = np.abs(got - want) > 1e-3 * np.maximum(np.abs(got), np.abs(want)) + 1e-5
        geo = fr["depth"] < 1.0
        px_bad = bad.any(axis=-1) & geo
        names = ["d.r", "d.g", "d.b", "rough", "s.r", "s.g", "s.b", "rayLen"]
        per = {n: float((bad[..., i] & geo).mean()) for i, n in enumerate(names)}
        ray_changed = bad[..., 7] & geo
        print(f"fast={fast} f{t}: bad pixels {px_bad.mean():.4e}; rayLength changed {ray_changed.mean():.4e}; "
              f"colour-only {(px_bad & ~ray_changed).mean():.4e}")
        print("   per channel:", {k: f"{v:.2e}" for k, v in per.items()})
        ys, xs = np.where(px_bad)
        for y, x in list(zip(ys, xs))[:6]:
            print("   px", y, x, "want", np.round(want[y, x], 5), "got", np.round(got[y, x], 5))
ctx.close()
