"""Mint the golden fixtures from the CPU oracle (run in the build container: python tests/golden/make_golden.py).

The reference has no golden vectors of its own, so these pin *self*-consistency of the
restatement: a later change of the oracle or of the kernels that alters results shows up against them.
Inputs are stored with the outputs so the fixtures do not depend on the synthetic generator's torch version.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import chain_harness as ch  # noqa: E402
import orc  # noqa: E402
from realism_effects_b200 import abi  # noqa: E402


def main():
    o = ch.Opts(steps=12, refine_steps=3)
    inp = ch.make_inputs(96, 54, 2)
    ref = ch.run_oracle_chain(inp, o)
    d = dict(env_map=inp.env_map, env_marginal=inp.env_marginal, env_conditional=inp.env_conditional, env_total=np.float64(inp.env_total))
    for t, fr in enumerate(inp.frames):
        for k in ("depth", "gbuffer", "velocity", "direct"):
            d[f"f{t}_{k}"] = fr[k]
        for k, v in fr["cam"].items():
            d[f"f{t}_cam_{k}"] = np.asarray(v)
        d[f"f{t}_moved"] = np.int32(fr["moved"])
        for k in ("ssgi", "tr0", "tr1", "dn0", "dn1", "composed"):
            d[f"f{t}_out_{k}"] = ref[t][k]
    # HBAO + AO compose + motion blur + TRAA compose on frame 1's planes
    f1 = inp.frames[1]
    H, W = f1["depth"].shape
    ao = orc.hbao(ch.hbao_params(f1["cam"], 4242), f1["depth"], inp.blue, np.zeros((H, W, 4), np.float16))
    d["hbao"] = ao
    d["ao_composed"] = orc.ao_compose(ch.ao_compose_params(), f1["depth"], ao, f1["direct"])
    vel = ch.rotation_velocity_field(W, H, f1["depth"])
    d["mb_velocity"] = vel
    d["motion_blur"] = orc.motion_blur(ch.motion_blur_params(W, H), vel, f1["direct"], inp.blue)
    d["traa_compose"] = orc.traa_compose(f1["direct"])
    np.savez_compressed(os.path.join(HERE, "chain_96x54.npz"), **d)
    print("wrote", os.path.join(HERE, "chain_96x54.npz"), os.path.getsize(os.path.join(HERE, "chain_96x54.npz")), "bytes")


if __name__ == "__main__":
    main()
