"""Mint the golden fixtures from THE REFERENCE'S OWN SHADERS, compiled for the CPU (tests/refglsl.py: the GLSL under
/root/reference/src, assembled like the reference's JS assembles it, run on oracle/ref/glsl_rt.h).

    python tests/golden/make_golden.py          # needs the reference checkout (/root/reference); run in the build container

The reference ships no golden vectors of its own; these are outputs of the reference itself run here.
tests/test_oracle_chain_cpu.py checks the C++ oracle against them on any machine (no checkout needed), the `-m gpu` tests check
the CUDA engine against them.  Inputs are stored with the outputs so the fixtures do not depend on the synthetic generator.
  chain_96x54.npz      SSGI chain (K1 -> K2 -> K3 x2 -> K4), 2 frames, steps 12 / refine 3  + K5..K9 and the AO denoise on frame 1
  chain_ssr_64x36.npz  SSR chain (mode "ssr": 1-plane K2/K3, TYPE_SPECULAR compose), 3 frames
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import chain_harness as ch  # noqa: E402
import refglsl as ref  # noqa: E402  (the reference shaders; NOT the oracle)
from realism_effects_b200 import abi  # noqa: E402


def store_inputs(d, inp):
    d.update(env_map=inp.env_map, env_marginal=inp.env_marginal, env_conditional=inp.env_conditional, env_total=np.float64(inp.env_total))
    for t, fr in enumerate(inp.frames):
        for k in ("depth", "gbuffer", "velocity", "direct"):
            d[f"f{t}_{k}"] = fr[k]
        for k, v in fr["cam"].items():
            d[f"f{t}_cam_{k}"] = np.asarray(v)
        d[f"f{t}_moved"] = np.int32(fr["moved"])


def fog_params(cam_u, exp2):
    p = abi.SsgiComposeParams()
    p.use_fog, p.fog_exp2, p.perspective = 1, int(exp2), 1
    p.fog_color[:] = [0.6, 0.7, 0.8]
    p.fog_near, p.fog_far, p.fog_density = 2.0, 30.0, 0.05
    p.camera_near, p.camera_far = float(cam_u["near"]), float(cam_u["far"])
    return p


def ao_denoise(m, f1, blue, ao):
    """AOEffect's denoiser: 2 Poisson passes over one plane with the velocity-layout normals (non-GBUFFER branch)"""
    H, W = f1["depth"].shape
    tgtA, tgtB = np.zeros((H, W, 4), np.float16), np.zeros((H, W, 4), np.float16)
    for i in range(2):
        p = ch.poisson_params(ch.Opts(), 1234568 + i, False)
        p.texture_count, p.gbuffer_texture, p.input_linear = 1, 0, 1
        p.is_texture_specular[:] = [0, 0]
        p.normal_phi, p.depth_phi, p.roughness_phi, p.specular_phi = 3.25, 2.0, 0.0, 0.0
        out, _ = m.poisson_denoise(p, f1["depth"], f1["velocity"], ao if i == 0 else tgtA, None, blue, tgtA if i == 0 else tgtB, None)
        if i == 0:
            tgtA = out
        else:
            tgtB = out
    return tgtA, tgtB


def traa_two_frames(m, f0, f1):
    z = np.zeros(f0["direct"].shape, np.float16)
    p0 = ch.traa_temporal_params(abi.make_camera(f0["cam"]), f0["cam"]["position"], f0["cam"], 0.0)
    h0, _ = m.temporal_reproject(p0, f0["direct"], f0["velocity"], z, None, z, None, out_half=True)
    p1 = ch.traa_temporal_params(abi.make_camera(f1["cam"]), f1["cam"]["position"], f0["cam"], 1.0)
    h1, _ = m.temporal_reproject(p1, f1["direct"], f1["velocity"], h0, None, h0, None, out_half=True)
    return h0, h1


def main():
    assert ref.assemble.available(), "the reference checkout is needed"
    o = ch.Opts(steps=12, refine_steps=3)
    inp = ch.make_inputs(96, 54, 2)
    out = ch.run_oracle_chain(inp, o, impl=ref)
    d = {}
    store_inputs(d, inp)
    for t in range(2):
        for k in ("ssgi", "tr0", "tr1", "dn0", "dn1", "composed"):
            d[f"f{t}_out_{k}"] = out[t][k]
    f0, f1 = inp.frames
    H, W = f1["depth"].shape
    ao = ref.hbao(ch.hbao_params(f1["cam"], 4242), f1["depth"], inp.blue, np.zeros((H, W, 4), np.float16))
    d["hbao"] = ao
    d["ao_composed"] = ref.ao_compose(ch.ao_compose_params(), f1["depth"], ao, f1["direct"])
    vel = ch.rotation_velocity_field(W, H, f1["depth"])
    d["mb_velocity"] = vel
    d["motion_blur"] = ref.motion_blur(ch.motion_blur_params(W, H), vel, f1["direct"], inp.blue)
    d["traa_compose"] = ref.traa_compose(f1["direct"])
    d["ao_dn_a"], d["ao_dn_b"] = ao_denoise(ref, f1, inp.blue, ao)
    d["traa_h0"], d["traa_h1"] = traa_two_frames(ref, f0, f1)
    d["k5_plain"] = ref.ssgi_compose(f1["depth"], out[1]["composed"], f1["direct"])
    d["k5_fog"] = ref.ssgi_compose(f1["depth"], out[1]["composed"], f1["direct"], fog_params(f1["cam"], False))
    d["k5_fog_exp2"] = ref.ssgi_compose(f1["depth"], out[1]["composed"], f1["direct"], fog_params(f1["cam"], True))
    path = os.path.join(HERE, "chain_96x54.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes")

    o = ch.Opts(mode=abi.MODE_SSR)
    inp = ch.make_inputs(64, 36, 3)
    out = ch.run_oracle_chain(inp, o, impl=ref)
    d = {}
    store_inputs(d, inp)
    for t in range(3):
        for k in ("ssgi", "tr0", "dn0", "composed"):
            d[f"f{t}_out_{k}"] = out[t][k]
    path = os.path.join(HERE, "chain_ssr_64x36.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
