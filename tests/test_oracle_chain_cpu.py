"""CPU tests: the oracle against the committed golden fixtures + behavioural properties the GLSL implies.

The fixtures under tests/golden/ are OUTPUTS OF THE REFERENCE'S OWN SHADERS (compiled for the CPU from /root/reference by
tests/golden/make_golden.py -> tests/refglsl.py); the oracle has to reproduce them bit for bit.  No reference checkout is
needed to run these tests."""
import os

import numpy as np
import pytest

import chain_harness as ch
import orc
from realism_effects_b200 import abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_96x54.npz")
GOLD_SSR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_ssr_64x36.npz")


def load_golden(path=GOLD, n_frames=2, size=(96, 54)):
    g = np.load(path)
    frames = []
    for t in range(n_frames):
        cam = {k[len(f"f{t}_cam_"):]: g[k] for k in g.files if k.startswith(f"f{t}_cam_")}
        frames.append(dict(depth=g[f"f{t}_depth"], gbuffer=g[f"f{t}_gbuffer"], velocity=g[f"f{t}_velocity"], direct=g[f"f{t}_direct"], cam=cam,
                           moved=bool(g[f"f{t}_moved"])))
    from realism_effects_b200 import synth

    inp = ch.Inputs(size[0], size[1], frames, g["env_map"], g["env_marginal"], g["env_conditional"], float(g["env_total"]), synth.load_blue_noise())
    return g, inp


def bits_equal(a, b) -> bool:
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def test_oracle_reproduces_golden_chain():
    """SSGI chain, 2 frames: every plane of the oracle equals the reference shaders' output bit for bit"""
    g, inp = load_golden()
    ref = ch.run_oracle_chain(inp, ch.Opts(steps=12, refine_steps=3))
    for t in range(2):
        for k in ("ssgi", "tr0", "tr1", "dn0", "dn1", "composed"):
            assert bits_equal(g[f"f{t}_out_{k}"], ref[t][k]), (t, k, ch.compare(g[f"f{t}_out_{k}"], ref[t][k], packed=(k == "ssgi")))


def test_oracle_reproduces_golden_ssr_chain():
    """mode "ssr" (1-plane K2 / K3, TYPE_SPECULAR compose), 3 frames with history"""
    g, inp = load_golden(GOLD_SSR, 3, (64, 36))
    ref = ch.run_oracle_chain(inp, ch.Opts(mode=abi.MODE_SSR))
    for t in range(3):
        for k in ("ssgi", "tr0", "dn0", "composed"):
            assert bits_equal(g[f"f{t}_out_{k}"], ref[t][k]), (t, k)


def test_oracle_reproduces_golden_ao_denoise_traa_and_fog():
    """the AO denoiser (non-GBUFFER Poisson branch; reference text + the one documented repair, oracle/ref/assemble.py), the TRAA form
    of K2 over two frames, and K5 with three.js Fog / FogExp2"""
    g, inp = load_golden()
    f0, f1 = inp.frames
    a, b = ch.ao_denoise(orc, f1, inp.blue, g["hbao"])
    assert bits_equal(a, g["ao_dn_a"]) and bits_equal(b, g["ao_dn_b"])
    h0, h1 = ch.traa_two_frames(orc, f0, f1)
    assert bits_equal(h0, g["traa_h0"]) and bits_equal(h1, g["traa_h1"])
    comp = g["f1_out_composed"]
    assert bits_equal(orc.ssgi_compose(f1["depth"], comp, f1["direct"]), g["k5_plain"])
    assert bits_equal(orc.ssgi_compose(f1["depth"], comp, f1["direct"], ch.fog_params(f1["cam"], False)), g["k5_fog"])
    assert bits_equal(orc.ssgi_compose(f1["depth"], comp, f1["direct"], ch.fog_params(f1["cam"], True)), g["k5_fog_exp2"])


def test_oracle_reproduces_golden_post_passes():
    g, inp = load_golden()
    f1 = inp.frames[1]
    H, W = f1["depth"].shape
    ao = orc.hbao(ch.hbao_params(f1["cam"], 4242), f1["depth"], inp.blue, np.zeros((H, W, 4), np.float16))
    assert ch.compare(g["hbao"], ao)["n_bad"] == 0
    assert ch.compare(g["ao_composed"], orc.ao_compose(ch.ao_compose_params(), f1["depth"], ao, f1["direct"]))["n_bad"] == 0
    mb = orc.motion_blur(ch.motion_blur_params(W, H), g["mb_velocity"], f1["direct"], inp.blue)
    assert ch.compare(g["motion_blur"], mb)["n_bad"] == 0
    assert ch.compare(g["traa_compose"], orc.traa_compose(f1["direct"]))["n_bad"] == 0


def test_discard_keeps_stale_texels_and_background_rules():
    """K2/K3/K4 `discard` pure-background pixels (depth == 1 and a flat 2x2 quad) and leave the target untouched;
    a background pixel whose quad touches geometry is processed."""
    _, inp = load_golden()
    fr = inp.frames[0]
    H, W = fr["depth"].shape
    o = ch.Opts(steps=12, refine_steps=3)
    rec = ch.run_oracle_chain(inp, o)[0]
    d = fr["depth"]
    bg = d == 1.0
    quad_flat = np.zeros_like(bg)
    for y in range(H):
        for x in range(W):
            quad_flat[y, x] = (d[y, x ^ 1] == d[y, x]) and (d[y ^ 1, x] == d[y, x])
    discarded = bg & quad_flat
    assert discarded.any() and (bg & ~quad_flat).any()
    sentinel = np.full((H, W, 4), 7.0, np.float32)
    cp = rec["_k4_params"]
    out = orc.gi_compose(cp, d, fr["gbuffer"], rec["dn0"], rec["dn1"], sentinel)
    assert (out[discarded] == 7.0).all()          # untouched
    assert (out[~discarded][:, 3] == 1.0).all()   # written: alpha = 1
    k3 = rec["_k3"][0]
    s16 = np.full((H, W, 4), 5.0, np.float16)
    o0, o1 = orc.poisson_denoise(k3["params"], d, fr["gbuffer"], k3["in0"], k3["in1"], inp.blue, s16, s16)
    assert (o0[discarded] == 5.0).all() and (o1[discarded] == 5.0).all()
    assert not (o0[~discarded] == 5.0).all(axis=-1).any()


def test_ssgi_background_and_sentinel():
    """K1: background pixels carry packTwoVec4(directLight, directLight); pixels without a diffuse sample carry the -1 sentinel."""
    _, inp = load_golden()
    fr = inp.frames[0]
    rec = ch.run_oracle_chain(inp, ch.Opts(steps=12, refine_steps=3))[0]
    s = ch.unpack_halves(rec["ssgi"])
    bg = fr["depth"] == 1.0
    dl = fr["direct"].astype(np.float32)
    want = (dl + 1e-4).astype(np.float16).astype(np.float32) - 1e-4  # pack adds the offset, unpack_halves does not remove it
    got = s[bg]
    assert np.allclose(got[:, :4], (dl[bg] + np.float32(1e-4)).astype(np.float16).astype(np.float32), atol=0)
    assert np.allclose(got[:, 4:], got[:, :4], atol=0)
    fg = ~bg
    sent = s[fg][:, 0] < -0.5
    assert 0.02 < sent.mean() < 0.98                      # the diffuse/specular lottery picks both
    assert np.allclose(s[fg][sent][:, :3], -1.0 + 1e-4, atol=1e-3)


def test_oracle_is_deterministic_and_thread_count_invariant():
    _, inp = load_golden()
    o = ch.Opts(steps=12, refine_steps=3)
    a = ch.run_oracle_chain(inp, o)
    os.environ["OMP_NUM_THREADS"] = "1"
    b = ch.run_oracle_chain(inp, o)
    for t in range(2):
        for k in ("ssgi", "tr0", "dn1", "composed"):
            assert a[t][k].tobytes() == b[t][k].tobytes()


@pytest.mark.parametrize("iters", [0, 1, 2])
def test_denoise_iteration_count(iters):
    """denoiseIterations = n => 2n Poisson passes (PoissonDenoisePass.js:135-149); n = 0 never writes renderTargetB (A8)."""
    _, inp = load_golden()
    rec = ch.run_oracle_chain(inp, ch.Opts(steps=8, refine_steps=2, denoise_iterations=iters))[0]
    assert len(rec["_k3"]) == 2 * iters
    if iters == 0:
        assert not rec["dn0"].any() and not rec["dn1"].any()
