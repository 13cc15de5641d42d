"""Pins the oracle against the reference's own shaders run on the CPU (tests/refglsl.py), and writes the golden fixtures.

    python tools/pin_oracle.py                 # the matrix below, small frames; prints a table, exits 1 on any mismatch
    python tools/pin_oracle.py --golden        # + writes tests/golden/ref_glsl/*.npz (outputs of the REFERENCE shaders)
    python tools/pin_oracle.py --size 1920 1080 --frames 3   # one chain at a BASELINE size (evidence for profiles/)

Needs the reference checkout to build the shaders (here: /root/reference); TEST INFRASTRUCTURE.
Every case drives the same inputs through (a) oracle/rfx_oracle.cpp and (b) the compiled reference GLSL, pass by pass, with the
reference's frame logic (tests/chain_harness.py), and compares every output plane bit for bit.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import chain_harness as ch  # noqa: E402
import orc  # noqa: E402
import refglsl  # noqa: E402
from realism_effects_b200 import abi  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden", "ref_glsl")
PLANES = ("ssgi", "tr0", "tr1", "dn0", "dn1", "composed")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint16) if a.dtype == np.float16 else (a if a.dtype == np.uint8 else a.view(np.uint32))


def diff(a, b) -> dict:
    """bit-level comparison (NaN payloads included)"""
    ne = _bits(a) != _bits(b)
    px = ne.reshape(ne.shape[0], ne.shape[1], -1).any(-1)
    return dict(n_px=int(px.sum()), frac=float(px.mean()))


CHAIN_CASES = {
    # name: (W, H, frames, Opts kwargs, make_inputs kwargs)
    "ssgi_default": (96, 64, 3, {}, {}),
    "ssgi_iter2_odd": (97, 61, 2, dict(denoise_iterations=2), {}),
    "ssgi_noenv": (64, 48, 2, dict(use_envmap=False, importance_sampling=False), {}),
    "ssgi_no_is": (64, 48, 2, dict(importance_sampling=False), {}),
    "ssgi_missed_rays": (64, 48, 2, dict(missed_rays=True, steps=8, refine_steps=2), {}),
    "ssgi_static": (64, 48, 3, {}, dict(static=True)),
    "ssgi_refine0": (64, 48, 1, dict(refine_steps=0, steps=12), {}),
    "ssr": (96, 64, 3, dict(mode=abi.MODE_SSR), {}),
    "ssgi_full_temporal": (64, 48, 3, dict(denoise_mode=abi.DENOISE_FULL_TEMPORAL), {}),
    "ssgi_temporal": (64, 48, 3, dict(denoise_mode=abi.DENOISE_TEMPORAL), {}),
    "ssr_full_temporal": (64, 48, 3, dict(mode=abi.MODE_SSR, denoise_mode=abi.DENOISE_FULL_TEMPORAL), {}),
    "ssgi_scale_0.5": (96, 56, 3, dict(resolution_scale=0.5), {}),
    "ssgi_scale_0.75": (96, 56, 2, dict(resolution_scale=0.75), {}),
    "ssr_scale_0.5": (96, 56, 2, dict(mode=abi.MODE_SSR, resolution_scale=0.5), {}),
    # OrthographicCamera: the #else branches of PERSPECTIVE_CAMERA in K1 (ray origin / direction), getViewZ, K2's and K4's view directions
    "ssgi_orthographic": (80, 54, 3, {}, dict(orthographic=True)),
    "ssr_orthographic": (80, 54, 2, dict(mode=abi.MODE_SSR), dict(orthographic=True)),
    "ssgi_ortho_missed": (64, 48, 2, dict(missed_rays=True, importance_sampling=False), dict(orthographic=True)),
}


def run_chain_case(name, golden: bool):
    W, H, F, okw, ikw = CHAIN_CASES[name]
    o = ch.Opts(**okw)
    inp = ch.make_inputs(W, H, F, **ikw)
    planes = [p for p in PLANES if not (o.mode == abi.MODE_SSR and p in ("tr1", "dn1")) and not (o.denoise_mode != 0 and p in ("dn0", "dn1"))]
    a = ch.run_oracle_chain(inp, o, capture=planes, lean=True)
    t = time.time()
    b = ch.run_oracle_chain(inp, o, capture=planes, lean=True, impl=refglsl)
    dt = time.time() - t
    rows, worst = [], 0
    for f, (x, y) in enumerate(zip(a, b)):
        for k in planes:
            d = diff(x[k], y[k])
            worst = max(worst, d["n_px"])
            rows.append((f"{name} f{f} {k}", d))
    if golden:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        rec = {f"f{f}_{k}": _bits(y[k]) for f, y in enumerate(b) for k in planes}
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"chain_{name}.npz"), **rec)
    return rows, worst, dt


def run_effect_cases(golden: bool):
    """K5 (+fog, debug), K6 + the 1-plane Poisson passes + K7, K8, the TRAA form of K2 + K9 — inputs as tests/test_gpu_effects.py"""
    rows, rec = [], {}
    inp = ch.make_inputs(96, 54, 2)
    f0, f1 = inp.frames
    W, H = 96, 54
    z = np.zeros((H, W, 4), np.float16)

    def both(tag, fn):
        a, b = fn(orc), fn(refglsl)
        for i, (x, y) in enumerate(zip(a, b)):
            rows.append((f"{tag}[{i}]", diff(x, y)))
            rec[f"{tag}_{i}"] = _bits(y)
        return b

    # K6 -> K3 x2 (single plane, velocity-layout normals) -> K7
    def hbao_chain(m):
        ao = m.hbao(ch.hbao_params(f1["cam"], 778), f1["depth"], inp.blue, z)
        tgtA, tgtB = z.copy(), z.copy()
        idx = 1234567
        for i in range(2):
            idx += 1
            p = ch.poisson_params(ch.Opts(), idx, False)
            p.texture_count, p.gbuffer_texture, p.input_linear = 1, 0, 1
            p.is_texture_specular[:] = [0, 0]
            p.normal_phi, p.depth_phi, p.roughness_phi, p.specular_phi = 3.25, 2.0, 0.0, 0.0
            out, _ = m.poisson_denoise(p, f1["depth"], f1["velocity"], ao if i == 0 else tgtA, None, inp.blue, tgtA if i == 0 else tgtB, None)
            if i == 0:
                tgtA = out
            else:
                tgtB = out
        k7 = m.ao_compose(ch.ao_compose_params(), f1["depth"], tgtB, f1["direct"])
        return ao, tgtA, tgtB, k7

    both("hbao_chain", hbao_chain)

    # TRAA: K2 one plane fp16 + K9
    def traa(m):
        p0 = ch.traa_temporal_params(abi.make_camera(f0["cam"]), f0["cam"]["position"], f0["cam"], 0.0)
        h0, _ = m.temporal_reproject(p0, f0["direct"], f0["velocity"], z, None, z, None, out_half=True)
        p1 = ch.traa_temporal_params(abi.make_camera(f1["cam"]), f1["cam"]["position"], f0["cam"], 1.0)
        h1, _ = m.temporal_reproject(p1, f1["direct"], f1["velocity"], h0, None, h0, None, out_half=True)
        return h0, h1, m.traa_compose(h1)

    both("traa", traa)

    # K8
    vel = ch.rotation_velocity_field(W, H, f1["depth"])
    both("motion_blur", lambda m: (m.motion_blur(ch.motion_blur_params(W, H, frame=7), vel, f1["direct"], inp.blue),
                                   m.motion_blur(ch.motion_blur_params(W, H, frame=0, samples=8, resolution=(1280, 720)), vel, f1["direct"], inp.blue)))

    # K5
    rng = np.random.default_rng(3)
    gi = rng.uniform(0, 2, (H, W, 4)).astype(np.float32)

    def k5(m):
        outs = [m.ssgi_compose(f1["depth"], gi, f1["direct"])]
        for exp2 in (0, 1):
            p = abi.SsgiComposeParams()
            p.use_fog, p.fog_exp2, p.perspective = 1, exp2, 1
            p.fog_color[:] = [0.6, 0.7, 0.8]
            p.fog_near, p.fog_far, p.fog_density = 2.0, 30.0, 0.05
            p.camera_near, p.camera_far = float(f1["cam"]["near"]), float(f1["cam"]["far"])
            outs.append(m.ssgi_compose(f1["depth"], gi, f1["direct"], p))
        p = abi.SsgiComposeParams()
        p.is_debug, p.perspective = 1, 1
        outs.append(m.ssgi_compose(f1["depth"], gi, f1["direct"], p))
        return outs

    both("ssgi_compose", k5)

    # cosmetic effects (EffectPass merges) and TAAPass
    both("cosmetic_effects", lambda m: [m.effects(ch.fx_params(f1["cam"], effs, sp), f1["direct"], f1["depth"], f1["velocity"]) for effs, sp in ch.FX_CASES])
    hist = np.random.default_rng(1).integers(0, 256, (H, W, 4), dtype=np.uint8)
    both("taa_pass", lambda m: [m.taa(p, f1["direct"], hist) for p in ch.taa_cases()])
    if golden:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "effects.npz"), **rec)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", action="store_true")
    ap.add_argument("--size", type=int, nargs=2)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--json", type=str)
    a = ap.parse_args()
    if not refglsl.assemble.available():
        sys.exit("the reference checkout is needed to build the shaders")
    report = {}
    bad = 0
    if a.size:
        W, H = a.size
        CHAIN_CASES[f"c_{W}x{H}"] = (W, H, a.frames, dict(denoise_iterations=2 if W >= 3840 else 1), {})
        cases = [f"c_{W}x{H}"]
    else:
        cases = list(CHAIN_CASES)
    for name in cases:
        rows, worst, dt = run_chain_case(name, a.golden and not a.size)
        bad += worst
        report[name] = dict(planes_compared=len(rows), worst_differing_pixels=worst, reference_glsl_seconds=round(dt, 1))
        print(f"{name:18s} {len(rows):3d} planes  worst plane: {worst} differing pixels   (reference GLSL {dt:.1f} s)")
        for tag, d in rows:
            if d["n_px"]:
                print(f"    {tag}: {d['n_px']} px ({d['frac']:.2e})")
    if not a.size:
        rows = run_effect_cases(a.golden)
        n = sum(d["n_px"] for _, d in rows)
        bad += n
        report["effects"] = dict(planes_compared=len(rows), differing_pixels=n)
        print(f"{'effects':18s} {len(rows):3d} planes  {n} differing pixels in total")
        for tag, d in rows:
            if d["n_px"]:
                print(f"    {tag}: {d['n_px']} px ({d['frac']:.2e})")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(report, f, indent=1)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
