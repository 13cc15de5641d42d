"""CPU tests of the host-side logic that surrounds the kernels: camera math, blue-noise index sequence,
env-map CDF tables, the synthetic generator's plane layouts."""
import numpy as np
import torch

import chain_harness as ch
from realism_effects_b200 import abi, synth


def test_camera_matrices_follow_threejs_conventions():
    cam = synth.Camera(aspect=16 / 9)
    u = cam.uniforms()
    P = u["projection"].reshape(4, 4).T
    # three.js makePerspective: te[11] = -1, te[15] = 0, te[14] = -2fn/(f-n)
    assert P[3, 2] == -1 and P[3, 3] == 0
    assert np.isclose(P[2, 3], -2 * 250 * 0.01 / (250 - 0.01))
    assert np.isclose(P[1, 1], 1 / np.tan(np.radians(20)))
    assert np.allclose(P @ u["projection_inverse"].reshape(4, 4).T, np.eye(4), atol=1e-4)
    Wm = u["camera_matrix_world"].reshape(4, 4).T
    assert np.allclose(Wm @ u["view_matrix"].reshape(4, 4).T, np.eye(4), atol=1e-5)
    assert np.allclose(Wm[:3, 3], [0, 8.75, 25]) and np.allclose(Wm[:3, 2], [0, 0, 1])  # looks down -z


def test_blue_noise_index_sequence():
    """BlueNoiseUtils.js:19-28: index_n = (start + index_{n-1} + 1) % 0x7fffffff, index_0 = 0"""
    start, c, seen = 1234567, 0, []
    for _ in range(5):
        c = ch.next_blue(start, c)
        seen.append(c)
    assert seen == [1234568, 2469136, 3703704, 4938272, 6172840]
    assert ch.next_blue(0x7FFFFFF0, 0x7FFFFFF0) == (0x7FFFFFF0 * 2 + 1) % 0x7FFFFFFF


def test_env_cdf_tables():
    env = synth.synthetic_env(64, 32).astype(np.float32)
    marg, cond, total = synth.build_env_cdf(env)
    assert marg.shape == (32,) and cond.shape == (32, 64)
    assert (np.diff(marg) >= 0).all() and (np.diff(cond, axis=1) >= 0).all()
    assert ((marg * 32 - 0.5) % 1 == 0).all() and ((cond * 64 - 0.5) % 1 == 0).all()  # half-texel centred indices
    lum = 0.2126 * env[..., 0] + 0.7152 * env[..., 1] + 0.0722 * env[..., 2]
    assert np.isclose(total, lum.astype(np.float64).sum(), rtol=1e-6)
    # importance: the brightest row must own more of the [0,1) range than a dark one
    rows = ((marg * 32 - 0.5)).astype(int)
    counts = np.bincount(rows, minlength=32)
    assert counts[np.argmax(lum.sum(1))] >= counts[np.argmin(lum.sum(1))]
    # A4: with flipY the reference's in-place "un-flip" mirrors the top half
    m2, c2, t2 = synth.build_env_cdf(env, flip_y=True)
    mirrored = env.copy()
    for y in range(32):
        mirrored[31 - y] = mirrored[y]
    m3, c3, t3 = synth.build_env_cdf(mirrored)
    assert np.array_equal(m2, m3) and np.array_equal(c2, c3) and t2 == t3
    assert not np.array_equal(m2, marg)


def test_synthetic_planes_follow_reference_layouts():
    fr = synth.render_frame(64, 36, t=1)
    d, g, v = fr.depth.numpy(), fr.gbuffer.numpy(), fr.velocity.numpy()
    bg = d == 1.0
    assert 0.05 < bg.mean() < 0.5 and (d[~bg] < 1).all() and (d[~bg] > 0.99).all()
    assert (g[bg] == [0, 0, 0, 1]).all() and (v[bg] == [0, 0, 0, 1]).all()         # cleared targets
    assert np.array_equal(v[..., 3], d)                                             # velocity.a = depth
    assert np.array_equal(v[~bg][:, 2].view(np.uint32), g[~bg][:, 1].view(np.uint32))  # same packed oct normal
    assert np.abs(v[~bg][:, :2]).max() > 1e-5                                       # camera moved between t-1 and t
    st = synth.render_frame(64, 36, t=1, static=True)
    assert not st.velocity.numpy()[..., :2].any()
    # packed albedo byte 3 (alpha = 1.0 -> 254 after the ONE_SAFE clamp) keeps the float away from NaN/Inf bit patterns
    bits = g[~bg][:, 0].view(np.uint32)
    assert ((bits >> 24) == 254).all()
    assert fr.direct_light.dtype == torch.float16 and tuple(fr.direct_light.shape) == (36, 64, 4)


def test_chain_option_mirrors_defaults():
    """Opts defaults == src/ssgi/SSGIOptions.js:26-48"""
    o = ch.Opts()
    assert (o.distance, o.thickness, o.denoise_iterations, o.radius, o.phi, o.luma_phi, o.depth_phi, o.normal_phi, o.roughness_phi, o.specular_phi,
            o.env_blur, o.steps, o.refine_steps) == (10, 10, 1, 3, 0.5, 5, 2, 50, 50, 50, 0.5, 20, 5)
    assert o.flags == abi.SSGI_IMPORTANCE_SAMPLING | abi.SSGI_USE_DIRECT_LIGHT | abi.SSGI_USE_ENVMAP
    assert ch.max_mip_level(1024, 512) == 11.0
