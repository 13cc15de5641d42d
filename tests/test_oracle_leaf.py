"""CPU tests of the oracle's leaf functions against independent restatements (numpy / torch / pure Python).
The reference ships no vectors, so these pin the oracle's building blocks to independently
written code and to properties stated by the GLSL."""
import ctypes as C

import numpy as np
import torch

import orc
from realism_effects_b200 import synth


def test_half_conversion_matches_numpy_bit_exact():
    L = orc.lib()
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * 10.0 ** rng.integers(-9, 6, 20000),
        np.array([0.0, -0.0, 1.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 5.96e-8, 2.98e-8, 2.9802322e-8, 3e-8, 6.1e-5, 6.0975e-5, np.inf, -np.inf], np.float32),
        (np.arange(0, 70000, 7).astype(np.float32) / 3.0),
    ])
    want = vals.astype(np.float16).view(np.uint16)
    got = np.array([L.orc_float_to_half(C.c_float(float(v))) for v in vals], np.uint16)
    assert (got == want).all()
    halves = np.arange(0, 65536, 13, dtype=np.uint16)
    back = np.array([L.orc_half_to_float(C.c_uint16(int(h))) for h in halves], np.float32)
    ref = halves.view(np.float16).astype(np.float32)
    assert ((back == ref) | (np.isnan(back) & np.isnan(ref))).all()


def test_pack_gbuffer_matches_torch_restatement_bit_exact():
    """oracle packGBuffer (C++) vs synth.pack_gbuffer (torch): two independent restatements of gbuffer_packing.glsl:166-178"""
    L = orc.lib()
    rng = np.random.default_rng(1)
    n = 2000
    diffuse = np.concatenate([rng.integers(0, 256, (n, 3)) / 255.0, np.ones((n, 1))], 1).astype(np.float32)
    nrm = rng.standard_normal((n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    rough = rng.choice([0.0, 0.05, 0.3, 0.6, 1.0, 0.123], n).astype(np.float32)
    metal = rng.choice([0.0, 1.0, 0.5], n).astype(np.float32)
    emis = (rng.random((n, 3)) * rng.choice([0.5, 2.0, 7.0], (n, 1))).astype(np.float32)
    t = synth.pack_gbuffer(torch.tensor(diffuse), torch.tensor(nrm), torch.tensor(rough), torch.tensor(metal), torch.tensor(emis)).numpy()
    out = np.zeros(4, np.float32)
    bad = 0
    for i in range(n):
        L.orc_pack_gbuffer(diffuse[i].ctypes.data_as(C.c_void_p), nrm[i].ctypes.data_as(C.c_void_p), C.c_float(float(rough[i])), C.c_float(float(metal[i])),
                           emis[i].ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        # .a (RGBE) goes through log2/exp2 whose last-ulp behaviour may differ between torch and libm: compare r,g,b bit-exactly, a by decode
        if not (out.view(np.uint32)[:3] == t[i].view(np.uint32)[:3]).all():
            bad += 1
    assert bad == 0


def test_gbuffer_roundtrip_properties():
    L = orc.lib()
    rng = np.random.default_rng(2)
    out4, out12 = np.zeros(4, np.float32), np.zeros(12, np.float32)
    for _ in range(500):
        d = np.append(rng.integers(0, 256, 3) / 255.0, 1.0).astype(np.float32)
        n = rng.standard_normal(3).astype(np.float32)
        n /= np.linalg.norm(n)
        r, m = float(rng.integers(0, 257) / 256.0), float(rng.integers(0, 2))
        e = (rng.random(3) * 4 + 0.01).astype(np.float32)
        L.orc_pack_gbuffer(d.ctypes.data_as(C.c_void_p), n.ctypes.data_as(C.c_void_p), C.c_float(r), C.c_float(m), e.ctypes.data_as(C.c_void_p), out4.ctypes.data_as(C.c_void_p))
        L.orc_unpack_gbuffer(out4.ctypes.data_as(C.c_void_p), out12.ctypes.data_as(C.c_void_p))
        assert np.abs(out12[:3] - d[:3]).max() < 1 / 255 + 2e-4          # 8-bit albedo (truncation)
        assert np.abs(out12[4:7] - n).max() < 2e-3                        # oct-encoded fp16 normal
        assert abs(np.linalg.norm(out12[4:7]) - 1) < 1e-6
        assert abs(out12[8] - min(m, 0.999999)) < 1 / 256 + 2e-4
        if m == 0:  # with metalness 1 the packed value exceeds 2^24 and float2color's mod() loses the low bits (reference quirk, kept):
            assert abs(out12[7] - min(r, 0.999999)) < 1 / 256 + 2e-4
        else:       # roughness comes back within 2/256 or wraps to 0 (e.g. roughness 1, metalness 1 decodes as roughness 0)
            assert abs(out12[7] - min(r, 0.999999)) < 2 / 256 + 2e-4 or out12[7] < 2 / 256
        # RGBE8: floatToVec4 also subtracts 1e-4 from the exponent byte, so exp2(a*255-128) is ~1.8% low (reference quirk, kept)
        assert np.abs(out12[9:12] - e).max() / e.max() < 0.03 and (out12[9:12] <= e * 1.001).all()


def test_pack_two_vec4_roundtrip():
    L = orc.lib()
    rng = np.random.default_rng(3)
    a, b, e, a2, b2 = (np.zeros(4, np.float32) for _ in range(5))
    for _ in range(300):
        a[:] = rng.random(4) * 10
        b[:] = rng.random(4) * 10
        a[0] = -1.0 if rng.random() < 0.3 else a[0]  # the "no diffuse sample" sentinel survives the round trip as a negative value
        L.orc_pack_two_vec4(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p))
        L.orc_unpack_two_vec4(e.ctypes.data_as(C.c_void_p), a2.ctypes.data_as(C.c_void_p), b2.ctypes.data_as(C.c_void_p))
        assert np.allclose(a2, a, rtol=1e-3, atol=2e-4) and np.allclose(b2, b, rtol=1e-3, atol=2e-4)
        assert (a2[0] < 0) == (a[0] < 0)


def _pcg4d(v):
    M = 0xFFFFFFFF
    v = [(x * 1664525 + 1013904223) & M for x in v]
    v[0] = (v[0] + v[1] * v[3]) & M; v[1] = (v[1] + v[2] * v[0]) & M; v[2] = (v[2] + v[0] * v[1]) & M; v[3] = (v[3] + v[1] * v[2]) & M
    v = [x ^ (x >> 16) for x in v]
    v[0] = (v[0] + v[1] * v[3]) & M; v[1] = (v[1] + v[2] * v[0]) & M; v[2] = (v[2] + v[0] * v[1]) & M; v[3] = (v[3] + v[1] * v[2]) & M
    return v


def test_blue_noise_coordinates_match_python_bigint_restatement():
    """blue_noise.glsl:9-34 with Python integers (wraparound made explicit), incl. indices near 2^31 (int overflow of index*15843)."""
    L = orc.lib()
    sx, sy = C.c_int(), C.c_int()
    M = 0xFFFFFFFF
    for index in [1, 2, 77, 1234568, 2469136, 0x7FFFFFFE, 0x7FFFFFF0, 1 << 30, 123456789]:
        s1 = _pcg4d([index & M, (index * 15843) & M, (index * 31 + 4566) & M, (index * 2345 + 58585) & M])
        for (x, y) in [(0, 0), (5, 9), (3839, 2159), (127, 128), (7679, 4319)]:
            L.orc_blue_noise_coord(x, y, index, 128, C.byref(sx), C.byref(sy))
            assert sx.value == (x + s1[0] % 0x0FFFFFFF) % 128 and sy.value == (y + s1[1] % 0x0FFFFFFF) % 128


def test_blue_noise_asset_is_the_reference_texture():
    import hashlib

    bn = synth.load_blue_noise()
    assert bn.shape == (128, 128, 4) and bn.dtype == np.uint8
    assert hashlib.sha256(bn.tobytes()).hexdigest() == "705a8dcdaf4fc41b14c19cce5d62dafb99d68f6687ad5f917bf6d90df37c44c4"
    assert np.allclose(bn.reshape(-1, 4).mean(0), 127.5)  # all four channels have mean 127.5
