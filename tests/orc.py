"""ctypes binding of oracle/librfx_oracle.so — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  numpy arrays in, numpy arrays out.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from realism_effects_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "librfx_oracle.so")


class OrcEnv(C.Structure):
    _fields_ = [("mips", C.c_void_p * 16), ("mip_w", C.c_int32 * 16), ("mip_h", C.c_int32 * 16), ("levels", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("marginal", C.c_void_p), ("conditional", C.c_void_p),
                ("total_sum_whole", C.c_float), ("total_sum_decimal", C.c_float)]


_lib = None


def build(force: bool = False):
    src_m = max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("rfx_oracle.cpp", "glsl.h", "Makefile"))
    src_m = max(src_m, os.path.getmtime(os.path.join(ROOT, "include", "rfx.h")))
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < src_m:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], env={**os.environ, "CXX": "g++"})


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_half_to_float.restype = C.c_float
        _lib.orc_half_to_float.argtypes = [C.c_uint16]
        _lib.orc_float_to_half.restype = C.c_uint16
        _lib.orc_float_to_half.argtypes = [C.c_float]
        _lib.orc_pack_normal.restype = C.c_float
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    assert a.dtype == dtype, (a.dtype, dtype)
    return a


def f16bits(a):
    """float16 array -> uint16 view (contiguous)"""
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    return a.view(np.uint16) if a.dtype == np.float16 else _c(a, np.uint16)


class Env:
    """env map (H,W,4 float16) + CDF tables, with the oracle-built box-filter mip chain."""

    def __init__(self, map_f16: np.ndarray, marginal: np.ndarray, conditional: np.ndarray, total_sum: float):
        L = lib()
        self.h, self.w = map_f16.shape[:2]
        self.levels = [f16bits(map_f16).reshape(self.h, self.w, 4).copy()]
        w, h = self.w, self.h
        while w > 1 or h > 1:
            w1, h1 = max(1, w >> 1), max(1, h >> 1)
            dst = np.empty((h1, w1, 4), np.uint16)
            L.orc_env_downsample(_p(self.levels[-1]), C.c_int(w), C.c_int(h), _p(dst))
            self.levels.append(dst)
            w, h = w1, h1
        self.marginal = _c(marginal, np.float32)
        self.conditional = _c(conditional, np.float32)
        self.total_sum = float(total_sum)
        whole = float(int(self.total_sum))  # ~~totalSumValue  (EquirectHdrInfoUniform.js:346-349)
        self.whole, self.decimal = np.float32(whole), np.float32(self.total_sum - whole)
        s = OrcEnv()
        for i, l in enumerate(self.levels):
            s.mips[i] = l.ctypes.data
            s.mip_w[i] = l.shape[1]
            s.mip_h[i] = l.shape[0]
        s.levels = len(self.levels)
        s.width, s.height = self.w, self.h
        s.marginal = self.marginal.ctypes.data
        s.conditional = self.conditional.ctypes.data
        s.total_sum_whole, s.total_sum_decimal = float(self.whole), float(self.decimal)
        self.struct = s


def ssgi_trace(p: abi.SsgiParams, depth, gbuffer, velocity, direct_light, accumulated, env: Env | None, blue_noise, out_size=None):
    """out_size (W, H): the render target when it is smaller than the input planes (resolutionScale < 1); default = the planes' size"""
    TH, TW = depth.shape
    W, H = out_size or (TW, TH)
    out = np.zeros((H, W, 4), np.float32)
    bn = _c(blue_noise, np.uint8)
    lib().orc_ssgi_trace(C.byref(p), C.c_int(W), C.c_int(H), _p(_c(depth, np.float32)), _p(_c(gbuffer, np.float32)),
                         _p(_c(velocity, np.float32)), _p(f16bits(direct_light)), _p(_c(accumulated, np.float32)),
                         C.byref(env.struct) if env is not None else None, _p(bn), C.c_int(bn.shape[1]), C.c_int(bn.shape[0]), _p(out), C.c_int(TW), C.c_int(TH))
    return out


def temporal_reproject(p: abi.TemporalParams, inp, velocity, hist0, hist1, out0_prev, out1_prev, out_half=False):
    """out*_prev: previous contents of the targets (kept where the shader discards)."""
    H, W = velocity.shape[:2]
    input_half = inp.dtype in (np.float16, np.uint16)
    inp_c = f16bits(inp) if input_half else _c(inp, np.float32)
    o0 = np.array(f16bits(out0_prev) if out_half else out0_prev, copy=True)
    o1 = None if out1_prev is None else np.array(f16bits(out1_prev) if out_half else out1_prev, copy=True)
    hist_float = hist0.dtype == np.float32  # denoiseMode "full_temporal" / "temporal": RGBA32F history
    if hist_float:
        h0, h1 = _c(hist0, np.float32), _c(hist1, np.float32) if hist1 is not None else None
    else:
        h0, h1 = f16bits(hist0), f16bits(hist1) if hist1 is not None else None
    lib().orc_temporal_reproject(C.byref(p), C.c_int(W), C.c_int(H), _p(inp_c), C.c_int(int(input_half)), _p(_c(velocity, np.float32)),
                                 _p(h0), _p(h1), _p(o0), _p(o1), C.c_int(int(out_half)), C.c_int(int(hist_float)), C.c_int(inp.shape[1]), C.c_int(inp.shape[0]))
    if out_half:
        o0 = o0.view(np.float16)
        o1 = None if o1 is None else o1.view(np.float16)
    return o0, o1


def poisson_denoise(p: abi.PoissonParams, depth, gbuffer_or_normal, in0, in1, blue_noise, out0_prev, out1_prev):
    H, W = depth.shape
    in_half = in0.dtype in (np.float16, np.uint16)
    i0 = f16bits(in0) if in_half else _c(in0, np.float32)
    i1 = None if in1 is None else (f16bits(in1) if in_half else _c(in1, np.float32))
    o0 = np.array(f16bits(out0_prev), copy=True)
    o1 = None if out1_prev is None else np.array(f16bits(out1_prev), copy=True)
    bn = _c(blue_noise, np.uint8)
    lib().orc_poisson_denoise(C.byref(p), C.c_int(W), C.c_int(H), _p(_c(depth, np.float32)), _p(_c(gbuffer_or_normal, np.float32)), _p(i0),
                              _p(i1), C.c_int(int(in_half)), _p(bn), C.c_int(bn.shape[1]), C.c_int(bn.shape[0]), _p(o0), _p(o1))
    return o0.view(np.float16), None if o1 is None else o1.view(np.float16)


def gi_compose(p: abi.ComposeParams, depth, gbuffer, diffuse_gi, specular_gi, out_prev, scene=None):
    """diffuse_gi / specular_gi / scene may be None (null sampler)"""
    H, W = depth.shape
    out = np.array(out_prev, np.float32, copy=True)
    gi_float = any(a is not None and a.dtype == np.float32 for a in (diffuse_gi, specular_gi))
    cv = (lambda a: None if a is None else _c(a, np.float32)) if gi_float else f16bits
    lib().orc_gi_compose(C.byref(p), C.c_int(W), C.c_int(H), _p(_c(depth, np.float32)), _p(_c(gbuffer, np.float32)), _p(cv(diffuse_gi)),
                         _p(cv(specular_gi)), _p(f16bits(scene)), _p(out), C.c_int(int(gi_float)))
    return out


def ssgi_compose(depth, gi, scene, params=None):
    """params: abi.SsgiComposeParams (fog / debug) or None"""
    H, W = depth.shape
    out = np.zeros((H, W, 4), np.uint16)
    lib().orc_ssgi_compose(C.byref(params) if params is not None else None, C.c_int(W), C.c_int(H), _p(_c(depth, np.float32)), _p(_c(gi, np.float32)),
                           _p(f16bits(scene)), _p(out))
    return out.view(np.float16)


def hbao(p: abi.HbaoParams, depth, blue_noise, out_prev):
    H, W = depth.shape
    out = np.array(f16bits(out_prev), copy=True)
    bn = _c(blue_noise, np.uint8)
    lib().orc_hbao(C.byref(p), C.c_int(W), C.c_int(H), _p(_c(depth, np.float32)), _p(bn), C.c_int(bn.shape[1]), C.c_int(bn.shape[0]), _p(out))
    return out.view(np.float16)


def ao_compose(p: abi.AoComposeParams, depth, ao, inp):
    H, W = depth.shape
    out = np.zeros((H, W, 4), np.uint16)
    lib().orc_ao_compose(C.byref(p), C.c_int(W), C.c_int(H), _p(_c(depth, np.float32)), _p(f16bits(ao)), _p(f16bits(inp)), _p(out))
    return out.view(np.float16)


def motion_blur(p: abi.MotionBlurParams, velocity, inp, blue_noise):
    H, W = velocity.shape[:2]
    out = np.zeros((H, W, 4), np.uint16)
    bn = _c(blue_noise, np.uint8)
    lib().orc_motion_blur(C.byref(p), C.c_int(W), C.c_int(H), _p(_c(velocity, np.float32)), _p(f16bits(inp)), _p(bn), C.c_int(bn.shape[1]),
                          C.c_int(bn.shape[0]), _p(out))
    return out.view(np.float16)


_FMT_OF = {np.dtype(np.uint8): 3, np.dtype(np.float16): 2, np.dtype(np.uint16): 2, np.dtype(np.float32): 1}


def gbuffer_ingest(albedo, normal, material, emissive, motion, depth, *, motion_scale=(1.0, 1.0), normalize_normals=True):
    """-> (gBuffer, velocity) RGBA32F.  albedo / material uint8 | float16, normal / motion float16 | float32 (H,W,4); emissive float16 | None"""
    H, W = depth.shape
    p = abi.IngestParams()
    p.motion_scale[:] = [float(motion_scale[0]), float(motion_scale[1])]
    p.normalize_normals = int(bool(normalize_normals))
    gb, vel = np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32)
    c = lambda a: None if a is None else np.ascontiguousarray(a)  # noqa: E731
    albedo, normal, material, emissive, motion = c(albedo), c(normal), c(material), c(emissive), c(motion)
    lib().orc_gbuffer_ingest(C.byref(p), C.c_int(W), C.c_int(H), _p(albedo), C.c_int(_FMT_OF[albedo.dtype]), _p(normal), C.c_int(_FMT_OF[normal.dtype]),
                             _p(material), C.c_int(_FMT_OF[material.dtype]), _p(emissive), _p(motion), C.c_int(_FMT_OF[motion.dtype] if motion is not None else 1),
                             _p(_c(depth, np.float32)), _p(gb), _p(vel))
    return gb, vel


def effects(p: abi.EffectsParams, inp, depth, velocity):
    """merged cosmetic effects (EffectPass semantics) -> RGBA16F"""
    H, W = inp.shape[:2]
    out = np.zeros((H, W, 4), np.uint16)
    lib().orc_effects(C.byref(p), C.c_int(W), C.c_int(H), _p(f16bits(inp)), _p(None if depth is None else _c(depth, np.float32)),
                      _p(None if velocity is None else _c(velocity, np.float32)), _p(out))
    return out.view(np.float16)


def taa(p: abi.TaaParams, inp, history):
    """TAAPass -> RGBA8"""
    H, W = inp.shape[:2]
    out = np.zeros((H, W, 4), np.uint8)
    lib().orc_taa(C.byref(p), C.c_int(W), C.c_int(H), _p(f16bits(inp)), _p(_c(history, np.uint8)), _p(out))
    return out


def traa_compose(acc):
    H, W = acc.shape[:2]
    out = np.zeros((H, W, 4), np.uint16)
    lib().orc_traa_compose(C.c_int(W), C.c_int(H), _p(f16bits(acc)), _p(out))
    return out.view(np.float16)
