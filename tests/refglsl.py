"""The reference's own fragment shaders, compiled for the CPU and run from numpy — TEST INFRASTRUCTURE.

oracle/ref/assemble.py reads the GLSL from the reference checkout and assembles it the way the reference's JS does,
oracle/ref/transpile.py re-spells it as C++ against the GLSL runtime oracle/ref/glsl_rt.h and builds oracle/_ref/*.so.
This module binds those libraries and wires each pass's uniforms / textures the way the reference's JS host code
does (citations per function), behind the same signatures as tests/orc.py so that the two can be swapped:
`refglsl.ssgi_trace(...)` is "the reference run here", `orc.ssgi_trace(...)` the restatement it pins.

Needs the reference checkout (RFX_REFERENCE_DIR, default /root/reference) to BUILD a shader; once built, the .so in
oracle/_ref/ is enough to run it (the GPU box has no checkout).  `available()` says whether pass `name` can run.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref"))
import assemble  # noqa: E402
import transpile  # noqa: E402

from realism_effects_b200 import abi  # noqa: E402

F_R32F, F_RGBA32F, F_RGBA16F, F_RGBA8 = 0, 1, 2, 3
_NP = {F_R32F: (np.float32, 1), F_RGBA32F: (np.float32, 4), F_RGBA16F: (np.uint16, 4), F_RGBA8: (np.uint8, 4)}
MANIFEST = os.path.join(transpile.OUT_DIR, "manifest.json")


def _manifest() -> dict:
    if os.path.exists(MANIFEST):
        with open(MANIFEST) as f:
            return json.load(f)
    return {}


def _key(name: str, kw: dict) -> str:
    return name + "(" + ",".join(f"{k}={kw[k]!r}" for k in sorted(kw)) + ")"


def available(name: str | None = None, **kw) -> bool:
    if assemble.available():
        return True
    if name is None:
        return bool(_manifest())
    e = _manifest().get(_key(name, kw))
    return bool(e) and os.path.exists(os.path.join(transpile.OUT_DIR, e["so"]))


class Shader:
    """One compiled reference fragment shader."""

    _cache: dict = {}

    def __init__(self, name: str, glsl: str | None = None, **kw):
        """name + kw: a pass of oracle/ref/assemble.py;  glsl: explicit shader text (the runtime's own tests)"""
        key = _key(name, kw)
        if glsl is not None:
            so, info = transpile.build(name, glsl)
            self.n_out = len(info["outputs"])
        elif assemble.available():
            glsl = getattr(assemble, name)(**kw)
            so, info = transpile.build(name, glsl)
            man = _manifest()
            if man.get(key, {}).get("so") != os.path.basename(so):
                man[key] = dict(so=os.path.basename(so), outputs=info["outputs"])
                with open(MANIFEST, "w") as f:
                    json.dump(man, f, indent=1, sort_keys=True)
            self.n_out = len(info["outputs"])
        else:
            e = _manifest().get(key)
            if not e:
                raise RuntimeError(f"reference shader {key} was not prebuilt and the reference checkout is absent")
            so, self.n_out = os.path.join(transpile.OUT_DIR, e["so"]), len(e["outputs"])
        self.lib = C.CDLL(so)
        L = self.lib
        L.rtx_create.restype = C.c_void_p
        L.rtx_destroy.argtypes = [C.c_void_p]
        L.rtx_uniform.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
        L.rtx_tex_create.restype = C.c_void_p
        L.rtx_tex_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.rtx_tex_add_mip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.rtx_tex_destroy.argtypes = [C.c_void_p]
        L.rtx_sampler.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.rtx_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        self.h = L.rtx_create()
        self._keep = []
        self._texs = []

    @classmethod
    def get(cls, name: str, **kw) -> "Shader":
        key = _key(name, kw)
        if key not in cls._cache:
            cls._cache[key] = cls(name, **kw)
        s = cls._cache[key]
        s.reset()
        return s

    def reset(self):
        for t in self._texs:
            self.lib.rtx_tex_destroy(t)
        self._texs, self._keep = [], []
        self.lib.rtx_destroy(self.h)
        self.h = self.lib.rtx_create()

    def set(self, optional=(), **uniforms):
        """values: number / sequence (matrices column-major, 16 numbers).  Names in `optional` may be absent from the shader
        (the JS sets uniforms the shader text does not declare; GL drops them)."""
        for name, v in uniforms.items():
            a = np.ascontiguousarray(np.asarray(v, np.float64).reshape(-1))
            r = self.lib.rtx_uniform(self.h, name.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
            if r == -1 and name in optional:
                continue
            if r != 0:
                raise KeyError(f"uniform {name}: {'not declared by the shader' if r == -1 else 'wrong component count'}")

    def tex(self, name: str, data, fmt: int, *, linear=False, repeat=False, mips=None, optional=False):
        """data: numpy array (H, W[, 4]) or None = null sampler.  mips: list of arrays (levels 0..n) for a trilinear texture."""
        if data is None:
            return
        dt, ch = _NP[fmt]
        a = np.ascontiguousarray(data)
        if a.dtype == np.float16:
            a = a.view(np.uint16)
        assert a.dtype == dt, (name, a.dtype, dt)
        h, w = a.shape[:2]
        t = self.lib.rtx_tex_create(a.ctypes.data_as(C.c_void_p), w, h, fmt, int(linear), int(repeat))
        self._keep.append(a)
        self._texs.append(t)
        for m in mips or []:
            m = np.ascontiguousarray(m)
            if m.dtype == np.float16:
                m = m.view(np.uint16)
            self._keep.append(m)
            self.lib.rtx_tex_add_mip(t, m.ctypes.data_as(C.c_void_p), m.shape[1], m.shape[0])
        r = self.lib.rtx_sampler(self.h, name.encode(), t)
        if r != 0 and not optional:
            raise KeyError(f"sampler {name} is not declared by the shader")

    def run(self, W: int, H: int, outs):
        """outs: list of (fmt, previous contents or None) -> list of arrays (discarded pixels keep the previous contents)."""
        assert len(outs) == self.n_out, (len(outs), self.n_out)
        arrs = []
        for fmt, prev in outs:
            dt, ch = _NP[fmt]
            if prev is None:
                a = np.zeros((H, W, ch) if ch > 1 else (H, W), dt)
            else:
                a = np.array(prev.view(np.uint16) if prev.dtype == np.float16 else prev, dtype=dt, copy=True)
            arrs.append(a)
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        fmts = (C.c_int * len(arrs))(*[f for f, _ in outs])
        self.lib.rtx_run(self.h, W, H, len(arrs), ptrs, fmts)
        return [a.view(np.float16) if f == F_RGBA16F else a for a, (f, _) in zip(arrs, outs)]


# the variants the tests, smoke() and bench.py use: built by __graft_entry__.build() while the checkout is there, so that they are
# available where it is not (the GPU box)
STANDARD_VARIANTS = [
    ("ssgi", dict(steps=20, refine_steps=5, mode=0, importance_sampling=True, missed_rays=False, use_direct_light=True, use_envmap=True, perspective=True)),
    ("ssgi", dict(steps=12, refine_steps=3, mode=0, importance_sampling=True, missed_rays=False, use_direct_light=True, use_envmap=True, perspective=True)),
    ("ssgi", dict(steps=20, refine_steps=5, mode=1, importance_sampling=True, missed_rays=False, use_direct_light=True, use_envmap=True, perspective=True)),
    ("temporal_reproject", dict(texture_count=2, input_type=0, neighborhood_clamp=(False, True), reproject_specular=(False, True), log_transform=True,
                                confidence_power=0.75, perspective=True)),
    ("temporal_reproject", dict(texture_count=1, input_type=2, neighborhood_clamp=(True,), reproject_specular=(True,), log_transform=True,
                                confidence_power=0.75, perspective=True)),
    ("temporal_reproject", dict(texture_count=1, input_type=1, neighborhood_clamp=(True,), reproject_specular=(False,), log_transform=True,
                                confidence_power=4.0, perspective=True)),
    ("poisson_denoise", dict(input_type="diffuseSpecular", gbuffer=True)),
    ("poisson_denoise", dict(input_type="specular", gbuffer=True)),
    ("poisson_denoise", dict(input_type="diffuse", gbuffer=False)),
    ("gi_compose", dict(input_type=0, perspective=True)),
    ("gi_compose", dict(input_type=2, perspective=True)),
    ("hbao", dict(spp=8)),
    ("motion_blur", dict(samples=16)),
    ("ao_compose", {}),
    ("traa_compose", {}),
    ("ssgi_compose", dict(fog=False, fog_exp2=False, perspective=True)),
    ("ssgi_compose", dict(fog=True, fog_exp2=False, perspective=True)),
    ("ssgi_compose", dict(fog=True, fog_exp2=True, perspective=True)),
    ("sharpness", {}),
    ("lens_distortion", {}),
    ("gradual_background", dict(perspective=True)),
    ("sparkle", dict(perspective=False)),
    ("sparkle", dict(perspective=True)),
    ("taa", dict(srgb_output=True)),
    ("taa", dict(srgb_output=False)),
]


def chain_available(mode: int = 0, steps: int = 20, refine_steps: int = 5) -> bool:
    """can the default-option SSGI / SSR chain (K1, K2, K3, K4) run through the reference shaders here?"""
    if assemble.available():
        return True
    need = [v for v in STANDARD_VARIANTS if v[0] in ("ssgi", "temporal_reproject", "poisson_denoise", "gi_compose")]
    if mode == 0:
        need = [need[i] for i in (0, 3, 6, 9)] if (steps, refine_steps) == (20, 5) else None
    else:
        need = [need[i] for i in (2, 4, 7, 10)] if (steps, refine_steps) == (20, 5) else None
    return bool(need) and all(available(n, **kw) for n, kw in need)


def prebuild(verbose: bool = False) -> int:
    """compile every standard variant (needs the reference checkout); returns how many libraries exist afterwards"""
    n = 0
    for name, kw in STANDARD_VARIANTS:
        Shader.get(name, **kw)
        n += 1
        if verbose:
            print("reference shader", _key(name, kw))
    return n


# ------------------------------------------------------------------------------------------------ uniform wiring (the JS side)
def _cam(s: Shader, cam: abi.CameraS, names=("projectionMatrix", "projectionMatrixInverse", "cameraMatrixWorld", "viewMatrix")):
    src = dict(projectionMatrix=cam.projection, projectionMatrixInverse=cam.projection_inverse, cameraMatrixWorld=cam.camera_matrix_world,
               viewMatrix=cam.view_matrix)
    s.set(optional=names, **{n: list(src[n]) for n in names})


def _blue(s: Shader, blue_noise, index: int):
    """src/utils/BlueNoiseUtils.js:9-33: 128x128 RGBA8, nearest, repeat"""
    s.tex("blueNoiseTexture", blue_noise, F_RGBA8, repeat=True)
    s.set(blueNoiseSize=[blue_noise.shape[1], blue_noise.shape[0]], blueNoiseIndex=index)


def ssgi_trace(p: abi.SsgiParams, depth, gbuffer, velocity, direct_light, accumulated, env, blue_noise, out_size=None):
    """SSGIPass.render  src/ssgi/pass/SSGIPass.js:68-95; material uniforms SSGIMaterial.js:15-42; env SSGIEffect.js:309-366 +
    EquirectHdrInfoUniform.js:287-303,346-349.  `env` is a tests/orc.Env (host-built tables + mip chain) or None.
    out_size (W, H): renderTarget.setSize(width * resolutionScale, ...) and the `resolution` uniform (SSGIPass.js:52-57)."""
    H, W = depth.shape
    if out_size:
        W, H = out_size
    f = int(p.flags)
    use_env = bool(f & abi.SSGI_USE_ENVMAP) and env is not None
    s = Shader.get("ssgi", steps=int(p.steps), refine_steps=int(p.refine_steps), mode=int(p.mode),
                   importance_sampling=bool(f & abi.SSGI_IMPORTANCE_SAMPLING) and use_env, missed_rays=bool(f & abi.SSGI_MISSED_RAYS),
                   use_direct_light=bool(f & abi.SSGI_USE_DIRECT_LIGHT), use_envmap=use_env, perspective=bool(p.cam.perspective))
    _cam(s, p.cam)
    near, far = float(p.cam.near_plane), float(p.cam.far_plane)
    s.set(optional=("cameraNear", "cameraFar", "nearMinusFar", "nearMulFar", "farMinusNear", "backgroundColor", "maxEnvMapMipLevel", "envBlur"),
          cameraNear=near, cameraFar=far, nearMinusFar=near - far, farMinusNear=far - near, nearMulFar=near * far,
          rayDistance=float(p.ray_distance), thickness=float(p.thickness), envBlur=float(p.env_blur),
          maxEnvMapMipLevel=float(p.max_env_map_mip_level), backgroundColor=[0, 0, 0], resolution=[W, H])
    s.tex("depthTexture", depth, F_R32F)
    s.tex("gBufferTexture", gbuffer, F_RGBA32F)
    s.tex("velocityTexture", velocity, F_RGBA32F)
    s.tex("directLightTexture", direct_light, F_RGBA16F, linear=True)
    s.tex("accumulatedTexture", accumulated, F_RGBA32F)
    _blue(s, blue_noise, int(p.blue_noise_index))
    if use_env:
        s.tex("envMapInfo.map", env.levels[0], F_RGBA16F, linear=True, mips=env.levels)
        s.tex("envMapInfo.marginalWeights", env.marginal.reshape(1, -1), F_R32F)
        s.tex("envMapInfo.conditionalWeights", env.conditional.reshape(env.h, env.w), F_R32F)
        s.set(**{"envMapInfo.size": [env.w, env.h], "envMapInfo.totalSumWhole": float(env.whole), "envMapInfo.totalSumDecimal": float(env.decimal)})
    return s.run(W, H, [(F_RGBA32F, None)])[0]


def temporal_reproject(p: abi.TemporalParams, inp, velocity, hist0, hist1, out0_prev, out1_prev, out_half=False):
    """TemporalReprojectPass.render  src/temporal-reproject/TemporalReprojectPass.js:162-214; ctor :76-117; setSize :132-152"""
    H, W = velocity.shape[:2]
    tc = int(p.texture_count)
    rs = [bool(p.reproject_specular[i]) for i in range(tc)]
    # neighborhoodClamp: SSGI [false, true], SSR / TRAA true (SSGIEffect.js:69-77, TRAAEffect.js:21-31): it equals reprojectSpecular
    # for the SSGI form and is all-true for the one-texture forms
    nc = rs if tc == 2 else [True]
    s = Shader.get("temporal_reproject", texture_count=tc, input_type=int(p.input_type), neighborhood_clamp=tuple(nc), reproject_specular=tuple(rs),
                   log_transform=bool(p.log_transform), confidence_power=float(p.confidence_power), perspective=bool(p.cam.perspective))
    _cam(s, p.cam)
    s.set(optional=("cameraNear", "cameraFar", "prevCameraPos", "cameraPos"),
          prevViewMatrix=list(p.prev_view_matrix), prevCameraMatrixWorld=list(p.prev_camera_matrix_world), prevProjectionMatrix=list(p.prev_projection),
          prevProjectionMatrixInverse=list(p.prev_projection_inverse), cameraPos=list(p.camera_pos), prevCameraPos=list(p.prev_camera_pos),
          maxBlend=float(p.max_blend), neighborhoodClampIntensity=float(p.neighborhood_clamp_intensity), keepData=float(p.keep_data),
          fullAccumulate=int(p.full_accumulate), invTexSize=[1.0 / W, 1.0 / H], cameraNear=float(p.cam.near_plane), cameraFar=float(p.cam.far_plane))
    half_in = inp.dtype in (np.float16, np.uint16)
    s.tex("inputTexture", inp, F_RGBA16F if half_in else F_RGBA32F, linear=half_in)
    s.tex("velocityTexture", velocity, F_RGBA32F)
    lin = bool(p.history_linear)
    hfmt = F_RGBA32F if hist0.dtype == np.float32 else F_RGBA16F  # RGBA32F: the FramebufferTexture of denoiseMode "full_temporal" / "temporal"
    s.tex("accumulatedTexture0", hist0, hfmt, linear=lin)
    if tc == 2:
        s.tex("accumulatedTexture1", hist1, hfmt, linear=lin)
    fmt = F_RGBA16F if out_half else F_RGBA32F
    outs = s.run(W, H, [(fmt, out0_prev)] + ([(fmt, out1_prev)] if tc == 2 else []))
    return outs[0], (outs[1] if tc == 2 else out1_prev)  # a second target that is not bound keeps its contents


def poisson_denoise(p: abi.PoissonParams, depth, gbuffer_or_normal, in0, in1, blue_noise, out0_prev, out1_prev):
    """PoissonDenoisePass.render  src/denoise/pass/PoissonDenoisePass.js:135-149; uniforms :48-69; setGBufferPass :108-117"""
    H, W = depth.shape
    tc = int(p.texture_count)
    spec = (bool(p.is_texture_specular[0]), bool(p.is_texture_specular[1]))
    input_type = "diffuseSpecular" if tc == 2 else ("specular" if spec[0] else "diffuse")
    s = Shader.get("poisson_denoise", input_type=input_type, gbuffer=bool(p.gbuffer_texture))
    s.set(optional=("projectionMatrix", "projectionMatrixInverse", "cameraMatrixWorld"), radius=float(p.radius), phi=float(p.phi), lumaPhi=float(p.luma_phi),
          depthPhi=float(p.depth_phi), normalPhi=float(p.normal_phi), roughnessPhi=float(p.roughness_phi), specularPhi=float(p.specular_phi), resolution=[W, H])
    s.tex("depthTexture", depth, F_R32F)
    s.tex("gBufferTexture" if p.gbuffer_texture else "normalTexture", gbuffer_or_normal, F_RGBA32F)
    half = in0.dtype in (np.float16, np.uint16)
    fmt = F_RGBA16F if half else F_RGBA32F
    s.tex("inputTexture", in0, fmt, linear=bool(p.input_linear))
    if tc == 2:
        s.tex("inputTexture2", in1, fmt, linear=bool(p.input_linear))
    _blue(s, blue_noise, int(p.blue_noise_index))
    outs = s.run(W, H, [(F_RGBA16F, out0_prev)] + ([(F_RGBA16F, out1_prev)] if tc == 2 else []))
    return outs[0], (outs[1] if tc == 2 else out1_prev)


def gi_compose(p: abi.ComposeParams, depth, gbuffer, diffuse_gi, specular_gi, out_prev, scene=None):
    """DenoiserComposePass.render  src/denoise/pass/DenoiserComposePass.js:125-131; uniforms :88-99; textures :23-33"""
    H, W = depth.shape
    s = Shader.get("gi_compose", input_type=int(p.input_type), perspective=bool(p.cam.perspective))
    _cam(s, p.cam)
    s.set(cameraNear=float(p.cam.near_plane), cameraFar=float(p.cam.far_plane))
    s.tex("depthTexture", depth, F_R32F)
    s.tex("gBufferTexture", gbuffer, F_RGBA32F)
    gi_float = any(a is not None and a.dtype == np.float32 for a in (diffuse_gi, specular_gi))  # "full_temporal": the temporal pass's FloatType NEAREST targets
    s.tex("diffuseGiTexture", diffuse_gi, F_RGBA32F if gi_float else F_RGBA16F, linear=not gi_float)
    s.tex("specularGiTexture", specular_gi, F_RGBA32F if gi_float else F_RGBA16F, linear=not gi_float)
    s.tex("sceneTexture", scene, F_RGBA16F, linear=True)
    return s.run(W, H, [(F_RGBA32F, np.asarray(out_prev, np.float32))])[0]


def ssgi_compose(depth, gi, scene, params=None):
    """SSGIEffect.update  src/ssgi/SSGIEffect.js:386-404 (uniform wiring) + :34-64"""
    H, W = depth.shape
    pr = params
    s = Shader.get("ssgi_compose", fog=bool(pr and pr.use_fog), fog_exp2=bool(pr and pr.fog_exp2), perspective=bool(pr.perspective) if pr else True)
    s.set(optional=("fogColor", "fogNear", "fogFar", "fogDensity"), isDebug=int(bool(pr and pr.is_debug)),
          cameraNear=float(pr.camera_near) if pr else 0.0, cameraFar=float(pr.camera_far) if pr else 0.0,
          fogColor=list(pr.fog_color) if pr else [0, 0, 0], fogNear=float(pr.fog_near) if pr else 0.0, fogFar=float(pr.fog_far) if pr else 0.0,
          fogDensity=float(pr.fog_density) if pr else 0.0)
    s.tex("depthTexture", depth, F_R32F)
    s.tex("inputTexture", gi, F_RGBA32F)
    s.tex("sceneTexture", scene, F_RGBA16F, linear=True)
    return s.run(W, H, [(F_RGBA16F, None)])[0]


def hbao(p: abi.HbaoParams, depth, blue_noise, out_prev):
    """AOPass.render  src/ao/AOPass.js:85-110; uniforms :36-54"""
    H, W = depth.shape
    s = Shader.get("hbao", spp=int(p.spp))
    s.set(optional=("frame", "blueNoiseRepeat", "cameraNear", "cameraFar", "viewMatrix"), projectionViewMatrix=list(p.projection_view),
          projectionMatrixInverse=list(p.projection_inverse), cameraMatrixWorld=list(p.camera_matrix_world), aoDistance=float(p.ao_distance),
          distancePower=float(p.distance_power), bias=float(p.bias), thickness=float(p.thickness), resolution=[W, H], frame=0, blueNoiseRepeat=[W / 128, H / 128])
    s.tex("depthTexture", depth, F_R32F)
    _blue(s, blue_noise, int(p.blue_noise_index))
    return s.run(W, H, [(F_RGBA16F, out_prev)])[0]


def ao_compose(p: abi.AoComposeParams, depth, ao, inp):
    """AOEffect  src/ao/AOEffect.js:27-35,171"""
    H, W = depth.shape
    s = Shader.get("ao_compose")
    s.set(power=float(p.power), color=list(p.color))
    s.tex("depthTexture", depth, F_R32F)
    s.tex("inputTexture", ao, F_RGBA16F, linear=True)
    s.tex("inputBuffer", inp, F_RGBA16F, linear=True)
    return s.run(W, H, [(F_RGBA16F, None)])[0]


def motion_blur(p: abi.MotionBlurParams, velocity, inp, blue_noise):
    """MotionBlurEffect.update  src/motion-blur/MotionBlurEffect.js:87-102"""
    H, W = velocity.shape[:2]
    s = Shader.get("motion_blur", samples=int(p.samples))
    s.set(optional=("texSize",), intensity=float(p.intensity), jitter=float(p.jitter), deltaTime=float(p.delta_time), frame=int(p.frame),
          resolution=list(p.resolution))
    s.tex("velocityTexture", velocity, F_RGBA32F)
    s.tex("inputTexture", inp, F_RGBA16F, linear=True)
    s.tex("inputBuffer", inp, F_RGBA16F, linear=True)
    s.tex("blueNoiseTexture", blue_noise, F_RGBA8, repeat=True)
    s.set(blueNoiseSize=[blue_noise.shape[1], blue_noise.shape[0]], optional=("blueNoiseIndex",), blueNoiseIndex=0)
    return s.run(W, H, [(F_RGBA16F, None)])[0]


def traa_compose(acc):
    """TRAAEffect  src/traa/TRAAEffect.js:11-15,63"""
    H, W = acc.shape[:2]
    s = Shader.get("traa_compose")
    s.tex("accumulatedTexture", acc, F_RGBA16F, linear=True)
    return s.run(W, H, [(F_RGBA16F, None)])[0]


_FX = {abi.FX_SHARPNESS: "sharpness", abi.FX_LENS_DISTORTION: "lens_distortion", abi.FX_GRADUAL_BACKGROUND: "gradual_background", abi.FX_SPARKLE: "sparkle"}


def effects(p: abi.EffectsParams, inp, depth, velocity):
    """An EffectPass with p.effects in order: postprocessing merges their mainImage()s into one shader; here each effect's own shader runs in turn,
    reading the SAME input buffer through `inputTexture` and the previous effect's colour through `inputBuffer` (which is what the merged shader's
    colour chaining amounts to: every effect samples `inputBuffer` only at vUv).  Intermediate colours stay fp32 (registers in the merged shader)."""
    H, W = inp.shape[:2]
    cur = None  # the chained colour (fp32 plane); None = the pass input
    for i in range(int(p.n_effects)):
        e = int(p.effects[i])
        kw = {}
        if e == abi.FX_GRADUAL_BACKGROUND:
            kw = dict(perspective=bool(p.cam.perspective))
        if e == abi.FX_SPARKLE:
            kw = dict(perspective=bool(p.sparkle_perspective))
        s = Shader.get(_FX[e], **kw)
        s.set(optional=("resolution", "texelSize", "cameraNear", "cameraFar"), resolution=[W, H], texelSize=[1.0 / W, 1.0 / H],
              cameraNear=float(p.cam.near_plane), cameraFar=float(p.cam.far_plane))
        if cur is None:
            s.tex("inputBuffer", inp, F_RGBA16F, linear=True)
        else:
            s.tex("inputBuffer", cur, F_RGBA32F)
        s.tex("inputTexture", inp, F_RGBA16F, linear=True, optional=True)
        if e == abi.FX_SHARPNESS:
            s.set(sharpness=float(p.sharpness))
        elif e == abi.FX_LENS_DISTORTION:
            s.set(alphax=float(p.alphax), alphay=float(p.alphay), aberration=float(p.aberration))
        elif e == abi.FX_GRADUAL_BACKGROUND:
            _cam(s, p.cam, ("projectionMatrix", "projectionMatrixInverse", "cameraMatrixWorld"))
            s.set(backgroundColor=list(p.background_color), maxDistance=float(p.max_distance))
            s.tex("depthTexture", depth, F_R32F)
        elif e == abi.FX_SPARKLE:
            _cam(s, p.cam)
            s.set(optional=("backgroundColor",), spread=float(p.spread), intensity=float(p.intensity), backgroundColor=[0, 0, 0])
            s.tex("velocityTexture", velocity, F_RGBA32F)
        cur = s.run(W, H, [(F_RGBA32F, None)])[0]
    return cur.astype(np.float16) if cur is not None else np.asarray(inp)


def taa(p: abi.TaaParams, inp, history):
    """TAAPass.render  src/taa/TAAPass.js:68-94 (renders to the screen; the FramebufferTexture copy is RGBA8)"""
    H, W = inp.shape[:2]
    s = Shader.get("taa", srgb_output=bool(p.srgb_output))
    s.set(optional=("invTexSize",), cameraNotMovedFrames=float(p.camera_not_moved_frames))
    s.tex("inputTexture", inp, F_RGBA16F, linear=True)
    s.tex("acculumatedTexture", history, F_RGBA8)
    return s.run(W, H, [(F_RGBA8, None)])[0]
